// Host build of csrc/qt_core.inl (every cooperative loop runs serially): lets the CPU test-suite
// check the level-synchronous quadtree against the literal oracle (oracle/orb_oracle.c).
// Test infrastructure only; built by tests/test_quadtree_model.py.
#define QT_HOST
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "../a-simple-stereo-slam-system-with-deep-loop-closing_b200/csrc/qt_core.inl"

extern "C" int qt_model(const uint32_t *cand, int n, int width, int height, int N, int nCols, int wCell, int hCell,
                        uint32_t *out, int out_cap) {
    int nIni = (int)roundf((float)width / (float)height);
    if (nIni < 1) return 0;
    int ncap = (4 * nIni > N + 3 ? 4 * nIni : N + 3) + 1;
    int P = 1; while (P < ncap) P <<= 1;
    std::vector<uint16_t> cnode(n + 1), cnt0(ncap), cnt1(ncap);
    std::vector<uint8_t> cq(n + 1);
    std::vector<int16_t> box0(ncap * 4), box1(ncap * 4);
    std::vector<int> childcnt(ncap * 4), ord(ncap), cpre(ncap), spre(ncap), cbase(ncap), ubase(ncap), scratch(64);
    std::vector<uint32_t> keys(P);
    QtCtx c;
    c.cand = cand; c.n = n; c.cnode = cnode.data(); c.cq = cq.data();
    c.box[0] = box0.data(); c.box[1] = box1.data(); c.cnt[0] = cnt0.data(); c.cnt[1] = cnt1.data();
    c.ncap = ncap; c.childcnt = childcnt.data(); c.ord = ord.data(); c.cpre = cpre.data(); c.spre = spre.data();
    c.cbase = cbase.data(); c.ubase = ubase.data(); c.keys = keys.data(); c.scratch = scratch.data();
    c.width = width; c.height = height; c.N = N; c.nCols = nCols; c.wCell = wCell; c.hCell = hCell;
    return qt_distribute(c, out, out_cap);
}
