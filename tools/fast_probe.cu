// Probe: device sb_fast_maybe / sb_fast_score (orb_core.inl) from global and from shared memory vs brute force.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../a-simple-stereo-slam-system-with-deep-loop-closing_b200/csrc/orb_core.inl"
#define W 64
#define H 48
__global__ void k_global(const uint8_t *img, int *out) {
    for (int i = threadIdx.x; i < W * H; i += blockDim.x) {
        int x = i % W, y = i / W;
        int v = -1000;
        if (x >= 3 && x < W - 3 && y >= 3 && y < H - 3) {
            const uint8_t *p = img + y * W + x;
            v = sb_fast_maybe(p, W, 7) ? sb_fast_score(p, W) : -999;
        }
        out[i] = v;
    }
}
__global__ void k_shared(const uint8_t *img, int *out, int pitch_rt) {
    __shared__ uint8_t t[W * H + 16];
    for (int i = threadIdx.x; i < W * H; i += blockDim.x) t[i] = img[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int y = 3 + warp; y < H - 3; y += 4)
        for (int x = 3 + lane; x < W - 3; x += 32) {
            const uint8_t *p = t + y * pitch_rt + x;
            out[y * W + x] = sb_fast_maybe(p, pitch_rt, 7) ? sb_fast_score(p, pitch_rt) : -999;
        }
}
static const int RX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
int main() {
    std::vector<uint8_t> img(W * H);
    srand(1);
    for (auto &v : img) v = rand() % 3 == 0 ? rand() % 256 : 100 + rand() % 20;
    for (int y = 10; y < 30; y++) for (int x = 20; x < 40; x++) img[y * W + x] = 220;
    uint8_t *d; int *o; cudaMalloc(&d, W * H); cudaMalloc(&o, W * H * 4);
    cudaMemcpy(d, img.data(), W * H, cudaMemcpyHostToDevice);
    std::vector<int> g(W * H), s(W * H, -1000);
    k_global<<<1, 128>>>(d, o); cudaMemcpy(g.data(), o, W * H * 4, cudaMemcpyDeviceToHost);
    cudaMemset(o, 0x80, W * H * 4);
    k_shared<<<1, 128>>>(d, o, W); cudaMemcpy(s.data(), o, W * H * 4, cudaMemcpyDeviceToHost);
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    int badg = 0, bads = 0, corners = 0;
    for (int y = 3; y < H - 3; y++) for (int x = 3; x < W - 3; x++) {
        int v = img[y * W + x], best = -256;
        for (int st = 0; st < 16; st++) {
            int mn = 1000, mx = -1000;
            for (int k = 0; k < 9; k++) { int dd = v - img[(y + RY[(st + k) & 15]) * W + x + RX[(st + k) & 15]]; if (dd < mn) mn = dd; if (dd > mx) mx = dd; }
            if (mn > best) best = mn; if (-mx > best) best = -mx;
        }
        int want = best - 1;
        if (want >= 7) corners++;
        int gg = g[y * W + x], ss = s[y * W + x];
        if (want >= 7 ? gg != want : (gg != -999 && gg != want)) { if (badg < 5) printf("global (%d,%d) got %d want %d\n", x, y, gg, want); badg++; }
        if (want >= 7 ? ss != want : (ss != -999 && ss != want)) { if (bads < 5) printf("shared (%d,%d) got %d want %d\n", x, y, ss, want); bads++; }
    }
    printf("corners %d bad_global %d bad_shared %d\n", corners, badg, bads);
    return 0;
}
