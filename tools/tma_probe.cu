// Standalone probe of the TMA path used by orb.cu (descriptor in __grid_constant__ params vs global memory).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../a-simple-stereo-slam-system-with-deep-loop-closing_b200/csrc/common.cuh"
void sb_set_error(const char *fmt, ...) {}
void sb_clear_error() {}
#include <stdarg.h>
typedef CUresult (*fn_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Maps { CUtensorMap m[4]; };
#define BW 144
#define BH 38
__global__ void probe_param(const __grid_constant__ Maps maps, int idx, int x, int y, int z, uint8_t *out) {
    __shared__ __align__(128) uint8_t tile[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        sb_mbar_init(&bar, 1);
        sb_mbar_expect_tx(&bar, BW * BH);
        sb_tma_load_3d(tile, &maps.m[idx], x, y, z, &bar);
    }
    __syncthreads();
    sb_mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
__global__ void probe_global(const CUtensorMap *maps, int idx, int x, int y, int z, uint8_t *out) {
    __shared__ __align__(128) uint8_t tile[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        sb_mbar_init(&bar, 1);
        sb_mbar_expect_tx(&bar, BW * BH);
        sb_tma_load_3d(tile, &maps[idx], x, y, z, &bar);
    }
    __syncthreads();
    sb_mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}
int main() {
    const int W = 1241, H = 376, P = 1280, B = 2;
    size_t slab = (size_t)P * H + 4096;
    uint8_t *d; cudaMalloc(&d, slab * B);
    std::vector<uint8_t> h(slab * B);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + i / P);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    fn_t fn = (fn_t)p;
    Maps maps;
    for (int i = 0; i < 4; i++) {
        cuuint64_t dims[3] = {W, H, B}; cuuint64_t str[2] = {P, slab}; cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
        CUresult r = fn(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d\n", i, (int)r);
    }
    uint8_t *out; cudaMalloc(&out, BW * BH);
    CUtensorMap *dm; cudaMalloc(&dm, sizeof(maps)); cudaMemcpy(dm, &maps, sizeof(maps), cudaMemcpyHostToDevice);
    std::vector<uint8_t> o(BW * BH);
    int tests[][3] = {{0, -3, 0}, {-16, -3, 1}, {1200, 350, 1}, {1232, 370, 0}, {13, 5, 1}};
    for (auto &t : tests) {
        for (int mode = 0; mode < 2; mode++) {
            cudaMemset(out, 0xee, BW * BH);
            if (mode == 0) probe_param<<<1, 128>>>(maps, 1, t[0], t[1], t[2], out);
            else probe_global<<<1, 128>>>(dm, 1, t[0], t[1], t[2], out);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(o.data(), out, BW * BH, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < BH; r++) for (int c = 0; c < BW; c++) {
                int gx = t[0] + c, gy = t[1] + r;
                uint8_t want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0 : h[t[2] * slab + (size_t)gy * P + gx];
                if (o[r * BW + c] != want) bad++;
            }
            printf("mode %s coord (%d,%d,%d): %s, mismatches %d\n", mode ? "global" : "param", t[0], t[1], t[2], cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
