// common.cuh — shared helpers of libslamb200 (status codes, error text, checks, TMA/mbarrier PTX).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/slamb200.h"

// NVTX ranges (tracing): every C-ABI entry point is one range named after the function, every
// kernel stage of the extractor one nested range.  NVTX 3 is header-only and a no-op unless a tool (Nsight
// Systems / Compute) is attached to the process.
struct SbNvtxRange {
    explicit SbNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~SbNvtxRange() { nvtxRangePop(); }
    SbNvtxRange(const SbNvtxRange &) = delete;
    SbNvtxRange &operator=(const SbNvtxRange &) = delete;
};
#define SB_NVTX_FN() SbNvtxRange sb_nvtx_range__(__func__)
#define SB_NVTX_RANGE(name) SbNvtxRange sb_nvtx_range2__(name)

#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)

void sb_set_error(const char *fmt, ...);
void sb_clear_error();

#define SB_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            sb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return SB_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define SB_REQUIRE(cond, msg)                                                     \
    do {                                                                          \
        if (!(cond)) {                                                            \
            sb_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return SB_ERR_INVALID;                                                \
        }                                                                         \
    } while (0)

#define SB_TRY(call)                 \
    do {                             \
        int rc__ = (call);           \
        if (rc__ != SB_OK) return rc__; \
    } while (0)

static inline int sb_div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t sb_align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Selects `device` for the calling thread; SB_ERR_CUDA when there is no usable CUDA device
// (there is deliberately no CPU fallback anywhere in this library).
int sb_use_device(int device);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// dims/strides innermost first; strides_bytes has rank-1 entries (dimension 0 is dense).
int sb_make_tensor_map_u8(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                          const uint64_t *strides_bytes, const uint32_t *box);

int sb_make_tensor_map_u8_sw128(CUtensorMap *map, const void *base, int rank, const uint64_t *dims,
                                const uint64_t *strides_bytes, const uint32_t *box);

// Programmatic dependent launch (SLAMB200_NO_PDL=1 turns it off for A/B runs)
bool sb_pdl_enabled();

#ifdef __CUDACC__
// ---- programmatic dependent launch ---------------------------------------------------------------
// The kernels of one extraction follow one another on a stream, each needing ALL of its predecessor's output.  Launched
// with the programmatic-stream-serialisation attribute, a kernel's CTAs are scheduled while the predecessor's last wave
// is still running and block in griddepcontrol.wait (first statement of the kernel) until that grid has completed and its
// writes are visible: the launch latency and the ramp of the block scheduler disappear behind the predecessor's tail.
// Every kernel lets ITS dependent start launching as soon as all of its own CTAs are resident.  Without the attribute
// both instructions are no-ops.
static __device__ __forceinline__ void sb_pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline cudaError_t sb_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = sb_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// ---- mbarrier + TMA (cp.async.bulk.tensor) -------------------------------------------------------
static __device__ __forceinline__ uint32_t sb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void sb_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void sb_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void sb_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(sb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled load: box at element coordinates (x, y, z) -> dense smem tile, completion on `bar`.
static __device__ __forceinline__ void sb_tma_load_3d(void *smem_dst, const CUtensorMap *map, int x, int y, int z,
                                                      uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            sb_smem_u32(smem_dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(sb_smem_u32(bar))
        : "memory");
}
static __device__ __forceinline__ void sb_tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif
