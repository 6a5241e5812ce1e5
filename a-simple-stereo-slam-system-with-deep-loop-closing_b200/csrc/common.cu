// common.cu — error text, version, device selection, tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void sb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void sb_clear_error() { g_err[0] = 0; }

extern "C" const char *sb_last_error(void) { return g_err; }
extern "C" const char *sb_version(void) { return "slamb200 0.1 (sm_100a, CUDA " SB_STR(CUDART_VERSION) ")"; }

int sb_use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        sb_set_error("no CUDA device available (%s); libslamb200 has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return SB_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        sb_set_error("device %d out of range (%d CUDA devices)", device, n);
        return SB_ERR_INVALID;
    }
    SB_CUDA(cudaSetDevice(device));
    return SB_OK;
}

typedef CUresult (*sb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);

static int make_tensor_map_u8(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                              const uint32_t *box, CUtensorMapSwizzle swizzle) {
    static sb_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) {
            sb_set_error("cuTensorMapEncodeTiled not available from this driver");
            return SB_ERR_CUDA;
        }
        fn = (sb_encode_tiled_fn)p;
    }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i + 1 < rank) gstr[i] = strides_bytes[i];
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        sb_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", (int)r, rank,
                     (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1);
        return SB_ERR_CUDA;
    }
    return SB_OK;
}

int sb_make_tensor_map_u8(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                          const uint32_t *box) {
    return make_tensor_map_u8(map, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

// 128-byte swizzle: the shared-memory image a tcgen05 K-major SWIZZLE_128B operand descriptor expects (box[0] = 128 bytes)
int sb_make_tensor_map_u8_sw128(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                                const uint32_t *box) {
    return make_tensor_map_u8(map, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Page-locked host memory for the asynchronous host-pointer entry points (sb_stereo_submit, sb_ba_submit): copies from / to
// pageable memory make cudaMemcpyAsync synchronous, which silently serialises "submit ... wait" pairs.
extern "C" int sb_host_alloc(void **ptr, size_t bytes) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(ptr && bytes > 0, "null pointer or zero size");
    *ptr = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        sb_set_error("no CUDA device available; libslamb200 has no CPU fallback");
        return SB_ERR_CUDA;
    }
    SB_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return SB_OK;
}
extern "C" int sb_host_free(void *ptr) {
    SB_NVTX_FN();
    if (ptr) cudaFreeHost(ptr);
    return SB_OK;
}

bool sb_pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("SLAMB200_NO_PDL");
        return !(e && e[0] == '1');
    }();
    return on;
}
