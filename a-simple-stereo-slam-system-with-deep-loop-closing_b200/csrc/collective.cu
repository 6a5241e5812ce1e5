// collective.cu — the one exchange step of the multi-GPU path (BASELINE config 5): an all-gather of
// the keyframe poses each rank owns, before PoseGraphOptimization (reference src/loopclosing.cpp:537-646 reads every
// keyframe's pose; in the sharded replay rank r holds keyframes r, r + world, ...).
//
// NCCL is reached through the caller's process: the symbols are resolved at first use from the libnccl.so.2 that is
// already loaded (the one that created the caller's ncclComm_t), else from the global scope (statically linked NCCL),
// else by loading libnccl.so.2.  libslamb200.so therefore has no link-time NCCL dependency and single-GPU users need
// no NCCL at all; a missing library is reported as SB_ERR_CUDA with the loader's text.
//
// Everything is one ncclAllGather on the caller's communicator and stream: each rank contributes a fixed-size record
// [cap][7] doubles + its count, so the counts need no second collective.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "common.cuh"

namespace {
struct nccl_api {
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int *);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *);
    ncclResult_t (*CommCuDevice)(const ncclComm_t, int *);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GetVersion)(int *);
    bool ok;
};
nccl_api g_nccl = {};

void *find_sym(void *lib, const char *name) { return lib ? dlsym(lib, name) : dlsym(RTLD_DEFAULT, name); }

int load_nccl() {
    if (g_nccl.ok) return SB_OK;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the caller already uses, whatever its path
    if (!lib && !dlsym(RTLD_DEFAULT, "ncclAllGather")) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib && !dlsym(RTLD_DEFAULT, "ncclAllGather")) {
        sb_set_error("NCCL is not available in this process: %s", dlerror());
        return SB_ERR_CUDA;
    }
#define SB_NCCL_SYM(field, name)                                                     \
    *(void **)(&g_nccl.field) = find_sym(lib, name);                                 \
    if (!g_nccl.field) {                                                             \
        sb_set_error("NCCL symbol %s not found", name);                              \
        return SB_ERR_CUDA;                                                          \
    }
    SB_NCCL_SYM(AllGather, "ncclAllGather")
    SB_NCCL_SYM(CommCount, "ncclCommCount")
    SB_NCCL_SYM(CommUserRank, "ncclCommUserRank")
    SB_NCCL_SYM(CommCuDevice, "ncclCommCuDevice")
    SB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
    SB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    SB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    SB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    SB_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef SB_NCCL_SYM
    g_nccl.ok = true;
    return SB_OK;
}

#define SB_NCCL(call)                                                                               \
    do {                                                                                            \
        ncclResult_t r__ = (call);                                                                  \
        if (r__ != ncclSuccess) {                                                                   \
            sb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return SB_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

// record of one rank in the gathered buffer: cap * 7 pose doubles, then the count as a double
inline size_t record_doubles(int cap) { return (size_t)cap * 7 + 1; }

// d_all_records [world][cap*7+1] -> d_all [world][cap][7] (dense) and d_counts [world]
__global__ void k_unpack_records(const double *__restrict__ rec, int world, int cap, double *__restrict__ all,
                                 int32_t *__restrict__ counts) {
    const size_t rd = (size_t)cap * 7 + 1, per = (size_t)cap * 7;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per * world; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / per, k = i % per;
        all[i] = rec[r * rd + k];
    }
    if (blockIdx.x == 0)
        for (int r = threadIdx.x; r < world; r += blockDim.x) counts[r] = (int32_t)rec[r * rd + per];
}
__global__ void k_pack_record(const double *__restrict__ local, int n_local, int cap, double *__restrict__ rec) {
    const size_t per = (size_t)cap * 7;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x)
        rec[i] = i < (size_t)n_local * 7 ? local[i] : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) rec[per] = (double)n_local;
}
}  // namespace

extern "C" int sb_nccl_version(int *version) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(version, "null pointer");
    SB_TRY(load_nccl());
    SB_NCCL(g_nccl.GetVersion(version));
    return SB_OK;
}

// Bootstrap helpers for a host that has no communicator yet (the reference has none: it is single-process).
// Rank 0 calls sb_nccl_unique_id and ships the 128 bytes to the other ranks over its own channel; every rank then
// calls sb_nccl_comm_init on its device.
extern "C" int sb_nccl_unique_id(uint8_t id128[128]) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(id128, "null pointer");
    SB_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    SB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return SB_OK;
}
extern "C" int sb_nccl_comm_init(void **comm, int device, int world, int rank, const uint8_t id128[128]) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(comm && id128, "null pointer");
    SB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
    *comm = nullptr;
    SB_TRY(load_nccl());
    SB_TRY(sb_use_device(device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t c;
    SB_NCCL(g_nccl.CommInitRank(&c, world, id, rank));
    *comm = c;
    return SB_OK;
}
extern "C" int sb_nccl_comm_destroy(void *comm) {
    SB_NVTX_FN();
    sb_clear_error();
    if (!comm) return SB_OK;
    SB_TRY(load_nccl());
    SB_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
    return SB_OK;
}

// Device-pointer form: enqueues pack -> ncclAllGather -> unpack on `stream`, no synchronisation.
//   d_local   [cap][7] (first n_local rows valid)          d_all    [world][cap][7]
//   d_counts  [world] int32                                 d_scratch (1 + world) * (cap * 7 + 1) doubles
extern "C" int sb_allgather_kf_poses_dev(void *nccl_comm, void *stream, const double *d_local, int n_local, double *d_all,
                                         int32_t *d_counts, int cap, double *d_scratch) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(nccl_comm && d_local && d_all && d_counts && d_scratch, "null pointer");
    SB_REQUIRE(cap >= 1 && cap <= (1 << 24) && n_local >= 0 && n_local <= cap, "n_local out of range [0, cap]");
    SB_TRY(load_nccl());
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    int world = 0;
    SB_NCCL(g_nccl.CommCount(comm, &world));
    SB_REQUIRE(world >= 1 && world <= 1024, "communicator size out of range");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t rd = record_doubles(cap);
    double *rec_local = d_scratch, *rec_all = d_scratch + rd;
    k_pack_record<<<sb_div_up(cap * 7, 256), 256, 0, s>>>(d_local, n_local, cap, rec_local);
    SB_CUDA(cudaGetLastError());
    SB_NCCL(g_nccl.AllGather(rec_local, rec_all, rd, ncclDouble, comm, s));
    const size_t total = (size_t)cap * 7 * world;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    k_unpack_records<<<blocks, 256, 0, s>>>(rec_all, world, cap, d_all, d_counts);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// Host-pointer form: local [cap][7] -> all [world][cap][7], counts [world]; stages through device memory
// on the communicator's device, runs on `stream` (NULL = the default stream) and synchronises it before returning.
extern "C" int sb_allgather_kf_poses(void *nccl_comm, void *stream, const double *local, int n_local, double *all, int *counts,
                                     int cap) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(nccl_comm && local && all && counts, "null pointer");
    SB_REQUIRE(cap >= 1 && cap <= (1 << 24) && n_local >= 0 && n_local <= cap, "n_local out of range [0, cap]");
    SB_TRY(load_nccl());
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    int world = 0, dev = -1;
    SB_NCCL(g_nccl.CommCount(comm, &world));
    SB_NCCL(g_nccl.CommCuDevice(comm, &dev));
    SB_REQUIRE(world >= 1 && world <= 1024, "communicator size out of range");
    SB_TRY(sb_use_device(dev));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t rd = record_doubles(cap), per = (size_t)cap * 7;
    double *d = nullptr;  // [local: per][all: world * per][scratch: (1 + world) * rd][counts: world int32]
    const size_t n_d = per + per * world + (1 + (size_t)world) * rd;
    SB_CUDA(cudaMalloc((void **)&d, n_d * 8 + (size_t)world * 4));
    double *d_local = d, *d_all = d + per, *d_scratch = d_all + per * world;
    int32_t *d_counts = (int32_t *)(d_scratch + (1 + (size_t)world) * rd);
    int rc = SB_OK;
    cudaError_t e = cudaMemcpyAsync(d_local, local, (size_t)n_local * 56, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) rc = sb_allgather_kf_poses_dev(nccl_comm, stream, d_local, n_local, d_all, d_counts, cap, d_scratch);
    if (e == cudaSuccess && rc == SB_OK) e = cudaMemcpyAsync(all, d_all, per * world * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && rc == SB_OK) e = cudaMemcpyAsync(counts, d_counts, (size_t)world * 4, cudaMemcpyDeviceToHost, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(d);
    if (rc != SB_OK) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) {
        sb_set_error("sb_allgather_kf_poses: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return SB_ERR_CUDA;
    }
    return SB_OK;
}
