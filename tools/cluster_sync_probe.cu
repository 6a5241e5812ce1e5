// cluster_sync_probe.cu — cost of a thread-block-cluster barrier on this GPU, with and without __threadfence(), and of a
// global-memory hand-over between CTAs of the cluster (store -> barrier -> load on another SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cluster_sync_probe tools/cluster_sync_probe.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
template <int MODE> __global__ void k(long long *out, double *buf, int iters) {
    cg::cluster_group c = cg::this_cluster();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    double acc = 0;
    c.sync();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (MODE == 2) { buf[gtid] = acc + i; }
        if (MODE >= 1) __threadfence();
        c.sync();
        if (MODE == 2) acc += buf[(gtid + blockDim.x * 3 + 17) % gsz];   // written by another CTA
        if (MODE == 3) { buf[gtid] = acc + i; c.sync(); acc += buf[(gtid + blockDim.x * 3 + 17) % gsz]; }
    }
    long long t1 = clock64();
    if (gtid == 0) out[0] = t1 - t0;
    if (acc == 12345.678) out[1] = 1;
}
template <int MODE> void run(const char *name, int cl, int threads) {
    long long *out; double *buf;
    cudaMalloc(&out, 16); cudaMalloc(&buf, 16 * 1024 * 8);
    cudaMemset(buf, 0, 16 * 1024 * 8);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cl); cfg.blockDim = dim3(threads);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int iters = 2000;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, out, buf, iters);
    long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-46s cluster %2d x %3d threads: %8.0f cycles per iteration  (%s)\n", name, cl, threads, (double)h / iters, cudaGetErrorString(e ? e : cudaGetLastError()));
    cudaFree(out); cudaFree(buf);
}
int main() {
    for (int cl : {8, 16})
        for (int th : {256, 512}) {
            run<0>("cluster.sync()", cl, th);
            run<1>("__threadfence(); cluster.sync()", cl, th);
            run<2>("store; __threadfence(); cluster.sync(); load", cl, th);
            run<3>("store; cluster.sync(); load (no fence)", cl, th);
        }
    return 0;
}
