// Microbenchmark (not part of the product): cost of a cluster-wide barrier vs __syncthreads on B200, 512 threads per CTA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(long long* out, int iters, double* sink) {
    __shared__ double s[512];
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        s[threadIdx.x] = v;
        if (MODE == 0) __syncthreads(); else cluster_sync();
        v = s[(threadIdx.x + 33) & 511] * 1.0000001 + 1.0;
        if (MODE == 0) __syncthreads(); else cluster_sync();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (v == 12345.678) *sink = v;
}
int main() {
    long long* d; double* sink; cudaMalloc(&d, 296 * 8); cudaMalloc(&sink, 8);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 200 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = mode == 0 ? 1 : (mode == 1 ? 2 : 4); at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        auto fn = mode == 0 ? k<0> : k<1>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (mode > 0) { int nc = 0; cudaOccupancyMaxActiveClusters(&nc, fn, &cfg); printf("max active clusters of %d (1 CTA/SM, 200 KB smem): %d\n", at[0].val.clusterDim.x, nc); }
        if (mode == 2) cfg.gridDim = dim3(144);
        cudaError_t e = cudaLaunchKernelEx(&cfg, fn, d, iters, sink);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, 144 * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 144; ++i) avg += h[i]; avg /= 144;
        printf("mode %d (%s): %s, %.1f cycles per barrier (incl. smem round trip)\n", mode, mode == 0 ? "__syncthreads" : (mode == 1 ? "cluster of 2" : "cluster of 4"), cudaGetErrorString(e), avg / (2.0 * iters));
    }
    return 0;
}
