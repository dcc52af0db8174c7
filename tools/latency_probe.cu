// Dependent-issue latencies on B200 that bound the launch-latency regime kernels (one thread / one CTA):
// DFMA, DADD, DMUL chains, sqrt / division / rsqrt (double), shared-memory load-to-use, warp shuffle + add,
// __syncthreads at several CTA sizes, cluster barrier.  nvcc -arch=sm_100a -O3 tools/latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void chains(double* out, long long* t, double x0, int n) {
    __shared__ double sh[64];
    double x = x0, y = 1.0000001;
    long long c0, c1;
    if (threadIdx.x == 0) for (int i = 0; i < 64; i++) sh[i] = (double)((i + 1) % 64);
    __syncthreads();
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
    c1 = clock64(); if (threadIdx.x == 0) t[0] = (c1 - c0) / (4 * n);
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { x = x + y; x = x + y; x = x + y; x = x + y; }
    c1 = clock64(); if (threadIdx.x == 0) t[1] = (c1 - c0) / (4 * n);
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { x = sqrt(x + 2.0); }
    c1 = clock64(); if (threadIdx.x == 0) t[2] = (c1 - c0) / n;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { x = 1.0 / (x + 2.0); }
    c1 = clock64(); if (threadIdx.x == 0) t[3] = (c1 - c0) / n;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { x = rsqrt(x + 2.0); }
    c1 = clock64(); if (threadIdx.x == 0) t[4] = (c1 - c0) / n;
    int idx = (int)x0 & 63;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { idx = (int)sh[idx]; idx = (int)sh[idx]; idx = (int)sh[idx]; idx = (int)sh[idx]; }
    c1 = clock64(); if (threadIdx.x == 0) t[5] = (c1 - c0) / (4 * n);      // LDS + F2I
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { x += __shfl_xor_sync(0xffffffffu, x, 1); x += __shfl_xor_sync(0xffffffffu, x, 2); }
    c1 = clock64(); if (threadIdx.x == 0) t[6] = (c1 - c0) / (2 * n);      // shuffle (2 x 32 bit) + DADD
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { __syncthreads(); __syncthreads(); }
    c1 = clock64(); if (threadIdx.x == 0) t[7] = (c1 - c0) / (2 * n);
    float f = (float)x0;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { f = fmaf(f, 1.0001f, 1e-9f); f = fmaf(f, 1.0001f, 1e-9f); f = fmaf(f, 1.0001f, 1e-9f); f = fmaf(f, 1.0001f, 1e-9f); }
    c1 = clock64(); if (threadIdx.x == 0) t[8] = (c1 - c0) / (4 * n);
    int q = idx + 12345, dv = (int)x0 + 7;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { q = q / dv + 100000; q = q / dv + 100000; }
    c1 = clock64(); if (threadIdx.x == 0) t[9] = (c1 - c0) / (2 * n);      // int division + add
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) { sh[threadIdx.x & 63] = x; __syncthreads(); x += sh[(threadIdx.x + 1) & 63]; __syncthreads(); }
    c1 = clock64(); if (threadIdx.x == 0) t[10] = (c1 - c0) / n;          // STS, BAR, LDS, DADD, BAR
    out[threadIdx.x] = x + idx + f + q;
}

__global__ void cluster_bar(long long* t, int n) {
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    long long c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++) cl.sync();
    long long c1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) t[0] = (c1 - c0) / n;
}

int main() {
    double* out; long long* t;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&t, 16 * 8);
    const char* names[] = {"DFMA", "DADD", "sqrt", "div", "rsqrt", "LDS+F2I", "SHFL+DADD", "__syncthreads", "FFMA", "IDIV+IADD", "STS,BAR,LDS,DADD,BAR"};
    for (int threads : {32, 128, 256, 512, 1024}) {
        long long h[16] = {0};
        chains<<<1, threads>>>(out, t, 3.0, 200);
        chains<<<1, threads>>>(out, t, 3.0, 200);
        cudaDeviceSynchronize();
        cudaMemcpy(h, t, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %4d:", threads);
        for (int i = 0; i < 11; i++) printf(" %s %lld", names[i], h[i]);
        printf("\n");
    }
    for (int C : {2, 4, 8}) {
        for (int threads : {128, 512}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(C); cfg.blockDim = dim3(threads);
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            long long h = 0;
            cudaLaunchKernelEx(&cfg, cluster_bar, t, 200);
            cudaLaunchKernelEx(&cfg, cluster_bar, t, 200);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, t, 8, cudaMemcpyDeviceToHost);
            printf("cluster %d x %d threads: cluster.sync %lld cycles (%s)\n", C, threads, h, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
