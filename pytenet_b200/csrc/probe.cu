// Peak probes for the roofline denominators (diagnostics, not on the product path):
// register-resident DMMA.8x8x4 and DFMA loops on every SM.  bench.py / tools use them
// to measure the FP64 pipe peak of the B200 the run landed on.
#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) dmma_8x8x4(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" {

// Launch `blocks` CTAs of 256 threads; each warp issues iters*16 DMMA.8x8x4
// (512 flop each) or each thread iters*16 DFMA (2 flop each).  `out` needs
// blocks*256 doubles.  Returns the flop count through *flops.
int ptb_probe_fp64_pipe(int use_dmma, int blocks, int iters, double* out, double* flops, void* stream) {
    if (blocks <= 0 || iters <= 0 || !out || !flops) return PTB_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (use_dmma) {
        dmma_probe_kernel<<<blocks, 256, 0, st>>>(out, iters);
        *flops = (double)blocks * 8.0 * iters * 16.0 * 512.0;
    } else {
        dfma_probe_kernel<<<blocks, 256, 0, st>>>(out, iters);
        *flops = (double)blocks * 256.0 * iters * 16.0 * 2.0;
    }
    return cuda_status(cudaGetLastError());
}

}  // extern "C"
