// Shared device helpers for the pytenet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <initializer_list>
#include "../../include/pytenet_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pytenet_b200 kernels are written for sm_100a (B200) only"
#endif

namespace ptb {

// status codes of the C ABI: PTB_OK / PTB_ERR_* macros from include/pytenet_b200.h
// positive return values are cudaError_t codes
static inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? PTB_OK : (int)e; }

#define PTB_CUDA_TRY(expr)                              \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) return (int)_e;          \
    } while (0)

// ---- per-device launch configuration ------------------------------------------------
// Function attributes (opt-in dynamic shared memory) and the SM count belong to a device, not to the
// process: a host that drives several GPUs from one process (or several threads) must get them per
// device.  `DeviceFlags` is one flag per device ordinal, set after the attribute call succeeded; a
// race between two threads only repeats an idempotent cudaFuncSetAttribute.
constexpr int PTB_MAX_DEVICES = 64;

struct DeviceFlags {
    std::atomic<unsigned char> done[PTB_MAX_DEVICES];
};

static inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev;
}

// number of SMs of the current device (cached per device ordinal)
static inline int device_sm_count() {
    static std::atomic<int> sms[PTB_MAX_DEVICES];
    const int dev = current_device();
    const int slot = dev < PTB_MAX_DEVICES ? dev : PTB_MAX_DEVICES - 1;
    int n = dev < PTB_MAX_DEVICES ? sms[slot].load(std::memory_order_relaxed) : 0;
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        if (dev < PTB_MAX_DEVICES) sms[slot].store(n, std::memory_order_relaxed);
    }
    return n;
}

// opt a kernel into `bytes` of dynamic shared memory on the current device (once per device)
template <class Kernel>
static inline int ensure_dynamic_smem(DeviceFlags& flags, Kernel kern, int bytes) {
    const int dev = current_device();
    if (dev < PTB_MAX_DEVICES && flags.done[dev].load(std::memory_order_acquire)) return PTB_OK;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    if (dev < PTB_MAX_DEVICES) flags.done[dev].store(1, std::memory_order_release);
    return PTB_OK;
}

#define PTB_TRY(expr)                                   \
    do {                                                \
        int _s = (expr);                                \
        if (_s != PTB_OK) return _s;                    \
    } while (0)

// ---- async copy (LDGSTS) with zero fill -----------------------------------------
// cp-size 16 or 8 bytes; src_bytes <= cp-size, the remainder of the destination is
// zero-filled, src_bytes == 0 reads nothing.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4) * B(4x8), SASS DMMA.8x8x4 -------------
// fragment ownership (lane = 4*g + q):  A[g][q],  B[q][g],  C[g][2q], C[g][2q+1]
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- deterministic block reduction (sum of doubles) ----------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// all threads must call; result valid in thread 0. `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double s = 0.0;
    if (wid == 0) {
        s = lane < nw ? red[lane] : 0.0;
        s = warp_sum(s);
    }
    return s;
}

// fused small-D matvec (csrc/heff_small.cu)
bool heff_small_applicable(bool cplx, bool has_w, bool w_cplx, int64_t Dl, int64_t d, int64_t Dr, int64_t cl,
                           int64_t cr, int64_t dout, int64_t Dlp, int64_t Drp);
int heff_small_launch(bool cplx, const void* a, const void* w, bool w_cplx, const void* l, const void* r, void* out,
                      int64_t Dl, int64_t d, int64_t Dr, int64_t cl, int64_t cr, int64_t dout, int64_t Dlp, int64_t Drp,
                      cudaStream_t st);

}  // namespace ptb
