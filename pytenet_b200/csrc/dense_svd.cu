// Large dense SVD for the two-site splits (pytenet/bond_ops.py:41-54 -> block_sparse_util.py:294): cuSOLVER's
// polar-decomposition SVD (cusolverDnXgesvdp: QDWH polar factor + Hermitian eigensolve -- GEMM / QR shaped work
// that runs at tensor-pipe speed) behind the C ABI.  north_star leaves QR / SVD to cuSOLVER ("not the optimisation
// target"); what matters on the path is that the 2048 x 2048 complex split of a two-site step no longer costs
// more than the 25 Lanczos matvecs before it (round 1: gesvd 0.45 s vs 0.25 s of Lanczos).
//
// libcusolver is resolved at run time (dlopen of the copy already loaded into the process by the host
// framework, else the system one), so the library carries no link-time dependency on it.
#include <cusolverDn.h>
#include <dlfcn.h>

#include <mutex>

#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

struct Api {
    decltype(&cusolverDnCreate) create = nullptr;
    decltype(&cusolverDnSetStream) set_stream = nullptr;
    decltype(&cusolverDnCreateParams) create_params = nullptr;
    decltype(&cusolverDnXgesvdp_bufferSize) buffer_size = nullptr;
    decltype(&cusolverDnXgesvdp) gesvdp = nullptr;
    bool ok = false;
};

Api g_api;
std::once_flag g_api_once;
std::mutex g_handle_mu;
cusolverDnHandle_t g_handle[PTB_MAX_DEVICES] = {};
cusolverDnParams_t g_params[PTB_MAX_DEVICES] = {};

void load_api() {
    void* h = dlopen("libcusolver.so.11", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libcusolver.so.11", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcusolver.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_api.create = reinterpret_cast<decltype(g_api.create)>(dlsym(h, "cusolverDnCreate"));
    g_api.set_stream = reinterpret_cast<decltype(g_api.set_stream)>(dlsym(h, "cusolverDnSetStream"));
    g_api.create_params = reinterpret_cast<decltype(g_api.create_params)>(dlsym(h, "cusolverDnCreateParams"));
    g_api.buffer_size = reinterpret_cast<decltype(g_api.buffer_size)>(dlsym(h, "cusolverDnXgesvdp_bufferSize"));
    g_api.gesvdp = reinterpret_cast<decltype(g_api.gesvdp)>(dlsym(h, "cusolverDnXgesvdp"));
    g_api.ok = g_api.create && g_api.set_stream && g_api.create_params && g_api.buffer_size && g_api.gesvdp;
}

// per-device handle (cuSOLVER handles are bound to the device current at creation)
int handle_for_current_device(cusolverDnHandle_t* h, cusolverDnParams_t* p) {
    std::call_once(g_api_once, load_api);
    if (!g_api.ok) return PTB_ERR_NOT_INITIALISED;
    const int dev = current_device();
    if (dev >= PTB_MAX_DEVICES) return PTB_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(g_handle_mu);
    if (!g_handle[dev]) {
        if (g_api.create(&g_handle[dev]) != CUSOLVER_STATUS_SUCCESS) return PTB_ERR_NOT_INITIALISED;
        if (g_api.create_params(&g_params[dev]) != CUSOLVER_STATUS_SUCCESS) return PTB_ERR_NOT_INITIALISED;
    }
    *h = g_handle[dev];
    *p = g_params[dev];
    return PTB_OK;
}

}  // namespace

extern "C" {

int ptb_svd_polar_workspace_bytes(int dtype, int64_t rows, int64_t cols, size_t* device_bytes, size_t* host_bytes) {
    if (!device_bytes || !host_bytes || rows <= 0 || cols <= 0) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    cusolverDnHandle_t h;
    cusolverDnParams_t p;
    PTB_TRY(handle_for_current_device(&h, &p));
    const cudaDataType t = dtype == PTB_COMPLEX128 ? CUDA_C_64F : CUDA_R_64F;
    const int64_t k = rows < cols ? rows : cols;
    (void)k;
    cusolverStatus_t st = g_api.buffer_size(h, p, CUSOLVER_EIG_MODE_VECTOR, 1, rows, cols, t, nullptr, rows, CUDA_R_64F,
                                            nullptr, t, nullptr, rows, t, nullptr, cols, t, device_bytes, host_bytes);
    return st == CUSOLVER_STATUS_SUCCESS ? PTB_OK : PTB_ERR_BAD_ARG;
}

int ptb_svd_polar(int dtype, int64_t rows, int64_t cols, void* a, int64_t lda, double* s, void* u, int64_t ldu, void* v,
                  int64_t ldv, void* device_ws, size_t device_bytes, void* host_ws, size_t host_bytes, int* info,
                  double* err_sigma, void* stream) {
    if (!a || !s || !u || !v || !info || !err_sigma || rows <= 0 || cols <= 0) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    if (lda < rows || ldu < rows || ldv < cols) return PTB_ERR_BAD_ARG;
    cusolverDnHandle_t h;
    cusolverDnParams_t p;
    PTB_TRY(handle_for_current_device(&h, &p));
    const cudaDataType t = dtype == PTB_COMPLEX128 ? CUDA_C_64F : CUDA_R_64F;
    std::lock_guard<std::mutex> lock(g_handle_mu);          // one stream binding at a time per process
    if (g_api.set_stream(h, static_cast<cudaStream_t>(stream)) != CUSOLVER_STATUS_SUCCESS) return PTB_ERR_BAD_ARG;
    cusolverStatus_t st = g_api.gesvdp(h, p, CUSOLVER_EIG_MODE_VECTOR, 1, rows, cols, t, a, lda, CUDA_R_64F, s, t, u, ldu,
                                       t, v, ldv, t, device_ws, device_bytes, host_ws, host_bytes, info, err_sigma);
    if (st == CUSOLVER_STATUS_INVALID_VALUE) return PTB_ERR_BAD_ARG;
    return st == CUSOLVER_STATUS_SUCCESS ? PTB_OK : PTB_ERR_WORKSPACE;
}

}  // extern "C"
