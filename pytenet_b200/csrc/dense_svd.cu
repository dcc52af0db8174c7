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
#include <thread>
#include <vector>

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

// Worker slots of the batched entry: own cuSOLVER handle, parameter object and stream per slot and device, so the
// (latency-bound: a fixed chain of small kernels and host synchronisations per call) factorisations of the
// independent sector blocks of one split run concurrently.
constexpr int MAX_WORKERS = 8;
struct Worker {
    cusolverDnHandle_t handle = nullptr;
    cusolverDnParams_t params = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    std::vector<unsigned char> host_ws;
};
Worker g_workers[PTB_MAX_DEVICES][MAX_WORKERS];
std::mutex g_batch_mu;

int worker_ready(int dev, int w) {
    Worker& wk = g_workers[dev][w];
    if (wk.handle) return PTB_OK;
    if (g_api.create(&wk.handle) != CUSOLVER_STATUS_SUCCESS) return PTB_ERR_NOT_INITIALISED;
    if (g_api.create_params(&wk.params) != CUSOLVER_STATUS_SUCCESS) return PTB_ERR_NOT_INITIALISED;
    if (cudaStreamCreateWithFlags(&wk.stream, cudaStreamNonBlocking) != cudaSuccess) return PTB_ERR_NOT_INITIALISED;
    if (cudaEventCreateWithFlags(&wk.done, cudaEventDisableTiming) != cudaSuccess) return PTB_ERR_NOT_INITIALISED;
    if (g_api.set_stream(wk.handle, wk.stream) != CUSOLVER_STATUS_SUCCESS) return PTB_ERR_NOT_INITIALISED;
    return PTB_OK;
}

}  // namespace

extern "C" {

int ptb_svd_polar_batch(int dtype, int njobs, ptb_svd_job* jobs, int max_workers, void* stream) {
    if (njobs < 0 || (njobs > 0 && !jobs)) return PTB_ERR_BAD_ARG;
    if (njobs == 0) return PTB_OK;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    std::call_once(g_api_once, load_api);
    if (!g_api.ok) return PTB_ERR_NOT_INITIALISED;
    const int dev = current_device();
    if (dev >= PTB_MAX_DEVICES) return PTB_ERR_BAD_ARG;
    for (int j = 0; j < njobs; j++) {
        const ptb_svd_job& jb = jobs[j];
        if (!jb.a || !jb.s || !jb.u || !jb.v || !jb.info || jb.rows <= 0 || jb.cols <= 0 || jb.lda < jb.rows ||
            jb.ldu < jb.rows || jb.ldv < jb.cols)
            return PTB_ERR_BAD_ARG;
    }
    int nw = max_workers <= 0 ? MAX_WORKERS : (max_workers < MAX_WORKERS ? max_workers : MAX_WORKERS);
    if (nw > njobs) nw = njobs;
    std::lock_guard<std::mutex> lock(g_batch_mu);           // the worker slots serve one batch at a time
    for (int w = 0; w < nw; w++) PTB_TRY(worker_ready(dev, w));
    cudaStream_t main = static_cast<cudaStream_t>(stream);
    cudaEvent_t start;
    if (cudaEventCreateWithFlags(&start, cudaEventDisableTiming) != cudaSuccess) return PTB_ERR_NOT_INITIALISED;
    cudaEventRecord(start, main);
    const cudaDataType t = dtype == PTB_COMPLEX128 ? CUDA_C_64F : CUDA_R_64F;
    auto run = [&](int w) {
        cudaSetDevice(dev);
        Worker& wk = g_workers[dev][w];
        cudaStreamWaitEvent(wk.stream, start, 0);
        for (int j = w; j < njobs; j += nw) {
            ptb_svd_job& jb = jobs[j];
            size_t db = 0, hb = 0;
            cusolverStatus_t st = g_api.buffer_size(wk.handle, wk.params, CUSOLVER_EIG_MODE_VECTOR, 1, jb.rows, jb.cols,
                                                    t, nullptr, jb.lda, CUDA_R_64F, nullptr, t, nullptr, jb.ldu, t,
                                                    nullptr, jb.ldv, t, &db, &hb);
            if (st != CUSOLVER_STATUS_SUCCESS || db > jb.device_bytes) {
                jb.status = PTB_ERR_WORKSPACE;
                continue;
            }
            if (wk.host_ws.size() < hb + 16) wk.host_ws.resize(hb + 16);
            st = g_api.gesvdp(wk.handle, wk.params, CUSOLVER_EIG_MODE_VECTOR, 1, jb.rows, jb.cols, t, jb.a, jb.lda,
                              CUDA_R_64F, jb.s, t, jb.u, jb.ldu, t, jb.v, jb.ldv, t, jb.device_ws, jb.device_bytes,
                              wk.host_ws.data(), hb, jb.info, &jb.err_sigma);
            jb.status = st == CUSOLVER_STATUS_SUCCESS ? PTB_OK
                        : (st == CUSOLVER_STATUS_INVALID_VALUE ? PTB_ERR_BAD_ARG : PTB_ERR_WORKSPACE);
        }
        cudaEventRecord(wk.done, wk.stream);
    };
    if (nw == 1) {
        run(0);
    } else {
        std::vector<std::thread> threads;
        for (int w = 0; w < nw; w++) threads.emplace_back(run, w);
        for (auto& th : threads) th.join();
    }
    for (int w = 0; w < nw; w++) cudaStreamWaitEvent(main, g_workers[dev][w].done, 0);
    cudaEventDestroy(start);
    int worst = PTB_OK;
    for (int j = 0; j < njobs; j++)
        if (jobs[j].status != PTB_OK) worst = jobs[j].status;
    return worst == PTB_OK ? cuda_status(cudaGetLastError()) : worst;
}

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
