// Host-buffer form of apply_local_hamiltonian (pytenet/chain_ops.py:237-279): the call a NumPy user of the
// reference makes, with every tensor in HOST memory.  The three steps of the device path are kept, but the
// operands are moved in slices so that the PCIe copies overlap the tensor-pipe work:
//
//   copy-in stream :  w | (a[:, :, k0], r[k0]) | (a[:, :, k1], r[k1]) | ... | l
//   compute stream :        t1  = a[..k0] r[k0]   t1 += a[..k1] r[k1]  ...   t2 = W t1   out[m0] = l[:, m0]^T t2  ...
//   copy-out stream:                                                                       out[m0] -> host  ...
//
// Step 1 is split along its contraction index (the right bond j): slice c needs only columns k_c of `a`
// (a strided 2-D copy) and rows k_c of `r` (contiguous), accumulated into t1 by the GEMM's `accumulate`
// flag.  Step 3 is split along the rows of `out` (the left bra bond i'), whose row blocks are contiguous in
// the result, so every block travels back while the next one is computed.  What stays exposed is the
// first input slice and the last output block.
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <algorithm>
#include <vector>

#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

constexpr int MAX_SLICES = 8;
constexpr int64_t CSR_MAX_NNZ = 4096;          // beyond this the dense GEMM W step is the better kernel

struct HostPipe {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t entry = nullptr, l_ready = nullptr, done = nullptr;
    cudaEvent_t slice[MAX_SLICES] = {};
    cudaEvent_t block[MAX_SLICES] = {};
    bool ok = false;
    std::mutex busy;            // one host-buffer call at a time per device (the call is synchronous anyway)
};

// ---- pageable host memory: staged through a page-locked ring ----------------------------------------------
// A drop-in NumPy caller passes ordinary (pageable) arrays.  cudaMemcpyAsync from / to pageable memory is staged by
// the driver through one internal bounce buffer on the calling thread (~10 GB/s, and it serialises with the
// enqueueing of the compute work).  Here the library owns a small ring of page-locked chunks per device; a few
// worker threads copy the caller's pages into a chunk while the DMA engine drains the previous one, so a pageable
// call approaches the page-locked one.  Page-locked inputs (detected with cudaPointerGetAttributes) bypass all this.
class CopyPool {
public:
    explicit CopyPool(int nthreads) : n_(nthreads) {
        for (int i = 0; i < n_; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // dst[r * dpitch .. +width) = src[r * spitch .. +width) for r < rows, split over the workers by rows / bytes
    void copy2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t rows) {
        if (width * rows < (size_t(1) << 20)) {
            for (size_t r = 0; r < rows; r++) memcpy(dst + r * dpitch, src + r * spitch, width);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst; src_ = src; dp_ = dpitch; sp_ = spitch; w_ = width; rows_ = rows;
            pending_ = n_;
            gen_++;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void loop(int id) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            char* dst = dst_; const char* src = src_;
            const size_t dp = dp_, sp = sp_, w = w_, rows = rows_;
            lk.unlock();
            if (rows == 1) {
                const size_t per = ((w + n_ - 1) / n_ + 63) & ~size_t(63);
                const size_t b = std::min(w, per * id), e = std::min(w, per * (id + 1));
                if (e > b) memcpy(dst + b, src + b, e - b);
            } else {
                const size_t r0 = rows * id / n_, r1 = rows * (id + 1) / n_;
                if (dp == w && sp == w) {
                    if (r1 > r0) memcpy(dst + r0 * w, src + r0 * w, (r1 - r0) * w);
                } else {
                    for (size_t r = r0; r < r1; r++) memcpy(dst + r * dp, src + r * sp, w);
                }
            }
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    unsigned long long gen_ = 0;
    bool stop_ = false;
    int pending_ = 0;
    char* dst_ = nullptr; const char* src_ = nullptr;
    size_t dp_ = 0, sp_ = 0, w_ = 0, rows_ = 0;
};

constexpr int STAGE_CHUNKS = 4;
constexpr size_t STAGE_BYTES = size_t(32) << 20;

struct StageRing {
    char* buf[STAGE_CHUNKS] = {};
    cudaEvent_t ev[STAGE_CHUNKS] = {};
    bool busy[STAGE_CHUNKS] = {};
    int next = 0;
    bool ok = false;
    int init() {
        if (ok) return PTB_OK;
        for (int i = 0; i < STAGE_CHUNKS; i++) {
            PTB_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&buf[i]), STAGE_BYTES, cudaHostAllocDefault));
            PTB_CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        ok = true;
        return PTB_OK;
    }
    // next chunk, free for the host to write (waits for the DMA / host copy that used it last)
    int acquire(char** p) {
        const int c = next;
        next = (next + 1) % STAGE_CHUNKS;
        if (busy[c]) PTB_CUDA_TRY(cudaEventSynchronize(ev[c]));
        busy[c] = false;
        *p = buf[c];
        return c;
    }
};

CopyPool* copy_pool() {
    static CopyPool pool([] {
        unsigned hc = std::thread::hardware_concurrency();
        int n = hc >= 16 ? 6 : (hc >= 8 ? 4 : 2);
        return n;
    }());
    return &pool;
}

inline bool is_pageable(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

std::mutex g_mu;
HostPipe g_pipe[64];
StageRing g_ring_in[64], g_ring_out[64];

// Host -> device copy of a (possibly strided) 2-D region on `st`; pageable sources go through the ring.
int h2d_2d(StageRing* ring, char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t rows,
           cudaStream_t st) {
    if (!ring) {
        if (rows == 1 || (dpitch == width && spitch == width))
            return cuda_status(cudaMemcpyAsync(dst, src, width * rows, cudaMemcpyHostToDevice, st));
        return cuda_status(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, st));
    }
    if (rows > 1 && dpitch == width && spitch == width) { width *= rows; rows = 1; }
    if (rows == 1) {
        for (size_t off = 0; off < width; off += STAGE_BYTES) {
            const size_t n = std::min(STAGE_BYTES, width - off);
            char* chunk;
            const int c = ring->acquire(&chunk);
            if (c < 0) return c;
            copy_pool()->copy2d(chunk, n, src + off, n, n, 1);
            PTB_CUDA_TRY(cudaMemcpyAsync(dst + off, chunk, n, cudaMemcpyHostToDevice, st));
            PTB_CUDA_TRY(cudaEventRecord(ring->ev[c], st));
            ring->busy[c] = true;
        }
        return PTB_OK;
    }
    const size_t rows_per = std::max<size_t>(1, STAGE_BYTES / width);
    for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
        const size_t nr = std::min(rows_per, rows - r0);
        char* chunk;
        const int c = ring->acquire(&chunk);
        if (c < 0) return c;
        copy_pool()->copy2d(chunk, width, src + r0 * spitch, spitch, width, nr);      // gather into a dense chunk
        PTB_CUDA_TRY(cudaMemcpy2DAsync(dst + r0 * dpitch, dpitch, chunk, width, width, nr, cudaMemcpyHostToDevice, st));
        PTB_CUDA_TRY(cudaEventRecord(ring->ev[c], st));
        ring->busy[c] = true;
    }
    return PTB_OK;
}

int pipe_for_current_device(HostPipe** pp) {
    int dev = 0;
    PTB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return PTB_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lock(g_mu);
    HostPipe& p = g_pipe[dev];
    if (!p.ok) {
        PTB_CUDA_TRY(cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking));
        PTB_CUDA_TRY(cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking));
        PTB_CUDA_TRY(cudaEventCreateWithFlags(&p.entry, cudaEventDisableTiming));
        PTB_CUDA_TRY(cudaEventCreateWithFlags(&p.l_ready, cudaEventDisableTiming));
        PTB_CUDA_TRY(cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
        for (int i = 0; i < MAX_SLICES; i++) {
            PTB_CUDA_TRY(cudaEventCreateWithFlags(&p.slice[i], cudaEventDisableTiming));
            PTB_CUDA_TRY(cudaEventCreateWithFlags(&p.block[i], cudaEventDisableTiming));
        }
        p.ok = true;
    }
    *pp = &p;
    return PTB_OK;
}

inline size_t up256(size_t x) { return (x + 255) & ~size_t(255); }

struct Layout {
    size_t a, r, l, t1, t2, out, part, w, rowptr, col, val, total;
    size_t part_bytes;
};

Layout layout(int dtype, int w_cplx, int64_t Dl, int64_t d, int64_t Dr, int64_t cl, int64_t cr, int64_t dout,
              int64_t Dlp, int64_t Drp) {
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    const size_t ws = w_cplx ? 16 : 8;
    Layout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up256(bytes); return o; };
    L.a = take((size_t)Dl * d * Dr * es);
    L.r = take((size_t)Dr * cr * Drp * es);
    L.l = take((size_t)Dl * cl * Dlp * es);
    L.t1 = take((size_t)Dl * d * cr * Drp * es);
    L.t2 = take((size_t)Dl * cl * dout * Drp * es);
    L.out = take((size_t)Dlp * dout * Drp * es);
    L.part_bytes = (size_t)Dlp * dout * Drp * es * 8;       // split-K partial tiles of the row blocks of step 3
    L.part = take(L.part_bytes);
    const size_t nw = (size_t)cl * dout * d * cr;
    L.w = take(nw * ws);
    L.rowptr = take(((size_t)cl * dout + 1) * 4);
    L.col = take((size_t)CSR_MAX_NNZ * 4);
    L.val = take((size_t)CSR_MAX_NNZ * ws);
    L.total = off;
    return L;
}

// Slice boundaries of an index range moved through the copy/compute pipeline.  What stays exposed is the
// first slice on the way in (`small_first`) or the last one on the way out, so that one is made small; every
// later slice is large enough that its GEMM hides the next copy, and few enough that the per-slice GEMM
// overhead (pipeline fill, partial waves: ~0.5 ms each at D = 2048) stays below the copy time saved.
// Small problems (slices under 256 indices or 16 MiB) are not pipelined.
int slice_bounds(int64_t extent, size_t bytes, bool small_first, int64_t* bounds) {
    int n = 1;
    while (n < 4 && extent / (2 * n) >= 256 && bytes / (2 * n) >= (size_t(16) << 20)) n *= 2;
    bounds[0] = 0;
    if (n == 1) { bounds[1] = extent; return 1; }
    if (n == 2) { bounds[1] = ((extent / 2) + 15) & ~int64_t(15); bounds[2] = extent; return 2; }
    const int64_t e8 = ((extent / 8) + 15) & ~int64_t(15);
    if (small_first) { bounds[1] = e8; bounds[2] = 3 * e8; }          // 1/8, 1/4, 5/8
    else { bounds[1] = 4 * e8; bounds[2] = extent - e8; }             // 1/2, 3/8, 1/8
    bounds[3] = extent;
    return 3;
}

}  // namespace

extern "C" {

size_t ptb_apply_local_hamiltonian_host_workspace_bytes(int dtype, int w_is_complex, int64_t Dl, int64_t d_in,
                                                        int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out,
                                                        int64_t Dlp, int64_t Drp) {
    return layout(dtype, w_is_complex, Dl, d_in, Dr, chi_l, chi_r, d_out, Dlp, Drp).total;
}

int ptb_apply_local_hamiltonian_host(int dtype, int w_is_complex, const void* a, const void* w, const void* l,
                                     const void* r, void* out, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l,
                                     int64_t chi_r, int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    if (!a || !w || !l || !r || !out) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    if (dtype == PTB_REAL64 && w_is_complex) return PTB_ERR_BAD_DTYPE;
    for (int64_t x : {Dl, d_in, Dr, chi_l, chi_r, d_out, Dlp, Drp})
        if (x <= 0 || x > 0x7fffffffLL) return PTB_ERR_BAD_ARG;
    const int64_t d = d_in, cl = chi_l, cr = chi_r, dout = d_out;
    const Layout L = layout(dtype, w_is_complex, Dl, d, Dr, cl, cr, dout, Dlp, Drp);
    if (!workspace || workspace_bytes < L.total) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PTB_ERR_ALIGNMENT;
    HostPipe* pp = nullptr;
    int rc = pipe_for_current_device(&pp);
    if (rc) return rc;
    HostPipe& P = *pp;
    std::lock_guard<std::mutex> one_call(P.busy);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* base = static_cast<char*>(workspace);
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    const size_t wes = w_is_complex ? 16 : 8;
    char *a_d = base + L.a, *r_d = base + L.r, *l_d = base + L.l, *t1 = base + L.t1, *t2 = base + L.t2,
         *out_d = base + L.out, *part = base + L.part, *w_d = base + L.w;
    const char* a_h = static_cast<const char*>(a);
    const char* r_h = static_cast<const char*>(r);
    char* out_h = static_cast<char*>(out);

    // W: dense copy, plus its CSR form (built here on the host) when it is sparse enough for the HBM-bound
    // sparse kernel (csrc/wapply.cu)
    const int64_t wrows = cl * dout, wcols = d * cr;
    std::vector<int32_t> rowptr((size_t)wrows + 1, 0), col;
    std::vector<double> val;
    bool use_csr = true;
    {
        const double* wv = static_cast<const double*>(w);
        const int we = w_is_complex ? 2 : 1;
        for (int64_t m = 0; m < wrows && use_csr; m++) {
            for (int64_t c = 0; c < wcols; c++) {
                const double* e = wv + (m * wcols + c) * we;
                if (e[0] != 0.0 || (we == 2 && e[1] != 0.0)) {
                    if ((int64_t)col.size() >= CSR_MAX_NNZ) { use_csr = false; break; }
                    col.push_back((int32_t)c);
                    val.push_back(e[0]);
                    if (we == 2) val.push_back(e[1]);
                }
            }
            rowptr[(size_t)m + 1] = (int32_t)col.size();
        }
        if (col.empty()) { col.push_back(0); val.assign((size_t)we, 0.0); }
    }

    // pageable operands / result: staged through the device's page-locked rings (see above)
    int devid = 0;
    PTB_CUDA_TRY(cudaGetDevice(&devid));
    StageRing* ring_a = is_pageable(a) ? &g_ring_in[devid] : nullptr;
    StageRing* ring_r = is_pageable(r) ? &g_ring_in[devid] : nullptr;
    StageRing* ring_l = is_pageable(l) ? &g_ring_in[devid] : nullptr;
    StageRing* ring_o = is_pageable(out) ? &g_ring_out[devid] : nullptr;
    if (ring_a || ring_r || ring_l) PTB_TRY(g_ring_in[devid].init());
    if (ring_o) PTB_TRY(g_ring_out[devid].init());

    // all work of this call is ordered after what the caller already enqueued on `stream`
    PTB_CUDA_TRY(cudaEventRecord(P.entry, st));
    PTB_CUDA_TRY(cudaStreamWaitEvent(P.in, P.entry, 0));
    PTB_CUDA_TRY(cudaStreamWaitEvent(P.out, P.entry, 0));

    if (use_csr) {
        PTB_CUDA_TRY(cudaMemcpyAsync(base + L.rowptr, rowptr.data(), rowptr.size() * 4, cudaMemcpyHostToDevice, P.in));
        PTB_CUDA_TRY(cudaMemcpyAsync(base + L.col, col.data(), col.size() * 4, cudaMemcpyHostToDevice, P.in));
        PTB_CUDA_TRY(cudaMemcpyAsync(base + L.val, val.data(), val.size() * 8, cudaMemcpyHostToDevice, P.in));
    } else {
        PTB_CUDA_TRY(cudaMemcpyAsync(w_d, w, (size_t)wrows * wcols * wes, cudaMemcpyHostToDevice, P.in));
    }

    // ---- step 1, sliced along the contraction index j                              chain_ops.py:273
    int64_t kb[MAX_SLICES + 1];
    const int ns = slice_bounds(Dr, ((size_t)Dl * d * Dr + (size_t)Dr * cr * Drp) * es, true, kb);
    for (int c = 0; c < ns; c++) {
        const int64_t k0 = kb[c], kc = kb[c + 1] - kb[c];
        // columns [k0, k0+kc) of a viewed as (Dl*d) x Dr, same pitch on both sides
        PTB_TRY(h2d_2d(ring_a, a_d + k0 * es, (size_t)Dr * es, a_h + k0 * es, (size_t)Dr * es, (size_t)kc * es,
                       (size_t)Dl * d, P.in));
        PTB_TRY(h2d_2d(ring_r, r_d + (size_t)k0 * cr * Drp * es, 0, r_h + (size_t)k0 * cr * Drp * es, 0,
                       (size_t)kc * cr * Drp * es, 1, P.in));
        PTB_CUDA_TRY(cudaEventRecord(P.slice[c], P.in));
        PTB_CUDA_TRY(cudaStreamWaitEvent(st, P.slice[c], 0));
        rc = ptb_gemm_splitk(dtype, 0, 0, 0, Dl * d, cr * Drp, kc, a_d + k0 * es, Dr, r_d + (size_t)k0 * cr * Drp * es,
                             cr * Drp, t1, cr * Drp, 1, 0, 0, 0, c > 0 ? 1 : 0, 0, part, L.part_bytes, st);
        if (rc) return rc;
    }
    PTB_TRY(h2d_2d(ring_l, l_d, 0, static_cast<const char*>(l), 0, (size_t)Dl * cl * Dlp * es, 1, P.in));
    PTB_CUDA_TRY(cudaEventRecord(P.l_ready, P.in));

    // ---- step 2                                                                      chain_ops.py:276
    if (use_csr) {
        rc = ptb_wapply_csr(dtype, w_is_complex, wrows, wcols, Drp, reinterpret_cast<const int32_t*>(base + L.rowptr),
                            reinterpret_cast<const int32_t*>(base + L.col), base + L.val, t1, t2, Dl, st);
    } else if (dtype == PTB_COMPLEX128 && !w_is_complex) {
        // real W times complex t1: real GEMM on (re, im)-interleaved columns, half the flops of a zgemm
        rc = ptb_gemm(PTB_REAL64, 0, 0, 0, wrows, 2 * Drp, wcols, w_d, wcols, t1, 2 * Drp, t2, 2 * Drp, Dl, 0,
                      2 * wcols * Drp, 2 * wrows * Drp, 0, st);
    } else {
        rc = ptb_gemm(dtype, 0, 0, 0, wrows, Drp, wcols, w_d, wcols, t1, Drp, t2, Drp, Dl, 0, wcols * Drp, wrows * Drp,
                      0, st);
    }
    if (rc) return rc;

    // ---- step 3 in row blocks of out, each copied back while the next is computed    chain_ops.py:278
    PTB_CUDA_TRY(cudaStreamWaitEvent(st, P.l_ready, 0));
    int64_t mb[MAX_SLICES + 1];
    const int nb = slice_bounds(Dlp, (size_t)Dlp * dout * Drp * es * 4, false, mb);
    for (int b = 0; b < nb; b++) {
        const int64_t m0 = mb[b], mc = mb[b + 1] - mb[b];
        rc = ptb_gemm_splitk(dtype, 1, 0, 0, mc, dout * Drp, Dl * cl, l_d + m0 * es, Dlp, t2, dout * Drp,
                             out_d + (size_t)m0 * dout * Drp * es, dout * Drp, 1, 0, 0, 0, 0, 0, part, L.part_bytes, st);
        if (rc) return rc;
        PTB_CUDA_TRY(cudaEventRecord(P.block[b], st));
        PTB_CUDA_TRY(cudaStreamWaitEvent(P.out, P.block[b], 0));
        if (!ring_o)
            PTB_CUDA_TRY(cudaMemcpyAsync(out_h + (size_t)m0 * dout * Drp * es, out_d + (size_t)m0 * dout * Drp * es,
                                         (size_t)mc * dout * Drp * es, cudaMemcpyDeviceToHost, P.out));
    }
    if (ring_o) {
        // pageable result: every row block travels device -> page-locked chunk on the copy-out stream (ordered
        // after its GEMM by the event waits above: the blocks were enqueued in order) and is copied to the
        // caller's pages by the worker threads while the next chunk is in flight
        struct Pend { int c; char* dst; size_t n; };
        std::vector<Pend> pend;
        auto drain = [&](size_t keep) -> int {
            while (pend.size() > keep) {
                const Pend q = pend.front();
                pend.erase(pend.begin());
                PTB_CUDA_TRY(cudaEventSynchronize(ring_o->ev[q.c]));
                copy_pool()->copy2d(q.dst, q.n, ring_o->buf[q.c], q.n, q.n, 1);
                ring_o->busy[q.c] = false;
            }
            return PTB_OK;
        };
        const size_t total = (size_t)Dlp * dout * Drp * es;
        for (size_t off = 0; off < total; off += STAGE_BYTES) {
            const size_t n = std::min(STAGE_BYTES, total - off);
            PTB_TRY(drain(STAGE_CHUNKS - 1));
            const int c = ring_o->next;
            ring_o->next = (ring_o->next + 1) % STAGE_CHUNKS;
            PTB_CUDA_TRY(cudaMemcpyAsync(ring_o->buf[c], out_d + off, n, cudaMemcpyDeviceToHost, P.out));
            PTB_CUDA_TRY(cudaEventRecord(ring_o->ev[c], P.out));
            pend.push_back({c, out_h + off, n});
        }
        PTB_TRY(drain(0));
    }
    // later work on the caller's stream may reuse the workspace: order it after the last copy-out
    PTB_CUDA_TRY(cudaEventRecord(P.done, P.out));
    PTB_CUDA_TRY(cudaStreamWaitEvent(st, P.done, 0));
    PTB_CUDA_TRY(cudaStreamSynchronize(P.out));      // `out` (host memory) is complete on return
    return PTB_OK;
}

}  // extern "C"
