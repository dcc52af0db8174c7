// C-ABI entry points of the chain contractions (see include/pytenet_b200.h).
//
// Each contraction is a chain of three (bond contraction: two) GEMMs on the
// FP64 tensor pipe, launched back to back on the caller's stream.  The
// intermediates t1 / t2 live in the caller's workspace and never leave the
// device; no operand is ever transposed or copied (the reference transposes and
// copies before every one of its tensordot calls, pytenet/chain_ops.py:276,278).
//
// Intermediate layouts (chosen so that every step is a plain row-major GEMM):
//   matvec / step_right:  t1[i, s, kappa, j']   t2[i, k, s', j']
//   step_left:            t [i, k, s',   j']    t2[i, s, kappa, j']
// The W step is a GEMM batched over the left bond index i.  When w is real and
// the state complex, t1[i] is reinterpreted as a real matrix with 2*Dr' columns,
// so the step runs as a real DGEMM at half the flops of the reference's upcast
// zgemm.
#include "../../include/pytenet_b200.h"
#include "gemm_dmma.cuh"
#include "gemm_ws.cuh"

using namespace ptb;

namespace {

inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

inline bool bad_dims(std::initializer_list<int64_t> v) {
    for (int64_t x : v)
        if (x <= 0 || x > 0x7fffffffLL) return true;
    return false;
}

inline bool fits_int(std::initializer_list<int64_t> v) {
    for (int64_t x : v)
        if (x > 0x7fffffffLL) return false;
    return true;
}

// ---- tiny products (launch-latency regime: merges, gauge absorption, environment updates at D <~ 32) ----
// The tensor-pipe engines are persistent kernels with ~200 KB of shared memory, tensor maps and an mbarrier
// pipeline: ~10 us per launch whatever the size.  Below TINY_GEMM_MACS multiply-adds one thread per output element
// with plain FP64 FMAs over L1/L2-resident operands finishes in 2-3 us.
constexpr long long TINY_GEMM_MACS = 1LL << 18;

template <bool CPLX>
__global__ void __launch_bounds__(128) gemm_tiny_kernel(const GemmParams p, int ta, int tb, int cj) {
    constexpr int E = CPLX ? 2 : 1;
    const long long per = (long long)p.M * p.N;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per * p.batch) return;
    const int bz = (int)(idx / per);
    const int rem = (int)(idx - (long long)bz * per);
    const int m = rem / p.N, n = rem - m * p.N;
    const double* A = p.A + (size_t)bz * p.sA * E;
    const double* B = p.B + (size_t)bz * p.sB * E;
    const size_t a0 = ta ? (size_t)m : (size_t)m * p.lda, as = ta ? (size_t)p.lda : 1;
    const size_t b0 = tb ? (size_t)n * p.ldb : (size_t)n, bs = tb ? 1 : (size_t)p.ldb;
    double re = 0.0, im = 0.0;
    for (int k = 0; k < p.K; k++) {
        if (CPLX) {
            const double ar = A[(a0 + k * as) * 2], ai = A[(a0 + k * as) * 2 + 1];
            const double br = B[(b0 + k * bs) * 2];
            double bi = B[(b0 + k * bs) * 2 + 1];
            if (cj) bi = -bi;
            re = fma(ar, br, re); re = fma(-ai, bi, re);
            im = fma(ar, bi, im); im = fma(ai, br, im);
        } else {
            re = fma(A[a0 + k * as], B[b0 + k * bs], re);
        }
    }
    double* c = p.C + ((size_t)bz * p.sC + (size_t)m * p.ldc + n) * E;
    if (p.accumulate) { re += c[0]; if (CPLX) im += c[1]; }
    c[0] = re;
    if (CPLX) c[1] = im;
}

// split_k: 1 = never split; 0 = choose automatically when `part_ws` is given; >1 = forced
template <bool CPLX>
int gemm(int ta, int tb, int cj, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
         int64_t ldb, void* C, int64_t ldc, int64_t batch, int64_t sA, int64_t sB, int64_t sC, int acc,
         cudaStream_t st, int split_k = 1, void* part_ws = nullptr, size_t part_ws_bytes = 0, int engine = 0) {
    if (!fits_int({M, N, K, batch})) return PTB_ERR_TOO_LARGE;
    GemmParams p;
    p.A = static_cast<const double*>(A);
    p.B = static_cast<const double*>(B);
    p.C = static_cast<double*>(C);
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.sA = sA; p.sB = sB; p.sC = sC;
    p.batch = (int)batch;
    p.accumulate = acc ? 1 : 0;
    p.tiles_m = p.tiles_n = 0;
    if (M == 0 || N == 0 || batch == 0) return PTB_OK;
    if (engine == 0 && split_k <= 1 && (long long)M * N * batch <= (1LL << 22) &&
        (long long)M * N * batch * (K > 0 ? K : 1) <= TINY_GEMM_MACS) {
        const long long total = (long long)M * N * batch;
        gemm_tiny_kernel<CPLX><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p, ta, tb, cj);
        return cuda_status(cudaGetLastError());
    }
    // engine selection (an argument, never process state): 0 = auto (warp-specialised TMA kernel when it
    // applies), 1 = first-generation cp.async kernel only, 2 = warp-specialised kernel required
    if (engine != 1 && K > 0) {
        const int rc = try_launch_ws<CPLX>(ta, tb, cj, p, st, split_k, part_ws, part_ws_bytes);
        if (rc != 1) return rc;
        if (engine == 2) return PTB_ERR_ALIGNMENT;
    }
    return launch_gemm<CPLX>(ta, tb, cj, p, st);
}

// t2[i] = op(W) * t[i] for i in [0, batch): rows_out x Drp per batch.
//   trans_w = 0: W is (rows_out x rows_in);  1: W stored (rows_in x rows_out)
template <bool CPLX>
int apply_w(bool w_cplx, int trans_w, int64_t rows_out, int64_t rows_in, int64_t ldw, int64_t Drp,
            const void* w, const void* t_in, void* t_out, int64_t batch, cudaStream_t st) {
    if (CPLX && !w_cplx) {
        // real W times complex t: real GEMM on (re, im)-interleaved columns
        return gemm<false>(trans_w, 0, 0, rows_out, 2 * Drp, rows_in, w, ldw, t_in, 2 * Drp, t_out, 2 * Drp, batch,
                           0, 2 * rows_in * Drp, 2 * rows_out * Drp, 0, st);
    }
    return gemm<CPLX>(trans_w, 0, 0, rows_out, Drp, rows_in, w, ldw, t_in, Drp, t_out, Drp, batch, 0,
                      rows_in * Drp, rows_out * Drp, 0, st);
}

// Workspace = intermediate 1 | intermediate 2 | split-K partial tiles of the last GEMM (up to
// MAX_SPLIT copies of the output; used when the output has too few tiles to fill 148 SMs).
constexpr int MAX_SPLIT = 8;

constexpr int FIRST_SPLIT = 4;                       // split factor budgeted for the first GEMM
constexpr int64_t FIRST_SPLIT_MAX_ELEMS = 6 * 148 * 128 * 64;   // workspace budget only: fewer than ~6 waves of tiles on a 148-SM part

size_t two_buffers(int dtype, int64_t n1, int64_t n2, int64_t nout = 0, int64_t nfirst = 0) {
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    size_t part = (size_t)nout * es * MAX_SPLIT;
    if (nfirst > 0 && nfirst < FIRST_SPLIT_MAX_ELEMS) {
        const size_t p1 = (size_t)nfirst * es * FIRST_SPLIT;
        if (p1 > part) part = p1;
    }
    return align16((size_t)n1 * es) + align16((size_t)n2 * es) + align16(part);
}

template <bool CPLX>
int apply_local_hamiltonian_impl(const void* a, const void* w, bool w_cplx, const void* l, const void* r,
                                 void* out, int64_t Dl, int64_t d, int64_t Dr, int64_t cl, int64_t cr,
                                 int64_t dout, int64_t Dlp, int64_t Drp, void* ws, size_t ws_bytes,
                                 cudaStream_t st) {
    if (!a || !w || !l || !r || !out) return PTB_ERR_BAD_ARG;
    if (bad_dims({Dl, d, Dr, cl, cr, dout, Dlp, Drp})) return PTB_ERR_BAD_ARG;
    const size_t es = CPLX ? 16 : 8;
    const size_t n1 = (size_t)Dl * d * cr * Drp, n2 = (size_t)Dl * cl * dout * Drp;
    if (ws_bytes < align16(n1 * es) + align16(n2 * es) || !ws) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(ws) % 16) return PTB_ERR_ALIGNMENT;
    // launch-latency regime: the whole contraction in one kernel (csrc/heff_small.cu)
    if (heff_small_applicable(CPLX, true, w_cplx, Dl, d, Dr, cl, cr, dout, Dlp, Drp))
        return heff_small_launch(CPLX, a, w, w_cplx, l, r, out, Dl, d, Dr, cl, cr, dout, Dlp, Drp, st);
    char* t1 = static_cast<char*>(ws);
    char* t2 = t1 + align16(n1 * es);
    char* part = t2 + align16(n2 * es);
    const size_t part_bytes = ws_bytes - (size_t)(part - t1);
    int rc;
    // (1) t1[(i,s),(kappa,j')] = a[(i,s),j] r[j,(kappa,j')]                 chain_ops.py:273
    rc = gemm<CPLX>(0, 0, 0, Dl * d, cr * Drp, Dr, a, Dr, r, cr * Drp, t1, cr * Drp, 1, 0, 0, 0, 0, st, 0, part,
                    part_bytes);
    if (rc) return rc;
    // (2) t2[i][(k,s'),j'] = w[(k,s'),(s,kappa)] t1[i][(s,kappa),j']        chain_ops.py:276
    rc = apply_w<CPLX>(w_cplx, 0, cl * dout, d * cr, d * cr, Drp, w, t1, t2, Dl, st);
    if (rc) return rc;
    // (3) out[i',(s',j')] = l[(i,k),i']^T t2[(i,k),(s',j')]                 chain_ops.py:278
    return gemm<CPLX>(1, 0, 0, Dlp, dout * Drp, Dl * cl, l, Dlp, t2, dout * Drp, out, dout * Drp, 1, 0, 0, 0, 0, st,
                      0, part, part_bytes);
}

// Same contraction with the W step done by the sparse CSR kernel (csrc/wapply.cu): for the 5-17 % dense MPO
// tensors of local Hamiltonians step 2 is then pure data movement (read t1, write t2 once).
template <bool CPLX>
int apply_local_hamiltonian_csr_impl(const void* a, const int32_t* w_rowptr, const int32_t* w_col, const void* w_val,
                                     bool w_cplx, const void* l, const void* r, void* out, int64_t Dl, int64_t d,
                                     int64_t Dr, int64_t cl, int64_t cr, int64_t dout, int64_t Dlp, int64_t Drp,
                                     void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!a || !w_rowptr || !w_col || !w_val || !l || !r || !out) return PTB_ERR_BAD_ARG;
    if (bad_dims({Dl, d, Dr, cl, cr, dout, Dlp, Drp})) return PTB_ERR_BAD_ARG;
    if (!CPLX && w_cplx) return PTB_ERR_BAD_DTYPE;
    const size_t es = CPLX ? 16 : 8;
    const size_t n1 = (size_t)Dl * d * cr * Drp, n2 = (size_t)Dl * cl * dout * Drp;
    if (ws_bytes < align16(n1 * es) + align16(n2 * es) || !ws) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(ws) % 16) return PTB_ERR_ALIGNMENT;
    char* t1 = static_cast<char*>(ws);
    char* t2 = t1 + align16(n1 * es);
    char* part = t2 + align16(n2 * es);
    const size_t part_bytes = ws_bytes - (size_t)(part - t1);
    int rc = gemm<CPLX>(0, 0, 0, Dl * d, cr * Drp, Dr, a, Dr, r, cr * Drp, t1, cr * Drp, 1, 0, 0, 0, 0, st, 0, part,
                        part_bytes);
    if (rc) return rc;
    rc = ptb_wapply_csr(CPLX ? PTB_COMPLEX128 : PTB_REAL64, w_cplx ? 1 : 0, cl * dout, d * cr, Drp, w_rowptr, w_col,
                        w_val, t1, t2, Dl, st);
    if (rc) return rc;
    return gemm<CPLX>(1, 0, 0, Dlp, dout * Drp, Dl * cl, l, Dlp, t2, dout * Drp, out, dout * Drp, 1, 0, 0, 0, 0, st,
                      0, part, part_bytes);
}

template <bool CPLX>
int bond_impl(const void* c, const void* l, const void* r, void* out, int64_t Dl, int64_t Dr, int64_t chi,
              int64_t Dlp, int64_t Drp, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!c || !l || !r || !out) return PTB_ERR_BAD_ARG;
    if (bad_dims({Dl, Dr, chi, Dlp, Drp})) return PTB_ERR_BAD_ARG;
    const size_t es = CPLX ? 16 : 8;
    const size_t n1 = (size_t)Dl * chi * Drp;
    if (ws_bytes < align16(n1 * es) || !ws) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(ws) % 16) return PTB_ERR_ALIGNMENT;
    if (heff_small_applicable(CPLX, false, false, Dl, 1, Dr, chi, chi, 1, Dlp, Drp))
        return heff_small_launch(CPLX, c, nullptr, false, l, r, out, Dl, 1, Dr, chi, chi, 1, Dlp, Drp, st);
    // (1) t[i,(k,j')] = c[i,j] r[j,(k,j')]                                  chain_ops.py:314
    int rc = gemm<CPLX>(0, 0, 0, Dl, chi * Drp, Dr, c, Dr, r, chi * Drp, ws, chi * Drp, 1, 0, 0, 0, 0, st);
    if (rc) return rc;
    // (2) out[i',j'] = l[(i,k),i']^T t[(i,k),j']                            chain_ops.py:316
    char* part = static_cast<char*>(ws) + align16(n1 * es);
    const size_t part_bytes = ws_bytes - align16(n1 * es);
    return gemm<CPLX>(1, 0, 0, Dlp, Drp, Dl * chi, l, Dlp, ws, Drp, out, Drp, 1, 0, 0, 0, 0, st, 0, part,
                      part_bytes);
}

template <bool CPLX>
int step_right_impl(const void* a, const void* b, const void* w, bool w_cplx, const void* r, void* r_next,
                    int64_t Dl, int64_t d, int64_t Dr, int64_t cl, int64_t cr, int64_t dout, int64_t Dlp,
                    int64_t Drp, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!a || !b || !w || !r || !r_next) return PTB_ERR_BAD_ARG;
    if (bad_dims({Dl, d, Dr, cl, cr, dout, Dlp, Drp})) return PTB_ERR_BAD_ARG;
    const size_t es = CPLX ? 16 : 8;
    const size_t n1 = (size_t)Dl * d * cr * Drp, n2 = (size_t)Dl * cl * dout * Drp;
    if (ws_bytes < align16(n1 * es) + align16(n2 * es) || !ws) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(ws) % 16) return PTB_ERR_ALIGNMENT;
    char* t1 = static_cast<char*>(ws);
    char* t2 = t1 + align16(n1 * es);
    int rc;
    // (1) t1 = a r                                                           chain_ops.py:50
    rc = gemm<CPLX>(0, 0, 0, Dl * d, cr * Drp, Dr, a, Dr, r, cr * Drp, t1, cr * Drp, 1, 0, 0, 0, 0, st);
    if (rc) return rc;
    // (2) t2[i] = w t1[i]  -> t2[i,k,s',j'] (the transpose at :54 is this layout)   chain_ops.py:52-54
    rc = apply_w<CPLX>(w_cplx, 0, cl * dout, d * cr, d * cr, Drp, w, t1, t2, Dl, st);
    if (rc) return rc;
    // (3) r_next[(i,k),i'] = t2[(i,k),(s',j')] conj(b[i',(s',j')])^T         chain_ops.py:56
    char* part = t2 + align16(n2 * es);
    return gemm<CPLX>(0, 1, 1, Dl * cl, Dlp, dout * Drp, t2, dout * Drp, b, dout * Drp, r_next, Dlp, 1, 0, 0, 0, 0,
                      st, 0, part, ws_bytes - (size_t)(part - t1));
}

template <bool CPLX>
int step_left_impl(const void* a, const void* b, const void* w, bool w_cplx, const void* l, void* l_next,
                   int64_t Dl, int64_t d, int64_t Dr, int64_t cl, int64_t cr, int64_t dout, int64_t Dlp,
                   int64_t Drp, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!a || !b || !w || !l || !l_next) return PTB_ERR_BAD_ARG;
    if (bad_dims({Dl, d, Dr, cl, cr, dout, Dlp, Drp})) return PTB_ERR_BAD_ARG;
    const size_t es = CPLX ? 16 : 8;
    const size_t n1 = (size_t)Dl * cl * dout * Drp, n2 = (size_t)Dl * d * cr * Drp;
    // same total as ptb_env_step_workspace_bytes (the two buffers swap roles)
    if (ws_bytes < align16(n1 * es) + align16(n2 * es) || !ws) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(ws) % 16) return PTB_ERR_ALIGNMENT;
    char* t = static_cast<char*>(ws);
    char* t2 = t + align16(n1 * es);
    int rc;
    // (1) t[(i,k),(s',j')] = l[(i,k),i'] conj(b[i',(s',j')])                 chain_ops.py:94
    rc = gemm<CPLX>(0, 0, 1, Dl * cl, dout * Drp, Dlp, l, Dlp, b, dout * Drp, t, dout * Drp, 1, 0, 0, 0, 0, st);
    if (rc) return rc;
    // (2) t2[i][(s,kappa),j'] = w[(k,s'),(s,kappa)]^T t[i][(k,s'),j']        chain_ops.py:96
    rc = apply_w<CPLX>(w_cplx, 1, d * cr, cl * dout, d * cr, Drp, w, t, t2, Dl, st);
    if (rc) return rc;
    // (3) l_next[j,(kappa,j')] = a[(i,s),j]^T t2[(i,s),(kappa,j')]           chain_ops.py:98
    char* part = t2 + align16(n2 * es);
    return gemm<CPLX>(1, 0, 0, Dr, cr * Drp, Dl * d, a, Dr, t2, cr * Drp, l_next, cr * Drp, 1, 0, 0, 0, 0, st, 0,
                      part, ws_bytes - (size_t)(part - t));
}

}  // namespace

extern "C" {

int ptb_version(void) { return 101; }


const char* ptb_status_string(int status) {
    switch (status) {
        case PTB_OK: return "ok";
        case PTB_ERR_BAD_ARG: return "bad argument (null pointer or non-positive / oversized extent)";
        case PTB_ERR_BAD_DTYPE: return "unsupported dtype";
        case PTB_ERR_WORKSPACE: return "workspace missing or too small";
        case PTB_ERR_ALIGNMENT: return "pointer not sufficiently aligned (8 B float64 / 16 B complex128)";
        case PTB_ERR_TOO_LARGE: return "extent exceeds the 31-bit index range of the kernels";
        case PTB_ERR_NOT_INITIALISED: return "communicator not initialised";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown status";
}

int ptb_gemm(int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k, const void* a,
             int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch, int64_t stride_a,
             int64_t stride_b, int64_t stride_c, int accumulate, void* stream) {
    if (!a || !b || !c) return PTB_ERR_BAD_ARG;
    if (m < 0 || n < 0 || k < 0 || batch < 0) return PTB_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == PTB_COMPLEX128)
        return gemm<true>(trans_a, trans_b, conj_b, m, n, k, a, lda, b, ldb, c, ldc, batch, stride_a, stride_b,
                          stride_c, accumulate, st);
    if (dtype == PTB_REAL64)
        return gemm<false>(trans_a, trans_b, 0, m, n, k, a, lda, b, ldb, c, ldc, batch, stride_a, stride_b,
                           stride_c, accumulate, st);
    return PTB_ERR_BAD_DTYPE;
}

int ptb_gemm_engine(int engine, int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k,
                    const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch,
                    int64_t stride_a, int64_t stride_b, int64_t stride_c, int accumulate, void* stream) {
    if (!a || !b || !c) return PTB_ERR_BAD_ARG;
    if (engine < 0 || engine > 2 || m < 0 || n < 0 || k < 0 || batch < 0) return PTB_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == PTB_COMPLEX128)
        return gemm<true>(trans_a, trans_b, conj_b, m, n, k, a, lda, b, ldb, c, ldc, batch, stride_a, stride_b,
                          stride_c, accumulate, st, 1, nullptr, 0, engine);
    if (dtype == PTB_REAL64)
        return gemm<false>(trans_a, trans_b, 0, m, n, k, a, lda, b, ldb, c, ldc, batch, stride_a, stride_b,
                           stride_c, accumulate, st, 1, nullptr, 0, engine);
    return PTB_ERR_BAD_DTYPE;
}

int ptb_gemm_splitk(int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k, const void* a,
                    int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch, int64_t stride_a,
                    int64_t stride_b, int64_t stride_c, int accumulate, int split_k, void* workspace,
                    size_t workspace_bytes, void* stream) {
    if (!a || !b || !c) return PTB_ERR_BAD_ARG;
    if (m < 0 || n < 0 || k < 0 || batch < 0 || split_k < 0 || split_k > 64) return PTB_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == PTB_COMPLEX128)
        return gemm<true>(trans_a, trans_b, conj_b, m, n, k, a, lda, b, ldb, c, ldc, batch, stride_a, stride_b,
                          stride_c, accumulate, st, split_k, workspace, workspace_bytes);
    if (dtype == PTB_REAL64)
        return gemm<false>(trans_a, trans_b, 0, m, n, k, a, lda, b, ldb, c, ldc, batch, stride_a, stride_b,
                           stride_c, accumulate, st, split_k, workspace, workspace_bytes);
    return PTB_ERR_BAD_DTYPE;
}

int ptb_gemm_banded(int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k, const void* a,
                    int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch, int64_t stride_a,
                    int64_t stride_b, int64_t stride_c, int accumulate, const int32_t* ktab, void* stream) {
    if (!a || !b || !c || !ktab) return PTB_ERR_BAD_ARG;
    if (m <= 0 || n <= 0 || k <= 0 || batch <= 0 || !fits_int({m, n, k, batch})) return PTB_ERR_BAD_ARG;
    if (accumulate < 0 || accumulate > 2) return PTB_ERR_BAD_ARG;
    GemmParams p;
    p.A = static_cast<const double*>(a);
    p.B = static_cast<const double*>(b);
    p.C = static_cast<double*>(c);
    p.M = (int)m; p.N = (int)n; p.K = (int)k;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.sA = stride_a; p.sB = stride_b; p.sC = stride_c;
    p.batch = (int)batch;
    p.accumulate = accumulate;
    p.tiles_m = p.tiles_n = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (dtype == PTB_COMPLEX128)
        rc = try_launch_ws<true>(trans_a, trans_b, conj_b, p, st, 1, nullptr, 0, ktab);
    else if (dtype == PTB_REAL64)
        rc = try_launch_ws<false>(trans_a, trans_b, 0, p, st, 1, nullptr, 0, ktab);
    else
        return PTB_ERR_BAD_DTYPE;
    return rc == 1 ? PTB_ERR_ALIGNMENT : rc;
}

int ptb_gemm_segmented(int dtype, int conj_b, int64_t m, int64_t n, int64_t k, const void* a, int64_t lda,
                       const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch, int64_t stride_a,
                       int64_t stride_b, int64_t stride_c, int accumulate, const int32_t* seg_ptr, const int32_t* segs,
                       const int64_t* sel_off, void* stream) {
    if (!a || !b || !c || !seg_ptr || !segs || !sel_off) return PTB_ERR_BAD_ARG;
    if (m <= 0 || n <= 0 || k <= 0 || batch <= 0 || !fits_int({m, n, k, batch})) return PTB_ERR_BAD_ARG;
    GemmParams p;
    p.A = static_cast<const double*>(a);
    p.B = static_cast<const double*>(b);
    p.C = static_cast<double*>(c);
    p.M = (int)m; p.N = (int)n; p.K = (int)k;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.sA = stride_a; p.sB = stride_b; p.sC = stride_c;
    p.batch = (int)batch;
    p.accumulate = accumulate ? 1 : 0;
    p.tiles_m = p.tiles_n = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static_assert(sizeof(long long) == sizeof(int64_t), "offset table is int64");
    const long long* off = reinterpret_cast<const long long*>(sel_off);
    if (dtype == PTB_COMPLEX128) return launch_ws_segmented<true>(conj_b, p, st, seg_ptr, segs, off);
    if (dtype == PTB_REAL64) return launch_ws_segmented<false>(0, p, st, seg_ptr, segs, off);
    return PTB_ERR_BAD_DTYPE;
}

int ptb_gemm_sector(int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k, const void* a,
                    int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch, int64_t stride_a,
                    int64_t stride_b, int64_t stride_c, int accumulate, const ptb_sector_tables* t, void* stream) {
    if (!a || !b || !c || !t) return PTB_ERR_BAD_ARG;
    if (m <= 0 || n <= 0 || k <= 0 || batch <= 0 || !fits_int({m, n, k, batch})) return PTB_ERR_BAD_ARG;
    const bool seg = t->seg_ptr != nullptr;
    if (seg ? (!t->segs || !t->sel_off || t->ktab || trans_a != 1 || trans_b != 0) : !t->ktab) return PTB_ERR_BAD_ARG;
    if (accumulate < 0 || accumulate > 2 || (seg && accumulate == 2)) return PTB_ERR_BAD_ARG;
    GemmParams p;
    p.A = static_cast<const double*>(a);
    p.B = static_cast<const double*>(b);
    p.C = static_cast<double*>(c);
    p.M = (int)m; p.N = (int)n; p.K = (int)k;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.sA = stride_a; p.sB = stride_b; p.sC = stride_c;
    p.batch = (int)batch;
    p.accumulate = accumulate;
    p.tiles_m = p.tiles_n = 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool cplx = dtype == PTB_COMPLEX128;
    if (!cplx && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    if (seg) {
        const long long* off = reinterpret_cast<const long long*>(t->sel_off);
        return cplx ? launch_ws_segmented<true>(conj_b, p, st, t->seg_ptr, t->segs, off, t->order)
                    : launch_ws_segmented<false>(0, p, st, t->seg_ptr, t->segs, off, t->order);
    }
    const int rc = cplx ? try_launch_ws<true>(trans_a, trans_b, conj_b, p, st, 1, nullptr, 0, t->ktab, t->order)
                        : try_launch_ws<false>(trans_a, trans_b, 0, p, st, 1, nullptr, 0, t->ktab, t->order);
    return rc == 1 ? PTB_ERR_ALIGNMENT : rc;
}

int ptb_gemm_tile_shape(int dtype, int* bm, int* bn, int* bk) {
    if (!bm || !bn || !bk) return PTB_ERR_BAD_ARG;
    if (dtype == PTB_COMPLEX128) { *bm = WsCfg<true>::BM; *bn = WsCfg<true>::BN; *bk = WsCfg<true>::BK; return PTB_OK; }
    if (dtype == PTB_REAL64) { *bm = WsCfg<false>::BM; *bn = WsCfg<false>::BN; *bk = WsCfg<false>::BK; return PTB_OK; }
    return PTB_ERR_BAD_DTYPE;
}

size_t ptb_apply_local_hamiltonian_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l,
                                                   int64_t chi_r, int64_t d_out, int64_t Dlp, int64_t Drp) {
    (void)Dr;
    return two_buffers(dtype, Dl * d_in * chi_r * Drp, Dl * chi_l * d_out * Drp, Dlp * d_out * Drp,
                       Dl * d_in * chi_r * Drp);
}

int ptb_apply_local_hamiltonian_z(const void* a, const void* w, int w_is_complex, const void* l, const void* r,
                                  void* out, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                                  int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    return apply_local_hamiltonian_impl<true>(a, w, w_is_complex != 0, l, r, out, Dl, d_in, Dr, chi_l, chi_r, d_out,
                                              Dlp, Drp, workspace, workspace_bytes,
                                              static_cast<cudaStream_t>(stream));
}

int ptb_apply_local_hamiltonian_d(const void* a, const void* w, const void* l, const void* r, void* out,
                                  int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                                  int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    return apply_local_hamiltonian_impl<false>(a, w, false, l, r, out, Dl, d_in, Dr, chi_l, chi_r, d_out, Dlp, Drp,
                                               workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ptb_apply_local_hamiltonian_csr_z(const void* a, const int32_t* w_rowptr, const int32_t* w_col, const void* w_val,
                                      int w_is_complex, const void* l, const void* r, void* out, int64_t Dl,
                                      int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out,
                                      int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                                      void* stream) {
    return apply_local_hamiltonian_csr_impl<true>(a, w_rowptr, w_col, w_val, w_is_complex != 0, l, r, out, Dl, d_in, Dr,
                                                  chi_l, chi_r, d_out, Dlp, Drp, workspace, workspace_bytes,
                                                  static_cast<cudaStream_t>(stream));
}

int ptb_apply_local_hamiltonian_csr_d(const void* a, const int32_t* w_rowptr, const int32_t* w_col, const void* w_val,
                                      const void* l, const void* r, void* out, int64_t Dl, int64_t d_in, int64_t Dr,
                                      int64_t chi_l, int64_t chi_r, int64_t d_out, int64_t Dlp, int64_t Drp,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    return apply_local_hamiltonian_csr_impl<false>(a, w_rowptr, w_col, w_val, false, l, r, out, Dl, d_in, Dr, chi_l,
                                                   chi_r, d_out, Dlp, Drp, workspace, workspace_bytes,
                                                   static_cast<cudaStream_t>(stream));
}

size_t ptb_apply_local_bond_contraction_workspace_bytes(int dtype, int64_t Dl, int64_t Dr, int64_t chi,
                                                        int64_t Dlp, int64_t Drp) {
    (void)Dr;
    return two_buffers(dtype, Dl * chi * Drp, 0, Dlp * Drp);
}

int ptb_apply_local_bond_contraction_z(const void* c, const void* l, const void* r, void* out, int64_t Dl,
                                       int64_t Dr, int64_t chi, int64_t Dlp, int64_t Drp, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    return bond_impl<true>(c, l, r, out, Dl, Dr, chi, Dlp, Drp, workspace, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

int ptb_apply_local_bond_contraction_d(const void* c, const void* l, const void* r, void* out, int64_t Dl,
                                       int64_t Dr, int64_t chi, int64_t Dlp, int64_t Drp, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    return bond_impl<false>(c, l, r, out, Dl, Dr, chi, Dlp, Drp, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
}

size_t ptb_env_step_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                                    int64_t d_out, int64_t Dlp, int64_t Drp) {
    // the final GEMM writes (Dr, chi_r, Drp) for step_left and (Dl, chi_l, Dlp) for step_right
    const int64_t o1 = Dr * chi_r * Drp, o2 = Dl * chi_l * Dlp;
    return two_buffers(dtype, Dl * d_in * chi_r * Drp, Dl * chi_l * d_out * Drp, o1 > o2 ? o1 : o2);
}

int ptb_env_step_left_z(const void* a, const void* b, const void* w, int w_is_complex, const void* l, void* l_next,
                        int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out,
                        int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes, void* stream) {
    return step_left_impl<true>(a, b, w, w_is_complex != 0, l, l_next, Dl, d_in, Dr, chi_l, chi_r, d_out, Dlp, Drp,
                                workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ptb_env_step_left_d(const void* a, const void* b, const void* w, const void* l, void* l_next, int64_t Dl,
                        int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out, int64_t Dlp,
                        int64_t Drp, void* workspace, size_t workspace_bytes, void* stream) {
    return step_left_impl<false>(a, b, w, false, l, l_next, Dl, d_in, Dr, chi_l, chi_r, d_out, Dlp, Drp, workspace,
                                 workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ptb_env_step_right_z(const void* a, const void* b, const void* w, int w_is_complex, const void* r,
                         void* r_next, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                         int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                         void* stream) {
    return step_right_impl<true>(a, b, w, w_is_complex != 0, r, r_next, Dl, d_in, Dr, chi_l, chi_r, d_out, Dlp, Drp,
                                 workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ptb_env_step_right_d(const void* a, const void* b, const void* w, const void* r, void* r_next, int64_t Dl,
                         int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out, int64_t Dlp,
                         int64_t Drp, void* workspace, size_t workspace_bytes, void* stream) {
    return step_right_impl<false>(a, b, w, false, r, r_next, Dl, d_in, Dr, chi_l, chi_r, d_out, Dlp, Drp, workspace,
                                  workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
