// One kernel per local problem for the launch-latency regime (BASELINE config 1: README XXZ D <= 28; config 5:
// METTS at bond dimension ~4; the edges of every chain).
//
// pytenet/tdvp.py:223-238 / dmrg.py:181-189 hand krylov.py:12-57 a closure; per local problem the reference runs
// numiter x (apply_local_hamiltonian + three-term orthogonalisation), the k x k eigenproblem and the combination
// of the Lanczos vectors (krylov.py:110-139).  On a GPU that is ~3 numiter + 3 launches of kernels that each finish
// in a few microseconds.  Here a SINGLE CTA runs the whole local step: start normalisation, all Lanczos iterations
// (the three contraction steps of chain_ops.py:273-278 as plain FP64 FMA loops over L2-resident operands -- the
// tiles are far too small for the tensor pipe), the tridiagonal problem (tridiag.cuh) and exp(-dt H_eff) v as the
// combination of the Lanczos vectors.  Intermediates and Lanczos vectors live in a global workspace that never
// leaves L1/L2; CTA barriers order the phases.  Chosen by the host side when one matvec is below
// PTB_SMALL_RUN_MAX_MACS multiply-adds (a single SM then needs a few microseconds per iteration).
//
// The zero-site problem (apply_local_bond_contraction, chain_ops.py:282-317) is the site problem with a
// one-dimensional physical index and no W step: w == nullptr.
#include "../../include/pytenet_b200.h"
#include "common.cuh"
#include "tridiag.cuh"

using namespace ptb;

namespace {

constexpr int RUN_THREADS = 512;
constexpr int W_SMEM = 1024;                     // doubles of the MPO tensor kept in shared memory
constexpr long long SMALL_RUN_MAX_MACS = 160000; // (complex) multiply-adds of one matvec

struct RunParams {
    const double* x;        // start vector (n elements)
    const double* w;        // (cl, d, d, cr) or nullptr (zero-site problem: d == 1, cl == cr)
    const double* l;        // (Dl, cl, Dl)
    const double* r;        // (Dr, cr, Dr)
    int w_cplx;
    int Dl, d, Dr, cl, cr;
    int numiter;
    double* V;              // numiter x n Lanczos vectors
    double* wv;             // n: H v
    double* t1;             // Dl d cr Dr
    double* t2;             // Dl cl d Dr
    double* scal;           // [|x|, alpha[0:k], beta[0:k-1]]
    int apply_expm;         // also out = sum_j coeff_j V_j  (expm_krylov), else the Lanczos run only
    double dt_re, dt_im;
    int out_cplx;
    double* out;
    double* coeff;          // 2 * TRIDIAG_MAX doubles + the int k_eff behind them
    double thresh;
};

// y = H_eff v for the whole CTA; E = doubles per element of the state (1: float64, 2: complex128)
template <bool CPLX>
__device__ __forceinline__ void matvec(const RunParams& p, const double* __restrict__ v, const double* ws, bool w_smem) {
    constexpr int E = CPLX ? 2 : 1;
    const int Dl = p.Dl, d = p.d, Dr = p.Dr, cl = p.cl, cr = p.cr;
    const int tid = threadIdx.x;
    // step 1: t1[(i,s), (K,j')] = sum_j v[(i,s), j] r[j, (K,j')]                       chain_ops.py:273
    const int n1 = Dl * d * cr * Dr, ncol = cr * Dr;
    for (int idx = tid; idx < n1; idx += RUN_THREADS) {
        const int row = idx / ncol, col = idx - row * ncol;
        const double* vr = v + (size_t)row * Dr * E;
        const double* rc = p.r + (size_t)col * E;
        double re = 0.0, im = 0.0;
        for (int j = 0; j < Dr; j++) {
            if (CPLX) {
                const double ar = vr[2 * j], ai = vr[2 * j + 1];
                const double br = rc[(size_t)j * ncol * 2], bi = rc[(size_t)j * ncol * 2 + 1];
                re = fma(ar, br, re); re = fma(-ai, bi, re);
                im = fma(ar, bi, im); im = fma(ai, br, im);
            } else {
                re = fma(vr[j], rc[(size_t)j * ncol], re);
            }
        }
        p.t1[(size_t)idx * E] = re;
        if (CPLX) p.t1[(size_t)idx * E + 1] = im;
    }
    __syncthreads();
    const double* t2 = p.t1;                       // zero-site problem: no W step, (i, K, j') is already (i, k, j')
    if (p.w != nullptr) {
        // step 2: t2[i, k, s', j'] = sum_{s,K} w[k, s', s, K] t1[i, s, K, j']          chain_ops.py:276
        const int n2 = Dl * cl * d * Dr;
        const int WE = p.w_cplx ? 2 : 1;
        const double* wsrc = w_smem ? ws : p.w;
        for (int idx = tid; idx < n2; idx += RUN_THREADS) {
            const int jp = idx % Dr;
            int rest = idx / Dr;
            const int sp = rest % d; rest /= d;
            const int k = rest % cl;
            const int i = rest / cl;
            const double* wrow = wsrc + (size_t)((k * d + sp) * d) * cr * WE;       // [s][K]
            const double* tin = p.t1 + ((size_t)i * d * cr * Dr + jp) * E;          // + (s * cr + K) * Dr * E
            double re = 0.0, im = 0.0;
            for (int sk = 0; sk < d * cr; sk++) {
                const double wr = wrow[sk * WE];
                const double wi = WE == 2 ? wrow[sk * 2 + 1] : 0.0;
                if (wr == 0.0 && wi == 0.0) continue;
                const double tr = tin[(size_t)sk * Dr * E];
                if (CPLX) {
                    const double ti = tin[(size_t)sk * Dr * E + 1];
                    re = fma(wr, tr, re); im = fma(wr, ti, im);
                    if (WE == 2) { re = fma(-wi, ti, re); im = fma(wi, tr, im); }
                } else {
                    re = fma(wr, tr, re);
                }
            }
            p.t2[(size_t)idx * E] = re;
            if (CPLX) p.t2[(size_t)idx * E + 1] = im;
        }
        __syncthreads();
        t2 = p.t2;
    }
    // step 3: y[i', s', j'] = sum_{i,k} l[i, k, i'] t2[i, k, s', j']                    chain_ops.py:278
    const int n3 = Dl * d * Dr, nk = Dl * cl, dDr = d * Dr;
    for (int idx = tid; idx < n3; idx += RUN_THREADS) {
        const int ip = idx / dDr, sj = idx - ip * dDr;
        const double* lc = p.l + (size_t)ip * E;                                    // + (i * cl + k) * Dl * E
        const double* tc = t2 + (size_t)sj * E;                                     // + (i * cl + k) * d * Dr * E
        double re = 0.0, im = 0.0;
        for (int ik = 0; ik < nk; ik++) {
            if (CPLX) {
                const double ar = lc[(size_t)ik * Dl * 2], ai = lc[(size_t)ik * Dl * 2 + 1];
                const double br = tc[(size_t)ik * dDr * 2], bi = tc[(size_t)ik * dDr * 2 + 1];
                re = fma(ar, br, re); re = fma(-ai, bi, re);
                im = fma(ar, bi, im); im = fma(ai, br, im);
            } else {
                re = fma(lc[(size_t)ik * Dl], tc[(size_t)ik * dDr], re);
            }
        }
        p.wv[(size_t)idx * E] = re;
        if (CPLX) p.wv[(size_t)idx * E + 1] = im;
    }
    __syncthreads();
}

template <bool CPLX>
__global__ void __launch_bounds__(RUN_THREADS) lanczos_small_kernel(const RunParams p) {
    constexpr int E = CPLX ? 2 : 1;
    __shared__ double red[32];
    __shared__ double bc;
    __shared__ double ws[W_SMEM];
    const int tid = threadIdx.x;
    const int n = p.Dl * p.d * p.Dr;
    const int nd = n * E;
    const int k = p.numiter;
    double* nrm = p.scal;
    double* alpha = p.scal + 1;
    double* beta = alpha + k;
    const int wn = p.w ? p.cl * p.d * p.d * p.cr * (p.w_cplx ? 2 : 1) : 0;
    const bool w_smem = wn > 0 && wn <= W_SMEM;
    if (w_smem)
        for (int i = tid; i < wn; i += RUN_THREADS) ws[i] = p.w[i];

    // v_0 = x / |x|                                                                      krylov.py:31-33
    double acc = 0.0;
    for (int i = tid; i < nd; i += RUN_THREADS) acc += p.x[i] * p.x[i];
    acc = block_sum(acc, red);
    if (tid == 0) { bc = sqrt(acc); *nrm = bc; }
    __syncthreads();
    {
        const double sc = bc;
        for (int i = tid; i < nd; i += RUN_THREADS) p.V[i] = p.x[i] / sc;
    }
    __syncthreads();

    for (int j = 0; j < k; j++) {
        double* vj = p.V + (size_t)j * nd;
        matvec<CPLX>(p, vj, ws, w_smem);
        double* w = p.wv;
        // alpha_j = Re <v_j, w>                                                          krylov.py:41
        acc = 0.0;
        for (int i = tid; i < nd; i += RUN_THREADS) acc += w[i] * vj[i];
        acc = block_sum(acc, red);
        if (tid == 0) { bc = acc; alpha[j] = acc; }
        __syncthreads();
        if (j == k - 1) break;                     // the closing matvec only contributes alpha (krylov.py:53-56)
        const double al = bc;
        const double* vjm1 = j > 0 ? vj - nd : nullptr;
        const double bp = j > 0 ? beta[j - 1] : 0.0;
        acc = 0.0;
        for (int i = tid; i < nd; i += RUN_THREADS) {
            double sub = al * vj[i];               // same association as krylov.py:42
            if (vjm1 != nullptr) sub = sub + bp * vjm1[i];
            const double rr = w[i] - sub;
            w[i] = rr;
            acc += rr * rr;
        }
        acc = block_sum(acc, red);                 // (starts with a barrier: `bc` has been read by all)
        if (tid == 0) { bc = sqrt(acc); beta[j] = bc; }
        __syncthreads();
        const double be = bc;
        for (int i = tid; i < nd; i += RUN_THREADS) vj[nd + i] = w[i] / be;      // v_{j+1}, own elements only
        __syncthreads();
    }
    if (!p.apply_expm) return;

    // coeff = U (|x| exp(dt w) U[0, :]); out = sum_{j < k_eff} coeff_j v_j               krylov.py:122-136
    __threadfence_block();
    __syncthreads();
    int* keff = reinterpret_cast<int*>(p.coeff + 2 * TRIDIAG_MAX);
    tridiag_expm_coeff(p.scal, k, p.thresh, p.dt_re, p.dt_im, p.coeff, keff);
    const int ke = *keff;
    for (int i = tid; i < n; i += RUN_THREADS) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < ke; j++) {
            const double cr = p.coeff[2 * j], ci = p.coeff[2 * j + 1];
            if (CPLX) {
                const double xr = p.V[((size_t)j * n + i) * 2], xi = p.V[((size_t)j * n + i) * 2 + 1];
                re += cr * xr - ci * xi;
                im += cr * xi + ci * xr;
            } else {
                const double x = p.V[(size_t)j * n + i];
                re += cr * x;
                if (p.out_cplx) im += ci * x;
            }
        }
        if (p.out_cplx) { p.out[2 * (size_t)i] = re; p.out[2 * (size_t)i + 1] = im; }
        else p.out[i] = re;
    }
}

inline size_t up16(size_t x) { return (x + 15) & ~size_t(15); }

}  // namespace

extern "C" {

int ptb_local_step_small_fits(int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter) {
    if (Dl <= 0 || d <= 0 || Dr <= 0 || chi_l <= 0 || chi_r <= 0 || numiter < 1 || numiter > TRIDIAG_MAX) return 0;
    const long long macs = (long long)Dl * d * Dr * chi_r * Dr + (long long)Dl * chi_l * d * Dr * d * chi_r +
                           (long long)Dl * d * Dr * Dl * chi_l;
    return macs <= SMALL_RUN_MAX_MACS ? 1 : 0;
}

size_t ptb_local_step_small_workspace_bytes(int dtype, int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r) {
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    return up16((size_t)Dl * d * Dr * es) + up16((size_t)Dl * d * chi_r * Dr * es) +
           up16((size_t)Dl * chi_l * d * Dr * es) + up16((2 * TRIDIAG_MAX + 2) * sizeof(double));
}

int ptb_local_step_small(int dtype, const void* x, const void* w, int w_is_complex, const void* l, const void* r,
                         int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter, void* V,
                         double* scal, int apply_expm, double dt_re, double dt_im, int out_is_complex, void* out,
                         void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !l || !r || !V || !scal || !workspace) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    const bool cplx = dtype == PTB_COMPLEX128;
    if (!cplx && w_is_complex) return PTB_ERR_BAD_DTYPE;
    if (!ptb_local_step_small_fits(Dl, d, Dr, chi_l, chi_r, numiter)) return PTB_ERR_TOO_LARGE;
    if (!w && (d != 1 || chi_l != chi_r)) return PTB_ERR_BAD_ARG;
    if (apply_expm && (!out || (!out_is_complex && (cplx || dt_im != 0.0)))) return PTB_ERR_BAD_ARG;
    if (workspace_bytes < ptb_local_step_small_workspace_bytes(dtype, Dl, d, Dr, chi_l, chi_r)) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16 || reinterpret_cast<uintptr_t>(V) % 16) return PTB_ERR_ALIGNMENT;
    const size_t es = cplx ? 16 : 8;
    const int64_t n = Dl * d * Dr;
    RunParams p;
    p.x = static_cast<const double*>(x); p.w = static_cast<const double*>(w);
    p.l = static_cast<const double*>(l); p.r = static_cast<const double*>(r);
    p.w_cplx = w_is_complex ? 1 : 0;
    p.Dl = (int)Dl; p.d = (int)d; p.Dr = (int)Dr; p.cl = (int)chi_l; p.cr = (int)chi_r;
    p.numiter = numiter;
    p.V = static_cast<double*>(V);
    char* wsb = static_cast<char*>(workspace);
    p.wv = reinterpret_cast<double*>(wsb); wsb += up16((size_t)n * es);
    p.t1 = reinterpret_cast<double*>(wsb); wsb += up16((size_t)Dl * d * chi_r * Dr * es);
    p.t2 = reinterpret_cast<double*>(wsb); wsb += up16((size_t)Dl * chi_l * d * Dr * es);
    p.coeff = reinterpret_cast<double*>(wsb);
    p.scal = scal;
    p.apply_expm = apply_expm ? 1 : 0;
    p.dt_re = dt_re; p.dt_im = dt_im;
    p.out_cplx = out_is_complex ? 1 : 0;
    p.out = static_cast<double*>(out);
    p.thresh = 100.0 * (double)n * 2.220446049250313e-16;      // krylov.py:44
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cplx) lanczos_small_kernel<true><<<1, RUN_THREADS, 0, st>>>(p);
    else lanczos_small_kernel<false><<<1, RUN_THREADS, 0, st>>>(p);
    return cuda_status(cudaGetLastError());
}

}  // extern "C"
