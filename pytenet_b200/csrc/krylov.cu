// Device-resident Lanczos vector operations (see include/pytenet_b200.h).
//
// Restates the BLAS-1 glue of pytenet/krylov.py:26-56 as HBM-bound kernels whose
// scalars (alpha, beta, norms) stay in device memory, so a whole Lanczos run is
// enqueued without a single host synchronisation; the host only reads the k
// alphas / betas once at the end to solve the k x k tridiagonal problem.
//
// Reductions are deterministic: fixed grid, per-block partial sums written to a
// scratch buffer, and the last block to finish (atomic ticket) adds the partials
// in index order.
#include "../../include/pytenet_b200.h"
#include "common.cuh"
#include "tridiag.cuh"

using namespace ptb;

namespace {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1184;  // 148 SMs x 8 resident CTAs
constexpr size_t SCRATCH_DOUBLES = RED_MAX_BLOCKS + 8;

struct Scratch {
    double* partial;         // RED_MAX_BLOCKS doubles
    unsigned int* ticket;    // 1 counter (kept zero between kernels)
};

inline Scratch scratch_of(void* p) {
    Scratch s;
    s.partial = static_cast<double*>(p);
    s.ticket = reinterpret_cast<unsigned int*>(s.partial + RED_MAX_BLOCKS);
    return s;
}

inline int red_blocks(int64_t n_doubles) {
    int64_t b = (n_doubles + (int64_t)RED_THREADS * 8 - 1) / ((int64_t)RED_THREADS * 8);
    if (b < 1) b = 1;
    if (b > RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
    return (int)b;
}

// Finish a grid-wide sum: every block contributes `v` (valid in thread 0); the last
// block sums the partials in order and calls `fin(total)` from thread 0.
template <typename Fin>
__device__ __forceinline__ void grid_sum_finish(double v, Scratch s, double* red, Fin fin) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        s.partial[blockIdx.x] = v;
        __threadfence();
        const unsigned int t = atomicAdd(s.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) acc += s.partial[i];
        // fixed association: strided partial sums, then the block tree
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) {
            *s.ticket = 0u;
            fin(acc);
        }
    }
}

// ---- sum of squares -> nrm (optionally) ------------------------------------------------
// x viewed as nd doubles (complex vectors are 2n doubles: |z|^2 = re^2 + im^2)
__global__ void __launch_bounds__(RED_THREADS) sumsq_kernel(const double* __restrict__ x, int64_t nd,
                                                            double* __restrict__ out_sqrt, Scratch s) {
    __shared__ double red[32];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0);
    for (; i + 1 < nd; i += stride) {
        double2 v;
        if (vec_ok) v = *reinterpret_cast<const double2*>(x + i);
        else { v.x = x[i]; v.y = x[i + 1]; }
        acc += v.x * v.x + v.y * v.y;
    }
    if (i < nd) acc += x[i] * x[i];
    acc = block_sum(acc, red);
    grid_sum_finish(acc, s, red, [=](double tot) { *out_sqrt = sqrt(tot); });
}

// ---- out = x / *scale ------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) scale_inv_kernel(const double* __restrict__ x, int64_t nd,
                                                                const double* __restrict__ scale,
                                                                double* __restrict__ out) {
    const double sc = *scale;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // division (not multiplication by the reciprocal) to match v = w / beta (krylov.py:29,51);
    // 16-byte accesses, two independent pairs in flight per thread
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    const int64_t npair = vec_ok ? nd / 2 : 0;
    const double2* __restrict__ x2 = reinterpret_cast<const double2*>(x);
    double2* __restrict__ o2 = reinterpret_cast<double2*>(out);
    int64_t i = tid;
    for (; i + stride < npair; i += 2 * stride) {
        const double2 a = x2[i], b = x2[i + stride];
        o2[i] = make_double2(a.x / sc, a.y / sc);
        o2[i + stride] = make_double2(b.x / sc, b.y / sc);
    }
    if (i < npair) {
        const double2 a = x2[i];
        o2[i] = make_double2(a.x / sc, a.y / sc);
    }
    for (int64_t j = 2 * npair + tid; j < nd; j += stride) out[j] = x[j] / sc;
}

// ---- alpha = Re <w, v>  = sum over doubles of w_i v_i ------------------------------------
__global__ void __launch_bounds__(RED_THREADS) dot_real_kernel(const double* __restrict__ w,
                                                               const double* __restrict__ v, int64_t nd,
                                                               double* __restrict__ out, Scratch s) {
    __shared__ double red[32];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) acc += w[i] * v[i];
    acc = block_sum(acc, red);
    grid_sum_finish(acc, s, red, [=](double tot) { *out = tot; });
}

// ---- w -= alpha v_j + beta_prev v_jm1 ;  beta = |w| ---------------------------------------
__global__ void __launch_bounds__(RED_THREADS) axpy_norm_kernel(double* __restrict__ w, const double* __restrict__ vj,
                                                                const double* __restrict__ vjm1, int64_t nd,
                                                                const double* __restrict__ alpha,
                                                                const double* __restrict__ beta_prev,
                                                                double* __restrict__ beta_out, Scratch s) {
    __shared__ double red[32];
    const double al = *alpha;
    const double bp = (vjm1 != nullptr && beta_prev != nullptr) ? *beta_prev : 0.0;
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
        // same association as the reference: w - (alpha*v_j + beta*v_jm1)   (krylov.py:42)
        double sub = al * vj[i];
        if (vjm1 != nullptr) sub = sub + bp * vjm1[i];
        const double r = w[i] - sub;
        w[i] = r;
        acc += r * r;
    }
    acc = block_sum(acc, red);
    grid_sum_finish(acc, s, red, [=](double tot) { *beta_out = sqrt(tot); });
}

// ---- out[n] = sum_j coeff[j] V[j, :] --------------------------------------------------------
// VC: V complex; CC: coeff complex.  out complex iff VC || CC.
template <bool VC, bool CC>
__global__ void __launch_bounds__(256) combine_kernel(const double* __restrict__ v, int64_t ldv, int64_t n, int k,
                                                      const double* __restrict__ coeff, double* __restrict__ out) {
    extern __shared__ double cs[];  // k coefficients (re, im)
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        cs[2 * j] = CC ? coeff[2 * j] : coeff[j];
        cs[2 * j + 1] = CC ? coeff[2 * j + 1] : 0.0;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < k; j++) {
            const double cr = cs[2 * j], ci = cs[2 * j + 1];
            if (VC) {
                const double2 x = *reinterpret_cast<const double2*>(v + 2 * ((int64_t)j * ldv + i));
                re += cr * x.x - ci * x.y;
                im += cr * x.y + ci * x.x;
            } else {
                const double x = v[(int64_t)j * ldv + i];
                re += cr * x;
                if (CC) im += ci * x;
            }
        }
        if (VC || CC) {
            *reinterpret_cast<double2*>(out + 2 * i) = make_double2(re, im);
        } else {
            out[i] = re;
        }
    }
}

// ---- single-CTA forms for short vectors (launch-latency-bound local problems, D <~ 64) --------------
// One launch instead of two / three; same arithmetic per element, reductions in one block (deterministic).
constexpr int SMALL_THREADS = 512;
constexpr int64_t SMALL_ND = 16384;      // doubles; 128 KiB stays in L1/L2 between the passes

__global__ void __launch_bounds__(SMALL_THREADS) start_small_kernel(const double* __restrict__ x, int nd,
                                                                    double* __restrict__ nrm_out,
                                                                    double* __restrict__ v0) {
    __shared__ double red[32];
    __shared__ double bc;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nd; i += SMALL_THREADS) acc += x[i] * x[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) { bc = sqrt(acc); *nrm_out = bc; }
    __syncthreads();
    const double sc = bc;
    for (int i = threadIdx.x; i < nd; i += SMALL_THREADS) v0[i] = x[i] / sc;
}

__global__ void __launch_bounds__(SMALL_THREADS) ortho_small_kernel(double* w, const double* __restrict__ vj,
                                                                    const double* __restrict__ vjm1, int nd,
                                                                    const double* __restrict__ beta_prev,
                                                                    double* __restrict__ alpha_out,
                                                                    double* __restrict__ beta_out, double* v_next) {
    __shared__ double red[32];
    __shared__ double bc;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nd; i += SMALL_THREADS) acc += w[i] * vj[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) { bc = acc; *alpha_out = acc; }
    __syncthreads();
    const double al = bc;
    const double bp = (vjm1 != nullptr && beta_prev != nullptr) ? *beta_prev : 0.0;
    acc = 0.0;
    for (int i = threadIdx.x; i < nd; i += SMALL_THREADS) {
        double sub = al * vj[i];                       // same association as krylov.py:42
        if (vjm1 != nullptr) sub = sub + bp * vjm1[i];
        const double r = w[i] - sub;
        w[i] = r;
        acc += r * r;
    }
    acc = block_sum(acc, red);                         // (starts with a barrier: `bc` has been read by all)
    if (threadIdx.x == 0) { bc = sqrt(acc); *beta_out = bc; }
    __syncthreads();
    const double be = bc;
    for (int i = threadIdx.x; i < nd; i += SMALL_THREADS) v_next[i] = w[i] / be;   // own elements only
}

// ---- k x k tridiagonal problem of expm_krylov on the device ------------------------------------------
// krylov.py:122-136: coeff = U (|vec| exp(dt w) * U[0, :]) with (w, U) the eigen-decomposition of the Lanczos
// tridiagonal matrix (krylov.py:142-150).  Solved here so that a TDVP local step needs no device->host round
// trip: implicit symmetric QL iteration (the classic tql2 recurrence).  Every thread runs the scalar
// recurrence redundantly (identical values, no communication) and applies the plane rotations to ITS row of the
// eigenvector matrix; thread r ends up holding U[r, :].  The breakdown rule of krylov.py:44-50 is applied to
// the betas first: k_eff = first j with beta[j] < thresh, plus one.

__global__ void __launch_bounds__(2 * TRIDIAG_MAX) expm_coeff_kernel(const double* __restrict__ scal, int numiter,
                                                                 double thresh, double dt_re, double dt_im,
                                                                 double* __restrict__ coeff, int* __restrict__ keff_out) {
    __shared__ __align__(16) double scratch[TAYLOR_SCRATCH_DOUBLES];
    tridiag_expm_solve(scal, numiter, thresh, dt_re, dt_im, coeff, keff_out, scratch);
}

// out[n] = sum_{j < *k_eff} coeff[j] V[j, :]  (complex coefficients; OC: out complex, else the real part is
// stored -- real vectors with a real time step, where the imaginary parts are exactly zero)
template <bool VC, bool OC>
__global__ void __launch_bounds__(256) combine_devk_kernel(const double* __restrict__ v, int64_t ldv, int64_t n,
                                                           const int* __restrict__ k_eff,
                                                           const double* __restrict__ coeff, double* __restrict__ out) {
    __shared__ double cs[2 * TRIDIAG_MAX];
    const int k = *k_eff;
    for (int j = threadIdx.x; j < 2 * k; j += blockDim.x) cs[j] = coeff[j];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double re = 0.0, im = 0.0;
        for (int j = 0; j < k; j++) {
            const double cr = cs[2 * j], ci = cs[2 * j + 1];
            if (VC) {
                const double2 x = *reinterpret_cast<const double2*>(v + 2 * ((int64_t)j * ldv + i));
                re += cr * x.x - ci * x.y;
                im += cr * x.y + ci * x.x;
            } else {
                const double x = v[(int64_t)j * ldv + i];
                re += cr * x;
                if (OC) im += ci * x;
            }
        }
        if (OC) *reinterpret_cast<double2*>(out + 2 * i) = make_double2(re, im);
        else out[i] = re;
    }
}

int start_impl(int64_t n, int e, const void* x, void* v0, double* nrm, void* scratch, cudaStream_t st) {
    if (n <= 0 || !x || !v0 || !nrm || !scratch) return PTB_ERR_BAD_ARG;
    const int64_t nd = n * e;
    if (nd <= SMALL_ND) {
        start_small_kernel<<<1, SMALL_THREADS, 0, st>>>(static_cast<const double*>(x), (int)nd, nrm,
                                                        static_cast<double*>(v0));
        return cuda_status(cudaGetLastError());
    }
    Scratch s = scratch_of(scratch);
    const int gb = red_blocks(nd);
    sumsq_kernel<<<gb, RED_THREADS, 0, st>>>(static_cast<const double*>(x), nd, nrm, s);
    scale_inv_kernel<<<gb, RED_THREADS, 0, st>>>(static_cast<const double*>(x), nd, nrm, static_cast<double*>(v0));
    return cuda_status(cudaGetLastError());
}

int ortho_impl(int64_t n, int e, void* w, const void* vj, const void* vjm1, const double* beta_prev,
               double* alpha_out, double* beta_out, void* v_next, void* scratch, cudaStream_t st) {
    if (n <= 0 || !w || !vj || !alpha_out || !beta_out || !v_next || !scratch) return PTB_ERR_BAD_ARG;
    if ((vjm1 == nullptr) != (beta_prev == nullptr)) return PTB_ERR_BAD_ARG;
    const int64_t nd = n * e;
    double* wd = static_cast<double*>(w);
    if (nd <= SMALL_ND) {
        ortho_small_kernel<<<1, SMALL_THREADS, 0, st>>>(wd, static_cast<const double*>(vj),
                                                        static_cast<const double*>(vjm1), (int)nd, beta_prev,
                                                        alpha_out, beta_out, static_cast<double*>(v_next));
        return cuda_status(cudaGetLastError());
    }
    Scratch s = scratch_of(scratch);
    const int gb = red_blocks(nd);
    dot_real_kernel<<<gb, RED_THREADS, 0, st>>>(wd, static_cast<const double*>(vj), nd, alpha_out, s);
    axpy_norm_kernel<<<gb, RED_THREADS, 0, st>>>(wd, static_cast<const double*>(vj),
                                                 static_cast<const double*>(vjm1), nd, alpha_out, beta_prev,
                                                 beta_out, s);
    scale_inv_kernel<<<gb, RED_THREADS, 0, st>>>(wd, nd, beta_out, static_cast<double*>(v_next));
    return cuda_status(cudaGetLastError());
}

int alpha_impl(int64_t n, int e, const void* w, const void* vj, double* alpha_out, void* scratch, cudaStream_t st) {
    if (n <= 0 || !w || !vj || !alpha_out || !scratch) return PTB_ERR_BAD_ARG;
    const int64_t nd = n * e;
    dot_real_kernel<<<red_blocks(nd), RED_THREADS, 0, st>>>(static_cast<const double*>(w),
                                                            static_cast<const double*>(vj), nd, alpha_out,
                                                            scratch_of(scratch));
    return cuda_status(cudaGetLastError());
}

}  // namespace

extern "C" {

size_t ptb_lanczos_scratch_bytes(void) { return SCRATCH_DOUBLES * sizeof(double); }

int ptb_lanczos_start_d(int64_t n, const void* x, void* v0, double* nrm, void* scratch, void* stream) {
    return start_impl(n, 1, x, v0, nrm, scratch, static_cast<cudaStream_t>(stream));
}
int ptb_lanczos_start_z(int64_t n, const void* x, void* v0, double* nrm, void* scratch, void* stream) {
    return start_impl(n, 2, x, v0, nrm, scratch, static_cast<cudaStream_t>(stream));
}
int ptb_lanczos_ortho_step_d(int64_t n, void* w, const void* v_j, const void* v_jm1, const double* beta_prev,
                             double* alpha_out, double* beta_out, void* v_next, void* scratch, void* stream) {
    return ortho_impl(n, 1, w, v_j, v_jm1, beta_prev, alpha_out, beta_out, v_next, scratch,
                      static_cast<cudaStream_t>(stream));
}
int ptb_lanczos_ortho_step_z(int64_t n, void* w, const void* v_j, const void* v_jm1, const double* beta_prev,
                             double* alpha_out, double* beta_out, void* v_next, void* scratch, void* stream) {
    return ortho_impl(n, 2, w, v_j, v_jm1, beta_prev, alpha_out, beta_out, v_next, scratch,
                      static_cast<cudaStream_t>(stream));
}
int ptb_lanczos_alpha_d(int64_t n, const void* w, const void* v_j, double* alpha_out, void* scratch, void* stream) {
    return alpha_impl(n, 1, w, v_j, alpha_out, scratch, static_cast<cudaStream_t>(stream));
}
int ptb_lanczos_alpha_z(int64_t n, const void* w, const void* v_j, double* alpha_out, void* scratch, void* stream) {
    return alpha_impl(n, 2, w, v_j, alpha_out, scratch, static_cast<cudaStream_t>(stream));
}

int ptb_krylov_combine(int v_dtype, int coeff_dtype, int64_t n, int64_t k, const void* v, int64_t ldv,
                       const void* coeff, void* out, void* stream) {
    if (n <= 0 || k <= 0 || k > 4096 || !v || !coeff || !out || ldv < n) return PTB_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t gb64 = (n + 255) / 256;
    const int gb = (int)(gb64 > 148 * 16 ? 148 * 16 : gb64);
    const size_t sh = (size_t)k * 2 * sizeof(double);
    const double* vd = static_cast<const double*>(v);
    const double* cd = static_cast<const double*>(coeff);
    double* od = static_cast<double*>(out);
    const bool vc = v_dtype == PTB_COMPLEX128, cc = coeff_dtype == PTB_COMPLEX128;
    if ((v_dtype != PTB_REAL64 && !vc) || (coeff_dtype != PTB_REAL64 && !cc)) return PTB_ERR_BAD_DTYPE;
    if (vc && cc) combine_kernel<true, true><<<gb, 256, sh, st>>>(vd, ldv, n, (int)k, cd, od);
    else if (vc && !cc) combine_kernel<true, false><<<gb, 256, sh, st>>>(vd, ldv, n, (int)k, cd, od);
    else if (!vc && cc) combine_kernel<false, true><<<gb, 256, sh, st>>>(vd, ldv, n, (int)k, cd, od);
    else combine_kernel<false, false><<<gb, 256, sh, st>>>(vd, ldv, n, (int)k, cd, od);
    return cuda_status(cudaGetLastError());
}

int ptb_krylov_expm_apply(int v_dtype, int64_t n, int numiter, const void* v, int64_t ldv, const double* scal,
                          double dt_re, double dt_im, int out_is_complex, void* coeff_ws, void* out,
                          void* stream) {
    if (n <= 0 || numiter < 1 || numiter > TRIDIAG_MAX || !v || !scal || !coeff_ws || !out || ldv < n)
        return PTB_ERR_BAD_ARG;
    const bool vc = v_dtype == PTB_COMPLEX128;
    if (v_dtype != PTB_REAL64 && !vc) return PTB_ERR_BAD_DTYPE;
    if (!out_is_complex && (vc || dt_im != 0.0)) return PTB_ERR_BAD_DTYPE;
    if (reinterpret_cast<uintptr_t>(coeff_ws) % 16 || reinterpret_cast<uintptr_t>(out) % 16) return PTB_ERR_ALIGNMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* coeff = static_cast<double*>(coeff_ws);
    int* keff = reinterpret_cast<int*>(coeff + 2 * TRIDIAG_MAX);
    const double thresh = 100.0 * (double)n * 2.220446049250313e-16;      // krylov.py:44
    expm_coeff_kernel<<<1, 2 * TRIDIAG_MAX, 0, st>>>(scal, numiter, thresh, dt_re, dt_im, coeff, keff);
    int64_t gb64 = (n + 255) / 256;
    const int gb = (int)(gb64 > 148 * 16 ? 148 * 16 : gb64);
    const double* vd = static_cast<const double*>(v);
    double* od = static_cast<double*>(out);
    if (vc) combine_devk_kernel<true, true><<<gb, 256, 0, st>>>(vd, ldv, n, keff, coeff, od);
    else if (out_is_complex) combine_devk_kernel<false, true><<<gb, 256, 0, st>>>(vd, ldv, n, keff, coeff, od);
    else combine_devk_kernel<false, false><<<gb, 256, 0, st>>>(vd, ldv, n, keff, coeff, od);
    return cuda_status(cudaGetLastError());
}

size_t ptb_krylov_expm_workspace_bytes(void) { return (2 * TRIDIAG_MAX + 2) * sizeof(double); }

}  // extern "C"
