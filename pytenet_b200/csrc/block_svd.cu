// Batched per-sector SVD of a block-sparse matrix (pytenet/block_sparse_util.py:244-319), SURVEY section 8(f) rank 1.
//
// One launch factorises all sector blocks that fit in shared memory, one CTA per block: one-sided (Hestenes) Jacobi
// on the gathered block -- columns are orthogonalised pairwise by plane rotations (round-robin ordering, one warp
// per pair), the same rotations accumulate the right singular vectors; at convergence the column norms are the
// singular values (high relative accuracy), which are sorted in descending order as LAPACK returns them, and U,
// sigma, V^H are scattered straight into the block-sparse outputs at the sector's position on the new bond.
// A block with fewer rows than columns is factorised through its conjugate transpose.  Larger blocks are left to
// the caller (cuSOLVER).
#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

// threads per CTA: one warp per column pair of a round; blocks with more than 16 columns get the full 32 warps
// (a round of k/2 pairs is then at most two pair-steps deep up to k = 128; the kernel is latency bound)
constexpr int SVD_MAX_THREADS = 1024;
constexpr int SVD_MAX_SWEEPS = 60;
constexpr int SVD_MAX_K = 1024;

struct Cz {
    double re, im;
};

template <bool CPLX>
__device__ __forceinline__ Cz ldz(const double* s, int idx) {
    if (CPLX) return {s[2 * idx], s[2 * idx + 1]};
    return {s[idx], 0.0};
}
template <bool CPLX>
__device__ __forceinline__ void stz(double* s, int idx, Cz v) {
    if (CPLX) { s[2 * idx] = v.re; s[2 * idx + 1] = v.im; }
    else s[idx] = v.re;
}

// rotate columns p, q of a column-major matrix with `rows` rows:  x_p' = cs x_p - sn conj(ph) x_q,
//                                                              x_q' = sn ph x_p + cs x_q
template <bool CPLX>
__device__ __forceinline__ void rotate_columns(double* mat, int rows, int p, int q, double cs, double sn, Cz ph,
                                               int lane) {
    for (int i = lane; i < rows; i += 32) {
        const Cz xp = ldz<CPLX>(mat, p * rows + i), xq = ldz<CPLX>(mat, q * rows + i);
        // conj(ph) * xq and ph * xp
        const Cz a = {ph.re * xq.re + ph.im * xq.im, ph.re * xq.im - ph.im * xq.re};
        const Cz b = {ph.re * xp.re - ph.im * xp.im, ph.re * xp.im + ph.im * xp.re};
        stz<CPLX>(mat, p * rows + i, Cz{cs * xp.re - sn * a.re, cs * xp.im - sn * a.im});
        stz<CPLX>(mat, q * rows + i, Cz{sn * b.re + cs * xq.re, sn * b.im + cs * xq.im});
    }
}

// meta per sector: {m, n, row_off, col_off, pos, 0, 0, 0}
template <bool CPLX>
__global__ void __launch_bounds__(SVD_MAX_THREADS) sector_svd_kernel(const double* __restrict__ A, int64_t lda,
                                                                 const int* __restrict__ meta,
                                                                 const int* __restrict__ rowidx,
                                                                 const int* __restrict__ colidx, double* __restrict__ U,
                                                                 int64_t ldu, double* __restrict__ S,
                                                                 double* __restrict__ VH, int64_t ldv) {
    constexpr int E = CPLX ? 2 : 1;
    extern __shared__ double smem[];
    __shared__ int rotated;
    const int* mt = meta + 8 * blockIdx.x;
    const int m = mt[0], n = mt[1], pos = mt[4];
    const int* ri = rowidx + mt[2];
    const int* ci = colidx + mt[3];
    const bool tall = m >= n;                    // factorise A (tall) or A^H (wide)
    const int rows = tall ? m : n, k = tall ? n : m;
    double* g = smem;                            // rows x k, column-major
    double* v = g + (size_t)rows * k * E;        // k x k, column-major
    double* sig = v + (size_t)k * k * E;         // k
    int* rank = reinterpret_cast<int*>(sig + k); // k
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int SVD_THREADS = blockDim.x, SVD_WARPS = blockDim.x >> 5;

    for (int idx = tid; idx < m * n; idx += SVD_THREADS) {
        const int i = idx / n, j = idx - i * n;
        const double* src = A + ((int64_t)ri[i] * lda + ci[j]) * E;
        const Cz x = CPLX ? Cz{src[0], src[1]} : Cz{src[0], 0.0};
        if (tall) stz<CPLX>(g, j * rows + i, x);
        else stz<CPLX>(g, i * rows + j, Cz{x.re, -x.im});          // (A^H)[j, i] = conj(A[i, j])
    }
    for (int idx = tid; idx < k * k; idx += SVD_THREADS) {
        const int i = idx % k, j = idx / k;
        stz<CPLX>(v, idx, Cz{i == j ? 1.0 : 0.0, 0.0});
    }
    __syncthreads();

    // ---- one-sided Jacobi sweeps, round-robin pairing over kk = k rounded up to even players ----
    const int kk = (k + 1) & ~1;
    // rotate while |<g_p, g_q>| > tol |g_p| |g_q|, tol = eps sqrt(rows) as in LAPACK's xGESVJ
    const double eps = 2.220446049250313e-16 * sqrt((double)rows);
    for (int sweep = 0; sweep < SVD_MAX_SWEEPS && k > 1; sweep++) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int r = 0; r < kk - 1; r++) {
            for (int pi = warp; pi < kk / 2; pi += SVD_WARPS) {
                int p, q;
                if (pi == 0) { p = r; q = kk - 1; }
                else { p = (r + pi) % (kk - 1); q = (r - pi + kk - 1) % (kk - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                if (q >= k) continue;                               // dummy player of an odd k
                double a = 0.0, b = 0.0, cr = 0.0, cim = 0.0;
                for (int i = lane; i < rows; i += 32) {
                    const Cz xp = ldz<CPLX>(g, p * rows + i), xq = ldz<CPLX>(g, q * rows + i);
                    a += xp.re * xp.re + xp.im * xp.im;
                    b += xq.re * xq.re + xq.im * xq.im;
                    cr += xp.re * xq.re + xp.im * xq.im;            // conj(xp) * xq
                    cim += xp.re * xq.im - xp.im * xq.re;
                }
                a = warp_sum(a); b = warp_sum(b); cr = warp_sum(cr); cim = warp_sum(cim);
                a = __shfl_sync(0xffffffffu, a, 0); b = __shfl_sync(0xffffffffu, b, 0);
                cr = __shfl_sync(0xffffffffu, cr, 0); cim = __shfl_sync(0xffffffffu, cim, 0);
                // |c|^2 > (eps |g_p| |g_q|)^2 instead of two square roots; reciprocal square roots (one special-function
                // step + Newton, ~1 ulp) instead of sqrt + division on the dependent chain of the rotation parameters
                const double c2 = cr * cr + cim * cim;
                bool rot;
                double inv;
                if (c2 > 1e-280 && c2 < 1e280 && a < 1e140 && b < 1e140) {
                    rot = c2 > (eps * eps) * (a * b);
                    inv = rsqrt(c2);
                } else {                                             // squares leave the double range: scaled forms
                    const double absc = hypot(cr, cim);
                    rot = absc > eps * (sqrt(a) * sqrt(b)) && absc > 0.0;
                    inv = rot ? 1.0 / absc : 0.0;
                }
                if (rot) {
                    const Cz ph = {cr * inv, cim * inv};
                    const double zeta = 0.5 * (b - a) * inv;
                    const double z1 = 1.0 + zeta * zeta;
                    // (1 + zeta^2 overflows only for column norms ~1e150 apart: the rotation angle is then zero)
                    const double t = z1 < 1e300 ? copysign(1.0, zeta) / (fabs(zeta) + z1 * rsqrt(z1)) : 0.0;
                    const double cs = rsqrt(1.0 + t * t), sn = cs * t;
                    rotate_columns<CPLX>(g, rows, p, q, cs, sn, ph, lane);
                    rotate_columns<CPLX>(v, k, p, q, cs, sn, ph, lane);
                    if (lane == 0) rotated = 1;
                }
            }
            __syncthreads();
        }
        const int again = rotated;
        __syncthreads();
        if (!again) break;
    }

    // ---- singular values = column norms; descending order (ties by column index) ----
    for (int j = warp; j < k; j += SVD_WARPS) {
        double a = 0.0;
        for (int i = lane; i < rows; i += 32) {
            const Cz x = ldz<CPLX>(g, j * rows + i);
            a += x.re * x.re + x.im * x.im;
        }
        a = warp_sum(a);
        if (lane == 0) sig[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = tid; j < k; j += SVD_THREADS) {
        int rk = 0;
        const double sj = sig[j];
        for (int i = 0; i < k; i++) rk += (sig[i] > sj || (sig[i] == sj && i < j)) ? 1 : 0;
        rank[j] = rk;
        S[pos + rk] = sj;
    }
    __syncthreads();

    // ---- scatter:  A = Uo diag(sig) VHo.  tall: Uo = g / sig, VHo = v^H;  wide: Uo = v, VHo = (g / sig)^H ----
    for (int idx = tid; idx < rows * k; idx += SVD_THREADS) {
        const int i = idx % rows, j = idx / rows;
        const double sj = sig[j];
        Cz x = ldz<CPLX>(g, idx);
        if (sj > 0.0) { x.re /= sj; x.im /= sj; } else { x.re = 0.0; x.im = 0.0; }
        if (tall) {
            double* dst = U + ((int64_t)ri[i] * ldu + pos + rank[j]) * E;
            dst[0] = x.re;
            if (CPLX) dst[1] = x.im;
        } else {
            double* dst = VH + ((int64_t)(pos + rank[j]) * ldv + ci[i]) * E;
            dst[0] = x.re;
            if (CPLX) dst[1] = -x.im;
        }
    }
    for (int idx = tid; idx < k * k; idx += SVD_THREADS) {
        const int i = idx % k, j = idx / k;
        const Cz x = ldz<CPLX>(v, idx);
        if (tall) {
            double* dst = VH + ((int64_t)(pos + rank[j]) * ldv + ci[i]) * E;
            dst[0] = x.re;
            if (CPLX) dst[1] = -x.im;
        } else {
            double* dst = U + ((int64_t)ri[i] * ldu + pos + rank[j]) * E;
            dst[0] = x.re;
            if (CPLX) dst[1] = x.im;
        }
    }
}

}  // namespace

extern "C" {

size_t ptb_block_svd_max_block_bytes(void) { return 200 * 1024; }

int ptb_block_svd(int dtype, const void* a, int64_t lda, int nsec, const int32_t* meta, int max_work_elems,
                  const int32_t* rowidx, const int32_t* colidx, void* u, int64_t ldu, double* s, void* vh, int64_t ldv,
                  void* stream) {
    if (!a || !meta || !rowidx || !colidx || !u || !s || !vh || nsec < 0 || max_work_elems < 0) return PTB_ERR_BAD_ARG;
    if (nsec == 0) return PTB_OK;
    const bool cplx = dtype == PTB_COMPLEX128;
    if (!cplx && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    const size_t es = cplx ? 16 : 8;
    if ((size_t)max_work_elems * es > ptb_block_svd_max_block_bytes()) return PTB_ERR_TOO_LARGE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // shared memory: G (rows x k) + V (k x k) = max_work_elems elements, then k singular values and k ranks
    const size_t smem_need = (size_t)max_work_elems * es + (size_t)SVD_MAX_K * 12 + 64;
    // max_work_elems >= 2 k^2 for the largest block: k <= sqrt(max_work_elems / 2)
    int pairs = 1;
    while ((size_t)(2 * pairs) * (2 * pairs) * 2 < (size_t)max_work_elems && pairs < 32) pairs++;
    const int SVD_THREADS = pairs <= 8 ? 256 : (pairs <= 16 ? 512 : SVD_MAX_THREADS);
    if (cplx) {
        static DeviceFlags configured;
        PTB_TRY(ensure_dynamic_smem(configured, sector_svd_kernel<true>, 220 * 1024));
        sector_svd_kernel<true><<<nsec, SVD_THREADS, smem_need, st>>>(
            static_cast<const double*>(a), lda, meta, rowidx, colidx, static_cast<double*>(u), ldu, s,
            static_cast<double*>(vh), ldv);
    } else {
        static DeviceFlags configured;
        PTB_TRY(ensure_dynamic_smem(configured, sector_svd_kernel<false>, 220 * 1024));
        sector_svd_kernel<false><<<nsec, SVD_THREADS, smem_need, st>>>(
            static_cast<const double*>(a), lda, meta, rowidx, colidx, static_cast<double*>(u), ldu, s,
            static_cast<double*>(vh), ldv);
    }
    return cuda_status(cudaGetLastError());
}

}  // extern "C"
