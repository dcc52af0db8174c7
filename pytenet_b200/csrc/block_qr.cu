// Batched per-sector QR of a block-sparse matrix (pytenet/block_sparse_util.py:106-180), SURVEY section 8(f) rank 1.
//
// The reference loops over the quantum-number sectors and calls LAPACK's QR on each gathered block; the sectors of
// an MPS bond are many and small (median 12-36 rows at D = 2048), so one library call per sector is dominated by
// launch latency.  Here ONE launch handles all sectors that fit in shared memory: CTA s gathers its block
// A[rows_s, cols_s] into shared memory, runs the unblocked Householder factorisation with LAPACK's conventions
// (zgeqr2 / zlarfg: beta = -sign(Re alpha) |(alpha, x)|, real diagonal of R; zung2r for Q), and scatters R and Q
// straight into the block-sparse outputs at the sector's position on the intermediate bond.  Larger blocks are left
// to the caller (cuSOLVER).
#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

// threads per CTA: one warp per column of the trailing update (blocks of up to 8 columns: 256 threads, else 1024 --
// the kernel is latency bound, a warp that has to walk over several columns serialises them)
constexpr int QR_MAX_THREADS = 1024;

struct Cx {
    double re, im;
};
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cx cmulc(Cx a, Cx b) {      // conj(a) * b
    return {a.re * b.re + a.im * b.im, a.re * b.im - a.im * b.re};
}

// element access: column-major in shared memory, leading dimension m (E doubles per element)
template <bool CPLX>
__device__ __forceinline__ Cx ld(const double* s, int idx) {
    if (CPLX) return {s[2 * idx], s[2 * idx + 1]};
    return {s[idx], 0.0};
}
template <bool CPLX>
__device__ __forceinline__ void st(double* s, int idx, Cx v) {
    if (CPLX) { s[2 * idx] = v.re; s[2 * idx + 1] = v.im; }
    else s[idx] = v.re;
}

// column c of `a` (column-major, m rows) <- (I - f0 v v^H) column c for the reflector v = (1, sc * x), x = the
// UNSCALED sub-diagonal part of column j as the factorisation left it: one warp, rows j .. m-1.  Keeping x unscaled
// lets the dot product x^H a_c start before the reflector scalars (a square root and two divisions) are known, and
// removes the scaling pass with its barrier.  Not inlined, loops not unrolled: the kernel runs once per launch with a
// cold instruction cache.
template <bool CPLX>
__device__ __noinline__ void apply_reflector(double* a, int m, int j, int c, Cx sc, Cx f0, int lane) {
    Cx w = {0.0, 0.0};
#pragma unroll 1
    for (int i = j + 1 + lane; i < m; i += 32) {
        const Cx p = cmulc(ld<CPLX>(a, j * m + i), ld<CPLX>(a, c * m + i));
        w.re += p.re; w.im += p.im;
    }
    w.re = warp_sum(w.re);
    if (CPLX) w.im = warp_sum(w.im);
    w = cmulc(sc, w);                                     // v^H a_c = conj(sc) x^H a_c + a_c[j]
    // the row-j element of column c is read and written by lane 0 only (broadcast by shuffle)
    Cx top = {0.0, 0.0};
    if (lane == 0) top = ld<CPLX>(a, c * m + j);
    top.re = __shfl_sync(0xffffffffu, top.re, 0);
    if (CPLX) top.im = __shfl_sync(0xffffffffu, top.im, 0);
    w.re += top.re; w.im += top.im;                       // v_j = 1
    const Cx f = cmul(f0, w);
    if (lane == 0) st<CPLX>(a, c * m + j, Cx{top.re - f.re, top.im - f.im});
    const Cx fs = cmul(f, sc);
#pragma unroll 1
    for (int i = j + 1 + lane; i < m; i += 32) {
        const Cx g = cmul(fs, ld<CPLX>(a, j * m + i));
        const Cx o = ld<CPLX>(a, c * m + i);
        st<CPLX>(a, c * m + i, Cx{o.re - g.re, o.im - g.im});
    }
}

// meta per sector: {m, n, row_off, col_off, pos, 0, 0, 0}
template <bool CPLX>
__global__ void __launch_bounds__(QR_MAX_THREADS) sector_qr_kernel(const double* __restrict__ A, int64_t lda,
                                                                   const int* __restrict__ meta,
                                                                   const int* __restrict__ rowidx,
                                                                   const int* __restrict__ colidx, double* __restrict__ Q,
                                                                   int64_t ldq, double* __restrict__ R, int64_t ldr) {
    constexpr int E = CPLX ? 2 : 1;
    extern __shared__ double smem[];
    __shared__ double ssh[2];
    const int* mt = meta + 8 * blockIdx.x;
    const int m = mt[0], n = mt[1], pos = mt[4];
    const int* ri = rowidx + mt[2];
    const int* ci = colidx + mt[3];
    const int kmax = m < n ? m : n;
    double* a = smem;                         // m x n column-major
    double* tau = smem + (size_t)m * n * E;   // kmax reflector factors tau_j
    double* scl = tau + (size_t)kmax * E;     // kmax scalings: v_j = (1, scl_j x_j)
    double* dg = scl + (size_t)kmax * E;      // kmax diagonal entries of R (beta_j, real)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int QR_THREADS = blockDim.x, QR_WARPS = blockDim.x >> 5;

    // gather (consecutive threads walk along a row of the global matrix: columns are the contiguous direction)
#pragma unroll 1
    for (int idx = tid; idx < m * n; idx += QR_THREADS) {
        const int i = idx / n, j = idx - i * n;
        const double* src = A + ((int64_t)ri[i] * lda + ci[j]) * E;
        Cx v = CPLX ? Cx{src[0], src[1]} : Cx{src[0], 0.0};
        st<CPLX>(a, j * m + i, v);
    }
    __syncthreads();

    // ---- Householder factorisation (zgeqr2) ----
    // ONE CTA barrier per column: every warp that owns a trailing column derives the reflector scalars itself
    // (identical values) from the squared norm of the sub-diagonal part -- computed by warp 0 right after it updated
    // that column in the previous step -- and applies the reflector to its columns; nothing of column j is
    // overwritten (the scaling lives in scl_j, the diagonal of R in dg_j), so no warp waits for another inside a step.
    if (warp == 0) {
        double ss = 0.0;
#pragma unroll 1
        for (int i = 1 + lane; i < m; i += 32) {
            const Cx x = ld<CPLX>(a, i);
            ss += x.re * x.re + x.im * x.im;
        }
        ss = warp_sum(ss);
        if (lane == 0) ssh[0] = ss;
    }
    __syncthreads();
    for (int j = 0; j < kmax; j++) {
        if (warp == 0 || j + 1 + warp < n) {
            const double ss = ssh[j & 1];
            const Cx alpha = ld<CPLX>(a, j * m + j);
            Cx t = {0.0, 0.0}, scal = {0.0, 0.0};
            double beta = alpha.re;
            if (!(ss == 0.0 && alpha.im == 0.0)) {
                beta = -copysign(sqrt(alpha.re * alpha.re + alpha.im * alpha.im + ss), alpha.re);
                t = {(beta - alpha.re) / beta, -alpha.im / beta};
                // 1 / (alpha - beta)
                const double dr = alpha.re - beta, di = alpha.im;
                const double den = dr * dr + di * di;
                scal = {dr / den, -di / den};
            }
            if (tid == 0) {
                st<CPLX>(tau, j, t);
                st<CPLX>(scl, j, scal);
                dg[j] = beta;
            }
            if (t.re != 0.0 || t.im != 0.0) {
                // apply H^H = I - conj(tau) v v^H to columns j+1 .. n-1 (one warp per column)
                const Cx tc = {t.re, -t.im};
#pragma unroll 1
                for (int c = j + 1 + warp; c < n; c += QR_WARPS) apply_reflector<CPLX>(a, m, j, c, scal, tc, lane);
            }
            if (warp == 0 && j + 1 < kmax) {
                // squared norm of the next column below its diagonal (warp 0 owns column j+1 in the loop above)
                __syncwarp();
                double sn = 0.0;
#pragma unroll 1
                for (int i = j + 2 + lane; i < m; i += 32) {
                    const Cx x = ld<CPLX>(a, (j + 1) * m + i);
                    sn += x.re * x.re + x.im * x.im;
                }
                sn = warp_sum(sn);
                if (lane == 0) ssh[(j + 1) & 1] = sn;
            }
        }
        __syncthreads();
    }

    // ---- R: upper trapezoid, rows 0..kmax-1, scattered to r[pos + row, ci[col]] ----
#pragma unroll 1
    for (int idx = tid; idx < kmax * n; idx += QR_THREADS) {
        const int i = idx / n, j = idx - i * n;
        if (j < i) continue;
        const Cx v = j == i ? Cx{dg[i], 0.0} : ld<CPLX>(a, j * m + i);
        double* dst = R + ((int64_t)(pos + i) * ldr + ci[j]) * E;
        dst[0] = v.re;
        if (CPLX) dst[1] = v.im;
    }
    __syncthreads();

    // ---- Q = H_0 H_1 ... H_{kmax-1} restricted to kmax columns (zung2r, in place) ----
    // One CTA barrier per column: column j+1 is turned from its reflector into its Q form
    // (0, ..., 0, 1 - tau, -tau v) by the warp that owns it, right before that warp applies H_j to it; the other
    // warps read the reflector of column j only.
    for (int j = kmax - 1; j >= 0; j--) {
        const Cx t = ld<CPLX>(tau, j), sc = ld<CPLX>(scl, j);
        const bool reflect = t.re != 0.0 || t.im != 0.0;
#pragma unroll 1
        for (int c = j + 1 + warp; c < kmax; c += QR_WARPS) {
            if (c == j + 1) {
                const Cx tc1 = ld<CPLX>(tau, c);
                const Cx ts1 = cmul(tc1, ld<CPLX>(scl, c));
#pragma unroll 1
                for (int i = lane; i < m; i += 32) {
                    Cx v;
                    if (i < c) v = {0.0, 0.0};
                    else if (i == c) v = {1.0 - tc1.re, -tc1.im};
                    else {
                        const Cx p = cmul(ts1, ld<CPLX>(a, c * m + i));
                        v = {-p.re, -p.im};
                    }
                    st<CPLX>(a, c * m + i, v);
                }
                __syncwarp();
            }
            if (reflect) apply_reflector<CPLX>(a, m, j, c, sc, t, lane);   // H_j = I - tau v v^H on the formed column c
        }
        __syncthreads();
    }
    if (kmax > 0) {
        const Cx t0 = ld<CPLX>(tau, 0);
        const Cx ts0 = cmul(t0, ld<CPLX>(scl, 0));
#pragma unroll 1
        for (int i = tid; i < m; i += QR_THREADS) {
            Cx v;
            if (i == 0) v = {1.0 - t0.re, -t0.im};
            else {
                const Cx p = cmul(ts0, ld<CPLX>(a, i));
                v = {-p.re, -p.im};
            }
            st<CPLX>(a, i, v);
        }
    }
    __syncthreads();

    // ---- scatter Q to q[ri[row], pos + col] ----
#pragma unroll 1
    for (int idx = tid; idx < m * kmax; idx += QR_THREADS) {
        const int i = idx / kmax, j = idx - i * kmax;
        const Cx v = ld<CPLX>(a, j * m + i);
        double* dst = Q + ((int64_t)ri[i] * ldq + pos + j) * E;
        dst[0] = v.re;
        if (CPLX) dst[1] = v.im;
    }
}

}  // namespace

extern "C" {

size_t ptb_block_qr_max_block_bytes(void) { return 200 * 1024; }

int ptb_block_qr(int dtype, const void* a, int64_t lda, int nsec, const int32_t* meta, int max_block_elems,
                 const int32_t* rowidx, const int32_t* colidx, void* q, int64_t ldq, void* r, int64_t ldr, void* stream) {
    if (!a || !meta || !rowidx || !colidx || !q || !r || nsec < 0 || max_block_elems < 0) return PTB_ERR_BAD_ARG;
    if (nsec == 0) return PTB_OK;
    const bool cplx = dtype == PTB_COMPLEX128;
    if (!cplx && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    const size_t es = cplx ? 16 : 8;
    if ((size_t)max_block_elems * es > ptb_block_qr_max_block_bytes()) return PTB_ERR_TOO_LARGE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // shared memory: the gathered block + one tau per reflector (kmax <= sqrt(m n) <= max_block_elems, capped)
    size_t ntau = 1;                           // kmax = min(m, n) <= sqrt(m n) <= sqrt(max_block_elems)
    while (ntau * ntau < (size_t)max_block_elems) ntau++;
    const size_t smem_need = ((size_t)max_block_elems + 3 * ntau) * es;      // block + tau, scl, dg
    // more than 8 columns are only possible with more than 64 elements (the larger extent is at least as long)
    const int threads = max_block_elems > 64 ? QR_MAX_THREADS : 256;
    if (cplx) {
        static DeviceFlags configured;
        PTB_TRY(ensure_dynamic_smem(configured, sector_qr_kernel<true>, 220 * 1024));
        sector_qr_kernel<true><<<nsec, threads, smem_need, st>>>(
            static_cast<const double*>(a), lda, meta, rowidx, colidx, static_cast<double*>(q), ldq,
            static_cast<double*>(r), ldr);
    } else {
        static DeviceFlags configured;
        PTB_TRY(ensure_dynamic_smem(configured, sector_qr_kernel<false>, 220 * 1024));
        sector_qr_kernel<false><<<nsec, threads, smem_need, st>>>(
            static_cast<const double*>(a), lda, meta, rowidx, colidx, static_cast<double*>(q), ldq,
            static_cast<double*>(r), ldr);
    }
    return cuda_status(cudaGetLastError());
}

}  // extern "C"
