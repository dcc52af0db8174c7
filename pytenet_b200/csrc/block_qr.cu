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

constexpr int QR_THREADS = 256;
constexpr int QR_WARPS = QR_THREADS / 32;

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

// meta per sector: {m, n, row_off, col_off, pos, 0, 0, 0}
template <bool CPLX>
__global__ void __launch_bounds__(QR_THREADS) sector_qr_kernel(const double* __restrict__ A, int64_t lda,
                                                               const int* __restrict__ meta,
                                                               const int* __restrict__ rowidx,
                                                               const int* __restrict__ colidx, double* __restrict__ Q,
                                                               int64_t ldq, double* __restrict__ R, int64_t ldr) {
    constexpr int E = CPLX ? 2 : 1;
    extern __shared__ double smem[];
    __shared__ double red[32];
    __shared__ double bc[4];
    const int* mt = meta + 8 * blockIdx.x;
    const int m = mt[0], n = mt[1], pos = mt[4];
    const int* ri = rowidx + mt[2];
    const int* ci = colidx + mt[3];
    const int kmax = m < n ? m : n;
    double* a = smem;                         // m x n column-major
    double* tau = smem + (size_t)m * n * E;   // kmax entries
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // gather (consecutive threads walk along a row of the global matrix: columns are the contiguous direction)
    for (int idx = tid; idx < m * n; idx += QR_THREADS) {
        const int i = idx / n, j = idx - i * n;
        const double* src = A + ((int64_t)ri[i] * lda + ci[j]) * E;
        Cx v = CPLX ? Cx{src[0], src[1]} : Cx{src[0], 0.0};
        st<CPLX>(a, j * m + i, v);
    }
    __syncthreads();

    // ---- Householder factorisation (zgeqr2) ----
    for (int j = 0; j < kmax; j++) {
        double ss = 0.0;
        for (int i = j + 1 + tid; i < m; i += QR_THREADS) {
            const Cx x = ld<CPLX>(a, j * m + i);
            ss += x.re * x.re + x.im * x.im;
        }
        ss = block_sum(ss, red);
        if (tid == 0) {
            const Cx alpha = ld<CPLX>(a, j * m + j);
            const double xnorm = sqrt(ss);
            Cx t = {0.0, 0.0}, scal = {0.0, 0.0};
            double beta = alpha.re;
            if (!(xnorm == 0.0 && alpha.im == 0.0)) {
                beta = -copysign(sqrt(alpha.re * alpha.re + alpha.im * alpha.im + ss), alpha.re);
                t = {(beta - alpha.re) / beta, -alpha.im / beta};
                // 1 / (alpha - beta)
                const double dr = alpha.re - beta, di = alpha.im;
                const double den = dr * dr + di * di;
                scal = {dr / den, -di / den};
                st<CPLX>(a, j * m + j, Cx{beta, 0.0});
            }
            bc[0] = t.re; bc[1] = t.im; bc[2] = scal.re; bc[3] = scal.im;
            st<CPLX>(tau, j, t);
        }
        __syncthreads();
        const Cx t = {bc[0], bc[1]};
        const Cx scal = {bc[2], bc[3]};
        if (t.re != 0.0 || t.im != 0.0) {
            for (int i = j + 1 + tid; i < m; i += QR_THREADS) st<CPLX>(a, j * m + i, cmul(scal, ld<CPLX>(a, j * m + i)));
        }
        __syncthreads();
        if (t.re != 0.0 || t.im != 0.0) {
            // apply H^H = I - conj(tau) v v^H to columns j+1 .. n-1 (one warp per column)
            const Cx tc = {t.re, -t.im};
            for (int c = j + 1 + warp; c < n; c += QR_WARPS) {
                Cx w = {0.0, 0.0};
                for (int i = j + 1 + lane; i < m; i += 32) {
                    const Cx p = cmulc(ld<CPLX>(a, j * m + i), ld<CPLX>(a, c * m + i));
                    w.re += p.re; w.im += p.im;
                }
                w.re = warp_sum(w.re); w.im = warp_sum(w.im);
                // the row-j element of column c is read and written by lane 0 only (broadcast by shuffle)
                Cx top = {0.0, 0.0};
                if (lane == 0) top = ld<CPLX>(a, c * m + j);
                top.re = __shfl_sync(0xffffffffu, top.re, 0); top.im = __shfl_sync(0xffffffffu, top.im, 0);
                w.re += top.re; w.im += top.im;                       // v_j = 1
                const Cx f = cmul(tc, w);
                if (lane == 0) st<CPLX>(a, c * m + j, Cx{top.re - f.re, top.im - f.im});
                for (int i = j + 1 + lane; i < m; i += 32) {
                    const Cx g = cmul(f, ld<CPLX>(a, j * m + i));
                    const Cx o = ld<CPLX>(a, c * m + i);
                    st<CPLX>(a, c * m + i, Cx{o.re - g.re, o.im - g.im});
                }
            }
        }
        __syncthreads();
    }

    // ---- R: upper trapezoid, rows 0..kmax-1, scattered to r[pos + row, ci[col]] ----
    for (int idx = tid; idx < kmax * n; idx += QR_THREADS) {
        const int i = idx / n, j = idx - i * n;
        if (j < i) continue;
        const Cx v = ld<CPLX>(a, j * m + i);
        double* dst = R + ((int64_t)(pos + i) * ldr + ci[j]) * E;
        dst[0] = v.re;
        if (CPLX) dst[1] = v.im;
    }
    __syncthreads();

    // ---- Q = H_0 H_1 ... H_{kmax-1} restricted to kmax columns (zung2r, in place) ----
    for (int j = kmax - 1; j >= 0; j--) {
        const Cx t = ld<CPLX>(tau, j);
        // apply H_j = I - tau v v^H to the already formed columns j+1 .. kmax-1 (rows j .. m-1)
        if (t.re != 0.0 || t.im != 0.0) {
            for (int c = j + 1 + warp; c < kmax; c += QR_WARPS) {
                Cx w = {0.0, 0.0};
                for (int i = j + 1 + lane; i < m; i += 32) {
                    const Cx p = cmulc(ld<CPLX>(a, j * m + i), ld<CPLX>(a, c * m + i));
                    w.re += p.re; w.im += p.im;
                }
                w.re = warp_sum(w.re); w.im = warp_sum(w.im);
                Cx top = {0.0, 0.0};
                if (lane == 0) top = ld<CPLX>(a, c * m + j);
                top.re = __shfl_sync(0xffffffffu, top.re, 0); top.im = __shfl_sync(0xffffffffu, top.im, 0);
                w.re += top.re; w.im += top.im;
                const Cx f = cmul(t, w);
                if (lane == 0) st<CPLX>(a, c * m + j, Cx{top.re - f.re, top.im - f.im});
                for (int i = j + 1 + lane; i < m; i += 32) {
                    const Cx g = cmul(f, ld<CPLX>(a, j * m + i));
                    const Cx o = ld<CPLX>(a, c * m + i);
                    st<CPLX>(a, c * m + i, Cx{o.re - g.re, o.im - g.im});
                }
            }
        }
        __syncthreads();
        // column j of Q: (0, ..., 0, 1 - tau, -tau v_{j+1..})
        for (int i = tid; i < m; i += QR_THREADS) {
            Cx v;
            if (i < j) v = {0.0, 0.0};
            else if (i == j) v = {1.0 - t.re, -t.im};
            else {
                const Cx x = ld<CPLX>(a, j * m + i);
                const Cx p = cmul(t, x);
                v = {-p.re, -p.im};
            }
            st<CPLX>(a, j * m + i, v);
        }
        __syncthreads();
    }

    // ---- scatter Q to q[ri[row], pos + col] ----
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
    const size_t ntau = max_block_elems < 1024 ? (size_t)max_block_elems : 1024;
    const size_t smem_need = ((size_t)max_block_elems + ntau) * es;
    if (cplx) {
        static DeviceFlags configured;
        PTB_TRY(ensure_dynamic_smem(configured, sector_qr_kernel<true>, 220 * 1024));
        sector_qr_kernel<true><<<nsec, QR_THREADS, smem_need, st>>>(
            static_cast<const double*>(a), lda, meta, rowidx, colidx, static_cast<double*>(q), ldq,
            static_cast<double*>(r), ldr);
    } else {
        static DeviceFlags configured;
        PTB_TRY(ensure_dynamic_smem(configured, sector_qr_kernel<false>, 220 * 1024));
        sector_qr_kernel<false><<<nsec, QR_THREADS, smem_need, st>>>(
            static_cast<const double*>(a), lda, meta, rowidx, colidx, static_cast<double*>(q), ldq,
            static_cast<double*>(r), ldr);
    }
    return cuda_status(cudaGetLastError());
}

}  // extern "C"
