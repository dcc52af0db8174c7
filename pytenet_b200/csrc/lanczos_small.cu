// One kernel per local problem for the launch-latency regime (BASELINE config 1: README XXZ D <= 28; config 5:
// METTS at bond dimension ~4; the edges of every chain).
//
// pytenet/tdvp.py:223-238 / dmrg.py:181-189 hand krylov.py:12-57 a closure; per local problem the reference runs
// numiter x (apply_local_hamiltonian + three-term orthogonalisation), the k x k eigenproblem and the combination
// of the Lanczos vectors (krylov.py:110-139).  On a GPU that is ~3 numiter + 3 launches of kernels that each finish
// in a few microseconds.  Here ONE launch runs the whole local step: start normalisation, all Lanczos iterations
// (the three contraction steps of chain_ops.py:273-278 as plain FP64 FMA loops -- the tiles are far too small for
// the tensor pipe), the tridiagonal problem (tridiag.cuh) and exp(-dt H_eff) v as the combination of the Lanczos
// vectors.
//
// Layout of the work (second generation of this kernel): a thread-block CLUSTER of C = 1, 2, 4 or 8 CTAs, CTA c
// owning the slice j' in [c Dr / C, (c+1) Dr / C) of the right bond index of the OUTPUT.  All three contraction steps
// are local to that slice -- t1[:, :, :, j'] needs the whole vector and r[:, :, j'], the W step and l^T t2 act on the
// other indices -- so the operands sit in SHARED memory (l whole, r by slice, W, the current and the previous
// Lanczos vector whole, t1 / t2 by slice) and an iteration needs exactly three exchanges between the CTAs: the two
// scalar reductions (alpha, beta: partial sums pushed into every peer's shared memory, summed in rank order so all
// CTAs hold bit-identical values) and the new Lanczos vector (each CTA pushes its slice into every peer's copy
// through distributed shared memory).  H v, the three-term update and the normalisation of a thread's own elements
// stay in registers.  C = 1 (METTS, chain edges) is the same code with CTA barriers only; nothing but the results
// (Lanczos vectors, alpha / beta, out) touches global memory after the operands were read once.
//
// The zero-site problem (apply_local_bond_contraction, chain_ops.py:282-317) is the site problem with a
// one-dimensional physical index and no W step: w == nullptr.
#include <cooperative_groups.h>

#include "../../include/pytenet_b200.h"
#include "common.cuh"
#include "tridiag.cuh"

namespace cg = cooperative_groups;
using namespace ptb;

namespace {

constexpr int MAX_CLUSTER = 8;
constexpr int MAX_OWN = 2;                          // output elements per thread (kept in registers)
constexpr int MAX_SLOTS = 2;                        // t1 / t2 elements per thread
constexpr int MAX_THREADS = 512;
constexpr long long SMALL_RUN_MAX_MACS = 400000;    // (complex) multiply-adds of one matvec, whole cluster
constexpr long long CTA_TARGET_MACS = 20000;        // add CTAs until one CTA's share of a matvec is below this
constexpr size_t SMEM_BUDGET = 176 * 1024;          // dynamic shared memory (tridiag.cuh adds ~36 KB static)
constexpr int MISC_DOUBLES = 1 + TAYLOR_SCRATCH_DOUBLES + 64 + 2 * MAX_CLUSTER + 132 + 2 * TRIDIAG_MAX + 2;

struct RunParams {
    const double* x;        // start vector (n elements)
    const double* w;        // (cl, d, d, cr) or nullptr (zero-site problem: d == 1, cl == cr)
    const double* l;        // (Dl, cl, Dl)
    const double* r;        // (Dr, cr, Dr)
    int w_cplx;
    int Dl, d, Dr, cl, cr;
    int numiter;
    double* V;              // numiter x n Lanczos vectors
    double* scal;           // [|x|, alpha[0:k], beta[0:k-1]]
    int apply_expm;         // also out = sum_j coeff_j V_j  (expm_krylov), else the Lanczos run only
    double dt_re, dt_im;
    int out_cplx;
    double* out;
    double thresh;
    int C;                  // CTAs of the cluster
    int wjmax;              // ceil(Dr / C): slice width the shared-memory layout is sized for
    int l_smem, w_smem;     // operands staged in shared memory (else read through L1/L2)
    int ksplit;             // lanes sharing one output element in step 3 (1, 2, 4 or 8)
};

struct Plan {
    int C, threads, wjmax, l_smem, w_smem, ksplit;
    size_t smem_bytes;
    bool ok;
};

// E = 2 (complex128) is assumed when `cplx` is unknown (ptb_local_step_small_fits has no dtype argument)
Plan make_plan(int64_t Dl, int64_t d, int64_t Dr, int64_t cl, int64_t cr, bool cplx, bool has_w, bool w_cplx) {
    Plan pl{};
    pl.ok = false;
    const long long macs = (long long)Dl * d * Dr * cr * Dr + (has_w ? (long long)Dl * cl * d * Dr * d * cr : 0) +
                           (long long)Dl * d * Dr * Dl * cl;
    if (macs > SMALL_RUN_MAX_MACS) return pl;
    int C = 1;
    while (C < MAX_CLUSTER && 2 * C <= Dr && macs / C > CTA_TARGET_MACS) C *= 2;
    const size_t E = cplx ? 2 : 1;
    const size_t wn = has_w ? (size_t)(cl * d * d * cr) * (w_cplx ? 2 : 1) : 0;
    for (; C <= MAX_CLUSTER && C <= (Dr > 1 ? Dr : 1); C *= 2) {
        const int64_t wj = (Dr + C - 1) / C;
        const int64_t nown = Dl * d * wj;
        if (nown > (int64_t)MAX_OWN * MAX_THREADS) continue;
        const size_t base = (size_t)(2 * Dl * d * Dr + Dr * cr * wj + Dl * d * cr * wj + (has_w ? Dl * cl * d * wj : 0)) * E +
                            MISC_DOUBLES;
        const size_t ls = (size_t)(Dl * cl * Dl) * E;
        for (int variant = 0; variant < 4; variant++) {
            const bool lsm = !(variant & 1), wsm = wn > 0 && !(variant & 2);
            if ((variant & 2) && wn == 0) continue;
            const size_t bytes = (base + (lsm ? ls : 0) + (wsm ? wn : 0)) * sizeof(double);
            if (bytes <= SMEM_BUDGET) {
                const int64_t n1 = Dl * d * cr * wj, n2 = has_w ? Dl * cl * d * wj : 0;
                const int64_t most = n1 > n2 ? (n1 > nown ? n1 : nown) : (n2 > nown ? n2 : nown);
                pl.C = C; pl.wjmax = (int)wj; pl.l_smem = lsm; pl.w_smem = wsm; pl.smem_bytes = bytes;
                pl.threads = most <= 128 ? 128 : (most <= 256 ? 256 : MAX_THREADS);
                while ((int64_t)pl.threads * MAX_OWN < nown) pl.threads *= 2;
                pl.ok = pl.threads <= MAX_THREADS && n1 <= (int64_t)MAX_SLOTS * pl.threads &&
                        n2 <= (int64_t)MAX_SLOTS * pl.threads;
                pl.ksplit = 1;
                while (pl.ksplit < 8 && nown * pl.ksplit * 2 <= pl.threads && Dl * cl >= 4 * pl.ksplit) pl.ksplit *= 2;
                if (!pl.ok) continue;
                return pl;
            }
        }
    }
    return pl;
}

struct Cx {
    double re, im;
};

// dot product of `len` (complex) elements a[i * sa] b[i * sb] (strides in elements), two independent accumulator
// sets so that the dependent FMA chain is len / 2 long
// (not inlined, loops not unrolled beyond the two accumulator sets: these kernels run with a cold instruction cache
// -- ncu: "no instruction" is the largest stall reason of the small problems -- so the code executed per iteration
// is kept short)
template <bool CPLX>
__device__ __noinline__ Cx dot_strided(const double* __restrict__ a, int sa, const double* __restrict__ b, int sb,
                                          int len) {
    if (CPLX) {
        double rr0 = 0.0, ii0 = 0.0, ri0 = 0.0, ir0 = 0.0, rr1 = 0.0, ii1 = 0.0, ri1 = 0.0, ir1 = 0.0;
        int i = 0;
#pragma unroll 1
        for (; i + 1 < len; i += 2) {
            const double2 a0 = *reinterpret_cast<const double2*>(a + (size_t)i * sa * 2);
            const double2 b0 = *reinterpret_cast<const double2*>(b + (size_t)i * sb * 2);
            const double2 a1 = *reinterpret_cast<const double2*>(a + (size_t)(i + 1) * sa * 2);
            const double2 b1 = *reinterpret_cast<const double2*>(b + (size_t)(i + 1) * sb * 2);
            rr0 = fma(a0.x, b0.x, rr0); ii0 = fma(a0.y, b0.y, ii0); ri0 = fma(a0.x, b0.y, ri0); ir0 = fma(a0.y, b0.x, ir0);
            rr1 = fma(a1.x, b1.x, rr1); ii1 = fma(a1.y, b1.y, ii1); ri1 = fma(a1.x, b1.y, ri1); ir1 = fma(a1.y, b1.x, ir1);
        }
        if (i < len) {
            const double2 a0 = *reinterpret_cast<const double2*>(a + (size_t)i * sa * 2);
            const double2 b0 = *reinterpret_cast<const double2*>(b + (size_t)i * sb * 2);
            rr0 = fma(a0.x, b0.x, rr0); ii0 = fma(a0.y, b0.y, ii0); ri0 = fma(a0.x, b0.y, ri0); ir0 = fma(a0.y, b0.x, ir0);
        }
        return {(rr0 + rr1) - (ii0 + ii1), (ri0 + ri1) + (ir0 + ir1)};
    } else {
        double s0 = 0.0, s1 = 0.0;
        int i = 0;
#pragma unroll 1
        for (; i + 1 < len; i += 2) {
            s0 = fma(a[(size_t)i * sa], b[(size_t)i * sb], s0);
            s1 = fma(a[(size_t)(i + 1) * sa], b[(size_t)(i + 1) * sb], s1);
        }
        if (i < len) s0 = fma(a[(size_t)i * sa], b[(size_t)i * sb], s0);
        return {s0 + s1, 0.0};
    }
}

// Sum of one double per thread over the CTA and, for C > 1, over the cluster.  Every thread of every CTA returns the
// same bits: warp partials are summed in warp order by every thread, CTA partials in rank order by every thread.
// `slot` alternates between consecutive calls (0: alpha, 1: beta / norms) so that a partial is never overwritten
// before every reader passed a later barrier.
__device__ __noinline__ double all_sum(double v, int slot, double* red, double* part, int C, unsigned rank) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[slot * 32 + warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll 1
    for (int i = 0; i < nwarp; i++) s += red[slot * 32 + i];
    if (C == 1) return s;
    cg::cluster_group cluster = cg::this_cluster();
    if ((int)threadIdx.x < C) {
        double* remote = cluster.map_shared_rank(part, threadIdx.x);
        remote[slot * MAX_CLUSTER + rank] = s;
    }
    cluster.sync();
    s = 0.0;
#pragma unroll 1
    for (int c = 0; c < C; c++) s += part[slot * MAX_CLUSTER + c];
    return s;
}

#ifdef PTB_LS_PROFILE
// phase clocks of thread 0 of CTA 0 (tools/local_step_probe.py builds a private copy of this file with the macro)
__device__ long long g_ls_prof[16];
#define LSPROF(i)                                                       \
    do {                                                                \
        if (tid == 0 && rank == 0) {                                    \
            const long long t_ = clock64();                             \
            g_ls_prof[i] += t_ - tlast;                                 \
            tlast = t_;                                                 \
        }                                                               \
    } while (0)
#else
#define LSPROF(i) do { } while (0)
#endif

template <bool CPLX>
__global__ void __launch_bounds__(MAX_THREADS) lanczos_small_kernel(const RunParams p) {
    constexpr int E = CPLX ? 2 : 1;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int C = p.C;
    const unsigned rank = C > 1 ? cg::this_cluster().block_rank() : 0u;
#ifdef PTB_LS_PROFILE
    long long tlast = clock64();
#endif
    const int Dl = p.Dl, d = p.d, Dr = p.Dr, cl = p.cl, cr = p.cr, k = p.numiter;
    const int n = Dl * d * Dr;
    const int j0 = (int)(((long long)rank * Dr) / C), j1 = (int)(((long long)(rank + 1) * Dr) / C);
    const int wj = j1 - j0;                       // this CTA's slice of the right bond index
    const int wjm = p.wjmax;
    const int nown = Dl * d * wj;                 // output elements owned by this CTA
    const bool has_w = p.w != nullptr;
    const int WE = p.w_cplx ? 2 : 1;
    const int wn = has_w ? cl * d * d * cr * WE : 0;

    // ---- shared-memory layout (identical in every CTA of the cluster: sized with wjmax) ----
    double* vfull = sm;                                            // 2 x n: current / previous Lanczos vector
    double* r_s = vfull + (size_t)2 * n * E;                       // r[j, K, j0 + jl] as [j][K][jl]
    double* t1_s = r_s + (size_t)Dr * cr * wjm * E;                // [(i, s)][K][jl]
    double* t2_s = t1_s + (size_t)Dl * d * cr * wjm * E;           // [i][k][s'][jl]
    double* l_s = t2_s + (has_w ? (size_t)Dl * cl * d * wjm * E : 0);
    double* tay = l_s + (p.l_smem ? (size_t)Dl * cl * Dl * E : 0); // scratch of the k x k solve (16-byte aligned)
    if ((tay - sm) & 1) tay++;                                     // float64 layouts can end on an odd double
    double* red = tay + TAYLOR_SCRATCH_DOUBLES;                    // 2 x 32 warp partials
    double* part = red + 64;                                       // 2 x MAX_CLUSTER CTA partials
    double* scal_s = part + 2 * MAX_CLUSTER;                       // [|x|, alpha, beta] for the k x k solve
    double* coeff_s = scal_s + 132;                                // 2 TRIDIAG_MAX doubles + k_eff behind them
    double* w_s = coeff_s + 2 * TRIDIAG_MAX + 2;                   // W (real: possibly an odd number of doubles)

    // ---- operands into shared memory ----
    if (p.l_smem) {
        const int nl = Dl * cl * Dl * E;
#pragma unroll 2
        for (int i = tid; i < nl; i += NT) l_s[i] = p.l[i];
    }
    {
        const int nr = Dr * cr * wj;
#pragma unroll 1
        for (int idx = tid; idx < nr; idx += NT) {
            const int jl = idx % wj, jk = idx / wj;                // jk = j * cr + K
            const double* src = p.r + ((size_t)jk * Dr + j0 + jl) * E;
            r_s[(size_t)idx * E] = src[0];
            if (CPLX) r_s[(size_t)idx * E + 1] = src[1];
        }
    }
    if (p.w_smem) {
#pragma unroll 1
        for (int i = tid; i < wn; i += NT) w_s[i] = p.w[i];
    }
    const double* lsrc = p.l_smem ? l_s : p.l;
    const double* wsrc = p.w_smem ? w_s : p.w;

    // ---- per-thread index tables (the same in every iteration: no integer division inside the Lanczos loop) ----
    // step 1 / step 2 outputs idx = tid + q NT
    const int ncol = cr * wj, n1 = Dl * d * ncol, n2 = has_w ? Dl * cl * d * wj : 0;
    static_assert(MAX_SLOTS == 2 && MAX_OWN == 2, "slot tables are two scalars each (selected, not indexed)");
    int o1a0 = -1, o1a1 = -1, o1b0 = 0, o1b1 = 0, o2w0 = -1, o2w1 = -1, o2t0 = 0, o2t1 = 0;
#pragma unroll 1
    for (int q = 0; q < MAX_SLOTS; q++) {
        const int idx = tid + q * NT;
        int a1 = -1, b1 = 0, w2 = -1, t2o = 0;
        if (idx < n1) {
            const int row = idx / ncol, col = idx - row * ncol;
            a1 = row * Dr * E;
            b1 = col * E;
        }
        if (idx < n2) {
            const int q1 = idx / wj, jl = idx - q1 * wj;
            const int q2 = q1 / d, sp = q1 - q2 * d;
            const int i = q2 / cl, kk = q2 - i * cl;
            w2 = ((kk * d + sp) * d) * cr * WE;                    // w[kk, sp, :, :] as [s][K]
            t2o = (i * d * cr * wj + jl) * E;                      // t1[i, :, :, jl]: + (s * cr + K) * wj * E
        }
        if (q == 0) { o1a0 = a1; o1b0 = b1; o2w0 = w2; o2t0 = t2o; }
        else { o1a1 = a1; o1b1 = b1; o2w1 = w2; o2t1 = t2o; }
    }
    // own elements (step 3 and the vector operations): KS lanes share one element when the CTA has threads to
    // spare (the l^T t2 sum is split over them), else up to MAX_OWN elements per thread.  Local index
    // e = sj * Dl + i' with sj = s' * wj + jl (lanes run along i', the contiguous index of l); g = global element
    // index (i', s', j0 + jl).  Only the first lane of a group (`owner`) stores and contributes to the sums.
    // Lane layout of a split element: a warp holds EPW = 32 / KS consecutive elements, lane = kpart * EPW + element,
    // so the eight lanes served together by a 128-bit shared-memory load read eight consecutive i' of one row of l
    // (conflict free) and one t2 entry (broadcast); with kpart as the fast lane index the KS partial sums of an
    // element sat on the same banks (ncu: 262 k load bank conflicts per launch at the README bulk site).
    const int KS = p.ksplit;
    const int EPW = 32 / KS;
    const int kpart = (tid & 31) / EPW;
    const bool owner = kpart == 0;
    int gidx[MAX_OWN], o3l[MAX_OWN], o3t[MAX_OWN];
#pragma unroll
    for (int m = 0; m < MAX_OWN; m++) {
        const int e = KS > 1 ? (m == 0 ? (tid >> 5) * EPW + (tid & 31) % EPW : nown) : tid + m * NT;
        gidx[m] = -1; o3l[m] = 0; o3t[m] = 0;
        if (e < nown) {
            const int ip = e % Dl, sj = e / Dl;
            const int sp = sj / wj, jl = sj - sp * wj;
            gidx[m] = (ip * d + sp) * Dr + j0 + jl;
            o3l[m] = ip * E;                                       // l[:, :, i']: + (i * cl + k) * Dl * E
            o3t[m] = sj * E;                                       // t2[:, :, s', jl]: + (i * cl + k) * d * wj * E
        }
    }

    // ---- v_0 = x / |x|  (krylov.py:31-33); every CTA normalises the whole vector (no exchange needed) ----
    double acc = 0.0;
#pragma unroll 1
    for (int i = tid; i < n * E; i += NT) acc += p.x[i] * p.x[i];
    {
        // CTA-level sum only: all CTAs read the same x with the same thread count, hence the same bits
        const double nrm = sqrt(all_sum(acc, 1, red, part, 1, 0u));
        const double ninv = 1.0 / nrm;
#pragma unroll 1
        for (int i = tid; i < n * E; i += NT) vfull[i] = p.x[i] * ninv;
        if (tid == 0) {
            scal_s[0] = nrm;
            if (rank == 0) p.scal[0] = nrm;
        }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MAX_OWN; m++) {
        if (gidx[m] >= 0 && owner) {
            const size_t g = (size_t)gidx[m] * E;
            p.V[g] = vfull[g];
            if (CPLX) p.V[g + 1] = vfull[g + 1];
        }
    }
    // every CTA of the cluster must be running before the first store into a peer's shared memory
    if (C > 1) cg::this_cluster().sync();
    double* alpha_g = p.scal + 1;
    double* beta_g = alpha_g + k;
    const int nk3 = Dl * cl;                                       // length of the step-3 sum
    const int len3 = (nk3 - kpart + KS - 1) / KS;                  // this lane's share of it
    LSPROF(0);

    for (int j = 0; j < k; j++) {
        const double* cur = vfull + (size_t)(j & 1) * n * E;
        double* oth = vfull + (size_t)((j + 1) & 1) * n * E;       // v_{j-1}; receives v_{j+1}
        // step 1: t1[(i,s), (K,jl)] = sum_j v[(i,s), j] r[j, (K, j0+jl)]                  chain_ops.py:273
#pragma unroll 1
        for (int q = 0; q < MAX_SLOTS; q++) {
            const int oa = q == 0 ? o1a0 : o1a1, ob = q == 0 ? o1b0 : o1b1;
            if (oa >= 0) {
                const Cx s = dot_strided<CPLX>(cur + oa, 1, r_s + ob, ncol, Dr);
                const int idx = tid + q * NT;
                t1_s[(size_t)idx * E] = s.re;
                if (CPLX) t1_s[(size_t)idx * E + 1] = s.im;
            }
        }
        __syncthreads();
        LSPROF(1);
        const double* t2 = t1_s;                  // zero-site problem: no W step, (i, K, jl) is already (i, k, jl)
        if (has_w) {
            // step 2: t2[i, k, s', jl] = sum_{s,K} w[k, s', s, K] t1[i, s, K, jl]          chain_ops.py:276
            // (all d cr terms, zero or not: a branch-free loop of independent loads)
            const int nsk = d * cr;
#pragma unroll 1
            for (int q = 0; q < MAX_SLOTS; q++) {
                const int ow = q == 0 ? o2w0 : o2w1, ot = q == 0 ? o2t0 : o2t1;
                if (ow >= 0) {
                    const double* wrow = wsrc + ow;
                    const double* tin = t1_s + ot;
                    double re = 0.0, im = 0.0, re2 = 0.0, im2 = 0.0;
#pragma unroll 2
                    for (int sk = 0; sk < nsk; sk++) {
                        const double wr = wrow[sk * WE];
                        const double tr = tin[(size_t)sk * wj * E];
                        re = fma(wr, tr, re);
                        if (CPLX) {
                            const double ti = tin[(size_t)sk * wj * E + 1];
                            im = fma(wr, ti, im);
                            if (WE == 2) {
                                const double wi = wrow[sk * 2 + 1];
                                re2 = fma(wi, ti, re2); im2 = fma(wi, tr, im2);
                            }
                        }
                    }
                    const int idx = tid + q * NT;
                    t2_s[(size_t)idx * E] = re - re2;
                    if (CPLX) t2_s[(size_t)idx * E + 1] = im + im2;
                }
            }
            __syncthreads();
            t2 = t2_s;
        }
        LSPROF(2);
        // step 3: y[i', s', jl] = sum_{i,k} l[i, k, i'] t2[i, k, s', jl]  (chain_ops.py:278), own elements in
        // registers, and alpha_j = Re <v_j, y>  (krylov.py:41)
        Cx y[MAX_OWN];
        acc = 0.0;
#pragma unroll
        for (int m = 0; m < MAX_OWN; m++) {
            y[m] = {0.0, 0.0};
            if (gidx[m] >= 0)
                y[m] = dot_strided<CPLX>(lsrc + o3l[m] + (size_t)kpart * Dl * E, Dl * KS,
                                         t2 + o3t[m] + (size_t)kpart * d * wj * E, d * wj * KS, len3);
            if (m == 0 && KS > 1) {
                for (int o = EPW; o < 32; o <<= 1) {                // all lanes of the warp take part
                    y[0].re += __shfl_xor_sync(0xffffffffu, y[0].re, o);
                    if (CPLX) y[0].im += __shfl_xor_sync(0xffffffffu, y[0].im, o);
                }
            }
            if (gidx[m] >= 0 && owner) {
                const size_t g = (size_t)gidx[m] * E;
                acc = fma(y[m].re, cur[g], acc);
                if (CPLX) acc = fma(y[m].im, cur[g + 1], acc);
            }
        }
        LSPROF(3);
        const double al = all_sum(acc, 0, red, part, C, rank);
        if (tid == 0) {
            scal_s[1 + j] = al;
            if (rank == 0) alpha_g[j] = al;
        }
        LSPROF(4);
        if (j == k - 1) break;                     // the closing matvec only contributes alpha (krylov.py:53-56)
        const double bp = j > 0 ? scal_s[1 + k + j - 1] : 0.0;
        acc = 0.0;
#pragma unroll
        for (int m = 0; m < MAX_OWN; m++) {
            if (gidx[m] >= 0 && owner) {
                const size_t g = (size_t)gidx[m] * E;
                double sub = al * cur[g];          // same association as krylov.py:42
                if (j > 0) sub = sub + bp * oth[g];
                y[m].re = y[m].re - sub;
                acc = fma(y[m].re, y[m].re, acc);
                if (CPLX) {
                    double subi = al * cur[g + 1];
                    if (j > 0) subi = subi + bp * oth[g + 1];
                    y[m].im = y[m].im - subi;
                    acc = fma(y[m].im, y[m].im, acc);
                }
            }
        }
        const double be = sqrt(all_sum(acc, 1, red, part, C, rank));
        if (tid == 0) {
            scal_s[1 + k + j] = be;
            if (rank == 0) beta_g[j] = be;
        }
        LSPROF(5);
        // v_{j+1} = y / beta_j: own slice into every CTA's copy (distributed shared memory) and into V.
        // `oth` (v_{j-1}) is free: every CTA read it before arriving at the barrier inside the beta reduction.
        double* vout = p.V + (size_t)(j + 1) * n * E;
        const double binv = 1.0 / be;
#pragma unroll
        for (int m = 0; m < MAX_OWN; m++) {
            if (gidx[m] >= 0 && owner) {
                const size_t g = (size_t)gidx[m] * E;
                const double vr = y[m].re * binv, vi = CPLX ? y[m].im * binv : 0.0;
                oth[g] = vr;
                if (CPLX) oth[g + 1] = vi;
                vout[g] = vr;
                if (CPLX) vout[g + 1] = vi;
                if (C > 1) {
                    cg::cluster_group cluster = cg::this_cluster();
                    for (int c = 1; c < C; c++) {
                        double* peer = cluster.map_shared_rank(oth, (rank + c) % C);
                        peer[g] = vr;
                        if (CPLX) peer[g + 1] = vi;
                    }
                }
            }
        }
        if (C > 1) cg::this_cluster().sync();
        else __syncthreads();
        LSPROF(6);
    }
    if (!p.apply_expm) return;

    // coeff = U (|x| exp(dt w) U[0, :]); out = sum_{j < k_eff} coeff_j v_j               krylov.py:122-136
    // (every CTA solves the k x k problem redundantly from its bit-identical copy of the scalars)
    __syncthreads();
    int* keff = reinterpret_cast<int*>(coeff_s + 2 * TRIDIAG_MAX);
    tridiag_expm_solve(scal_s, k, p.thresh, p.dt_re, p.dt_im, coeff_s, keff, tay);
    LSPROF(7);
    const int ke = *keff;
#pragma unroll
    for (int m = 0; m < MAX_OWN; m++) {
        if (gidx[m] < 0 || !owner) continue;
        const size_t g = (size_t)gidx[m];
        double re = 0.0, im = 0.0;
        for (int j = 0; j < ke; j++) {
            const double cre = coeff_s[2 * j], cim = coeff_s[2 * j + 1];
            if (CPLX) {
                // own elements of V: written by this very thread
                const double xr = p.V[((size_t)j * n + g) * 2], xi = p.V[((size_t)j * n + g) * 2 + 1];
                re += cre * xr - cim * xi;
                im += cre * xi + cim * xr;
            } else {
                const double x = p.V[(size_t)j * n + g];
                re += cre * x;
                if (p.out_cplx) im += cim * x;
            }
        }
        if (p.out_cplx) { p.out[2 * g] = re; p.out[2 * g + 1] = im; }
        else p.out[g] = re;
    }
    LSPROF(8);
}

template <bool CPLX>
int launch(const RunParams& p, const Plan& pl, cudaStream_t st) {
    static DeviceFlags configured;
    PTB_TRY(ensure_dynamic_smem(configured, lanczos_small_kernel<CPLX>, (int)SMEM_BUDGET));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.C, 1, 1);
    cfg.blockDim = dim3(pl.threads, 1, 1);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pl.C > 1 ? 1 : 0;
    return cuda_status(cudaLaunchKernelEx(&cfg, lanczos_small_kernel<CPLX>, p));
}

}  // namespace

#ifdef PTB_LS_PROFILE
extern "C" int lsprobe_read(long long* host16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host16, g_ls_prof, sizeof(long long) * 16);
    if (reset) {
        long long z[16] = {0};
        cudaMemcpyToSymbol(g_ls_prof, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" {

int ptb_local_step_small_fits(int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter) {
    if (Dl <= 0 || d <= 0 || Dr <= 0 || chi_l <= 0 || chi_r <= 0 || numiter < 1 || numiter > TRIDIAG_MAX) return 0;
    if (Dl > 4096 || d > 4096 || Dr > 4096 || chi_l > 4096 || chi_r > 4096) return 0;
    // dtype and MPO tensor unknown here: complex128 operands and a complex W are the largest layout
    return make_plan(Dl, d, Dr, chi_l, chi_r, true, true, true).ok ? 1 : 0;
}

size_t ptb_local_step_small_workspace_bytes(int dtype, int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r) {
    (void)dtype; (void)Dl; (void)d; (void)Dr; (void)chi_l; (void)chi_r;
    return 16;      // intermediates live in shared memory; the argument is kept for ABI stability
}

int ptb_local_step_small(int dtype, const void* x, const void* w, int w_is_complex, const void* l, const void* r,
                         int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter, void* V,
                         double* scal, int apply_expm, double dt_re, double dt_im, int out_is_complex, void* out,
                         void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    if (!x || !l || !r || !V || !scal) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    const bool cplx = dtype == PTB_COMPLEX128;
    if (!cplx && w_is_complex) return PTB_ERR_BAD_DTYPE;
    if (!ptb_local_step_small_fits(Dl, d, Dr, chi_l, chi_r, numiter)) return PTB_ERR_TOO_LARGE;
    if (!w && (d != 1 || chi_l != chi_r)) return PTB_ERR_BAD_ARG;
    if (apply_expm && (!out || (!out_is_complex && (cplx || dt_im != 0.0)))) return PTB_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(V) % 16) return PTB_ERR_ALIGNMENT;
    if (cplx && (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(l) % 16 ||
                 reinterpret_cast<uintptr_t>(r) % 16))
        return PTB_ERR_ALIGNMENT;
    const Plan pl = make_plan(Dl, d, Dr, chi_l, chi_r, cplx, w != nullptr, w_is_complex != 0);
    if (!pl.ok) return PTB_ERR_TOO_LARGE;
    const int64_t n = Dl * d * Dr;
    RunParams p;
    p.x = static_cast<const double*>(x); p.w = static_cast<const double*>(w);
    p.l = static_cast<const double*>(l); p.r = static_cast<const double*>(r);
    p.w_cplx = w_is_complex ? 1 : 0;
    p.Dl = (int)Dl; p.d = (int)d; p.Dr = (int)Dr; p.cl = (int)chi_l; p.cr = (int)chi_r;
    p.numiter = numiter;
    p.V = static_cast<double*>(V);
    p.scal = scal;
    p.apply_expm = apply_expm ? 1 : 0;
    p.dt_re = dt_re; p.dt_im = dt_im;
    p.out_cplx = out_is_complex ? 1 : 0;
    p.out = static_cast<double*>(out);
    p.thresh = 100.0 * (double)n * 2.220446049250313e-16;      // krylov.py:44
    p.C = pl.C; p.wjmax = pl.wjmax; p.l_smem = pl.l_smem; p.w_smem = pl.w_smem; p.ksplit = pl.ksplit;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return cplx ? launch<true>(p, pl, st) : launch<false>(p, pl, st);
}

}  // extern "C"
