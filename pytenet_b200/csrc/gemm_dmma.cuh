// FP64 / complex128 GEMM engine on the sm_100a FP64 tensor pipe (DMMA.8x8x4).
//
//   C[b] (M x N, row-major, ldc)  (+)=  op(A[b]) (M x K)  *  op(B[b]) (K x N)
//
// Every contraction of the effective-Hamiltonian path (reference
// pytenet/chain_ops.py:50-56, 94-98, 273-278, 314-316) is expressed as one of
// these GEMMs on the *original* C-ordered tensors: the axis permutations the
// reference performs on the host (np.tensordot transposes + copies) become the
// operand layout flags below, so no transposed copy of any tensor ever exists.
//
//   A_KC : A[m,k] at A[m*lda + k]  ("N", k contiguous)   else A[k*lda + m] ("T")
//   B_KC : B[k,n] at B[n*ldb + k]  ("T", k contiguous)   else B[k*ldb + n] ("N")
//   CONJB: use conj(B)   (complex only)
//
// Complex arithmetic is built from four real DMMA sub-products per 8x8x4 tile on
// de-interleaved fragments: HBM and shared memory keep NumPy's interleaved
// complex128 layout; one 16-byte LDS fetches (re, im) of a fragment element and
// the sign flip for the imaginary product (and for conj) is a register negate.
//
// Tiling: CTA = 256 threads = 8 warps (4 along M x 2 along N); warp tile
// 32 x 32 complex (or 32 x 64 real) = 64 FP64 accumulators per thread; CTA tile
// 128 x 64 complex / 128 x 128 real; K step 8 complex / 16 real per stage;
// 4-stage cp.async (LDGSTS) ring with zero-fill predication, so arbitrary
// (ragged, tiny, odd) extents are handled by the same kernel.
// Shared-memory layouts are chosen per operand so that the 8x4 fragment reads
// are bank-conflict free:
//   k-contiguous operand  -> [k/4][mn][k%4]   (a warp fragment is one 512 B line)
//   mn-contiguous operand -> [k][mn + pad]    (pad = 32 B)
#pragma once
#include "common.cuh"

namespace ptb {

struct GemmParams {
    const double* A;
    const double* B;
    double* C;
    int M, N, K;
    int64_t lda, ldb, ldc;  // leading dimensions in elements (complex elements if CPLX)
    int64_t sA, sB, sC;     // batch strides in elements
    int batch;
    int accumulate;  // C += A*B instead of C = A*B
    int tiles_m, tiles_n;
};

template <bool CPLX>
struct GemmCfg {
    static constexpr int E = CPLX ? 2 : 1;  // doubles per element
    static constexpr int THREADS = 256;
    static constexpr int BM = 128;
    static constexpr int BN = CPLX ? 64 : 128;
    static constexpr int BK = CPLX ? 8 : 16;
    static constexpr int WTM = 32;
    static constexpr int WTN = CPLX ? 32 : 64;
    static constexpr int MT = WTM / 8;
    static constexpr int NT = WTN / 8;
    static constexpr int PAD = CPLX ? 2 : 4;  // elements (= 32 bytes)
    static constexpr int STAGES = 4;
    static constexpr int SA = BK * (BM + PAD) * E;  // doubles per A stage (max of both layouts)
    static constexpr int SB = BK * (BN + PAD) * E;
    static constexpr int SMEM_BYTES = STAGES * (SA + SB) * 8;
};

// Copy one operand tile (MN_T x BK) global -> shared with zero fill outside [MN) x [K).
// VEC_D = doubles per cp.async (2 -> 16 B, 1 -> 8 B).
template <bool CPLX, bool KC, int MN_T, int VEC_D>
__device__ __forceinline__ void load_tile(double* __restrict__ smem, const double* __restrict__ g, int64_t ld,
                                          int mn0, int k0, int MN, int K, int tid) {
    using Cfg = GemmCfg<CPLX>;
    constexpr int E = Cfg::E;
    constexpr int BK = Cfg::BK;
    constexpr int PAD = Cfg::PAD;
    if (KC) {
        constexpr int CPR = BK * E / VEC_D;  // chunks per mn-row
        constexpr int TOTAL = MN_T * CPR;
#pragma unroll
        for (int c0 = 0; c0 < TOTAL; c0 += Cfg::THREADS) {
            const int c = c0 + tid;
            if (TOTAL % Cfg::THREADS != 0 && c >= TOTAL) break;
            const int row = c / CPR;
            const int kd = (c % CPR) * VEC_D;  // offset in doubles inside the row
            const int ke = kd / E;             // element index along k
            const int mn = mn0 + row;
            int nd = (K - k0) * E - kd;  // valid doubles from here
            nd = nd < 0 ? 0 : (nd > VEC_D ? VEC_D : nd);
            if (mn >= MN) nd = 0;
            const double* src = nd > 0 ? g + ((int64_t)mn * ld + k0) * E + kd : g;
            double* dst = smem + ((((ke >> 2) * MN_T + row) << 2) + (ke & 3)) * E + (kd % E);
            if (VEC_D == 2)
                cp_async_16(dst, src, nd * 8);
            else
                cp_async_8(dst, src, nd * 8);
        }
    } else {
        constexpr int CPR = MN_T * E / VEC_D;  // chunks per k-row
        constexpr int TOTAL = BK * CPR;
#pragma unroll
        for (int c0 = 0; c0 < TOTAL; c0 += Cfg::THREADS) {
            const int c = c0 + tid;
            if (TOTAL % Cfg::THREADS != 0 && c >= TOTAL) break;
            const int kr = c / CPR;
            const int md = (c % CPR) * VEC_D;  // offset in doubles along mn
            int nd = (MN - mn0) * E - md;
            nd = nd < 0 ? 0 : (nd > VEC_D ? VEC_D : nd);
            if (k0 + kr >= K) nd = 0;
            const double* src = nd > 0 ? g + ((int64_t)(k0 + kr) * ld + mn0) * E + md : g;
            double* dst = smem + kr * (MN_T + PAD) * E + md;
            if (VEC_D == 2)
                cp_async_16(dst, src, nd * 8);
            else
                cp_async_8(dst, src, nd * 8);
        }
    }
}

template <bool CPLX, bool A_KC, bool B_KC, bool CONJB, int VEC_D>
__global__ void __launch_bounds__(256, 1) gemm_dmma_kernel(const GemmParams p) {
    using Cfg = GemmCfg<CPLX>;
    constexpr int E = Cfg::E, BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
    constexpr int MT = Cfg::MT, NT = Cfg::NT, PAD = Cfg::PAD, STAGES = Cfg::STAGES;
    constexpr int SA = Cfg::SA, SB = Cfg::SB;

    extern __shared__ __align__(16) double smem[];
    double* sA = smem;
    double* sB = smem + STAGES * SA;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;

    // grouped tile order: consecutive CTAs share B column panels / A row panels in L2
    constexpr int GROUP = 8;
    const int tile = blockIdx.x;
    const int per_group = GROUP * p.tiles_n;
    const int grp = tile / per_group;
    const int first_m = grp * GROUP;
    const int gsize = min(p.tiles_m - first_m, GROUP);
    const int tm = first_m + (tile % per_group) % gsize;
    const int tn = (tile % per_group) / gsize;
    const int m0 = tm * BM, n0 = tn * BN;

    const int64_t bz = blockIdx.y;
    const double* __restrict__ Ag = p.A + bz * p.sA * E;
    const double* __restrict__ Bg = p.B + bz * p.sB * E;
    double* __restrict__ Cg = p.C + bz * p.sC * E;

    const int KT = (p.K + BK - 1) / BK;

    // accumulators: real: acc[mt][nt][2]; complex: re at [..][0..1], im at [..][2..3]
    double acc[MT][NT][2 * E];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int e = 0; e < 2 * E; e++) acc[i][j][e] = 0.0;

    auto issue = [&](int kt) {
        const int s = kt % STAGES;
        load_tile<CPLX, A_KC, BM, VEC_D>(sA + s * SA, Ag, p.lda, m0, kt * BK, p.M, p.K, tid);
        load_tile<CPLX, B_KC, BN, VEC_D>(sB + s * SB, Bg, p.ldb, n0, kt * BK, p.N, p.K, tid);
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < KT) issue(s);
        cp_async_commit();
    }

    // per-thread fragment base offsets (in doubles) inside a stage
    const int a_row = wm * Cfg::WTM + g;
    const int b_col = wn * Cfg::WTN + g;
    const int a_off = A_KC ? ((a_row << 2) + q) * E : (q * (BM + PAD) + a_row) * E;
    const int b_off = B_KC ? ((b_col << 2) + q) * E : (q * (BN + PAD) + b_col) * E;
    constexpr int A_KS = A_KC ? BM * 4 * E : 4 * (BM + PAD) * E;  // stride per k4 step
    constexpr int B_KS = B_KC ? BN * 4 * E : 4 * (BN + PAD) * E;
    constexpr int A_MT = A_KC ? 8 * 4 * E : 8 * E;  // stride per 8-row m-tile
    constexpr int B_NT = B_KC ? 8 * 4 * E : 8 * E;

    for (int kt = 0; kt < KT; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) issue(nk);
            cp_async_commit();
        }
        const double* As = sA + (kt % STAGES) * SA + a_off;
        const double* Bs = sB + (kt % STAGES) * SB + b_off;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ks++) {
            if constexpr (CPLX) {
                double2 af[MT], bf[NT];
                double bneg[NT];
#pragma unroll
                for (int i = 0; i < MT; i++)
                    af[i] = *reinterpret_cast<const double2*>(As + ks * A_KS + i * A_MT);
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    bf[j] = *reinterpret_cast<const double2*>(Bs + ks * B_KS + j * B_NT);
                    bneg[j] = -bf[j].y;
                }
#pragma unroll
                for (int i = 0; i < MT; i++)
#pragma unroll
                    for (int j = 0; j < NT; j++) {
                        // re += ar*br - ai*bi (conj: + ai*bi);  im += ar*bi + ai*br (conj: - ar*bi)
                        const double bi_re = CONJB ? bf[j].y : bneg[j];
                        const double bi_im = CONJB ? bneg[j] : bf[j].y;
                        dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
                        dmma_8x8x4(acc[i][j][2], acc[i][j][3], af[i].x, bi_im);
                        dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i].y, bi_re);
                        dmma_8x8x4(acc[i][j][2], acc[i][j][3], af[i].y, bf[j].x);
                    }
            } else {
                double af[MT], bf[NT];
#pragma unroll
                for (int i = 0; i < MT; i++) af[i] = As[ks * A_KS + i * A_MT];
#pragma unroll
                for (int j = 0; j < NT; j++) bf[j] = Bs[ks * B_KS + j * B_NT];
#pragma unroll
                for (int i = 0; i < MT; i++)
#pragma unroll
                    for (int j = 0; j < NT; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns C[row][col], C[row][col+1] per 8x8 tile (col = 2q)
#pragma unroll
    for (int i = 0; i < MT; i++) {
        const int row = m0 + wm * Cfg::WTM + i * 8 + g;
        if (row >= p.M) continue;
        double* crow = Cg + (int64_t)row * p.ldc * E;
#pragma unroll
        for (int j = 0; j < NT; j++) {
            const int col = n0 + wn * Cfg::WTN + j * 8 + 2 * q;
            if constexpr (CPLX) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if (col + e < p.N) {
                        double2* dst = reinterpret_cast<double2*>(crow + (int64_t)(col + e) * 2);
                        double2 v = make_double2(acc[i][j][e], acc[i][j][2 + e]);
                        if (p.accumulate) {
                            const double2 old = *dst;
                            v.x += old.x;
                            v.y += old.y;
                        }
                        *dst = v;
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if (col + e < p.N) {
                        double v = acc[i][j][e];
                        if (p.accumulate) v += crow[col + e];
                        crow[col + e] = v;
                    }
                }
            }
        }
    }
}

// ---- host-side dispatch ---------------------------------------------------------------

template <bool CPLX, bool A_KC, bool B_KC, bool CONJB, int VEC_D>
static int launch_gemm_inst(const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<CPLX>;
    auto kern = gemm_dmma_kernel<CPLX, A_KC, B_KC, CONJB, VEC_D>;
    static DeviceFlags configured;  // per instantiation and device
    PTB_TRY(ensure_dynamic_smem(configured, kern, Cfg::SMEM_BYTES));
    GemmParams q = p;
    int done = 0;
    while (done < p.batch) {  // gridDim.y limit
        const int nb = (p.batch - done) > 65535 ? 65535 : (p.batch - done);
        q.A = p.A + (int64_t)done * p.sA * Cfg::E;
        q.B = p.B + (int64_t)done * p.sB * Cfg::E;
        q.C = p.C + (int64_t)done * p.sC * Cfg::E;
        q.batch = nb;
        dim3 grid((unsigned)(p.tiles_m * p.tiles_n), (unsigned)nb, 1);
        kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(q);
        PTB_CUDA_TRY(cudaGetLastError());
        done += nb;
    }
    return PTB_OK;
}

// transA: 0 = "N" (A is M x K row-major), 1 = "T" (A stored K x M row-major)
// transB: 0 = "N" (B is K x N row-major), 1 = "T" (B stored N x K row-major)
template <bool CPLX>
static int launch_gemm(int transA, int transB, int conjB, GemmParams p, cudaStream_t stream) {
    using Cfg = GemmCfg<CPLX>;
    if (p.M < 0 || p.N < 0 || p.K < 0 || p.batch < 0) return PTB_ERR_BAD_ARG;
    if (p.M == 0 || p.N == 0 || p.batch == 0) return PTB_OK;
    p.tiles_m = (p.M + Cfg::BM - 1) / Cfg::BM;
    p.tiles_n = (p.N + Cfg::BN - 1) / Cfg::BN;
    if ((int64_t)p.tiles_m * p.tiles_n > 0x7fffffffLL) return PTB_ERR_TOO_LARGE;
    // 16-byte cp.async needs 16-byte aligned global chunks: always true for complex128;
    // for float64 it needs even leading dimensions / strides and aligned bases.
    bool vec2 = true;
    if (!CPLX) {
        vec2 = ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.B)) % 16 == 0) &&
               (p.lda % 2 == 0) && (p.ldb % 2 == 0) && (p.sA % 2 == 0) && (p.sB % 2 == 0);
    } else {
        if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.B) |
             reinterpret_cast<uintptr_t>(p.C)) % 16 != 0)
            return PTB_ERR_ALIGNMENT;
    }
    if ((reinterpret_cast<uintptr_t>(p.A) | reinterpret_cast<uintptr_t>(p.B) | reinterpret_cast<uintptr_t>(p.C)) % 8 != 0)
        return PTB_ERR_ALIGNMENT;
    const bool a_kc = (transA == 0), b_kc = (transB != 0);
    const bool cj = CPLX && conjB;
    constexpr int V2 = 2;
    constexpr int V1 = CPLX ? 2 : 1;  // complex elements are always 16-byte chunks
    const int sel = (a_kc ? 4 : 0) | (b_kc ? 2 : 0) | (cj ? 1 : 0);
    if (vec2) {
        switch (sel) {
            case 0: return launch_gemm_inst<CPLX, false, false, false, V2>(p, stream);
            case 1: return launch_gemm_inst<CPLX, false, false, CPLX, V2>(p, stream);
            case 2: return launch_gemm_inst<CPLX, false, true, false, V2>(p, stream);
            case 3: return launch_gemm_inst<CPLX, false, true, CPLX, V2>(p, stream);
            case 4: return launch_gemm_inst<CPLX, true, false, false, V2>(p, stream);
            case 5: return launch_gemm_inst<CPLX, true, false, CPLX, V2>(p, stream);
            case 6: return launch_gemm_inst<CPLX, true, true, false, V2>(p, stream);
            case 7: return launch_gemm_inst<CPLX, true, true, CPLX, V2>(p, stream);
        }
    } else {
        switch (sel) {
            case 0: return launch_gemm_inst<CPLX, false, false, false, V1>(p, stream);
            case 2: return launch_gemm_inst<CPLX, false, true, false, V1>(p, stream);
            case 4: return launch_gemm_inst<CPLX, true, false, false, V1>(p, stream);
            case 6: return launch_gemm_inst<CPLX, true, true, false, V1>(p, stream);
        }
    }
    return PTB_ERR_BAD_ARG;
}

}  // namespace ptb
