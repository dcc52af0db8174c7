// Grouped FP64 / complex128 GEMM for the sector-packed block-sparse path (BASELINE config 3).
//
// One launch executes a list of independent output tiles, each with its own operand pointers, leading
// dimensions and extents (`ptb_group_tile`, include/pytenet_b200.h): C(m x n) = A^T (m x k) B (k x n) with A stored
// k x m and B stored k x n (both "k-major": the contraction index is the row index), C row-major.  The
// packed layouts of pytenet_b200/sector_packed.py are built so that every quantum-number sector group is
// exactly such a product with TWO large, stacked extents and only one sector-sized one, so the fixed
// 128 x 64 (complex) tile of the dense engine wastes little: visited / exact flops = 1.2 at the config-3
// shape, where the work lists over dense-layout tensors visited 3.8 x the exact flops.
//
// Same machine model as gemm_ws.cuh, whose primitives it reuses: persistent grid, one CTA per SM; a producer
// warp stages operand tiles with cp.async.bulk (one bulk copy per k-row per operand -- the operands are
// k-major with arbitrary per-tile base pointers, so no tensor map is needed) into a full/empty mbarrier ring;
// 16 consumer warps hold 32 x 16 complex warp tiles and issue DMMA.8x8x4.  Differences: the tile list is a
// device table (sorted by decreasing k on the host: longest-processing-time-first over the persistent CTAs),
// the k-tail is processed at the DMMA granularity of 4 (not BK = 16), and warps whose whole 32 x 16 sub-tile
// lies outside the tile's valid m x n extent skip their MMAs.
#pragma once
#include "gemm_ws.cuh"

namespace ptb {

struct GroupTile {      // layout of ptb_group_tile (64 bytes); offsets in elements from the launch's base pointers
    long long a_off;    // A element (k, m) at A[(a_off + k * lda + m) * E]: the tile's first column
    long long b_off;    // B element (k, n) at B[(b_off + k * ldb + n) * E]
    long long c_off;    // C element (m, n) at C[(c_off + m * ldc + n) * E]: the tile's origin
    int lda, ldb, ldc;
    int m, n, k;        // valid extents of this tile (m <= BM, n <= BN) and the contraction length
    int accumulate;     // C += instead of C =
    int pad_[3];
};
static_assert(sizeof(GroupTile) == 64, "ptb_group_tile must be 64 bytes");

// Small-tile configuration: 64 x 32 complex (64 x 64 real) CTA tiles, warp tiles 16 x 8 (16 x 16), for groups whose
// sector-sized extents leave most of a 128 x 64 tile empty (the zero-site problem: both output extents of step 3 are
// sector sized; fragmented sector profiles with a median of ~14 rows).  Fewer DMMA per fragment load than the large
// tile (the 16 warps of a CTA are always all busy, though), more and shorter tiles per launch, a deeper stage ring.
template <bool CPLX>
struct GroupSmallCfg {
    static constexpr int E = CPLX ? 2 : 1;
    static constexpr int BM = 64;
    static constexpr int BN = CPLX ? 32 : 64;
    static constexpr int BK = 16;
    static constexpr int WTM = 16;
    static constexpr int WTN = CPLX ? 8 : 16;
    static constexpr int MT = WTM / 8;
    static constexpr int NT = WTN / 8;
    static constexpr int PAD = CPLX ? 2 : 4;
    static constexpr int STAGES = 8;
    static constexpr int CONSUMER_WARPS = 16;
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
    static constexpr int PRODUCER_REGS = 24;
    static constexpr int CONSUMER_REGS = 112;
    static constexpr int SA = BK * (BM + PAD) * E;
    static constexpr int SB = BK * (BN + PAD) * E;
    static constexpr int SMEM_BYTES = STAGES * (SA + SB) * 8 + 2 * STAGES * 8 + 128;
};

template <bool CPLX, class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
gemm_grouped_kernel(const double* __restrict__ Abase, const double* __restrict__ Bbase, double* __restrict__ Cbase,
                    const GroupTile* __restrict__ tiles, int ntiles) {
    constexpr int E = Cfg::E, BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK;
    constexpr int MT = Cfg::MT, NT = Cfg::NT, PAD = Cfg::PAD, STAGES = Cfg::STAGES;
    constexpr int SA = Cfg::SA, SB = Cfg::SB;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    double* sA = reinterpret_cast<double*>(base);
    double* sB = sA + STAGES * SA;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * SB);
    uint64_t* empty = full + STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;

    if (warp >= Cfg::CONSUMER_WARPS) {
        // ===================== producer =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(Cfg::PRODUCER_REGS));
        if (warp != Cfg::CONSUMER_WARPS) return;
        for (int u = blockIdx.x; u < ntiles; u += gridDim.x) {
            const GroupTile t = tiles[u];
            const uint32_t a_bytes = (uint32_t)(t.m * E * 8), b_bytes = (uint32_t)(t.n * E * 8);
            for (int k0 = 0; k0 < t.k; k0 += BK) {
                mbar_wait(&empty[stage], phase ^ 1);
                double* a_st = sA + stage * SA;
                double* b_st = sB + stage * SB;
                const int rows = min(BK, t.k - k0);
                const int rows4 = (rows + 3) & ~3;          // consumers read whole k-steps of 4
                if (rows4 > rows) {
                    // k-tail rows multiply valid data of the other operand: they must read as zeros
                    for (int r = rows; r < rows4; r++) {
                        double2* za = reinterpret_cast<double2*>(a_st + r * (BM + PAD) * E);
                        double2* zb = reinterpret_cast<double2*>(b_st + r * (BN + PAD) * E);
                        for (int i = lane; i < BM * E / 2; i += 32) za[i] = make_double2(0.0, 0.0);
                        for (int i = lane; i < BN * E / 2; i += 32) zb[i] = make_double2(0.0, 0.0);
                    }
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_expect_tx(&full[stage], (uint32_t)rows * (a_bytes + b_bytes));
                __syncwarp();
                // one bulk copy per k-row and operand, spread over the lanes of the producer warp
                for (int r = lane; r < 2 * rows; r += 32) {
                    if (r < rows)
                        bulk_load_1d(a_st + r * (BM + PAD) * E, Abase + (t.a_off + (int64_t)(k0 + r) * t.lda) * E, a_bytes,
                                     &full[stage]);
                    else
                        bulk_load_1d(b_st + (r - rows) * (BN + PAD) * E,
                                     Bbase + (t.b_off + (int64_t)(k0 + r - rows) * t.ldb) * E, b_bytes, &full[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ===================== consumers =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Cfg::CONSUMER_REGS));
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int a_off = (q * (BM + PAD) + wm * Cfg::WTM + g) * E;
    const int b_off = (q * (BN + PAD) + wn * Cfg::WTN + g) * E;
    constexpr int A_KS = 4 * (BM + PAD) * E, B_KS = 4 * (BN + PAD) * E;
    constexpr int A_MT = 8 * E, B_NT = 8 * E;

    for (int u = blockIdx.x; u < ntiles; u += gridDim.x) {
        const GroupTile t = tiles[u];
        const bool active = (wm * Cfg::WTM < t.m) && (wn * Cfg::WTN < t.n);

        double acc[MT][NT][2 * E];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++)
#pragma unroll
                for (int e = 0; e < 2 * E; e++) acc[i][j][e] = 0.0;

        for (int k0 = 0; k0 < t.k; k0 += BK) {
            mbar_wait(&full[stage], phase);
            if (active) {
                const double* As = sA + stage * SA + a_off;
                const double* Bs = sB + stage * SB + b_off;
                const int nks = (min(BK, t.k - k0) + 3) >> 2;
#pragma unroll
                for (int ks = 0; ks < BK / 4; ks++) {
                    if (ks < nks) {
                        if constexpr (CPLX) {
                            double2 af[MT], bf[NT];
#pragma unroll
                            for (int i = 0; i < MT; i++)
                                af[i] = *reinterpret_cast<const double2*>(As + ks * A_KS + i * A_MT);
#pragma unroll
                            for (int j = 0; j < NT; j++)
                                bf[j] = *reinterpret_cast<const double2*>(Bs + ks * B_KS + j * B_NT);
#pragma unroll
                            for (int i = 0; i < MT; i++)
#pragma unroll
                                for (int j = 0; j < NT; j++) {
                                    const double nbi = -bf[j].y;
                                    dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
                                    dmma_8x8x4(acc[i][j][2], acc[i][j][3], af[i].x, bf[j].y);
                                    dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i].y, nbi);
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
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }

        if (!active) continue;
        // epilogue: bounds at the tile's valid extent (overlaps the producer's prefetch of the next tile)
#pragma unroll
        for (int i = 0; i < MT; i++) {
            const int row = wm * Cfg::WTM + i * 8 + g;
            if (row >= t.m) continue;
            double* crow = Cbase + (t.c_off + (int64_t)row * t.ldc) * E;
#pragma unroll
            for (int j = 0; j < NT; j++) {
                const int col = wn * Cfg::WTN + j * 8 + 2 * q;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if (col + e < t.n) {
                        if constexpr (CPLX) {
                            double2* dst = reinterpret_cast<double2*>(crow + (int64_t)(col + e) * 2);
                            double2 v = make_double2(acc[i][j][e], acc[i][j][2 + e]);
                            if (t.accumulate) {
                                const double2 old = *dst;
                                v.x += old.x;
                                v.y += old.y;
                            }
                            *dst = v;
                        } else {
                            double v = acc[i][j][e];
                            if (t.accumulate) v += crow[col + e];
                            crow[col + e] = v;
                        }
                    }
                }
            }
        }
    }
}

template <bool CPLX, class Cfg>
static int launch_grouped(const double* a, const double* b, double* c, const GroupTile* tiles, int ntiles,
                          cudaStream_t stream) {
    if (ntiles <= 0) return PTB_OK;
    auto kern = gemm_grouped_kernel<CPLX, Cfg>;
    static DeviceFlags configured;
    PTB_TRY(ensure_dynamic_smem(configured, kern, Cfg::SMEM_BYTES));
    const int num_sms = device_sm_count();
    const int grid = ntiles < num_sms ? ntiles : num_sms;
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a, b, c, tiles, ntiles);
    return cuda_status(cudaGetLastError());
}

}  // namespace ptb
