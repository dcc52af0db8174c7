// Warp-specialised persistent FP64 / complex128 GEMM for sm_100a (second-generation engine).
//
// Same contract and operand-layout flags as gemm_dmma.cuh (C = op(A) op(B), row-major C,
// no operand is ever transposed in memory), but built the Hopper/Blackwell way:
//
//   * persistent grid: one CTA per SM walks the tile list; 16 consumer warps
//     (4 per SM sub-partition) + 1 producer warp.
//   * the producer stages operand tiles with the TMA unit and the bulk-copy engine:
//       k-contiguous operand  -> cp.async.bulk.tensor (3-D tensor map, box = 4 k x MN rows;
//                                out-of-range rows / columns are zero-filled by hardware)
//       mn-contiguous operand -> one cp.async.bulk per k-row into a padded row
//                                (conflict-free fragment reads need a 32 B row skew that a
//                                dense TMA box cannot give)
//     completion is tracked with mbarrier transaction counts; a STAGES-deep full/empty
//     mbarrier ring replaces every __syncthreads of the first-generation kernel, and the
//     producer runs ahead across tile boundaries, so the prologue of tile i+1 overlaps the
//     epilogue of tile i.
//   * consumers hold 32 x 16 complex (32 x 32 real) warp tiles = 32 FP64 accumulators per
//     thread, which leaves room for 16 resident warps per SM; fragments are 16-byte LDS of
//     interleaved (re, im); complex products are 4 real DMMA.8x8x4 with the sign folded into
//     the instruction's operand-negate modifier.
//
// FP64 has no tcgen05 / TMEM path on Blackwell (the UMMA kinds are f16, tf32, i8, f8f6f4 and
// the block-scaled formats), so the FP64 tensor pipe is driven by warp-level mma.sync.
#pragma once
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "gemm_dmma.cuh"

namespace ptb {

template <bool CPLX>
struct WsCfg {
    static constexpr int E = CPLX ? 2 : 1;
    static constexpr int BM = 128;
    static constexpr int BN = CPLX ? 64 : 128;
    static constexpr int BK = 16;
    static constexpr int WTM = 32;
    static constexpr int WTN = CPLX ? 16 : 32;
    static constexpr int MT = WTM / 8;
    static constexpr int NT = WTN / 8;
    static constexpr int PAD = CPLX ? 2 : 4;  // elements = 32 bytes
    static constexpr int STAGES = CPLX ? 4 : 6;
    static constexpr int CONSUMER_WARPS = 16;
    // registers are granted per group of 4 warps: the producer gets its own warpgroup (one
    // active warp) and hands most of its registers to the consumers with setmaxnreg
    static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
    static constexpr int PRODUCER_REGS = 24;
    static constexpr int CONSUMER_REGS = 112;  // 16 x 512 extra registers <= the 72 x 128 the producer group releases
    static constexpr int SA = BK * (BM + PAD) * E;  // doubles per stage
    static constexpr int SB = BK * (BN + PAD) * E;
    static constexpr int SMEM_BYTES = STAGES * (SA + SB) * 8 + 2 * STAGES * 8 + 128;
};

struct WsParams {
    const double* A;
    const double* B;
    double* C;
    int M, N, K;
    int64_t lda, ldb, ldc;
    int64_t sA, sB, sC;
    int batch;
    int accumulate;
    int tiles_m, tiles_n;
    int batched_a, batched_b;  // 0 when the batch stride is 0 (operand shared by all batches)
    // split-K: every output tile is computed by `split_k` work units over disjoint k ranges that write
    // partial tiles to Cpart ([batch][split][M][N], dense); a second kernel sums them in fixed order
    int split_k;
    double* Cpart;
    // tail splitting: the tiles of the last, partially filled wave (tile index >= tail_begin) are each
    // computed by `tail_split` work units over disjoint k ranges, so the wave finishes in 1/tail_split
    // of a tile time; partial tiles go to Cpart as dense [unit][BM][BN] buffers and are summed in fixed
    // order by a second kernel.  tail_split == 1 disables it.
    int tail_split;
    long long tail_begin;
    // sector-banded GEMM: per (batch, tile) range [lo, hi) of k-tiles that can be non-zero given the
    // quantum-number sectors of the operands (nullptr = full range); hi <= lo skips the tile's main loop
    const int2* ktab;
    // segmented GEMM (SEG kernels): tile t accumulates the segments segs[seg_ptr[t] .. seg_ptr[t+1]); a segment is
    // {first k-tile, end k-tile, selector, unused}; sel_off[2*selector + {0,1}] are element offsets added to the
    // A / B base pointers for that segment (mn-contiguous operands only).  Used for sums over an outer index
    // (e.g. the left MPO bond in step 3 of the sector path) that would otherwise need one accumulating launch each.
    const int* seg_ptr;
    const int4* segs;
    const long long* sel_off;
    // optional permutation of the tile indices (sector path): work unit u processes tile order[u].  The host sorts
    // the tiles by decreasing work so that the round-robin assignment of units to the persistent CTAs is balanced
    // although the per-tile k ranges differ widely.
    const int* order;
    // tile rows per L2 super-tile of the grouped tile order (ws_tile_coords): the A panels of `group` tile rows stay in
    // L2 while the B panels stream past them
    int group;
};

// Rows per super-tile: a wave of 148 tiles then covers `group` tile rows x 148/group tile columns; the A panels of a
// group (group x BM x K elements) must fit in the 126 MB L2 next to the streaming B panels and the C write-back.
static inline int ws_group_rows(int tiles_m, int K, bool cplx) {
    // measured at the headline shape (ncu dram__bytes, profiles/r02_gemm_ncu.md): 8 rows 6.8 GB per launch, 16 rows
    // 10.0 GB, 24 rows 12.9 GB, 32 rows 21.2 GB at equal duration -- beyond 8 rows the A panels no longer survive in
    // L2 next to the streaming B panels and the C write-back
    (void)tiles_m; (void)K; (void)cplx;
    return 8;
}

// ---- mbarrier / TMA primitives ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error becomes a trap (reported as a launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 2 GHz
    }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
            "r"(smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Producer side for one operand tile of one k-tile.  Returns the transaction bytes that the
// asynchronous copies will signal.  `issue` selects between the accounting pass (all lanes,
// also zero-fills k-tail rows of mn-contiguous operands) and the copy pass (lane 0 only).
template <bool CPLX, bool KC, int MN_T>
__device__ __forceinline__ uint32_t producer_operand(bool issue, double* stage, const CUtensorMap* map,
                                                     const double* g, int64_t ld, int mn0, int k0, int MN, int K,
                                                     int bcoord, uint64_t* bar, int lane) {
    using Cfg = WsCfg<CPLX>;
    constexpr int E = Cfg::E, BK = Cfg::BK, PAD = Cfg::PAD;
    if (KC) {
        if (issue) {
#pragma unroll
            for (int c = 0; c < BK / 4; c++)
                tma_load_3d(stage + c * MN_T * 4 * E, map, bar, (k0 + 4 * c) * E, mn0, bcoord);
        }
        return (uint32_t)(MN_T * BK * E * 8);
    } else {
        const int rows = min(BK, K - k0);
        const int cols = min(MN_T, MN - mn0);
        const uint32_t row_bytes = (uint32_t)(cols * E * 8);
        if (issue) {
            for (int r = 0; r < rows; r++)
                bulk_load_1d(stage + r * (MN_T + PAD) * E, g + ((int64_t)(k0 + r) * ld + mn0) * E, row_bytes, bar);
        } else if (rows < BK) {
            // k-tail: rows beyond K must read as zeros (they multiply valid data of the other operand)
            double2* z = reinterpret_cast<double2*>(stage + rows * (MN_T + PAD) * E);
            const int n16 = (BK - rows) * (MN_T + PAD) * E / 2;
            for (int i = lane; i < n16; i += 32) z[i] = make_double2(0.0, 0.0);
        }
        return (uint32_t)rows * row_bytes;
    }
}

// Work unit -> (tile, k-split index, number of splits, kind of destination)
struct WsUnit {
    long long tile;   // global tile index (batch-major)
    int sk, nsplit;   // this unit covers k-tiles [KT*sk/nsplit, KT*(sk+1)/nsplit)
    int dest;         // 0 = C, 1 = full-matrix split-K partial, 2 = tail tile buffer
    long long slot;   // tile-buffer index for dest == 2
};

__device__ __forceinline__ WsUnit ws_decode_unit(const WsParams& p, long long u) {
    WsUnit w;
    if (p.split_k > 1) {
        w.tile = u / p.split_k;
        w.sk = (int)(u - w.tile * p.split_k);
        w.nsplit = p.split_k;
        w.dest = 1;
        w.slot = 0;
    } else if (p.tail_split > 1 && u >= p.tail_begin) {
        const long long v = u - p.tail_begin;
        w.tile = p.tail_begin + v / p.tail_split;
        w.sk = (int)(v % p.tail_split);
        w.nsplit = p.tail_split;
        w.dest = 2;
        w.slot = v;
    } else {
        w.tile = p.order != nullptr ? (long long)p.order[u] : u;
        w.sk = 0;
        w.nsplit = 1;
        w.dest = 0;
        w.slot = 0;
    }
    return w;
}

// tile index -> (batch, tile row, tile column) in the grouped order (GROUP tile rows share B panels in L2)
__device__ __forceinline__ void ws_tile_coords(const WsParams& p, long long t, int& bz, int& tm, int& tn) {
    const int GROUP = p.group;
    const int tiles_per_batch = p.tiles_m * p.tiles_n;
    bz = (int)(t / tiles_per_batch);
    const int tile = (int)(t - (long long)bz * tiles_per_batch);
    if (p.order != nullptr) {      // scheduled tiles are given in table order (batch, tile row, tile column)
        tm = tile / p.tiles_n;
        tn = tile - tm * p.tiles_n;
        return;
    }
    const int per_group = GROUP * p.tiles_n;
    const int grp = tile / per_group;
    const int first_m = grp * GROUP;
    const int gsize = min(p.tiles_m - first_m, GROUP);
    tm = first_m + (tile % per_group) % gsize;
    tn = (tile % per_group) / gsize;
}

template <bool CPLX, bool A_KC, bool B_KC, bool CONJB, bool SEG = false>
__global__ void __launch_bounds__(WsCfg<CPLX>::THREADS, 1)
gemm_ws_kernel(const WsParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    using Cfg = WsCfg<CPLX>;
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

    const int KT_all = (p.K + BK - 1) / BK;
    const int tiles_per_batch = p.tiles_m * p.tiles_n;
    const long long total_tiles = (long long)tiles_per_batch * p.batch;
    const long long total = p.split_k > 1 ? total_tiles * p.split_k
                          : (p.tail_split > 1 ? p.tail_begin + (total_tiles - p.tail_begin) * p.tail_split
                                              : total_tiles);

    int stage = 0;
    uint32_t phase = 0;

    if (warp >= Cfg::CONSUMER_WARPS) {
        // ===================== producer warpgroup (one working warp) =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(Cfg::PRODUCER_REGS));
        if (warp != Cfg::CONSUMER_WARPS) return;
        for (long long u = blockIdx.x; u < total; u += gridDim.x) {
            const WsUnit un = ws_decode_unit(p, u);
            int kt_begin = (int)(((long long)KT_all * un.sk) / un.nsplit);
            int kt_end = (int)(((long long)KT_all * (un.sk + 1)) / un.nsplit);
            int bz, tm, tn;
            ws_tile_coords(p, un.tile, bz, tm, tn);
            const int m0 = tm * BM, n0 = tn * BN;
            if (p.ktab != nullptr) {
                const int2 kr = p.ktab[(long long)bz * tiles_per_batch + (long long)tm * p.tiles_n + tn];
                kt_begin = max(kr.x, 0);
                kt_end = min(kr.y, KT_all);
            }
            const double* Ag = p.A + (int64_t)bz * p.sA * E;
            const double* Bg = p.B + (int64_t)bz * p.sB * E;
            const int ba = p.batched_a ? bz : 0, bb = p.batched_b ? bz : 0;
            int seg = 0, seg_end = 1;
            if constexpr (SEG) {
                const long long tix = (long long)bz * tiles_per_batch + (long long)tm * p.tiles_n + tn;
                seg = p.seg_ptr[tix];
                seg_end = p.seg_ptr[tix + 1];
            }
            for (; seg < seg_end; seg++) {
                const double* Ags = Ag;
                const double* Bgs = Bg;
                if constexpr (SEG) {
                    const int4 sg = p.segs[seg];
                    kt_begin = max(sg.x, 0);
                    kt_end = min(sg.y, KT_all);
                    Ags += p.sel_off[2 * sg.z] * E;
                    Bgs += p.sel_off[2 * sg.z + 1] * E;
                }
                for (int kt = kt_begin; kt < kt_end; kt++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    double* a_st = sA + stage * SA;
                    double* b_st = sB + stage * SB;
                    const int k0 = kt * BK;
                    uint32_t bytes = producer_operand<CPLX, A_KC, BM>(false, a_st, &tmA, Ags, p.lda, m0, k0, p.M, p.K,
                                                                      ba, &full[stage], lane);
                    bytes += producer_operand<CPLX, B_KC, BN>(false, b_st, &tmB, Bgs, p.ldb, n0, k0, p.N, p.K, bb,
                                                              &full[stage], lane);
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&full[stage], bytes);
                        producer_operand<CPLX, A_KC, BM>(true, a_st, &tmA, Ags, p.lda, m0, k0, p.M, p.K, ba,
                                                         &full[stage], lane);
                        producer_operand<CPLX, B_KC, BN>(true, b_st, &tmB, Bgs, p.ldb, n0, k0, p.N, p.K, bb,
                                                         &full[stage], lane);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(Cfg::CONSUMER_REGS));
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int a_row = wm * Cfg::WTM + g;
    const int b_col = wn * Cfg::WTN + g;
    const int a_off = A_KC ? ((a_row << 2) + q) * E : (q * (BM + PAD) + a_row) * E;
    const int b_off = B_KC ? ((b_col << 2) + q) * E : (q * (BN + PAD) + b_col) * E;
    constexpr int A_KS = A_KC ? BM * 4 * E : 4 * (BM + PAD) * E;
    constexpr int B_KS = B_KC ? BN * 4 * E : 4 * (BN + PAD) * E;
    constexpr int A_MT = A_KC ? 8 * 4 * E : 8 * E;
    constexpr int B_NT = B_KC ? 8 * 4 * E : 8 * E;

    for (long long u = blockIdx.x; u < total; u += gridDim.x) {
        const WsUnit un = ws_decode_unit(p, u);
        int kt_begin = (int)(((long long)KT_all * un.sk) / un.nsplit);
        int kt_end = (int)(((long long)KT_all * (un.sk + 1)) / un.nsplit);
        int bz, tm, tn;
        ws_tile_coords(p, un.tile, bz, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        if (p.ktab != nullptr) {
            const int2 kr = p.ktab[(long long)bz * tiles_per_batch + (long long)tm * p.tiles_n + tn];
            kt_begin = max(kr.x, 0);
            kt_end = min(kr.y, KT_all);
        }

        // banded GEMM with accumulate == 2: tiles with an empty k range are left untouched (the caller keeps the
        // structurally empty part of C zeroed once, instead of rewriting the zeros on every call)
        if (!SEG && p.ktab != nullptr && p.accumulate == 2 && kt_end <= kt_begin) continue;

        double acc[MT][NT][2 * E];
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++)
#pragma unroll
                for (int e = 0; e < 2 * E; e++) acc[i][j][e] = 0.0;

        int seg = 0, seg_end = 1;
        if constexpr (SEG) {
            const long long tix = (long long)bz * tiles_per_batch + (long long)tm * p.tiles_n + tn;
            seg = p.seg_ptr[tix];
            seg_end = p.seg_ptr[tix + 1];
        }
        for (; seg < seg_end; seg++) {
        if constexpr (SEG) {
            const int4 sg = p.segs[seg];
            kt_begin = max(sg.x, 0);
            kt_end = min(sg.y, KT_all);
        }
        for (int kt = kt_begin; kt < kt_end; kt++) {
            mbar_wait(&full[stage], phase);
            const double* As = sA + stage * SA + a_off;
            const double* Bs = sB + stage * SB + b_off;
#pragma unroll
            for (int ks = 0; ks < BK / 4; ks++) {
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
                            const double bi_re = CONJB ? bf[j].y : nbi;   // multiplies a.im into re
                            const double bi_im = CONJB ? nbi : bf[j].y;   // multiplies a.re into im
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        }   // segments

        // epilogue (overlaps the producer's prefetch of the next tile)
        if (un.dest == 2) {
            // tail unit: the whole BM x BN partial tile goes to its dense buffer (no bounds: the reduction
            // kernel only reads the in-range part)
            double* __restrict__ Tb = p.Cpart + un.slot * (long long)(BM * BN * E);
#pragma unroll
            for (int i = 0; i < MT; i++) {
                const int row = wm * Cfg::WTM + i * 8 + g;
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    const int col = wn * Cfg::WTN + j * 8 + 2 * q;
                    if constexpr (CPLX) {
                        double2* dst = reinterpret_cast<double2*>(Tb + ((int64_t)row * BN + col) * 2);
                        dst[0] = make_double2(acc[i][j][0], acc[i][j][2]);
                        dst[1] = make_double2(acc[i][j][1], acc[i][j][3]);
                    } else {
                        *reinterpret_cast<double2*>(Tb + (int64_t)row * BN + col) = make_double2(acc[i][j][0], acc[i][j][1]);
                    }
                }
            }
            continue;
        }
        const bool partial = un.dest == 1;
        const int64_t ldc_eff = partial ? (int64_t)p.N : p.ldc;
        const bool accum = !partial && p.accumulate == 1;
        {
            double* __restrict__ Cg =
                partial ? p.Cpart + ((int64_t)bz * p.split_k + un.sk) * (int64_t)p.M * p.N * E
                        : p.C + (int64_t)bz * p.sC * E;
#pragma unroll
            for (int i = 0; i < MT; i++) {
                const int row = m0 + wm * Cfg::WTM + i * 8 + g;
                if (row >= p.M) continue;
                double* crow = Cg + (int64_t)row * ldc_eff * E;
#pragma unroll
                for (int j = 0; j < NT; j++) {
                    const int col = n0 + wn * Cfg::WTN + j * 8 + 2 * q;
                    if constexpr (CPLX) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            if (col + e < p.N) {
                                double2* dst = reinterpret_cast<double2*>(crow + (int64_t)(col + e) * 2);
                                double2 v = make_double2(acc[i][j][e], acc[i][j][2 + e]);
                                if (accum) {
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
                                if (accum) v += crow[col + e];
                                crow[col + e] = v;
                            }
                        }
                    }
                }
            }
        }
    }
}

// Sum the tail-unit partial tiles in fixed order and store them at their place in C.
template <bool CPLX>
__global__ void __launch_bounds__(256) tail_reduce_kernel(const WsParams p) {
    using Cfg = WsCfg<CPLX>;
    constexpr int E = Cfg::E, BM = Cfg::BM, BN = Cfg::BN;
    const long long t = p.tail_begin + blockIdx.x;
    int bz, tm, tn;
    ws_tile_coords(p, t, bz, tm, tn);
    const int m0 = tm * BM, n0 = tn * BN;
    const double* __restrict__ src = p.Cpart + (long long)blockIdx.x * p.tail_split * (BM * BN * E);
    double* __restrict__ Cg = p.C + (int64_t)bz * p.sC * E;
    for (int idx = threadIdx.x; idx < BM * BN * E; idx += blockDim.x) {
        const int row = idx / (BN * E);
        const int cd = idx - row * (BN * E);          // column in doubles
        if (m0 + row >= p.M || n0 * E + cd >= p.N * E) continue;
        double acc = 0.0;
        for (int sidx = 0; sidx < p.tail_split; sidx++) acc += src[(long long)sidx * (BM * BN * E) + idx];
        double* dst = Cg + ((int64_t)(m0 + row) * p.ldc + n0) * E + cd;
        if (p.accumulate) acc += *dst;
        *dst = acc;
    }
}

// Sum the split-K partial tiles in fixed order: C[b][m][n] (+)= sum_s part[b][s][m][n]  (doubles; a complex
// matrix is 2N doubles per row).
static __global__ void __launch_bounds__(256) splitk_reduce_kernel(const double* __restrict__ part, double* __restrict__ C,
                                                            int M, int ND, int64_t ldcD, int64_t sCD, int SK,
                                                            int accumulate) {
    const int64_t per = (int64_t)M * ND;
    const int b = blockIdx.y;
    const double* pb = part + (int64_t)b * SK * per;
    double* cb = C + (int64_t)b * sCD;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < per;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = idx / ND;
        const int64_t nd = idx - m * ND;
        double acc = 0.0;
        for (int sidx = 0; sidx < SK; sidx++) acc += pb[(int64_t)sidx * per + idx];
        double* dst = cb + m * ldcD + nd;
        if (accumulate) acc += *dst;
        *dst = acc;
    }
}

// ---- host side -----------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// Tensor map over a k-contiguous operand X[mn][k] (leading dimension ld, batch stride sx, all in
// elements): dims (fastest first) = {K*E doubles, MN rows, batches}; box = {4*E doubles, MN_T rows, 1}.
template <bool CPLX>
static bool make_kc_map(CUtensorMap* map, const double* ptr, int MN, int K, int64_t ld, int64_t sx, int batch,
                        int mn_tile) {
    constexpr int E = CPLX ? 2 : 1;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    const bool batched = (sx != 0) && batch > 1;
    cuuint64_t dims[3] = {(cuuint64_t)K * E, (cuuint64_t)MN, (cuuint64_t)(batched ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * E * 8, (cuuint64_t)(batched ? sx * E * 8 : (int64_t)ld * E * 8 * MN)};
    if (strides[0] % 16 != 0 || strides[1] % 16 != 0) return false;
    if (strides[0] >= (1ULL << 40) || strides[1] >= (1ULL << 40)) return false;
    if (strides[1] == 0) strides[1] = 16;
    cuuint32_t box[3] = {(cuuint32_t)(4 * E), (cuuint32_t)mn_tile, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <bool CPLX, bool A_KC, bool B_KC, bool CONJB, bool SEG = false>
static int launch_ws_inst(const WsParams& p, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
    using Cfg = WsCfg<CPLX>;
    auto kern = gemm_ws_kernel<CPLX, A_KC, B_KC, CONJB, SEG>;
    static DeviceFlags configured;  // per instantiation and device
    PTB_TRY(ensure_dynamic_smem(configured, kern, Cfg::SMEM_BYTES));
    const int num_sms = device_sm_count();
    const long long total_tiles = (long long)p.tiles_m * p.tiles_n * p.batch;
    const long long total = p.split_k > 1 ? total_tiles * p.split_k
                          : (p.tail_split > 1 ? p.tail_begin + (total_tiles - p.tail_begin) * p.tail_split
                                              : total_tiles);
    const int grid = (int)(total < num_sms ? total : num_sms);
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(p, ta, tb);
    PTB_CUDA_TRY(cudaGetLastError());
    if (p.split_k == 1 && p.tail_split > 1) {
        tail_reduce_kernel<CPLX><<<(unsigned)(total_tiles - p.tail_begin), 256, 0, stream>>>(p);
        PTB_CUDA_TRY(cudaGetLastError());
    }
    if (p.split_k > 1) {
        const int ND = p.N * Cfg::E;
        const int64_t per = (int64_t)p.M * ND;
        int gx = (int)((per + 255) / 256);
        if (gx > num_sms * 8) gx = num_sms * 8;
        if (gx < 1) gx = 1;
        splitk_reduce_kernel<<<dim3(gx, p.batch), 256, 0, stream>>>(p.Cpart, p.C, p.M, ND, p.ldc * Cfg::E,
                                                                     p.sC * Cfg::E, p.split_k, p.accumulate);
        PTB_CUDA_TRY(cudaGetLastError());
    }
    return PTB_OK;
}

// Split factor for the persistent grid.  Splitting K by S raises the wave efficiency
// eff(S) = units / (waves * SMs) but costs S partial tiles written and read back (about 32 S bytes
// per output element of a complex GEMM at HBM speed).  Per output element the time saved is
// (8 K / peak) (1/eff(1) - 1/eff(S)); S is chosen to maximise the net gain and only accepted when
// the saving is at least twice the extra traffic.  At least 4 k-tiles (one pipeline depth) per unit.
static int choose_split_k(long long tiles, int KT, int K, int num_sms) {
    if (tiles >= 6LL * num_sms) return 1;
    // seconds per output element at the FP64 tensor peak: 128 flop / clk / SM (DMMA), ~1.96 GHz boost
    const double flop_time = 8.0 * (double)K / (128.0 * 1.96e9 * num_sms);
    const double byte_time = 32.0 / 6.0e12;               // write + read of one complex partial element
    auto eff = [&](int sk) {
        const long long units = tiles * sk;
        const long long waves = (units + num_sms - 1) / num_sms;
        return (double)units / (double)(waves * num_sms);
    };
    const double e1 = eff(1);
    double best_net = 0.0;
    int best = 1;
    for (int sk = 2; sk <= 8; sk++) {
        if (KT / sk < 4) break;
        const double saved = flop_time * (1.0 / e1 - 1.0 / eff(sk));
        const double cost = byte_time * sk;
        if (saved > 2.0 * cost && saved - cost > best_net) { best_net = saved - cost; best = sk; }
    }
    return best;
}

template <bool CPLX>
static int try_launch_ws(int transA, int transB, int conjB, const GemmParams& gp, cudaStream_t stream,
                         int split_k = 1, void* part_ws = nullptr,
                         size_t part_ws_bytes = 0, const int32_t* ktab = nullptr, const int32_t* order = nullptr) {
    using Cfg = WsCfg<CPLX>;
    constexpr int E = Cfg::E;
    const bool a_kc = (transA == 0), b_kc = (transB != 0);
    // 16-byte granularity of the bulk / tensor copies
    auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    if (!al16(gp.A) || !al16(gp.B)) return 1;
    if (CPLX && !al16(gp.C)) return 1;
    if (!CPLX) {
        if ((gp.lda | gp.ldb | gp.sA | gp.sB) & 1) return 1;
        if (!a_kc && (gp.M & 1)) return 1;  // partial mn-rows must stay multiples of 16 bytes
        if (!b_kc && (gp.N & 1)) return 1;
    }
    if (gp.K < 1) return 1;
    WsParams p;
    p.A = gp.A; p.B = gp.B; p.C = gp.C;
    p.M = gp.M; p.N = gp.N; p.K = gp.K;
    p.lda = gp.lda; p.ldb = gp.ldb; p.ldc = gp.ldc;
    p.sA = gp.sA; p.sB = gp.sB; p.sC = gp.sC;
    p.batch = gp.batch;
    p.accumulate = gp.accumulate;
    p.tiles_m = (gp.M + Cfg::BM - 1) / Cfg::BM;
    p.tiles_n = (gp.N + Cfg::BN - 1) / Cfg::BN;
    p.batched_a = (gp.sA != 0 && gp.batch > 1) ? 1 : 0;
    p.batched_b = (gp.sB != 0 && gp.batch > 1) ? 1 : 0;
    p.split_k = 1;
    p.Cpart = nullptr;
    p.tail_split = 1;
    p.tail_begin = 0;
    p.ktab = reinterpret_cast<const int2*>(ktab);
    p.seg_ptr = nullptr; p.segs = nullptr; p.sel_off = nullptr;
    p.order = order;
    p.group = ws_group_rows(p.tiles_m, gp.K, CPLX);
    if (ktab != nullptr && (reinterpret_cast<uintptr_t>(ktab) % 8) != 0) return PTB_ERR_ALIGNMENT;
    if (split_k != 1 && part_ws != nullptr && ktab == nullptr) {
        const int KT = (gp.K + Cfg::BK - 1) / Cfg::BK;
        const int num_sms = device_sm_count();
        int sk = split_k > 1 ? split_k : choose_split_k((long long)p.tiles_m * p.tiles_n * p.batch, KT, gp.K, num_sms);
        if (sk > KT) sk = KT;
        const size_t need = (size_t)p.batch * sk * p.M * p.N * E * 8;
        if (sk > 1 && need <= part_ws_bytes && al16(part_ws)) {
            p.split_k = sk;
            p.Cpart = static_cast<double*>(part_ws);
        }
        // no global split: split only the tiles of the last, partially filled wave
        if (p.split_k == 1 && split_k == 0 && al16(part_ws)) {
            const long long total_tiles = (long long)p.tiles_m * p.tiles_n * p.batch;
            const long long tail = total_tiles % num_sms;
            if (total_tiles > num_sms && tail > 0 && tail <= num_sms / 2) {
                int ts = (int)(num_sms / tail);
                if (ts > KT / 4) ts = KT / 4;
                if (ts > 16) ts = 16;
                const size_t need_tail = (size_t)tail * ts * Cfg::BM * Cfg::BN * E * 8;
                if (ts >= 2 && need_tail <= part_ws_bytes) {
                    p.tail_split = ts;
                    p.tail_begin = total_tiles - tail;
                    p.Cpart = static_cast<double*>(part_ws);
                }
            }
        }
    }
    CUtensorMap ta, tb;
    memset(&ta, 0, sizeof(ta));
    memset(&tb, 0, sizeof(tb));
    if (a_kc && !make_kc_map<CPLX>(&ta, gp.A, gp.M, gp.K, gp.lda, gp.sA, gp.batch, Cfg::BM)) return 1;
    if (b_kc && !make_kc_map<CPLX>(&tb, gp.B, gp.N, gp.K, gp.ldb, gp.sB, gp.batch, Cfg::BN)) return 1;
    (void)E;
    const bool cj = CPLX && conjB;
    const int sel = (a_kc ? 4 : 0) | (b_kc ? 2 : 0) | (cj ? 1 : 0);
    switch (sel) {
        case 0: return launch_ws_inst<CPLX, false, false, false>(p, ta, tb, stream);
        case 1: return launch_ws_inst<CPLX, false, false, CPLX>(p, ta, tb, stream);
        case 2: return launch_ws_inst<CPLX, false, true, false>(p, ta, tb, stream);
        case 3: return launch_ws_inst<CPLX, false, true, CPLX>(p, ta, tb, stream);
        case 4: return launch_ws_inst<CPLX, true, false, false>(p, ta, tb, stream);
        case 5: return launch_ws_inst<CPLX, true, false, CPLX>(p, ta, tb, stream);
        case 6: return launch_ws_inst<CPLX, true, true, false>(p, ta, tb, stream);
        case 7: return launch_ws_inst<CPLX, true, true, CPLX>(p, ta, tb, stream);
    }
    return 1;
}

// Segmented GEMM on mn-contiguous operands (A stored K x M, B stored K x N): see WsParams::seg_ptr.
template <bool CPLX>
static int launch_ws_segmented(int conjB, const GemmParams& gp, cudaStream_t stream, const int* seg_ptr,
                               const int* segs, const long long* sel_off, const int* order = nullptr) {
    using Cfg = WsCfg<CPLX>;
    auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    if (!al16(gp.A) || !al16(gp.B) || !al16(gp.C) || !al16(segs)) return PTB_ERR_ALIGNMENT;
    if (!CPLX && (((gp.lda | gp.ldb | gp.sA | gp.sB) & 1) || (gp.M & 1) || (gp.N & 1))) return PTB_ERR_ALIGNMENT;
    if (gp.K < 1) return PTB_ERR_BAD_ARG;
    WsParams p;
    p.A = gp.A; p.B = gp.B; p.C = gp.C;
    p.M = gp.M; p.N = gp.N; p.K = gp.K;
    p.lda = gp.lda; p.ldb = gp.ldb; p.ldc = gp.ldc;
    p.sA = gp.sA; p.sB = gp.sB; p.sC = gp.sC;
    p.batch = gp.batch;
    p.accumulate = gp.accumulate;
    p.tiles_m = (gp.M + Cfg::BM - 1) / Cfg::BM;
    p.tiles_n = (gp.N + Cfg::BN - 1) / Cfg::BN;
    p.batched_a = (gp.sA != 0 && gp.batch > 1) ? 1 : 0;
    p.batched_b = (gp.sB != 0 && gp.batch > 1) ? 1 : 0;
    p.split_k = 1; p.Cpart = nullptr; p.tail_split = 1; p.tail_begin = 0;
    p.ktab = nullptr;
    p.seg_ptr = seg_ptr;
    p.segs = reinterpret_cast<const int4*>(segs);
    p.sel_off = sel_off;
    p.order = order;
    p.group = 8;
    CUtensorMap ta, tb;
    memset(&ta, 0, sizeof(ta));
    memset(&tb, 0, sizeof(tb));
    if (CPLX && conjB) return launch_ws_inst<CPLX, false, false, CPLX, true>(p, ta, tb, stream);
    return launch_ws_inst<CPLX, false, false, false, true>(p, ta, tb, stream);
}

}  // namespace ptb
