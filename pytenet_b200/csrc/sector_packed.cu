// C-ABI entries of the sector-packed block-sparse path (see include/pytenet_b200.h and
// pytenet_b200/sector_packed.py): the grouped GEMM (gemm_grouped.cuh) and the block gather that serves the W
// step, the packing of the operands and the unpacking of the result.
#include "../../include/pytenet_b200.h"
#include "gemm_grouped.cuh"

using namespace ptb;

namespace {

static_assert(sizeof(ptb_group_tile) == sizeof(GroupTile), "tile descriptor layouts must agree");
static_assert(sizeof(ptb_gather_chunk) == 32 && sizeof(ptb_gather_term) == 32, "gather table layouts");

constexpr int GATHER_THREADS = 256;

// dst[r, c] = sum_t coef_t * src[off_t + r * rs_t + c * cs_t]   for the rows of one work item.
// HBM / L2 bound: every destination element is written once; the terms of a chunk (a handful: the non-zero
// MPO entries that connect two sector blocks) are warp-uniform broadcast loads.
template <bool CPLX>
__global__ void __launch_bounds__(GATHER_THREADS) block_gather_kernel(const double* __restrict__ src,
                                                                      double* __restrict__ dst,
                                                                      const ptb_gather_chunk* __restrict__ chunks,
                                                                      const ptb_gather_term* __restrict__ terms,
                                                                      const int4* __restrict__ work) {
    const int4 wk = work[blockIdx.x];
    const ptb_gather_chunk ch = chunks[wk.x];
    const int row0 = wk.y, nrows = wk.z;
    const int total = nrows * ch.cols;
    const double sgn = (ch.flags & 1) ? -1.0 : 1.0;        // conjugated source (bra tensors)
    for (int idx = threadIdx.x; idx < total; idx += GATHER_THREADS) {
        const int r = row0 + idx / ch.cols;
        const int c = idx % ch.cols;
        double re = 0.0, im = 0.0;
        for (int t = ch.term_begin; t < ch.term_end; t++) {
            const ptb_gather_term tm = terms[t];
            const int64_t s = tm.src_off + (int64_t)r * tm.src_rs + (int64_t)c * tm.src_cs;
            if (CPLX) {
                double2 v = *reinterpret_cast<const double2*>(src + 2 * s);
                v.y *= sgn;
                re += tm.coef_re * v.x - tm.coef_im * v.y;
                im += tm.coef_re * v.y + tm.coef_im * v.x;
            } else {
                re += tm.coef_re * src[s];
            }
        }
        const int64_t d = ch.dst_off + (int64_t)r * ch.dst_ld + c;
        if (CPLX) *reinterpret_cast<double2*>(dst + 2 * d) = make_double2(re, im);
        else dst[d] = re;
    }
}

}  // namespace

extern "C" {

int ptb_gemm_grouped_tile_shape(int dtype, int variant, int* bm, int* bn) {
    if (!bm || !bn || variant < 0 || variant > 1) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    const bool cplx = dtype == PTB_COMPLEX128;
    if (variant == 0) { *bm = cplx ? WsCfg<true>::BM : WsCfg<false>::BM; *bn = cplx ? WsCfg<true>::BN : WsCfg<false>::BN; }
    else { *bm = cplx ? GroupSmallCfg<true>::BM : GroupSmallCfg<false>::BM; *bn = cplx ? GroupSmallCfg<true>::BN : GroupSmallCfg<false>::BN; }
    return PTB_OK;
}

int ptb_gemm_grouped_v(int dtype, int variant, const void* a, const void* b, void* c, const ptb_group_tile* tiles,
                       int ntiles, void* stream) {
    if (ntiles < 0 || variant < 0 || variant > 1) return PTB_ERR_BAD_ARG;
    if (ntiles == 0) return PTB_OK;
    if (!a || !b || !c || !tiles) return PTB_ERR_BAD_ARG;
    auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    if (!al16(a) || !al16(b) || !al16(c) || !al16(tiles)) return PTB_ERR_ALIGNMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const GroupTile* gt = reinterpret_cast<const GroupTile*>(tiles);
    const double* ad = static_cast<const double*>(a);
    const double* bd = static_cast<const double*>(b);
    double* cd = static_cast<double*>(c);
    if (dtype == PTB_COMPLEX128)
        return variant == 0 ? launch_grouped<true, WsCfg<true>>(ad, bd, cd, gt, ntiles, st)
                            : launch_grouped<true, GroupSmallCfg<true>>(ad, bd, cd, gt, ntiles, st);
    if (dtype == PTB_REAL64)
        return variant == 0 ? launch_grouped<false, WsCfg<false>>(ad, bd, cd, gt, ntiles, st)
                            : launch_grouped<false, GroupSmallCfg<false>>(ad, bd, cd, gt, ntiles, st);
    return PTB_ERR_BAD_DTYPE;
}

int ptb_gemm_grouped(int dtype, const void* a, const void* b, void* c, const ptb_group_tile* tiles, int ntiles,
                     void* stream) {
    return ptb_gemm_grouped_v(dtype, 0, a, b, c, tiles, ntiles, stream);
}

int ptb_block_gather(int dtype, const void* src, void* dst, const ptb_gather_chunk* chunks,
                     const ptb_gather_term* terms, const int32_t* work, int nwork, void* stream) {
    if (nwork < 0) return PTB_ERR_BAD_ARG;
    if (nwork == 0) return PTB_OK;
    if (!src || !dst || !chunks || !terms || !work) return PTB_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(work) % 16 || reinterpret_cast<uintptr_t>(chunks) % 8 ||
        reinterpret_cast<uintptr_t>(terms) % 8)
        return PTB_ERR_ALIGNMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int4* wk = reinterpret_cast<const int4*>(work);
    if (dtype == PTB_COMPLEX128) {
        if (reinterpret_cast<uintptr_t>(src) % 16 || reinterpret_cast<uintptr_t>(dst) % 16) return PTB_ERR_ALIGNMENT;
        block_gather_kernel<true><<<nwork, GATHER_THREADS, 0, st>>>(static_cast<const double*>(src),
                                                                   static_cast<double*>(dst), chunks, terms, wk);
    } else if (dtype == PTB_REAL64) {
        block_gather_kernel<false><<<nwork, GATHER_THREADS, 0, st>>>(static_cast<const double*>(src),
                                                                    static_cast<double*>(dst), chunks, terms, wk);
    } else {
        return PTB_ERR_BAD_DTYPE;
    }
    return cuda_status(cudaGetLastError());
}

}  // extern "C"
