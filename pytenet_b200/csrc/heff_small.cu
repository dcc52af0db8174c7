// Fused effective-Hamiltonian matvec for small bond dimensions (launch-latency regime: README config, METTS,
// chain edges).  pytenet/chain_ops.py:237-279 (and :282-317 for the zero-site form).
//
// At D <~ 64 the three GEMM-shaped steps of  out = l . W . a . r  take a few microseconds each and the path is bound
// by kernel launches.  Here ONE launch does the whole contraction: CTA j' owns output column j' (the right bra bond
// index), for which no intermediate has to be shared between CTAs:
//   phase 1   t1[i,s,K]   = sum_j    a[i,s,j] r[j,K,j']             (shared memory)
//   phase 2   t2[i,k,s']  = sum_sK   w[k,s',s,K] t1[i,s,K]          (shared memory; identity when w == nullptr)
//   phase 3   out[i',s',j'] = sum_ik l[i,k,i'] t2[i,k,s']           (coalesced reads of l along i')
// Plain FP64 FMAs (the tiles are far too small for the tensor pipe); a and l are re-read by every CTA from L2.
#include <cstdlib>

#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

constexpr int HS_THREADS = 256;
constexpr int HS_MAX_DOUT = 16;

template <bool CPLX>
__device__ __forceinline__ void ldn(const double* p, int64_t idx, double& re, double& im) {
    if (CPLX) { re = p[2 * idx]; im = p[2 * idx + 1]; }
    else { re = p[idx]; im = 0.0; }
}

// WC: w complex.  w == nullptr: identity in (k, K), d_out = d_in = 1 (zero-site contraction; chi_l == chi_r)
template <bool CPLX, bool WC>
__global__ void __launch_bounds__(HS_THREADS) heff_small_kernel(const double* __restrict__ a, const double* __restrict__ w,
                                                                const double* __restrict__ l,
                                                                const double* __restrict__ r, double* __restrict__ out,
                                                                int Dl, int d, int Dr, int cl, int cr, int dout, int Dlp,
                                                                int Drp) {
    constexpr int E = CPLX ? 2 : 1;
    extern __shared__ double sm[];
    const int jp = blockIdx.x;
    const int tid = threadIdx.x;
    const int n1 = Dl * d * cr, n2 = Dl * cl * dout;
    const int G = HS_THREADS / Dlp > 0 ? HS_THREADS / Dlp : 1;
    double* rs = sm;                         // r[:, :, jp]      Dr * cr
    double* t1 = rs + (size_t)Dr * cr * E;   // n1
    double* t2 = t1 + (size_t)n1 * E;        // n2
    double* red = t2 + (size_t)n2 * E;       // G * dout * Dlp
    double* ws = red + (size_t)G * dout * Dlp * E;   // W as (cl*dout) x (d*cr), complex iff WC

    for (int idx = tid; idx < Dr * cr; idx += HS_THREADS) {
        double re, im;
        ldn<CPLX>(r, (int64_t)idx * Drp + jp, re, im);
        rs[E * idx] = re;
        if (CPLX) rs[E * idx + 1] = im;
    }
    if (w != nullptr) {
        const int nw = cl * dout * d * cr * (WC ? 2 : 1);
        for (int idx = tid; idx < nw; idx += HS_THREADS) ws[idx] = w[idx];
    }
    __syncthreads();

    // phase 1: t1[(i,s), K] = sum_j a[(i,s), j] rs[j, K]
    for (int o = tid; o < n1; o += HS_THREADS) {
        const int row = o / cr, K = o - row * cr;
        const double* arow = a + (int64_t)row * Dr * E;
        double re = 0.0, im = 0.0;
        for (int j = 0; j < Dr; j++) {
            if (CPLX) {
                const double ar = arow[2 * j], ai = arow[2 * j + 1];
                const double br = rs[2 * (j * cr + K)], bi = rs[2 * (j * cr + K) + 1];
                re += ar * br - ai * bi;
                im += ar * bi + ai * br;
            } else {
                re += arow[j] * rs[j * cr + K];
            }
        }
        t1[E * o] = re;
        if (CPLX) t1[E * o + 1] = im;
    }
    __syncthreads();

    // phase 2: t2[i, (k,s')] = sum_c W[(k,s'), c] t1[i, c],  c = (s, K)
    if (w != nullptr) {
        const int rin = d * cr, rout = cl * dout;
        for (int o = tid; o < n2; o += HS_THREADS) {
            const int i = o / rout, mrow = o - i * rout;
            double re = 0.0, im = 0.0;
            for (int c = 0; c < rin; c++) {
                double xr = t1[E * (i * rin + c)], xi = CPLX ? t1[E * (i * rin + c) + 1] : 0.0;
                if (WC) {
                    const double wr = ws[2 * (mrow * rin + c)], wi = ws[2 * (mrow * rin + c) + 1];
                    re += wr * xr - wi * xi;
                    im += wr * xi + wi * xr;
                } else {
                    const double wv = ws[mrow * rin + c];
                    re += wv * xr;
                    im += wv * xi;
                }
            }
            t2[E * o] = re;
            if (CPLX) t2[E * o + 1] = im;
        }
    } else {
        for (int o = tid; o < n2 * E; o += HS_THREADS) t2[o] = t1[o];       // identity: n1 == n2
    }
    __syncthreads();

    // phase 3: out[i', s', jp] = sum_p l[p, i'] t2[p, s'],  p = (i, k); thread = (group g, i'), groups split p
    const int ip = tid % Dlp, g = tid / Dlp;
    double accr[HS_MAX_DOUT], acci[HS_MAX_DOUT];
#pragma unroll
    for (int s = 0; s < HS_MAX_DOUT; s++) { accr[s] = 0.0; acci[s] = 0.0; }
    if (g < G) {
        const int np = Dl * cl;
        for (int p = g; p < np; p += G) {
            double lr, li;
            ldn<CPLX>(l, (int64_t)p * Dlp + ip, lr, li);
#pragma unroll
            for (int s = 0; s < HS_MAX_DOUT; s++) {
                if (s < dout) {
                    const double xr = t2[E * (p * dout + s)], xi = CPLX ? t2[E * (p * dout + s) + 1] : 0.0;
                    accr[s] += lr * xr - li * xi;
                    if (CPLX) acci[s] += lr * xi + li * xr;
                }
            }
        }
#pragma unroll
        for (int s = 0; s < HS_MAX_DOUT; s++) {
            if (s < dout) {
                red[E * ((g * dout + s) * Dlp + ip)] = accr[s];
                if (CPLX) red[E * ((g * dout + s) * Dlp + ip) + 1] = acci[s];
            }
        }
    }
    __syncthreads();
    for (int o = tid; o < dout * Dlp; o += HS_THREADS) {
        const int s = o / Dlp, i2 = o - s * Dlp;
        double re = 0.0, im = 0.0;
        for (int gg = 0; gg < G; gg++) {
            re += red[E * ((gg * dout + s) * Dlp + i2)];
            if (CPLX) im += red[E * ((gg * dout + s) * Dlp + i2) + 1];
        }
        double* dst = out + (((int64_t)i2 * dout + s) * Drp + jp) * E;
        dst[0] = re;
        if (CPLX) dst[1] = im;
    }
}

bool small_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = std::getenv("PYTENET_B200_SMALL_HEFF");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

}  // namespace

namespace ptb {

static size_t heff_small_smem(bool cplx, bool has_w, bool w_cplx, int64_t Dl, int64_t d, int64_t Dr, int64_t cl,
                              int64_t cr, int64_t dout, int64_t Dlp) {
    const size_t es = cplx ? 16 : 8;
    const int64_t n1 = Dl * d * cr, n2 = Dl * cl * dout;
    const int64_t G = HS_THREADS / Dlp > 0 ? HS_THREADS / Dlp : 1;
    return ((size_t)(Dr * cr + n1 + n2 + G * dout * Dlp)) * es +
           (has_w ? (size_t)(cl * dout * d * cr) * (w_cplx ? 16 : 8) : 0);
}

// True when the fused small-D kernel applies (and is expected to beat the three-GEMM path).
bool heff_small_applicable(bool cplx, bool has_w, bool w_cplx, int64_t Dl, int64_t d, int64_t Dr, int64_t cl,
                           int64_t cr, int64_t dout, int64_t Dlp, int64_t Drp) {
    if (!small_enabled()) return false;
    if (Dlp > HS_THREADS || dout > HS_MAX_DOUT || Dl > 256 || Dr > 256 || Drp > 256 || Dlp < 1) return false;
    if (!has_w && (cl != cr || d != 1 || dout != 1)) return false;
    if (!cplx && w_cplx) return false;
    if (heff_small_smem(cplx, has_w, w_cplx, Dl, d, Dr, cl, cr, dout, Dlp) > 160 * 1024) return false;
    // work per CTA (multiply-adds of the three phases): beyond this the tensor-pipe GEMMs win
    const int64_t n1 = Dl * d * cr, n2 = Dl * cl * dout;
    const int64_t macs = n1 * Dr + (has_w ? n2 * (d * cr) : 0) + n2 * Dlp;
    return macs <= 150000;
}

int heff_small_launch(bool cplx, const void* a, const void* w, bool w_cplx, const void* l, const void* r, void* out,
                      int64_t Dl, int64_t d, int64_t Dr, int64_t cl, int64_t cr, int64_t dout, int64_t Dlp, int64_t Drp,
                      cudaStream_t st) {
    const size_t smem = heff_small_smem(cplx, w != nullptr, w_cplx, Dl, d, Dr, cl, cr, dout, Dlp);
    auto launch = [&](auto kern, int slot) -> int {
        static DeviceFlags configured[4];
        PTB_TRY(ensure_dynamic_smem(configured[slot], kern, 160 * 1024));
        kern<<<(unsigned)Drp, HS_THREADS, smem, st>>>(static_cast<const double*>(a), static_cast<const double*>(w),
                                                      static_cast<const double*>(l), static_cast<const double*>(r),
                                                      static_cast<double*>(out), (int)Dl, (int)d, (int)Dr, (int)cl,
                                                      (int)cr, (int)dout, (int)Dlp, (int)Drp);
        return cuda_status(cudaGetLastError());
    };
    if (cplx && w_cplx) return launch(heff_small_kernel<true, true>, 3);
    if (cplx) return launch(heff_small_kernel<true, false>, 2);
    return launch(heff_small_kernel<false, false>, 0);
}

}  // namespace ptb
