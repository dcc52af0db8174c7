// Sparse W step of the effective-Hamiltonian contraction (HBM-bound kernel).
//
//   t_out[b, m, n] = sum_c W[m, c] t_in[b, c, n]        b < batch (left bond index i), n < N (right bra bond)
//
// MPO tensors of local Hamiltonians are tiny and 5-17 % dense (SURVEY.md headline 4: Heisenberg 12/100,
// Fermi-Hubbard 28/576 non-zeros), so for them the W step (pytenet/chain_ops.py:276) is pure data
// movement: read t1 once, write t2 once.  W is passed in CSR form (row pointers, column indices, values,
// a few hundred entries, served from L1/constant cache as warp-uniform broadcasts); every thread owns
// one column n of one batch and walks the CSR rows, so global loads and stores are fully coalesced
// 16-byte accesses along n.  The input tile of a CTA (R_in x 128 columns) stays in L1 across the rows.
// Algorithmic bytes: 16 (R_in + R_out) N batch.
#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

// TC: t is complex128 (else float64);  WC: W values complex
template <bool TC, bool WC>
__global__ void __launch_bounds__(128) wapply_csr_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                         const double* __restrict__ val,
                                                         const double* __restrict__ tin, double* __restrict__ tout,
                                                         int r_out, int r_in, int64_t n_cols, int batch0,
                                                         const unsigned char* __restrict__ active, int batch_block) {
    constexpr int E = TC ? 2 : 1;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_cols) return;
    const int64_t b = (int64_t)blockIdx.y + batch0;
    // row-activity flags (sector path): per (batch block, 128-column block) one byte per input row and per output
    // row.  Structurally zero input rows are not read, structurally zero output rows are not written (they keep
    // the zeros the caller initialised once).  The flags are uniform over the thread block.
    const unsigned char* __restrict__ fl =
        active ? active + ((b / batch_block) * gridDim.x + blockIdx.x) * (int64_t)(r_in + r_out) : nullptr;
    const double* __restrict__ x = tin + (b * r_in * n_cols + n) * E;
    double* __restrict__ y = tout + (b * r_out * n_cols + n) * E;
    int p = __ldg(rowptr);
    for (int m = 0; m < r_out; m++) {
        const int pe = __ldg(rowptr + m + 1);
        if (fl != nullptr && !fl[r_in + m]) { p = pe; continue; }
        double re = 0.0, im = 0.0;
        for (; p < pe; p++) {
            const int c = __ldg(col + p);
            if (fl != nullptr && !fl[c]) continue;
            if (TC) {
                const double2 xv = *reinterpret_cast<const double2*>(x + (int64_t)c * n_cols * 2);
                if (WC) {
                    const double wr = __ldg(val + 2 * p), wi = __ldg(val + 2 * p + 1);
                    re += wr * xv.x - wi * xv.y;
                    im += wr * xv.y + wi * xv.x;
                } else {
                    const double wv = __ldg(val + p);
                    re += wv * xv.x;
                    im += wv * xv.y;
                }
            } else {
                re += __ldg(val + p) * x[(int64_t)c * n_cols];
            }
        }
        if (TC)
            *reinterpret_cast<double2*>(y + (int64_t)m * n_cols * 2) = make_double2(re, im);
        else
            y[(int64_t)m * n_cols] = re;
    }
}

}  // namespace

extern "C" {

static int wapply_launch(int t_dtype, int w_is_complex, int64_t r_out, int64_t r_in, int64_t n_cols,
                         const int32_t* rowptr, const int32_t* col, const void* val, const void* t_in, void* t_out,
                         int64_t batch, const unsigned char* active, int64_t batch_block, void* stream) {
    if (!rowptr || !col || !val || !t_in || !t_out) return PTB_ERR_BAD_ARG;
    if (r_out <= 0 || r_in <= 0 || n_cols <= 0 || batch <= 0 || r_out > 0x7fffffffLL || r_in > 0x7fffffffLL)
        return PTB_ERR_BAD_ARG;
    if (active != nullptr && (batch_block <= 0 || batch_block > 0x7fffffffLL)) return PTB_ERR_BAD_ARG;
    const bool tc = t_dtype == PTB_COMPLEX128;
    if (!tc && t_dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    if (!tc && w_is_complex) return PTB_ERR_BAD_DTYPE;   // complex W on a real tensor: promote the tensor first
    if (tc && ((reinterpret_cast<uintptr_t>(t_in) | reinterpret_cast<uintptr_t>(t_out)) % 16)) return PTB_ERR_ALIGNMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned gx = (unsigned)((n_cols + 127) / 128);
    const double* v = static_cast<const double*>(val);
    const double* x = static_cast<const double*>(t_in);
    double* y = static_cast<double*>(t_out);
    const int bb = active ? (int)batch_block : 1;
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
        const unsigned gy = (unsigned)((batch - b0) > 65535 ? 65535 : (batch - b0));
        dim3 grid(gx, gy, 1);
        if (tc && w_is_complex)
            wapply_csr_kernel<true, true><<<grid, 128, 0, st>>>(rowptr, col, v, x, y, (int)r_out, (int)r_in, n_cols, (int)b0, active, bb);
        else if (tc)
            wapply_csr_kernel<true, false><<<grid, 128, 0, st>>>(rowptr, col, v, x, y, (int)r_out, (int)r_in, n_cols, (int)b0, active, bb);
        else
            wapply_csr_kernel<false, false><<<grid, 128, 0, st>>>(rowptr, col, v, x, y, (int)r_out, (int)r_in, n_cols, (int)b0, active, bb);
    }
    return cuda_status(cudaGetLastError());
}

int ptb_wapply_csr(int t_dtype, int w_is_complex, int64_t r_out, int64_t r_in, int64_t n_cols, const int32_t* rowptr,
                   const int32_t* col, const void* val, const void* t_in, void* t_out, int64_t batch, void* stream) {
    return wapply_launch(t_dtype, w_is_complex, r_out, r_in, n_cols, rowptr, col, val, t_in, t_out, batch, nullptr, 1,
                         stream);
}

int ptb_wapply_csr_masked(int t_dtype, int w_is_complex, int64_t r_out, int64_t r_in, int64_t n_cols,
                          const int32_t* rowptr, const int32_t* col, const void* val, const void* t_in, void* t_out,
                          int64_t batch, const uint8_t* active, int64_t batch_block, void* stream) {
    if (!active) return PTB_ERR_BAD_ARG;
    return wapply_launch(t_dtype, w_is_complex, r_out, r_in, n_cols, rowptr, col, val, t_in, t_out, batch, active,
                         batch_block, stream);
}

}  // extern "C"
