// A whole Lanczos run on the local effective Hamiltonian, enqueued by ONE call.
//
// pytenet/krylov.py:12-57 drives the closures of tdvp.py:223-238 / dmrg.py:181-189: per iteration one
// apply_local_hamiltonian (or apply_local_bond_contraction) and the three-term orthogonalisation.  For
// small bond dimensions (README config, METTS, chain edges) every kernel takes a few microseconds and the
// path is bound by launch latency and by the host code between launches, so the numiter iterations are
// issued here back to back from C: no Python, no allocation, no host synchronisation in between.  The
// kernels are the same as those behind the per-step entry points, so results are identical to calling
// ptb_lanczos_start / ptb_apply_local_hamiltonian / ptb_lanczos_ortho_step in a loop.
#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

namespace {

inline size_t up16(size_t x) { return (x + 15) & ~size_t(15); }

// scal = [ |x|, alpha[0..k), beta[0..k-1) ]  (2k doubles)
template <typename Matvec>
int lanczos_run(int dtype, int64_t n, const void* x, int numiter, void* V, double* scal, void* scratch, char* wvec,
                void* stream, Matvec matvec) {
    const bool cplx = dtype == PTB_COMPLEX128;
    const size_t row = (size_t)n * (cplx ? 16 : 8);
    char* v = static_cast<char*>(V);
    double* nrm = scal;
    double* alpha = scal + 1;
    double* beta = alpha + numiter;
    int rc = cplx ? ptb_lanczos_start_z(n, x, v, nrm, scratch, stream) : ptb_lanczos_start_d(n, x, v, nrm, scratch, stream);
    if (rc) return rc;
    for (int j = 0; j < numiter; j++) {
        char* vj = v + (size_t)j * row;
        rc = matvec(vj, wvec);
        if (rc) return rc;
        if (j == numiter - 1) {      // closing matvec only contributes alpha (krylov.py:53-56)
            return cplx ? ptb_lanczos_alpha_z(n, wvec, vj, alpha + j, scratch, stream)
                        : ptb_lanczos_alpha_d(n, wvec, vj, alpha + j, scratch, stream);
        }
        const void* vjm1 = j > 0 ? vj - row : nullptr;
        const double* bprev = j > 0 ? beta + (j - 1) : nullptr;
        rc = cplx ? ptb_lanczos_ortho_step_z(n, wvec, vj, vjm1, bprev, alpha + j, beta + j, vj + row, scratch, stream)
                  : ptb_lanczos_ortho_step_d(n, wvec, vj, vjm1, bprev, alpha + j, beta + j, vj + row, scratch, stream);
        if (rc) return rc;
    }
    return PTB_OK;
}

}  // namespace

extern "C" {

size_t ptb_heff_lanczos_workspace_bytes(int dtype, int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r) {
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    return up16((size_t)Dl * d * Dr * es) +
           ptb_apply_local_hamiltonian_workspace_bytes(dtype, Dl, d, Dr, chi_l, chi_r, d, Dl, Dr);
}

int ptb_heff_lanczos(int dtype, const void* x, const void* w, int w_is_complex, const int32_t* w_rowptr,
                     const int32_t* w_col, const void* w_val, const void* l, const void* r, int64_t Dl, int64_t d,
                     int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter, void* V, double* scal, void* scratch,
                     void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !l || !r || !V || !scal || !scratch || numiter < 1) return PTB_ERR_BAD_ARG;
    if (!w && !(w_rowptr && w_col && w_val)) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    if (dtype == PTB_REAL64 && w_is_complex) return PTB_ERR_BAD_DTYPE;
    for (int64_t e : {Dl, d, Dr, chi_l, chi_r})
        if (e <= 0 || e > 0x7fffffffLL) return PTB_ERR_BAD_ARG;
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    const size_t nvec = up16((size_t)Dl * d * Dr * es);
    if (!workspace || workspace_bytes < ptb_heff_lanczos_workspace_bytes(dtype, Dl, d, Dr, chi_l, chi_r))
        return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PTB_ERR_ALIGNMENT;
    char* wvec = static_cast<char*>(workspace);
    char* mws = wvec + nvec;
    const size_t mws_bytes = workspace_bytes - nvec;
    const bool cplx = dtype == PTB_COMPLEX128;
    // small bond dimensions: the dense-w entry dispatches to the fused one-kernel matvec (csrc/heff_small.cu)
    const bool small = w != nullptr &&
                       heff_small_applicable(cplx, true, w_is_complex != 0, Dl, d, Dr, chi_l, chi_r, d, Dl, Dr);
    const bool csr = !small && w_rowptr && w_col && w_val;
    auto matvec = [&](const void* vin, void* vout) -> int {
        if (csr) {
            return cplx ? ptb_apply_local_hamiltonian_csr_z(vin, w_rowptr, w_col, w_val, w_is_complex, l, r, vout, Dl, d,
                                                            Dr, chi_l, chi_r, d, Dl, Dr, mws, mws_bytes, stream)
                        : ptb_apply_local_hamiltonian_csr_d(vin, w_rowptr, w_col, w_val, l, r, vout, Dl, d, Dr, chi_l,
                                                            chi_r, d, Dl, Dr, mws, mws_bytes, stream);
        }
        return cplx ? ptb_apply_local_hamiltonian_z(vin, w, w_is_complex, l, r, vout, Dl, d, Dr, chi_l, chi_r, d, Dl, Dr,
                                                    mws, mws_bytes, stream)
                    : ptb_apply_local_hamiltonian_d(vin, w, l, r, vout, Dl, d, Dr, chi_l, chi_r, d, Dl, Dr, mws,
                                                    mws_bytes, stream);
    };
    return lanczos_run(dtype, Dl * d * Dr, x, numiter, V, scal, scratch, wvec, stream, matvec);
}

size_t ptb_bond_lanczos_workspace_bytes(int dtype, int64_t Dl, int64_t Dr, int64_t chi) {
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    return up16((size_t)Dl * Dr * es) + ptb_apply_local_bond_contraction_workspace_bytes(dtype, Dl, Dr, chi, Dl, Dr);
}

int ptb_bond_lanczos(int dtype, const void* c, const void* l, const void* r, int64_t Dl, int64_t Dr, int64_t chi,
                     int numiter, void* V, double* scal, void* scratch, void* workspace, size_t workspace_bytes,
                     void* stream) {
    if (!c || !l || !r || !V || !scal || !scratch || numiter < 1) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    for (int64_t e : {Dl, Dr, chi})
        if (e <= 0 || e > 0x7fffffffLL) return PTB_ERR_BAD_ARG;
    const size_t es = dtype == PTB_COMPLEX128 ? 16 : 8;
    const size_t nvec = up16((size_t)Dl * Dr * es);
    if (!workspace || workspace_bytes < ptb_bond_lanczos_workspace_bytes(dtype, Dl, Dr, chi)) return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PTB_ERR_ALIGNMENT;
    char* wvec = static_cast<char*>(workspace);
    char* mws = wvec + nvec;
    const size_t mws_bytes = workspace_bytes - nvec;
    const bool cplx = dtype == PTB_COMPLEX128;
    auto matvec = [&](const void* vin, void* vout) -> int {
        return cplx ? ptb_apply_local_bond_contraction_z(vin, l, r, vout, Dl, Dr, chi, Dl, Dr, mws, mws_bytes, stream)
                    : ptb_apply_local_bond_contraction_d(vin, l, r, vout, Dl, Dr, chi, Dl, Dr, mws, mws_bytes, stream);
    };
    return lanczos_run(dtype, Dl * Dr, c, numiter, V, scal, scratch, wvec, stream, matvec);
}

}  // extern "C"
