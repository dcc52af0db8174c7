// MPO-bond-sharded entry points of the C ABI (SURVEY.md section 8(b) minimum list: `*_comm_init / _destroy` and
// sharded variants of the four contractions; section 8(e): BASELINE config 4, molecular Hamiltonians).
//
//   H_eff = sum_{k,kappa} l_k^T (x) w[k,:,:,kappa] (x) r_kappa      is a sum over MPO-bond index pairs.
//
// Rank g of G owns the range kappa_g of the RIGHT MPO bond (zero padded to P = ceil(chi_r / G)) and, once per
// site, contracts the full left block with its slice of the MPO tensor (ptb_sharded_precontract):
//   LW_g[(i, s, kappa_loc), s', i'] = sum_k w[k, s', s, kappa] l[i, k, i'].
// One matvec is then two GEMMs on the rank's range and ONE all-reduce of the result:
//   t1_g = a . r_g,   out_g = LW_g^T . t1_g,   out = sum_g out_g            (ptb_apply_local_hamiltonian_sharded)
// the next left block is born sharded over the right bond with no communication:
//   l_next,g[j, kappa_loc, j'] = a^T ( LW_g conj(b) )                       (ptb_env_step_left_sharded)
// and the zero-site contraction sums over the shard of the shared bond:
//   out = sum_g l_g^T (c . r_g)                                            (ptb_apply_local_bond_contraction_sharded)
// The right-to-left direction uses the same entries on mirrored tensors (a -> a^T(2,1,0), w -> w^T(3,1,2,0)).
//
// The communicator is an opaque handle around an NCCL communicator (one process per GPU; NCCL over NVLink 5 /
// NVSwitch).  libnccl is resolved at run time (the copy already loaded by the host framework, else the system
// one), so a host without torch can drive the multi-GPU path through this ABI alone.  `comm == NULL` means a
// single rank: the all-reduce is skipped.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "../../include/pytenet_b200.h"
#include "common.cuh"

using namespace ptb;

struct ptb_comm {
    ncclComm_t nccl;
    int nranks, rank, device;
};

namespace {

struct NcclApi {
    decltype(&ncclGetUniqueId) get_id = nullptr;
    decltype(&ncclCommInitRank) init_rank = nullptr;
    decltype(&ncclAllReduce) all_reduce = nullptr;
    decltype(&ncclCommDestroy) destroy = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_nccl.get_id = reinterpret_cast<decltype(g_nccl.get_id)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.init_rank = reinterpret_cast<decltype(g_nccl.init_rank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.all_reduce = reinterpret_cast<decltype(g_nccl.all_reduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.destroy = reinterpret_cast<decltype(g_nccl.destroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.ok = g_nccl.get_id && g_nccl.init_rank && g_nccl.all_reduce && g_nccl.destroy;
}

inline bool nccl_ready() {
    std::call_once(g_nccl_once, load_nccl);
    return g_nccl.ok;
}

inline size_t up16(size_t x) { return (x + 15) & ~size_t(15); }
inline size_t esize(int dtype) { return dtype == PTB_COMPLEX128 ? 16 : 8; }

// in-place sum over the ranks of `count` elements (complex = 2 doubles); no-op for a single rank
int allreduce(ptb_comm* comm, int dtype, void* buf, int64_t count, cudaStream_t st) {
    if (!comm || comm->nranks <= 1) return PTB_OK;
    const size_t n = (size_t)count * (dtype == PTB_COMPLEX128 ? 2 : 1);
    const ncclResult_t r = g_nccl.all_reduce(buf, buf, n, ncclDouble, ncclSum, comm->nccl, st);
    return r == ncclSuccess ? PTB_OK : PTB_ERR_NOT_INITIALISED;
}

constexpr int MAX_SPLIT = 8;

}  // namespace

extern "C" {

int ptb_comm_unique_id(void* id128) {
    if (!id128) return PTB_ERR_BAD_ARG;
    if (!nccl_ready()) return PTB_ERR_NOT_INITIALISED;
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    return g_nccl.get_id(static_cast<ncclUniqueId*>(id128)) == ncclSuccess ? PTB_OK : PTB_ERR_NOT_INITIALISED;
}

int ptb_comm_init(ptb_comm** comm, int nranks, int rank, const void* id128) {
    if (!comm || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return PTB_ERR_BAD_ARG;
    if (!nccl_ready()) return PTB_ERR_NOT_INITIALISED;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ptb_comm* c = new ptb_comm();
    c->nranks = nranks;
    c->rank = rank;
    c->device = current_device();
    if (g_nccl.init_rank(&c->nccl, nranks, id, rank) != ncclSuccess) {
        delete c;
        return PTB_ERR_NOT_INITIALISED;
    }
    *comm = c;
    return PTB_OK;
}

int ptb_comm_destroy(ptb_comm* comm) {
    if (!comm) return PTB_OK;
    const ncclResult_t r = g_nccl.ok ? g_nccl.destroy(comm->nccl) : ncclSuccess;
    delete comm;
    return r == ncclSuccess ? PTB_OK : PTB_ERR_NOT_INITIALISED;
}

int ptb_comm_info(const ptb_comm* comm, int* nranks, int* rank) {
    if (nranks) *nranks = comm ? comm->nranks : 1;
    if (rank) *rank = comm ? comm->rank : 0;
    return PTB_OK;
}

int ptb_allreduce_sum(ptb_comm* comm, int dtype, void* buf, int64_t count, void* stream) {
    if (!buf || count < 0) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    return allreduce(comm, dtype, buf, count, static_cast<cudaStream_t>(stream));
}

// LW[i, r, i'] = sum_k w3[r, k] l[i, k, i']   (r = (s, kappa_loc, s') runs over R rows); w3 real or complex
int ptb_sharded_precontract(int dtype, int w_is_complex, const void* w3, const void* l, void* lw, int64_t Dl,
                            int64_t chi_l, int64_t Dlp, int64_t R, void* stream) {
    if (!w3 || !l || !lw) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    if (dtype == PTB_REAL64 && w_is_complex) return PTB_ERR_BAD_DTYPE;
    for (int64_t x : {Dl, chi_l, Dlp, R})
        if (x <= 0 || x > 0x7fffffffLL) return PTB_ERR_BAD_ARG;
    if (dtype == PTB_COMPLEX128 && !w_is_complex)
        // real W on complex l: real GEMM over (re, im)-interleaved columns, half the flops of a zgemm
        return ptb_gemm(PTB_REAL64, 0, 0, 0, R, 2 * Dlp, chi_l, w3, chi_l, l, 2 * Dlp, lw, 2 * Dlp, Dl, 0,
                        2 * chi_l * Dlp, 2 * R * Dlp, 0, stream);
    return ptb_gemm(dtype, 0, 0, 0, R, Dlp, chi_l, w3, chi_l, l, Dlp, lw, Dlp, Dl, 0, chi_l * Dlp, R * Dlp, 0, stream);
}

size_t ptb_apply_local_hamiltonian_sharded_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t Dr, int64_t P,
                                                           int64_t d_out, int64_t Dlp, int64_t Drp) {
    (void)Dr;
    const size_t es = esize(dtype);
    return up16((size_t)Dl * d_in * P * Drp * es) + up16((size_t)MAX_SPLIT * d_out * Dlp * Drp * es);
}

// out[i', s', j'] = sum over ranks of  sum_{(i,s,kappa_loc)} lw[(i,s,kappa_loc), s', i'] (a r_shard)[(i,s), (kappa_loc, j')]
int ptb_apply_local_hamiltonian_sharded(ptb_comm* comm, int dtype, const void* a, const void* lw, const void* r_shard,
                                        void* out, int64_t Dl, int64_t d_in, int64_t Dr, int64_t P, int64_t d_out,
                                        int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                                        void* stream) {
    if (!a || !lw || !r_shard || !out) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    for (int64_t x : {Dl, d_in, Dr, P, d_out, Dlp, Drp})
        if (x <= 0 || x > 0x7fffffffLL) return PTB_ERR_BAD_ARG;
    if (!workspace ||
        workspace_bytes < ptb_apply_local_hamiltonian_sharded_workspace_bytes(dtype, Dl, d_in, Dr, P, d_out, Dlp, Drp))
        return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PTB_ERR_ALIGNMENT;
    const size_t es = esize(dtype);
    char* t1 = static_cast<char*>(workspace);
    char* part = t1 + up16((size_t)Dl * d_in * P * Drp * es);
    const size_t part_bytes = workspace_bytes - (size_t)(part - t1);
    // step 1 on this rank's kappa range; its row-major memory is also [(i, s, kappa_loc), j']
    PTB_TRY(ptb_gemm(dtype, 0, 0, 0, Dl * d_in, P * Drp, Dr, a, Dr, r_shard, P * Drp, t1, P * Drp, 1, 0, 0, 0, 0, stream));
    // one GEMM batched over s' with the long contraction index split over work units (the output has few tiles)
    PTB_TRY(ptb_gemm_splitk(dtype, 1, 0, 0, Dlp, Drp, Dl * d_in * P, lw, d_out * Dlp, t1, Drp, out, d_out * Drp, d_out,
                            Dlp, 0, Drp, 0, 0, part, part_bytes, stream));
    return allreduce(comm, dtype, out, Dlp * d_out * Drp, static_cast<cudaStream_t>(stream));
}

size_t ptb_env_step_left_sharded_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t P, int64_t Drb) {
    return up16((size_t)Dl * d_in * P * Drb * esize(dtype));
}

// l_next[j, kappa_loc, j'] = sum_{i,s} a[i,s,j] X[(i,s,kappa_loc), j'],   X = sum_{s'} lw[:, s', :] conj(b[:, s', :])
int ptb_env_step_left_sharded(int dtype, const void* a, const void* b, const void* lw, void* l_next, int64_t Dl,
                              int64_t d_in, int64_t Dr, int64_t P, int64_t d_out, int64_t Dlp, int64_t Drb,
                              void* workspace, size_t workspace_bytes, void* stream) {
    if (!a || !b || !lw || !l_next) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    for (int64_t x : {Dl, d_in, Dr, P, d_out, Dlp, Drb})
        if (x <= 0 || x > 0x7fffffffLL) return PTB_ERR_BAD_ARG;
    if (!workspace || workspace_bytes < ptb_env_step_left_sharded_workspace_bytes(dtype, Dl, d_in, P, Drb))
        return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PTB_ERR_ALIGNMENT;
    const size_t es = esize(dtype);
    const char* lwc = static_cast<const char*>(lw);
    const char* bc = static_cast<const char*>(b);
    const int64_t rows = Dl * d_in * P;
    // X (rows x Drb): one accumulating GEMM per s' (no permuted copy of b: the slice b[:, s', :] has row stride d_out*Drb)
    for (int64_t sp = 0; sp < d_out; sp++)
        PTB_TRY(ptb_gemm(dtype, 0, 0, 1, rows, Drb, Dlp, lwc + (size_t)sp * Dlp * es, d_out * Dlp,
                         bc + (size_t)sp * Drb * es, d_out * Drb, workspace, Drb, 1, 0, 0, 0, sp > 0 ? 1 : 0, stream));
    // l_next (Dr x P*Drb) = a^T X  with X viewed as (Dl*d) x (P*Drb)
    return ptb_gemm(dtype, 1, 0, 0, Dr, P * Drb, Dl * d_in, a, Dr, workspace, P * Drb, l_next, P * Drb, 1, 0, 0, 0, 0,
                    stream);
}

size_t ptb_apply_local_bond_contraction_sharded_workspace_bytes(int dtype, int64_t Dl, int64_t P, int64_t Drp) {
    return up16((size_t)Dl * P * Drp * esize(dtype));
}

// out[i', j'] = sum over ranks of  sum_{i,k_loc} l_shard[i, k_loc, i'] (c r_shard)[i, (k_loc, j')]
int ptb_apply_local_bond_contraction_sharded(ptb_comm* comm, int dtype, const void* c, const void* l_shard,
                                             const void* r_shard, void* out, int64_t Dl, int64_t Dr, int64_t P,
                                             int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    if (!c || !l_shard || !r_shard || !out) return PTB_ERR_BAD_ARG;
    if (dtype != PTB_COMPLEX128 && dtype != PTB_REAL64) return PTB_ERR_BAD_DTYPE;
    for (int64_t x : {Dl, Dr, P, Dlp, Drp})
        if (x <= 0 || x > 0x7fffffffLL) return PTB_ERR_BAD_ARG;
    if (!workspace || workspace_bytes < ptb_apply_local_bond_contraction_sharded_workspace_bytes(dtype, Dl, P, Drp))
        return PTB_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PTB_ERR_ALIGNMENT;
    PTB_TRY(ptb_gemm(dtype, 0, 0, 0, Dl, P * Drp, Dr, c, Dr, r_shard, P * Drp, workspace, P * Drp, 1, 0, 0, 0, 0, stream));
    PTB_TRY(ptb_gemm(dtype, 1, 0, 0, Dlp, Drp, Dl * P, l_shard, Dlp, workspace, Drp, out, Drp, 1, 0, 0, 0, 0, stream));
    return allreduce(comm, dtype, out, Dlp * Drp, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
