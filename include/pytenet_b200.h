/*
 * pytenet_b200 -- C ABI of the B200-native effective-Hamiltonian path.
 *
 * Drop-in boundary for the hot path of cmendl/pytenet v1.3.0.  The reference is
 * pure Python (no FFI of its own), so every entry point cites the *Python*
 * function it replaces; INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers (void*), 64-bit extents, no torch / C++ types.
 *   - all tensors are C-ordered (row-major) and dense, exactly as NumPy holds
 *     them: MPS tensor (Dl, d, Dr); MPO tensor (chi_l, d_out, d_in, chi_r);
 *     environment blocks (ket bond, MPO bond, bra bond).
 *   - suffix _d: every tensor float64.  suffix _z: a, b, c, l, r, out are
 *     complex128 (interleaved re, im, 16 bytes); the MPO tensor w is float64
 *     when w_is_complex == 0 (what pytenet/mpo.py:132 allocates) or complex128.
 *   - the library never allocates: the caller passes `out` and a `workspace` of
 *     at least *_workspace_bytes(...) bytes (device memory, 16-byte aligned).
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous with
 *     respect to the host and re-entrant across streams.
 *   - return value: 0 = ok; negative = PTB_ERR_* (bad argument); positive = a
 *     cudaError_t / ncclResult_t code.  No exceptions, no exit().
 *   - outputs never alias inputs.
 */
#ifndef PYTENET_B200_H
#define PYTENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_OK 0
#define PTB_ERR_BAD_ARG (-1)
#define PTB_ERR_BAD_DTYPE (-2)
#define PTB_ERR_WORKSPACE (-3)
#define PTB_ERR_ALIGNMENT (-4)
#define PTB_ERR_TOO_LARGE (-5)
#define PTB_ERR_NOT_INITIALISED (-6)

#define PTB_REAL64 0
#define PTB_COMPLEX128 1

/* library version (major*10000 + minor*100 + patch) and status text */
int ptb_version(void);
const char* ptb_status_string(int status);

/* ---------------------------------------------------------------------------
 * Building block: strided-batched GEMM on the FP64 tensor pipe (DMMA.8x8x4).
 *   C[b] (M x N, ldc) (+)= op(A[b]) * op(B[b]),  b = 0..batch-1
 *   trans_a = 0: A is M x K row-major (lda);  1: A is stored K x M row-major
 *   trans_b = 0: B is K x N row-major (ldb);  1: B is stored N x K row-major
 *   conj_b     : use conj(B) (complex only)
 * Leading dimensions and batch strides are in ELEMENTS of `dtype`.
 * Replaces the np.tensordot -> OpenBLAS zgemm/dgemm calls at
 * pytenet/chain_ops.py:50,52,56,94,96,98,273,276,278,314,316.
 * ------------------------------------------------------------------------- */
int ptb_gemm(int dtype, int trans_a, int trans_b, int conj_b,
             int64_t m, int64_t n, int64_t k,
             const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc,
             int64_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_c,
             int accumulate, void* stream);

/* ptb_gemm with split-K: when the output has too few tiles to fill the 148 SMs, every tile is
 * computed by `split_k` work units over disjoint k ranges (partials in `workspace`, at least
 * batch*split_k*m*n elements) and summed in fixed order by a second kernel (deterministic).
 * split_k = 0 chooses the factor automatically (1..8), 1 disables splitting. */
int ptb_gemm_splitk(int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k,
                    const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc,
                    int64_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_c, int accumulate,
                    int split_k, void* workspace, size_t workspace_bytes, void* stream);

/* Sector-banded GEMM (quantum-number block sparsity): like ptb_gemm, but every output tile only
 * visits the k-tiles in [ktab[2*t], ktab[2*t+1]) where t = (batch * tiles_m + tile_m) * tiles_n + tile_n
 * (device array of int32 pairs, tile shape from ptb_gemm_tile_shape).  The caller derives the ranges
 * from the quantum numbers of the operand indices: contributions outside them are exact zeros
 * (pytenet/block_sparse_util.py:47-53), so the result equals the dense product.  An empty range
 * leaves the tile zero (or untouched when accumulating).  accumulate == 2: overwrite like 0, but leave
 * tiles with an empty range untouched (the caller keeps the structurally empty part of C zeroed once
 * instead of having the zeros rewritten by every call).  TMA engine only. */
int ptb_gemm_banded(int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k,
                    const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc,
                    int64_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_c, int accumulate,
                    const int32_t* ktab, void* stream);
int ptb_gemm_tile_shape(int dtype, int* bm, int* bn, int* bk);
/* General form of the two sector GEMMs with an optional tile schedule.  Exactly one of `ktab` (banded mode, as
 * ptb_gemm_banded; accumulate 0 / 1 / 2) and `seg_ptr` + `segs` + `sel_off` (segmented mode, as ptb_gemm_segmented;
 * needs trans_a = 1, trans_b = 0) is given.  `order` (optional) is a permutation of the tile indices: work unit u of
 * the persistent grid processes tile order[u].  The k ranges of the tiles differ widely, so the caller sorts the
 * tiles by decreasing work, which balances the round-robin assignment of units to the 148 CTAs.  All tables are
 * device arrays. */
typedef struct {
    const int32_t* ktab;
    const int32_t* seg_ptr;
    const int32_t* segs;
    const int64_t* sel_off;
    const int32_t* order;
} ptb_sector_tables;
int ptb_gemm_sector(int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k, const void* a,
                    int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch, int64_t stride_a,
                    int64_t stride_b, int64_t stride_c, int accumulate, const ptb_sector_tables* tables, void* stream);

/* Segmented GEMM (sums over an outer index inside ONE launch):  C[b] (+)= sum_seg A_seg^T B_seg  with both
 * operands stored K-major (A is K x M, B is K x N, row-major: trans_a = 1, trans_b = 0 in ptb_gemm terms).
 * Output tile t (same numbering as ptb_gemm_banded) accumulates the segments
 * segs[seg_ptr[t] .. seg_ptr[t+1]); a segment is 4 int32 {first k-tile, end k-tile, selector, 0}, and
 * sel_off[2*selector], sel_off[2*selector+1] are ELEMENT offsets added to the a / b base pointers for that
 * segment.  Step 3 of the sector path,  out[i',s',j'] = sum_k sum_i l[i,k,i'] t2[i,k,s',j'],  uses one selector per
 * left MPO index k (offsets k*Dlp and k*d*Drp, lda = chi_l*Dlp, ldb = chi_l*d*Drp) and per tile only the k-tile
 * ranges the quantum numbers allow -- one launch instead of chi_l accumulating ones.  A tile without segments is
 * zero (or untouched when accumulating).  All tables are device arrays (segs 16-byte aligned).  TMA engine only. */
int ptb_gemm_segmented(int dtype, int conj_b, int64_t m, int64_t n, int64_t k, const void* a, int64_t lda,
                       const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch, int64_t stride_a,
                       int64_t stride_b, int64_t stride_c, int accumulate, const int32_t* seg_ptr, const int32_t* segs,
                       const int64_t* sel_off, void* stream);


/* ptb_gemm with an explicit kernel generation (for A/B measurements and tests; an argument, so the
 * library keeps no process-wide mode -- every other entry point selects automatically):
 *   0 = automatic: warp-specialised TMA/mbarrier kernel when operands meet its 16-byte
 *       granularity, else the cp.async kernel;  1 = cp.async kernel only;
 *   2 = warp-specialised kernel required (PTB_ERR_ALIGNMENT if it cannot run). */
int ptb_gemm_engine(int engine, int dtype, int trans_a, int trans_b, int conj_b, int64_t m, int64_t n, int64_t k,
                    const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t batch,
                    int64_t stride_a, int64_t stride_b, int64_t stride_c, int accumulate, void* stream);

/* ---------------------------------------------------------------------------
 * apply_local_hamiltonian(a, w, l, r)            pytenet/chain_ops.py:237-279
 *   out[i',s',j'] = sum l[i,k,i'] w[k,s',s,kappa] a[i,s,j] r[j,kappa,j']
 *   a (Dl,d_in,Dr)  w (chi_l,d_out,d_in,chi_r)  l (Dl,chi_l,Dlp)  r (Dr,chi_r,Drp)
 *   out (Dlp,d_out,Drp)
 * ------------------------------------------------------------------------- */
size_t ptb_apply_local_hamiltonian_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t Dr,
                                                   int64_t chi_l, int64_t chi_r, int64_t d_out,
                                                   int64_t Dlp, int64_t Drp);
int ptb_apply_local_hamiltonian_z(const void* a, const void* w, int w_is_complex, const void* l, const void* r,
                                  void* out, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                                  int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace,
                                  size_t workspace_bytes, void* stream);
int ptb_apply_local_hamiltonian_d(const void* a, const void* w, const void* l, const void* r, void* out,
                                  int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                                  int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* Same contraction with the MPO tensor given in CSR form: W = w reshaped to (chi_l*d_out) x (d_in*chi_r),
 * row pointers / column indices as device int32 arrays, values float64 (or complex128 when
 * w_is_complex).  MPO tensors of local Hamiltonians are 5-17 % dense, so the W step becomes the
 * HBM-bound sparse kernel below instead of a small dense GEMM.  Workspace as above. */
int ptb_apply_local_hamiltonian_csr_z(const void* a, const int32_t* w_rowptr, const int32_t* w_col, const void* w_val,
                                      int w_is_complex, const void* l, const void* r, void* out, int64_t Dl,
                                      int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out,
                                      int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                                      void* stream);
int ptb_apply_local_hamiltonian_csr_d(const void* a, const int32_t* w_rowptr, const int32_t* w_col, const void* w_val,
                                      const void* l, const void* r, void* out, int64_t Dl, int64_t d_in, int64_t Dr,
                                      int64_t chi_l, int64_t chi_r, int64_t d_out, int64_t Dlp, int64_t Drp,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* Host-buffer form (the call a NumPy user of the reference makes): a, w, l, r and out are HOST pointers with
 * the layouts above (page-locked memory gives full-speed DMA; pageable memory works, staged by the driver);
 * `workspace` is DEVICE memory of at least ptb_apply_local_hamiltonian_host_workspace_bytes(...).  All state
 * tensors have `dtype`; w is float64 unless w_is_complex.  Operands are copied in slices on internal copy
 * streams so that the transfers overlap the three contraction steps: step 1 is sliced along its contraction
 * index (strided 2-D copies of a, contiguous row blocks of r, accumulated), step 3 along the rows of out, which
 * are copied back block by block.  Ordered after prior work on `stream`; SYNCHRONOUS: `out` is complete on
 * return.  Sparse w (<= 4096 non-zeros) takes the CSR W step, otherwise the dense one. */
size_t ptb_apply_local_hamiltonian_host_workspace_bytes(int dtype, int w_is_complex, int64_t Dl, int64_t d_in,
                                                        int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out,
                                                        int64_t Dlp, int64_t Drp);
int ptb_apply_local_hamiltonian_host(int dtype, int w_is_complex, const void* a, const void* w, const void* l,
                                     const void* r, void* out, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l,
                                     int64_t chi_r, int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* W step alone (pytenet/chain_ops.py:276, :52, :96):  t_out[b, m, n] = sum_c W[m, c] t_in[b, c, n],
 * W in CSR form (r_out rows), t_in (batch, r_in, n_cols), t_out (batch, r_out, n_cols) dense, dtype
 * t_dtype.  One pass over t_in and t_out: algorithmic bytes = elem_size (r_in + r_out) n_cols batch. */
int ptb_wapply_csr(int t_dtype, int w_is_complex, int64_t r_out, int64_t r_in, int64_t n_cols, const int32_t* rowptr,
                   const int32_t* col, const void* val, const void* t_in, void* t_out, int64_t batch, void* stream);

/* Same with row-activity flags (sector path): for every (batch block b / batch_block, 128-column block n / 128),
 * in that order, r_in + r_out bytes: flag[c] == 0 marks input row c as structurally zero (not read), flag[r_in + m]
 * == 0 marks output row m as structurally zero (not written: it keeps the zeros the caller initialised once). */
int ptb_wapply_csr_masked(int t_dtype, int w_is_complex, int64_t r_out, int64_t r_in, int64_t n_cols,
                          const int32_t* rowptr, const int32_t* col, const void* val, const void* t_in, void* t_out,
                          int64_t batch, const uint8_t* active, int64_t batch_block, void* stream);

/* ---------------------------------------------------------------------------
 * apply_local_bond_contraction(c, l, r)          pytenet/chain_ops.py:282-317
 *   out[i',j'] = sum l[i,k,i'] c[i,j] r[j,k,j']
 *   c (Dl,Dr)  l (Dl,chi,Dlp)  r (Dr,chi,Drp)  out (Dlp,Drp)
 * ------------------------------------------------------------------------- */
size_t ptb_apply_local_bond_contraction_workspace_bytes(int dtype, int64_t Dl, int64_t Dr, int64_t chi,
                                                        int64_t Dlp, int64_t Drp);
int ptb_apply_local_bond_contraction_z(const void* c, const void* l, const void* r, void* out, int64_t Dl,
                                       int64_t Dr, int64_t chi, int64_t Dlp, int64_t Drp, void* workspace,
                                       size_t workspace_bytes, void* stream);
int ptb_apply_local_bond_contraction_d(const void* c, const void* l, const void* r, void* out, int64_t Dl,
                                       int64_t Dr, int64_t chi, int64_t Dlp, int64_t Drp, void* workspace,
                                       size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * contraction_operator_step_left(a, b, w, l)     pytenet/chain_ops.py:60-99
 *   l_next[j,kappa,j'] = sum l[i,k,i'] conj(b[i',s',j']) w[k,s',s,kappa] a[i,s,j]
 *   a (Dl,d_in,Dr)  b (Dlp,d_out,Drp)  l (Dl,chi_l,Dlp)  l_next (Dr,chi_r,Drp)
 * contraction_operator_step_right(a, b, w, r)    pytenet/chain_ops.py:16-57
 *   r_next[i,k,i'] = sum a[i,s,j] r[j,kappa,j'] w[k,s',s,kappa] conj(b[i',s',j'])
 *   r (Dr,chi_r,Drp)  r_next (Dl,chi_l,Dlp)
 * (one workspace query serves both directions)
 * ------------------------------------------------------------------------- */
size_t ptb_env_step_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l,
                                    int64_t chi_r, int64_t d_out, int64_t Dlp, int64_t Drp);
int ptb_env_step_left_z(const void* a, const void* b, const void* w, int w_is_complex, const void* l,
                        void* l_next, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                        int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                        void* stream);
int ptb_env_step_left_d(const void* a, const void* b, const void* w, const void* l, void* l_next, int64_t Dl,
                        int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out, int64_t Dlp,
                        int64_t Drp, void* workspace, size_t workspace_bytes, void* stream);
int ptb_env_step_right_z(const void* a, const void* b, const void* w, int w_is_complex, const void* r,
                         void* r_next, int64_t Dl, int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r,
                         int64_t d_out, int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                         void* stream);
int ptb_env_step_right_d(const void* a, const void* b, const void* w, const void* r, void* r_next, int64_t Dl,
                         int64_t d_in, int64_t Dr, int64_t chi_l, int64_t chi_r, int64_t d_out, int64_t Dlp,
                         int64_t Drp, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Device-resident Lanczos vector operations        pytenet/krylov.py:12-57
 * All scalars stay on the device (double*), nothing synchronises with the host.
 *
 * ptb_lanczos_scratch_bytes: size of the reduction scratch buffer.
 *
 * ptb_lanczos_start_{d,z}:  v0 = x / |x|;  *nrm = |x|            krylov.py:26-29,36
 *
 * ptb_lanczos_ortho_step_{d,z}: one three-term step               krylov.py:41-43,51
 *     alpha      = Re <w, v_j>
 *     w         -= alpha v_j + beta_prev v_jm1      (v_jm1 / beta_prev may be NULL for j = 0)
 *     beta       = |w|
 *     v_next     = w / beta                         (v_next may alias w)
 *   alpha_out, beta_out, beta_prev are device doubles.
 *
 * ptb_lanczos_alpha_{d,z}: closing alpha = Re <w, v_j>            krylov.py:53-56
 *
 * ptb_krylov_combine: out[n] = sum_j coeff[j] V[j,:], V (k x n) row-major,
 *   coeff on the device.  v_dtype / coeff_dtype in {PTB_REAL64, PTB_COMPLEX128};
 *   out has the promoted dtype.                     krylov.py:118 and :136
 * ------------------------------------------------------------------------- */
size_t ptb_lanczos_scratch_bytes(void);
int ptb_lanczos_start_d(int64_t n, const void* x, void* v0, double* nrm, void* scratch, void* stream);
int ptb_lanczos_start_z(int64_t n, const void* x, void* v0, double* nrm, void* scratch, void* stream);
int ptb_lanczos_ortho_step_d(int64_t n, void* w, const void* v_j, const void* v_jm1, const double* beta_prev,
                             double* alpha_out, double* beta_out, void* v_next, void* scratch, void* stream);
int ptb_lanczos_ortho_step_z(int64_t n, void* w, const void* v_j, const void* v_jm1, const double* beta_prev,
                             double* alpha_out, double* beta_out, void* v_next, void* scratch, void* stream);
int ptb_lanczos_alpha_d(int64_t n, const void* w, const void* v_j, double* alpha_out, void* scratch, void* stream);
int ptb_lanczos_alpha_z(int64_t n, const void* w, const void* v_j, double* alpha_out, void* scratch, void* stream);
int ptb_krylov_combine(int v_dtype, int coeff_dtype, int64_t n, int64_t k, const void* v, int64_t ldv,
                       const void* coeff, void* out, void* stream);

/* expm_krylov's k x k problem and the final combination, on the device (pytenet/krylov.py:122-136, :142-150):
 *   out = V^T U (|vec| exp(dt w) * U[0, :]),  (w, U) = eigen-decomposition of tridiag(alpha, beta)
 * `scal` is the [ |vec|, alpha[0..numiter), beta[0..numiter-1) ] array written by the Lanczos entry points;
 * the breakdown rule of krylov.py:44-50 (beta[j] < 100 n eps) truncates the problem on the device.  No
 * device->host transfer: a TDVP local step stays asynchronous.  numiter <= 64; `out` (n elements) is complex128,
 * or float64 when out_is_complex == 0 (allowed for float64 vectors with dt_im == 0, NumPy's result dtype);
 * coeff_ws: ptb_krylov_expm_workspace_bytes() of device memory, receives the 2*numiter coefficient doubles
 * followed (at double offset 128) by the int32 k_eff. */
size_t ptb_krylov_expm_workspace_bytes(void);
int ptb_krylov_expm_apply(int v_dtype, int64_t n, int numiter, const void* v, int64_t ldv, const double* scal,
                          double dt_re, double dt_im, int out_is_complex, void* coeff_ws, void* out, void* stream);

/* ---------------------------------------------------------------------------
 * Batched per-sector QR of a block-sparse matrix        pytenet/block_sparse_util.py:106-180
 * One launch factorises `nsec` gathered blocks A[rows_s, cols_s] (one CTA each, block held in shared memory,
 * unblocked Householder QR with LAPACK's zgeqr2 / zung2r conventions: real diagonal of R) and scatters
 *   q[rows_s[i], pos_s + j] = Q_s[i, j],   r[pos_s + i, cols_s[j]] = R_s[i, j]   (j >= i),   i, j < min(m_s, n_s)
 * into the caller's zero-initialised row-major outputs q (ldq) and r (ldr).  a is row-major with leading dimension
 * lda.  meta: 8 int32 per sector {m, n, offset into rowidx, offset into colidx, pos, 0, 0, 0}; every block must
 * satisfy m*n <= max_block_elems and max_block_elems * element size <= ptb_block_qr_max_block_bytes() (larger
 * blocks are the caller's: cuSOLVER).  All tables are device arrays.
 * ------------------------------------------------------------------------- */
size_t ptb_block_qr_max_block_bytes(void);
int ptb_block_qr(int dtype, const void* a, int64_t lda, int nsec, const int32_t* meta, int max_block_elems,
                 const int32_t* rowidx, const int32_t* colidx, void* q, int64_t ldq, void* r, int64_t ldr, void* stream);

/* ---------------------------------------------------------------------------
 * Batched per-sector SVD of a block-sparse matrix        pytenet/block_sparse_util.py:244-319
 * One launch factorises `nsec` gathered blocks A[rows_s, cols_s] = U_s diag(sigma_s) V_s^H (one CTA each, one-sided
 * Jacobi in shared memory, singular values in descending order) and scatters
 *   u[rows_s[i], pos_s + j] = U_s[i, j],   s[pos_s + j] = sigma_s[j],   vh[pos_s + j, cols_s[c]] = conj(V_s[c, j])
 * for j < min(m_s, n_s) into the caller's zero-initialised row-major outputs u (ldu), vh (ldv) and the device
 * vector s.  meta as for ptb_block_qr; every block must satisfy
 *   max(m,n)*min(m,n) + min(m,n)^2 <= max_work_elems,  max_work_elems * element size <= ptb_block_svd_max_block_bytes().
 * Columns belonging to exactly zero singular values are left zero.  All tables are device arrays.
 * ------------------------------------------------------------------------- */
size_t ptb_block_svd_max_block_bytes(void);
int ptb_block_svd(int dtype, const void* a, int64_t lda, int nsec, const int32_t* meta, int max_work_elems,
                  const int32_t* rowidx, const int32_t* colidx, void* u, int64_t ldu, double* s, void* vh, int64_t ldv,
                  void* stream);

/* ---------------------------------------------------------------------------
 * MPO-bond-sharded contractions (SURVEY.md 8(b) minimum list, 8(e); BASELINE config 4)
 *   the sum over MPO-bond index pairs of pytenet/chain_ops.py:237-279 / :282-317 / :60-99 split over the ranks
 *   of a communicator; see csrc/sharded.cu for the scheme.  One process per GPU.
 *
 * ptb_comm: opaque handle around an NCCL communicator (libnccl resolved at run time).  Rank 0 obtains a 128-byte
 * id with ptb_comm_unique_id and distributes it by any means (MPI, sockets, a file, torch.distributed); every rank
 * then calls ptb_comm_init on its current device (collective).  comm == NULL everywhere means a single rank.
 * This handle is the only process-level state of the library; all entries are asynchronous on `stream`.
 *
 * Shapes: a (Dl, d_in, Dr); lw (Dl*d_in*P, d_out, Dlp) from ptb_sharded_precontract with
 * w3[(s, kappa_loc, s'), k] = w[k, s', s, kappa] on this rank's zero-padded range of P = ceil(chi_r / nranks)
 * right-bond indices and l (Dl, chi_l, Dlp) the FULL left block; r_shard (Dr, P, Drp) the rank's range of the right
 * block; out (Dlp, d_out, Drp), identical on every rank after the all-reduce.
 * env_step_left_sharded: b (Dlp, d_out, Drb) the bra tensor; l_next (Dr, P, Drb) = this rank's range of the next
 * left block -- no communication.  The right-to-left direction applies the same entries to mirrored tensors.
 * bond contraction: c (Dl, Dr), l_shard (Dl, P, Dlp), r_shard (Dr, P, Drp) ranges of the SAME MPO bond.
 * ------------------------------------------------------------------------- */
typedef struct ptb_comm ptb_comm;
int ptb_comm_unique_id(void* id128);
int ptb_comm_init(ptb_comm** comm, int nranks, int rank, const void* id128);
int ptb_comm_destroy(ptb_comm* comm);
int ptb_comm_info(const ptb_comm* comm, int* nranks, int* rank);
int ptb_allreduce_sum(ptb_comm* comm, int dtype, void* buf, int64_t count, void* stream);
int ptb_sharded_precontract(int dtype, int w_is_complex, const void* w3, const void* l, void* lw, int64_t Dl,
                            int64_t chi_l, int64_t Dlp, int64_t R, void* stream);
size_t ptb_apply_local_hamiltonian_sharded_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t Dr, int64_t P,
                                                           int64_t d_out, int64_t Dlp, int64_t Drp);
int ptb_apply_local_hamiltonian_sharded(ptb_comm* comm, int dtype, const void* a, const void* lw, const void* r_shard,
                                        void* out, int64_t Dl, int64_t d_in, int64_t Dr, int64_t P, int64_t d_out,
                                        int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                                        void* stream);
size_t ptb_env_step_left_sharded_workspace_bytes(int dtype, int64_t Dl, int64_t d_in, int64_t P, int64_t Drb);
int ptb_env_step_left_sharded(int dtype, const void* a, const void* b, const void* lw, void* l_next, int64_t Dl,
                              int64_t d_in, int64_t Dr, int64_t P, int64_t d_out, int64_t Dlp, int64_t Drb,
                              void* workspace, size_t workspace_bytes, void* stream);
size_t ptb_apply_local_bond_contraction_sharded_workspace_bytes(int dtype, int64_t Dl, int64_t P, int64_t Drp);
int ptb_apply_local_bond_contraction_sharded(ptb_comm* comm, int dtype, const void* c, const void* l_shard,
                                             const void* r_shard, void* out, int64_t Dl, int64_t Dr, int64_t P,
                                             int64_t Dlp, int64_t Drp, void* workspace, size_t workspace_bytes,
                                             void* stream);

/* ---------------------------------------------------------------------------
 * Large dense SVD of a two-site split          pytenet/bond_ops.py:41-54, block_sparse_util.py:294
 * A (rows x cols, COLUMN-major, leading dimension lda, destroyed) = U diag(s) V^H by cuSOLVER's polar-decomposition
 * driver (cusolverDnXgesvdp, economy size): u is rows x min (column-major, ldu), v is cols x min (column-major,
 * ldv; V itself, not V^H), s descending.  A row-major matrix M (m x n) is passed as its transpose (rows = n, cols =
 * m): then the `u` buffer read row-major is V_M^H and the `v` buffer read row-major is U_M^H.  `info` is a device
 * int, `err_sigma` a host double (0 when the result is accurate to working precision).  Workspaces as reported
 * by ptb_svd_polar_workspace_bytes (device and host part).  cuSOLVER is resolved at run time (dlopen).
 * ------------------------------------------------------------------------- */
int ptb_svd_polar_workspace_bytes(int dtype, int64_t rows, int64_t cols, size_t* device_bytes, size_t* host_bytes);
int ptb_svd_polar(int dtype, int64_t rows, int64_t cols, void* a, int64_t lda, double* s, void* u, int64_t ldu, void* v,
                  int64_t ldv, void* device_ws, size_t device_bytes, void* host_ws, size_t host_bytes, int* info,
                  double* err_sigma, void* stream);

/* Batched form: the independent sector blocks of ONE split (pytenet/block_sparse_util.py:276-300 loops over the
 * sectors and calls LAPACK per block) factorised concurrently -- every job as ptb_svd_polar, run on up to
 * `max_workers` (<= 8; <= 0: 8) internal worker streams with their own cuSOLVER handles; the work is ordered after
 * everything enqueued on `stream` so far, and `stream` waits for all jobs before the call returns.  Per job:
 * `device_ws` of at least the size ptb_svd_polar_workspace_bytes reports, `info` a device int; `err_sigma` and
 * `status` (PTB_OK or an error code) are written on return.  Returns the last non-OK job status, else PTB_OK. */
typedef struct ptb_svd_job {
    int64_t rows, cols;
    void* a; int64_t lda;              /* column-major rows x cols, destroyed */
    double* s;                         /* min(rows, cols) */
    void* u; int64_t ldu;              /* column-major rows x min */
    void* v; int64_t ldv;              /* column-major cols x min */
    void* device_ws; size_t device_bytes;
    int* info;                         /* device */
    double err_sigma;                  /* out */
    int status;                        /* out */
    int reserved;
} ptb_svd_job;
int ptb_svd_polar_batch(int dtype, int njobs, ptb_svd_job* jobs, int max_workers, void* stream);

/* ---------------------------------------------------------------------------
 * A whole Lanczos run on the local effective Hamiltonian in ONE call
 *   pytenet/krylov.py:12-57 driven by the closures tdvp.py:223-229 (site), tdvp.py:232-238 (bond),
 *   dmrg.py:181-189.
 * Enqueues  v0 = x/|x|;  for j < numiter: w = H_eff v_j, three-term step  -- the same kernels as the
 * per-step entry points above, issued back to back from C (no host code, allocation or synchronisation
 * between the iterations: small-bond local problems are launch-latency-bound).  H_eff is square
 * (Dlp = Dl, Drp = Dr, d_out = d_in).
 *   V    : numiter x n Lanczos vectors (row-major), n = Dl*d*Dr (site) or Dl*Dr (bond)
 *   scal : 2*numiter device doubles  [ |x|, alpha[0..numiter), beta[0..numiter-1) ]
 *   w    : dense MPO tensor, or NULL when the CSR form (w_rowptr, w_col, w_val) is given; when both
 *          are given the CSR form is used.
 * The breakdown test of krylov.py:44-50 is the caller's, applied to the betas afterwards (entries
 * before the breakdown index do not depend on later steps).
 * ------------------------------------------------------------------------- */
size_t ptb_heff_lanczos_workspace_bytes(int dtype, int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r);
int ptb_heff_lanczos(int dtype, const void* x, const void* w, int w_is_complex, const int32_t* w_rowptr,
                     const int32_t* w_col, const void* w_val, const void* l, const void* r, int64_t Dl, int64_t d,
                     int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter, void* V, double* scal, void* scratch,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t ptb_bond_lanczos_workspace_bytes(int dtype, int64_t Dl, int64_t Dr, int64_t chi);
int ptb_bond_lanczos(int dtype, const void* c, const void* l, const void* r, int64_t Dl, int64_t Dr, int64_t chi,
                     int numiter, void* V, double* scal, void* scratch, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---------------------------------------------------------------------------
 * One kernel per local problem (launch-latency regime: README config, METTS, chain edges)
 *   pytenet/tdvp.py:223-238, dmrg.py:181-189 (closures) + krylov.py:12-57 (Lanczos) + krylov.py:110-139 (expm_krylov)
 * ONE launch runs the start normalisation, all `numiter` Lanczos iterations on H_eff = (l, w, r) -- the three
 * contraction steps of chain_ops.py:273-278 as FP64 FMA loops over operands staged in shared memory --, and, with
 * `apply_expm`, the numiter x numiter tridiagonal problem and out = exp(dt H_eff) x as the combination of the
 * Lanczos vectors (dt = dt_re + i dt_im; the drivers pass -dt).  The launch is a thread-block cluster of 1, 2, 4 or
 * 8 CTAs, each owning a slice of the right bond index of the output; the Lanczos vector and the two scalars of an
 * iteration are exchanged through distributed shared memory.  `w == NULL` is the zero-site problem
 * (apply_local_bond_contraction, chain_ops.py:282-317: d == 1, chi_l == chi_r).  V (numiter x n elements) receives the
 * Lanczos vectors, scal = [|x|, alpha[0:numiter], beta[0:numiter-1]] as ptb_heff_lanczos -- `scal` is only written
 * and may point to page-locked host memory (mapped): the drivers let the kernel deposit its scalars in the ring their
 * deferred checks read.  All numiter steps are executed, the breakdown rule of krylov.py:44-50 is applied to the
 * betas inside the k x k solve (and by the caller to `scal`).  The k x k solve: Krylov spaces up to 16 by the
 * shifted, scaled Taylor series of the tridiagonal matrix with squaring (same result as the eigenvector formula of
 * krylov.py:136 to a few 2^s eps; refused when |Re dt| |T - mean| > 1.5 or more than ten squarings are needed),
 * otherwise the implicit QL iteration.  ptb_local_step_small_fits: 1 when one matvec is at most 400 000 complex
 * multiply-adds, the operands fit the shared memory of the cluster and numiter <= 64; otherwise
 * ptb_local_step_small returns PTB_ERR_TOO_LARGE.  `workspace` is unused (kept for ABI stability; may be NULL).
 * ------------------------------------------------------------------------- */
int ptb_local_step_small_fits(int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter);
size_t ptb_local_step_small_workspace_bytes(int dtype, int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r);
int ptb_local_step_small(int dtype, const void* x, const void* w, int w_is_complex, const void* l, const void* r,
                         int64_t Dl, int64_t d, int64_t Dr, int64_t chi_l, int64_t chi_r, int numiter, void* V,
                         double* scal, int apply_expm, double dt_re, double dt_im, int out_is_complex, void* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Sector-packed block-sparse path (BASELINE config 3: quantum-number sectors as a device-side grouped GEMM)
 *   sector structure: pytenet/block_sparse_util.py:47-53 (sparsity rule), :151-169 (sector order);
 *   contraction: pytenet/chain_ops.py:237-279.
 * With bonds grouped by sector, `a`, `l`, `r` consist of dense blocks.  pytenet_b200/sector_packed.py packs
 * them so that the three steps become
 *   (1) per right sector beta:   T1_beta (M x N) = AT_beta^T (M x n_beta) RB_beta (n_beta x N)
 *   (2) T2 = W . T1 as block gathers with the non-zero MPO entries as coefficients
 *   (3) per left sector alpha':  O_alpha' (N' x n_alpha') = T2_alpha'^T (N' x K') LP_alpha' (K' x n_alpha')
 * -- per group two large stacked extents and one sector-sized one, no structural zero stored or multiplied.
 *
 * ptb_gemm_grouped: one launch over `ntiles` independent output tiles (device table): C(m x n) (+)= A^T B with A
 * stored k x m (element (k,m) at a[a_off + k*lda + m]), B stored k x n, C row-major; a_off / b_off / c_off are ELEMENT
 * offsets from the three base pointers to the tile's origin (so one table serves any buffers of the same layout);
 * m <= 128, n <= 64 (complex128) / 128 (float64) = ptb_gemm_tile_shape; k >= 1.  The persistent CTAs take the
 * tiles round-robin in table order (sort by decreasing k).  complex128: any extents; float64: m, n, lda, ldb and
 * the offsets must be even (16-byte rows).
 *
 * ptb_block_gather: dst chunk (rows x cols at dst + dst_off, leading dimension dst_ld) = sum over its terms of
 * coef * src[src_off + r*src_rs + c*src_cs]; `work` lists {chunk, first row, number of rows, 0} per CTA.  Serves the
 * W step (coefficients = MPO entries), the packing of a / l / r and the unpacking of the result (one term, coef 1,
 * strides express transposition).  coef_im is ignored for float64 data.  A chunk with flags bit 0 set reads the
 * complex conjugate of its source elements (bra tensors of the environment updates, pytenet/chain_ops.py:55,93).
 * The environment updates run on the same two kernels (pytenet_b200/sector_packed.py: PackedEnvPlan):
 *   step_right:  T1^T_beta = RB_beta^T AT_beta,  T2^T = W . T1^T (block gather),
 *                r_next as LP-layout blocks  (K' x n_alpha') = T2^T_alpha'^T (K' x N') conj(BT_alpha') (N' x n_alpha')
 *   step_left:   the same on the mirrored tensors (strides of the gather tables, no copies).
 * ------------------------------------------------------------------------- */
typedef struct ptb_group_tile {
    int64_t a_off, b_off, c_off;
    int32_t lda, ldb, ldc;
    int32_t m, n, k;
    int32_t accumulate;
    int32_t reserved[3];
} ptb_group_tile;                       /* 64 bytes */
int ptb_gemm_grouped(int dtype, const void* a, const void* b, void* c, const ptb_group_tile* tiles, int ntiles,
                     void* stream);
/* Tile variants of the grouped GEMM: 0 = the engine's 128 x 64 (complex) / 128 x 128 (real) tile (ptb_gemm_grouped),
 * 1 = 64 x 32 / 64 x 64 for groups whose sector-sized extents would leave most of the large tile empty (zero-site
 * problem, fragmented sector profiles).  The tile table must be built for the variant's shape
 * (ptb_gemm_grouped_tile_shape); same table format and operand conventions. */
int ptb_gemm_grouped_tile_shape(int dtype, int variant, int* bm, int* bn);
int ptb_gemm_grouped_v(int dtype, int variant, const void* a, const void* b, void* c, const ptb_group_tile* tiles,
                       int ntiles, void* stream);

typedef struct ptb_gather_chunk {
    int64_t dst_off;
    int32_t dst_ld, rows, cols, term_begin, term_end, flags;      /* flags bit 0: conjugate the source */
} ptb_gather_chunk;                     /* 32 bytes */
typedef struct ptb_gather_term {
    int64_t src_off;
    int32_t src_rs, src_cs;
    double coef_re, coef_im;
} ptb_gather_term;                      /* 32 bytes */
int ptb_block_gather(int dtype, const void* src, void* dst, const ptb_gather_chunk* chunks,
                     const ptb_gather_term* terms, const int32_t* work, int nwork, void* stream);

/* ---------------------------------------------------------------------------
 * Diagnostics: register-resident DMMA.8x8x4 (use_dmma != 0) or DFMA loop on
 * `blocks` CTAs x 256 threads, to measure the FP64 pipe peak that the roofline
 * fractions are quoted against.  `out`: blocks*256 device doubles; *flops
 * (host) receives the number of floating-point operations issued.
 * ------------------------------------------------------------------------- */
int ptb_probe_fp64_pipe(int use_dmma, int blocks, int iters, double* out, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PYTENET_B200_H */
