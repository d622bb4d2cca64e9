/*
 * grief_b200.h -- C ABI of the B200-native GP-GRIEF hot path (libgrief_b200.so).
 *
 * The reference (scwolof/gp_grief) is pure Python; it has no FFI of its own.  The boundary this
 * library drops in behind is therefore the set of Python methods listed beside each entry point
 * (file:line relative to the reference's gp_grief/ package).  INTEGRATION.md shows the ctypes stubs
 * a maintainer of the reference would add at those lines.
 *
 * Conventions
 *   - every function returns 0 on success or one of the GRIEF_ERR_* codes; the message of the last
 *     failure on the calling thread is returned by grief_last_error().
 *   - pointers named *_dev are CUDA device pointers on the current device, *_host are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - dimension order is INPUT-dimension order i = 0..d-1 everywhere in this ABI.  (The reference
 *     stores its Kronecker factors reversed, kern/grid_kernel.py:109,174; the Python host layer does
 *     that bookkeeping.)  The one exception is grief_topk_kron, which takes the factor list in the
 *     order it is given, exactly like KronMatrix.find_extremum_eigs.
 *   - all floating point data is IEEE double; matrices are dense row-major unless stated.
 *   - the library never keeps a pointer to caller memory after a call returns.
 */
#ifndef GRIEF_B200_H
#define GRIEF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRIEF_OK 0
#define GRIEF_ERR_BAD_ARG 1      /* -> ValueError / AssertionError on the Python side            */
#define GRIEF_ERR_CUDA 2         /* -> RuntimeError                                             */
#define GRIEF_ERR_NOT_PD 3       /* -> numpy.linalg.LinAlgError (what scipy cho_factor raises)   */
#define GRIEF_ERR_UNSUPPORTED 4  /* -> NotImplementedError                                      */
#define GRIEF_ERR_LIBRARY 5      /* -> RuntimeError (driver entry point missing)                */
#define GRIEF_ERR_NCCL 6         /* -> RuntimeError (NCCL not loadable, or a failed collective)  */

#define GRIEF_KERN_RBF 0         /* kern/stationary.py:108-134 */
#define GRIEF_KERN_EXPONENTIAL 1 /* kern/stationary.py:161-175 */
#define GRIEF_KERN_MATERN32 2    /* kern/stationary.py:202-216 */
#define GRIEF_KERN_MATERN52 3    /* kern/stationary.py:243-258 */

/* indices into the `scalars_host` array filled by grief_solve_lml */
#define GRIEF_SC_LML 0
#define GRIEF_SC_YT_ALPHA 1
#define GRIEF_SC_LOGDET 2
#define GRIEF_SC_GRAD_NOISE 3
#define GRIEF_SC_RTB 4
#define GRIEF_SC_ALPHA_SQ 5
#define GRIEF_SC_TRACE 6
#define GRIEF_SC_COUNT 7

typedef struct grief_plan grief_plan; /* one basis (one value of the kernel hyper-parameters) */
typedef struct grief_ctx grief_ctx;   /* per-model scratch: dense-stage workspaces             */

int grief_version(void);
const char* grief_last_error(void);
/* number of CUDA kernels this library has launched on the calling thread since the last reset */
int grief_launch_count(void);
void grief_launch_count_reset(void);

/*
 * Optional per-kernel timing for benchmarks: CUDA events are recorded on the launch stream around the kernels
 * of each slot.  Slots (grief_profile_slots() = 10): 0 Gram GEMM (pass 1), 1 Phi*B GEMM (pass 2), 2 table prepass, 3 backward
 * sweep of the gradient pass (k_contract_back), 4 top-p select, 5 p x p stage, 6 stand-alone Phi^T v, 7 gradient tail (K_xu / F buffer +
 * k_contract_tail), 8 Phi^T slab builder (pass 1, incl. the fused Phi^T y and the row maxima), 9 Phi slab builder (pass 2, incl. the
 * fused residual).  grief_profile_read synchronises,
 * writes the accumulated milliseconds and launch counts of every slot (arrays of grief_profile_slots()) and resets.
 */
void grief_profile_enable(int on);
int grief_profile_slots(void);
void grief_profile_read(double* ms_out, int* count_out);

int grief_ctx_create(grief_ctx** ctx);
void grief_ctx_destroy(grief_ctx* ctx);

/*
 * Top-p selection over the Kronecker product of eigenvalue vectors.
 * Replaces KronMatrix.find_extremum_eigs(n_eigs, mode='largest', log_expand=True, sort=True)
 * (tensors/kron_matrix.py:369-446) as called from GriefKernel._setup_inducing_cov
 * (kern/grief_kernel.py:184).
 *   d, m_host[d]        number of factors and their sizes (factor k <-> KronMatrix.K[k])
 *   raw0_host[m[0]]     raw eigenvalues of factor 0 (selection key of the first step, :407)
 *   logeig_host[sum m]  np.log of every factor's eigenvalues, concatenated, computed by the caller
 *                       with the same NumPy call the reference uses (linalg.py:86-89)
 *   p                   n_eigs; must not exceed prod(m)
 *   idx_dev[p*d]        out, int32 row-major (p, d): eig_loc, column k <-> factor k
 *   loglam_dev[p]       out: log eigenvalue products, descending
 *   n_out_host          out: number of entries written (== p)
 */
int grief_topk_kron(int d, const int32_t* m_host, const double* raw0_host, const double* logeig_host, int p,
                    int32_t* idx_dev, double* loglam_dev, int* n_out_host, void* stream);

/*
 * Describe one basis.  Replaces the state GriefKernel._setup_inducing_cov leaves behind
 * (kern/grief_kernel.py:168-190: _Quu, _log_lam, _Sp) in the form the row kernels consume.
 *   m[d], kernel_id[d], variance[d], lengthscale[d], grid_concat[sum m]   per input dimension
 *                        kernel_id: GRIEF_KERNEL_RBF / _EXPONENTIAL / _MATERN32 / _MATERN52 are evaluated on the device;
 *                        GRIEF_KERNEL_HOST marks a dimension whose kernel only the caller can evaluate (a GPy kernel, a kernel
 *                        with children, any Python cov: kern/gpy_kernel.py:45-58) -- its K_xu columns come in through
 *                        grief_build_tables_kxu, variance / lengthscale are ignored
 *   u[d]                 number of distinct selected eigen-indices of the dimension
 *                        (SelectionMatrixSparse.unique, tensors/selection_matrix.py:78-79)
 *   qs_concat            per dimension an (m_i, u_i) row-major matrix: column k is the Schur vector of
 *                        the k-th unique eigen-index divided by sqrt(its eigenvalue)
 *                        (tensors/tensors.py:118 with kern/grief_kernel.py:104 folded in per dimension)
 *   p, uinv[p*d]         per basis column and dimension the position in the unique list
 *                        (SelectionMatrixSparse.unique_inverse)
 *   width_cap            0 = default; upper bound on table entries per row (testing knob)
 * All inputs are host pointers and are copied.
 */
int grief_plan_create(grief_plan** plan, int d, const int32_t* m, const int32_t* kernel_id,
                      const double* variance, const double* lengthscale, const double* grid_concat,
                      const int32_t* u, const double* qs_concat, int p, const int32_t* uinv, int width_cap);
void grief_plan_destroy(grief_plan* plan);
#define GRIEF_KERNEL_RBF 0
#define GRIEF_KERNEL_EXPONENTIAL 1
#define GRIEF_KERNEL_MATERN32 2
#define GRIEF_KERNEL_MATERN52 3
#define GRIEF_KERNEL_HOST 4
/* layout queries: what = 0 groups G, 1 table width, 2 row stride (doubles), 3 p, 4 p_pad, 5 d, 6 sum of the grid sizes m_i
 * (columns of K_xu in grief_build_tables_kxu; dimension i starts at m_0 + ... + m_{i-1}), 7 number of host-evaluated dimensions */
int grief_plan_info(const grief_plan* plan, int what);
/* rows the table buffer must hold for n data rows (n rounded up to the MMA chunk) */
int64_t grief_table_rows(int64_t n);

/*
 * Prepass: group tables of n data rows.  Replaces GridKernel.cov_kr (kern/grid_kernel.py:148-179),
 * the per-dimension projection of expand_SKC (tensors/tensors.py:118) and the first levels of its
 * product (tensors/tensors.py:120-124).
 *   X_dev       (n, ldx) row-major inputs, column i = input dimension i
 *   T_dev       out, (grief_table_rows(n), stride) row-major; rows >= n are written as zeros
 */
int grief_build_tables(const grief_plan* plan, const double* X_dev, int64_t ldx, int64_t n, double* T_dev, void* stream);

/*
 * Same tables with the kernel of input dimension `dim` replaced by its derivative with respect to x: every product formed from
 * them (grief_phi_rows, grief_phi_vec, ...) is then d/dx[:, dim] of the corresponding quantity.  grief_phi_vec on these tables with
 * v = alpha_p gives GPGriefModel.d_Yhat_d_x (models/gp_grief_model.py:127-134; the reference needs GPy for the kernel derivative,
 * kern/grid_kernel.py:181-206).
 */
int grief_build_tables_dx(const grief_plan* plan, const double* X_dev, int64_t ldx, int64_t n, int dim, double* T_dev, void* stream);

/*
 * Tables for a plan with host-evaluated dimensions (GRIEF_KERNEL_HOST): GridKernel.cov_kr (kern/grid_kernel.py:148-179) calls
 * kern.cov(x[:, i], U_i) per dimension -- for kernels that exist only as host code (GPyKernel, kern/gpy_kernel.py:45-58) the caller
 * evaluates those (n, m_i) blocks, uploads them, and the device does the rest (projection on the scaled eigenvectors, group
 * products).  Dimensions with device kernels are still evaluated on the device from X_dev.
 *   Kxu_dev, ldk   (n, ldk) row-major, ldk >= sum m_i; columns [off_i, off_i + m_i) = K_xu,i of a host dimension i (others unused)
 *   deriv_dim      -1, or the dimension whose kernel is replaced by its x-derivative (a host dimension then passes d K_xu,i / d x)
 * Row chunks may be built one call at a time (T_dev + r0 * stride, r0 a multiple of 128).
 */
int grief_build_tables_kxu(const grief_plan* plan, const double* X_dev, int64_t ldx, const double* Kxu_dev, int64_t ldk, int64_t n,
                           int deriv_dim, double* T_dev, void* stream);

/* Phi (n, p) row-major from the tables: GriefKernel.cov(x)[0] (kern/grief_kernel.py:68-111). Small n only. */
int grief_phi_rows(const grief_plan* plan, const double* T_dev, int64_t n, double* Phi_dev, void* stream);

/*
 * Options of the two O(n p^2) products (A = Phi^T Phi and Z = Phi B).  Every plan carries its own values; a new plan copies the
 * CALLING THREAD's defaults (grief_set_default_option), and grief_plan_set_option changes one plan.  Nothing is process-global.
 * Workspace sizes depend on the options: query them after changing one.
 *   GRIEF_OPT_GEMM_MODE    arithmetic:
 *       0  FP64 DMMA GEMM (k_gemm_nt), 36 TFLOP/s
 *       1  FP64 emulated on the INT8 tensor cores (tcgen05 kind::i8): every operand row is scaled by a power of two and cut
 *          into D balanced 8-bit digits (8 D - 2 bits + sign, round to nearest); D (D + 1) / 2 exact int8 x int8 -> int32 digit
 *          products; element errors are bounded by ~2^(1 - 8 D) of the product of the operands' row maxima times K
 *          (D = 7: the accuracy class of DGEMM); one CTA per 128 x 128 tile
 *       3  (default) as 1, with CTA pairs (thread-block clusters of 2) computing 256 x 128 tiles through tcgen05.mma.cta_group::2
 *          wherever the product has at least 16 x 16 tiles (p >= 2048): each CTA stages half of the B digits, a quarter less
 *          L2 -> SM traffic per MMA.  Same arithmetic, bit-identical results; +12 % (Gram) / +5 % (Phi P^-1) at C3 and +16 % at
 *          p = 8192 with the round-2 digit counts (round 1, 7 digits: no gain under the power cap); smaller products keep mode 1
 *       5  as 3, with CTA pairs for every product with at least two row tiles (testing)
 *   GRIEF_OPT_DIGITS_GRAM  D of A = Phi^T Phi (3..7, default 6: 46-bit operands, 22 digit products).  A feeds a Cholesky
 *                          factorisation; measured against the FP64 mode at n = 10^6..10^7, p = 4096: LML identical to 1e-15 with
 *                          D = 6 and D = 7, 1.5e-13 with D = 5
 *   GRIEF_OPT_DIGITS_Z     D of Zp = Phi P^-1 in grief_grad_theta (3..7, default 4: 30-bit operands, 10 digit products).  The
 *                          gradient sums n p products of Zp, so its truncation errors average out: 1.5e-13 of the largest gradient
 *                          component with D = 4, 2e-15 with D >= 5, 6e-11 with D = 3 (same measurement)
 *   GRIEF_OPT_DIGITS_VAR   D of Z = Phi B in grief_quadform_rows (3..7, default 6): a predictive variance sums only p products
 *   GPGriefModel audits these choices a posteriori against the FP64 mode on a row sample (models/gp_grief_model.py, arithmetic_audit).
 *   GRIEF_OPT_SLAB_BUDGET  bytes of HBM for the Phi^T slab that pass 1 stages per GEMM launch (default 1.25 GiB of FP64-equivalent rows, rounded down to whole waves of the slab builder -- 37888 rows at p = 4096; a K split of the digit planes then passes through L2 in pieces that the tiles in flight share; 0 restores it).  Smaller
 *                          budgets mean more, shorter launches; results are identical up to the order of the fixed-order accumulation
 * grief_set_slab_budget / grief_set_gemm_mode / grief_get_gemm_mode are the round-1 spellings of the default setters.
 */
#define GRIEF_OPT_GEMM_MODE 0
#define GRIEF_OPT_DIGITS_GRAM 1
#define GRIEF_OPT_DIGITS_Z 2
#define GRIEF_OPT_SLAB_BUDGET 3
#define GRIEF_OPT_DIGITS_VAR 4
int grief_set_default_option(int what, int64_t value);
int64_t grief_get_default_option(int what);
int grief_plan_set_option(grief_plan* plan, int what, int64_t value);
int64_t grief_plan_get_option(const grief_plan* plan, int what);
void grief_set_slab_budget(size_t bytes);
void grief_set_gemm_mode(int mode);
int grief_get_gemm_mode(void);

/*
 * Gram matrix A = Phi^T Phi (models/gp_grief_model.py:148-149).  Phi is never held whole: one slab of rows at a time is staged in
 * the workspace (as int8 digit planes in arithmetic mode 1, as FP64 in mode 0) and consumed by the library's GEMM.
 *   A_dev       out, (p, lda) row-major, full symmetric matrix
 *   workspace   device scratch of at least grief_gram_workspace_bytes(plan, n) bytes
 */
size_t grief_gram_workspace_bytes(const grief_plan* plan, int64_t n);
int grief_gram(const grief_plan* plan, const double* T_dev, int64_t n, double* A_dev, int64_t lda,
               void* workspace_dev, size_t workspace_bytes, void* stream);
/*
 * A = Phi^T Phi and r = Phi^T y (models/gp_grief_model.py:148-149 and :234) from ONE sweep over the rows: the builder that stages a
 * slab of Phi has every element in a register, so y^T Phi is accumulated there (fixed order) instead of a second pass over Phi.
 *   y_dev (n), r_dev (p) out.  r_dev == NULL: same as grief_gram.
 *   rowmax_dev  out or NULL: int32[grief_table_rows(n)], the high word of max_j |Phi[row, j]| of every row (the same sweep sees every
 *               element).  Handing it to grief_grad_theta for the SAME tables saves that call one of its two sweeps over Phi.
 */
int grief_gram_ry(const grief_plan* plan, const double* T_dev, int64_t n, const double* y_dev, double* A_dev, int64_t lda,
                  double* r_dev, int32_t* rowmax_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* out[j] = sum_n v[n] Phi[n,j]  (r = Phi^T y, models/gp_grief_model.py:234).  ws: grief_phi_t_vec_workspace_bytes */
size_t grief_phi_t_vec_workspace_bytes(const grief_plan* plan, int64_t n);
int grief_phi_t_vec(const grief_plan* plan, const double* T_dev, int64_t n, const double* v_dev, double* out_dev,
                    void* workspace_dev, void* stream);
/* out[n] = sum_j Phi[n,j] v[j]  (predictive mean Phi* alpha_p, models/gp_grief_model.py:119) */
int grief_phi_vec(const grief_plan* plan, const double* T_dev, int64_t n, const double* v_dev, double* out_dev, void* stream);
/* out[0] = sum_n y[n]^2 ; ws_dev >= 4096 bytes */
int grief_sumsq(const double* y_dev, int64_t n, double* out_dev, void* ws_dev, void* stream);

/*
 * The p x p stage (models/gp_grief_model.py:152-153 cho_factor, :234 cho_solve, :243-245 log-det,
 * :212-213 LML, :171-180 d/dw, :185-191 d/d noise_var), from the reduced statistics A, r, y^T y.
 *   n_rows          total number of data rows (over all GPUs)
 *   L_dev           out (p,p): upper Cholesky factor U of P = A + diag(noise_var/w), P = U^T U, row-major,
 *                   strictly lower part zero (what scipy.linalg.cho_factor + np.triu give)
 *   b_dev           out (p): P^-1 r  (== alpha_p of models/gp_grief_model.py:97)
 *   Pinv_dev        out (p,p) or NULL: P^-1 (needed for the gradients)
 *   grad_w_dev      out (p) or NULL
 *   G2_dev          out (p,p) or NULL: -(P^-1 + b b^T / noise_var) (the round-1 operand of pass 2; grief_grad_theta takes P^-1 now)
 *   scalars_host    out double[GRIEF_SC_COUNT]
 *   info_host       out: 0, or the order of the leading minor that is not positive definite
 */
int grief_solve_lml(grief_ctx* ctx, int p, const double* A_dev, int64_t lda, const double* r_dev,
                    const double* yty_dev, const double* w_dev, double noise_var, int64_t n_rows, double* L_dev,
                    double* b_dev, double* Pinv_dev, double* grad_w_dev, double* G2_dev, double* scalars_host,
                    int* info_host, void* stream);

/*
 * Analytic hyper-parameter gradient (pass 2).  The reference has no counterpart: for kernel
 * hyper-parameters it falls back to forward finite differences (models/gp_grief_model.py:71-74,
 * :194-196; models/basemodel.py:328-361), i.e. (#free + 1) complete evaluations.
 *
 * grief_grad_setup: declare the parameters to differentiate.
 *   dims[a], kinds[a]   input dimension and kind (0 = variance, 1 = lengthscale) of active parameter a
 *   dqs_concat          per active parameter an (m_i, u_i) row-major matrix d(qs_i)/d(theta_a): derivative of
 *                       the scaled eigenvectors passed to grief_plan_create (host-side eigen-perturbation)
 * grief_grad_theta: grad_dev[a] = d LML / d theta_a for the rows given (a row shard when multi-GPU: sum over
 *   ranks), holding the selected eigen-index set fixed (as finite differences implicitly do).
 *   d LML / d theta = sum_n sum_j Zt[n,j] dPhi[n,j]/dtheta with Zt = -Phi P^-1 + a b^T, a = (y - Phi b) / noise_var: only Phi P^-1
 *   goes through the O(n p^2) GEMM (arithmetic and digits per the plan's options), the rank-one part is formed in FP64.
 *   T_dev          tables of the n rows (grief_build_tables)
 *   Pinv_dev, ldp  P^-1 from grief_solve_lml, symmetric
 *   b_dev          P^-1 r from grief_solve_lml
 *   rowmax_dev     NULL, or the row maxima grief_gram_ry recorded for these tables (INT8 arithmetic: the row scales of the digit
 *                  planes then need no sweep of their own; results are identical either way)
 */
int grief_grad_setup(grief_plan* plan, int n_active, const int32_t* dims, const int32_t* kinds, const double* dqs_concat);
size_t grief_grad_workspace_bytes(const grief_plan* plan, int64_t n);
int grief_grad_theta(const grief_plan* plan, const double* T_dev, const double* X_dev, int64_t ldx, const double* y_dev,
                     int64_t n, const double* Pinv_dev, int64_t ldp, const double* b_dev, double noise_var,
                     const int32_t* rowmax_dev, double* grad_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * q[n] = phi(x_n)^T B phi(x_n) for a symmetric (p,p) matrix B, phi built on the fly.
 * With B = P^-1 this is the diagonal of Phi* P^-1 Phi*^T of models/gp_grief_model.py:122-124
 * (predictive variance = noise_var * (q + 1)).
 */
size_t grief_quadform_workspace_bytes(const grief_plan* plan, int64_t n);
int grief_quadform_rows(const grief_plan* plan, const double* T_dev, int64_t n, const double* B_dev, int64_t ldb,
                        double* q_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * Row-shard exchange for hosts without torch.distributed (rows of X shard over the GPUs of one box, one process per
 * GPU; only the packed statistics (A | r | s) -- one contiguous buffer, p^2 + p + 1 doubles -- and the kernel-parameter gradient are
 * all-reduced).  NCCL is loaded with dlopen at the first call: the library has no link-time dependency on it.  The Python layer
 * uses the caller's torch.distributed process group instead (gp_grief_b200/models/gp_grief_model.py: distributed=True).
 *   grief_comm_unique_id      rank 0: 128 bytes to hand to every rank (file, socket, MPI_Bcast, ...)      [ncclGetUniqueId]
 *   grief_comm_create         every rank, with its CUDA device current                                   [ncclCommInitRank]
 *   grief_comm_allreduce_sum  in place, FP64, on the caller's stream; a no-op for world_size 1           [ncclAllReduce]
 */
typedef struct grief_comm grief_comm;
int grief_comm_unique_id(char* id_out128);
int grief_comm_create(grief_comm** comm, const char* id128, int world_size, int rank);
int grief_comm_allreduce_sum(grief_comm* comm, double* buf_dev, int64_t count, void* stream);
void grief_comm_destroy(grief_comm* comm);

/*
 * Row- and column-partitioned Khatri-Rao mat-vec, RowColKhatriRaoMatrix.__mul__ (tensors/khatri_rao_matrix.py:156-167, with get_rows
 * :110-138 fused in): y[i] = sum_j x[j] * prod_t (R_t C_t)[i, j].  No row block of the product is materialised.
 *   d, m[d]        number of factors (<= 32) and their inner sizes
 *   R_dev[d]       host array of DEVICE pointers: R_t (rows, m_t) row-major, or NULL for a factor given by ridx_dev[t]
 *   ridx_dev[d]    host array of DEVICE pointers (or NULL): int32 (rows,), R_t is the selection matrix whose row i picks row
 *                  ridx_t[i] of C_t (SelectionMatrixSparse, tensors/selection_matrix.py:55-107); exactly one of R_dev[t] / ridx_dev[t]
 *   C_dev[d]       C_t (m_t, cols) row-major (K_t C_t when the product has a Kronecker middle factor, khatri_rao_matrix.py:77-82)
 *   x_dev (cols), y_dev (rows)
 */
int grief_rowcol_kr_matvec(int d, const int32_t* m, const double* const* R_dev, const int32_t* const* ridx_dev,
                           const double* const* C_dev, int64_t rows, int64_t cols, const double* x_dev, double* y_dev, void* stream);

/*
 * C (M x N, ldc) = beta * C + alpha * A (M x K, lda) * B (N x K, ldb)^T, all row-major on the device; the TMA-fed FP64 DMMA
 * GEMM both O(n p^2) passes and the dense stage are built on.  Any M, N, K >= 1; A and B need 16-byte aligned rows (even lda,
 * ldb, aligned base).  Stands in for the numpy `dot` calls on
 * p x p and M x p operands (models/gp_grief_model.py:122-124 full predictive covariance).
 */
int grief_gemm_nt(const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double* C_dev, int64_t ldc, int M,
                  int N, int K, double alpha, double beta, void* stream);
/* The same product stored transposed: Ct (N x M, ldct) = beta * Ct + alpha * (A B^T)^T.  The Kronecker mat-vec (tensors/kron_matrix.py:52-97)
 * applies one factor per step to a running matrix and needs the result transposed for the next factor: this saves that pass. */
int grief_gemm_nt_t(const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double* Ct_dev, int64_t ldct, int M,
                    int N, int K, double alpha, double beta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRIEF_B200_H */
