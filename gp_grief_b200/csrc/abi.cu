// extern "C" surface of libgrief_b200.so -- see include/grief_b200.h for the contract.
#include "plan.h"

namespace grief {
const char* last_error_cstr();
void prof_enable(bool on);
void prof_read(double* ms, int* count);
struct SolveCtx;
int solve_ctx_create(SolveCtx** out);
void solve_ctx_destroy(SolveCtx* c);
int solve_lml(SolveCtx* ctx, int p, const double* A, int64_t lda, const double* r, const double* yty, const double* w,
              double noise, int64_t n_rows, double* L, double* b, double* Pinv, double* grad_w, double* G2,
              double* scalars_host, int* info_out, cudaStream_t stream, int* launches);
int launch_tables(const Plan* pl, const double* X, int64_t ldx, int64_t n, int64_t n_pad, double* T, cudaStream_t stream, int deriv_dim,
                  const double* Kxu, int64_t ldk);
int launch_phi_rows(const Plan* pl, const double* T, int64_t n, double* Phi, cudaStream_t stream);
int phi_t_vec_blocks(int64_t n);
int launch_phi_t_vec(const Plan* pl, const double* T, int64_t n, const double* v, double* out, double* ws, cudaStream_t stream);
int launch_phi_vec(const Plan* pl, const double* T, int64_t n, const double* v, double* out, cudaStream_t stream);
int launch_sumsq(const double* y, int64_t n, double* out, double* ws, cudaStream_t stream);
size_t gram_workspace_bytes(const Plan* pl, int64_t n_pad, int sms);
int launch_gram(const Plan* pl, const double* T, int64_t n_pad, double* A, int64_t lda, const double* y, int64_t n, double* r,
                int* rowmax_out, void* workspace, size_t ws_bytes, int sms, cudaStream_t stream, int* launches);
int launch_topk(int d, const int32_t* m_host, const double* raw0_host, const double* logeig_host, int p, int32_t* idx_dev,
                double* loglam_dev, int* n_out_host, cudaStream_t stream, int* launches);

int grad_desc_create(GradDesc** out, const Plan* pl, int n_active, const int32_t* dims, const int32_t* kinds, const double* dqs_concat);
int grad_desc_n_active(const GradDesc* gd);
int contract_acc_len(const Plan* pl);
int contract_total_warps(const Plan* pl, int sms);
size_t contract_w_doubles(const Plan* pl, int64_t rows);
size_t kf_doubles(const Plan* pl, int64_t rows);
int launch_contract(const Plan* pl, const double* Zt, int64_t ldz, const double* T, const double* X, int64_t ldx, const double* a,
                    const double* bvec, int64_t rows_blk, int64_t rows_valid, double* Wbuf, double* KFbuf, double* acc, int sms,
                    cudaStream_t stream);
int launch_grad_finish(const Plan* pl, const GradDesc* gd, const double* acc, int sms, double* out, cudaStream_t stream);
int launch_rowdot(const Plan* pl, const double* Zt, int64_t ldz, const double* T, int64_t rows_blk, int64_t rows_valid, double* out,
                  cudaStream_t stream);
int launch_permute_b(const Plan* pl, const double* B, int64_t ldb, double* Bperm, cudaStream_t stream);
int launch_permute_vec(const Plan* pl, const double* in, double scale, double* out, cudaStream_t stream);
size_t zgemm_scratch_bytes(const Plan* pl, int64_t slab_rows);
int launch_zgemm_prepare(const Plan* pl, const double* Bperm, int64_t slab_rows_max, void* scratch, int digits, cudaStream_t stream);
struct ResidualArgs {      // optional by-product of the slab builder: a = (y - Phi bvec) / noise_var for the slab's rows
  const double* bvec = nullptr; const double* y = nullptr; int64_t y_rows = 0; double inv_noise = 0.0; double* a_out = nullptr;
  const int* row_hi = nullptr;
};
int launch_zgemm(const Plan* pl, const double* T_slab, int64_t slab_rows, const double* Bperm, void* scratch, int64_t slab_rows_max, double* Z,
                 int64_t ldz, int digits, const ResidualArgs* res, cudaStream_t stream, int* launches);
struct Comm;
int comm_unique_id(char* id_out);
int comm_create(Comm** out, const char* id, int world, int rank);
int comm_allreduce_sum(Comm* c, double* buf, int64_t count, cudaStream_t stream);
void comm_destroy(Comm* c);
int launch_rowcol_kr_matvec(int d, const int32_t* m, const double* const* R, const int32_t* const* ridx, const double* const* C, int64_t rows,
                            int64_t cols, const double* x, double* y, cudaStream_t stream);

static thread_local int g_launches = 0;

constexpr int64_t kRowBlock = 128;
static int64_t slab_rows_for(int64_t n, int sms) {
  const int64_t n128 = (n + kRowBlock - 1) / kRowBlock * kRowBlock;
  return std::min<int64_t>(n128, (int64_t)sms * kRowBlock * 2);
}
static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

// a plan's buffers live on the device that was current when it was created
#define GRIEF_PLAN_DEVICE(pl_)                                                                                          \
  do {                                                                                                                  \
    int dev__ = -1;                                                                                                     \
    cudaGetDevice(&dev__);                                                                                              \
    GRIEF_REQUIRE(dev__ == (pl_)->device, "plan was created on device %d, the current device is %d", (pl_)->device, dev__); \
  } while (0)

static int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}
}  // namespace grief

using namespace grief;

struct grief_plan { Plan* impl; };
struct grief_ctx { SolveCtx* solve; };

namespace grief {
int gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N, int K,
            double alpha, double beta, bool lower_only, bool store_t, cudaStream_t stream, int* launches, bool tri_k);
}

extern "C" {

int grief_version(void) { return 100; }
const char* grief_last_error(void) { return last_error_cstr(); }
int grief_launch_count(void) { return g_launches; }
void grief_launch_count_reset(void) { g_launches = 0; }

void grief_profile_enable(int on) { prof_enable(on != 0); }
int grief_profile_slots(void) { return PROF_COUNT; }
void grief_profile_read(double* ms_out, int* count_out) { prof_read(ms_out, count_out); }

int grief_ctx_create(grief_ctx** ctx) {
  GRIEF_REQUIRE(ctx != nullptr, "grief_ctx_create: null");
  SolveCtx* s = nullptr;
  int rc = solve_ctx_create(&s);
  if (rc != GRIEF_OK) return rc;
  *ctx = new grief_ctx{s};
  return GRIEF_OK;
}
void grief_ctx_destroy(grief_ctx* ctx) {
  if (!ctx) return;
  solve_ctx_destroy(ctx->solve);
  delete ctx;
}

int grief_topk_kron(int d, const int32_t* m_host, const double* raw0_host, const double* logeig_host, int p,
                    int32_t* idx_dev, double* loglam_dev, int* n_out_host, void* stream) {
  GRIEF_REQUIRE(m_host && raw0_host && logeig_host && idx_dev && loglam_dev, "grief_topk_kron: null pointer");
  return launch_topk(d, m_host, raw0_host, logeig_host, p, idx_dev, loglam_dev, n_out_host, (cudaStream_t)stream, &g_launches);
}

int grief_plan_create(grief_plan** plan, int d, const int32_t* m, const int32_t* kernel_id, const double* variance,
                      const double* lengthscale, const double* grid_concat, const int32_t* u, const double* qs_concat,
                      int p, const int32_t* uinv, int width_cap) {
  GRIEF_REQUIRE(plan && m && kernel_id && variance && lengthscale && grid_concat && u && qs_concat && uinv,
                "grief_plan_create: null pointer");
  Plan* pl = nullptr;
  int rc = plan_create(&pl, d, m, kernel_id, variance, lengthscale, grid_concat, u, qs_concat, p, uinv, width_cap);
  if (rc != GRIEF_OK) return rc;
  *plan = new grief_plan{pl};
  return GRIEF_OK;
}
void grief_plan_destroy(grief_plan* plan) {
  if (!plan) return;
  delete plan->impl;
  delete plan;
}
int grief_plan_info(const grief_plan* plan, int what) {
  if (!plan) return -1;
  const Plan* pl = plan->impl;
  switch (what) {
    case 0: return pl->n_groups;
    case 1: return pl->width;
    case 2: return pl->stride;
    case 3: return pl->p;
    case 4: return pl->p_pad;
    case 5: return pl->d;
    case 6: return pl->sum_m;
    case 7: return pl->n_host_dims;
    default: return -1;
  }
}
int64_t grief_table_rows(int64_t n) { return (n + kRowBlock - 1) / kRowBlock * kRowBlock; }

int grief_build_tables(const grief_plan* plan, const double* X_dev, int64_t ldx, int64_t n, double* T_dev, void* stream) {
  GRIEF_REQUIRE(plan && (n == 0 || (T_dev && X_dev)), "grief_build_tables: null pointer");
  GRIEF_PLAN_DEVICE(plan->impl);
  GRIEF_REQUIRE(n >= 0 && ldx >= plan->impl->d, "grief_build_tables: n=%lld ldx=%lld d=%d", (long long)n, (long long)ldx, plan->impl->d);
  int rc = launch_tables(plan->impl, X_dev, ldx, n, grief_table_rows(n), T_dev, (cudaStream_t)stream, -1, nullptr, 0);
  if (rc == GRIEF_OK && n > 0) g_launches += 1;
  return rc;
}

int grief_build_tables_kxu(const grief_plan* plan, const double* X_dev, int64_t ldx, const double* Kxu_dev, int64_t ldk, int64_t n,
                           int deriv_dim, double* T_dev, void* stream) {
  GRIEF_REQUIRE(plan && (n == 0 || (T_dev && X_dev && Kxu_dev)), "grief_build_tables_kxu: null pointer");
  GRIEF_PLAN_DEVICE(plan->impl);
  GRIEF_REQUIRE(n >= 0 && ldx >= plan->impl->d && ldk >= plan->impl->sum_m, "grief_build_tables_kxu: n=%lld ldx=%lld ldk=%lld (d=%d, sum of grid sizes=%d)",
                (long long)n, (long long)ldx, (long long)ldk, plan->impl->d, plan->impl->sum_m);
  GRIEF_REQUIRE(deriv_dim >= -1 && deriv_dim < plan->impl->d, "grief_build_tables_kxu: deriv_dim=%d outside [-1,%d)", deriv_dim, plan->impl->d);
  int rc = launch_tables(plan->impl, X_dev, ldx, n, grief_table_rows(n), T_dev, (cudaStream_t)stream, deriv_dim, Kxu_dev, ldk);
  if (rc == GRIEF_OK && n > 0) g_launches += 1;
  return rc;
}

int grief_build_tables_dx(const grief_plan* plan, const double* X_dev, int64_t ldx, int64_t n, int dim, double* T_dev, void* stream) {
  GRIEF_REQUIRE(plan && (n == 0 || (T_dev && X_dev)), "grief_build_tables_dx: null pointer");
  GRIEF_PLAN_DEVICE(plan->impl);
  GRIEF_REQUIRE(n >= 0 && ldx >= plan->impl->d, "grief_build_tables_dx: n=%lld ldx=%lld d=%d", (long long)n, (long long)ldx, plan->impl->d);
  GRIEF_REQUIRE(dim >= 0 && dim < plan->impl->d, "grief_build_tables_dx: dim=%d outside [0,%d)", dim, plan->impl->d);
  int rc = launch_tables(plan->impl, X_dev, ldx, n, grief_table_rows(n), T_dev, (cudaStream_t)stream, dim, nullptr, 0);
  if (rc == GRIEF_OK && n > 0) g_launches += 1;
  return rc;
}

int grief_phi_rows(const grief_plan* plan, const double* T_dev, int64_t n, double* Phi_dev, void* stream) {
  GRIEF_REQUIRE(plan && (n == 0 || (T_dev && Phi_dev)), "grief_phi_rows: null pointer");
  int rc = launch_phi_rows(plan->impl, T_dev, n, Phi_dev, (cudaStream_t)stream);
  if (rc == GRIEF_OK && n > 0) g_launches += 1;
  return rc;
}

// ---- options: thread-local defaults for new plans, and per-plan values ----
static int set_opt(PlanOpts& o, int what, int64_t value) {
  switch (what) {
    case GRIEF_OPT_GEMM_MODE:
      GRIEF_REQUIRE(value == 0 || value == 1 || value == 3 || value == 5, "option gemm_mode: %lld is not 0, 1, 3 or 5", (long long)value);
      o.gemm_mode = (int)(value & 1); o.cluster = value == 3 ? 1 : (value == 5 ? 2 : 0);
      return GRIEF_OK;
    case GRIEF_OPT_DIGITS_GRAM:
    case GRIEF_OPT_DIGITS_Z:
    case GRIEF_OPT_DIGITS_VAR:
      GRIEF_REQUIRE(value >= kOzMinDigits && value <= kOzMaxDigits, "option digits: %lld outside [%d,%d]", (long long)value, kOzMinDigits, kOzMaxDigits);
      (what == GRIEF_OPT_DIGITS_GRAM ? o.digits_gram : (what == GRIEF_OPT_DIGITS_Z ? o.digits_z : o.digits_var)) = (int)value;
      return GRIEF_OK;
    case GRIEF_OPT_SLAB_BUDGET:
      GRIEF_REQUIRE(value >= 0, "option slab_budget: %lld", (long long)value);
      o.slab_budget = value ? (size_t)value : ((size_t)5 << 28);
      return GRIEF_OK;
    default: return fail(GRIEF_ERR_BAD_ARG, "unknown option %d", what);
  }
}
static int64_t get_opt(const PlanOpts& o, int what) {
  switch (what) {
    case GRIEF_OPT_GEMM_MODE: return o.gemm_mode | (o.cluster == 1 ? 2 : (o.cluster == 2 ? 4 : 0));
    case GRIEF_OPT_DIGITS_GRAM: return o.digits_gram;
    case GRIEF_OPT_DIGITS_Z: return o.digits_z;
    case GRIEF_OPT_DIGITS_VAR: return o.digits_var;
    case GRIEF_OPT_SLAB_BUDGET: return (int64_t)o.slab_budget;
    default: return -1;
  }
}
int grief_set_default_option(int what, int64_t value) { return set_opt(default_plan_opts(), what, value); }
int64_t grief_get_default_option(int what) { return get_opt(default_plan_opts(), what); }
int grief_plan_set_option(grief_plan* plan, int what, int64_t value) {
  GRIEF_REQUIRE(plan != nullptr, "grief_plan_set_option: null plan");
  return set_opt(plan->impl->opts, what, value);
}
int64_t grief_plan_get_option(const grief_plan* plan, int what) { return plan ? get_opt(plan->impl->opts, what) : -1; }
void grief_set_slab_budget(size_t bytes) { set_opt(default_plan_opts(), GRIEF_OPT_SLAB_BUDGET, (int64_t)bytes); }
void grief_set_gemm_mode(int mode) { set_opt(default_plan_opts(), GRIEF_OPT_GEMM_MODE, mode); }
int grief_get_gemm_mode(void) { return default_plan_opts().gemm_mode; }

size_t grief_gram_workspace_bytes(const grief_plan* plan, int64_t n) {
  return gram_workspace_bytes(plan->impl, grief_table_rows(n), sm_count());
}
int grief_gram(const grief_plan* plan, const double* T_dev, int64_t n, double* A_dev, int64_t lda, void* workspace_dev,
               size_t workspace_bytes, void* stream) {
  return grief_gram_ry(plan, T_dev, n, nullptr, A_dev, lda, nullptr, nullptr, workspace_dev, workspace_bytes, stream);
}
int grief_gram_ry(const grief_plan* plan, const double* T_dev, int64_t n, const double* y_dev, double* A_dev, int64_t lda, double* r_dev,
                  int32_t* rowmax_dev, void* workspace_dev, size_t workspace_bytes, void* stream) {
  GRIEF_REQUIRE(plan && A_dev && workspace_dev && (T_dev || n == 0), "grief_gram: null pointer");
  GRIEF_REQUIRE(r_dev == nullptr || y_dev != nullptr || n == 0, "grief_gram_ry: r needs y");
  GRIEF_REQUIRE(lda >= plan->impl->p, "grief_gram: lda=%lld < p=%d", (long long)lda, plan->impl->p);
  GRIEF_PLAN_DEVICE(plan->impl);
  int rc = launch_gram(plan->impl, T_dev, grief_table_rows(n), A_dev, lda, y_dev, n, r_dev, rowmax_dev, workspace_dev, workspace_bytes, sm_count(),
                       (cudaStream_t)stream, &g_launches);
  if (rc == GRIEF_OK && plan->impl->opts.gemm_mode == 1) rc = ozaki_check(plan->impl->d_err, (cudaStream_t)stream);
  return rc;
}

size_t grief_phi_t_vec_workspace_bytes(const grief_plan* plan, int64_t n) {
  return (size_t)phi_t_vec_blocks(n) * plan->impl->p * sizeof(double);
}
int grief_phi_t_vec(const grief_plan* plan, const double* T_dev, int64_t n, const double* v_dev, double* out_dev,
                    void* workspace_dev, void* stream) {
  GRIEF_REQUIRE(plan && out_dev && workspace_dev, "grief_phi_t_vec: null pointer");
  int rc = launch_phi_t_vec(plan->impl, T_dev, n, v_dev, out_dev, (double*)workspace_dev, (cudaStream_t)stream);
  if (rc == GRIEF_OK) g_launches += 2;
  return rc;
}
int grief_phi_vec(const grief_plan* plan, const double* T_dev, int64_t n, const double* v_dev, double* out_dev, void* stream) {
  GRIEF_REQUIRE(plan && (n == 0 || (T_dev && v_dev && out_dev)), "grief_phi_vec: null pointer");
  int rc = launch_phi_vec(plan->impl, T_dev, n, v_dev, out_dev, (cudaStream_t)stream);
  if (rc == GRIEF_OK && n > 0) g_launches += 1;
  return rc;
}
int grief_sumsq(const double* y_dev, int64_t n, double* out_dev, void* ws_dev, void* stream) {
  GRIEF_REQUIRE(out_dev && ws_dev, "grief_sumsq: null pointer");
  int rc = launch_sumsq(y_dev, n, out_dev, (double*)ws_dev, (cudaStream_t)stream);
  if (rc == GRIEF_OK) g_launches += 2;
  return rc;
}

int grief_solve_lml(grief_ctx* ctx, int p, const double* A_dev, int64_t lda, const double* r_dev, const double* yty_dev,
                    const double* w_dev, double noise_var, int64_t n_rows, double* L_dev, double* b_dev, double* Pinv_dev,
                    double* grad_w_dev, double* G2_dev, double* scalars_host, int* info_host, void* stream) {
  GRIEF_REQUIRE(ctx && A_dev && r_dev && yty_dev && w_dev && L_dev && b_dev && scalars_host, "grief_solve_lml: null pointer");
  return solve_lml(ctx->solve, p, A_dev, lda, r_dev, yty_dev, w_dev, noise_var, n_rows, L_dev, b_dev, Pinv_dev, grad_w_dev,
                   G2_dev, scalars_host, info_host, (cudaStream_t)stream, &g_launches);
}

int grief_grad_setup(grief_plan* plan, int n_active, const int32_t* dims, const int32_t* kinds, const double* dqs_concat) {
  GRIEF_REQUIRE(plan && (n_active == 0 || (dims && kinds && dqs_concat)), "grief_grad_setup: null pointer");
  Plan* pl = plan->impl;
  GRIEF_REQUIRE(pl->n_host_dims == 0 || n_active == 0, "grief_grad_setup: the analytic kernel-parameter gradient needs device kernels in every "
                "dimension (%d dimension(s) are host-evaluated): use finite differences", pl->n_host_dims);
  if (pl->grad) { grad_desc_destroy(pl->grad); pl->grad = nullptr; }
  return grad_desc_create(&pl->grad, pl, n_active, dims, kinds, dqs_concat);
}

// workspace of grief_grad_theta: [Zp^T slab] [GEMM scratch] [W slab] [K_xu | F slab] [a slab] [per-warp accumulators] [b in sorted order]
// [permuted P^-1]
size_t grief_grad_workspace_bytes(const grief_plan* plan, int64_t n) {
  const Plan* pl = plan->impl;
  if (!pl->grad) return 0;
  const int sms = sm_count();
  const int64_t slab = slab_rows_for(n, sms);
  return align256((size_t)slab * pl->p_pad * sizeof(double)) + align256(zgemm_scratch_bytes(pl, slab)) +
         align256(contract_w_doubles(pl, slab) * sizeof(double)) + align256(kf_doubles(pl, slab) * sizeof(double)) +
         align256((size_t)slab * sizeof(double)) +
         align256((size_t)contract_total_warps(pl, sms) * contract_acc_len(pl) * sizeof(double)) + align256((size_t)pl->p_pad * sizeof(double)) +
         align256((size_t)pl->p_pad * pl->p_pad * sizeof(double));
}

int grief_grad_theta(const grief_plan* plan, const double* T_dev, const double* X_dev, int64_t ldx, const double* y_dev, int64_t n,
                     const double* Pinv_dev, int64_t ldp, const double* b_dev, double noise_var, const int32_t* rowmax_dev, double* grad_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream_) {
  GRIEF_REQUIRE(plan && plan->impl->grad, "grief_grad_theta: call grief_grad_setup first");
  GRIEF_REQUIRE(Pinv_dev && b_dev && grad_dev && workspace_dev && (n == 0 || (T_dev && X_dev && y_dev)), "grief_grad_theta: null pointer");
  GRIEF_REQUIRE(workspace_bytes >= grief_grad_workspace_bytes(plan, n), "grief_grad_theta: workspace too small");
  GRIEF_REQUIRE(noise_var > 0.0, "grief_grad_theta: noise_var=%g", noise_var);
  GRIEF_PLAN_DEVICE(plan->impl);
  const Plan* pl = plan->impl;
  const GradDesc* gd = pl->grad;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int sms = sm_count();
  const int na = grad_desc_n_active(gd);
  const int64_t slab = slab_rows_for(n, sms);
  const size_t acc_doubles = (size_t)contract_total_warps(pl, sms) * contract_acc_len(pl);
  char* q = reinterpret_cast<char*>(workspace_dev);
  double* Zt = reinterpret_cast<double*>(q); q += align256((size_t)slab * pl->p_pad * sizeof(double));
  void* zscr = q; q += align256(zgemm_scratch_bytes(pl, slab));
  double* Wbuf = reinterpret_cast<double*>(q); q += align256(contract_w_doubles(pl, slab) * sizeof(double));
  double* KFbuf = reinterpret_cast<double*>(q); q += align256(kf_doubles(pl, slab) * sizeof(double));
  double* avec = reinterpret_cast<double*>(q); q += align256((size_t)slab * sizeof(double));
  double* acc = reinterpret_cast<double*>(q); q += align256(acc_doubles * sizeof(double));
  double* bvec = reinterpret_cast<double*>(q); q += align256((size_t)pl->p_pad * sizeof(double));
  double* Bperm = reinterpret_cast<double*>(q);
  if (na == 0) return GRIEF_OK;
  GRIEF_CUDA(cudaMemsetAsync(acc, 0, acc_doubles * sizeof(double), stream));
  int rc = launch_permute_vec(pl, b_dev, 1.0, bvec, stream);                  // b in sorted column order
  if (rc != GRIEF_OK) return rc;
  rc = launch_permute_b(pl, Pinv_dev, ldp, Bperm, stream);
  if (rc == GRIEF_OK) rc = launch_zgemm_prepare(pl, Bperm, slab, zscr, pl->opts.digits_z, stream);
  if (rc != GRIEF_OK) return rc;
  g_launches += 2;
  const int64_t n128 = grief_table_rows(n);
  for (int64_t r0 = 0; r0 < n128; r0 += slab) {
    const int64_t rows_blk = std::min(slab, n128 - r0);        // multiple of 128, covered by the zero-padded tables
    const int64_t rows_valid = std::max<int64_t>(0, std::min(rows_blk, n - r0));
    ResidualArgs res;                                           // a = (y - Phi b) / noise_var comes out of the slab builder's sweep
    res.bvec = bvec; res.y = y_dev + r0; res.y_rows = rows_valid; res.inv_noise = 1.0 / noise_var; res.a_out = avec;
    res.row_hi = rowmax_dev ? rowmax_dev + r0 : nullptr;
    rc = launch_zgemm(pl, T_dev + (size_t)r0 * pl->stride, rows_blk, Bperm, zscr, slab, Zt, slab, pl->opts.digits_z, &res, stream, &g_launches);
    if (rc != GRIEF_OK) return rc;
    rc = launch_contract(pl, Zt, slab, T_dev + (size_t)r0 * pl->stride, X_dev + (size_t)r0 * ldx, ldx, avec, bvec, rows_blk, rows_valid, Wbuf,
                         KFbuf, acc, sms, stream);
    if (rc != GRIEF_OK) return rc;
    g_launches += 3;
  }
  rc = launch_grad_finish(pl, gd, acc, sms, grad_dev, stream);
  if (rc == GRIEF_OK) g_launches += 1;
  if (rc == GRIEF_OK && pl->opts.gemm_mode == 1) rc = ozaki_check(pl->d_err, stream);
  return rc;
}

size_t grief_quadform_workspace_bytes(const grief_plan* plan, int64_t n) {
  const Plan* pl = plan->impl;
  const int64_t slab = slab_rows_for(n, sm_count());
  return align256((size_t)slab * pl->p_pad * sizeof(double)) + align256(zgemm_scratch_bytes(pl, slab)) + align256((size_t)pl->p_pad * pl->p_pad * sizeof(double));
}

int grief_quadform_rows(const grief_plan* plan, const double* T_dev, int64_t n, const double* B_dev, int64_t ldb, double* q_dev,
                        void* workspace_dev, size_t workspace_bytes, void* stream_) {
  GRIEF_REQUIRE(plan && B_dev && workspace_dev && (n == 0 || (T_dev && q_dev)), "grief_quadform_rows: null pointer");
  GRIEF_REQUIRE(workspace_bytes >= grief_quadform_workspace_bytes(plan, n), "grief_quadform_rows: workspace too small");
  GRIEF_PLAN_DEVICE(plan->impl);
  const Plan* pl = plan->impl;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int sms = sm_count();
  const int64_t slab = slab_rows_for(n, sms);
  double* Z = reinterpret_cast<double*>(workspace_dev);
  void* zscr = reinterpret_cast<char*>(workspace_dev) + align256((size_t)slab * pl->p_pad * sizeof(double));
  double* Bperm = reinterpret_cast<double*>(reinterpret_cast<char*>(zscr) + align256(zgemm_scratch_bytes(pl, slab)));
  {
    int rc0 = launch_permute_b(pl, B_dev, ldb, Bperm, stream);
    if (rc0 == GRIEF_OK) rc0 = launch_zgemm_prepare(pl, Bperm, slab, zscr, pl->opts.digits_var, stream);
    if (rc0 != GRIEF_OK) return rc0;
    g_launches += 1;
  }
  const int64_t n128 = grief_table_rows(n);
  for (int64_t r0 = 0; r0 < n128; r0 += slab) {
    const int64_t rows_blk = std::min(slab, n128 - r0);
    const int64_t rows_valid = std::max<int64_t>(0, std::min(rows_blk, n - r0));
    int rc = launch_zgemm(pl, T_dev + (size_t)r0 * pl->stride, rows_blk, Bperm, zscr, slab, Z, slab, pl->opts.digits_var, nullptr, stream, &g_launches);
    if (rc != GRIEF_OK) return rc;
    rc = launch_rowdot(pl, Z, slab, T_dev + (size_t)r0 * pl->stride, rows_blk, rows_valid, q_dev + r0, stream);
    if (rc != GRIEF_OK) return rc;
    g_launches += 1;
  }
  return pl->opts.gemm_mode == 1 ? ozaki_check(pl->d_err, stream) : GRIEF_OK;
}

int grief_comm_unique_id(char* id_out) {
  GRIEF_REQUIRE(id_out != nullptr, "grief_comm_unique_id: null pointer");
  return comm_unique_id(id_out);
}
int grief_comm_create(grief_comm** comm, const char* id, int world_size, int rank) {
  GRIEF_REQUIRE(comm && id, "grief_comm_create: null pointer");
  Comm* c = nullptr;
  int rc = comm_create(&c, id, world_size, rank);
  if (rc != GRIEF_OK) return rc;
  *comm = reinterpret_cast<grief_comm*>(c);
  return GRIEF_OK;
}
int grief_comm_allreduce_sum(grief_comm* comm, double* buf_dev, int64_t count, void* stream) {
  GRIEF_REQUIRE(comm != nullptr, "grief_comm_allreduce_sum: null communicator");
  return comm_allreduce_sum(reinterpret_cast<Comm*>(comm), buf_dev, count, (cudaStream_t)stream);
}
void grief_comm_destroy(grief_comm* comm) { comm_destroy(reinterpret_cast<Comm*>(comm)); }

int grief_rowcol_kr_matvec(int d, const int32_t* m, const double* const* R_dev, const int32_t* const* ridx_dev, const double* const* C_dev,
                           int64_t rows, int64_t cols, const double* x_dev, double* y_dev, void* stream) {
  GRIEF_REQUIRE(m && C_dev && (R_dev || ridx_dev) && (rows == 0 || cols == 0 || (x_dev && y_dev)), "grief_rowcol_kr_matvec: null pointer");
  int rc = launch_rowcol_kr_matvec(d, m, R_dev, ridx_dev, C_dev, rows, cols, x_dev, y_dev, (cudaStream_t)stream);
  if (rc == GRIEF_OK && rows > 0) g_launches += 1;
  return rc;
}

int grief_gemm_nt(const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double* C_dev, int64_t ldc, int M, int N, int K,
                  double alpha, double beta, void* stream_) {
  GRIEF_REQUIRE(A_dev && B_dev && C_dev, "grief_gemm_nt: null pointer");
  GRIEF_REQUIRE(M >= 0 && N >= 0 && K >= 1 && lda >= K && ldb >= K && ldc >= N, "grief_gemm_nt: M=%d N=%d K=%d lda=%lld ldb=%lld ldc=%lld",
                M, N, K, (long long)lda, (long long)ldb, (long long)ldc);
  return gemm_nt(A_dev, lda, B_dev, ldb, C_dev, ldc, M, N, K, alpha, beta, false, false, (cudaStream_t)stream_, &g_launches, false);
}

int grief_gemm_nt_t(const double* A_dev, int64_t lda, const double* B_dev, int64_t ldb, double* Ct_dev, int64_t ldct, int M, int N, int K,
                    double alpha, double beta, void* stream_) {
  GRIEF_REQUIRE(A_dev && B_dev && Ct_dev, "grief_gemm_nt_t: null pointer");
  GRIEF_REQUIRE(M >= 0 && N >= 0 && K >= 1 && lda >= K && ldb >= K && ldct >= M, "grief_gemm_nt_t: M=%d N=%d K=%d lda=%lld ldb=%lld ldct=%lld",
                M, N, K, (long long)lda, (long long)ldb, (long long)ldct);
  return gemm_nt(A_dev, lda, B_dev, ldb, Ct_dev, ldct, M, N, K, alpha, beta, false, true, (cudaStream_t)stream_, &g_launches, false);
}

}  // extern "C"
