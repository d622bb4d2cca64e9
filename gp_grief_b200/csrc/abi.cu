// extern "C" surface of libgrief_b200.so -- see include/grief_b200.h for the contract.
#include "plan.h"

namespace grief {
const char* last_error_cstr();
struct SolveCtx;
int solve_ctx_create(SolveCtx** out);
void solve_ctx_destroy(SolveCtx* c);
int solve_lml(SolveCtx* ctx, int p, const double* A, int64_t lda, const double* r, const double* yty, const double* w,
              double noise, int64_t n_rows, double* L, double* b, double* Pinv, double* grad_w, double* G2,
              double* scalars_host, int* info_out, cudaStream_t stream, int* launches);
int launch_tables(const Plan* pl, const double* X, int64_t ldx, int64_t n, int64_t n_pad, double* T, cudaStream_t stream);
int launch_phi_rows(const Plan* pl, const double* T, int64_t n, double* Phi, cudaStream_t stream);
int phi_t_vec_blocks(int64_t n);
int launch_phi_t_vec(const Plan* pl, const double* T, int64_t n, const double* v, double* out, double* ws, cudaStream_t stream);
int launch_phi_vec(const Plan* pl, const double* T, int64_t n, const double* v, double* out, cudaStream_t stream);
int launch_sumsq(const double* y, int64_t n, double* out, double* ws, cudaStream_t stream);
size_t gram_workspace_bytes(const Plan* pl, int64_t n_pad, int sms);
int launch_gram(const Plan* pl, const double* T, int64_t n_pad, double* A, int64_t lda, void* workspace, size_t ws_bytes,
                int sms, cudaStream_t stream, int* launches);
int launch_topk(int d, const int32_t* m_host, const double* raw0_host, const double* logeig_host, int p, int32_t* idx_dev,
                double* loglam_dev, int* n_out_host, cudaStream_t stream, int* launches);

static thread_local int g_launches = 0;

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

extern "C" {

int grief_version(void) { return 100; }
const char* grief_last_error(void) { return last_error_cstr(); }
int grief_launch_count(void) { return g_launches; }
void grief_launch_count_reset(void) { g_launches = 0; }

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
    default: return -1;
  }
}
int64_t grief_table_rows(int64_t n) { return (n + kChunk - 1) / kChunk * kChunk; }

int grief_build_tables(const grief_plan* plan, const double* X_dev, int64_t ldx, int64_t n, double* T_dev, void* stream) {
  GRIEF_REQUIRE(plan && T_dev && (X_dev || n == 0), "grief_build_tables: null pointer");
  GRIEF_REQUIRE(n >= 0 && ldx >= plan->impl->d, "grief_build_tables: n=%lld ldx=%lld d=%d", (long long)n, (long long)ldx, plan->impl->d);
  int rc = launch_tables(plan->impl, X_dev, ldx, n, grief_table_rows(n), T_dev, (cudaStream_t)stream);
  if (rc == GRIEF_OK && n > 0) g_launches += 1;
  return rc;
}

int grief_phi_rows(const grief_plan* plan, const double* T_dev, int64_t n, double* Phi_dev, void* stream) {
  GRIEF_REQUIRE(plan && (n == 0 || (T_dev && Phi_dev)), "grief_phi_rows: null pointer");
  int rc = launch_phi_rows(plan->impl, T_dev, n, Phi_dev, (cudaStream_t)stream);
  if (rc == GRIEF_OK && n > 0) g_launches += 1;
  return rc;
}

size_t grief_gram_workspace_bytes(const grief_plan* plan, int64_t n) {
  return gram_workspace_bytes(plan->impl, grief_table_rows(n), sm_count());
}
int grief_gram(const grief_plan* plan, const double* T_dev, int64_t n, double* A_dev, int64_t lda, void* workspace_dev,
               size_t workspace_bytes, void* stream) {
  GRIEF_REQUIRE(plan && A_dev && workspace_dev && (T_dev || n == 0), "grief_gram: null pointer");
  GRIEF_REQUIRE(lda >= plan->impl->p, "grief_gram: lda=%lld < p=%d", (long long)lda, plan->impl->p);
  return launch_gram(plan->impl, T_dev, grief_table_rows(n), A_dev, lda, workspace_dev, workspace_bytes, sm_count(),
                     (cudaStream_t)stream, &g_launches);
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

}  // extern "C"
