// G3: the p x p stage on device -- P = A + diag(sigma^2/w), Cholesky, b = P^-1 r, log-determinant,
// log marginal likelihood, d/dw and d/dsigma^2, and the pass-2 operand G2 = -(P^-1 + b b^T/sigma^2).
//
// Reference being replaced (models/gp_grief_model.py):
//   :152-153  P, cho_factor            :234  cho_solve(Pchol, Phi^T y)        :243-245 log-det
//   :212-213  LML                      :171-180 d/dw      :185-191 d/dsigma^2
// The reference's three extra passes over Phi and its p-right-hand-side solve are removed with
// the identities of SURVEY.md 7.1 (all follow from A = P - D, D = diag(sigma^2/w)):
//   Phi^T alpha = b / w                      alpha_p = b
//   Y^T alpha   = (s - r^T b)/sigma^2        alpha^T alpha = (s - 2 r^T b + b^T A b)/sigma^4
//   b^T A b     = b^T r - sum_j D_j b_j^2    tr(P^-1 A) = p - sum_j D_j (P^-1)_jj
//   diag(A) - colsum(A o P^-1 A) = D_j (1 - D_j (P^-1)_jj)
// The dense factorisation, the triangular solves and the inverse are hand-written too (dense.cu: blocked right-looking
// Cholesky on a TMA-fed FP64 DMMA GEMM, block forward substitution for L^-1, P^-1 = L^-T L^-1); no cuSOLVER / cuBLAS.
#include "plan.h"

namespace grief {

struct DenseWork;
DenseWork* dense_work_new();
void dense_work_delete(DenseWork* wk);
int dense_work_reserve(DenseWork* wk, int q);
double* dense_work_factor(DenseWork* wk);
double* dense_work_inverse(DenseWork* wk);
double* dense_work_tmp(DenseWork* wk);
int dense_potrf(DenseWork* wk, int q, int* d_info, cudaStream_t stream, int* launches);
int dense_inverse_from_factor(DenseWork* wk, int q, cudaStream_t stream, int* launches);
int dense_trsv_pair(DenseWork* wk, int q, const double* r, int p, double* b, double* tmp, cudaStream_t stream, int* launches);
int launch_form_P_padded(const double* A, int64_t lda, const double* w, double noise, int p, int q, double* P, cudaStream_t stream);
int launch_copy_block(const double* in, int q, int p, bool transpose, bool upper_only, double* out, int64_t ldo, cudaStream_t stream);

struct SolveCtx {
  DenseWork* work = nullptr;
  int* d_info = nullptr;
  double* d_scalars = nullptr;   // SC_COUNT doubles
  ~SolveCtx() {
    if (work) dense_work_delete(work);
    cudaFree(d_info);
    cudaFree(d_scalars);
  }
};

enum { SC_LML = 0, SC_YT_ALPHA, SC_LOGDET, SC_GRAD_NOISE, SC_RTB, SC_ALPHA_SQ, SC_TRACE, SC_COUNT };

int solve_ctx_create(SolveCtx** out) {
  SolveCtx* c = new SolveCtx();
  c->work = dense_work_new();
  if (cudaMalloc(&c->d_info, sizeof(int)) != cudaSuccess || cudaMalloc(&c->d_scalars, SC_COUNT * sizeof(double)) != cudaSuccess) {
    delete c;
    return fail(GRIEF_ERR_CUDA, "solve_ctx_create: cudaMalloc failed");
  }
  *out = c;
  return GRIEF_OK;
}
void solve_ctx_destroy(SolveCtx* c) { delete c; }

// G2 = -(Pinv + b b^T / noise)
__global__ void k_form_G2(const double* __restrict__ Pinv, const double* __restrict__ b, double noise, int p,
                          double* __restrict__ G2) {
  const int64_t total = (int64_t)p * p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / p), j = (int)(e - (int64_t)i * p);
    G2[e] = -(Pinv[e] + b[i] * b[j] / noise);
  }
}

__device__ double block_sum_1024(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    s = warp_sum(s);
  }
  __syncthreads();
  if (threadIdx.x == 0) sh[32] = s;
  __syncthreads();
  return sh[32];
}

// One block of 1024 threads: every O(p) reduction of the likelihood and its w / sigma^2 gradient.
__global__ void __launch_bounds__(1024)
k_assemble(const double* __restrict__ L, const double* __restrict__ Pinv /* may be null */,
           const double* __restrict__ r, const double* __restrict__ b, const double* __restrict__ w,
           const double* __restrict__ yty, double noise, double n_rows, int p, double* __restrict__ grad_w,
           double* __restrict__ scalars) {
  __shared__ double sh[40];
  double s_logL = 0, s_logw = 0, s_rtb = 0, s_db2 = 0, s_dpinv = 0;
  for (int j = threadIdx.x; j < p; j += blockDim.x) {
    const double wj = w[j], bj = b[j], Dj = noise / wj;
    s_logL += log(L[(size_t)j * p + j]);
    s_logw += log(wj);
    s_rtb += r[j] * bj;
    s_db2 += Dj * bj * bj;
    if (Pinv) {
      const double pjj = Pinv[(size_t)j * p + j];
      s_dpinv += Dj * pjj;
      if (grad_w) {
        const double t = bj / wj;
        grad_w[j] = 0.5 * t * t - 0.5 * (1.0 - Dj * pjj) / wj;
      }
    }
  }
  const double logL = block_sum_1024(s_logL, sh);
  const double logw = block_sum_1024(s_logw, sh);
  const double rtb = block_sum_1024(s_rtb, sh);
  const double db2 = block_sum_1024(s_db2, sh);
  const double dpinv = block_sum_1024(s_dpinv, sh);
  if (threadIdx.x == 0) {
    const double s = yty[0];
    const double yta = (s - rtb) / noise;
    const double logdet = 2.0 * logL + logw + (n_rows - (double)p) * log(noise);
    scalars[SC_LML] = -0.5 * (yta + logdet + n_rows * log(2.0 * 3.14159265358979323846));
    scalars[SC_YT_ALPHA] = yta;
    scalars[SC_LOGDET] = logdet;
    scalars[SC_RTB] = rtb;
    const double bAb = rtb - db2;
    const double alpha_sq = (s - 2.0 * rtb + bAb) / (noise * noise);
    const double trace = (double)p - dpinv;
    scalars[SC_ALPHA_SQ] = alpha_sq;
    scalars[SC_TRACE] = trace;
    scalars[SC_GRAD_NOISE] = Pinv ? 0.5 * alpha_sq - 0.5 * (n_rows - trace) / noise : 0.0;
  }
}

// Returns GRIEF_ERR_NOT_PD with the failing leading-minor order in *info_out when P is not PD.
// L_out receives the factor in scipy's cho_factor convention for a row-major array: upper U with P = U^T U (lower part 0).
int solve_lml(SolveCtx* ctx, int p, const double* A, int64_t lda, const double* r, const double* yty,
              const double* w, double noise, int64_t n_rows, double* L /* p*p */, double* b /* p */,
              double* Pinv /* p*p or null */, double* grad_w /* p or null */, double* G2 /* p*p or null */,
              double* scalars_host /* SC_COUNT */, int* info_out, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(p >= 1, "solve_lml: p=%d", p);
  GRIEF_REQUIRE(noise > 0.0, "solve_lml: noise_var=%g must be positive", noise);
  GRIEF_REQUIRE(G2 == nullptr || Pinv != nullptr, "solve_lml: G2 needs Pinv");
  const int q = (p + 127) / 128 * 128;               // P is padded with an identity block to a multiple of the tile
  int rc = dense_work_reserve(ctx->work, q);
  if (rc != GRIEF_OK) return rc;
  const unsigned eb = (unsigned)std::min<int64_t>(((int64_t)p * p + 255) / 256, 148 * 8);
  prof_begin(PROF_SOLVE, stream);
  GRIEF_CUDA(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), stream));
  rc = launch_form_P_padded(A, lda, w, noise, p, q, dense_work_factor(ctx->work), stream);
  if (rc == GRIEF_OK) rc = dense_potrf(ctx->work, q, ctx->d_info, stream, launches);
  if (rc != GRIEF_OK) return rc;
  int info = 0;
  GRIEF_CUDA(cudaMemcpyAsync(&info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, stream));
  GRIEF_CUDA(cudaStreamSynchronize(stream));
  if (info_out) *info_out = info;
  if (info != 0) return fail(GRIEF_ERR_NOT_PD, "Cholesky failed: leading minor of order %d of P = A + diag(noise_var/w) is not positive definite", info);
  rc = launch_copy_block(dense_work_factor(ctx->work), q, p, true, true, L, p, stream);
  if (rc == GRIEF_OK) rc = dense_trsv_pair(ctx->work, q, r, p, b, dense_work_tmp(ctx->work), stream, launches);
  if (rc != GRIEF_OK) return rc;
  if (Pinv) {
    rc = dense_inverse_from_factor(ctx->work, q, stream, launches);
    if (rc == GRIEF_OK) rc = launch_copy_block(dense_work_inverse(ctx->work), q, p, false, false, Pinv, p, stream);
    if (rc != GRIEF_OK) return rc;
  }
  k_assemble<<<1, 1024, 0, stream>>>(L, Pinv, r, b, w, yty, noise, (double)n_rows, p, grad_w, ctx->d_scalars);
  GRIEF_CUDA(cudaGetLastError());
  if (G2) {
    k_form_G2<<<eb, 256, 0, stream>>>(Pinv, b, noise, p, G2);
    GRIEF_CUDA(cudaGetLastError());
  }
  prof_end(PROF_SOLVE, stream);
  GRIEF_CUDA(cudaMemcpyAsync(scalars_host, ctx->d_scalars, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, stream));
  GRIEF_CUDA(cudaStreamSynchronize(stream));
  if (launches) *launches += 4 + (Pinv ? 1 : 0) + (G2 ? 1 : 0);
  return GRIEF_OK;
}

}  // namespace grief
