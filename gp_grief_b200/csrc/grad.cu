// G4: analytic hyper-parameter gradient (replaces the reference's finite-difference loop,
// models/basemodel.py:328-361, which costs (#free + 1) full Phi / Gram rebuilds).
//
//   dLML/dtheta = sum_n sum_j Zt[n,j] * dPhi[n,j]/dtheta,    Zt = Phi * G2 + y g^T,
//   G2 = -(P^-1 + b b^T/sigma^2),  g = b/sigma^2                                
//
// Z = Phi*G2 comes from launch_zgemm (phi_stage.cu: slab builder + GEMM), one slab of rows at a time, with its columns in the
// plan's SORTED order (so do gvec and the slot table handed to k_contract / k_rowdot).  For a parameter theta of
// input dimension i (group g of the table layout) only the factor of dimension i changes:
//   dPhi[n,j]/dtheta = DT_theta[n, t_g(j)] * prod_{g' != g} H_g'[n, t_g'(j)]
//   DT_theta[n,t]    = dF_i[n,k_i(t)]/dtheta * prod_{i' in g, i' != i} F_i'[n,k_i'(t)]
//   dF_i[n,k]/dtheta = sum_u dK_i(x_n,u)/dtheta * Qs_i[u,k] + K_i(x_n,u) * dQs_i[u,k]/dtheta
// with dQs/dtheta (eigenvector / eigenvalue perturbation of the m_i x m_i grid problem) supplied by the host.
// k_dtables evaluates DT for a slab of rows; k_contract folds Z, H and DT into per-parameter sums.
#include <vector>

#include "plan.h"

namespace grief {

struct GradDesc {            // device-resident description of the active parameters
  int n_active = 0;
  int dt_width = 0;          // doubles per DT row
  int sum_du = 0;            // sum over active params of u_dim
  int max_np = 0;
  int* d_a_dim = nullptr;    // [n_active]
  int* d_a_kind = nullptr;   // 0 = variance, 1 = lengthscale
  int* d_a_group = nullptr;
  int* d_a_kk = nullptr;     // index among the active params of its group
  int* d_a_qoff = nullptr;   // offset into d_dqs
  int* d_a_foff = nullptr;   // offset into the per-row dF scratch
  int* d_g_np = nullptr;     // [G] active params per group
  int* d_g_dtoff = nullptr;  // [G+1] offset of the group's block inside a DT row
  int* d_g_slot0 = nullptr;  // [G]
  int* d_g_size = nullptr;   // [G]
  int* d_ga = nullptr;       // [G][max_np] -> active index
  double* d_dqs = nullptr;
  std::vector<int> g_np_h, g_dtoff_h;
  ~GradDesc() {
    cudaFree(d_a_dim); cudaFree(d_a_kind); cudaFree(d_a_group); cudaFree(d_a_kk); cudaFree(d_a_qoff); cudaFree(d_a_foff);
    cudaFree(d_g_np); cudaFree(d_g_dtoff); cudaFree(d_g_slot0); cudaFree(d_g_size); cudaFree(d_ga); cudaFree(d_dqs);
  }
};

template <typename T>
static int upload_vec(T** dst, const std::vector<T>& src) {
  size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
  GRIEF_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), bytes));
  if (!src.empty()) GRIEF_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return GRIEF_OK;
}

// dims[a], kinds[a] (0 variance, 1 lengthscale); dqs_concat: per active param an (m_i x u_i) row-major matrix.
int grad_desc_create(GradDesc** out, const Plan* pl, int n_active, const int32_t* dims, const int32_t* kinds,
                     const double* dqs_concat) {
  GRIEF_REQUIRE(n_active >= 0 && n_active <= 2 * kMaxDims, "grad_desc: n_active=%d", n_active);
  GradDesc* gd = new GradDesc();
  gd->n_active = n_active;
  const int G = pl->n_groups;
  std::vector<int> a_dim(n_active), a_kind(n_active), a_group(n_active), a_kk(n_active), a_qoff(n_active), a_foff(n_active);
  std::vector<int> g_np(G, 0);
  int qoff = 0, foff = 0;
  for (int a = 0; a < n_active; ++a) {
    const int i = dims[a];
    if (i < 0 || i >= pl->d || kinds[a] < 0 || kinds[a] > 1) {
      delete gd;
      return fail(GRIEF_ERR_BAD_ARG, "grad_desc: parameter %d has dim=%d kind=%d", a, i, kinds[a]);
    }
    int g = 0;
    while (!(pl->group_begin[g] <= i && i < pl->group_begin[g + 1])) ++g;
    a_dim[a] = i; a_kind[a] = kinds[a]; a_group[a] = g; a_kk[a] = g_np[g]++;
    a_qoff[a] = qoff; qoff += pl->dims[i].m * pl->dims[i].u;
    a_foff[a] = foff; foff += pl->dims[i].u;
  }
  gd->sum_du = foff;
  std::vector<int> g_dtoff(G + 1, 0);
  int max_np = 1;
  for (int g = 0; g < G; ++g) {
    g_dtoff[g + 1] = g_dtoff[g] + pl->group_size[g] * g_np[g];
    max_np = std::max(max_np, g_np[g]);
  }
  gd->dt_width = g_dtoff[G];
  gd->max_np = max_np;
  std::vector<int> ga((size_t)G * max_np, -1);
  for (int a = 0; a < n_active; ++a) ga[(size_t)a_group[a] * max_np + a_kk[a]] = a;
  gd->g_np_h = g_np;
  gd->g_dtoff_h = g_dtoff;
  std::vector<double> dqs(dqs_concat, dqs_concat + qoff);
  int rc = upload_vec(&gd->d_a_dim, a_dim);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_a_kind, a_kind);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_a_group, a_group);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_a_kk, a_kk);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_a_qoff, a_qoff);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_a_foff, a_foff);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_g_np, g_np);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_g_dtoff, g_dtoff);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_g_slot0, pl->group_slot0);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_g_size, pl->group_size);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_ga, ga);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_dqs, dqs);
  if (rc != GRIEF_OK) { delete gd; return rc; }
  *out = gd;
  return GRIEF_OK;
}
void grad_desc_destroy(GradDesc* gd) { delete gd; }
int grad_desc_n_active(const GradDesc* gd) { return gd->n_active; }
int grad_desc_dt_width(const GradDesc* gd) { return gd->dt_width; }

__global__ void k_scale_vec(const double* __restrict__ in, double scale, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * scale;
}
int launch_scale_vec(const double* in, double scale, int n, double* out, cudaStream_t stream) {
  k_scale_vec<<<(n + 255) / 256, 256, 0, stream>>>(in, scale, n, out);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// kernel value and its lengthscale derivative (variance derivative is K / variance)
__device__ __forceinline__ void kern_eval_d(int kernel, double x, double u, double variance, double ls, double& k,
                                            double& dk_dls) {
  const double diff = x - u;
  const double d2 = diff * diff;
  switch (kernel) {
    case KERN_RBF: {
      if (ls < 1e-6) { k = d2 == 0.0 ? variance : 0.0; dk_dls = 0.0; return; }
      k = variance * exp(-0.5 * d2 / (ls * ls));
      dk_dls = k * d2 / (ls * ls * ls);
      return;
    }
    case KERN_EXPONENTIAL: {
      const double r = sqrt(d2) / ls;
      k = variance * exp(-r);
      dk_dls = k * r / ls;
      return;
    }
    case KERN_MATERN32: {
      const double s3 = 1.7320508075688772;
      const double r = sqrt(d2) / ls;
      const double e = exp(-s3 * r);
      k = variance * (1.0 + s3 * r) * e;
      dk_dls = variance * 3.0 * r * r * e / ls;
      return;
    }
    default: {
      const double s5 = 2.23606797749979;
      const double r2 = d2 / (ls * ls);
      const double r = sqrt(r2);
      const double e = exp(-s5 * r);
      k = variance * (1.0 + s5 * r + (5.0 / 3.0) * r2) * e;
      dk_dls = variance * e * (5.0 / 3.0) * r2 * (1.0 + s5 * r) / ls;
      return;
    }
  }
}

struct DtParams {
  const DimDesc* dims; const double* grid; const double* qs; const uint8_t* slot_k; const int* group_begin;
  int d, sum_m, sum_u, max_group_dims, G;
  // gradient description
  int n_active, dt_width, sum_du, max_np;
  const int *a_dim, *a_kind, *a_qoff, *a_foff, *g_np, *g_dtoff, *g_slot0, *ga;
  const double* dqs;
  const double* X; int64_t ldx; int64_t n_valid;   // rows >= n_valid produce zeros
  double* DT; int RB;
};

// One block = RB rows.  smem: sK[RB][sum_m], sKl[RB][sum_m], sF[RB][sum_u], sdF[RB][sum_du].
__global__ void __launch_bounds__(256) k_dtables(const DtParams P, int64_t rows_total) {
  extern __shared__ double sm[];
  double* sK = sm;
  double* sKl = sK + (size_t)P.RB * P.sum_m;
  double* sF = sKl + (size_t)P.RB * P.sum_m;
  double* sdF = sF + (size_t)P.RB * P.sum_u;
  const int64_t row0 = (int64_t)blockIdx.x * P.RB;
  const int rows = (int)min((int64_t)P.RB, rows_total - row0);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int task = tid; task < rows * P.sum_m; task += nt) {
    const int r = task / P.sum_m, c = task - r * P.sum_m;
    int i = 0;
    while (i + 1 < P.d && P.dims[i + 1].grid_off <= c) ++i;
    const DimDesc dd = P.dims[i];
    const int64_t row = row0 + r;
    double k = 0.0, dk = 0.0;
    if (row < P.n_valid) kern_eval_d(dd.kernel, P.X[row * P.ldx + i], P.grid[c], dd.variance, dd.lengthscale, k, dk);
    sK[(size_t)r * P.sum_m + c] = k;
    sKl[(size_t)r * P.sum_m + c] = dk;
  }
  __syncthreads();
  for (int task = tid; task < rows * P.sum_u; task += nt) {
    const int r = task / P.sum_u, c = task - r * P.sum_u;
    int i = 0;
    while (i + 1 < P.d && P.dims[i + 1].f_off <= c) ++i;
    const DimDesc dd = P.dims[i];
    const int k = c - dd.f_off;
    const double* kv = sK + (size_t)r * P.sum_m + dd.grid_off;
    const double* q = P.qs + dd.q_off + k;
    double acc = 0.0;
    for (int g = 0; g < dd.m; ++g) acc = fma(kv[g], q[(size_t)g * dd.u], acc);
    sF[(size_t)r * P.sum_u + c] = acc;
  }
  for (int task = tid; task < rows * P.sum_du; task += nt) {
    const int r = task / P.sum_du, c = task - r * P.sum_du;
    int a = 0;
    while (a + 1 < P.n_active && P.a_foff[a + 1] <= c) ++a;
    const DimDesc dd = P.dims[P.a_dim[a]];
    const int k = c - P.a_foff[a];
    const double* kv = sK + (size_t)r * P.sum_m + dd.grid_off;
    const double* kl = sKl + (size_t)r * P.sum_m + dd.grid_off;
    const double* q = P.qs + dd.q_off + k;
    const double* dq = P.dqs + P.a_qoff[a] + k;
    const bool is_ls = P.a_kind[a] == 1;
    const double inv_var = 1.0 / dd.variance;
    double acc = 0.0;
    for (int g = 0; g < dd.m; ++g) {
      const double dk = is_ls ? kl[g] : kv[g] * inv_var;
      acc = fma(dk, q[(size_t)g * dd.u], acc);
      acc = fma(kv[g], dq[(size_t)g * dd.u], acc);
    }
    sdF[(size_t)r * P.sum_du + c] = acc;
  }
  __syncthreads();
  for (int task = tid; task < rows * P.dt_width; task += nt) {
    const int r = task / P.dt_width, c = task - r * P.dt_width;
    int g = 0;
    while (g + 1 < P.G && P.g_dtoff[g + 1] <= c) ++g;
    const int np = P.g_np[g];
    const int local = c - P.g_dtoff[g];
    const int tl = local / np, kk = local - tl * np;
    const int a = P.ga[(size_t)g * P.max_np + kk];
    const int ia = P.a_dim[a];
    const int s = P.g_slot0[g] + tl;
    const uint8_t* ks = P.slot_k + (size_t)s * P.max_group_dims;
    const int b0 = P.group_begin[g], b1 = P.group_begin[g + 1];
    double v = sdF[(size_t)r * P.sum_du + P.a_foff[a] + ks[ia - b0]];
    for (int i = b0; i < b1; ++i)
      if (i != ia) v *= sF[(size_t)r * P.sum_u + P.dims[i].f_off + ks[i - b0]];
    P.DT[(row0 + r) * (int64_t)P.dt_width + c] = v;
  }
}

struct ContractParams {
  const double* Z; int64_t ldz;      // slab rows x ldz
  const double* T; int stride;       // slab rows x stride
  const double* DT; int dt_width;    // slab rows x dt_width
  const double* y;                   // slab rows
  const double* gvec;                // p: b / sigma^2
  const uint16_t* col_slot;          // p_pad x G
  const int *g_np, *g_dtoff, *g_slot0;
  int p, max_np;
  int64_t rows;                      // valid rows in this slab
  double* partial;                   // [gridDim.x][n_active], accumulated (+=) across slabs
  const int* ga; int n_active;
};

constexpr int kContractK = 4;   // parameters per group accumulated in registers per sweep (measured at C3: 4 -> 36.2, 6 -> 38.4, 8 -> 41.5 ms per 1M rows)

// One warp per data row, ONE sweep over the columns for all groups (a second sweep only if a group has more than
// kContractK active parameters): per column the leave-one-group-out products L_g = Zt * prod_{g' != g} H_g' come from
// prefix / suffix products, and L_g * DT[g][t_g(j)][kk] is accumulated per (group, parameter) in registers.
template <int G>
__global__ void __launch_bounds__(256) k_contract(const ContractParams P) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double* sRow = sm + (size_t)warp * (P.stride + P.dt_width);          // [stride] table row, [dt_width] DT row
  double* sAcc = sm + (size_t)nw * (P.stride + P.dt_width) + (size_t)warp * P.n_active;
  for (int a = lane; a < P.n_active; a += 32) sAcc[a] = 0.0;
  int np[G], slot0[G], dtoff[G];
#pragma unroll
  for (int g = 0; g < G; ++g) { np[g] = P.g_np[g]; slot0[g] = P.g_slot0[g]; dtoff[g] = P.g_dtoff[g]; }
  for (int64_t row = (int64_t)blockIdx.x * nw + warp; row < P.rows; row += (int64_t)gridDim.x * nw) {
    __syncwarp();
    for (int e = lane; e < P.stride; e += 32) sRow[e] = P.T[row * P.stride + e];
    for (int e = lane; e < P.dt_width; e += 32) sRow[P.stride + e] = P.DT[row * (int64_t)P.dt_width + e];
    __syncwarp();
    const double yr = P.y[row];
    const double* zrow = P.Z + row * P.ldz;
    const double* sDT = sRow + P.stride;
    for (int k0 = 0; k0 < P.max_np; k0 += kContractK) {
      double acc[G][kContractK];
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int kk = 0; kk < kContractK; ++kk) acc[g][kk] = 0.0;
      for (int j = lane; j < P.p; j += 32) {
        const double zt = fma(yr, P.gvec[j], zrow[j]);
        const uint16_t* cs = P.col_slot + (size_t)j * G;
        int sl[G];
        double h[G], L[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { sl[g] = cs[g]; h[g] = sRow[sl[g]]; }
        double run = zt;                       // prefix pass: L[g] = zt * prod_{g' < g} h
#pragma unroll
        for (int g = 0; g < G; ++g) { L[g] = run; run *= h[g]; }
        run = 1.0;                             // suffix pass: L[g] *= prod_{g' > g} h
#pragma unroll
        for (int g = G - 1; g >= 0; --g) { L[g] *= run; run *= h[g]; }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const double* dt = sDT + dtoff[g] + (size_t)(sl[g] - slot0[g]) * np[g] + k0;
#pragma unroll
          for (int kk = 0; kk < kContractK; ++kk)
            if (k0 + kk < np[g]) acc[g][kk] = fma(L[g], dt[kk], acc[g][kk]);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int kk = 0; kk < kContractK; ++kk)
          if (k0 + kk < np[g]) {               // warp-uniform
            const double v = warp_sum(acc[g][kk]);
            if (lane == 0) sAcc[P.ga[(size_t)g * P.max_np + k0 + kk]] += v;
          }
    }
  }
  __syncthreads();
  // block partial = sum over warps (fixed order)
  for (int a = threadIdx.x; a < P.n_active; a += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += sm[(size_t)nw * (P.stride + P.dt_width) + (size_t)w * P.n_active + a];
    P.partial[(size_t)blockIdx.x * P.n_active + a] += s;
  }
}

__global__ void k_reduce_partials(const double* __restrict__ partial, int nblk, int n_active, double* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_active) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * n_active + a];
  out[a] = s;
}

// out[n] = sum_j Z[n,j] * Phi[n,j]   (diag of Phi* P^-1 Phi*^T for the predictive variance)
template <int G>
__global__ void __launch_bounds__(256) k_rowdot(const double* __restrict__ Z, int64_t ldz, const double* __restrict__ T, int stride,
                                                const uint16_t* __restrict__ col_slot, int p, int64_t rows, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < rows; row += nwarps) {
    const double* t = T + row * stride;
    const double* z = Z + row * ldz;
    double acc = 0.0;
    for (int j = lane; j < p; j += 32) {
      double ph = t[col_slot[(size_t)j * G]];
#pragma unroll
      for (int g = 1; g < G; ++g) ph *= t[col_slot[(size_t)j * G + g]];
      acc = fma(ph, z[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
  }
}

int launch_dtables(const Plan* pl, const GradDesc* gd, const double* X, int64_t ldx, int64_t n_valid, int64_t rows_total,
                   double* DT, cudaStream_t stream) {
  if (rows_total == 0 || gd->dt_width == 0) return GRIEF_OK;
  const size_t per_row = (size_t)(2 * pl->sum_m + pl->sum_u + gd->sum_du) * sizeof(double);
  int RB = (int)std::min<size_t>(32, (160 * 1024) / per_row);
  if (RB < 1) return fail(GRIEF_ERR_UNSUPPORTED, "dtables: per-row scratch of %zu bytes exceeds shared memory", per_row);
  DtParams P;
  P.dims = pl->d_dims; P.grid = pl->d_grid; P.qs = pl->d_qs; P.slot_k = pl->d_slot_k; P.group_begin = pl->d_group_begin;
  P.d = pl->d; P.sum_m = pl->sum_m; P.sum_u = pl->sum_u; P.max_group_dims = pl->max_group_dims; P.G = pl->n_groups;
  P.n_active = gd->n_active; P.dt_width = gd->dt_width; P.sum_du = gd->sum_du; P.max_np = gd->max_np;
  P.a_dim = gd->d_a_dim; P.a_kind = gd->d_a_kind; P.a_qoff = gd->d_a_qoff; P.a_foff = gd->d_a_foff;
  P.g_np = gd->d_g_np; P.g_dtoff = gd->d_g_dtoff; P.g_slot0 = gd->d_g_slot0; P.ga = gd->d_ga; P.dqs = gd->d_dqs;
  P.X = X; P.ldx = ldx; P.n_valid = n_valid; P.DT = DT; P.RB = RB;
  const size_t smem = per_row * RB;
  GRIEF_CUDA(cudaFuncSetAttribute(k_dtables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (rows_total + RB - 1) / RB;
  prof_begin(PROF_DTABLES, stream);
  k_dtables<<<(unsigned)blocks, 256, smem, stream>>>(P, rows_total);
  prof_end(PROF_DTABLES, stream);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

int contract_blocks(int sms) { return sms * 2; }

template <int G>
static int launch_contract_g(const ContractParams& P, int blocks, size_t smem, cudaStream_t stream) {
  GRIEF_CUDA(cudaFuncSetAttribute(k_contract<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_contract<G><<<blocks, 256, smem, stream>>>(P);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

int launch_contract(const Plan* pl, const GradDesc* gd, const double* Z, int64_t ldz, const double* T, const double* DT,
                    const double* y, const double* gvec, int64_t rows, double* partial, int sms, cudaStream_t stream) {
  if (rows == 0 || gd->n_active == 0) return GRIEF_OK;
  ContractParams P;
  P.Z = Z; P.ldz = ldz; P.T = T; P.stride = pl->stride; P.DT = DT; P.dt_width = gd->dt_width; P.y = y; P.gvec = gvec;
  P.col_slot = pl->d_sorted_gslot; P.g_np = gd->d_g_np; P.g_dtoff = gd->d_g_dtoff; P.g_slot0 = gd->d_g_slot0;
  // Z, gvec and the slot table are in the sorted column order (padding columns contribute 0)
  P.p = pl->p_pad; P.max_np = gd->max_np; P.rows = rows; P.partial = partial; P.ga = gd->d_ga; P.n_active = gd->n_active;
  const int nw = 8;
  const size_t smem = ((size_t)nw * (pl->stride + gd->dt_width) + (size_t)nw * gd->n_active) * sizeof(double);
  if (smem > 200 * 1024) return fail(GRIEF_ERR_UNSUPPORTED, "contract: %zu bytes of shared memory per block", smem);
  const int blocks = contract_blocks(sms);
  int rc;
  prof_begin(PROF_CONTRACT, stream);
  switch (pl->n_groups) {
    case 1: rc = launch_contract_g<1>(P, blocks, smem, stream); break;
    case 2: rc = launch_contract_g<2>(P, blocks, smem, stream); break;
    case 3: rc = launch_contract_g<3>(P, blocks, smem, stream); break;
    case 4: rc = launch_contract_g<4>(P, blocks, smem, stream); break;
    case 5: rc = launch_contract_g<5>(P, blocks, smem, stream); break;
    case 6: rc = launch_contract_g<6>(P, blocks, smem, stream); break;
    case 7: rc = launch_contract_g<7>(P, blocks, smem, stream); break;
    case 8: rc = launch_contract_g<8>(P, blocks, smem, stream); break;
    default: return fail(GRIEF_ERR_UNSUPPORTED, "contract: %d groups", pl->n_groups);
  }
  prof_end(PROF_CONTRACT, stream);
  return rc;
}

int launch_reduce_partials(const double* partial, int nblk, int n_active, double* out, cudaStream_t stream) {
  if (n_active == 0) return GRIEF_OK;
  k_reduce_partials<<<(n_active + 127) / 128, 128, 0, stream>>>(partial, nblk, n_active, out);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

int launch_rowdot(const Plan* pl, const double* Z, int64_t ldz, const double* T, int64_t rows, double* out, cudaStream_t stream) {
  if (rows == 0) return GRIEF_OK;
  const unsigned blocks = (unsigned)std::min<int64_t>((rows + 7) / 8, 148 * 8);
#define RD(Gv) k_rowdot<Gv><<<blocks, 256, 0, stream>>>(Z, ldz, T, pl->stride, pl->d_sorted_gslot, pl->p_pad, rows, out)
  switch (pl->n_groups) {
    case 1: RD(1); break; case 2: RD(2); break; case 3: RD(3); break; case 4: RD(4); break;
    case 5: RD(5); break; case 6: RD(6); break; case 7: RD(7); break; case 8: RD(8); break;
    default: return fail(GRIEF_ERR_UNSUPPORTED, "rowdot: %d groups", pl->n_groups);
  }
#undef RD
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

}  // namespace grief
