// Row kernels (HBM-bound): the group-table prepass, Phi materialisation for small n,
// and the two Phi mat-vecs.  Reference semantics:
//   K_xu,i[n,g]   = k_i(x[n,i], U_i[g])                      kern/grid_kernel.py:171, kern/stationary.py:126-127
//   F_i[n,k]      = sum_g K_xu,i[n,g] * Qs_i[g,k]            tensors/tensors.py:118 (with the lambda^-1/2 of
//                                                             kern/grief_kernel.py:104 folded into Qs per dimension)
//   Phi[n,j]      = prod_i F_i[n, uinv[j,i]]                 tensors/tensors.py:116-124
// The table row of data row n holds, for every group of dimensions, the product of the F_i over
// the group's dimensions for each distinct index sub-tuple (plan.h).
#include "plan.h"

namespace grief {

__device__ __forceinline__ double kern_eval(int kernel, double x, double u, double variance, double lengthscale) {
  const double diff = x - u;
  const double d2 = diff * diff;
  switch (kernel) {
    case KERN_RBF: {
      if (lengthscale < 1e-6) return d2 == 0.0 ? variance : 0.0;   // kern/stationary.py:121-122
      return variance * exp(-0.5 * d2 / (lengthscale * lengthscale));
    }
    case KERN_EXPONENTIAL: {
      const double r = sqrt(d2) / lengthscale;
      return variance * exp(-r);
    }
    case KERN_MATERN32: {
      const double r = sqrt(d2) / lengthscale;
      const double s3 = 1.7320508075688772;
      return variance * (1.0 + s3 * r) * exp(-s3 * r);
    }
    default: {  // KERN_MATERN52
      const double r2 = d2 / (lengthscale * lengthscale);
      const double r = sqrt(r2);
      const double s5 = 2.23606797749979;
      return variance * (1.0 + s5 * r + (5.0 / 3.0) * r2) * exp(-s5 * r);
    }
  }
}

// d k(x, u) / d x  (kern/stationary.py grad_x of the in-house kernels; the reference needs GPy's gradients_X for this,
// kern/grid_kernel.py:196-199)
__device__ __forceinline__ double kern_eval_dx(int kernel, double x, double u, double variance, double lengthscale) {
  const double diff = x - u;
  const double l2 = lengthscale * lengthscale;
  switch (kernel) {
    case KERN_RBF:
      if (lengthscale < 1e-6) return 0.0;
      return -variance * exp(-0.5 * diff * diff / l2) * diff / l2;
    case KERN_EXPONENTIAL: {
      const double sgn = diff > 0.0 ? 1.0 : (diff < 0.0 ? -1.0 : 0.0);
      return -variance * exp(-fabs(diff) / lengthscale) * sgn / lengthscale;
    }
    case KERN_MATERN32: {
      const double s3 = 1.7320508075688772;
      return -variance * 3.0 * diff * exp(-s3 * fabs(diff) / lengthscale) / l2;
    }
    default: {  // KERN_MATERN52
      const double s5 = 2.23606797749979;
      const double r = fabs(diff) / lengthscale;
      return -variance * (5.0 / 3.0) * diff * (1.0 + s5 * r) * exp(-s5 * r) / l2;
    }
  }
}

// One block = RB data rows.  smem: sK[RB][sum_m], sF[RB][sum_u].
// deriv_dim >= 0: the kernel of that input dimension is replaced by its derivative with respect to x, so that the
// table products become d Phi / d x[:, deriv_dim] (models/gp_grief_model.py:127-134, kern/grief_kernel.py:113-126).
// Kxu != nullptr: (n x sum_m) row-major, ldk doubles per row; the columns of every dimension whose kernel id is KERN_HOST hold
// K_xu,i evaluated by the caller (any BaseKernel whose cov is host code: kern/gpy_kernel.py:45-58, kern/grid_kernel.py:148-179);
// with deriv_dim = i they hold d K_xu,i / d x instead.
// KF != nullptr (RB must be 32): the kernel values and the factors of the block's rows are also written as
// KF[block][c][row in block], c < sum_m: K_xu, then sum_u columns of F -- the layout the lane = row gradient kernels read coalesced
// (grad.cu: k_contract_tail); T == nullptr skips the group products.
__global__ void __launch_bounds__(256)
k_tables(const DimDesc* __restrict__ dims, const double* __restrict__ grid, const double* __restrict__ qs,
         const uint8_t* __restrict__ slot_k, const int* __restrict__ slot_group, const int* __restrict__ group_begin,
         int d, int sum_m, int sum_u, int width, int stride, int max_group_dims, const double* __restrict__ X,
         int64_t ldx, int64_t n, int64_t n_pad, double* __restrict__ T, int RB, int deriv_dim,
         const double* __restrict__ Kxu, int64_t ldk, double* __restrict__ KF) {
  extern __shared__ double sm[];
  double* sK = sm;
  double* sF = sm + (size_t)RB * sum_m;
  DimDesc* sD = reinterpret_cast<DimDesc*>(sF + (size_t)RB * sum_u);          // [d] the dimension descriptors
  uint8_t* dim_of_m = reinterpret_cast<uint8_t*>(sD + d);                       // [sum_m] dimension of a grid column
  uint8_t* dim_of_u = dim_of_m + sum_m;                                         // [sum_u] dimension of a factor column
  const int64_t row0 = (int64_t)blockIdx.x * RB;
  const int rows = (int)min((int64_t)RB, n_pad - row0);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < d; i += nt) sD[i] = dims[i];
  __syncthreads();
  for (int i = tid; i < d; i += nt) {               // lookup tables instead of a linear search per task (d is up to 64)
    const DimDesc dd = sD[i];
    for (int g = 0; g < dd.m; ++g) dim_of_m[dd.grid_off + g] = (uint8_t)i;
    for (int k = 0; k < dd.u; ++k) dim_of_u[dd.f_off + k] = (uint8_t)i;
  }
  __syncthreads();

  // phase 1a: kernel values against every grid point
  for (int task = tid; task < rows * sum_m; task += nt) {
    const int r = task / sum_m;
    const int c = task - r * sum_m;
    const int i = dim_of_m[c];
    const DimDesc dd = sD[i];
    const int64_t row = row0 + r;
    double v = 0.0;
    if (row < n) {
      if (dd.kernel == KERN_HOST) v = Kxu[row * ldk + c];
      else
        v = (i == deriv_dim) ? kern_eval_dx(dd.kernel, X[row * ldx + i], grid[c], dd.variance, dd.lengthscale)
                             : kern_eval(dd.kernel, X[row * ldx + i], grid[c], dd.variance, dd.lengthscale);
    }
    sK[(size_t)r * sum_m + c] = v;
  }
  __syncthreads();
  // phase 1b: project on the scaled eigenvectors
  for (int task = tid; task < rows * sum_u; task += nt) {
    const int r = task / sum_u;
    const int c = task - r * sum_u;
    const int i = dim_of_u[c];
    const int m_i = sD[i].m, u_i = sD[i].u;
    const int k = c - sD[i].f_off;
    const double* kv = sK + (size_t)r * sum_m + sD[i].grid_off;
    const double* q = qs + sD[i].q_off + k;
    double acc = 0.0;
    for (int g = 0; g < m_i; ++g) acc = fma(kv[g], __ldg(q + (size_t)g * u_i), acc);
    sF[(size_t)r * sum_u + c] = acc;
  }
  __syncthreads();
  if (KF != nullptr) {
    double* dst = KF + (size_t)blockIdx.x * (sum_m + sum_u) * 32;
    for (int task = tid; task < (sum_m + sum_u) * 32; task += nt) {
      const int c = task >> 5, r = task & 31;
      dst[task] = r < rows ? (c < sum_m ? sK[(size_t)r * sum_m + c] : sF[(size_t)r * sum_u + (c - sum_m)]) : 0.0;
    }
  }
  if (T == nullptr) return;
  // phase 2: group products, coalesced row-major store (pad rows >= n are written as zeros)
  for (int task = tid; task < rows * stride; task += nt) {
    const int r = task / stride;
    const int s = task - r * stride;
    const int64_t row = row0 + r;
    double v;
    if (s >= width || row >= n) v = 0.0;
    else if (s == 0) v = 1.0;
    else if (s == 1) v = 0.0;
    else {
      const int g = __ldg(slot_group + s);
      const int a = __ldg(group_begin + g), b = __ldg(group_begin + g + 1);
      const uint8_t* ks = slot_k + (size_t)s * max_group_dims;
      v = 1.0;
      for (int i = a; i < b; ++i) v *= sF[(size_t)r * sum_u + sD[i].f_off + __ldg(ks + i - a)];
    }
    T[row * stride + s] = v;
  }
}

int launch_tables(const Plan* pl, const double* X, int64_t ldx, int64_t n, int64_t n_pad, double* T,
                  cudaStream_t stream, int deriv_dim, const double* Kxu, int64_t ldk) {
  if (pl->n_host_dims > 0 && n > 0)
    GRIEF_REQUIRE(Kxu != nullptr && ldk >= pl->sum_m, "tables: %d dimension(s) have host-evaluated kernels: pass K_xu (n x %d) through "
                  "grief_build_tables_kxu", pl->n_host_dims, pl->sum_m);
  const size_t per_row = (size_t)(pl->sum_m + pl->sum_u) * sizeof(double);
  const size_t lookup = (size_t)pl->d * sizeof(DimDesc) + (size_t)(pl->sum_m + pl->sum_u) + 16;
  int RB = (int)std::min<size_t>(32, (160 * 1024 - lookup) / per_row);
  if (RB < 1) return fail(GRIEF_ERR_UNSUPPORTED, "tables: %d grid points + %d factors per row exceed shared memory",
                          pl->sum_m, pl->sum_u);
  const size_t smem = per_row * RB + lookup;
  GRIEF_CUDA(cudaFuncSetAttribute(k_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = (n_pad + RB - 1) / RB;
  if (blocks == 0) return GRIEF_OK;
  prof_begin(PROF_TABLES, stream);
  k_tables<<<(unsigned)blocks, 256, smem, stream>>>(pl->d_dims, pl->d_grid, pl->d_qs, pl->d_slot_k, pl->d_slot_group,
                                                    pl->d_group_begin, pl->d, pl->sum_m, pl->sum_u, pl->width,
                                                    pl->stride, pl->max_group_dims, X, ldx, n, n_pad, T, RB, deriv_dim, Kxu, ldk, nullptr);
  prof_end(PROF_TABLES, stream);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// K_xu and F of rows_blk rows (a multiple of 32; rows >= n_valid are written as zeros) in the [row group][column][32] layout.
size_t kf_doubles(const Plan* pl, int64_t rows) { return (size_t)rows * (pl->sum_m + pl->sum_u); }
int launch_kf(const Plan* pl, const double* X, int64_t ldx, int64_t n_valid, int64_t rows_blk, double* KF, cudaStream_t stream) {
  if (rows_blk == 0) return GRIEF_OK;
  GRIEF_REQUIRE(rows_blk % 32 == 0, "kf: rows=%lld is not a multiple of 32", (long long)rows_blk);
  const size_t per_row = (size_t)(pl->sum_m + pl->sum_u) * sizeof(double);
  const size_t lookup = (size_t)pl->d * sizeof(DimDesc) + (size_t)(pl->sum_m + pl->sum_u) + 16;
  const size_t smem = per_row * 32 + lookup;
  if (smem > 224 * 1024)
    return fail(GRIEF_ERR_UNSUPPORTED, "kf: %d grid points + %d factors per row exceed shared memory", pl->sum_m, pl->sum_u);
  GRIEF_CUDA(cudaFuncSetAttribute(k_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tables<<<(unsigned)(rows_blk / 32), 256, smem, stream>>>(pl->d_dims, pl->d_grid, pl->d_qs, pl->d_slot_k, pl->d_slot_group,
                                                            pl->d_group_begin, pl->d, pl->sum_m, pl->sum_u, pl->width, pl->stride,
                                                            pl->max_group_dims, X, ldx, n_valid, rows_blk, nullptr, 32, -1, nullptr, 0, KF);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// ---- Phi[n, j] materialisation (row-major n x p) for API parity / small n -------------------------
__global__ void __launch_bounds__(256)
k_phi_rows(const double* __restrict__ T, int stride, const uint16_t* __restrict__ col_slot, int G, int p, int64_t n,
           double* __restrict__ Phi) {
  const int64_t total = n * (int64_t)p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / p;
    const int j = (int)(e - row * p);
    const double* t = T + row * stride;
    double v = t[col_slot[(size_t)j * G]];
    for (int g = 1; g < G; ++g) v *= t[col_slot[(size_t)j * G + g]];
    Phi[e] = v;
  }
}

int launch_phi_rows(const Plan* pl, const double* T, int64_t n, double* Phi, cudaStream_t stream) {
  if (n == 0) return GRIEF_OK;
  const int64_t total = n * (int64_t)pl->p;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
  k_phi_rows<<<blocks, 256, 0, stream>>>(T, pl->stride, pl->d_col_slot, pl->n_groups, pl->p, n, Phi);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// ---- out[j] = sum_n v[n] * Phi[n, j]   (r = Phi^T y, reference models/gp_grief_model.py:234) -------
// Block b handles a contiguous row range; thread t owns columns t, t+256, ...  Partials go to
// ws[b][p]; a second kernel reduces them in block order (deterministic).
constexpr int kPtvCols = 8;  // columns per thread per sweep
__global__ void __launch_bounds__(256)
k_phi_t_vec(const double* __restrict__ T, int stride, const uint16_t* __restrict__ col_slot, int G, int p, int64_t n,
            int64_t rows_per_block, const double* __restrict__ v, double* __restrict__ ws) {
  extern __shared__ double sm[];  // [32][stride] table rows + [32] v
  const int RB = 32;
  double* sT = sm;
  double* sV = sm + (size_t)RB * stride;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(n, r_begin + rows_per_block);
  for (int j0 = 0; j0 < p; j0 += 256 * kPtvCols) {
    double acc[kPtvCols];
    int slots[kPtvCols][kMaxGroups];
#pragma unroll
    for (int c = 0; c < kPtvCols; ++c) {
      acc[c] = 0.0;
      const int j = j0 + c * 256 + threadIdx.x;
      for (int g = 0; g < G; ++g) slots[c][g] = (j < p) ? col_slot[(size_t)j * G + g] : 1;
    }
    for (int64_t rb = r_begin; rb < r_end; rb += RB) {
      const int rows = (int)min((int64_t)RB, r_end - rb);
      __syncthreads();
      for (int e = threadIdx.x; e < rows * stride; e += 256) sT[e] = T[rb * stride + e];
      if (threadIdx.x < rows) sV[threadIdx.x] = v[rb + threadIdx.x];
      __syncthreads();
      for (int r = 0; r < rows; ++r) {
        const double* t = sT + (size_t)r * stride;
        const double vr = sV[r];
#pragma unroll
        for (int c = 0; c < kPtvCols; ++c) {
          double ph = t[slots[c][0]];
          for (int g = 1; g < G; ++g) ph *= t[slots[c][g]];
          acc[c] = fma(ph, vr, acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kPtvCols; ++c) {
      const int j = j0 + c * 256 + threadIdx.x;
      if (j < p) ws[(size_t)blockIdx.x * p + j] = acc[c];
    }
  }
}

__global__ void k_reduce_rows(const double* __restrict__ ws, int nblk, int p, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += ws[(size_t)b * p + j];
  out[j] = s;
}

int phi_t_vec_blocks(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(148 * 2, (n + 511) / 512)); }

int launch_phi_t_vec(const Plan* pl, const double* T, int64_t n, const double* v, double* out, double* ws,
                     cudaStream_t stream) {
  const int nblk = phi_t_vec_blocks(n);
  int64_t rpb = (n + nblk - 1) / nblk;
  rpb = (rpb + 31) / 32 * 32;
  const size_t smem = ((size_t)32 * pl->stride + 32) * sizeof(double);
  GRIEF_CUDA(cudaFuncSetAttribute(k_phi_t_vec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(PROF_PHITY, stream);
  k_phi_t_vec<<<nblk, 256, smem, stream>>>(T, pl->stride, pl->d_col_slot, pl->n_groups, pl->p, n, rpb, v, ws);
  GRIEF_CUDA(cudaGetLastError());
  k_reduce_rows<<<(pl->p + 255) / 256, 256, 0, stream>>>(ws, nblk, pl->p, out);
  prof_end(PROF_PHITY, stream);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// ---- out[n] = sum_j Phi[n, j] * v[j]   (predictive mean, reference models/gp_grief_model.py:119) ---
__global__ void __launch_bounds__(256)
k_phi_vec(const double* __restrict__ T, int stride, const uint16_t* __restrict__ col_slot, int G, int p, int64_t n,
          const double* __restrict__ v, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < n; row += nwarps) {
    const double* t = T + row * stride;
    double acc = 0.0;
    for (int j = lane; j < p; j += 32) {
      double ph = t[col_slot[(size_t)j * G]];
      for (int g = 1; g < G; ++g) ph *= t[col_slot[(size_t)j * G + g]];
      acc = fma(ph, v[j], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
  }
}

int launch_phi_vec(const Plan* pl, const double* T, int64_t n, const double* v, double* out, cudaStream_t stream) {
  if (n == 0) return GRIEF_OK;
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 7) / 8, 148 * 8);
  k_phi_vec<<<blocks, 256, 0, stream>>>(T, pl->stride, pl->d_col_slot, pl->n_groups, pl->p, n, v, out);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// ---- s = sum_n y[n]^2 (deterministic two-stage) ------------------------------------------------------
__global__ void __launch_bounds__(256) k_sumsq_partial(const double* __restrict__ y, int64_t n, double* __restrict__ part) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc = fma(y[i], y[i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sh[w];
    part[blockIdx.x] = s;
  }
}
__global__ void k_sum_final(const double* __restrict__ part, int nparts, double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += part[i];
    out[0] = s;
  }
}

int launch_sumsq(const double* y, int64_t n, double* out, double* ws /* >= 296 doubles */, cudaStream_t stream) {
  const int nblk = 296;
  k_sumsq_partial<<<nblk, 256, 0, stream>>>(y, n, ws);
  GRIEF_CUDA(cudaGetLastError());
  k_sum_final<<<1, 32, 0, stream>>>(ws, nblk, out);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

}  // namespace grief
