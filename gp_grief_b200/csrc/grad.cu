// G4: analytic hyper-parameter gradient (replaces the reference's finite-difference loop,
// models/basemodel.py:328-361, which costs (#free + 1) full Phi / Gram rebuilds).
//
//   dLML/dtheta = sum_n sum_j Zt[n,j] * dPhi[n,j]/dtheta,
//   Zt = Phi * G2 + y g^T,  G2 = -(P^-1 + b b^T/sigma^2),  g = b/sigma^2                              
//      = -Phi P^-1 + a b^T,  a = (y - Phi b) / sigma^2   (the residual scaled by the noise: alpha of models/gp_grief_model.py:234)
// Only Zp = Phi P^-1 goes through the O(n p^2) GEMM (launch_zgemm, phi_stage.cu, one slab of rows at a time, stored TRANSPOSED with
// its columns in the plan's SORTED order); the rank-one part is added here in FP64 from f = Phi b.  (Round 1 multiplied by G2: its
// rows are dominated by b b^T/sigma^2, ~n times larger than P^-1, so the digit GEMM spent ~20 of its bits on a term that costs two
// multiplies per element to form exactly.)
//
// Reverse mode instead of a derivative table per parameter.  With Phi[n,j] = prod_g T[n, s_g(j)] (group tables, plan.h):
//   W[n,s]   = dL/dT[n,s]   = sum_{j: s_g(j) = s} Zt[n,j] * prod_{g' != g} T[n, s_g'(j)]
//   V_i[n,k] = dL/dF_i[n,k] = sum_{s in group(i): k_i(s) = k} W[n,s] * prod_{i' in group, i' != i} F_i'[n, k_i'(s)]
//   F_i[n,k] = sum_u K_i(x_n,u) Qs_i[u,k]   =>   dL/dtheta = sum_n sum_u dK_i(x_n,u)/dtheta * (sum_k V_i[n,k] Qs_i[u,k])
//                                                          + sum_{u,k} dQs_i[u,k]/dtheta * M_i[u,k],   M_i[u,k] = sum_n K_i(x_n,u) V_i[n,k]
// with dQs/dtheta (eigenvector / eigenvalue perturbation of the m_i x m_i grid problem) supplied by the host.
// Three kernels per slab of rows share the work (each at the occupancy its state allows):
//   k_build_phi (phi_stage.cu)   the slab builder that feeds the GEMM also forms f = Phi b from the values it has in registers and
//                                writes a = (y - f) / sigma^2 per row: no separate forward sweep.
//   k_contract_back              one warp per 32 data rows, LANE = ROW.  The sorted columns form a trie over the key positions
//                                (consecutive columns share their leading slots); every lane walks it in the same order, so there is
//                                no divergence and no cross-lane traffic: prefix products and per-level partial sums live in registers,
//                                W lives in shared memory as [slot][lane] (conflict-free) and is written out as [row group][slot][32].
//                                Per-row state (table row + W row, 1.6 KB) limits this kernel to 4 warps per SM, so it does nothing
//                                but the sweep over Zp^T.
//   k_contract_tail              W -> V -> parameter sums and the M_i matrices, 8 warps per SM (needs F, V and K_i of a row only).
// ~15 instructions per (row, column) instead of the ~110 of the round-1 kernel (warp per row, lane = column, one FMA per parameter),
// and the 4 KB/row derivative table of k_dtables is gone.
#include <vector>

#include "plan.h"

namespace grief {

struct GradDesc {            // device-resident description of the active parameters
  int n_active = 0;
  int* d_a_dim = nullptr;    // [n_active] input dimension
  int* d_a_kind = nullptr;   // 0 = variance, 1 = lengthscale
  int* d_a_qoff = nullptr;   // offset of the parameter's (m_i x u_i) block inside d_dqs
  double* d_dqs = nullptr;   // d(scaled eigenvectors) / d(theta), per active parameter
  ~GradDesc() { cudaFree(d_a_dim); cudaFree(d_a_kind); cudaFree(d_a_qoff); cudaFree(d_dqs); }
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
  std::vector<int> a_dim(n_active), a_kind(n_active), a_qoff(n_active);
  int qoff = 0;
  for (int a = 0; a < n_active; ++a) {
    const int i = dims[a];
    if (i < 0 || i >= pl->d || kinds[a] < 0 || kinds[a] > 1) {
      delete gd;
      return fail(GRIEF_ERR_BAD_ARG, "grad_desc: parameter %d has dim=%d kind=%d", a, i, kinds[a]);
    }
    a_dim[a] = i; a_kind[a] = kinds[a];
    a_qoff[a] = qoff; qoff += pl->dims[i].m * pl->dims[i].u;
  }
  std::vector<double> dqs(dqs_concat, dqs_concat + qoff);
  int rc = upload_vec(&gd->d_a_dim, a_dim);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_a_kind, a_kind);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_a_qoff, a_qoff);
  if (rc == GRIEF_OK) rc = upload_vec(&gd->d_dqs, dqs);
  if (rc != GRIEF_OK) { delete gd; return rc; }
  *out = gd;
  return GRIEF_OK;
}
void grad_desc_destroy(GradDesc* gd) { delete gd; }
int grad_desc_n_active(const GradDesc* gd) { return gd->n_active; }

// d k / d lengthscale from the kernel VALUE k (the tail kernel reads k from the K_xu buffer instead of evaluating exp again; the
// variance derivative is k / variance).  With r = |x - u| / l:
//   RBF          k = v exp(-r^2 / 2)                           dk/dl = k r^2 / l
//   Exponential  k = v exp(-r)                                 dk/dl = k r / l
//   Matern32     k = v (1 + s3 r) exp(-s3 r)                   dk/dl = v 3 r^2 exp(-s3 r) / l       = k 3 r^2 / ((1 + s3 r) l)
//   Matern52     k = v (1 + s5 r + 5/3 r^2) exp(-s5 r)         dk/dl = v 5/3 r^2 (1 + s5 r) exp(-s5 r) / l
//                                                                    = k 5/3 r^2 (1 + s5 r) / ((1 + s5 r + 5/3 r^2) l)
// (kern/stationary.py:108-258 has the kernel formulas; the reference differentiates them numerically.)
__device__ __forceinline__ double kern_dls_from_k(int kernel, double k, double x, double u, double ls) {
  const double diff = x - u;
  const double d2 = diff * diff;
  switch (kernel) {
    case KERN_RBF: return ls < 1e-6 ? 0.0 : k * d2 / (ls * ls * ls);
    case KERN_EXPONENTIAL: return k * (sqrt(d2) / ls) / ls;
    case KERN_MATERN32: {
      const double s3 = 1.7320508075688772;
      const double r = sqrt(d2) / ls;
      return k * 3.0 * r * r / ((1.0 + s3 * r) * ls);
    }
    default: {
      const double s5 = 2.23606797749979;
      const double r2 = d2 / (ls * ls);
      const double r = sqrt(r2);
      return k * (5.0 / 3.0) * r2 * (1.0 + s5 * r) / ((1.0 + s5 * r + (5.0 / 3.0) * r2) * ls);
    }
  }
}

// ---- backward sweep (lane = row; all control flow is warp-uniform) ----
// Sorted column c has its key-order slots packed one byte each in plan->d_sorted_pack ((G + 3) / 4 words, key position k in byte
// 3 - k % 4 of word k / 4); the first key position where c differs from c - 1 is a count-leading-zeros of the XOR.
template <int G>
struct Trie {
  double pfx[G > 1 ? G - 1 : 1];     // pfx[k] = h_0 * ... * h_k of the open path (levels 0 .. G-2)
  double hk[G > 1 ? G - 1 : 1];      // h_k of the open node at level k
  int sk[G > 1 ? G - 1 : 1];         // its slot
};

struct BackParams {
  const double* Zt; int64_t ldz;     // Zp^T slab: [p_pad][ldz], rows of the slab contiguous
  const double* T; int stride;       // slab rows x stride
  const double* a;                   // slab rows: (y - Phi b) / noise_var (0 for rows that do not exist), written by k_build_phi
  const double* bvec;                // p_pad: b in sorted column order (0 for padding columns)
  const uint32_t* pack;              // p_pad x NW packed key-order slots
  int p_pad;
  int n_rg;                          // 32-row groups in the slab
  double* W;                         // out: [n_rg][stride][32]  dL/dT of every row
};

// One warp per 32 data rows.  The table rows sit in shared memory as [lane][stride] (odd stride: conflict-free), W as [slot][lane].
// Every lane walks the trie of sorted columns in the same order: prefix products and per-level partial sums live in registers, only
// the last key position costs a shared-memory read-modify-write per column.  Two batches of NB columns of Zp^T are in flight per warp.
template <int G>
__global__ void __launch_bounds__(128) k_contract_back(const BackParams P) {
  extern __shared__ double sm[];
  constexpr int NW = (G + 3) / 4, NB = 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int stride = P.stride;
  double* sA = sm + (size_t)warp * (2 * stride * 32);                 // table rows [32][stride]
  double* sW = sA + (size_t)stride * 32;                              // [stride][32]
  const int wg = blockIdx.x * nw + warp, wtot = gridDim.x * nw;
  for (int rg = wg; rg < P.n_rg; rg += wtot) {
    __syncwarp();
    const int64_t row = (int64_t)rg * 32 + lane;
    {
      const double* src = P.T + (size_t)rg * 32 * stride;
      for (int e = lane; e < 32 * stride; e += 32) sA[e] = src[e];
      for (int e = lane; e < 32 * stride; e += 32) sW[e] = 0.0;
    }
    __syncwarp();
    const double* trow = sA + (size_t)lane * stride;
    const double a_n = P.a[row];
    Trie<G> tr;
    double S[G > 1 ? G - 1 : 1];
#pragma unroll
    for (int k = 0; k < (G > 1 ? G - 1 : 1); ++k) S[k] = 0.0;
    uint32_t prev[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) prev[w] = 0;
    // everything a batch needs from global memory (Zp^T from HBM; the packed slots and b from L1 / L2) is requested one batch
    // ahead: with four warps per SM nothing else hides those latencies
    const double* zcol = P.Zt + row;
    double znext[NB], bnext[NB];
    uint32_t wnext[NB][NW];
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      znext[e] = __ldcs(zcol + (size_t)e * P.ldz);
      bnext[e] = __ldg(P.bvec + e);
#pragma unroll
      for (int w = 0; w < NW; ++w) wnext[e][w] = __ldg(P.pack + (size_t)e * NW + w);
    }
    for (int c0 = 0; c0 < P.p_pad; c0 += NB) {
      uint32_t wd[NB][NW];
      double zt[NB];
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        zt[e] = fma(a_n, bnext[e], -znext[e]);
#pragma unroll
        for (int w = 0; w < NW; ++w) wd[e][w] = wnext[e][w];
      }
      if (c0 + NB < P.p_pad) {
#pragma unroll
        for (int e = 0; e < NB; ++e) {
          znext[e] = __ldcs(zcol + (size_t)(c0 + NB + e) * P.ldz);
          bnext[e] = __ldg(P.bvec + c0 + NB + e);
#pragma unroll
          for (int w = 0; w < NW; ++w) wnext[e][w] = __ldg(P.pack + (size_t)(c0 + NB + e) * NW + w);
        }
      }
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        const int sl = (int)((wd[e][(G - 1) >> 2] >> (8 * (3 - ((G - 1) & 3)))) & 0xFFu);
        if constexpr (G > 1) {
          int lv = G;                                  // first key position where this column differs from the previous one
#pragma unroll
          for (int w = NW - 1; w >= 0; --w) {
            const uint32_t x = wd[e][w] ^ prev[w];
            if (x) lv = 4 * w + (__clz(x) >> 3);
          }
          if (c0 + e == 0) lv = 0;
          if (lv < G - 1) {
            if (c0 + e > 0) {                          // close the open nodes of levels G-2 .. lv
#pragma unroll
              for (int k = G - 2; k >= 0; --k)
                if (k >= lv) {
                  const double up = k > 0 ? tr.pfx[k - 1] : 1.0;
                  sW[tr.sk[k] * 32 + lane] += up * S[k];
                  if (k > 0) S[k - 1] = fma(tr.hk[k], S[k], S[k - 1]);
                  S[k] = 0.0;
                }
            }
#pragma unroll
            for (int k = 0; k < G - 1; ++k)            // open levels lv .. G-2 for this column
              if (k >= lv) {
                tr.sk[k] = (int)((wd[e][k >> 2] >> (8 * (3 - (k & 3)))) & 0xFFu);
                tr.hk[k] = trow[tr.sk[k]];
                tr.pfx[k] = k > 0 ? tr.pfx[k - 1] * tr.hk[k] : tr.hk[k];
              }
          }
          sW[sl * 32 + lane] += tr.pfx[G - 2] * zt[e];
          S[G - 2] = fma(trow[sl], zt[e], S[G - 2]);
#pragma unroll
          for (int w = 0; w < NW; ++w) prev[w] = wd[e][w];
        } else {
          sW[sl * 32 + lane] += zt[e];
        }
      }
    }
    if constexpr (G > 1) {                             // close what is still open
#pragma unroll
      for (int k = G - 2; k >= 0; --k) {
        const double up = k > 0 ? tr.pfx[k - 1] : 1.0;
        sW[tr.sk[k] * 32 + lane] += up * S[k];
        if (k > 0) S[k - 1] = fma(tr.hk[k], S[k], S[k - 1]);
      }
    }
    __syncwarp();
    double* dst = P.W + (size_t)rg * stride * 32;
    for (int e = lane; e < 32 * stride; e += 32) dst[e] = sW[e];
  }
}

// ---- tail: W -> V (per-dimension factors) -> parameter sums ----
struct TailParams {
  const double* W;                   // [n_rg][stride][32]
  const double* KF;                  // [n_rg][sum_m + sum_u][32]: K_xu then F of the slab's rows (rows.cu: launch_kf)
  const double* X; int64_t ldx;      // slab rows x ldx
  const DimDesc* dims; const double* grid; const double* qs; const uint8_t* slot_k; const int* slot_group; const int* group_begin;
  int d, stride, width, sum_m, sum_u, max_group_dims, m_max, n_groups;
  int64_t rows_valid;                // rows of the slab that exist (the rest are zero-padded tables)
  int n_rg;
  double* acc; int acc_len;          // [total warps][acc_len]: M_i (sum m_i u_i doubles, layout of qs) then (gl_i, gv_i) per dimension
};

// One warp per 32 data rows, lane = row.  Per warp: V [sum_u][32] and K [m_max][33] in shared memory (4 warps per CTA, as many CTAs
// per SM as fit); per CTA the small index tables of the plan.  No kernel is evaluated here: K_xu and F come from the KF buffer
// (written by the table kernel at full occupancy; F is read in place, it stays in L1), d k / d lengthscale follows from k algebraically.
__global__ void __launch_bounds__(128) k_contract_tail(const TailParams P) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int per_warp = 32 * P.sum_u + 33 * P.m_max;
  double* sV = sm + (size_t)warp * per_warp;     // [sum_u][32]
  double* sK = sV + (size_t)P.sum_u * 32;        // [m_max][33]
  // CTA-wide copies of the index tables: f_off per dimension, group of a slot, group boundaries, per-slot factor indices
  int* s_foff = reinterpret_cast<int*>(sm + (size_t)nw * per_warp);             // [d]
  int* s_sgrp = s_foff + P.d;                                                    // [width]
  int* s_gbeg = s_sgrp + P.width;                                                // [groups + 1] (<= kMaxGroups + 1)
  uint8_t* s_slotk = reinterpret_cast<uint8_t*>(s_gbeg + kMaxGroups + 1);        // [width][max_group_dims]
  for (int i = threadIdx.x; i < P.d; i += blockDim.x) s_foff[i] = P.dims[i].f_off;
  for (int s = threadIdx.x; s < P.width; s += blockDim.x) s_sgrp[s] = P.slot_group[s];
  for (int g = threadIdx.x; g <= P.n_groups; g += blockDim.x) s_gbeg[g] = P.group_begin[g];
  for (int e = threadIdx.x; e < P.width * P.max_group_dims; e += blockDim.x) s_slotk[e] = P.slot_k[e];
  __syncthreads();
  const int wg = blockIdx.x * nw + warp, wtot = gridDim.x * nw;
  double* acc_out = P.acc + (size_t)wg * P.acc_len;
  for (int rg = wg; rg < P.n_rg; rg += wtot) {
    __syncwarp();
    const int64_t row = (int64_t)rg * 32 + lane;
    const bool valid = row < P.rows_valid;
    const double* Wl = P.W + (size_t)rg * P.stride * 32 + lane;
    const double* Kl = P.KF + (size_t)rg * (P.sum_m + P.sum_u) * 32 + lane;
    const double* Fl = Kl + (size_t)P.sum_m * 32;
    for (int e = 0; e < P.sum_u; ++e) sV[e * 32 + lane] = 0.0;
    constexpr int SB = 4;                          // W of the next SB slots is requested while SB slots are folded in
    double wnext[SB];
#pragma unroll
    for (int t = 0; t < SB; ++t) wnext[t] = (2 + t < P.width) ? __ldcs(Wl + (size_t)(2 + t) * 32) : 0.0;
    for (int s0 = 2; s0 < P.width; s0 += SB) {     // V_i[k_i(s)] += W[s] * prod_{i' != i} F_i'[k_i'(s)]
      double wcur[SB];
#pragma unroll
      for (int t = 0; t < SB; ++t) wcur[t] = wnext[t];
#pragma unroll
      for (int t = 0; t < SB; ++t) wnext[t] = (s0 + SB + t < P.width) ? __ldcs(Wl + (size_t)(s0 + SB + t) * 32) : 0.0;
#pragma unroll
      for (int t = 0; t < SB; ++t) {
        const int s = s0 + t;
        if (s >= P.width) break;
        const int g = s_sgrp[s];
        const int b0 = s_gbeg[g], b1 = s_gbeg[g + 1];
        const uint8_t* ks = s_slotk + (size_t)s * P.max_group_dims;
        for (int i = b0; i < b1; ++i) {
          double prod = wcur[t];
          for (int i2 = b0; i2 < b1; ++i2)
            if (i2 != i) prod *= __ldg(Fl + (size_t)(s_foff[i2] + (int)ks[i2 - b0]) * 32);
          sV[(s_foff[i] + (int)ks[i - b0]) * 32 + lane] += prod;
        }
      }
    }
    for (int i = 0; i < P.d; ++i) {                // parameter sums of dimension i and its M matrix
      const DimDesc dd = P.dims[i];
      const double x = valid ? P.X[row * P.ldx + i] : 0.0;
      double gl = 0.0, gv = 0.0;
      for (int g = 0; g < dd.m; ++g) {
        const double kv = __ldcs(Kl + (size_t)(dd.grid_off + g) * 32);       // 0 for rows that do not exist
        const double dk = kern_dls_from_k(dd.kernel, kv, x, __ldg(P.grid + dd.grid_off + g), dd.lengthscale);
        sK[g * 33 + lane] = kv;
        const double* q = P.qs + dd.q_off + (size_t)g * dd.u;
        double U = 0.0;
        for (int k = 0; k < dd.u; ++k) U = fma(sV[(dd.f_off + k) * 32 + lane], __ldg(q + k), U);
        gl = fma(dk, U, gl);
        gv = fma(kv, U, gv);
      }
      gl = warp_sum(gl);
      gv = warp_sum(gv) / dd.variance;
      const int ag = P.acc_len - 2 * P.d + 2 * i;
      if (lane == 0) { acc_out[ag] += gl; acc_out[ag + 1] += gv; }
      __syncwarp();
      for (int q = lane; q < dd.m * dd.u; q += 32) {      // M_i[g, k] += sum_rows K[g][row] V[k][row]
        const int g = q / dd.u, k = q - g * dd.u;
        const double* kr = sK + g * 33;
        const double* vr = sV + (size_t)(dd.f_off + k) * 32;
        double s = 0.0;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) s = fma(kr[r], vr[r], s);
        acc_out[dd.q_off + q] += s;
      }
      __syncwarp();
    }
  }
}

// grad[a] = sum over warps of (gl or gv of the parameter's dimension) + <dQs_a, sum over warps of M_dim>   (fixed order)
__global__ void __launch_bounds__(256) k_grad_finish(const double* __restrict__ acc, int n_warps, int acc_len, int d, int n_active,
                                                     const int* __restrict__ a_dim, const int* __restrict__ a_kind,
                                                     const int* __restrict__ a_qoff, const DimDesc* __restrict__ dims,
                                                     const double* __restrict__ dqs, double* __restrict__ grad) {
  __shared__ double red[256];
  const int a = blockIdx.x;
  if (a >= n_active) return;
  const DimDesc dd = dims[a_dim[a]];
  double part = 0.0;
  for (int q = threadIdx.x; q < dd.m * dd.u; q += blockDim.x) {
    double m = 0.0;
    for (int w = 0; w < n_warps; ++w) m += acc[(size_t)w * acc_len + dd.q_off + q];
    part = fma(dqs[a_qoff[a] + q], m, part);
  }
  if (threadIdx.x == 0) {
    const int ag = acc_len - 2 * d + 2 * a_dim[a] + (a_kind[a] == 1 ? 0 : 1);
    double g = 0.0;
    for (int w = 0; w < n_warps; ++w) g += acc[(size_t)w * acc_len + ag];
    part += g;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) grad[a] = red[0];
}

// out[row] = sum_c Zt[c][row] * Phi[row][c]   (diag of Phi* P^-1 Phi*^T for the predictive variance), lane = row, trie prefix products
template <int G>
__global__ void __launch_bounds__(128) k_rowdot_t(const double* __restrict__ Zt, int64_t ldz, const double* __restrict__ T, int stride,
                                                  const uint16_t* __restrict__ ss, const uint8_t* __restrict__ level, int p_pad,
                                                  int n_rg, int64_t rows_valid, double* __restrict__ out) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double* sA = sm + (size_t)warp * stride * 32;
  constexpr int NB = 16;
  for (int rg = blockIdx.x * nw + warp; rg < n_rg; rg += gridDim.x * nw) {
    __syncwarp();
    const double* src = T + (size_t)rg * 32 * stride;
    for (int e = lane; e < 32 * stride; e += 32) sA[e] = src[e];
    __syncwarp();
    const int64_t row = (int64_t)rg * 32 + lane;
    const double* trow = sA + (size_t)lane * stride;
    const double* zcol = Zt + row;
    double pfx[G > 1 ? G - 1 : 1];
    double acc = 0.0;
    double znext[NB];                                  // Z^T streams from HBM: the next batch is in flight while this one is consumed
#pragma unroll
    for (int e = 0; e < NB; ++e) znext[e] = __ldcs(zcol + (size_t)e * ldz);
    for (int c0 = 0; c0 < p_pad; c0 += NB) {
      int lv[NB], sl[NB];
      double z[NB];
#pragma unroll
      for (int e = 0; e < NB; ++e) z[e] = znext[e];
      if (c0 + NB < p_pad) {
#pragma unroll
        for (int e = 0; e < NB; ++e) znext[e] = __ldcs(zcol + (size_t)(c0 + NB + e) * ldz);
      }
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        lv[e] = (c0 + e == 0) ? 0 : (int)__ldg(level + c0 + e);
        sl[e] = __ldg(ss + (size_t)(c0 + e) * G + (G - 1));
      }
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        if constexpr (G > 1) {
          if (lv[e] < G - 1) {
#pragma unroll
            for (int k = 0; k < G - 1; ++k)
              if (k >= lv[e]) {
                const double h = trow[__ldg(ss + (size_t)(c0 + e) * G + k)];
                pfx[k] = k > 0 ? pfx[k - 1] * h : h;
              }
          }
          acc = fma(pfx[G - 2] * trow[sl[e]], z[e], acc);
        } else {
          acc = fma(trow[sl[e]], z[e], acc);
        }
      }
    }
    if (row < rows_valid) out[row] = acc;
  }
}

// ---- launchers ----
static int plan_m_max(const Plan* pl) {
  int mm = 1;
  for (auto& dd : pl->dims) mm = std::max(mm, dd.m);
  return mm;
}
// backward kernel: warps per CTA that fit shared memory (per warp: table rows [32][stride] and W [stride][32])
static int back_warps(const Plan* pl, size_t* smem_out) {
  const size_t per_warp = (size_t)2 * pl->stride * 32 * sizeof(double);
  const int nw = (int)std::min<size_t>(4, (224 * 1024) / per_warp);
  if (smem_out) *smem_out = per_warp * std::max(nw, 1);
  return nw;
}
// tail kernel: CTAs of 4 warps; per warp V [sum_u][32] and K [m_max][33], per CTA the index tables; as many CTAs per SM as fit
constexpr int kTailWarps = 4;
static size_t tail_smem(const Plan* pl) {
  const size_t per_warp = ((size_t)32 * pl->sum_u + (size_t)33 * plan_m_max(pl)) * sizeof(double);
  const size_t tables = (size_t)(pl->d + pl->width + kMaxGroups + 1) * sizeof(int) + (size_t)pl->width * pl->max_group_dims + 16;
  return per_warp * kTailWarps + tables;
}
static int tail_ctas_per_sm(const Plan* pl) {
  const size_t smem = tail_smem(pl) + 1024;          // + the per-CTA reservation of the driver
  return (int)std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / smem));
}
int contract_acc_len(const Plan* pl) {
  int qtot = 0;
  for (auto& dd : pl->dims) qtot += dd.m * dd.u;
  return qtot + 2 * pl->d;
}
// accumulator rows of the tail kernel (one per warp of its grid)
int contract_total_warps(const Plan* pl, int sms) { return sms * tail_ctas_per_sm(pl) * kTailWarps; }
// doubles of the W slab for a slab of `rows` table rows
size_t contract_w_doubles(const Plan* pl, int64_t rows) { return (size_t)rows * pl->stride; }

template <int G>
static int launch_back_g(const BackParams& P, int blocks, int threads, size_t smem, cudaStream_t stream) {
  GRIEF_CUDA(cudaFuncSetAttribute(k_contract_back<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_contract_back<G><<<blocks, threads, smem, stream>>>(P);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// Zt: Zp^T slab (p_pad x ldz); a: (y - Phi b) / noise_var of the slab's rows (from the builder); Wbuf: contract_w_doubles(rows_blk)
// doubles; KFbuf: kf_doubles(rows_blk) doubles; acc: contract_total_warps x contract_acc_len doubles, zeroed by the caller before the first slab.
int launch_kf(const Plan* pl, const double* X, int64_t ldx, int64_t n_valid, int64_t rows_blk, double* KF, cudaStream_t stream);
int launch_contract(const Plan* pl, const double* Zt, int64_t ldz, const double* T, const double* X, int64_t ldx, const double* a,
                    const double* bvec, int64_t rows_blk, int64_t rows_valid, double* Wbuf, double* KFbuf, double* acc, int sms,
                    cudaStream_t stream) {
  if (rows_blk == 0) return GRIEF_OK;
  size_t smem_b = 0;
  const size_t smem_t = tail_smem(pl);
  const int nwb = back_warps(pl, &smem_b);
  if (nwb < 1) return fail(GRIEF_ERR_UNSUPPORTED, "contract: a table row of %d entries does not fit shared memory", pl->stride);
  if (smem_t > 220 * 1024) return fail(GRIEF_ERR_UNSUPPORTED, "contract: %d factors per row do not fit shared memory", pl->sum_u);
  BackParams B;
  B.Zt = Zt; B.ldz = ldz; B.T = T; B.stride = pl->stride; B.a = a; B.bvec = bvec; B.pack = pl->d_sorted_pack; B.p_pad = pl->p_pad;
  B.n_rg = (int)(rows_blk / 32); B.W = Wbuf;
  int rc;
  prof_begin(PROF_CONTRACT, stream);
  switch (pl->n_groups) {
    case 1: rc = launch_back_g<1>(B, sms, nwb * 32, smem_b, stream); break;
    case 2: rc = launch_back_g<2>(B, sms, nwb * 32, smem_b, stream); break;
    case 3: rc = launch_back_g<3>(B, sms, nwb * 32, smem_b, stream); break;
    case 4: rc = launch_back_g<4>(B, sms, nwb * 32, smem_b, stream); break;
    case 5: rc = launch_back_g<5>(B, sms, nwb * 32, smem_b, stream); break;
    case 6: rc = launch_back_g<6>(B, sms, nwb * 32, smem_b, stream); break;
    case 7: rc = launch_back_g<7>(B, sms, nwb * 32, smem_b, stream); break;
    case 8: rc = launch_back_g<8>(B, sms, nwb * 32, smem_b, stream); break;
    default: return fail(GRIEF_ERR_UNSUPPORTED, "contract: %d groups", pl->n_groups);
  }
  prof_end(PROF_CONTRACT, stream);
  if (rc != GRIEF_OK) return rc;
  prof_begin(PROF_DTABLES, stream);
  rc = launch_kf(pl, X, ldx, rows_valid, rows_blk, KFbuf, stream);
  if (rc != GRIEF_OK) return rc;
  TailParams Q;
  Q.W = Wbuf; Q.KF = KFbuf; Q.X = X; Q.ldx = ldx;
  Q.dims = pl->d_dims; Q.grid = pl->d_grid; Q.qs = pl->d_qs; Q.slot_k = pl->d_slot_k; Q.slot_group = pl->d_slot_group; Q.group_begin = pl->d_group_begin;
  Q.d = pl->d; Q.stride = pl->stride; Q.width = pl->width; Q.sum_m = pl->sum_m; Q.sum_u = pl->sum_u; Q.max_group_dims = pl->max_group_dims; Q.m_max = plan_m_max(pl); Q.n_groups = pl->n_groups;
  Q.rows_valid = rows_valid; Q.n_rg = (int)(rows_blk / 32);
  Q.acc = acc; Q.acc_len = contract_acc_len(pl);
  GRIEF_CUDA(cudaFuncSetAttribute(k_contract_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
  k_contract_tail<<<sms * tail_ctas_per_sm(pl), kTailWarps * 32, smem_t, stream>>>(Q);
  prof_end(PROF_DTABLES, stream);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

int launch_grad_finish(const Plan* pl, const GradDesc* gd, const double* acc, int sms, double* out, cudaStream_t stream) {
  if (gd->n_active == 0) return GRIEF_OK;
  k_grad_finish<<<gd->n_active, 256, 0, stream>>>(acc, contract_total_warps(pl, sms), contract_acc_len(pl), pl->d, gd->n_active,
                                                  gd->d_a_dim, gd->d_a_kind, gd->d_a_qoff, pl->d_dims, gd->d_dqs, out);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

int launch_rowdot(const Plan* pl, const double* Zt, int64_t ldz, const double* T, int64_t rows_blk, int64_t rows_valid, double* out,
                  cudaStream_t stream) {
  if (rows_blk == 0 || rows_valid == 0) return GRIEF_OK;
  const size_t per_warp = (size_t)pl->stride * 32 * sizeof(double);
  const int nw = (int)std::min<size_t>(4, (220 * 1024) / per_warp);
  if (nw < 1) return fail(GRIEF_ERR_UNSUPPORTED, "rowdot: a table row of %d entries does not fit shared memory", pl->stride);
  const size_t smem = per_warp * nw;
  const int n_rg = (int)(rows_blk / 32);
  const unsigned blocks = (unsigned)std::min<int>((n_rg + nw - 1) / nw, 148 * 2);
#define RD(Gv)                                                                                                          \
  do {                                                                                                                  \
    GRIEF_CUDA(cudaFuncSetAttribute(k_rowdot_t<Gv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    k_rowdot_t<Gv><<<blocks, nw * 32, smem, stream>>>(Zt, ldz, T, pl->stride, pl->d_sorted_slot, pl->d_sorted_level,    \
                                                      pl->p_pad, n_rg, rows_valid, out);                                \
  } while (0)
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
