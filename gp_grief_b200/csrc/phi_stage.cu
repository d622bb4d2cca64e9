// The two O(n p^2) products of an evaluation, staged:  the basis matrix Phi of a slab of data rows is built by a
// bandwidth-bound kernel into HBM, then a plain TMA-fed FP64 DMMA GEMM (dense.cu, k_gemm_nt) consumes it.
//
//   pass 1 (models/gp_grief_model.py:148-149)   A = Phi^T Phi :  Phi^T slab (p_pad x R, data rows contiguous)
//                                               -> SYRK on the lower tiles, K = data rows split over gridDim.z
//   pass 2 / predictive var.    Z = Phi B     :  Phi slab (R x p_pad, sorted columns contiguous)
//                                               -> GEMM against the column-permuted symmetric B
//
// Two arithmetic modes (per plan, PlanOpts): 1 (default) -- the builders emit power-of-two row scales and 4..7 int8 digit planes
// and k_ozaki (ozaki.cu) multiplies them on the tcgen05 INT8 tensor cores; 0 -- the builders emit the FP64 slab for k_gemm_nt.
// r = Phi^T y is formed inside the pass-1 builder (the values are in registers there).
//
// Why staged and not fused (round-1 measurements, profiles/r01_gram_design_notes.md): DMUL, DFMA and DMMA share ONE FP64
// pipe per SM sub-partition.  With the Phi tiles built inside the GEMM CTAs the builder's DMULs queue behind the DMMAs
// (32 % of warp samples in stall_math) and the kernels stop at 25-28 TFLOP/s; the same DMMA loop fed by TMA alone runs at
// 36 TFLOP/s.  The slab costs 7-8 B written + read per element of Phi, a few % of the GEMM time, on an HBM that is otherwise
// idle during these passes.
#include <cuda.h>

#include <algorithm>

#include "plan.h"

namespace grief {

constexpr int kBuildRows = 128;       // table rows per builder CTA
constexpr int kBuildThreads = 512;

// table rows [rb*128, rb*128+128) -> shared memory (bulk async copy in 16-row pieces); `use` = how often this CTA has
// staged rows before (barrier parity).  The barrier is initialised by the caller (init_stage_barrier).
__device__ __forceinline__ void init_stage_barrier(uint64_t* bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
}
__device__ __forceinline__ void stage_table_rows(double* sT, uint64_t* bar, const double* T, int stride, int64_t rb, int use) {
  __syncthreads();                                   // everybody is done with the previous rows
  if (threadIdx.x == 0) {
    const uint32_t piece = (uint32_t)(16 * stride * sizeof(double));
    fence_proxy_async();
    mbar_arrive_expect_tx(bar, piece * (kBuildRows / 16));
    for (int i = 0; i < kBuildRows / 16; ++i)
      bulk_g2s(sT + (size_t)i * 16 * stride, T + ((size_t)rb * kBuildRows + i * 16) * stride, piece, bar);
  }
  mbar_wait(bar, (uint32_t)(use & 1));
}

// What a builder launch produces: the FP64 slab (DMMA path), or -- for the INT8 path, which never stores the FP64 slab --
// the `sd` balanced 8-bit digits of every element, scaled by the power of two of its operand row (exponents come first).
enum BuildOut : int { OUT_F64 = 0, OUT_DIGITS = 2 };

// high word of |v|: orders like |v|, and its exponent field is all the INT8 path needs of a row maximum
__device__ __forceinline__ int abs_hi(double v) { return __double2hiint(v) & 0x7fffffff; }
// frexp exponent e (|x| < 2^e) from the high word of the row maximum (0 for an all-zero row)
__device__ __forceinline__ int exp_from_hi(int hi) { return hi > 0 ? min(max((hi >> 20) - 1022, -900), 1024) : 0; }   // clamp: 2^(8 sd - 2 - e) stays finite
// 2^(8 sd - 2 - e): multiplies a value below 2^e in magnitude into the range of sd digits
__device__ __forceinline__ double digit_scale(int sd, int e) { return __hiloint2double((1023 + 8 * sd - 2 - e) << 20, 0); }
// digits d_s of q = rint(v * scale), scale = 2^(8 sd - 2 - e):  v 2^-e = sum_s d_s 2^(-6 - 8 s), d_s in [-128, 127], s < sd.
// Balanced base-256 digits without a carry chain: add 128 to every byte position (q + 0x80..80 is positive and below 2^(8 sd)),
// then byte k of the sum, minus 128 (= XOR 0x80 read as int8), is the digit of 256^k.  Digit 0 (most significant) is moved to byte 6.
__device__ __forceinline__ void store_digits(double v, double scale, int sd, int8_t* __restrict__ base, size_t plane_stride) {
  const unsigned long long bias = 0x0080808080808080ull >> (8 * (7 - sd));
  const unsigned long long w = (((unsigned long long)__double2ll_rn(v * scale) + bias) ^ bias) << (8 * (7 - sd));
  const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
  base[0 * plane_stride] = (int8_t)(hi >> 16);       // digit 0 = byte 6 (most significant)
  base[1 * plane_stride] = (int8_t)(hi >> 8);
  base[2 * plane_stride] = (int8_t)hi;
  if (sd > 3) base[3 * plane_stride] = (int8_t)(lo >> 24);
  if (sd > 4) base[4 * plane_stride] = (int8_t)(lo >> 16);
  if (sd > 5) base[5 * plane_stride] = (int8_t)(lo >> 8);
  if (sd > 6) base[6 * plane_stride] = (int8_t)lo;
}

// ---- exponents of the Phi^T operand rows (= basis columns) of a slab, without a pass over Phi ----
// |Phi[n][c]| = prod_g |T[n][slot_g(c)]| <= prod_g max_n |T[n][slot_g(c)]|: one sweep over the slab's TABLE rows (stride doubles per
// data row instead of p) gives per-slot maxima, and the product of a column's G maxima bounds the column.  The bound costs at
// most a bit or two of the 8 sd - 2 (the factors of a column peak at different rows); it replaces a second evaluation of all of Phi.
__global__ void __launch_bounds__(256) k_slot_hi(const double* __restrict__ T, int stride, int64_t total, int* __restrict__ slot_hi) {
  extern __shared__ int s_hi_slots[];
  for (int s = threadIdx.x; s < stride; s += blockDim.x) s_hi_slots[s] = 0;
  __syncthreads();
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int h = abs_hi(T[e]);
    const int s = (int)(e % stride);
    if (h > s_hi_slots[s]) atomicMax(s_hi_slots + s, h);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < stride; s += blockDim.x)
    if (s_hi_slots[s] > 0) atomicMax(slot_hi + s, s_hi_slots[s]);
}
// exps[c] from the slot maxima (c = sorted column); entries c in [n, n_pad) are zeroed
__global__ void k_col_exps(const int* __restrict__ slot_hi, const uint16_t* __restrict__ sorted_slot, int G, int n, int n_pad,
                           int* __restrict__ exps, int* __restrict__ err) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) {
    double bound = 1.0;
    bool bad = false;
    for (int g = 0; g < G; ++g) {
      const int h = slot_hi[sorted_slot[(size_t)c * G + g]];
      if (h >= 0x7ff00000 - 2) bad = true;            // Inf / NaN in a table entry
      bound *= h > 0 ? __hiloint2double(h + 2, 0) : 0.0;   // the slot maximum rounded UP in its high word (covers the product's roundings)
    }
    const int hb = abs_hi(bound);
    if (bad || hb >= 0x7ff00000) atomicExch(err, 4);
    exps[c] = exp_from_hi(hb);
  } else if (c < n_pad) exps[c] = 0;
}

// Phi^T slab, c = sorted column, one CTA per 128 table rows.  Lane = data row; each of the 16 warps owns p_pad / 16 consecutive
// sorted columns and walks them for the four 32-row groups.  Consecutive sorted columns share their leading slots: the product P
// of the first G-1 factors stays in a register per row group and is rebuilt only where sorted_level says a leading factor changed
// (warp-uniform branch) -- ~1.5 gathers and 1.3 DMULs per element.
//   DIG = false: out[c * ld + row]                                  (FP64 slab)
//   DIG = true:  planes[s][c][row] (row stride ld bytes, plane stride p_pad * ld), scaled by 2^(8 sd - 2 - exps[c])
// y != nullptr: the values are already in registers, so r_ws[block][c] = sum over the block's rows of y[row] * Phi[row][c] is formed
// here (fixed order: lanes by butterfly, row groups in sequence) instead of rebuilding Phi a second time for Phi^T y.
template <int G, bool DIG>
__global__ void __launch_bounds__(kBuildThreads) k_build_phi_t(const double* __restrict__ T, int stride,
                                                               const uint16_t* __restrict__ sorted_slot,
                                                               const uint8_t* __restrict__ sorted_level, int p_pad,
                                                               double* __restrict__ out, int64_t ld, const int* __restrict__ exps,
                                                               int8_t* __restrict__ planes, int sd, const double* __restrict__ y,
                                                               int64_t y_rows, double* __restrict__ r_ws) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sT = reinterpret_cast<double*>(smem_raw + 128);
  init_stage_barrier(bar);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cpw = p_pad / (kBuildThreads / 32);        // columns per warp (p_pad is a multiple of 128)
  const int c_begin = warp * cpw;
  constexpr int NB = 8, RG = kBuildRows / 32;
  const int64_t rb = blockIdx.x;
  stage_table_rows(sT, bar, T, stride, rb, 0);
  const size_t grow0 = (size_t)rb * kBuildRows + lane;
  static_assert(RG == 4, "four row groups of 32 rows");
  double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;       // y of this lane's row in each row group
  if (y != nullptr) {
    const int64_t row = (int64_t)grow0;
    y0 = row < y_rows ? y[row] : 0.0;
    y1 = row + 32 < y_rows ? y[row + 32] : 0.0;
    y2 = row + 64 < y_rows ? y[row + 64] : 0.0;
    y3 = row + 96 < y_rows ? y[row + 96] : 0.0;
  }
  double P0 = 1.0, P1 = 1.0, P2 = 1.0, P3 = 1.0;       // running prefix products, one per row group
  for (int c0 = c_begin; c0 < c_begin + cpw; c0 += NB) {
    int lv[NB], sl[NB];
    double dot[NB];
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      lv[e] = (c0 + e == c_begin) ? 0 : (int)__ldg(sorted_level + c0 + e);
      sl[e] = __ldg(sorted_slot + (size_t)(c0 + e) * G + (G - 1));
      dot[e] = 0.0;
    }
#pragma unroll 1
    for (int rg = 0; rg < RG; ++rg) {                   // not unrolled: one copy of the body keeps the kernel below 128 registers
      const double* trow = sT + (size_t)(rg * 32 + lane) * stride;
      double P = rg == 0 ? P0 : (rg == 1 ? P1 : (rg == 2 ? P2 : P3));
      const double yr = rg == 0 ? y0 : (rg == 1 ? y1 : (rg == 2 ? y2 : y3));
      double last[NB];
#pragma unroll
      for (int e = 0; e < NB; ++e) last[e] = trow[sl[e]];
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        if constexpr (G > 1) {
          if (lv[e] < G - 1) {
            double q = trow[__ldg(sorted_slot + (size_t)(c0 + e) * G)];
#pragma unroll
            for (int g = 1; g < G - 1; ++g) q *= trow[__ldg(sorted_slot + (size_t)(c0 + e) * G + g)];
            P = q;
          }
          last[e] *= P;
        }
      }
      if (rg == 0) P0 = P; else if (rg == 1) P1 = P; else if (rg == 2) P2 = P; else P3 = P;
      const size_t grow = grow0 + rg * 32;
      if constexpr (!DIG) {
#pragma unroll
        for (int e = 0; e < NB; ++e) out[(size_t)(c0 + e) * ld + grow] = last[e];
      } else {
#pragma unroll
        for (int e = 0; e < NB; ++e)
          store_digits(last[e], digit_scale(sd, __ldg(exps + c0 + e)), sd, planes + (size_t)(c0 + e) * ld + grow, (size_t)p_pad * ld);
      }
      if (y != nullptr) {                                // warp-uniform
#pragma unroll
        for (int e = 0; e < NB; ++e) dot[e] = fma(yr, last[e], dot[e]);
      }
    }
    if (y != nullptr) {                                  // warp-uniform
      double mine = 0.0;
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        const double t = warp_sum(dot[e]);
        if (lane == e) mine = t;
      }
      if (lane < NB) r_ws[(size_t)rb * p_pad + c0 + lane] = mine;
    }
  }
}

// r_acc[c] (+)= sum_b r_ws[b][c] over the nblk row blocks of a slab (fixed order: warp w sums b = w, w + 16, ..., then warps in sequence)
__global__ void __launch_bounds__(512) k_reduce_r(const double* __restrict__ r_ws, int nblk, int p_pad, int accumulate, double* __restrict__ r_acc) {
  __shared__ double part[16][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s = 0.0;
  for (int b = warp; b < nblk; b += 16) s += r_ws[(size_t)b * p_pad + c];
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 16; ++w) t += part[w][lane];
    r_acc[c] = accumulate ? r_acc[c] + t : t;
  }
}
// r[perm[c]] = r_acc[c]  (sorted column order -> the caller's column order)
__global__ void k_unpermute_vec(const double* __restrict__ in, const int* __restrict__ perm, int p_pad, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < p_pad && perm[c] >= 0) out[perm[c]] = in[c];
}

// Phi slab, row-major, lane = sorted column (coalesced stores); the G slots of a column are loaded once and reused for the
// warp's eight rows.
//   DIG = false: out[row * ldo + c]
//   DIG = true:  exps[row] from the row maximum (first sweep, registers only), then planes[s][row][c] (second sweep)
template <int G, bool DIG>
__global__ void __launch_bounds__(kBuildThreads) k_build_phi(const double* __restrict__ T, int stride,
                                                             const uint16_t* __restrict__ sorted_slot, int p_pad,
                                                             double* __restrict__ out, int64_t ldo, int* __restrict__ exps,
                                                             int8_t* __restrict__ planes, size_t plane_stride, int sd, int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sT = reinterpret_cast<double*>(smem_raw + 128);
  init_stage_barrier(bar);
  stage_table_rows(sT, bar, T, stride, blockIdx.x, 0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int RW = kBuildRows / (kBuildThreads / 32);   // 8 rows per warp
  const double* tbase = sT + (size_t)warp * RW * stride;
  const size_t row0 = (size_t)blockIdx.x * kBuildRows + warp * RW;
  double scale[RW];
  if constexpr (DIG) {
    int hi[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) hi[r] = 0;
    for (int c = lane; c < p_pad; c += 32) {
      int sl[G];
#pragma unroll
      for (int g = 0; g < G; ++g) sl[g] = __ldg(sorted_slot + (size_t)c * G + g);
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        double v = tbase[r * stride + sl[0]];
#pragma unroll
        for (int g = 1; g < G; ++g) v *= tbase[r * stride + sl[g]];
        hi[r] = max(hi[r], abs_hi(v));
      }
    }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int mh = __reduce_max_sync(0xffffffffu, hi[r]);
      const int e = exp_from_hi(mh);
      if (lane == 0) {
        exps[row0 + r] = e;
        if (mh >= 0x7ff00000) atomicExch(err, 4);   // Inf / NaN in the row
      }
      scale[r] = digit_scale(sd, e);
    }
  }
  for (int c = lane; c < p_pad; c += 32) {
    int sl[G];
#pragma unroll
    for (int g = 0; g < G; ++g) sl[g] = __ldg(sorted_slot + (size_t)c * G + g);
    double v[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) v[r] = tbase[r * stride + sl[0]];
#pragma unroll
    for (int g = 1; g < G; ++g)
#pragma unroll
      for (int r = 0; r < RW; ++r) v[r] *= tbase[r * stride + sl[g]];
    if constexpr (!DIG) {
#pragma unroll
      for (int r = 0; r < RW; ++r) out[(row0 + r) * ldo + c] = v[r];
    } else {
#pragma unroll
      for (int r = 0; r < RW; ++r) store_digits(v[r], scale[r], sd, planes + (row0 + r) * (size_t)p_pad + c, plane_stride);
    }
  }
}

struct BuildArgs {
  bool transposed = false;
  bool digits = false;
  int sd = kOzMaxDigits;                     // digits per element (INT8 path)
  double* out = nullptr; int64_t ld = 0;     // FP64 slab; ld also = bytes per plane row of the transposed digits
  int* exps = nullptr;                       // read (transposed digits) or written (row-major digits)
  int8_t* planes = nullptr; size_t plane_stride = 0;
  const double* y = nullptr; int64_t y_rows = 0; double* r_ws = nullptr;   // transposed only: fused Phi^T y partials
};

template <int G>
static int launch_build_g(const Plan* pl, const double* T, int64_t rows, const BuildArgs& a, cudaStream_t stream) {
  const size_t smem = 128 + (size_t)kBuildRows * pl->stride * sizeof(double);
  const unsigned grid = (unsigned)(rows / kBuildRows);
  GRIEF_REQUIRE(smem <= 227 * 1024, "build_phi: %zu bytes of shared memory", smem);
#define GRIEF_BT(DIG_)                                                                                                              \
  do {                                                                                                                              \
    GRIEF_CUDA(cudaFuncSetAttribute(k_build_phi_t<G, DIG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    k_build_phi_t<G, DIG_><<<grid, kBuildThreads, smem, stream>>>(T, pl->stride, pl->d_sorted_slot, pl->d_sorted_level, pl->p_pad,  \
                                                                 a.out, a.ld, a.exps, a.planes, a.sd, a.y, a.y_rows, a.r_ws);       \
  } while (0)
#define GRIEF_BN(DIG_)                                                                                                              \
  do {                                                                                                                              \
    GRIEF_CUDA(cudaFuncSetAttribute(k_build_phi<G, DIG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    k_build_phi<G, DIG_><<<grid, kBuildThreads, smem, stream>>>(T, pl->stride, pl->d_sorted_slot, pl->p_pad, a.out, a.ld, a.exps,   \
                                                               a.planes, a.plane_stride, a.sd, pl->d_err);                          \
  } while (0)
  if (a.transposed) {
    if (a.digits) GRIEF_BT(true); else GRIEF_BT(false);
  } else {
    if (a.digits) GRIEF_BN(true); else GRIEF_BN(false);
  }
#undef GRIEF_BT
#undef GRIEF_BN
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// rows: multiple of 128.
static int launch_build(const Plan* pl, const double* T, int64_t rows, const BuildArgs& a, cudaStream_t stream) {
  if (rows == 0) return GRIEF_OK;
  GRIEF_REQUIRE(rows % kBuildRows == 0, "build_phi: rows=%lld is not a multiple of %d", (long long)rows, kBuildRows);
  switch (pl->n_groups) {
    case 1: return launch_build_g<1>(pl, T, rows, a, stream);
    case 2: return launch_build_g<2>(pl, T, rows, a, stream);
    case 3: return launch_build_g<3>(pl, T, rows, a, stream);
    case 4: return launch_build_g<4>(pl, T, rows, a, stream);
    case 5: return launch_build_g<5>(pl, T, rows, a, stream);
    case 6: return launch_build_g<6>(pl, T, rows, a, stream);
    case 7: return launch_build_g<7>(pl, T, rows, a, stream);
    case 8: return launch_build_g<8>(pl, T, rows, a, stream);
    default: return fail(GRIEF_ERR_UNSUPPORTED, "build_phi: %d groups", pl->n_groups);
  }
}

static OzOpts oz_opts(const Plan* pl, int digits) {
  OzOpts o;
  o.digits = digits; o.cluster = pl->opts.cluster; o.err = pl->d_err;
  return o;
}

// ---- pass 1: A = Phi^T Phi (and r = Phi^T y from the same sweep) ----
struct GramSchedule {
  int nb, n_tiles, splits;
  int64_t slab_rows;
};

constexpr int kOzakiKRange = 16384;   // values of K per int32 accumulation in k_ozaki

GramSchedule gram_schedule(const Plan* pl, int64_t n_pad, int sms) {
  GramSchedule s;
  const int p_pad = pl->p_pad;
  s.nb = p_pad / kTileN;
  s.n_tiles = s.nb * (s.nb + 1) / 2;
  const int64_t budget_rows = (int64_t)(pl->opts.slab_budget / ((size_t)p_pad * 8)) / kBuildRows * kBuildRows;
  s.slab_rows = std::max<int64_t>(kBuildRows, std::min<int64_t>(n_pad, std::max<int64_t>(kBuildRows, budget_rows)));
  // K splits: fill whole waves of `sms` CTAs, keep >= 256 data rows per split
  const int64_t max_splits = std::max<int64_t>(1, std::min<int64_t>(s.slab_rows / 256, 64));
  int best = 1;
  double best_eff = -1.0;
  for (int64_t S = 1; S <= max_splits; ++S) {
    const int64_t items = S * s.n_tiles;
    const int64_t waves = (items + sms - 1) / sms;
    const double eff = (double)items / (double)(waves * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = (int)S; }
    if (eff >= 0.97) { best = (int)S; break; }
  }
  s.splits = best;
  if (pl->opts.gemm_mode == 1)      // at most 16384 rows per int32 accumulation; more splits when that is needed to fill the SMs
    s.splits = std::max(best, (int)((s.slab_rows + kOzakiKRange - 1) / kOzakiKRange));
  return s;
}

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

static size_t gram_exps_len(const Plan* pl) { return (size_t)(pl->p_pad + 255) / 256 * 256; }

// workspace: [Phi^T slab (FP64, DMMA path) | digit planes (INT8 path)] [split partials] [exps] [slot_hi] [r_ws] [r_acc]
size_t gram_workspace_bytes(const Plan* pl, int64_t n_pad, int sms) {
  const GramSchedule s = gram_schedule(pl, n_pad, sms);
  const size_t slab = pl->opts.gemm_mode == 1 ? align256(ozaki_plane_bytes(pl->p_pad, (int)s.slab_rows))
                                              : align256((size_t)pl->p_pad * s.slab_rows * sizeof(double));
  return slab + align256((size_t)s.splits * pl->p_pad * pl->p_pad * sizeof(double)) + 2 * align256(gram_exps_len(pl) * sizeof(int)) +
         align256((size_t)(s.slab_rows / kBuildRows) * pl->p_pad * sizeof(double)) + align256((size_t)pl->p_pad * sizeof(double));
}

// A[perm[i]][perm[j]] = sum_s part[s][i][j] over the lower tiles (fixed split order), mirrored bit-identically.
__global__ void __launch_bounds__(256)
k_gram_reduce(const double* __restrict__ part, const int* __restrict__ perm, int p_pad, int splits, int64_t lda, double* __restrict__ A) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const size_t pp = (size_t)p_pad * p_pad;
  for (int e = threadIdx.x; e < kTileN * kTileN; e += blockDim.x) {
    const int m = e / kTileN, n = e - m * kTileN;
    const int srow = bi * kTileN + m, scol = bj * kTileN + n;
    if (bi == bj && scol > srow) continue;
    const int row = perm[srow], col = perm[scol];
    if (row < 0 || col < 0) continue;                // padding columns
    double s = 0.0;
    for (int sp = 0; sp < splits; ++sp) s += part[sp * pp + (size_t)srow * p_pad + scol];
    A[(size_t)row * lda + col] = s;
    A[(size_t)col * lda + row] = s;
  }
}

// r != nullptr: r (p) = Phi^T y comes out of the builders' sweep (y: n valid rows).
int launch_gram(const Plan* pl, const double* T, int64_t n_pad, double* A, int64_t lda, const double* y, int64_t n, double* r,
                void* workspace, size_t ws_bytes, int sms, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(n_pad % kBuildRows == 0, "gram: n_pad=%lld must be a multiple of %d", (long long)n_pad, kBuildRows);
  GRIEF_REQUIRE(ws_bytes >= gram_workspace_bytes(pl, n_pad, sms), "gram: workspace too small");
  const GramSchedule s = gram_schedule(pl, n_pad, sms);
  const int pp = pl->p_pad;
  const bool i8 = pl->opts.gemm_mode == 1;
  const int sd = pl->opts.digits_gram;
  char* wq = reinterpret_cast<char*>(workspace);
  double* PhiT = reinterpret_cast<double*>(wq);
  int8_t* planes = reinterpret_cast<int8_t*>(wq);
  wq += i8 ? align256(ozaki_plane_bytes(pp, (int)s.slab_rows)) : align256((size_t)pp * s.slab_rows * sizeof(double));
  double* part = reinterpret_cast<double*>(wq); wq += align256((size_t)s.splits * pp * pp * sizeof(double));
  int* exps = reinterpret_cast<int*>(wq); wq += align256(gram_exps_len(pl) * sizeof(int));
  int* slot_hi = reinterpret_cast<int*>(wq); wq += align256(gram_exps_len(pl) * sizeof(int));
  double* r_ws = reinterpret_cast<double*>(wq); wq += align256((size_t)(s.slab_rows / kBuildRows) * pp * sizeof(double));
  double* r_acc = reinterpret_cast<double*>(wq);
  GRIEF_REQUIRE((size_t)pl->stride <= gram_exps_len(pl), "gram: table stride %d exceeds the slot scratch", pl->stride);
  const size_t part_doubles = (size_t)s.splits * pp * pp;
  if (n_pad == 0) {
    GRIEF_CUDA(cudaMemsetAsync(part, 0, part_doubles * sizeof(double), stream));
    if (r) GRIEF_CUDA(cudaMemsetAsync(r_acc, 0, (size_t)pp * sizeof(double), stream));
  }
  OzOpts oz_gram = oz_opts(pl, sd);
  oz_gram.diag_pair = 1;
  GemmOpts o;
  o.lower_only = true;
  o.splits = s.splits;
  o.c_split_stride = (int64_t)pp * pp;
  for (int64_t r0 = 0; r0 < n_pad; r0 += s.slab_rows) {
    const int64_t R = std::min(s.slab_rows, n_pad - r0);
    const double* Ts = T + (size_t)r0 * pl->stride;
    BuildArgs ba;
    ba.transposed = true;
    if (r) { ba.y = y + r0; ba.y_rows = std::max<int64_t>(0, n - r0); ba.r_ws = r_ws; }
    int rc;
    prof_begin(PROF_BUILD_T, stream);
    if (i8) {      // exponents from the slot maxima of the slab's tables, then the digit planes [sd][p_pad][R] (K = data rows)
      GRIEF_CUDA(cudaMemsetAsync(slot_hi, 0, gram_exps_len(pl) * sizeof(int), stream));
      const int64_t total = R * (int64_t)pl->stride;
      k_slot_hi<<<sms * 4, 256, (size_t)pl->stride * sizeof(int), stream>>>(Ts, pl->stride, total, slot_hi);
      const int n_e = (int)gram_exps_len(pl);
      k_col_exps<<<(n_e + 255) / 256, 256, 0, stream>>>(slot_hi, pl->d_sorted_slot, pl->n_groups, pp, n_e, exps, pl->d_err);
      ba.digits = true; ba.sd = sd; ba.exps = exps; ba.planes = planes; ba.ld = R;
      rc = launch_build(pl, Ts, R, ba, stream);
    } else {
      ba.out = PhiT; ba.ld = s.slab_rows;
      rc = launch_build(pl, Ts, R, ba, stream);
    }
    if (rc == GRIEF_OK && r) k_reduce_r<<<pp / 32, 512, 0, stream>>>(r_ws, (int)(R / kBuildRows), pp, r0 > 0 ? 1 : 0, r_acc);
    prof_end(PROF_BUILD_T, stream);
    if (rc != GRIEF_OK) return rc;
    prof_begin(PROF_GRAM, stream);
    if (i8)
      rc = ozaki_gemm(planes, pp, exps, pp, planes, pp, exps, pp, (int)R, part, pp, r0 > 0, true, s.splits, (int64_t)pp * pp, oz_gram, stream, launches);
    else
      rc = gemm_nt_ex(PhiT, s.slab_rows, PhiT, s.slab_rows, part, pp, pp, pp, (int)R, 1.0, r0 > 0 ? 1.0 : 0.0, o, stream, launches);
    prof_end(PROF_GRAM, stream);
    if (rc != GRIEF_OK) return rc;
    if (launches) *launches += (i8 ? 3 : 1) + (r ? 1 : 0);
  }
  k_gram_reduce<<<dim3(s.nb, s.nb), 256, 0, stream>>>(part, pl->d_perm, pp, s.splits, lda, A);
  if (r) k_unpermute_vec<<<(pp + 255) / 256, 256, 0, stream>>>(r_acc, pl->d_perm, pp, r);
  GRIEF_CUDA(cudaGetLastError());
  if (launches) *launches += r ? 2 : 1;
  return GRIEF_OK;
}

// ---- pass 2 / predictive variance: Z = Phi B ----
// B'[c][k] = B[perm[c]][perm[k]] (0 for padding rows / columns): both the K dimension of Z = Phi * B and the columns of Z are in
// the plan's SORTED column order, so neighbouring lanes of the consumers (k_contract, k_rowdot) share their leading slots.
__global__ void __launch_bounds__(256)
k_permute_sym(const double* __restrict__ B, int64_t ldb, const int* __restrict__ perm, int p_pad, double* __restrict__ out) {
  const int64_t total = (int64_t)p_pad * p_pad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / p_pad), k = (int)(e - (int64_t)c * p_pad);
    const int sc = perm[c], sk = perm[k];
    out[e] = (sc >= 0 && sk >= 0) ? B[(size_t)sc * ldb + sk] : 0.0;
  }
}

// Bperm: p_pad x p_pad doubles
int launch_permute_b(const Plan* pl, const double* B, int64_t ldb, double* Bperm, cudaStream_t stream) {
  const int64_t total = (int64_t)pl->p_pad * pl->p_pad;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
  k_permute_sym<<<blocks, 256, 0, stream>>>(B, ldb, pl->d_perm, pl->p_pad, Bperm);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// out[c] = perm[c] >= 0 ? scale * in[perm[c]] : 0   (a p-vector brought into the sorted column order)
__global__ void k_permute_vec(const double* __restrict__ in, double scale, const int* __restrict__ perm, int p_pad, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < p_pad) out[c] = perm[c] >= 0 ? scale * in[perm[c]] : 0.0;
}
int launch_permute_vec(const Plan* pl, const double* in, double scale, double* out, cudaStream_t stream) {
  k_permute_vec<<<(pl->p_pad + 255) / 256, 256, 0, stream>>>(in, scale, pl->d_perm, pl->p_pad, out);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// Scratch of the Z = Phi B product (carved out of the callers' workspaces)
size_t zgemm_scratch_bytes(const Plan* pl, int64_t slab_rows) {
  if (pl->opts.gemm_mode != 1) return align256((size_t)slab_rows * pl->p_pad * sizeof(double));                // Phi slab
  return align256(ozaki_plane_bytes(slab_rows, pl->p_pad)) + align256((size_t)slab_rows * sizeof(int)) +       // digits of the slab
         align256(ozaki_plane_bytes(pl->p_pad, pl->p_pad)) + align256(((size_t)pl->p_pad + 256) * sizeof(int));  // digits of B
}

struct ZScratch {
  double* Phi;
  int8_t* pa; int* ea;
  int8_t* pb; int* eb;
};
static ZScratch carve_zscratch(const Plan* pl, int64_t slab_rows, void* scratch) {
  ZScratch z{};
  char* q = reinterpret_cast<char*>(scratch);
  if (pl->opts.gemm_mode != 1) {
    z.Phi = reinterpret_cast<double*>(q);
  } else {
    z.pa = reinterpret_cast<int8_t*>(q); q += align256(ozaki_plane_bytes(slab_rows, pl->p_pad));
    z.ea = reinterpret_cast<int*>(q); q += align256((size_t)slab_rows * sizeof(int));
    z.pb = reinterpret_cast<int8_t*>(q); q += align256(ozaki_plane_bytes(pl->p_pad, pl->p_pad));
    z.eb = reinterpret_cast<int*>(q);
  }
  return z;
}

// Once per evaluation, after launch_permute_b: digit planes of B' for the INT8 path (no-op on the DMMA path).
// `digits`: int8 digits per operand of this product (PlanOpts::digits_z for the gradient pass, digits_var for the predictive variance).
int launch_zgemm_prepare(const Plan* pl, const double* Bperm, int64_t slab_rows_max, void* scratch, int digits, cudaStream_t stream) {
  if (pl->opts.gemm_mode != 1) return GRIEF_OK;
  ZScratch z = carve_zscratch(pl, slab_rows_max, scratch);
  return ozaki_slice(Bperm, pl->p_pad, pl->p_pad, pl->p_pad, z.eb, (pl->p_pad + 255) / 256 * 256, z.pb, digits, pl->d_err, stream);
}

// Zt (p_pad x ldz, TRANSPOSED: sorted column c of Z = Phi(slab) * B is row c of Zt, the slab's data rows are contiguous) -- the
// layout its consumers (k_contract_rows, k_rowdot_t: lane = data row) read with full coalescing.  B symmetric, given as Bperm
// (p_pad x p_pad, launch_permute_b).  scratch: zgemm_scratch_bytes(pl, slab_rows_max) bytes, prepared by launch_zgemm_prepare.
int launch_zgemm(const Plan* pl, const double* T_slab, int64_t slab_rows, const double* Bperm, void* scratch, int64_t slab_rows_max, double* Zt,
                 int64_t ldz, int digits, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(slab_rows % kBuildRows == 0, "zgemm: slab_rows=%lld is not a multiple of %d", (long long)slab_rows, kBuildRows);
  GRIEF_REQUIRE(ldz >= slab_rows, "zgemm: ldz=%lld must be >= slab_rows=%lld", (long long)ldz, (long long)slab_rows);
  if (slab_rows == 0) return GRIEF_OK;
  const bool i8 = pl->opts.gemm_mode == 1;
  ZScratch z = carve_zscratch(pl, slab_rows_max, scratch);
  BuildArgs ba;
  if (i8) {      // row exponents + digit planes [sd][slab_rows][p_pad] straight from the tables
    ba.digits = true; ba.sd = digits; ba.exps = z.ea; ba.planes = z.pa; ba.plane_stride = (size_t)slab_rows * pl->p_pad;
  } else {
    ba.out = z.Phi; ba.ld = pl->p_pad;
  }
  prof_begin(PROF_BUILD, stream);
  int rc = launch_build(pl, T_slab, slab_rows, ba, stream);
  prof_end(PROF_BUILD, stream);
  if (rc != GRIEF_OK) return rc;
  prof_begin(PROF_ZGEMM, stream);
  if (i8) {
    GRIEF_REQUIRE(pl->p_pad <= kOzakiKRange, "zgemm: p_pad=%d exceeds the INT8 path's K range of %d", pl->p_pad, kOzakiKRange);
    OzOpts oo = oz_opts(pl, digits);
    oo.store_t = 1;
    rc = ozaki_gemm(z.pa, slab_rows, z.ea, (int)slab_rows, z.pb, pl->p_pad, z.eb, pl->p_pad, pl->p_pad, Zt, ldz, false, false, 1, 0, oo, stream, launches);
  } else {
    GemmOpts o;
    o.store_t = true;
    rc = gemm_nt_ex(z.Phi, pl->p_pad, Bperm, pl->p_pad, Zt, ldz, (int)slab_rows, pl->p_pad, pl->p_pad, 1.0, 0.0, o, stream, launches);
  }
  prof_end(PROF_ZGEMM, stream);
  if (rc == GRIEF_OK && launches) *launches += 1;
  return rc;
}

}  // namespace grief
