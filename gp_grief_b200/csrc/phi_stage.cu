// The two O(n p^2) products of an evaluation, staged:  the basis matrix Phi of a slab of data rows is built by a row kernel into
// HBM (as int8 digit planes, or as FP64), then a GEMM consumes it (ozaki.cu on the INT8 tensor cores, or dense.cu on FP64 DMMA).
//
//   pass 1 (models/gp_grief_model.py:148-149)   A = Phi^T Phi :  Phi^T slab (p_pad x R, K = data rows contiguous)
//                                               -> SYRK on the lower tiles, K split over gridDim.z; r = Phi^T y and the row maxima
//                                                  of |Phi| fall out of the builder's sweep
//   pass 2 / predictive var.    Z = Phi B     :  Phi slab (R x p_pad, K = sorted columns contiguous)
//                                               -> GEMM against the column-permuted symmetric B, stored transposed; the residual
//                                                  a = (y - Phi b) / sigma^2 falls out of the builder's sweep
//
// Two arithmetic modes (per plan, PlanOpts): INT8 (default) -- the builders emit power-of-two row scales and 3..7 int8 digit planes
// and k_ozaki multiplies them exactly; FP64 -- the builders emit the FP64 slab for k_gemm_nt.
//
// Why staged and not fused (round-1 measurements, profiles/r01_gram_design_notes.md): DMUL, DFMA and DMMA share ONE FP64
// pipe per SM sub-partition.  With the Phi tiles built inside the GEMM CTAs the builder's DMULs queue behind the DMMAs
// (32 % of warp samples in stall_math) and the kernels stop at 25-28 TFLOP/s; the same DMMA loop fed by TMA alone runs at
// 36 TFLOP/s.  The slab costs 4-6 B written + read per element of Phi, a few % of the GEMM time, on an HBM that is otherwise
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

// digits d_s of q = rint(v * scale), scale = 2^(8 sd - 2 - e):  v 2^-e = sum_s d_s 2^(-6 - 8 s), d_s in [-128, 127], s < sd, as the bytes
// of one 64-bit word: digit s (0 = most significant) is byte 6 - s.  Balanced base-256 digits without a carry chain: add 128 to every
// byte position (q + 0x80..80 is positive and below 2^(8 sd)), then byte k of the sum, minus 128 (= XOR 0x80 read as int8), is the
// digit of 256^k.
// sd <= 6: |q| <= 2^46, so q is read off the mantissa of v * scale + 1.5 * 2^52 (one DFMA; the rounding to an integer is the
// same round-to-nearest-even as the conversion instruction, which takes a trip through a slower pipe); sd = 7 converts.
__device__ __forceinline__ unsigned long long digit_word(double v, double scale, int sd) {
  const unsigned long long bias = 0x0080808080808080ull >> (8 * (7 - sd));
  long long q;
  if (sd <= 6) q = __double_as_longlong(fma(v, scale, 6755399441055744.0)) - 0x4338000000000000ll;
  else q = __double2ll_rn(v * scale);
  return (((unsigned long long)q + bias) ^ bias) << (8 * (7 - sd));
}
// byte B (0..6) of four digit words, packed into one 32-bit word (word i -> byte i): 3 PRMT
template <int B>
__device__ __forceinline__ uint32_t gather_byte(const unsigned long long (&w)[4]) {
  const uint32_t x0 = B < 4 ? (uint32_t)w[0] : (uint32_t)(w[0] >> 32), x1 = B < 4 ? (uint32_t)w[1] : (uint32_t)(w[1] >> 32);
  const uint32_t x2 = B < 4 ? (uint32_t)w[2] : (uint32_t)(w[2] >> 32), x3 = B < 4 ? (uint32_t)w[3] : (uint32_t)(w[3] >> 32);
  constexpr uint32_t sel = (uint32_t)(B & 3) | ((uint32_t)(4 + (B & 3)) << 4);
  return __byte_perm(__byte_perm(x0, x1, sel), __byte_perm(x2, x3, sel), 0x5410);
}

// Phi^T slab, c = sorted column, one CTA per 128 table rows.  Lane = data row; each of the 16 warps owns p_pad / 16 consecutive
// sorted columns and walks them for the four 32-row groups at once.  Consecutive sorted columns share their leading slots: the
// product P of the first G-1 factors stays in a register per row group and is rebuilt only where sorted_level says a leading factor
// changed (warp-uniform branch) -- ~1.5 gathers and 1.3 DMULs per element.
//   DIG = false: out[c * ld + row]                                  (FP64 slab)
//   DIG = true:  planes[s][c][k] (row stride ld bytes, plane stride p_pad * ld), scaled by 2^(8 sd - 2 - exps[c]); within every block
//                of 128 data rows the K position is k = 4 * lane + row_group (data row = 32 * row_group + lane): a lane then stores
//                the digit s of its four rows as ONE 32-bit word and a warp writes 128 contiguous bytes per (column, digit).  The
//                Gram sums over K, and both operands of that product are this same slab: any fixed permutation of K is exact.
// y != nullptr: the values are already in registers, so r_ws[block][c] = sum over the block's rows of y[row] * Phi[row][c] is formed
// here (fixed order: the four row groups in sequence, lanes by butterfly) instead of rebuilding Phi a second time for Phi^T y.
template <int G, bool DIG>
__global__ void __launch_bounds__(kBuildThreads) k_build_phi_t(const double* __restrict__ T, int stride,
                                                               const uint16_t* __restrict__ sorted_slot,
                                                               const uint8_t* __restrict__ sorted_level, int p_pad,
                                                               double* __restrict__ out, int64_t ld, const int* __restrict__ exps,
                                                               int8_t* __restrict__ planes, int sd, const double* __restrict__ y,
                                                               int64_t y_rows, double* __restrict__ r_ws, int* __restrict__ row_hi_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sT = reinterpret_cast<double*>(smem_raw + 128);
  int* s_rowhi = reinterpret_cast<int*>(sT + (size_t)kBuildRows * stride);      // [128] row maxima (high words), row_hi_out only
  if (row_hi_out != nullptr && threadIdx.x < kBuildRows) s_rowhi[threadIdx.x] = 0;
  init_stage_barrier(bar);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cpw = p_pad / (kBuildThreads / 32);        // columns per warp (p_pad is a multiple of 128)
  const int c_begin = warp * cpw;
  constexpr int NB = 8, RG = kBuildRows / 32;
  static_assert(RG == 4, "four row groups of 32 rows");
  const int64_t rb = blockIdx.x;
  stage_table_rows(sT, bar, T, stride, rb, 0);
  const size_t grow0 = (size_t)rb * kBuildRows + lane;
  double yv[RG];                                        // y of this lane's row in each row group
#pragma unroll
  for (int rg = 0; rg < RG; ++rg) yv[rg] = (y != nullptr && (int64_t)grow0 + 32 * rg < y_rows) ? y[grow0 + 32 * rg] : 0.0;
  const double* trow[RG];
#pragma unroll
  for (int rg = 0; rg < RG; ++rg) trow[rg] = sT + (size_t)(rg * 32 + lane) * stride;
  double P[RG] = {1.0, 1.0, 1.0, 1.0};                  // running prefix products, one per row group
  int rhi[RG] = {0, 0, 0, 0};                           // row maxima over this warp's columns
  const size_t plane_stride = (size_t)p_pad * ld;
  for (int c0 = c_begin; c0 < c_begin + cpw; c0 += NB) {
    int lv[NB], sl[NB];
    double dot[NB];
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      lv[e] = (c0 + e == c_begin) ? 0 : (int)__ldg(sorted_level + c0 + e);
      sl[e] = __ldg(sorted_slot + (size_t)(c0 + e) * G + (G - 1));
    }
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      if constexpr (G > 1) {
        if (lv[e] < G - 1) {                             // warp-uniform: a leading factor changed
          int lead[G - 1];
#pragma unroll
          for (int g = 0; g < G - 1; ++g) lead[g] = __ldg(sorted_slot + (size_t)(c0 + e) * G + g);
#pragma unroll
          for (int rg = 0; rg < RG; ++rg) {
            double q = trow[rg][lead[0]];
#pragma unroll
            for (int g = 1; g < G - 1; ++g) q *= trow[rg][lead[g]];
            P[rg] = q;
          }
        }
      }
      double v[RG];
#pragma unroll
      for (int rg = 0; rg < RG; ++rg) v[rg] = G > 1 ? P[rg] * trow[rg][sl[e]] : trow[rg][sl[e]];
#pragma unroll
      for (int rg = 0; rg < RG; ++rg) rhi[rg] = max(rhi[rg], abs_hi(v[rg]));
      if constexpr (!DIG) {
#pragma unroll
        for (int rg = 0; rg < RG; ++rg) out[(size_t)(c0 + e) * ld + grow0 + rg * 32] = v[rg];
      } else {
        const double scale = digit_scale(sd, __ldg(exps + c0 + e));
        unsigned long long w[RG];
#pragma unroll
        for (int rg = 0; rg < RG; ++rg) w[rg] = digit_word(v[rg], scale, sd);
        uint32_t* dst = reinterpret_cast<uint32_t*>(planes + (size_t)(c0 + e) * ld + (size_t)rb * kBuildRows) + lane;
        const size_t ps4 = plane_stride / 4;             // ld is a multiple of 128
        dst[0 * ps4] = gather_byte<6>(w);
        dst[1 * ps4] = gather_byte<5>(w);
        dst[2 * ps4] = gather_byte<4>(w);
        if (sd > 3) dst[3 * ps4] = gather_byte<3>(w);
        if (sd > 4) dst[4 * ps4] = gather_byte<2>(w);
        if (sd > 5) dst[5 * ps4] = gather_byte<1>(w);
        if (sd > 6) dst[6 * ps4] = gather_byte<0>(w);
      }
      double dsum = 0.0;
#pragma unroll
      for (int rg = 0; rg < RG; ++rg) dsum = fma(yv[rg], v[rg], dsum);
      dot[e] = dsum;
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
  if (row_hi_out != nullptr) {                           // warp-uniform: max over the 16 warps' column ranges, one int per row
#pragma unroll
    for (int rg = 0; rg < RG; ++rg) atomicMax(s_rowhi + rg * 32 + lane, rhi[rg]);
    __syncthreads();
    if (threadIdx.x < kBuildRows) row_hi_out[(size_t)rb * kBuildRows + threadIdx.x] = s_rowhi[threadIdx.x];
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

// pack: plan->d_sorted_pack, (G + 3) / 4 words per sorted column, key position k in byte 3 - k % 4 of word k / 4
template <int G>
__device__ __forceinline__ int pack_slot(const uint32_t* w, int k) { return (int)((w[k >> 2] >> (8 * (3 - (k & 3)))) & 0xFFu); }

// Phi slab, row-major [row][sorted column], one CTA per 128 table rows.  A lane owns FOUR consecutive sorted columns (one 16-byte
// load brings their packed slots), a warp 8 rows: per row and digit plane the lane stores one 32-bit word and the warp 128 contiguous
// bytes (FP64 mode: 1 KB).  Consecutive sorted columns mostly share their leading slots, so columns 1..3 of a lane reuse the prefix
// product of their predecessor when the packed words agree above the last key position.
//   DIG = false: out[row * ldo + c]
//   DIG = true:  planes[s][row][c], scaled by 2^(8 sd - 2 - exps[row]).  The row scale needs max_c |Phi[row][c]|: taken from
//                row_hi (recorded by the pass-1 builder of the same tables, k_build_phi_t) or, without it, from a first sweep.
// bvec != nullptr: f[row] = sum_c Phi[row][c] * bvec[c] is accumulated along the way (the values are in registers) and
// a_out[row] = (y[row] - f[row]) * inv_noise is written -- the residual scaled by the noise that the contraction kernel needs
// (grad.cu); lanes are summed by butterfly (fixed order).
template <int G, bool DIG>
__global__ void __launch_bounds__(kBuildThreads) k_build_phi(const double* __restrict__ T, int stride,
                                                             const uint32_t* __restrict__ pack, int p_pad,
                                                             double* __restrict__ out, int64_t ldo, int* __restrict__ exps,
                                                             int8_t* __restrict__ planes, size_t plane_stride, int sd, int* __restrict__ err,
                                                             const double* __restrict__ bvec, const double* __restrict__ y, int64_t y_rows,
                                                             double inv_noise, double* __restrict__ a_out,
                                                             const int* __restrict__ row_hi) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sT = reinterpret_cast<double*>(smem_raw + 128);
  init_stage_barrier(bar);
  stage_table_rows(sT, bar, T, stride, blockIdx.x, 0);
  constexpr int NW = (G + 3) / 4;
  constexpr int RW = kBuildRows / (kBuildThreads / 32);                // 8 rows per warp
  constexpr uint32_t kLastMask = 0xFFu << (8 * (3 - ((G - 1) & 3)));   // the last key position inside the last packed word
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* tbase = sT + (size_t)warp * RW * stride;
  const size_t row0 = (size_t)blockIdx.x * kBuildRows + warp * RW;
  const bool want_f = bvec != nullptr;

  // values of this lane's four columns for table row `trow`; slots come from the packed words wd[q][w]
  auto four_values = [&](const double* trow, const uint32_t (&wd)[4][NW], double (&v)[4]) {
    double P = 1.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if constexpr (G > 1) {
        bool reuse = q > 0;
        if (q > 0) {
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            uint32_t x = wd[q][w] ^ wd[q - 1][w];
            if (w == NW - 1) x &= ~kLastMask;
            reuse = reuse && (x == 0);
          }
        }
        if (!reuse) {
          P = trow[pack_slot<G>(wd[q], 0)];
#pragma unroll
          for (int g = 1; g < G - 1; ++g) P *= trow[pack_slot<G>(wd[q], g)];
        }
        v[q] = P * trow[pack_slot<G>(wd[q], G - 1)];
      } else {
        v[q] = trow[pack_slot<G>(wd[q], 0)];
      }
    }
  };
  auto load_words = [&](int cbase, uint32_t (&wd)[4][NW]) {
    const uint4* src = reinterpret_cast<const uint4*>(pack + (size_t)(cbase + 4 * lane) * NW);
#pragma unroll
    for (int i = 0; i < NW; ++i) {                                     // 4 columns x NW words = NW 16-byte loads
      const uint4 t = __ldg(src + i);
      const uint32_t flat[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) wd[(4 * i + j) / NW][(4 * i + j) % NW] = flat[j];
    }
  };

  double scale[RW];
  if constexpr (DIG) {
    int mh[RW];
    if (row_hi != nullptr) {
#pragma unroll
      for (int r = 0; r < RW; ++r) mh[r] = __ldg(row_hi + row0 + r);
    } else {                                                           // first sweep: row maxima
#pragma unroll
      for (int r = 0; r < RW; ++r) mh[r] = 0;
      for (int cbase = 0; cbase < p_pad; cbase += 128) {
        uint32_t wd[4][NW];
        load_words(cbase, wd);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          double v[4];
          four_values(tbase + (size_t)r * stride, wd, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) mh[r] = max(mh[r], abs_hi(v[q]));
        }
      }
#pragma unroll
      for (int r = 0; r < RW; ++r) mh[r] = __reduce_max_sync(0xffffffffu, mh[r]);
    }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int e = exp_from_hi(mh[r]);
      if (lane == 0) {
        exps[row0 + r] = e;
        if (mh[r] >= 0x7ff00000) atomicExch(err, 4);                   // Inf / NaN in the row
      }
      scale[r] = digit_scale(sd, e);
    }
  }
  double facc[RW];
#pragma unroll
  for (int r = 0; r < RW; ++r) facc[r] = 0.0;
  for (int cbase = 0; cbase < p_pad; cbase += 128) {
    uint32_t wd[4][NW];
    load_words(cbase, wd);
    double bq[4] = {0.0, 0.0, 0.0, 0.0};
    if (want_f) {
      const double2* bsrc = reinterpret_cast<const double2*>(bvec + cbase + 4 * lane);
      const double2 b01 = __ldg(bsrc), b23 = __ldg(bsrc + 1);
      bq[0] = b01.x; bq[1] = b01.y; bq[2] = b23.x; bq[3] = b23.y;
    }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      double v[4];
      four_values(tbase + (size_t)r * stride, wd, v);
      if (want_f) facc[r] = fma(v[3], bq[3], fma(v[2], bq[2], fma(v[1], bq[1], fma(v[0], bq[0], facc[r]))));
      if constexpr (!DIG) {
        double2* dst = reinterpret_cast<double2*>(out + (row0 + r) * ldo + cbase + 4 * lane);
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
      } else {
        unsigned long long w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = digit_word(v[q], scale[r], sd);
        uint32_t* dst = reinterpret_cast<uint32_t*>(planes + (row0 + r) * (size_t)p_pad + cbase) + lane;
        const size_t ps4 = plane_stride / 4;
        dst[0 * ps4] = gather_byte<6>(w);
        dst[1 * ps4] = gather_byte<5>(w);
        dst[2 * ps4] = gather_byte<4>(w);
        if (sd > 3) dst[3 * ps4] = gather_byte<3>(w);
        if (sd > 4) dst[4 * ps4] = gather_byte<2>(w);
        if (sd > 5) dst[5 * ps4] = gather_byte<1>(w);
        if (sd > 6) dst[6 * ps4] = gather_byte<0>(w);
      }
    }
  }
  if (want_f) {                                                        // warp-uniform
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const double ft = warp_sum(facc[r]);
      const size_t grow = row0 + r;
      if (lane == 0) a_out[grow] = (int64_t)grow < y_rows ? (y[grow] - ft) * inv_noise : 0.0;
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
  const double* y = nullptr; int64_t y_rows = 0; double* r_ws = nullptr;   // transposed: fused Phi^T y partials (y, y_rows, r_ws)
  const double* bvec = nullptr; double inv_noise = 0.0; double* a_out = nullptr;   // row-major: a = (y - Phi bvec) * inv_noise (y, y_rows too)
  int* row_hi_out = nullptr;                 // transposed: per-row maxima (high words) of |Phi| for a later row-major build of the same tables
  const int* row_hi = nullptr;               // row-major digits: row maxima recorded earlier (skips the maximum sweep)
};

template <int G>
static int launch_build_g(const Plan* pl, const double* T, int64_t rows, const BuildArgs& a, cudaStream_t stream) {
  const size_t smem = 128 + (size_t)kBuildRows * pl->stride * sizeof(double) + (a.transposed ? kBuildRows * sizeof(int) : 0);
  const unsigned grid = (unsigned)(rows / kBuildRows);
  GRIEF_REQUIRE(smem <= 227 * 1024, "build_phi: %zu bytes of shared memory", smem);
#define GRIEF_BT(DIG_)                                                                                                              \
  do {                                                                                                                              \
    GRIEF_CUDA(cudaFuncSetAttribute(k_build_phi_t<G, DIG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    k_build_phi_t<G, DIG_><<<grid, kBuildThreads, smem, stream>>>(T, pl->stride, pl->d_sorted_slot, pl->d_sorted_level, pl->p_pad,  \
                                                                 a.out, a.ld, a.exps, a.planes, a.sd, a.y, a.y_rows, a.r_ws,        \
                                                                 a.row_hi_out);                                                     \
  } while (0)
#define GRIEF_BN(DIG_)                                                                                                              \
  do {                                                                                                                              \
    GRIEF_CUDA(cudaFuncSetAttribute(k_build_phi<G, DIG_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    k_build_phi<G, DIG_><<<grid, kBuildThreads, smem, stream>>>(T, pl->stride, pl->d_sorted_pack, pl->p_pad, a.out, a.ld, a.exps,   \
                                                               a.planes, a.plane_stride, a.sd, pl->d_err, a.bvec, a.y, a.y_rows,    \
                                                               a.inv_noise, a.a_out, a.row_hi);                                     \
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
  int64_t budget_rows = (int64_t)(pl->opts.slab_budget / ((size_t)p_pad * 8)) / kBuildRows * kBuildRows;
  const int64_t wave_rows = (int64_t)sms * kBuildRows;              // the builder runs one CTA of 128 rows per SM: whole waves per slab
  if (budget_rows >= wave_rows) budget_rows = budget_rows / wave_rows * wave_rows;
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
// rowmax_out != nullptr (n_pad ints): the high word of max_c |Phi[row][c]| of every table row, for launch_zgemm on the same tables.
int launch_gram(const Plan* pl, const double* T, int64_t n_pad, double* A, int64_t lda, const double* y, int64_t n, double* r,
                int* rowmax_out, void* workspace, size_t ws_bytes, int sms, cudaStream_t stream, int* launches) {
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
  oz_gram.tile_order = pl->d_gram_order; oz_gram.n_tile_order = pl->n_gram_order;
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
    if (rowmax_out) ba.row_hi_out = rowmax_out + r0;
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
// layout its consumers (k_contract_back, k_rowdot_t: lane = data row) read with full coalescing.  B symmetric, given as Bperm
// (p_pad x p_pad, launch_permute_b).  scratch: zgemm_scratch_bytes(pl, slab_rows_max) bytes, prepared by launch_zgemm_prepare.
struct ResidualArgs {      // optional by-product of the slab builder: a = (y - Phi bvec) / noise_var for the slab's rows
  const double* bvec = nullptr; const double* y = nullptr; int64_t y_rows = 0; double inv_noise = 0.0; double* a_out = nullptr;
  const int* row_hi = nullptr;      // optional input: row maxima of the slab's rows recorded by launch_gram
};
int launch_zgemm(const Plan* pl, const double* T_slab, int64_t slab_rows, const double* Bperm, void* scratch, int64_t slab_rows_max, double* Zt,
                 int64_t ldz, int digits, const ResidualArgs* res, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(slab_rows % kBuildRows == 0, "zgemm: slab_rows=%lld is not a multiple of %d", (long long)slab_rows, kBuildRows);
  GRIEF_REQUIRE(ldz >= slab_rows, "zgemm: ldz=%lld must be >= slab_rows=%lld", (long long)ldz, (long long)slab_rows);
  if (slab_rows == 0) return GRIEF_OK;
  const bool i8 = pl->opts.gemm_mode == 1;
  ZScratch z = carve_zscratch(pl, slab_rows_max, scratch);
  BuildArgs ba;
  if (res) { ba.bvec = res->bvec; ba.y = res->y; ba.y_rows = res->y_rows; ba.inv_noise = res->inv_noise; ba.a_out = res->a_out; ba.row_hi = res->row_hi; }
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
