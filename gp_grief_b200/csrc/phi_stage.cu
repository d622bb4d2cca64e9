// The two O(n p^2) products of an evaluation, staged:  the basis matrix Phi of a slab of data rows is built by a
// bandwidth-bound kernel into HBM, then a plain TMA-fed FP64 DMMA GEMM (dense.cu, k_gemm_nt) consumes it.
//
//   pass 1 (models/gp_grief_model.py:148-149)   A = Phi^T Phi :  Phi^T slab (p_pad x R, data rows contiguous)
//                                               -> SYRK on the lower tiles, K = data rows split over gridDim.z
//   pass 2 / predictive var.    Z = Phi B     :  Phi slab (R x p_pad, sorted columns contiguous)
//                                               -> GEMM against the column-permuted symmetric B
//
// Two arithmetic modes (grief_set_gemm_mode): 1 (default) -- the builders emit power-of-two row scales and seven int8 digit planes
// and k_ozaki (ozaki.cu) multiplies them on the tcgen05 INT8 tensor cores; 0 -- the builders emit the FP64 slab for k_gemm_nt.
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
// first the binary exponents of the operand rows' maxima, then the 7 balanced 8-bit digits of every element.
enum BuildOut : int { OUT_F64 = 0, OUT_EXP = 1, OUT_DIGITS = 2 };
constexpr int kDigits = 7;

// high word of |v|: orders like |v|, and its exponent field is all the INT8 path needs of a row maximum
__device__ __forceinline__ int abs_hi(double v) { return __double2hiint(v) & 0x7fffffff; }
// frexp exponent e (|x| < 2^e) from the high word of the row maximum (0 for an all-zero row)
__device__ __forceinline__ int exp_from_hi(int hi) { return hi > 0 ? min(max((hi >> 20) - 1022, -900), 1024) : 0; }   // clamp: 2^(54-e) stays finite
// digits d_s of q = trunc(v * scale), scale = 2^(54 - e):  v 2^-e = sum_s d_s 2^(-6 - 8 s), d_s in [-128, 127].
// Balanced base-256 digits without a carry chain: add 128 to every byte position (q + 0x80..80 is positive and below 2^56),
// then byte k of the sum, minus 128 (= XOR 0x80 read as int8), is the digit of 256^k.
__device__ __forceinline__ void store_digits(double v, double scale, int8_t* __restrict__ base, size_t plane_stride) {
  const unsigned long long w = ((unsigned long long)__double2ll_rz(v * scale) + 0x0080808080808080ull) ^ 0x0080808080808080ull;
  const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
  base[0 * plane_stride] = (int8_t)(hi >> 16);       // digit 0 = byte 6 (most significant)
  base[1 * plane_stride] = (int8_t)(hi >> 8);
  base[2 * plane_stride] = (int8_t)hi;
  base[3 * plane_stride] = (int8_t)(lo >> 24);
  base[4 * plane_stride] = (int8_t)(lo >> 16);
  base[5 * plane_stride] = (int8_t)(lo >> 8);
  base[6 * plane_stride] = (int8_t)lo;
}

// Phi^T slab, c = sorted column.  Lane = data row; a warp walks a run of consecutive sorted columns, which share their
// leading slots: the product P of the first G-1 factors stays in a register and is rebuilt only where sorted_level says a
// leading factor changed (warp-uniform branch) -- ~1.5 gathers and 1.3 DMULs per element.
//   OUT_F64:    out[c * ld + row]
//   OUT_EXP:    atomicMax(col_hi[c], high word of |Phi[row][c]|) over the slab's rows (one REDUX per column and warp)
//   OUT_DIGITS: planes[s][c][row] (row stride ld bytes, plane stride p_pad * ld), scaled by 2^(54 - exps[c])
template <int G, int MODE>
__global__ void __launch_bounds__(kBuildThreads) k_build_phi_t(const double* __restrict__ T, int stride,
                                                               const uint16_t* __restrict__ sorted_slot,
                                                               const uint8_t* __restrict__ sorted_level, int p_pad, int n_blocks,
                                                               double* __restrict__ out, int64_t ld, int* __restrict__ col_hi,
                                                               const int* __restrict__ exps, int8_t* __restrict__ planes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sT = reinterpret_cast<double*>(smem_raw + 128);
  int* s_hi = reinterpret_cast<int*>(smem_raw + 128 + (size_t)kBuildRows * stride * sizeof(double));   // OUT_EXP: p_pad column maxima
  if constexpr (MODE == OUT_EXP)
    for (int c = threadIdx.x; c < p_pad; c += kBuildThreads) s_hi[c] = 0;
  init_stage_barrier(bar);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = (warp & 3) * 32 + lane;
  const double* trow = sT + (size_t)row * stride;
  const int cpq = p_pad / 4;                           // four warps share a row group, a quarter of the columns each
  const int c_begin = (warp >> 2) * cpq;
  constexpr int NB = 8;
  int use = 0;
  for (int rb = blockIdx.x; rb < n_blocks; rb += gridDim.x, ++use) {   // persistent over 128-row blocks
    stage_table_rows(sT, bar, T, stride, rb, use);
    const size_t grow = (size_t)rb * kBuildRows + row;
    double P = 1.0;
    for (int c0 = c_begin; c0 < c_begin + cpq; c0 += NB) {
      double last[NB];
      int lv[NB];
#pragma unroll
      for (int e = 0; e < NB; ++e) {
        lv[e] = (c0 + e == c_begin) ? 0 : (int)__ldg(sorted_level + c0 + e);
        last[e] = trow[__ldg(sorted_slot + (size_t)(c0 + e) * G + (G - 1))];
      }
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
      if constexpr (MODE == OUT_F64) {
#pragma unroll
        for (int e = 0; e < NB; ++e) out[(size_t)(c0 + e) * ld + grow] = last[e];
      } else if constexpr (MODE == OUT_EXP) {
#pragma unroll
        for (int e = 0; e < NB; ++e) {
          const int m = __reduce_max_sync(0xffffffffu, abs_hi(last[e]));
          if (lane == 0 && m > s_hi[c0 + e]) atomicMax(s_hi + c0 + e, m);     // shared-memory maximum over this CTA's rows
        }
      } else {
#pragma unroll
        for (int e = 0; e < NB; ++e) {
          const double scale = __hiloint2double((1023 + 54 - __ldg(exps + c0 + e)) << 20, 0);
          store_digits(last[e], scale, planes + (size_t)(c0 + e) * ld + grow, (size_t)p_pad * ld);
        }
      }
    }
  }
  if constexpr (MODE == OUT_EXP) {                     // one global maximum per column and CTA
    __syncthreads();
    for (int c = threadIdx.x; c < p_pad; c += kBuildThreads)
      if (s_hi[c] > 0) atomicMax(col_hi + c, s_hi[c]);
  }
}

// exps[c] from the accumulated high words (and reset them for the next slab)
__global__ void k_exps_from_hi(int* __restrict__ hi, int n, int n_pad, int* __restrict__ exps, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (hi[i] >= 0x7ff00000) atomicExch(err, 4);      // Inf / NaN in the column
    exps[i] = exp_from_hi(hi[i]);
    hi[i] = 0;
  }
  else if (i < n_pad) exps[i] = 0;
}

// Phi slab, row-major, lane = sorted column (coalesced stores); the G slots of a column are loaded once and reused for the
// warp's eight rows.
//   OUT_F64:    out[row * ldo + c]
//   OUT_DIGITS: exps[row] from the row maximum (first sweep, registers only), then planes[s][row][c] (second sweep)
template <int G, int MODE>
__global__ void __launch_bounds__(kBuildThreads) k_build_phi(const double* __restrict__ T, int stride,
                                                             const uint16_t* __restrict__ sorted_slot, int p_pad,
                                                             double* __restrict__ out, int64_t ldo, int* __restrict__ exps,
                                                             int8_t* __restrict__ planes, size_t plane_stride, int* __restrict__ err) {
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
  if constexpr (MODE == OUT_DIGITS) {
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
      scale[r] = __hiloint2double((1023 + 54 - e) << 20, 0);
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
    if constexpr (MODE == OUT_F64) {
#pragma unroll
      for (int r = 0; r < RW; ++r) out[(row0 + r) * ldo + c] = v[r];
    } else {
#pragma unroll
      for (int r = 0; r < RW; ++r) store_digits(v[r], scale[r], planes + (row0 + r) * (size_t)p_pad + c, plane_stride);
    }
  }
}

struct BuildArgs {
  bool transposed = false;
  int mode = OUT_F64;
  double* out = nullptr; int64_t ld = 0;     // FP64 slab (OUT_F64); ld also = bytes per plane row of the transposed digits
  int* col_hi = nullptr;                     // OUT_EXP (transposed)
  int* exps = nullptr;                       // read (transposed OUT_DIGITS) or written (row-major OUT_DIGITS)
  int8_t* planes = nullptr; size_t plane_stride = 0;
};

template <int G>
static int launch_build_g(const Plan* pl, const double* T, int64_t rows, const BuildArgs& a, cudaStream_t stream) {
  const size_t smem_rows = 128 + (size_t)kBuildRows * pl->stride * sizeof(double);
  const int n_blocks = (int)(rows / kBuildRows);
  const unsigned grid = (unsigned)n_blocks;
#define GRIEF_BT(MODE_)                                                                                                             \
  do {                                                                                                                              \
    const size_t smem = smem_rows + (MODE_ == OUT_EXP ? (size_t)pl->p_pad * sizeof(int) : 0);                                       \
    const unsigned g = MODE_ == OUT_EXP ? std::min<unsigned>(grid, 148u * 2u) : grid;                                                \
    GRIEF_REQUIRE(smem <= 227 * 1024, "build_phi_t: %zu bytes of shared memory", smem);                                             \
    GRIEF_CUDA(cudaFuncSetAttribute(k_build_phi_t<G, MODE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
    k_build_phi_t<G, MODE_><<<g, kBuildThreads, smem, stream>>>(T, pl->stride, pl->d_sorted_slot, pl->d_sorted_level, pl->p_pad,    \
                                                               n_blocks, a.out, a.ld, a.col_hi, a.exps, a.planes);                  \
  } while (0)
#define GRIEF_BN(MODE_)                                                                                                             \
  do {                                                                                                                              \
    const size_t smem = smem_rows;                                                                                                  \
    GRIEF_CUDA(cudaFuncSetAttribute(k_build_phi<G, MODE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
    k_build_phi<G, MODE_><<<grid, kBuildThreads, smem, stream>>>(T, pl->stride, pl->d_sorted_slot, pl->p_pad, a.out, a.ld, a.exps,  \
                                                                a.planes, a.plane_stride, ozaki_err_flag());                        \
  } while (0)
  if (a.transposed) {
    if (a.mode == OUT_F64) GRIEF_BT(OUT_F64);
    else if (a.mode == OUT_EXP) GRIEF_BT(OUT_EXP);
    else GRIEF_BT(OUT_DIGITS);
  } else {
    if (a.mode == OUT_F64) GRIEF_BN(OUT_F64);
    else GRIEF_BN(OUT_DIGITS);
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

// ---- pass 1: A = Phi^T Phi ----
struct GramSchedule {
  int nb, n_tiles, splits;
  int64_t slab_rows;
};

static int g_gemm_mode = 1;            // INT8 tensor-core arithmetic by default; 0 = FP64 DMMA
int gemm_mode() { return g_gemm_mode; }
void set_gemm_mode(int mode) { g_gemm_mode = mode ? 1 : 0; }
constexpr int kOzakiKRange = 16384;   // values of K per int32 accumulation in k_ozaki

static size_t g_slab_budget_bytes = (size_t)4 << 30;      // bytes of Phi^T staged per pass-1 slab
void set_slab_budget(size_t bytes) { g_slab_budget_bytes = bytes ? bytes : ((size_t)4 << 30); }

GramSchedule gram_schedule(int p_pad, int64_t n_pad, int sms) {
  GramSchedule s;
  s.nb = p_pad / kTileN;
  s.n_tiles = s.nb * (s.nb + 1) / 2;
  const int64_t budget_rows = (int64_t)(g_slab_budget_bytes / ((size_t)p_pad * 8)) / kBuildRows * kBuildRows;
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
  if (g_gemm_mode == 1)      // at most 16384 rows per int32 accumulation; more splits when that is needed to fill the SMs
    s.splits = std::max(best, (int)((s.slab_rows + kOzakiKRange - 1) / kOzakiKRange));
  return s;
}

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

static size_t gram_exps_len(const Plan* pl) { return (size_t)(pl->p_pad + 255) / 256 * 256; }

// workspace: [Phi^T slab (FP64, DMMA path) | digit planes (INT8 path)] [split partials] [exps] [col_hi]
size_t gram_workspace_bytes(const Plan* pl, int64_t n_pad, int sms) {
  const GramSchedule s = gram_schedule(pl->p_pad, n_pad, sms);
  const size_t slab = g_gemm_mode == 1 ? align256(ozaki_plane_bytes(pl->p_pad, (int)s.slab_rows))
                                       : align256((size_t)pl->p_pad * s.slab_rows * sizeof(double));
  return slab + align256((size_t)s.splits * pl->p_pad * pl->p_pad * sizeof(double)) + 2 * align256(gram_exps_len(pl) * sizeof(int));
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

int launch_gram(const Plan* pl, const double* T, int64_t n_pad, double* A, int64_t lda, void* workspace, size_t ws_bytes,
                int sms, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(n_pad % kBuildRows == 0, "gram: n_pad=%lld must be a multiple of %d", (long long)n_pad, kBuildRows);
  GRIEF_REQUIRE(ws_bytes >= gram_workspace_bytes(pl, n_pad, sms), "gram: workspace too small");
  const GramSchedule s = gram_schedule(pl->p_pad, n_pad, sms);
  const int pp = pl->p_pad;
  const bool i8 = g_gemm_mode == 1;
  char* wq = reinterpret_cast<char*>(workspace);
  double* PhiT = reinterpret_cast<double*>(wq);
  int8_t* planes = reinterpret_cast<int8_t*>(wq);
  wq += i8 ? align256(ozaki_plane_bytes(pp, (int)s.slab_rows)) : align256((size_t)pp * s.slab_rows * sizeof(double));
  double* part = reinterpret_cast<double*>(wq); wq += align256((size_t)s.splits * pp * pp * sizeof(double));
  int* exps = reinterpret_cast<int*>(wq); wq += align256(gram_exps_len(pl) * sizeof(int));
  int* col_hi = reinterpret_cast<int*>(wq);
  const size_t part_doubles = (size_t)s.splits * pp * pp;
  if (n_pad == 0) GRIEF_CUDA(cudaMemsetAsync(part, 0, part_doubles * sizeof(double), stream));
  if (i8) GRIEF_CUDA(cudaMemsetAsync(col_hi, 0, gram_exps_len(pl) * sizeof(int), stream));
  GemmOpts o;
  o.lower_only = true;
  o.splits = s.splits;
  o.c_split_stride = (int64_t)pp * pp;
  for (int64_t r0 = 0; r0 < n_pad; r0 += s.slab_rows) {
    const int64_t R = std::min(s.slab_rows, n_pad - r0);
    const double* Ts = T + (size_t)r0 * pl->stride;
    BuildArgs ba;
    ba.transposed = true;
    int rc;
    prof_begin(PROF_BUILD_T, stream);
    if (i8) {      // exponents of the column maxima over the slab, then the digit planes [7][p_pad][R] (K = data rows)
      ba.mode = OUT_EXP; ba.col_hi = col_hi;
      rc = launch_build(pl, Ts, R, ba, stream);
      if (rc == GRIEF_OK) {
        const int n_e = (int)gram_exps_len(pl);
        k_exps_from_hi<<<(n_e + 255) / 256, 256, 0, stream>>>(col_hi, pp, n_e, exps, ozaki_err_flag());
        ba.mode = OUT_DIGITS; ba.exps = exps; ba.planes = planes; ba.ld = R;
        rc = launch_build(pl, Ts, R, ba, stream);
      }
    } else {
      ba.mode = OUT_F64; ba.out = PhiT; ba.ld = s.slab_rows;
      rc = launch_build(pl, Ts, R, ba, stream);
    }
    prof_end(PROF_BUILD_T, stream);
    if (rc != GRIEF_OK) return rc;
    prof_begin(PROF_GRAM, stream);
    if (i8)
      rc = ozaki_gemm(planes, pp, exps, pp, planes, pp, exps, pp, (int)R, part, pp, r0 > 0, true, s.splits, (int64_t)pp * pp, stream, launches);
    else
      rc = gemm_nt_ex(PhiT, s.slab_rows, PhiT, s.slab_rows, part, pp, pp, pp, (int)R, 1.0, r0 > 0 ? 1.0 : 0.0, o, stream, launches);
    prof_end(PROF_GRAM, stream);
    if (rc != GRIEF_OK) return rc;
    if (launches) *launches += i8 ? 3 : 1;
  }
  k_gram_reduce<<<dim3(s.nb, s.nb), 256, 0, stream>>>(part, pl->d_perm, pp, s.splits, lda, A);
  GRIEF_CUDA(cudaGetLastError());
  if (launches) *launches += 1;
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
  if (g_gemm_mode != 1) return align256((size_t)slab_rows * pl->p_pad * sizeof(double));                       // Phi slab
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
  if (g_gemm_mode != 1) {
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
int launch_zgemm_prepare(const Plan* pl, const double* Bperm, int64_t slab_rows_max, void* scratch, cudaStream_t stream) {
  if (g_gemm_mode != 1) return GRIEF_OK;
  ZScratch z = carve_zscratch(pl, slab_rows_max, scratch);
  return ozaki_slice(Bperm, pl->p_pad, pl->p_pad, pl->p_pad, z.eb, (pl->p_pad + 255) / 256 * 256, z.pb, stream);
}

// Z (slab_rows x ldz, columns in SORTED order) = Phi(slab) * B, B symmetric given as Bperm (p_pad x p_pad, launch_permute_b).
// scratch: zgemm_scratch_bytes(pl, slab_rows_max) bytes, prepared by launch_zgemm_prepare.
int launch_zgemm(const Plan* pl, const double* T_slab, int64_t slab_rows, const double* Bperm, void* scratch, int64_t slab_rows_max, double* Z,
                 int64_t ldz, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(slab_rows % kBuildRows == 0, "zgemm: slab_rows=%lld is not a multiple of %d", (long long)slab_rows, kBuildRows);
  GRIEF_REQUIRE(ldz >= pl->p_pad, "zgemm: ldz=%lld must be >= p_pad=%d", (long long)ldz, pl->p_pad);
  if (slab_rows == 0) return GRIEF_OK;
  ZScratch z = carve_zscratch(pl, slab_rows_max, scratch);
  BuildArgs ba;
  if (g_gemm_mode == 1) {      // row exponents + digit planes [7][slab_rows][p_pad] straight from the tables
    ba.mode = OUT_DIGITS; ba.exps = z.ea; ba.planes = z.pa; ba.plane_stride = (size_t)slab_rows * pl->p_pad;
  } else {
    ba.mode = OUT_F64; ba.out = z.Phi; ba.ld = pl->p_pad;
  }
  prof_begin(PROF_BUILD, stream);
  int rc = launch_build(pl, T_slab, slab_rows, ba, stream);
  prof_end(PROF_BUILD, stream);
  if (rc != GRIEF_OK) return rc;
  prof_begin(PROF_ZGEMM, stream);
  if (g_gemm_mode == 1) {
    GRIEF_REQUIRE(pl->p_pad <= kOzakiKRange, "zgemm: p_pad=%d exceeds the INT8 path's K range of %d", pl->p_pad, kOzakiKRange);
    rc = ozaki_gemm(z.pa, slab_rows, z.ea, (int)slab_rows, z.pb, pl->p_pad, z.eb, pl->p_pad, pl->p_pad, Z, ldz, false, false, 1, 0, stream, launches);
  } else {
    GemmOpts o;
    rc = gemm_nt_ex(z.Phi, pl->p_pad, Bperm, pl->p_pad, Z, ldz, (int)slab_rows, pl->p_pad, pl->p_pad, 1.0, 0.0, o, stream, launches);
  }
  prof_end(PROF_ZGEMM, stream);
  if (rc == GRIEF_OK && launches) *launches += 1;
  return rc;
}

}  // namespace grief
