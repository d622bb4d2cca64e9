// FP64 GEMM on the INT8 tensor cores (Ozaki splitting, tcgen05 kind::i8) -- the fast path of both O(n p^2) products.
//
//   C (M x N, f64) (+)= A (M x K, f64) * B (N x K, f64)^T
// Every operand row is scaled by a power of two so that |x| < 1 and cut into SD balanced 8-bit digits (SD = 3..7; 8 SD - 2 bits
// + sign, rounded to nearest):
//   x 2^-e = sum_s d_s 2^(-6 - 8 s),   A B^T = sum_{a,b} 2^(-12 - 8 (a + b)) D_a(A) D_b(B)^T,   pairs with a + b <= SD - 1 kept
//   (SD (SD + 1) / 2 products: 28 / 21 / 15 / 10; what is dropped is below ~2^(2 - 8 SD) of the row-scale product; for the
//   symmetric product A = X X^T with SD even the pair a = b = SD / 2 of the first dropped group is kept as well).  Each digit
//   product is an exact int8 x int8 -> int32 GEMM; one accumulation covers at most 16384 values of K (7 pairs x 2^14 x 2^14 < 2^31),
//   longer K is split over blockIdx.z.
// Kernel: one CTA per 128 x 128 tile (CL = 1) or one CTA PAIR per 256 x 128 tile (CL = 2, the default from 16 x 16 tiles on; see the
// template comment).  TMEM holds four 128-column int32 accumulators, one per significance group g = a + b;
// two sweeps over K (g = SD-1..SD-4 with all SD + SD digit planes, then g = SD-5..0 with digits 0..SD-5; the second sweep walks K
// backwards so that it starts on the chunks the first one left in L2), each followed by a drain; with SD <= 4 there is one sweep.
// All digit planes a sweep needs for one 32-byte K chunk sit in shared memory (3-D TMA boxes over [digit][row][32 B], SWIZZLE_32B;
// three 56 KB stages at SD = 7, more at fewer digits); warp 4 = TMA producer, warp 5 = MMA issuer (one thread,
// tcgen05.mma.cta_group::{1,2}.kind::i8, M = 128 or 256, N = 128, K = 32), warps 0-3 drain TMEM: tcgen05.ld -> f64 ->
// sum_g 2^(-12-8g) acc_g -> row / column scales -> transposed through shared memory (or stored transposed directly) -> added to C.
// The six tensor maps travel as __grid_constant__ kernel parameters (no descriptor is ever re-written in global memory).
// Launch order: full products are rasterised in groups of `group_n` column tiles so that the B digits of a group stay L2-resident
// while all row tiles (or row-tile pairs) pass; the lower-triangle pair tiles of the Gram follow a blocked order table (plan.cu).
// Measured (B200, C3, inside the benchmark under the 1000 W cap, profiles/r02_design_notes.md): Phi P^-1 (37888 x 4096 x 4096,
// 4 digits) 2.85 POP/s int8, Gram (6 digits) 2.4 POP/s, against 2.43 POP/s for the dense int8 GEMM of cuBLASLt in the same process;
// tensor pipe 67 % active (single CTAs: 55 %); with 7 digits max |C - C_dgemm| / max|C| ~ 2e-15 (round 1: 88 TFLOP/s FP64-equivalent
// alone, against 35 for cuBLAS DGEMM / 36.4 for the DMMA kernel in dense.cu).
// Variants measured and not adopted (profiles/r01_gram_design_notes.md): 128 x 256 tiles with two accumulators and five work
// items, CTA pairs with multicast B, 64-byte K stages for the light sweep, accumulator-interleaved MMA order (all slower or equal).
// Every barrier wait is bounded: a protocol failure raises an error flag (checked by the callers) instead of hanging the GPU.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "plan.h"

namespace grief {

constexpr int TM = 128, TN = 128, KC = 32;
constexpr int A_SLICE = TM * KC, B_SLICE = TN * KC;            // one digit plane of a tile and K chunk: 4 KB each
constexpr uint32_t SPIN_LIMIT = 1u << 26;
constexpr int kStageBudget = 196 * 1024;                       // bytes of shared memory for the operand ring

// Work items.  TMEM holds four 128-column int32 accumulators, one per significance group g = a + b (larger g = less significant).
// With g_top the last group kept (g_top = SD - 1, or SD when the diagonal pair is added, see OzOpts::diag_pair):
//   item 0: g = g_top .. g_top-3   all SD + SD digit planes   (SD = 7: g = 6..3, 22 MMAs per K chunk, 56 KB per stage)
//   item 1: g = g_top-4 .. 0       digits 0 .. g_top-4 of both (SD = 7: g = 2..0, 6 MMAs per K chunk, 24 KB per stage)
// Group g = SD (only with diag_pair, SD even) holds the single pair a = b = SD / 2.
// Two sweeps over K per tile, two drains.
__host__ __device__ constexpr int item_g_hi(int g_top, int it) { return it == 0 ? g_top : g_top - 4; }
__host__ __device__ constexpr int item_g_lo(int g_top, int it) { return it == 0 ? (g_top - 3 > 0 ? g_top - 3 : 0) : 0; }
__host__ __device__ constexpr int item_digits(int sd, int g_top, int it) { return it == 0 ? sd : g_top - 3; }

constexpr int oz_stage_bytes(int cl, int sd) { return sd * (A_SLICE + B_SLICE / cl); }
constexpr int oz_stages(int cl, int sd) { return kStageBudget / oz_stage_bytes(cl, sd) > 6 ? 6 : kStageBudget / oz_stage_bytes(cl, sd); }
constexpr size_t oz_smem_bytes(int cl, int sd) { return 1024 + 4096 + 20480 + 1024 + (size_t)oz_stages(cl, sd) * oz_stage_bytes(cl, sd); }

__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < SPIN_LIMIT; ++i) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// ---- row scales and slicing ----
// exps[r] = e with max_k |X[r][k]| < 2^e (0 for an all-zero row)
__global__ void k_row_exp(const double* __restrict__ X, int64_t ld, int K, int* __restrict__ exps, int* __restrict__ err) {
  const int r = blockIdx.x;
  double m = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const double a = fabs(X[(size_t)r * ld + k]);
    m = (a != a) ? 1.0 / 0.0 : fmax(m, a);            // NaN counts as non-finite
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]); __syncthreads(); }
  if (threadIdx.x == 0) {
    int e = 0;
    if (sm[0] > 0.0) frexp(sm[0], &e);
    if (!(sm[0] <= 1.7976931348623157e308)) { e = 0; atomicExch(err, 4); }     // Inf or NaN (fmax drops NaN: checked below too)
    exps[r] = e;
  }
}
// planes[s][r][k] = balanced digit s of rint(X[r][k] * 2^(8 sd - 2 - exps[r])):  x 2^-e = sum_s d_s 2^(-6 - 8 s)
__global__ void k_slice(const double* __restrict__ X, int64_t ld, int R, int K, int kp, int sd, const int* __restrict__ exps,
                        int8_t* __restrict__ planes) {
  const size_t plane = (size_t)R * kp;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / kp), k = (int)(e - (size_t)r * kp);
    long long v = (k < K) ? __double2ll_rn(ldexp(X[(size_t)r * ld + k], 8 * sd - 2 - exps[r])) : 0;    // |v| <= 2^(8 sd - 2), exact
    for (int s = sd - 1; s >= 0; --s) {
      const int d = (int)(int8_t)(v & 0xFF);
      planes[(size_t)s * plane + e] = (int8_t)d;
      v = (v - d) >> 8;
    }
  }
}

struct OzMaps { CUtensorMap m[6]; };   // A heavy, A light, B heavy, B light, B half rows heavy, B half rows light

struct OzParams {
  double* C; int64_t ldc;
  const int* ea; const int* eb;      // row exponents of A and B, padded with zeros to multiples of 128
  int chunks;                        // 32-byte K chunks in total
  int split_chunks;                  // chunks per blockIdx.z (== chunks when K is not split)
  int64_t c_split_stride;            // split z writes C + z * c_split_stride
  int m_valid, n_valid;              // elements of C that exist
  int lower_only;                    // skip tiles entirely above the diagonal
  int accumulate;                    // 1: C += result, 0: C = result
  int store_t;                       // 1: C is stored transposed, element (m, n) at C[n * ldc + m]
  int group_n;                       // column tiles per rasterisation group (0: plain row-major tile order)
  const uint32_t* tile_order;        // CL = 2, lower_only: (pair row << 16 | column tile) of cluster blockIdx.x / 2, or nullptr
  int* err;
};

// CL = CTAs per tile (1 or 2).  CL = 2: a 256 x 128 tile is computed by a PAIR of CTAs (thread-block cluster (2,1,1), adjacent in x)
// with tcgen05.mma.cta_group::2: each CTA stages its own 128 A rows and HALF of the B rows (64), the leader CTA issues the MMAs
// (M = 256, N = 128, K = 32), each CTA's TMEM receives its 128 rows of the four accumulators and each CTA drains its own rows.
// Both CTAs' TMA loads complete on the LEADER's full barrier (peer bit of the barrier address cleared), the leader's commits are
// multicast to both CTAs' empty / tfull barriers, the drain threads of both CTAs arrive on the leader's tfree barrier.
// SD = digits per operand, GT = last significance group kept (SD - 1, or SD with the diagonal pair).  Both are template
// parameters: the MMA sequence of a K chunk must be straight-line code -- one thread issues it, and with run-time group bounds
// that thread spent more cycles deciding than the tensor pipe spent multiplying (measured: 1650 instead of 2400 TOP/s).
template <int CL, int SD, int GT, int IT>
__device__ __forceinline__ void issue_chunk(uint32_t tmem, uint64_t da0, uint64_t db0, uint32_t idesc, bool first) {
  constexpr int nd = item_digits(SD, GT, IT), g_hi = item_g_hi(GT, IT), g_lo = item_g_lo(GT, IT);
  constexpr int kBSlice = B_SLICE / CL;
#pragma unroll
  for (int gi = 0; gi < 4; ++gi) {            // accumulator gi <-> group g_hi - gi, TMEM columns gi*128 ..
    const int g = g_hi - gi;
    if (g < g_lo) continue;
    bool fresh = first;
#pragma unroll
    for (int a = 0; a < SD; ++a) {
      const int b = g - a;
      if (a >= nd || b < 0 || b >= nd) continue;
      if (g == SD && a != b) continue;          // group SD: the diagonal pair only
      const uint64_t da = da0 + (uint64_t)(a * (A_SLICE >> 4)), db = db0 + (uint64_t)(b * (kBSlice >> 4));
      const uint32_t accf = fresh ? 0u : 1u;
      if (CL == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                         tmem + (uint32_t)(gi * TN)), "l"(da), "l"(db), "r"(idesc), "r"(accf) : "memory");
      else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                         tmem + (uint32_t)(gi * TN)), "l"(da), "l"(db), "r"(idesc), "r"(accf) : "memory");
      fresh = false;
    }
  }
}

template <int CL, int SD, int GT>
__global__ void __launch_bounds__(192, 1) k_ozaki(const __grid_constant__ OzMaps maps, const OzParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t full = base, empty = base + 64, tfull = base + 128, tfree = base + 136, slot = base + 144;
  const uint32_t cscale = base + 1024;               // 128 doubles: 2^eb of the tile's columns
  const uint32_t stagebuf = base + 4096;             // 4 warps x 32 rows x 17 doubles (transpose staging for coalesced stores)
  const uint32_t ring = base + 4096 + 20480;         // 1024-aligned
  constexpr int kBSlice = B_SLICE / CL;              // this CTA's share of a B digit plane: 128 or 64 rows of 32 bytes
  constexpr int kStageBytes = oz_stage_bytes(CL, SD);
  constexpr int kStages = oz_stages(CL, SD);
  constexpr int kNumItems = GT > 3 ? 2 : 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int bn = CL == 2 ? blockIdx.y : blockIdx.x, bm = CL == 2 ? blockIdx.x : blockIdx.y;   // CTA pairs are adjacent in x
  if (CL == 1 && prm.group_n > 0) {                  // grouped rasterisation: all row tiles pass over group_n column tiles at a time
    const int tiles_n = gridDim.x, tiles_m = gridDim.y;
    const int L = (int)blockIdx.y * tiles_n + (int)blockIdx.x;
    const int per_group = prm.group_n * tiles_m;
    const int grp = L / per_group;
    const int first = grp * prm.group_n;
    const int gw = min(prm.group_n, tiles_n - first);
    const int within = L - grp * per_group;
    bm = within / gw;
    bn = first + (within - bm * gw);
  }
  if (CL == 2 && prm.group_n > 0) {                  // the same for CTA pairs: a pair (adjacent in x) walks the column tiles of a group
    const int npairs = gridDim.x >> 1, tiles_n = gridDim.y;
    const int L = (int)blockIdx.y * npairs + ((int)blockIdx.x >> 1);      // cluster id in launch order
    const int per_group = prm.group_n * npairs;
    const int grp = L / per_group;
    const int first = grp * prm.group_n;
    const int gw = min(prm.group_n, tiles_n - first);
    const int within = L - grp * per_group;
    const int pm = within / gw;
    bn = first + (within - pm * gw);
    bm = 2 * pm + ((int)blockIdx.x & 1);
  }
  if (CL == 2 && prm.tile_order != nullptr) {        // blocked order of the lower-triangle pair tiles (plan.cu)
    const uint32_t t = __ldg(prm.tile_order + (blockIdx.x >> 1));
    bm = 2 * (int)(t >> 16) + ((int)blockIdx.x & 1);
    bn = (int)(t & 0xFFFFu);
  }
  if (prm.lower_only && bn > (CL == 2 ? (bm | 1) : bm)) return;      // uniform over the cluster
  uint32_t crank = 0;
  if (CL == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool leader = crank == 0;
  const int c_begin = (int)blockIdx.z * prm.split_chunks;
  const int nk = min(prm.split_chunks, prm.chunks - c_begin);
  double* const Cz = prm.C + (size_t)blockIdx.z * prm.c_split_stride;
  if (nk <= 0) {                                    // this split has no K range: its tile is zero
    if (!prm.accumulate)
      for (int e = tid; e < TM * TN; e += 192) {
        const int r = bm * TM + e / TN, c = bn * TN + e % TN;
        if (r < prm.m_valid && c < prm.n_valid) Cz[prm.store_t ? (size_t)c * prm.ldc + r : (size_t)r * prm.ldc + c] = 0.0;
      }
    return;
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8 * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty + 8 * s));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tfull));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tfree), "r"(128 * CL));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CL == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  for (int j = tid; j < TN; j += 192) {
    const double cs = ldexp(1.0, prm.eb[bn * TN + j]);
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(cscale + 8 * j), "d"(cs) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL == 2) {                                    // barriers of both CTAs exist before anything is signalled remotely
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  bool ok = true;
  if (tid == 128) {                                 // ---- TMA producer (warp 4) ----
    int q = 0;
    for (int it = 0; it < kNumItems && ok; ++it) {
      const int nd = item_digits(SD, GT, it);
      const uint32_t bytes = (uint32_t)(nd * (A_SLICE + kBSlice));
      const CUtensorMap* mA = &maps.m[it == 0 ? 0 : 1];
      const CUtensorMap* mB = &maps.m[CL == 2 ? (it == 0 ? 4 : 5) : (it == 0 ? 2 : 3)];
      for (int c = 0; c < nk && ok; ++c, ++q) {
        const int s = q % kStages;
        if (q >= kStages) ok = wait_bounded(empty + 8 * s, (uint32_t)(((q / kStages) - 1) & 1));
        if (!ok) break;
        const uint32_t dst = ring + s * kStageBytes, bar = full + 8 * s;
        const int kc = (c_begin + (it == 0 ? c : nk - 1 - c)) * KC;      // the second sweep walks K backwards (L2 reuse; integer sums commute)
        if (CL == 1) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                       "l"(mA), "r"(kc), "r"(bm * TM), "r"(0), "r"(bar) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                           dst + nd * A_SLICE), "l"(mB), "r"(kc), "r"(bn * TN), "r"(0), "r"(bar) : "memory");
        } else {                                    // both CTAs' bytes are counted on the leader's barrier
          const uint32_t lbar = bar & 0xFEFFFFFFu;
          if (leader) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * bytes) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                           dst), "l"(mA), "r"(kc), "r"(bm * TM), "r"(0), "r"(lbar) : "memory");
          asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                           dst + nd * A_SLICE), "l"(mB), "r"(kc), "r"(bn * TN + (int)crank * (TN / 2)), "r"(0), "r"(lbar) : "memory");
        }
      }
    }
    if (!ok) atomicExch(prm.err, 1);
  } else if (tid == 160 && leader) {                // ---- MMA issuer (warp 5; of the leader CTA when CL = 2) ----
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((CL * TM) >> 4) << 24);
    const uint64_t dhi = (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);   // K-major SWIZZLE_32B: LBO 1, SBO 256 B, v1
    int q = 0;
    for (int it = 0; it < kNumItems && ok; ++it) {
      const int nd = item_digits(SD, GT, it);
      if (it > 0) ok = wait_bounded(tfree, (uint32_t)((it - 1) & 1));     // accumulators drained
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = 0; c < nk && ok; ++c, ++q) {
        const int s = q % kStages;
        ok = wait_bounded(full + 8 * s, (uint32_t)((q / kStages) & 1));
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = ring + s * kStageBytes, sb = sa + nd * A_SLICE;
        const uint64_t da0 = dhi | (uint64_t)((sa >> 4) & 0x3FFF), db0 = dhi | (uint64_t)((sb >> 4) & 0x3FFF);
        if (it == 0) issue_chunk<CL, SD, GT, 0>(tmem, da0, db0, idesc, c == 0);
        else issue_chunk<CL, SD, GT, 1>(tmem, da0, db0, idesc, c == 0);
        if (CL == 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * s) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                           empty + 8 * s), "h"((uint16_t)3) : "memory");
      }
      if (CL == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tfull) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tfull),
                     "h"((uint16_t)3) : "memory");
    }
    if (!ok) atomicExch(prm.err, 2);
  }
  __syncwarp();
  // ---- drains (warps 0-3): thread = tile row (TMEM lane); FP64 tile accumulated in global memory through a transpose ----
  if (warp < 4) {
    const int row0 = bm * TM + warp * 32;
    const double rs = (row0 + lane < prm.m_valid) ? ldexp(1.0, prm.ea[row0 + lane]) : 0.0;
    const uint32_t stg = stagebuf + (uint32_t)warp * (32 * 17 * 8);
    bool live = true;
    for (int it = 0; it < kNumItems && live; ++it) {
      live = wait_bounded(tfull, (uint32_t)(it & 1));
      if (!live) { if (lane == 0) atomicExch(prm.err, 3); break; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int g_hi = item_g_hi(GT, it), ng = g_hi - item_g_lo(GT, it) + 1;
      double wgt[4];
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) wgt[gi] = gi < ng ? ldexp(1.0, -12 - 8 * (g_hi - gi)) * rs : 0.0;
      for (int c0 = 0; c0 < TN; c0 += 16) {
        double acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0;
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {            // least significant group first
          if (gi >= ng) break;
          uint32_t v[16];
          const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(gi * TN + c0);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                         "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                       : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fma(wgt[gi], (double)(int)v[j], acc[j]);
        }
        const bool keep = (it > 0) || prm.accumulate;
        if (prm.store_t) {                           // C^T: for a fixed column the 32 lanes (= 32 consecutive rows) are contiguous
          const int row = row0 + lane;
          if (row < prm.m_valid) {
            double old[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = bn * TN + c0 + j;
              old[j] = (keep && col < prm.n_valid) ? __ldcg(Cz + (size_t)col * prm.ldc + row) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = bn * TN + c0 + j;
              double cs;
              asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cs) : "r"(cscale + (uint32_t)(c0 + j) * 8));
              if (col < prm.n_valid) __stcg(Cz + (size_t)col * prm.ldc + row, old[j] + acc[j] * cs);
            }
          }
          continue;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j)                // own row, 16 columns -> staging [row][17]
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(stg + (uint32_t)(lane * 17 + j) * 8), "d"(acc[j]) : "memory");
        __syncwarp();
        // two rows per step, lanes 0-15 / 16-31 along the 16 columns: 128-byte segments; all loads before all stores
        const int j = lane & 15;
        double cs;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cs) : "r"(cscale + (uint32_t)(c0 + j) * 8));
        const int col = bn * TN + c0 + j, rbase = row0 + (lane >> 4);
        const bool col_ok = col < prm.n_valid;
        double* dst0 = Cz + (size_t)rbase * prm.ldc + col;
        double old[16], val[16];
#pragma unroll
        for (int h = 0; h < 16; ++h) old[h] = (keep && col_ok && rbase + 2 * h < prm.m_valid) ? __ldcg(dst0 + (size_t)(2 * h) * prm.ldc) : 0.0;
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const int r = 2 * h + (lane >> 4);
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(val[h]) : "r"(stg + (uint32_t)(r * 17 + j) * 8));
        }
#pragma unroll
        for (int h = 0; h < 16; ++h)
          if (col_ok && rbase + 2 * h < prm.m_valid) __stcg(dst0 + (size_t)(2 * h) * prm.ldc, old[h] + val[h] * cs);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (CL == 1) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tfree) : "memory");
      } else {                                      // the issuer lives in the leader CTA
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(tfree), "r"(0));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL == 2) {                                    // nobody leaves while the partner can still write or signal into this CTA
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (CL == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 encode_fn3() {
  static EncodeTiledFn3 fn = nullptr;            // a driver entry point: the same for every device and thread
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn3>(p);
  }
  return fn;
}

// planes: [digits][rows_alloc][kp] int8; rows beyond `rows` read as zeros (TMA out-of-bounds fill)
static int make_plane_map(CUtensorMap* m, const int8_t* planes, int rows, int64_t rows_alloc, int kp, int box_rows, int depth, int digits) {
  EncodeTiledFn3 enc = encode_fn3();
  if (!enc) return fail(GRIEF_ERR_LIBRARY, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[3] = {(cuuint64_t)kp, (cuuint64_t)rows, (cuuint64_t)digits};
  const cuuint64_t gstr[2] = {(cuuint64_t)kp, (cuuint64_t)kp * (cuuint64_t)rows_alloc};
  const cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)box_rows, (cuuint32_t)depth};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(GRIEF_ERR_CUDA, "cuTensorMapEncodeTiled (int8 planes) failed with code %d", (int)r);
  return GRIEF_OK;
}

// Planes are always laid out for kOzMaxDigits digits (buffer sizes do not depend on the digit count in use).
size_t ozaki_plane_bytes(int64_t rows, int K) { return (size_t)kOzMaxDigits * (size_t)rows * (size_t)((K + KC - 1) / KC * KC); }

// exps[r] (r < rows; entries up to exps_len are zeroed) and digit planes of X (rows x K, ld); planes: [digits][rows][kp], kp = K rounded up to 32
int ozaki_slice(const double* X, int64_t ld, int rows, int K, int* exps, int exps_len, int8_t* planes, int digits, int* err, cudaStream_t stream) {
  if (rows == 0) return GRIEF_OK;
  GRIEF_REQUIRE(digits >= kOzMinDigits && digits <= kOzMaxDigits, "ozaki_slice: %d digits outside [%d,%d]", digits, kOzMinDigits, kOzMaxDigits);
  const int kp = (K + KC - 1) / KC * KC;
  if (exps_len > rows) GRIEF_CUDA(cudaMemsetAsync(exps + rows, 0, (size_t)(exps_len - rows) * sizeof(int), stream));
  k_row_exp<<<rows, 256, 0, stream>>>(X, ld, K, exps, err);
  k_slice<<<148 * 8, 256, 0, stream>>>(X, ld, rows, K, kp, digits, exps, planes);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

template <int CL, int SD, int GT>
static int launch_ozaki(const OzMaps& maps, const OzParams& prm, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = oz_smem_bytes(CL, SD);
  static_assert(smem <= 227 * 1024, "k_ozaki: shared memory");
  // function attributes are per device: set on every launch (a host-side table lookup in the runtime)
  GRIEF_CUDA(cudaFuncSetAttribute(k_ozaki<CL, SD, GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CL == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    GRIEF_CUDA(cudaLaunchKernelEx(&cfg, k_ozaki<CL, SD, GT>, maps, prm));
  } else {
    k_ozaki<CL, SD, GT><<<grid, 192, smem, stream>>>(maps, prm);
  }
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// C (M x N, ldc) (+)= A B^T from digit planes.  pa: [digits][rows_a_alloc][kp] with M valid rows, ea: >= round_up(M,128) exponents;
// pb likewise with N rows and >= round_up(N,256) exponents.  K <= 16384 * splits.
int ozaki_gemm(const int8_t* pa, int64_t rows_a_alloc, const int* ea, int M, const int8_t* pb, int64_t rows_b_alloc, const int* eb, int N,
               int K, double* C, int64_t ldc, bool accumulate, bool lower_only, int splits, int64_t c_split_stride, const OzOpts& opt,
               cudaStream_t stream, int* launches) {
  if (M <= 0 || N <= 0) return GRIEF_OK;
  const int SD = opt.digits;
  GRIEF_REQUIRE(SD >= kOzMinDigits && SD <= kOzMaxDigits, "ozaki_gemm: %d digits outside [%d,%d]", SD, kOzMinDigits, kOzMaxDigits);
  GRIEF_REQUIRE(opt.err != nullptr, "ozaki_gemm: no error flag");
  const int kp = (K + KC - 1) / KC * KC;
  const int chunks = kp / KC;
  splits = std::max(1, splits);
  const int split_chunks = (chunks + splits - 1) / splits;
  GRIEF_REQUIRE(split_chunks * KC <= 16384, "ozaki_gemm: %d values of K per accumulation exceed the int32 budget of 16384", split_chunks * KC);
  alignas(64) OzMaps maps;
  memset(&maps, 0, sizeof(maps));
  // Symmetric product (A = Phi^T Phi: both operands are the same matrix) with an even digit count: the first dropped group g = SD
  // contains the pair a = b = SD / 2, whose product d_a(x) d_a(x') is a SQUARE on the diagonal (and positively correlated for
  // correlated columns): a bias of ~2^(2 - 8 SD) that does not average out over K, measured as 10x the random truncation error.
  // That one pair is kept (one more MMA per K chunk); for independent operands (Z = Phi B) it is zero-mean like the others.
  const int g_top = (opt.diag_pair && SD % 2 == 0) ? SD : SD - 1;
  const int light = g_top - 3;                       // digit planes of the second sweep
  {
    int rc = make_plane_map(&maps.m[0], pa, M, rows_a_alloc, kp, TM, SD, SD);
    if (rc == GRIEF_OK) rc = make_plane_map(&maps.m[2], pb, N, rows_b_alloc, kp, TN, SD, SD);
    if (rc == GRIEF_OK) rc = make_plane_map(&maps.m[4], pb, N, rows_b_alloc, kp, TN / 2, SD, SD);      // half of the B rows (CTA-pair path)
    if (light > 0) {
      if (rc == GRIEF_OK) rc = make_plane_map(&maps.m[1], pa, M, rows_a_alloc, kp, TM, light, SD);
      if (rc == GRIEF_OK) rc = make_plane_map(&maps.m[3], pb, N, rows_b_alloc, kp, TN, light, SD);
      if (rc == GRIEF_OK) rc = make_plane_map(&maps.m[5], pb, N, rows_b_alloc, kp, TN / 2, light, SD);
    }
    if (rc != GRIEF_OK) return rc;
  }
  OzParams prm;
  prm.C = C; prm.ldc = ldc; prm.ea = ea; prm.eb = eb;
  prm.chunks = chunks; prm.split_chunks = split_chunks; prm.c_split_stride = c_split_stride;
  prm.m_valid = M; prm.n_valid = N; prm.lower_only = lower_only ? 1 : 0; prm.accumulate = accumulate ? 1 : 0; prm.store_t = opt.store_t; prm.err = opt.err;
  const int tiles_m = (M + TM - 1) / TM, tiles_n = (N + TN - 1) / TN;
  // rasterisation: the B digit panels of one group of column tiles (128 rows x K of a CTA x SD bytes each) take <= ~60 MB of the 126 MB L2
  prm.group_n = 0;
  if (!lower_only && tiles_m > 1) {
    const int64_t panel = (int64_t)TN * std::min<int64_t>(kp, (int64_t)split_chunks * KC) * SD;
    const int gmax = (int)std::max<int64_t>(1, ((int64_t)60 << 20) / panel);
    const int ngroups = (tiles_n + gmax - 1) / gmax;
    prm.group_n = (tiles_n + ngroups - 1) / ngroups;
  }
  // CTA pairs adjacent in x (cluster (2,1,1)); an odd last pair runs one CTA on zero rows.  Pairs pay off on large grids (C3, C4, C5:
  // +5..16 %); on a few tiles (C2: p = 1024, 8 x 8) the coarser tiles cost more in wave quantisation than they save in L2 traffic.
  const bool pairs = tiles_m >= 2 && (opt.cluster == 2 || (opt.cluster == 1 && std::min(tiles_m, tiles_n) >= 16));
  if (!pairs && prm.group_n >= tiles_n) prm.group_n = 0;      // one group = plain order (single CTAs: x already walks the column tiles)
  const bool ordered = pairs && lower_only && opt.tile_order != nullptr && opt.n_tile_order > 0 && M == N;
  prm.tile_order = ordered ? opt.tile_order : nullptr;
  const dim3 grid = ordered ? dim3(2 * opt.n_tile_order, 1, splits)
                            : (pairs ? dim3((tiles_m + 1) / 2 * 2, tiles_n, splits) : dim3(tiles_n, tiles_m, splits));
  int rc;
#define GRIEF_OZ(SD_, GT_)                                                                                         \
  if (SD == SD_ && g_top == GT_)                                                                                   \
    rc = pairs ? launch_ozaki<2, SD_, GT_>(maps, prm, grid, stream) : launch_ozaki<1, SD_, GT_>(maps, prm, grid, stream)
  rc = fail(GRIEF_ERR_BAD_ARG, "ozaki_gemm: %d digits with last group %d", SD, g_top);
  GRIEF_OZ(3, 2); GRIEF_OZ(4, 3); GRIEF_OZ(4, 4); GRIEF_OZ(5, 4); GRIEF_OZ(6, 5); GRIEF_OZ(6, 6); GRIEF_OZ(7, 6);
#undef GRIEF_OZ
  if (rc != GRIEF_OK) return rc;
  if (launches) *launches += 1;
  return GRIEF_OK;
}

// Synchronises the stream and reports a pipeline failure inside any k_ozaki launch on this flag since the last check.
int ozaki_check(int* err, cudaStream_t stream) {
  if (!err) return GRIEF_OK;
  int flag = 0;
  GRIEF_CUDA(cudaMemcpyAsync(&flag, err, sizeof(int), cudaMemcpyDeviceToHost, stream));
  GRIEF_CUDA(cudaStreamSynchronize(stream));
  if (flag != 0) {
    cudaMemsetAsync(err, 0, sizeof(int), stream);
    if (flag == 4) return fail(GRIEF_ERR_BAD_ARG, "non-finite value (NaN / Inf) in the basis matrix Phi or in the p x p operand");
    return fail(GRIEF_ERR_CUDA, "k_ozaki: barrier wait timed out in role %d (1 = TMA producer, 2 = MMA issuer, 3 = drain)", flag);
  }
  return GRIEF_OK;
}

}  // namespace grief
