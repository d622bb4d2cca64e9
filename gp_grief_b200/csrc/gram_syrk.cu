// G2: fused Phi-tile builder + FP64 tensor-core SYRK,  A = Phi^T Phi  (lower block triangle).
//
// Reference being replaced: models/gp_grief_model.py:148-149 (`Phi = kern.cov(X)[0]`,
// `A = Phi.T.dot(Phi)`), where Phi (n x p) is materialised in host memory (328 GB at n=1e7,p=4096).
// Here Phi never exists in HBM: every CTA streams 16-row chunks of the group table (plan.h) into
// shared memory with bulk async copies (1-D TMA, mbarrier completion), builds the two 16 x 128
// Phi tiles its output tile needs with G gathers + (G-1) multiplies per element, and feeds them to
// m16n8k16 FP64 MMAs (DMMA) whose 128 x 128 accumulator tile stays in registers for the whole row
// range.
//
// Work item = (row split s, output tile (bi,bj), bi >= bj).  Items are ordered split-major so the
// ~148 CTAs running concurrently read the same table rows (L2 reuse).  Each item writes its
// partial tile to a workspace; k_gram_reduce sums the splits in a fixed order (deterministic) and
// mirrors the result into the full symmetric p x p matrix.
#include "plan.h"

namespace grief {

constexpr int kGramThreads = 512;                 // 16 warps, 4 x 4, warp tile 32 (M) x 32 (N); 4 warps per SM sub-partition
constexpr int kGramRows = 32;                     // data rows per pipeline chunk (two m16n8k16-equivalent K steps)
constexpr int kTStages = 3;                       // table-chunk ring (bulk async copies)
constexpr int kPhiLd = 34;                        // doubles per Phi-tile column in smem (32 rows + 2 pad: conflict-free LDS.128)
constexpr int kPhiStageDoubles = 2 * kTileN * kPhiLd;

template <int G> struct SlotPack;   // G u16 slot indices of one column, padded to a power of two
template <> struct SlotPack<1> { static constexpr int GP = 1; };
template <> struct SlotPack<2> { static constexpr int GP = 2; };
template <> struct SlotPack<3> { static constexpr int GP = 4; };
template <> struct SlotPack<4> { static constexpr int GP = 4; };
template <> struct SlotPack<5> { static constexpr int GP = 8; };
template <> struct SlotPack<6> { static constexpr int GP = 8; };
template <> struct SlotPack<7> { static constexpr int GP = 8; };
template <> struct SlotPack<8> { static constexpr int GP = 8; };

// Tile builder.  Lane = data row of the 32-row chunk, the warp walks a run of consecutive columns of the plan's
// SORTED order (smallest group = major key), so neighbouring columns share their leading slots: sLvl[c] is the
// first key position where column c differs from c-1 (0 at the start of a run).  The product P of the first G-1
// factors lives in a register and is rebuilt only when one of them changes (warp-uniform branch, rare); every
// element then costs one gather of the last factor and one DMUL.  C3: ~1.5 gathers / 1.3 DMULs per element
// instead of 4 / 3 -- the build phase is bound by shared-memory traffic (profiles/r01_gram_design_notes.md).
template <int G, int NB>
__device__ __forceinline__ void build_run(double& P, const double* __restrict__ trow, const uint16_t* __restrict__ sIdx,
                                          const uint8_t* __restrict__ sLvl, double* __restrict__ dst, int c0) {
  constexpr int GP = SlotPack<G>::GP;
  double last[NB];
  int lv[NB];
#pragma unroll
  for (int e = 0; e < NB; ++e) {                      // all last-factor gathers are issued up front
    lv[e] = sLvl[c0 + e];
    last[e] = trow[sIdx[(c0 + e) * GP + (G - 1)]];
  }
#pragma unroll
  for (int e = 0; e < NB; ++e) {
    if constexpr (G > 1) {
      if (lv[e] < G - 1) {                            // a leading factor changed: rebuild the prefix product
        double q = trow[sIdx[(c0 + e) * GP]];
#pragma unroll
        for (int g = 1; g < G - 1; ++g) q *= trow[sIdx[(c0 + e) * GP + g]];
        P = q;
      }
      last[e] *= P;
    }
  }
#pragma unroll
  for (int e = 0; e < NB; ++e) dst[(c0 + e) * kPhiLd] = last[e];
}

struct GramParams {
  const double* T;               // group table, n_pad rows x stride
  const uint16_t* sorted_slot;   // p_pad x G, sorted column order
  const uint8_t* sorted_level;   // p_pad
  const int2* tiles;             // n_tiles (bi, bj)
  double* ws;                    // n_items x 128 x 128 partial tiles
  int stride;
  int n_tiles;
  int n_items;
  int64_t n_chunks;              // total 32-row chunks
  int64_t chunks_per_split;
};

// Every warp does both jobs in two phases per 32-row chunk: build its share of the Phi tiles of chunk lc+1, then the
// DMMAs of chunk lc.  Measured alternatives (profiles/r01_gram_design_notes.md): dedicated builder warps and DMULs
// interleaved between DMMAs are both slower, because a DMUL entering the DMMA-busy FP64 pipe stalls for tens of cycles.
template <int G>
__global__ void __launch_bounds__(kGramThreads, 1) k_gram(const GramParams prm) {
  constexpr int GP = SlotPack<G>::GP;
  constexpr int NB = 8;                     // columns in flight per lane during the build phase
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* t_full = reinterpret_cast<uint64_t*>(smem_raw);                      // kTStages barriers (64 B reserved)
  uint16_t* sIdx = reinterpret_cast<uint16_t*>(smem_raw + 64);                   // 256 x GP
  uint8_t* sLvl = smem_raw + 64 + 256 * GP * sizeof(uint16_t);                   // 256
  double* sT = reinterpret_cast<double*>(smem_raw + 64 + 256 * GP * sizeof(uint16_t) + 256);
  const int stage_doubles = kGramRows * prm.stride;
  double* sPhi = sT + (size_t)kTStages * stage_doubles;                          // 2 x 256 x kPhiLd

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, t4 = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  // Warps 0-3 and 8-11 build first and multiply second, warps 4-7 and 12-15 the other way round: every SM
  // sub-partition (warp % 4) always has two warps with DMMAs in flight while the other two sit in the
  // latency-bound tile build.
  const bool build_first = ((warp >> 2) & 1) == 0;

  if (tid == 0) {
    for (int s = 0; s < kTStages; ++s) mbar_init(&t_full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const uint32_t stage_bytes = (uint32_t)stage_doubles * sizeof(double);
  uint64_t gq = 0;   // chunks consumed so far by this CTA (ring slot + mbarrier parity)

  for (int item = blockIdx.x; item < prm.n_items; item += gridDim.x) {
    const int split = item / prm.n_tiles;
    const int tile = item - split * prm.n_tiles;
    const int2 ij = prm.tiles[tile];
    const bool diag = (ij.x == ij.y);
    const int64_t c0 = (int64_t)split * prm.chunks_per_split;
    const int64_t c1 = min(prm.n_chunks, c0 + prm.chunks_per_split);
    const int nc = (int)max((int64_t)0, c1 - c0);
    const int ncols = diag ? kTileN : 2 * kTileN;
    const int boff = diag ? 0 : kTileN;
    const int cpw = ncols >> 4;                 // columns per warp per chunk: one run of sorted columns (16, or 8 on diagonal tiles)

    __syncthreads();
    for (int e = tid; e < ncols * G; e += kGramThreads) {
      const int c = e / G, g = e - c * G;
      const int col = (c < kTileN ? ij.x * kTileN + c : ij.y * kTileN + (c - kTileN));
      sIdx[c * GP + g] = prm.sorted_slot[(size_t)col * G + g];
    }
    for (int c = tid; c < ncols; c += kGramThreads) {
      const int col = (c < kTileN ? ij.x * kTileN + c : ij.y * kTileN + (c - kTileN));
      sLvl[c] = (c % cpw == 0) ? (uint8_t)0 : prm.sorted_level[col];
    }
    __syncthreads();

    double acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.0;

    auto issue = [&](int lc) {   // thread 0 only; the target ring slot was released by the last __syncthreads
      const uint64_t q = gq + lc;
      const int st = (int)(q % kTStages);
      fence_proxy_async();
      mbar_arrive_expect_tx(&t_full[st], stage_bytes);
      bulk_g2s(sT + (size_t)st * stage_doubles, prm.T + (size_t)(c0 + lc) * stage_doubles, stage_bytes, &t_full[st]);
    };
    auto build = [&](int lc) {   // chunk lc: table ring slot -> Phi tile buffer lc & 1
      const uint64_t q = gq + lc;
      const int st = (int)(q % kTStages);
      mbar_wait(&t_full[st], (uint32_t)((q / kTStages) & 1));
      const double* trow = sT + (size_t)st * stage_doubles + (size_t)lane * prm.stride;
      double* dst = sPhi + (size_t)(lc & 1) * kPhiStageDoubles + lane;
      double P = 1.0;
      for (int it = 0; it < cpw; it += NB) build_run<G, NB>(P, trow, sIdx, sLvl, dst, warp * cpw + it);
    };

    if (nc > 0) {
      if (tid == 0)
        for (int s = 0; s < kTStages - 1 && s < nc; ++s) issue(s);
      build(0);
      __syncthreads();
      for (int lc = 0; lc < nc; ++lc) {
        if (tid == 0 && lc + kTStages - 1 < nc) issue(lc + kTStages - 1);   // slot of chunk lc-1: free since last sync
        if (build_first && lc + 1 < nc) build(lc + 1);
        // ---- MMA phase (chunk lc).  K order inside each 16-row half is permuted (lane t supplies rows 4t..4t+3):
        //      one LDS.128 yields two K steps; K steps are outermost so consecutive DMMAs are independent ----
        const double* base = sPhi + (size_t)(lc & 1) * kPhiStageDoubles;
        const double* pa = base + (size_t)(wm * 32 + g4) * kPhiLd + 4 * t4;
        const double* pbp = base + (size_t)(boff + wn * 32 + g4) * kPhiLd + 4 * t4;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {                   // (16-row half, 8-row quarter pair) = offset 16*(q4>>1) + 2*(q4&1)
          const int off = 16 * (q4 >> 1) + 2 * (q4 & 1);
          double2 av[2][2], bv[4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) bv[nt] = *reinterpret_cast<const double2*>(pbp + nt * 8 * kPhiLd + off);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              av[mt][h] = *reinterpret_cast<const double2*>(pa + (mt * 16 + 8 * h) * kPhiLd + off);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma_16x8x4(acc[mt][nt], av[mt][0].x, av[mt][1].x, bv[nt].x);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma_16x8x4(acc[mt][nt], av[mt][0].y, av[mt][1].y, bv[nt].y);
          pipe_window();
        }
        if (!build_first && lc + 1 < nc) build(lc + 1);
        __syncthreads();
      }
      gq += nc;
    }

    // partial tile -> workspace (row-major 128 x 128: [m][n], m indexes block bi, n block bj)
    double* out = prm.ws + (size_t)item * (kTileN * kTileN);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int mrow = wm * 32 + mt * 16 + g4 + 8 * hh;
          const int ncol = wn * 32 + nt * 8 + 2 * t4;
          *reinterpret_cast<double2*>(out + (size_t)mrow * kTileN + ncol) =
              make_double2(acc[mt][nt][2 * hh], acc[mt][nt][2 * hh + 1]);
        }
  }
}

// A[perm[bi*128+m]][perm[bj*128+n]] = sum_s ws[s][tile][m][n] (fixed split order), mirrored across the diagonal.
// `perm` maps the sorted column order the tiles were computed in back to the caller's column order.
__global__ void __launch_bounds__(256)
k_gram_reduce(const double* __restrict__ ws, const int2* __restrict__ tiles, const int* __restrict__ perm, int n_tiles,
              int n_splits, int p, int64_t lda, double* __restrict__ A) {
  const int tile = blockIdx.x;
  const int2 ij = tiles[tile];
  for (int e = threadIdx.x; e < kTileN * kTileN; e += blockDim.x) {
    const int m = e / kTileN, n = e - m * kTileN;
    const int srow = ij.x * kTileN + m, scol = ij.y * kTileN + n;
    if (ij.x == ij.y && scol > srow) continue;       // diagonal tile: lower triangle only, mirrored below
    const int row = perm[srow], col = perm[scol];
    if (row < 0 || col < 0) continue;                // padding columns
    double s = 0.0;
    for (int sp = 0; sp < n_splits; ++sp) s += ws[((size_t)sp * n_tiles + tile) * (kTileN * kTileN) + e];
    A[(size_t)row * lda + col] = s;
    A[(size_t)col * lda + row] = s;
  }
}

struct GramSchedule {
  int nb, n_tiles, n_splits, n_items;
  int64_t n_chunks, chunks_per_split;
};

GramSchedule gram_schedule(int p_pad, int64_t n_pad, int sms) {
  GramSchedule s;
  s.nb = p_pad / kTileN;
  s.n_tiles = s.nb * (s.nb + 1) / 2;
  s.n_chunks = n_pad / kGramRows;
  // choose the number of row splits: fill whole waves of `sms` CTAs, keep >= 32 chunks (1024 rows) per split
  int64_t max_splits = std::max<int64_t>(1, std::min<int64_t>(s.n_chunks / 32, 4096 / std::max(1, s.n_tiles) + 1));
  int best = 1;
  double best_eff = -1.0;
  for (int64_t S = 1; S <= max_splits; ++S) {
    const int64_t items = S * s.n_tiles;
    const int64_t waves = (items + sms - 1) / sms;
    const double eff = (double)items / (double)(waves * sms);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = (int)S; }
    if (eff >= 0.985) { best = (int)S; break; }
  }
  s.n_splits = best;
  s.chunks_per_split = (s.n_chunks + best - 1) / best;
  if (s.chunks_per_split < 1) s.chunks_per_split = 1;
  s.n_items = s.n_splits * s.n_tiles;
  return s;
}

size_t gram_workspace_bytes(const Plan* pl, int64_t n_pad, int sms) {
  GramSchedule s = gram_schedule(pl->p_pad, n_pad, sms);
  return (size_t)s.n_items * kTileN * kTileN * sizeof(double) + (size_t)s.n_tiles * sizeof(int2) + 256;
}

template <int G>
static int launch_gram_g(const Plan* pl, const GramParams& prm, int grid, size_t smem, cudaStream_t stream) {
  GRIEF_CUDA(cudaFuncSetAttribute(k_gram<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gram<G><<<grid, kGramThreads, smem, stream>>>(prm);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

int launch_gram(const Plan* pl, const double* T, int64_t n_pad, double* A, int64_t lda, void* workspace,
                size_t ws_bytes, int sms, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(n_pad % kGramRows == 0, "gram: n_pad=%lld must be a multiple of %d", (long long)n_pad, kGramRows);
  GRIEF_REQUIRE(ws_bytes >= gram_workspace_bytes(pl, n_pad, sms), "gram: workspace too small");
  GramSchedule s = gram_schedule(pl->p_pad, n_pad, sms);
  // workspace layout: [tiles (int2 x n_tiles) padded to 256 B][partials]
  int2* d_tiles = reinterpret_cast<int2*>(workspace);
  size_t tiles_bytes = ((size_t)s.n_tiles * sizeof(int2) + 255) / 256 * 256;
  double* d_ws = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + tiles_bytes);
  std::vector<int2> tiles(s.n_tiles);
  int t = 0;
  for (int bi = 0; bi < s.nb; ++bi)
    for (int bj = 0; bj <= bi; ++bj) tiles[t++] = make_int2(bi, bj);
  GRIEF_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
  GRIEF_CUDA(cudaStreamSynchronize(stream));   // `tiles` is a host temporary

  GramParams prm;
  prm.T = T;
  prm.sorted_slot = pl->d_sorted_slot;
  prm.sorted_level = pl->d_sorted_level;
  prm.tiles = d_tiles;
  prm.ws = d_ws;
  prm.stride = pl->stride;
  prm.n_tiles = s.n_tiles;
  prm.n_items = s.n_items;
  prm.n_chunks = s.n_chunks;
  prm.chunks_per_split = s.chunks_per_split;
  const int G = pl->n_groups;
  const int GP = G <= 1 ? 1 : (G <= 2 ? 2 : (G <= 4 ? 4 : 8));
  const size_t smem = 64 + 256 * GP * sizeof(uint16_t) + 256 + (size_t)kTStages * kGramRows * pl->stride * sizeof(double) +
                      (size_t)2 * kPhiStageDoubles * sizeof(double);
  const int grid = std::min(sms, s.n_items);
  int rc = GRIEF_OK;
  if (s.n_chunks > 0) {
    prof_begin(PROF_GRAM, stream);
    switch (G) {
      case 1: rc = launch_gram_g<1>(pl, prm, grid, smem, stream); break;
      case 2: rc = launch_gram_g<2>(pl, prm, grid, smem, stream); break;
      case 3: rc = launch_gram_g<3>(pl, prm, grid, smem, stream); break;
      case 4: rc = launch_gram_g<4>(pl, prm, grid, smem, stream); break;
      case 5: rc = launch_gram_g<5>(pl, prm, grid, smem, stream); break;
      case 6: rc = launch_gram_g<6>(pl, prm, grid, smem, stream); break;
      case 7: rc = launch_gram_g<7>(pl, prm, grid, smem, stream); break;
      case 8: rc = launch_gram_g<8>(pl, prm, grid, smem, stream); break;
      default: return fail(GRIEF_ERR_UNSUPPORTED, "gram: %d groups", G);
    }
    prof_end(PROF_GRAM, stream);
    if (rc != GRIEF_OK) return rc;
  } else {
    GRIEF_CUDA(cudaMemsetAsync(d_ws, 0, (size_t)s.n_items * kTileN * kTileN * sizeof(double), stream));
  }
  k_gram_reduce<<<s.n_tiles, 256, 0, stream>>>(d_ws, d_tiles, pl->d_perm, s.n_tiles, s.n_splits, pl->p, lda, A);
  GRIEF_CUDA(cudaGetLastError());
  if (launches) *launches += 2;
  return GRIEF_OK;
}

}  // namespace grief
