// G4/G5 building block: Z = Phi * B for a slab of data rows, Phi built on the fly, B symmetric (p x p).
//
//   pass 2 of the hyper-parameter gradient:  B = G2 = -(P^-1 + b b^T / sigma^2)   
//   predictive variance:                     B = P^-1   (models/gp_grief_model.py:122-124, diagonal only)
//
// One CTA owns 128 data rows: their group-table rows stay resident in shared memory (bulk async
// copy), and the CTA walks all 128-column blocks of Z.  Per 16-column K chunk it builds the
// 128 x 16 Phi operand (G gathers, tree product), receives the 128 x 16 slice of B through a 2-D
// TMA tensor map (SWIZZLE_128B, so the LDS.128 fragment reads are bank-conflict free) and issues 64
// m16n8k4 FP64 MMAs per warp.  B is symmetric, so the "col-major" B fragment b[k][n] is read as B[n][k].
#include <cuda.h>

#include "plan.h"

namespace grief {

constexpr int kZThreads = 512;              // 16 warps, 4 x 4, warp tile 32 (rows) x 32 (cols); two anti-phase groups
constexpr int kZRows = 128;                 // data rows per CTA
constexpr int kZPhiLd = 18;                 // doubles per Phi-operand row in smem (16 + 2 pad)
constexpr int kZPhiStage = kZRows * kZPhiLd;
constexpr int kZBStageBytes = kTileN * kChunk * 8;   // 16 KB: 128 rows of B x 16 doubles, swizzled
constexpr int kZE = 4;                      // Phi elements per thread per K chunk (128 rows x 16 columns / 512 threads)

struct ZgemmParams {
  const double* T;               // table rows of the slab (row 0 = first slab row), stride doubles per row
  const uint16_t* sorted_slot;   // p_pad x G: K columns are walked in the plan's sorted order
  const uint8_t* sorted_level;   // p_pad
  double* Z;                     // out: slab_rows x ldz
  int64_t ldz;
  int stride;
  int n_row_blocks;              // slab_rows / 128
  int n_col_blocks;              // p_pad / 128
  int n_k_chunks;                // p_pad / 16
  int nb_stages;                 // B ring depth (2..4)
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// B'[c][k] = B[c][perm[k]] (0 for padding columns): the K dimension of Z = Phi * B is walked in sorted column order.
__global__ void __launch_bounds__(256)
k_permute_cols(const double* __restrict__ B, int64_t ldb, const int* __restrict__ perm, int p, int p_pad, double* __restrict__ out) {
  const int64_t total = (int64_t)p * p_pad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e / p_pad), k = (int)(e - (int64_t)c * p_pad);
    const int src = perm[k];
    out[e] = src >= 0 ? B[(size_t)c * ldb + src] : 0.0;
  }
}

template <int G>
__global__ void __launch_bounds__(kZThreads, 1) k_zgemm(const __grid_constant__ CUtensorMap mapB, const ZgemmParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem);          // nb_stages barriers
  uint64_t* t_full = b_full + 8;                                 // 1 barrier
  unsigned char* sB = smem + 1024;                               // nb_stages x 16 KB (1024-B aligned: SWIZZLE_128B)
  double* sPhi = reinterpret_cast<double*>(sB + (size_t)prm.nb_stages * kZBStageBytes);   // 2 x 128 x 18
  double* sT = sPhi + 2 * kZPhiStage;                            // 128 x stride

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, t4 = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int brow = tid & (kZRows - 1), bq4 = tid >> 7;           // builder: thread <-> data row, 4 of the 16 chunk columns
  const bool build_first = ((warp >> 2) & 1) == 0;               // anti-phase groups, two warps of each per SM sub-partition

  if (tid == 0) {
    for (int s = 0; s < prm.nb_stages; ++s) mbar_init(&b_full[s], 1);
    mbar_init(t_full, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int F = prm.n_col_blocks * prm.n_k_chunks;   // (column block, K chunk) pairs per row block
  uint64_t bq = 0;                                   // B tiles consumed so far (ring slot + parity)
  uint32_t t_parity = 0;

  for (int rb = blockIdx.x; rb < prm.n_row_blocks; rb += gridDim.x) {
    // ---- table rows of this row block -> smem (16-row pieces keep each bulk copy small) ----
    __syncthreads();
    if (tid == 0) {
      const uint32_t piece = (uint32_t)(16 * prm.stride * sizeof(double));
      fence_proxy_async();
      mbar_arrive_expect_tx(t_full, piece * (kZRows / 16));
      for (int i = 0; i < kZRows / 16; ++i)
        bulk_g2s(sT + (size_t)i * 16 * prm.stride, prm.T + ((size_t)rb * kZRows + i * 16) * prm.stride, piece, t_full);
    }
    auto issue_b = [&](int f) {   // thread 0 only
      const int cb = f / prm.n_k_chunks, kc = f - cb * prm.n_k_chunks;
      const uint64_t q = bq + f;
      const int st = (int)(q % prm.nb_stages);
      fence_proxy_async();
      mbar_arrive_expect_tx(&b_full[st], kZBStageBytes);
      tma_load_2d(sB + (size_t)st * kZBStageBytes, &mapB, kc * kChunk, cb * kTileN, &b_full[st]);
    };
    if (tid == 0)
      for (int f = 0; f < prm.nb_stages && f < F; ++f) issue_b(f);
    mbar_wait(t_full, t_parity);
    t_parity ^= 1;

    const double* trow = sT + (size_t)brow * prm.stride;
    // Phi operand of K chunk f: thread = (row, 4 consecutive sorted columns).  The product of the first G-1 factors
    // is rebuilt only where a leading slot changes (warp-uniform: all lanes of a warp share the columns).
    auto build = [&](int f) {
      const int kc = f % prm.n_k_chunks;
      const int col0 = kc * kChunk + bq4 * kZE;
      const uint16_t* cs = prm.sorted_slot + (size_t)col0 * G;
      double v[kZE];
#pragma unroll
      for (int e = 0; e < kZE; ++e) v[e] = trow[__ldg(cs + e * G + (G - 1))];
      if constexpr (G > 1) {
        double P = 1.0;
#pragma unroll
        for (int e = 0; e < kZE; ++e) {
          const int lv = (e == 0) ? 0 : (int)__ldg(prm.sorted_level + col0 + e);
          if (lv < G - 1) {
            double q = trow[__ldg(cs + e * G)];
#pragma unroll
            for (int g = 1; g < G - 1; ++g) q *= trow[__ldg(cs + e * G + g)];
            P = q;
          }
          v[e] *= P;
        }
      }
      double* dst = sPhi + (size_t)(f & 1) * kZPhiStage + (size_t)brow * kZPhiLd + bq4 * kZE;
      *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
    };

    double acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.0;

    build(0);
    __syncthreads();
    for (int f = 0; f < F; ++f) {
      if (build_first && f + 1 < F) build(f + 1);
      const uint64_t q = bq + f;
      const int st = (int)(q % prm.nb_stages);
      mbar_wait(&b_full[st], (uint32_t)((q / prm.nb_stages) & 1));
      const double* pa = sPhi + (size_t)(f & 1) * kZPhiStage + (size_t)(wm * 32 + g4) * kZPhiLd + 4 * t4;
      const unsigned char* pb = sB + (size_t)st * kZBStageBytes + (size_t)(wn * 32 + g4) * 128;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        double2 av[2][2], bv[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)     // row n = wn*32 + nt*8 + g4 (n & 7 == g4); 16-B unit (2*t4 + hf) ^ (n & 7)
          bv[nt] = *reinterpret_cast<const double2*>(pb + nt * 8 * 128 + (((2 * t4 + hf) ^ g4) << 4));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            av[mt][h] = *reinterpret_cast<const double2*>(pa + (mt * 16 + 8 * h) * kZPhiLd + 2 * hf);
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
      const int cb = f / prm.n_k_chunks, kc = f - cb * prm.n_k_chunks;
      if (kc == prm.n_k_chunks - 1) {       // column block finished: write the Z tile, restart the accumulators
        double* out = prm.Z + ((size_t)rb * kZRows) * prm.ldz + (size_t)cb * kTileN;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int mrow = wm * 32 + mt * 16 + g4 + 8 * hh;
              const int ncol = wn * 32 + nt * 8 + 2 * t4;
              *reinterpret_cast<double2*>(out + (size_t)mrow * prm.ldz + ncol) =
                  make_double2(acc[mt][nt][2 * hh], acc[mt][nt][2 * hh + 1]);
              acc[mt][nt][2 * hh] = 0.0;
              acc[mt][nt][2 * hh + 1] = 0.0;
            }
      }
      if (!build_first && f + 1 < F) build(f + 1);
      __syncthreads();
      if (tid == 0 && f + prm.nb_stages < F) issue_b(f + prm.nb_stages);
    }
    bq += F;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <int G>
static int launch_zgemm_g(const CUtensorMap& map, const ZgemmParams& prm, int grid, size_t smem, cudaStream_t stream) {
  GRIEF_CUDA(cudaFuncSetAttribute(k_zgemm<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_zgemm<G><<<grid, kZThreads, smem, stream>>>(map, prm);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

// Z (slab_rows x ldz) = Phi(slab rows) * B.   B: (p x p) symmetric, leading dimension ldb (even).
// slab_rows must be a multiple of 128 and T must hold that many (zero padded) rows.
// Bperm (p x p_pad doubles) must hold permute_b(B): call launch_permute_b once per B, then launch_zgemm per slab.
int launch_permute_b(const Plan* pl, const double* B, int64_t ldb, double* Bperm, cudaStream_t stream) {
  GRIEF_REQUIRE(ldb >= pl->p, "permute_b: ldb=%lld < p", (long long)ldb);
  const unsigned blocks = (unsigned)std::min<int64_t>(((int64_t)pl->p * pl->p_pad + 255) / 256, 148 * 8);
  k_permute_cols<<<blocks, 256, 0, stream>>>(B, ldb, pl->d_perm, pl->p, pl->p_pad, Bperm);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

int launch_zgemm(const Plan* pl, const double* T_slab, int64_t slab_rows, const double* Bperm, double* Z,
                 int64_t ldz, int sms, cudaStream_t stream, int* launches) {
  const double* B = Bperm;
  const int64_t ldb = pl->p_pad;
  GRIEF_REQUIRE(slab_rows % kZRows == 0, "zgemm: slab_rows=%lld is not a multiple of %d", (long long)slab_rows, kZRows);
  GRIEF_REQUIRE(ldb % 2 == 0 && ldb >= pl->p, "zgemm: ldb=%lld must be even and >= p", (long long)ldb);
  GRIEF_REQUIRE(ldz >= pl->p_pad && ldz % 2 == 0, "zgemm: ldz=%lld must be even and >= p_pad=%d", (long long)ldz, pl->p_pad);
  GRIEF_REQUIRE((reinterpret_cast<uintptr_t>(B) & 15) == 0, "zgemm: B must be 16-byte aligned");
  if (slab_rows == 0) return GRIEF_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(GRIEF_ERR_CUDA, "zgemm: cuTensorMapEncodeTiled is not available from the driver");
  alignas(64) CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)pl->p_pad, (cuuint64_t)pl->p};    // inner = sorted K columns, outer = rows (n)
  const cuuint64_t gstr[1] = {(cuuint64_t)ldb * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)kChunk, (cuuint32_t)kTileN};       // 16 doubles (128 B) x 128 rows
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(B), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(GRIEF_ERR_CUDA, "zgemm: cuTensorMapEncodeTiled failed with code %d", (int)cr);

  ZgemmParams prm;
  prm.T = T_slab;
  prm.sorted_slot = pl->d_sorted_slot;
  prm.sorted_level = pl->d_sorted_level;
  prm.Z = Z;
  prm.ldz = ldz;
  prm.stride = pl->stride;
  prm.n_row_blocks = (int)(slab_rows / kZRows);
  prm.n_col_blocks = pl->p_pad / kTileN;
  prm.n_k_chunks = pl->p_pad / kChunk;
  const size_t fixed = 1024 + 1024 + (size_t)2 * kZPhiStage * sizeof(double) + (size_t)kZRows * pl->stride * sizeof(double);
  const size_t max_smem = 227 * 1024;
  int nb = (int)((max_smem - fixed) / kZBStageBytes);
  if (nb > 4) nb = 4;
  if (nb < 2) return fail(GRIEF_ERR_UNSUPPORTED, "zgemm: table stride %d leaves no room for the B ring", pl->stride);
  prm.nb_stages = nb;
  const size_t smem = fixed + (size_t)nb * kZBStageBytes;
  const int grid = std::min(sms, prm.n_row_blocks);
  int rc;
  prof_begin(PROF_ZGEMM, stream);
  switch (pl->n_groups) {
    case 1: rc = launch_zgemm_g<1>(map, prm, grid, smem, stream); break;
    case 2: rc = launch_zgemm_g<2>(map, prm, grid, smem, stream); break;
    case 3: rc = launch_zgemm_g<3>(map, prm, grid, smem, stream); break;
    case 4: rc = launch_zgemm_g<4>(map, prm, grid, smem, stream); break;
    case 5: rc = launch_zgemm_g<5>(map, prm, grid, smem, stream); break;
    case 6: rc = launch_zgemm_g<6>(map, prm, grid, smem, stream); break;
    case 7: rc = launch_zgemm_g<7>(map, prm, grid, smem, stream); break;
    case 8: rc = launch_zgemm_g<8>(map, prm, grid, smem, stream); break;
    default: return fail(GRIEF_ERR_UNSUPPORTED, "zgemm: %d groups", pl->n_groups);
  }
  prof_end(PROF_ZGEMM, stream);
  if (rc == GRIEF_OK && launches) *launches += 1;
  return rc;
}

}  // namespace grief
