// Prototype 2: as ozaki_gemm.cu, but every 256 x 256 tile is computed by a PAIR of CTAs (thread-block cluster of 2 along M) with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 A rows and HALF of the B rows (128), the leader CTA issues the MMAs
// (M = 256, N = 256, K = 32), each CTA's TMEM receives its 128 rows of the two accumulators.  Per SM and MMA the tensor core then
// reads 4 KB of A + 4 KB of B from shared memory in 128 cycles = 64 B/clk instead of 128 B/clk.
// Prototype: FP64 GEMM emulated on the INT8 tensor cores (Ozaki splitting), sm_100a tcgen05.  Standalone test bed.
//   C[M][N] (f64) = A[M][K] (f64, K contiguous) * B[N][K]^T (f64, K contiguous)
// Each operand row is scaled by a power of two so that |x| < 1 and cut into S = 8 signed 7-bit slices (56 bits):
//   x = sum_k s_k 2^(-7(k+1)),   A B^T = sum_{a,b} 2^(-7(a+b+2)) S_a(A) S_b(B)^T,   pairs with a + b <= 7 kept (36 of 64).
// Every slice product is an exact int8 x int8 -> int32 GEMM (K <= 16384 per accumulation keeps 8 pairs within int32).
// Kernel: one CTA per 128 x 256 tile; four passes over K, each pass accumulates the pairs of TWO significance groups
// g = a + b into two TMEM accumulators (2 x 256 columns), all needed slices of a 32-byte K chunk resident in shared
// memory (TMA 3-D boxes, SWIZZLE_32B, 2-stage ring); after a pass the four warps drain TMEM into the FP64 tile.
// All waits are bounded (error flag instead of a hang).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ozaki tools/microbench/ozaki_gemm.cu -lcuda -lcublas
#include <cublas_v2.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int S = 7;                     // balanced 8-bit digits per operand (54 bits + sign)
constexpr int TM = 128, TN = 256, KC = 32, STAGES = 4;   // per CTA: 128 rows of A, and TN/2 = 128 rows of B
constexpr int A_SLICE = TM * KC, B_SLICE = (TN / 2) * KC;      // 4 KB each (B: this CTA's half)
constexpr int STAGE_BYTES = 40 * 1024;                         // largest work item: 5 A slices + 5 B half-slices
constexpr uint32_t SPIN_LIMIT = 1u << 24;

// Work items: a significance-group pair (gh, gh-1) accumulates into the two TMEM accumulators; the pair (6,5) is cut in two
// by A-slice range so that every stage fits 60 KB and three stages fit shared memory.
struct Item { int gh, a_lo, a_hi, b_lo, b_hi, first, last; };
__constant__ Item kItems[5] = {
    {6, 0, 3, 2, 6, 1, 0},   // g=6: a=0..3 (b=6..3); g=5: a=0..3 (b=5..2)            8 MMAs, 56 KB
    {6, 4, 6, 0, 2, 0, 1},   // g=6: a=4..6 (b=2..0); g=5: a=4..5 (b=1..0)            5 MMAs, 36 KB
    {4, 0, 4, 0, 4, 1, 1},   // g=4: 5 pairs; g=3: 4 pairs                            9 MMAs, 60 KB
    {2, 0, 2, 0, 2, 1, 1},   // g=2: 3 pairs; g=1: 2 pairs                            5 MMAs, 36 KB
    {0, 0, 0, 0, 0, 1, 1},   // g=0: 1 pair                                           1 MMA,  12 KB
};
constexpr int kNumItems = 5;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < SPIN_LIMIT; ++i) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
// K-major, SWIZZLE_32B: 8-row groups of 32-byte rows (SBO = 256 B), LBO = 1, descriptor version 1, layout type 6
__device__ __forceinline__ uint64_t make_desc32(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61);
}

// ---- row scales and slicing ----
// exps[r] = e with max_k |X[r][k]| < 2^e (0 for an all-zero row)
__global__ void k_row_exp(const double* __restrict__ X, int64_t ld, int K, int* __restrict__ exps) {
  const int r = blockIdx.x;
  double m = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) m = fmax(m, fabs(X[(size_t)r * ld + k]));
  __shared__ double sm[256];
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]); __syncthreads(); }
  if (threadIdx.x == 0) { int e = 0; if (sm[0] > 0.0) frexp(sm[0], &e); exps[r] = e; }
}
// planes[s][r][k] = balanced digit s of trunc(X[r][k] * 2^(54 - exps[r])):  x 2^-e = sum_s d_s 2^(-6 - 8 s)
__global__ void k_slice(const double* __restrict__ X, int64_t ld, int R, int K, const int* __restrict__ exps, int8_t* __restrict__ planes) {
  const size_t plane = (size_t)R * K;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / K), k = (int)(e - (size_t)r * K);
    long long v = __double2ll_rz(ldexp(X[(size_t)r * ld + k], 54 - exps[r]));    // |v| < 2^54, exact
#pragma unroll
    for (int s = S - 1; s >= 0; --s) {
      const int d = (int)(int8_t)(v & 0xFF);
      planes[(size_t)s * plane + e] = (int8_t)d;
      v = (v - d) >> 8;
    }
  }
}

struct OzParams {
  double* C; int64_t ldc;
  const int* ea; const int* eb;      // row exponents of A and B
  const CUtensorMap* maps;           // [0..5]: A with box depth 0..5 slices, [6..11]: B likewise (unused entries zero)
  int K;
  int* err;
};

__global__ void __launch_bounds__(192, 1) k_ozaki(const OzParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t full = base, empty = base + 32, tfull = base + 64, tfree = base + 72, slot = base + 80;
  const uint32_t cscale = base + 1024;               // 256 doubles: 2^eb of the tile's columns
  const uint32_t stagebuf = base + 4096;             // 4 warps x 32 rows x 17 doubles (transpose staging for coalesced stores)
  const uint32_t ring = base + 4096 + 20480;         // 1024-aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bm = blockIdx.x, bn = blockIdx.y;        // bm: 128-row block; the cluster pair is (bm even, bm odd), adjacent in x
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const bool leader = crank == 0;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8 * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty + 8 * s));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tfull));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(tfree));      // drain threads of BOTH CTAs arrive on the leader's
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int j = tid; j < TN; j += 192) {
    const double cs = ldexp(1.0, prm.eb[bn * TN + j]);
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(cscale + 8 * j), "d"(cs) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const int nk = prm.K / KC;
  bool ok = true;
  if (tid == 128) {                                 // ---- TMA producer (warp 4) ----
    int q = 0;
    for (int it = 0; it < kNumItems && ok; ++it) {
      const Item w = kItems[it];
      const int na = w.a_hi - w.a_lo + 1, nb = w.b_hi - w.b_lo + 1;
      const uint32_t bytes = (uint32_t)(na * A_SLICE + nb * B_SLICE);
      const CUtensorMap* mA = prm.maps + na;
      const CUtensorMap* mB = prm.maps + 6 + nb;
      for (int c = 0; c < nk && ok; ++c, ++q) {
        const int s = q % STAGES;
        if (q >= STAGES) ok = wait_bounded(empty + 8 * s, (uint32_t)(((q / STAGES) - 1) & 1));
        if (!ok) break;
        const uint32_t dst = ring + s * STAGE_BYTES, bar = (full + 8 * s) & 0xFEFFFFFFu;   // peer bit cleared: the leader's barrier
        if (leader) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * s), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                     "l"(mA), "r"(c * KC), "r"(bm * TM), "r"(w.a_lo), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         dst + na * A_SLICE), "l"(mB), "r"(c * KC), "r"(bn * TN + (int)crank * (TN / 2)), "r"(w.b_lo), "r"(bar) : "memory");
      }
    }
    if (!ok) atomicExch(prm.err, 1);
  } else if (tid == 160 && leader) {                // ---- MMA issuer (warp 5 of the leader CTA) ----
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);   // M = 256 over the pair
    int q = 0, drains = 0;
    for (int it = 0; it < kNumItems && ok; ++it) {
      const Item w = kItems[it];
      const int na = w.a_hi - w.a_lo + 1;
      if (w.first && drains > 0) ok = wait_bounded(tfree, (uint32_t)((drains - 1) & 1));   // accumulators drained
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = 0; c < nk && ok; ++c, ++q) {
        const int s = q % STAGES;
        ok = wait_bounded(full + 8 * s, (uint32_t)((q / STAGES) & 1));
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = ring + s * STAGE_BYTES, sb = sa + na * A_SLICE;
        for (int half = 0; half < 2; ++half) {      // group gh -> columns 0..255, group gh-1 -> columns 256..511
          const int g = w.gh - half;
          if (g < 0) break;
          bool fresh = (c == 0) && w.first;         // first MMA of this accumulator in this K range overwrites
          for (int a = w.a_lo; a <= w.a_hi; ++a) {
            const int b = g - a;
            if (b < w.b_lo || b > w.b_hi) continue;
            const uint64_t da = make_desc32(sa + (a - w.a_lo) * A_SLICE), db = make_desc32(sb + (b - w.b_lo) * B_SLICE);
            const uint32_t accf = fresh ? 0u : 1u;
            fresh = false;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                             tmem + (uint32_t)(half * TN)), "l"(da), "l"(db), "r"(idesc), "r"(accf) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(empty + 8 * s),
                     "h"((uint16_t)3) : "memory");
      }
      if (w.last) {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tfull),
                     "h"((uint16_t)3) : "memory");
        ++drains;
      }
    }
    if (!ok) atomicExch(prm.err, 2);
  }
  __syncwarp();
  // ---- drains (warps 0-3): thread = tile row (TMEM lane); FP64 tile accumulated in global memory through a transpose ----
  if (warp < 4) {
    const int row0 = bm * TM + warp * 32;
    const double rs = ldexp(1.0, prm.ea[row0 + lane]);
    const uint32_t stg = stagebuf + (uint32_t)warp * (32 * 17 * 8);
    int drains = 0;
    bool live = true;
    for (int it = 0; it < kNumItems && live; ++it) {
      const Item w = kItems[it];
      if (!w.last) continue;
      live = wait_bounded(tfull, (uint32_t)(drains & 1));
      if (!live) { if (lane == 0) atomicExch(prm.err, 3); break; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const double w_hi = ldexp(1.0, -12 - 8 * w.gh) * rs, w_lo = (w.gh > 0) ? ldexp(1.0, -12 - 8 * (w.gh - 1)) * rs : 0.0;
      for (int c0 = 0; c0 < TN; c0 += 16) {
        uint32_t hi[16], lo[16];
        const uint32_t t_hi = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, t_lo = t_hi + TN;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]), "=r"(hi[4]), "=r"(hi[5]), "=r"(hi[6]), "=r"(hi[7]), "=r"(hi[8]),
                       "=r"(hi[9]), "=r"(hi[10]), "=r"(hi[11]), "=r"(hi[12]), "=r"(hi[13]), "=r"(hi[14]), "=r"(hi[15])
                     : "r"(t_hi));
        if (w.gh > 0) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                       : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]), "=r"(lo[4]), "=r"(lo[5]), "=r"(lo[6]), "=r"(lo[7]), "=r"(lo[8]),
                         "=r"(lo[9]), "=r"(lo[10]), "=r"(lo[11]), "=r"(lo[12]), "=r"(lo[13]), "=r"(lo[14]), "=r"(lo[15])
                       : "r"(t_lo));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) lo[j] = 0;
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j) {              // own row, 16 columns -> staging [row][17]
          const double v = w_hi * (double)(int)hi[j] + w_lo * (double)(int)lo[j];
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(stg + (uint32_t)(lane * 17 + j) * 8), "d"(v) : "memory");
        }
        __syncwarp();
        // two rows per step, lanes 0-15 / 16-31 along the 16 columns: 128-byte segments; all loads before all stores
        const int j = lane & 15;
        double cs;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cs) : "r"(cscale + (uint32_t)(c0 + j) * 8));
        double* dst0 = prm.C + (size_t)(row0 + (lane >> 4)) * prm.ldc + (size_t)bn * TN + c0 + j;
        double old[16], val[16];
#pragma unroll
        for (int h = 0; h < 16; ++h) old[h] = (drains == 0) ? 0.0 : __ldcg(dst0 + (size_t)(2 * h) * prm.ldc);
#pragma unroll
        for (int h = 0; h < 16; ++h) {
          const int r = 2 * h + (lane >> 4);
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(val[h]) : "r"(stg + (uint32_t)(r * 17 + j) * 8));
        }
#pragma unroll
        for (int h = 0; h < 16; ++h) __stcg(dst0 + (size_t)(2 * h) * prm.ldc, old[h] + val[h] * cs);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(tfree), "r"(0));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
      }
      ++drains;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void make_map3(CUtensorMap* m, const int8_t* p, int rows, int K, int box_rows, int depth) {
  static EncodeFn enc = nullptr;
  if (!enc) {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    enc = (EncodeFn)f;
  }
  const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)S};
  const cuuint64_t gstr[2] = {(cuuint64_t)K, (cuuint64_t)K * rows};
  const cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)box_rows, (cuuint32_t)depth};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)p, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

static int run(int M, int N, int K, bool check, cublasHandle_t h) {
  std::vector<double> hA((size_t)M * K), hB((size_t)N * K);
  uint64_t st = 0x9E3779B97F4A7C15ull ^ (uint64_t)(M * 31 + N * 7 + K);
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
  for (auto& x : hA) x = (rnd() - 0.5) * exp(4.0 * (rnd() - 0.5));     // magnitudes spread over ~e^4
  for (auto& x : hB) x = (rnd() - 0.5) * exp(4.0 * (rnd() - 0.5));
  double *dA, *dB, *dC, *dR; int8_t *pA, *pB; int *eA, *eB, *dErr;
  CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8)); CK(cudaMalloc(&dC, (size_t)M * N * 8)); CK(cudaMalloc(&dR, (size_t)M * N * 8));
  CK(cudaMalloc(&pA, (size_t)S * M * K)); CK(cudaMalloc(&pB, (size_t)S * N * K)); CK(cudaMalloc(&eA, M * 4)); CK(cudaMalloc(&eB, N * 4)); CK(cudaMalloc(&dErr, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dErr, 0, 4));
  cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  alignas(64) CUtensorMap hmaps[12];
  memset(hmaps, 0, sizeof(hmaps));
  for (int dpt = 1; dpt <= 5; ++dpt) { make_map3(&hmaps[dpt], pA, M, K, TM, dpt); make_map3(&hmaps[6 + dpt], pB, N, K, TN / 2, dpt); }
  CUtensorMap* dmaps; CK(cudaMalloc(&dmaps, sizeof(hmaps))); CK(cudaMemcpy(dmaps, hmaps, sizeof(hmaps), cudaMemcpyHostToDevice));
  const size_t smem = 1024 + 4096 + 20480 + 1024 + (size_t)STAGES * STAGE_BYTES;
  CK(cudaFuncSetAttribute(k_ozaki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  OzParams prm{dC, N, eA, eB, dmaps, K, dErr};
  dim3 grid(M / TM, N / TN);
  float ms_slice = 0, ms_gemm = 0;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    k_row_exp<<<M, 256>>>(dA, K, K, eA); k_row_exp<<<N, 256>>>(dB, K, K, eB);
    k_slice<<<148 * 8, 256>>>(dA, K, M, K, eA, pA); k_slice<<<148 * 8, 256>>>(dB, K, N, K, eB, pB);
    CK(cudaEventRecord(e1));
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = grid; cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (rep == 0) {
        int ncl = -1; cudaError_t oe = cudaOccupancyMaxActiveClusters(&ncl, k_ozaki, &cfg);
        cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_ozaki);
        printf("  grid (%u,%u,%u) smem %zu regs %d maxActiveClusters %d (%s)\n", grid.x, grid.y, grid.z, smem, fa.numRegs, ncl, cudaGetErrorString(oe));
      }
      CK(cudaLaunchKernelEx(&cfg, k_ozaki, prm));
    }
    CK(cudaEventRecord(e2)); CK(cudaEventSynchronize(e2));
    CK(cudaEventElapsedTime(&ms_slice, e0, e1)); CK(cudaEventElapsedTime(&ms_gemm, e1, e2));
  }
  int herr = 0; CK(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
  // reference: cuBLAS DGEMM, column-major view: C^T (N x M) = B (N x K as col-major K x N transposed) ...
  const double one = 1.0, zero = 0.0;
  cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one, dB, K, dA, K, &zero, dR, N);   // dR row-major M x N
  CK(cudaEventRecord(e0));
  cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one, dB, K, dA, K, &zero, dR, N);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms_ref; CK(cudaEventElapsedTime(&ms_ref, e0, e1));
  double max_err = 0, max_ref = 0;
  if (check) {
    std::vector<double> hC((size_t)M * N), hR((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hR.data(), dR, hR.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < hC.size(); ++i) { max_err = fmax(max_err, fabs(hC[i] - hR[i])); max_ref = fmax(max_ref, fabs(hR[i])); }
  }
  printf("M=%d N=%d K=%d: slice %.3f ms, ozaki gemm %.3f ms = %.1f TFLOP/s fp64-equivalent (%.0f TOPS int8); cuBLAS DGEMM %.3f ms = %.1f TFLOP/s; err_flag=%d",
         M, N, K, ms_slice, ms_gemm, 2.0 * M * N * K / ms_gemm * 1e-9, 28 * 2.0 * M * N * K / ms_gemm * 1e-9, ms_ref, 2.0 * M * N * K / ms_ref * 1e-9, herr);
  if (check) printf("  max|C - C_dgemm| = %.3e (max|C| = %.3e, ratio %.2e)", max_err, max_ref, max_err / max_ref);
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR); cudaFree(pA); cudaFree(pB); cudaFree(eA); cudaFree(eB); cudaFree(dErr);
  return herr;
}

int main() {
  cublasHandle_t h; cublasCreate(&h);
  if (run(256, 256, 256, true, h)) return 1;
  if (run(512, 1024, 2048, true, h)) return 1;
  run(4096, 4096, 16384, true, h);
  run(37888, 4096, 4096, false, h);
  return 0;
}
