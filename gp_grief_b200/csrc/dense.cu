// Dense p x p building blocks of the G3 stage, hand-written for sm_100a (no cuSOLVER / cuBLAS):
//   k_gemm_nt      C = beta*C + alpha * A * B^T   both operands through 2-D TMA tensor maps (SWIZZLE_128B),
//                  4-stage mbarrier ring, FP64 DMMA core (16 warps, warp tile 32 x 32), optional lower-triangle-only
//                  tiles and transposed store
//   k_potf2_inv    Cholesky of one 128 x 128 diagonal block in shared memory + the inverse of its factor
//   k_trsv_step    one block step of a blocked triangular solve with one right-hand side, one CTA per dependent block row
// and the drivers built from them: blocked right-looking Cholesky, triangular inverse, P^-1 = W^T W.
// All matrices handled here are q x q with q a multiple of 128 (the caller pads P with an identity block), row-major.
// Reference being replaced: scipy.linalg.cho_factor / cho_solve at models/gp_grief_model.py:153, :175, :234.
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "plan.h"

namespace grief {

int gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N, int K,
            double alpha, double beta, bool lower_only, bool store_t, cudaStream_t stream, int* launches, bool tri_k = false);

constexpr int kDB = 128;                 // block size of the blocked algorithms = GEMM tile edge
constexpr int kGemmThreads = 512;
constexpr int kGemmStages = 4;
constexpr int kGemmStageBytes = 2 * kDB * kChunk * 8;   // A tile + B tile, 16 KB each

struct GemmParams {
  double* C;
  int64_t ldc;
  int K;
  double alpha, beta;
  int lower_only;      // skip tiles strictly above the block diagonal
  int store_t;         // write C^T: element (m, n) goes to C[n * ldc + m]
  int tri_k;           // operands are upper block triangular in K: start the K loop at block bm (needs bm >= bn)
  int split_chunks;    // > 0: blockIdx.z owns K chunks [z * split_chunks, (z + 1) * split_chunks) and the C copy z
  int64_t c_split_stride;
  int tiles_m, tiles_n;
  int m_valid, n_valid;  // C has m_valid x n_valid elements; the rest of the last tiles is not stored
  int group_n;         // rasterisation: CTAs walk all M tiles of a group of group_n N tiles before the next group, so the
                       // group's B panels stay in L2 and A is streamed tiles_n / group_n times
};

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// grid = (tiles along N, tiles along M, K splits); A is M x K and B is N x K, both row-major with K contiguous.
// Stage ring: full[s] is completed by the two TMA tile loads of a chunk, empty[s] by one arrival per warp once its
// fragment loads of that chunk are done.  No CTA-wide barrier in the main loop: warps drift by up to a stage, so the
// fragment-load latency of one warp hides behind the DMMAs of the others.  Thread 0 refills, at iteration c, the stage
// that was consumed at iteration c-1 (every warp has normally released it by then).
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_nt(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmParams prm) {
  int bn, bm;
  {
    const int per_group = prm.group_n * prm.tiles_m;
    const int g = (int)blockIdx.x / per_group, rem = (int)blockIdx.x - g * per_group;
    const int gcols = min(prm.group_n, prm.tiles_n - g * prm.group_n);
    bm = rem / gcols;
    bn = g * prm.group_n + (rem - bm * gcols);
  }
  if (prm.lower_only && bn > bm) return;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;      // 1024-B aligned: SWIZZLE_128B atoms
  const uint32_t full = base, empty = base + 64, ring = base + 1024;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, t4 = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  if (tid == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full + 8 * s), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty + 8 * s), "r"(kGemmThreads / 32));
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int c_first = prm.tri_k ? bm * (kDB / kChunk) : (int)blockIdx.z * prm.split_chunks;
  int nk = (prm.K + kChunk - 1) / kChunk - c_first;
  if (prm.split_chunks > 0 && nk > prm.split_chunks) nk = prm.split_chunks;
  double* const Cz = prm.C + (size_t)blockIdx.z * prm.c_split_stride;
  auto issue = [&](int c) {
    const int st = c % kGemmStages;
    const uint32_t dst = ring + (uint32_t)st * kGemmStageBytes;
    const uint32_t bar = full + 8 * st;
    fence_proxy_async();
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kGemmStageBytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(&mapA), "r"((c_first + c) * kChunk), "r"(bm * kDB), "r"(bar)
                 : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     dst + kGemmStageBytes / 2),
                 "l"(&mapB), "r"((c_first + c) * kChunk), "r"(bn * kDB), "r"(bar)
                 : "memory");
  };
  auto wait = [&](uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
  };
  if (tid == 0)
    for (int c = 0; c < kGemmStages && c < nk; ++c) issue(c);
  double acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.0;
  const uint32_t offA = (uint32_t)(wm * 32 + g4) * 128u, offB = (uint32_t)(kGemmStageBytes / 2) + (uint32_t)(wn * 32 + g4) * 128u;
  for (int c = 0; c < nk; ++c) {
    const int st = c % kGemmStages;
    if (tid == 0 && c >= 1 && c - 1 + kGemmStages < nk) {
      wait(empty + 8 * ((c - 1) % kGemmStages), (uint32_t)(((c - 1) / kGemmStages) & 1));
      issue(c - 1 + kGemmStages);
    }
    wait(full + 8 * st, (uint32_t)((c / kGemmStages) & 1));
    const uint32_t pa = ring + (uint32_t)st * kGemmStageBytes + offA;
    const uint32_t pb = ring + (uint32_t)st * kGemmStageBytes + offB;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const uint32_t unit = (uint32_t)(((2 * t4 + hf) ^ g4) << 4);      // SWIZZLE_128B: 16-byte unit index XOR (row & 7)
      double2 av[2][2], bv[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bv[nt] = lds128(pb + nt * 8 * 128 + unit);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) av[mt][h] = lds128(pa + (mt * 16 + 8 * h) * 128 + unit);
      if (hf == 1) {                                                     // this warp is done reading the stage
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty + 8 * st) : "memory");
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma_16x8x4(acc[mt][nt], av[mt][0].x, av[mt][1].x, bv[nt].x);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma_16x8x4(acc[mt][nt], av[mt][0].y, av[mt][1].y, bv[nt].y);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int m = bm * kDB + wm * 32 + mt * 16 + g4 + 8 * hh;
        const int n = bn * kDB + wn * 32 + nt * 8 + 2 * t4;
        double v0 = prm.alpha * acc[mt][nt][2 * hh], v1 = prm.alpha * acc[mt][nt][2 * hh + 1];
        if (m >= prm.m_valid || n >= prm.n_valid) continue;
        const bool two = n + 1 < prm.n_valid;
        if (prm.store_t) {
          double* d0 = Cz + (size_t)n * prm.ldc + m;
          double* d1 = d0 + prm.ldc;
          if (prm.beta != 0.0) { v0 += prm.beta * (*d0); if (two) v1 += prm.beta * (*d1); }
          *d0 = v0;
          if (two) *d1 = v1;
        } else {
          double* d0 = Cz + (size_t)m * prm.ldc + n;
          if (two && ((reinterpret_cast<uintptr_t>(d0) & 15) == 0)) {     // 16-byte aligned pair
            double2* d2 = reinterpret_cast<double2*>(d0);
            if (prm.beta != 0.0) { const double2 o = *d2; v0 += prm.beta * o.x; v1 += prm.beta * o.y; }
            *d2 = make_double2(v0, v1);
          } else {
            if (prm.beta != 0.0) { v0 += prm.beta * d0[0]; if (two) v1 += prm.beta * d0[1]; }
            d0[0] = v0;
            if (two) d0[1] = v1;
          }
        }
      }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  return fn;
}

static int make_map(CUtensorMap* map, const double* base, int rows, int cols, int64_t ld) {
  EncodeTiledFn2 enc = encode_fn();
  if (!enc) return fail(GRIEF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)kChunk, (cuuint32_t)kDB};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(GRIEF_ERR_CUDA, "cuTensorMapEncodeTiled failed with code %d", (int)cr);
  return GRIEF_OK;
}

// C (M x N, ldc) = beta*C + alpha * A (M x K, lda) * B (N x K, ldb)^T.   M, N multiples of 128; lda, ldb even.
// rows_a / rows_b: rows of A / B that exist in memory (the rest of the M / N range reads as zeros through the TMA
// out-of-bounds fill); splits > 1: split K over blockIdx.z, split z accumulating into C + z * c_split_stride.
int gemm_nt_ex(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N, int K, double alpha,
               double beta, const GemmOpts& o, cudaStream_t stream, int* launches) {
  if (M <= 0 || N <= 0) return GRIEF_OK;
  GRIEF_REQUIRE(K >= 0, "gemm_nt: K=%d", K);
  GRIEF_REQUIRE(!(o.tri_k && o.splits > 1), "gemm_nt: tri_k and split K do not combine");
  GRIEF_REQUIRE(lda % 2 == 0 && ldb % 2 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                "gemm_nt: operands need 16-byte aligned rows (even leading dimensions)");
  const int m_valid = M, n_valid = N;
  M = (M + kDB - 1) / kDB * kDB;                     // partial last tiles: absent rows read as zeros, stores are guarded
  N = (N + kDB - 1) / kDB * kDB;
  alignas(64) CUtensorMap mA, mB;
  const int Kmap = std::max(K, 2);
  int rc = make_map(&mA, A, o.rows_a > 0 ? o.rows_a : m_valid, Kmap, lda);
  if (rc == GRIEF_OK) rc = make_map(&mB, B, o.rows_b > 0 ? o.rows_b : n_valid, Kmap, ldb);
  if (rc != GRIEF_OK) return rc;
  GemmParams prm;
  prm.C = C; prm.ldc = ldc; prm.K = K; prm.alpha = alpha; prm.beta = beta;
  prm.lower_only = o.lower_only ? 1 : 0; prm.store_t = o.store_t ? 1 : 0; prm.tri_k = o.tri_k ? 1 : 0;
  const int splits = std::max(1, o.splits);
  const int chunks = (K + kChunk - 1) / kChunk;
  prm.split_chunks = splits > 1 ? (chunks + splits - 1) / splits : 0;
  prm.c_split_stride = o.c_split_stride;
  const size_t smem = 1024 + 1024 + (size_t)kGemmStages * kGemmStageBytes;
  // function attributes are per device: set on every launch (a host-side table lookup in the runtime)
  GRIEF_CUDA(cudaFuncSetAttribute(k_gemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prm.m_valid = m_valid;
  prm.n_valid = n_valid;
  prm.tiles_m = M / kDB;
  prm.tiles_n = N / kDB;
  // B panels of one group (group_n x 128 rows x the K range of a CTA) are sized to ~40 % of the 126 MB L2
  const int64_t k_cta = splits > 1 ? (int64_t)prm.split_chunks * kChunk : (int64_t)K;
  const int64_t panel_bytes = std::max<int64_t>(1, k_cta * kDB * 8);
  prm.group_n = (int)std::max<int64_t>(1, std::min<int64_t>(prm.tiles_n, ((int64_t)50 << 20) / panel_bytes));
  if (o.lower_only) prm.group_n = prm.tiles_n;       // triangle: plain row-major tile order
  dim3 grid((unsigned)(prm.tiles_m * prm.tiles_n), 1, splits);
  k_gemm_nt<<<grid, kGemmThreads, smem, stream>>>(mA, mB, prm);
  GRIEF_CUDA(cudaGetLastError());
  if (launches) *launches += 1;
  return GRIEF_OK;
}

int gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N, int K,
            double alpha, double beta, bool lower_only, bool store_t, cudaStream_t stream, int* launches, bool tri_k) {
  GemmOpts o;
  o.lower_only = lower_only; o.store_t = store_t; o.tri_k = tri_k;
  GRIEF_REQUIRE(K > 0, "gemm_nt: K=%d", K);
  return gemm_nt_ex(A, lda, B, ldb, C, ldc, M, N, K, alpha, beta, o, stream, launches);
}

// sum over k = k_begin, k_begin + 4, ... < k_end of a[k] * b[k], with four independent chains so that the shared-memory
// loads of the next terms are in flight while the previous FMAs retire
__device__ __forceinline__ double strided_dot4(const double* __restrict__ a, const double* __restrict__ b, int k_begin, int k_end) {
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  int k = k_begin;
  for (; k + 12 < k_end; k += 16) {
    const double a0 = a[k], a1 = a[k + 4], a2 = a[k + 8], a3 = a[k + 12];
    const double b0 = b[k], b1 = b[k + 4], b2 = b[k + 8], b3 = b[k + 12];
    t0 = fma(a0, b0, t0);
    t1 = fma(a1, b1, t1);
    t2 = fma(a2, b2, t2);
    t3 = fma(a3, b3, t3);
  }
  for (; k < k_end; k += 4) t0 = fma(a[k], b[k], t0);
  return (t0 + t1) + (t2 + t3);
}

// ---- diagonal block: Cholesky (lower, in place, strictly upper part zeroed) and the inverse of the factor ----
// info: 0, or (global index + 1) of the first non-positive pivot (only the first failure is recorded).
__global__ void __launch_bounds__(512) k_potf2_inv(double* __restrict__ A, int64_t lda, int k0, double* __restrict__ Linv,
                                                   int* __restrict__ info) {
  extern __shared__ double S[];                    // 128 x 129, then 128 reciprocal pivots, then the pivot being broadcast
  const int LD = kDB + 1;
  double* rdiag = S + kDB * LD;
  double* piv = rdiag + kDB;
  const int tid = threadIdx.x;
  for (int e = tid; e < kDB * kDB; e += 512) {
    const int i = e / kDB, j = e - i * kDB;
    S[i * LD + j] = A[(size_t)(k0 + i) * lda + k0 + j];
  }
  __syncthreads();
  // left-looking Cholesky: four lanes share one row and split the dot product over the finished columns; two barriers a column
  const int row = tid >> 2, part = tid & 3;
  for (int j = 0; j < kDB; ++j) {
    double t = 0.0;
    if (row >= j) t = strided_dot4(S + row * LD, S + j * LD, part, j);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t = S[row * LD + j] - t;
    if (tid == 4 * j) piv[0] = t;
    __syncthreads();
    const double pj = piv[0];
    if (!(pj > 0.0) && tid == 0) atomicCAS(info, 0, k0 + j + 1);       // also catches NaN; only the first failure sticks
    const double rd = rsqrt(pj);                   // one long-latency op per column instead of sqrt + divide
    if (part == 0) {
      if (row == j) {
        S[j * LD + j] = pj * rd;
        rdiag[j] = rd;
      } else if (row > j) {
        S[row * LD + j] = t * rd;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < kDB * kDB; e += 512) {
    const int i = e / kDB, j = e - i * kDB;
    A[(size_t)(k0 + i) * lda + k0 + j] = (j <= i) ? S[i * LD + j] : 0.0;
  }
  // inverse X = L^-1 by forward substitution, four lanes per column c.  X[i][c] (i > c) is kept in the free upper triangle
  // of S at S[c][i]; a column only ever reads its own row c of that triangle, so only the four lanes need to agree.
  {
    const int c = row;
    const double xcc = rdiag[c];
    for (int i = 1; i < kDB; ++i) {
      double acc = 0.0;
      if (i > c) {
        acc = strided_dot4(S + i * LD, S + c * LD, c + 1 + part, i);
        if (part == 0) acc = fma(S[i * LD + c], xcc, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (i > c && part == 0) S[c * LD + i] = -acc * rdiag[i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < kDB * kDB; e += 512) {
    const int i = e / kDB, c = e - i * kDB;
    Linv[e] = (i == c) ? rdiag[i] : (i > c ? S[c * LD + i] : 0.0);
  }
}

// ---- blocked triangular solves with one right-hand side ----
// One block step k, spread over the block rows that still depend on it (one CTA each): every CTA first recomputes the
// 128 solved unknowns x_k = Linv_kk v_k (forward) or Linv_kk^T v_k (backward) -- 16 K FMAs, cheaper than another launch --
// and then removes their contribution from its own 128 entries of the working right-hand side v.
//   forward:  L z = r      v_i -= L[i, k] x_k      for block rows i > k
//   backward: L^T b = z    v_i -= L[k, i]^T x_k    for block rows i < k
// CTA 0 also writes x_k to out (entries below n_out only).  Launched in block order on one stream.
__global__ void __launch_bounds__(256) k_trsv_step(const double* __restrict__ L, int64_t ld, const double* __restrict__ Linv, int q,
                                                   double* __restrict__ v, double* __restrict__ out, int n_out, int k, int backward) {
  __shared__ double rk[kDB], xk[kDB], part[kDB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = tid & (kDB - 1), half = tid >> 7;
  const int k0 = k * kDB;
  if (tid < kDB) rk[tid] = v[k0 + tid];
  __syncthreads();
  const double* Lk = Linv + (size_t)k * kDB * kDB;
  if (!backward) {
    double a[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {               // 64 independent loads per lane in flight
      const double* Lr = Lk + (warp * 16 + rr) * kDB;
      a[rr] = Lr[lane] * rk[lane] + Lr[lane + 32] * rk[lane + 32] + Lr[lane + 64] * rk[lane + 64] + Lr[lane + 96] * rk[lane + 96];
    }
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const double t = warp_sum(a[rr]);
      if (lane == rr) xk[warp * 16 + rr] = t;
    }
  } else {
    double a = 0.0;
#pragma unroll 16
    for (int j = half * 64; j < half * 64 + 64; ++j) a = fma(Lk[j * kDB + m], rk[j], a);
    if (half == 1) part[m] = a;
    __syncthreads();
    if (half == 0) xk[m] = a + part[m];
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid < kDB && k0 + tid < n_out) out[k0 + tid] = xk[tid];
  const int nb = q / kDB;
  const int nrem = backward ? k : nb - 1 - k;
  if ((int)blockIdx.x >= nrem) return;
  const int i0 = (backward ? (int)blockIdx.x : k + 1 + (int)blockIdx.x) * kDB;
  if (!backward) {
    double a[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const double* Lr = L + (size_t)(i0 + warp * 16 + rr) * ld + k0;
      a[rr] = Lr[lane] * xk[lane] + Lr[lane + 32] * xk[lane + 32] + Lr[lane + 64] * xk[lane + 64] + Lr[lane + 96] * xk[lane + 96];
    }
    const double mine = (lane < 16) ? v[i0 + warp * 16 + lane] : 0.0;
    double upd = 0.0;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const double t = warp_sum(a[rr]);
      if (lane == rr) upd = t;
    }
    if (lane < 16) v[i0 + warp * 16 + lane] = mine - upd;
  } else {
    double a = 0.0;
#pragma unroll 16
    for (int j = half * 64; j < half * 64 + 64; ++j) a = fma(L[(size_t)(k0 + j) * ld + i0 + m], xk[j], a);
    if (half == 1) part[m] = a;
    __syncthreads();
    if (half == 0) v[i0 + m] -= a + part[m];
  }
}

__global__ void k_pad_vec(const double* __restrict__ in, int n, int q, double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < q; i += gridDim.x * blockDim.x) out[i] = (i < n) ? in[i] : 0.0;
}

// P (q x q, ld q) <- [A + diag(noise/w), 0; 0, I]
__global__ void k_form_P_padded(const double* __restrict__ A, int64_t lda, const double* __restrict__ w, double noise, int p, int q,
                                double* __restrict__ P) {
  const int64_t total = (int64_t)q * q;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / q), j = (int)(e - (int64_t)i * q);
    double v = 0.0;
    if (i < p && j < p) {
      v = A[(size_t)i * lda + j];
      if (i == j) v += noise / w[i];
    } else if (i == j) {
      v = 1.0;
    }
    P[e] = v;
  }
}

// out (p x p, ldo) <- in (q x q)[0:p, 0:p], optionally transposed
__global__ void k_copy_block(const double* __restrict__ in, int q, int p, int transpose, double* __restrict__ out, int64_t ldo,
                             int upper_only = 0) {
  const int64_t total = (int64_t)p * p;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / p), j = (int)(e - (int64_t)i * p);
    const double v = transpose ? in[(size_t)j * q + i] : in[(size_t)i * q + j];
    out[(size_t)i * ldo + j] = (upper_only && j < i) ? 0.0 : v;
  }
}

// lower triangle -> full symmetric (q x q)
__global__ void k_mirror_lower(double* __restrict__ M, int q) {
  const int64_t total = (int64_t)q * q;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / q), j = (int)(e - (int64_t)i * q);
    if (j > i) M[e] = M[(size_t)j * q + i];
  }
}

struct DenseWork {       // q x q workspaces, grown on demand
  int q = 0;
  double *P = nullptr, *Wt = nullptr, *T1t = nullptr, *Linv = nullptr, *X = nullptr;
  ~DenseWork() { release(); }
  void release() {
    cudaFree(P); cudaFree(Wt); cudaFree(T1t); cudaFree(Linv); cudaFree(X);
    P = Wt = T1t = Linv = X = nullptr;
    q = 0;
  }
};

DenseWork* dense_work_new() { return new DenseWork(); }
void dense_work_delete(DenseWork* wk) { delete wk; }
double* dense_work_factor(DenseWork* wk) { return wk->P; }
double* dense_work_inverse(DenseWork* wk) { return wk->X; }
double* dense_work_tmp(DenseWork* wk) { return wk->T1t; }

int dense_work_reserve(DenseWork* wk, int q) {
  if (wk->q >= q) return GRIEF_OK;
  wk->release();
  const size_t qq = (size_t)q * q * sizeof(double);
  GRIEF_CUDA(cudaMalloc(&wk->P, qq));
  GRIEF_CUDA(cudaMalloc(&wk->Wt, qq));
  GRIEF_CUDA(cudaMalloc(&wk->X, qq));
  GRIEF_CUDA(cudaMalloc(&wk->T1t, (size_t)q * kDB * sizeof(double)));
  GRIEF_CUDA(cudaMalloc(&wk->Linv, (size_t)q * kDB * sizeof(double)));
  wk->q = q;
  return GRIEF_OK;
}

// Blocked right-looking Cholesky of wk->P (q x q), in place: lower factor L (strictly upper part of the diagonal blocks
// zeroed, blocks above the diagonal keep their input values and are never read).  Inverses of the diagonal blocks -> wk->Linv.
int dense_potrf(DenseWork* wk, int q, int* d_info, cudaStream_t stream, int* launches) {
  const size_t smem = ((size_t)kDB * (kDB + 1) + kDB + 1) * sizeof(double);
  GRIEF_CUDA(cudaFuncSetAttribute(k_potf2_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
  double* A = wk->P;
  for (int k0 = 0; k0 < q; k0 += kDB) {
    k_potf2_inv<<<1, 512, smem, stream>>>(A, q, k0, wk->Linv + (size_t)(k0 / kDB) * kDB * kDB, d_info);
    GRIEF_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    const int rem = q - k0 - kDB;
    if (rem <= 0) break;
    double* A21 = A + (size_t)(k0 + kDB) * q + k0;
    // L21 = A21 * inv(L11)^T   (in place: each CTA reads only the rows it later overwrites)
    int rc = gemm_nt(A21, q, wk->Linv + (size_t)(k0 / kDB) * kDB * kDB, kDB, A21, q, rem, kDB, kDB, 1.0, 0.0, false, false,
                     stream, launches);
    if (rc != GRIEF_OK) return rc;
    // A22 -= L21 * L21^T   (lower tiles)
    double* A22 = A + (size_t)(k0 + kDB) * q + (k0 + kDB);
    rc = gemm_nt(A21, q, A21, q, A22, q, rem, rem, kDB, -1.0, 1.0, true, false, stream, launches);
    if (rc != GRIEF_OK) return rc;
  }
  return GRIEF_OK;
}

// X = P^-1 (q x q, full symmetric) from the factor in wk->P:  W = L^-1 by right-looking block forward substitution,
// X = W^T W.  Only W^T is ever stored (wk->Wt, upper block triangle): both GEMM forms below want it K-contiguous.
//   step k:  W[k, j]   = -inv(L_kk) * Acc[k, j]           j < k      (Acc^T lives in Wt[j, k] until it is finalised here)
//            W[k, k]   =  inv(L_kk)
//            Acc[i, j] +=  L[i, k] * W[k, j]              i > k, j <= k   -> (nb-k-1) x (k+1) tiles, K = 128
int dense_inverse_from_factor(DenseWork* wk, int q, cudaStream_t stream, int* launches) {
  const size_t qq = (size_t)q * q * sizeof(double);
  GRIEF_CUDA(cudaMemsetAsync(wk->Wt, 0, qq, stream));
  const double* L = wk->P;
  const unsigned cb = 148 * 4;
  for (int k0 = 0; k0 < q; k0 += kDB) {
    const double* Lk = wk->Linv + (size_t)(k0 / kDB) * kDB * kDB;
    int rc;
    if (k0 > 0) {
      // out[m][n] = -sum_c Lk[m][c] * Acc[c][n], Acc[c][n] = Wt[n][k0 + c]; stored transposed over the tile it was read from
      rc = gemm_nt(Lk, kDB, wk->Wt + k0, q, wk->Wt + k0, q, kDB, k0, kDB, -1.0, 0.0, false, true, stream, launches);
      if (rc != GRIEF_OK) return rc;
    }
    k_copy_block<<<64, 256, 0, stream>>>(Lk, kDB, kDB, 1, wk->Wt + (size_t)k0 * q + k0, q, 0);
    GRIEF_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    const int rem = q - k0 - kDB;
    if (rem <= 0) break;
    // Acc[i][j] += sum_c L[i][k0 + c] * W[k0 + c][j], W[k0 + c][j] = Wt[j][k0 + c]; Acc^T accumulates in Wt[j][i]
    rc = gemm_nt(L + (size_t)(k0 + kDB) * q + k0, q, wk->Wt + k0, q, wk->Wt + (k0 + kDB), q, rem, k0 + kDB, kDB, 1.0, 1.0, false,
                 true, stream, launches);
    if (rc != GRIEF_OK) return rc;
  }
  // X = W^T W :  X[m][n] = sum_k Wt[m][k] * Wt[n][k]   (lower tiles, then mirrored)
  int rc = gemm_nt(wk->Wt, q, wk->Wt, q, wk->X, q, q, q, q, 1.0, 0.0, true, false, stream, launches, true);
  if (rc != GRIEF_OK) return rc;
  k_mirror_lower<<<cb, 256, 0, stream>>>(wk->X, q);
  GRIEF_CUDA(cudaGetLastError());
  if (launches) *launches += 1;
  return GRIEF_OK;
}

// b = P^-1 r through the factor: forward then backward substitution.  tmp: at least 2 q doubles of scratch.
int dense_trsv_pair(DenseWork* wk, int q, const double* r, int p, double* b, double* tmp, cudaStream_t stream, int* launches) {
  const int nb = q / kDB;
  double* v = tmp;
  double* z = tmp + q;
  k_pad_vec<<<(q + 255) / 256, 256, 0, stream>>>(r, p, q, v);
  for (int k = 0; k < nb; ++k) k_trsv_step<<<std::max(1, nb - 1 - k), 256, 0, stream>>>(wk->P, q, wk->Linv, q, v, z, q, k, 0);
  for (int k = nb - 1; k >= 0; --k) k_trsv_step<<<std::max(1, k), 256, 0, stream>>>(wk->P, q, wk->Linv, q, z, b, p, k, 1);
  GRIEF_CUDA(cudaGetLastError());
  if (launches) *launches += 1 + 2 * nb;
  return GRIEF_OK;
}

int launch_form_P_padded(const double* A, int64_t lda, const double* w, double noise, int p, int q, double* P, cudaStream_t stream) {
  k_form_P_padded<<<148 * 4, 256, 0, stream>>>(A, lda, w, noise, p, q, P);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}
int launch_copy_block(const double* in, int q, int p, bool transpose, bool upper_only, double* out, int64_t ldo, cudaStream_t stream) {
  k_copy_block<<<148 * 4, 256, 0, stream>>>(in, q, p, transpose ? 1 : 0, out, ldo, upper_only ? 1 : 0);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

}  // namespace grief
