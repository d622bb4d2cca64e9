// Exploration for the next round: INT8 tcgen05 (kind::i8) GEMM on sm_100a, the building block of an Ozaki-style FP64
// emulation (FP64 DMMA tops out at 37 TFLOP/s; dense INT8 is nominally 4.5 POPS).  Standalone, not part of the library.
//   C[M][N] (int32) = A[M][K] (int8, K contiguous) * B[N][K]^T (int8, K contiguous)
// One CTA per 128 x 256 tile, 128-byte K chunks (TMA, SWIZZLE_128B), 4-stage ring, accumulator in TMEM (256 columns),
// thread 0 = TMA producer, thread 32 = MMA issuer, all four warps drain TMEM.  Every wait is BOUNDED: a protocol bug
// shows up as an error flag and wrong numbers, never as a hung GPU.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/umma_i8 tools/microbench/umma_i8.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int TM = 128, TN = 256;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

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
template <int KC>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major; SWIZZLE_128B (SBO 1024 B, type 2) or SWIZZLE_32B (SBO 256 B, type 6)
  return KC == 128 ? ((uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61))
                   : ((uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (16ull << 32) | (1ull << 46) | (6ull << 61));
}

template <int KC, int STAGES>
__global__ void __launch_bounds__(128, 1)
k_umma_i8(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int32_t* __restrict__ C, int N, int K,
          int* __restrict__ err) {
  constexpr int A_BYTES = TM * KC, B_BYTES = TN * KC, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bars = base;                       // full[STAGES], empty[STAGES], tmem_full, tmem_slot
  const uint32_t full = bars, empty = bars + 256, tfull = bars + 512, slot = bars + 520;
  const uint32_t ring = base + 1024;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bn = blockIdx.x, bm = blockIdx.y;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8 * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty + 8 * s));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tfull));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const int nk = K / KC;
  bool ok = true;
  if (tid == 0) {                                   // ---- TMA producer ----
    for (int c = 0; c < nk && ok; ++c) {
      const int s = c % STAGES;
      if (c >= STAGES) ok = wait_bounded(empty + 8 * s, (uint32_t)(((c / STAGES) - 1) & 1));
      if (!ok) break;
      const uint32_t dst = ring + s * STAGE_BYTES, bar = full + 8 * s;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE_BYTES) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                   "l"(&mapA), "r"(c * KC), "r"(bm * TM), "r"(bar) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       dst + A_BYTES), "l"(&mapB), "r"(c * KC), "r"(bn * TN), "r"(bar) : "memory");
    }
    if (!ok) atomicExch(err, 1);
  } else if (tid == 32) {                           // ---- MMA issuer ----
    // instruction descriptor: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), K-major both, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    for (int c = 0; c < nk && ok; ++c) {
      const int s = c % STAGES;
      ok = wait_bounded(full + 8 * s, (uint32_t)((c / STAGES) & 1));
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = ring + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
      for (int k = 0; k < KC / 32; ++k) {           // K = 32 per instruction = 32 bytes = 2 descriptor address units
        const uint64_t da = make_desc<KC>(sa) + (uint64_t)(2 * k), db = make_desc<KC>(sb) + (uint64_t)(2 * k);
        const uint32_t acc = (c > 0 || k > 0) ? 1u : 0u;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * s) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tfull) : "memory");
    if (!ok) atomicExch(err, 2);
  }
  __syncwarp();
  // ---- epilogue: all four warps; warp w owns TMEM lanes 32w .. 32w+31 = tile rows ----
  bool got = wait_bounded(tfull, 0);
  if (!got && lane == 0) atomicExch(err, 3);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (got) {
    const int row = bm * TM + warp * 32 + lane;
    for (int c0 = 0; c0 < TN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
          "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      int4* dst = reinterpret_cast<int4*>(C + (size_t)row * N + bn * TN + c0);
#pragma unroll
      for (int q = 0; q < 8; ++q) dst[q] = make_int4((int)v[4 * q], (int)v[4 * q + 1], (int)v[4 * q + 2], (int)v[4 * q + 3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void make_map(CUtensorMap* m, const int8_t* p, int rows, int K, int box_rows, int KC) {
  static EncodeFn enc = nullptr;
  if (!enc) {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    enc = (EncodeFn)f;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)K};
  const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)p, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   KC == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

template <int KC, int STAGES>
static int run(int M, int N, int K, bool check) {
  constexpr int STAGE_BYTES = (TM + TN) * KC;
  std::vector<int8_t> hA((size_t)M * K), hB((size_t)N * K);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  for (auto& x : hA) x = (int8_t)((int)(rnd() % 255) - 127);
  for (auto& x : hB) x = (int8_t)((int)(rnd() % 255) - 127);
  int8_t *dA, *dB; int32_t* dC; int* dErr;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dC, (size_t)M * N * 4)); CK(cudaMalloc(&dErr, 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)M * N * 4)); CK(cudaMemset(dErr, 0, 4));
  alignas(64) CUtensorMap mA, mB;
  make_map(&mA, dA, M, K, TM, KC); make_map(&mB, dB, N, K, TN, KC);
  const size_t smem = 1024 + 1024 + (size_t)STAGES * STAGE_BYTES;
  CK(cudaFuncSetAttribute(k_umma_i8<KC, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(N / TN, M / TM);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_umma_i8<KC, STAGES><<<grid, 128, smem>>>(mA, mB, dC, N, K, dErr);
  CK(cudaDeviceSynchronize());
  int herr = 0; CK(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
  const int reps = check ? 1 : 5;
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) k_umma_i8<KC, STAGES><<<grid, 128, smem>>>(mA, mB, dC, N, K, dErr);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
  long long bad = 0;
  if (check) {
    std::vector<int32_t> hC((size_t)M * N);
    CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < M; i += 7)
      for (int j = 0; j < N; j += 5) {
        long long s = 0;
        for (int k = 0; k < K; ++k) s += (int)hA[(size_t)i * K + k] * (int)hB[(size_t)j * K + k];
        if ((int32_t)s != hC[(size_t)i * N + j]) { if (bad < 5) printf("  mismatch (%d,%d): got %d want %lld\n", i, j, hC[(size_t)i * N + j], s); ++bad; }
      }
  }
  printf("KC=%d stages=%d M=%d N=%d K=%d: %.3f ms  %.1f TOPS  err_flag=%d%s\n", KC, STAGES, M, N, K, ms, 2.0 * M * N * K / ms * 1e-9, herr,
         check ? (bad ? "  MISMATCH" : "  exact") : "");
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dErr);
  return (herr || bad) ? 1 : 0;
}

int main() {
  int rc = 0;
  rc |= run<128, 4>(128, 256, 128, true);
  rc |= run<128, 4>(256, 512, 1024, true);
  rc |= run<32, 16>(256, 512, 1024, true);
  if (rc) { printf("correctness failed; skipping the throughput runs\n"); return 1; }
  // same tile, same bytes in flight (192 KB): 128-byte K rows (4 MMAs per stage) against 32-byte K rows (1 MMA per stage)
  run<128, 4>(8192, 8192, 8192, false);
  run<32, 16>(8192, 8192, 8192, false);
  run<128, 4>(4096, 4096, 131072, false);
  run<32, 16>(4096, 4096, 131072, false);
  return 0;
}
