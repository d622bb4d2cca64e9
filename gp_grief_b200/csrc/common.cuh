// Shared helpers for the GP-GRIEF sm_100a kernels: error plumbing, PTX wrappers
// (mbarrier, bulk async copy = the 1-D TMA path, FP64 DMMA), small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "grief_b200.h"

namespace grief {

// ---- error plumbing -------------------------------------------------------------------------
// error codes: the public GRIEF_ERR_* macros of include/grief_b200.h

void set_last_error(const std::string& msg);
int fail(int code, const char* fmt, ...);

#define GRIEF_CUDA(expr)                                                                       \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::grief::fail(GRIEF_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, \
                           __LINE__, cudaGetErrorString(e__));                                 \
  } while (0)

#define GRIEF_REQUIRE(cond, ...)                                               \
  do {                                                                         \
    if (!(cond)) return ::grief::fail(GRIEF_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

// ---- optional per-kernel timing (CUDA events on the launch stream; off by default) --------------------
enum ProfSlot : int { PROF_GRAM = 0, PROF_ZGEMM, PROF_TABLES, PROF_CONTRACT, PROF_TOPK, PROF_SOLVE, PROF_PHITY, PROF_DTABLES, PROF_BUILD_T, PROF_BUILD, PROF_COUNT };
void prof_begin(int slot, cudaStream_t stream);
void prof_end(int slot, cudaStream_t stream);

// ---- shape constants shared by host and device ------------------------------------------------
constexpr int kMaxDims = 64;       // input dimensions
constexpr int kMaxGrid = 64;       // grid points per dimension
constexpr int kMaxGroups = 8;      // table groups gathered per basis column
constexpr int kTileN = 128;        // Gram / GEMM tile edge (columns of Phi per block)
constexpr int kChunk = 16;         // rows of Phi per pipeline stage (= one m16n8k16 K step)
constexpr int kTableCap = 116;     // max table row width (doubles): 3 x 32-row ring stages + two 32-row Phi tile buffers fit 227 KB

#ifdef __CUDACC__
// ---- PTX: mbarrier + bulk async copy ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- PTX: FP64 tensor-core MMA (lowers to 8 x DMMA.8x8x4 on sm_100a) ---------------------------
// D(16x8) += A(16x16,row) * B(16x8,col).  Fragment layout (lane = 4*g + t):
//   a[2*ks+h] = A[g + 8h][t + 4ks],  b[ks] = B[t + 4ks][g],  c[2*hh + e] = C[g + 8hh][2t + e].
__device__ __forceinline__ void dmma_16x8x16(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
      "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
        "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// D(16x8) += A(16x4,row) * B(4x8,col):  a[h] = A[g + 8h][t],  b = B[t][g]  (2 x DMMA.8x8x4, independent halves)
__device__ __forceinline__ void dmma_16x8x4(double (&c)[4], double a0, double a1, double b) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a0), "d"(a1), "d"(b));
}

// A short not-ready window for the calling warp (two dependent clock reads).  Issued by the MMA warps once per
// 32 DMMAs: without it the warps that build Phi tiles wait several hundred cycles for a slot in the FP64 pipe
// (ncu: builder DMULs in stall_math); measured +1..4 % on k_gram (profiles/r01_gram_design_notes.md).
__device__ __forceinline__ void pipe_window() {
  unsigned c1, c2;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c1));
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c2));
  if (c1 - c2 == 0x7fffffffu) asm volatile("trap;");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__

}  // namespace grief
