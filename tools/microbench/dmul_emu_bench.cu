// Correctness and co-issue behaviour of the integer-pipe double multiply (gp_grief_b200/csrc/dmul_emu.cuh).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dmul_emu_bench tools/microbench/dmul_emu_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "dmul_emu.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
using namespace grief;

__device__ __forceinline__ void mma1684(double (&c)[4], double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b));
}

template <bool EX>
__global__ void k_check(const double* a, const double* b, int n, unsigned long long* counts, double* worst) {
  unsigned long long mism = 0, off2 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double r = a[i] * b[i], q = dmul_emu<EX>(a[i], b[i]);
    const long long rb = __double_as_longlong(r), qb = __double_as_longlong(q);
    if (rb != qb && !(r != r && q != q)) {
      ++mism;
      long long d = rb - qb; if (d < 0) d = -d;
      if (d > 1) { ++off2; worst[0] = a[i]; worst[1] = b[i]; }
    }
  }
  atomicAdd(&counts[0], mism);
  atomicAdd(&counts[1], off2);
}

// MODE 0: warps 4..7 idle; 1: real DMUL; 2: dmul_emu exact; 3: dmul_emu approximate carry.   4 MMA warps + 4 multiplier warps.
template <int MODE>
__global__ void __launch_bounds__(256) k_split(double* out, int iters, double seed, int mul_per_iter) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp < 4) {
    double a0 = seed + 1e-9 * threadIdx.x, a1 = seed - 2e-9 * threadIdx.x, b = seed - 1e-9 * threadIdx.x;
    double c[8][4];
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) mma1684(c[j], a0, a1, b);
    }
    for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  } else if (MODE != 0) {
    double x[16];
    for (int i = 0; i < 16; ++i) x[i] = 1.0 + 1e-3 * i + 1e-7 * threadIdx.x;
    const double m = 1.0 + 1e-12 * seed;
    for (int it = 0; it < iters * mul_per_iter; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = (MODE == 1) ? x[i] * m : (MODE == 2 ? dmul_emu<true>(x[i], m) : dmul_emu<false>(x[i], m));
    }
    for (int i = 0; i < 16; ++i) s += x[i];
  }
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Anti-phase pattern of the tile kernels: 16 warps; every warp alternates a burst of NMUL multiplies (8 in flight, chained
// twice) with 64 DMMAs; half of the warps start with the burst, the other half with the DMMAs; __syncthreads per round.
template <int MODE>
__global__ void __launch_bounds__(512) k_phase(double* out, int iters, double seed, int nbatch) {
  const int warp = threadIdx.x >> 5;
  const bool first = ((warp >> 2) & 1) == 0;
  double a0 = seed + 1e-9 * threadIdx.x, a1 = seed - 2e-9 * threadIdx.x, b = seed - 1e-9 * threadIdx.x;
  double c[8][4];
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = 1.0 + 1e-3 * i + 1e-7 * threadIdx.x;
  const double m = 1.0 + 1e-12 * seed;
  auto burst = [&]() {
    for (int bt = 0; bt < nbatch; ++bt) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (MODE == 1) ? x[i] * m : (MODE == 2 ? dmul_emu<true>(x[i], m) : (MODE == 3 ? dmul_emu<false>(x[i], m) : x[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (MODE == 1) ? x[i] * m : (MODE == 2 ? dmul_emu<true>(x[i], m) : (MODE == 3 ? dmul_emu<false>(x[i], m) : x[i]));
    }
  };
  for (int it = 0; it < iters; ++it) {
    if (first) burst();
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) mma1684(c[j], a0, a1, b);
    if (!first) burst();
    __syncthreads();
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float time_ms(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s sms=%d\n", prop.name, sms);
  {
    const int n = 1 << 24;
    std::vector<double> ha(n), hb(n);
    uint64_t st = 88172645463325252ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    const double specials[] = {0.0, -0.0, 1.0, -1.0, 0.5, 2.0, 3.0, 1e-310, -1e-310, 1e308, 1e-308, 2.2250738585072014e-308,
                               1.7976931348623157e308, 1.0 / 0.0, -1.0 / 0.0, 0.0 / 0.0, 1.0000000000000002, 0.9999999999999999,
                               134217729.0, 134217727.0, 6755399441055744.0};
    const int ns = sizeof(specials) / sizeof(double);
    for (int i = 0; i < n; ++i) {
      uint64_t ua = rnd(), ub = rnd();
      if (i < n / 2) {   // moderate exponents (what the tables hold): 2^-40 .. 2^40
        ua = (ua & 0x800fffffffffffffull) | ((uint64_t)(1023 - 40 + (rnd() % 81)) << 52);
        ub = (ub & 0x800fffffffffffffull) | ((uint64_t)(1023 - 40 + (rnd() % 81)) << 52);
      }
      memcpy(&ha[i], &ua, 8); memcpy(&hb[i], &ub, 8);
      if (i >= n - ns * ns) { const int k = i - (n - ns * ns); ha[i] = specials[k / ns]; hb[i] = specials[k % ns]; }
      if (i >= n / 2 && i < n / 2 + 100000) {   // short significands: exact products and exact ties
        ha[i] = (double)(rnd() % (1u << 27)) + 1.0; hb[i] = (double)(rnd() % (1u << 27)) + 1.0;
      }
    }
    double *da, *db, *dw; unsigned long long* dc;
    CK(cudaMalloc(&da, n * 8)); CK(cudaMalloc(&db, n * 8)); CK(cudaMalloc(&dc, 16)); CK(cudaMalloc(&dw, 16));
    CK(cudaMemcpy(da, ha.data(), n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), n * 8, cudaMemcpyHostToDevice));
    for (int ex = 1; ex >= 0; --ex) {
      CK(cudaMemset(dc, 0, 16)); CK(cudaMemset(dw, 0, 16));
      if (ex) k_check<true><<<sms * 4, 256>>>(da, db, n, dc, dw); else k_check<false><<<sms * 4, 256>>>(da, db, n, dc, dw);
      CK(cudaDeviceSynchronize());
      unsigned long long hc[2]; double hw[2];
      CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hw, dw, 16, cudaMemcpyDeviceToHost));
      printf("check %s: %d pairs, %llu differ from a*b (1 ulp), %llu differ by more than 1 ulp (last such pair %.17g %.17g)\n",
             ex ? "exact-carry" : "approx-carry", n, hc[0], hc[1], hw[0], hw[1]);
    }
  }
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 512));
  const int iters = 20000;
  for (int mpi : {1, 2, 4}) {
    const int grid = sms * 2;
    float t0 = time_ms([&] { k_split<0><<<grid, 256>>>(out, iters, 1.0, mpi); });
    float t1 = time_ms([&] { k_split<1><<<grid, 256>>>(out, iters, 1.0, mpi); });
    float t2 = time_ms([&] { k_split<2><<<grid, 256>>>(out, iters, 1.0, mpi); });
    float t3 = time_ms([&] { k_split<3><<<grid, 256>>>(out, iters, 1.0, mpi); });
    printf("split 4 MMA warps (32 m16n8k4/iter) + 4 multiplier warps (%d x16 mul/iter), 2 CTAs/SM: none %.3f  DMUL %.3f  emu-exact %.3f  emu-approx %.3f ms\n",
           mpi, t0, t1, t2, t3);
  }
  for (int nb : {1, 2, 4}) {
    float t0 = time_ms([&] { k_phase<0><<<sms, 512>>>(out, iters, 1.0, nb); });
    float t1 = time_ms([&] { k_phase<1><<<sms, 512>>>(out, iters, 1.0, nb); });
    float t2 = time_ms([&] { k_phase<2><<<sms, 512>>>(out, iters, 1.0, nb); });
    float t3 = time_ms([&] { k_phase<3><<<sms, 512>>>(out, iters, 1.0, nb); });
    const double ideal = (double)iters * 64 * 4 * 32 / 1.965e6;   // 64 DMMA m16n8k4 per warp, 4 warps per sub-partition, 32 clk each
    printf("anti-phase 16 warps (64 m16n8k4 + %d x16 mul per round): none %.3f  DMUL %.3f  emu-exact %.3f  emu-approx %.3f ms  (pipe-bound ideal %.3f @1965MHz)\n",
           nb, t0, t1, t2, t3, ideal);
  }
  return 0;
}
