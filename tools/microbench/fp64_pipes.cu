// Microbenchmark: FP64 pipe characterisation on B200 (sm_100a).
//   * DMMA throughput per instruction shape (m8n8k4, m16n8k4, m16n8k8, m16n8k16)
//   * plain DFMA throughput
//   * DMMA + DMUL issued together (do they share the FP64 pipe?)
//   * fragment-layout check of m16n8k16 against a scalar reference
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

constexpr int NACC = 8;   // independent accumulator tiles per warp

template <int SHAPE>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double seed) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = seed + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = seed - 1e-9 * (threadIdx.x + i);
  double c[NACC][4];
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      if (SHAPE == 0) { double (&cc)[2] = *reinterpret_cast<double (*)[2]>(&c[j][0]); mma884(cc, a[0], b[0]); }
      if (SHAPE == 1) { const double (&aa)[2] = *reinterpret_cast<const double (*)[2]>(&a[0]); mma1684(c[j], aa, b[0]); }
      if (SHAPE == 2) { const double (&aa)[4] = *reinterpret_cast<const double (*)[4]>(&a[0]);
                        const double (&bb)[2] = *reinterpret_cast<const double (*)[2]>(&b[0]); mma1688(c[j], aa, bb); }
      if (SHAPE == 3) { mma16816(c[j], a, b); }
    }
  }
  double s = 0; for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double seed) {
  double x[16];
  for (int i = 0; i < 16; ++i) x[i] = seed + i;
  double m = 1.0 + 1e-12 * seed, ad = 1e-13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], m, ad);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: every warp issues NACC m16n8k16 MMAs plus NMUL independent DMULs per iteration
template <int NMUL>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double seed) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = seed + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = seed - 1e-9 * (threadIdx.x + i);
  double c[NACC][4];
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  double x[NMUL > 0 ? NMUL : 1];
  for (int i = 0; i < NMUL; ++i) x[i] = 1.0 + 1e-3 * i;
  double m = 1.0 + 1e-12 * seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma16816(c[j], a, b);
#pragma unroll
    for (int i = 0; i < NMUL; ++i) x[i] = x[i] * m;
  }
  double s = 0; for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  for (int i = 0; i < NMUL; ++i) s += x[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp-specialised mix: warps 0..3 MMA, warps 4..7 DMUL chains
__global__ void __launch_bounds__(256) k_split(double* out, int iters, double seed, int mul_per_iter) {
  int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp < 4) {
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = seed + 1e-9 * (threadIdx.x + i);
    for (int i = 0; i < 4; ++i) b[i] = seed - 1e-9 * (threadIdx.x + i);
    double c[NACC][4];
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) mma16816(c[j], a, b);
    }
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  } else {
    double x[16];
    for (int i = 0; i < 16; ++i) x[i] = 1.0 + 1e-3 * i;
    double m = 1.0 + 1e-12 * seed;
    for (int it = 0; it < iters * mul_per_iter; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] * m;
    }
    for (int i = 0; i < 16; ++i) s += x[i];
  }
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// layout check: one warp computes D = A(16x16,row) * B(16x8,col) with m16n8k16
__global__ void k_layout(const double* A, const double* B, double* D) {
  int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[8], b[4], c[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) a[i] = A[(g + 8 * (i & 1)) * 16 + (t + 4 * (i >> 1))];
  for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];      // B[k][n]
  mma16816(c, a, b);
  for (int i = 0; i < 4; ++i) D[(g + 8 * (i >> 1)) * 8 + 2 * t + (i & 1)] = c[i];
}

template <typename F>
static float time_ms(F launch, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s sms=%d clock=%d kHz\n", prop.name, sms, prop.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024));
  // layout check
  {
    std::vector<double> hA(256), hB(128), hD(128), ref(128, 0.0);
    for (int i = 0; i < 256; ++i) hA[i] = sin(0.37 * i + 0.1);
    for (int i = 0; i < 128; ++i) hB[i] = cos(0.11 * i + 0.3);
    for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < 16; ++k) s += hA[m * 16 + k] * hB[k * 8 + n]; ref[m * 8 + n] = s; }
    double *dA, *dB, *dD; CK(cudaMalloc(&dA, 256 * 8)); CK(cudaMalloc(&dB, 128 * 8)); CK(cudaMalloc(&dD, 128 * 8));
    CK(cudaMemcpy(dA, hA.data(), 256 * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), 128 * 8, cudaMemcpyHostToDevice));
    k_layout<<<1, 32>>>(dA, dB, dD); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, 128 * 8, cudaMemcpyDeviceToHost));
    double err = 0; for (int i = 0; i < 128; ++i) err = fmax(err, fabs(hD[i] - ref[i]));
    printf("layout_check m16n8k16 max_abs_err=%.3e\n", err);
  }
  const int iters = 20000;
  for (int cps = 1; cps <= 4; cps *= 2) {   // CTAs per SM (8 warps each)
    int grid = sms * cps;
    const char* names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    const double fmas[4] = {256, 1024, 2048, 4096};
    for (int s = 0; s < 4; ++s) {
      float ms = 0;
      if (s == 0) ms = time_ms([&] { k_dmma<0><<<grid, 256>>>(out, iters, 1.0); });
      if (s == 1) ms = time_ms([&] { k_dmma<1><<<grid, 256>>>(out, iters, 1.0); });
      if (s == 2) ms = time_ms([&] { k_dmma<2><<<grid, 256>>>(out, iters, 1.0); });
      if (s == 3) ms = time_ms([&] { k_dmma<3><<<grid, 256>>>(out, iters, 1.0); });
      double total = (double)grid * 8 * NACC * iters * fmas[s];
      printf("dmma %-9s ctas/sm=%d  %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM @1965MHz)\n", names[s], cps, ms,
             2 * total / ms * 1e-9, total / sms / (ms * 1e-3 * 1.965e9));
    }
    {
      float ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0); });
      double total = (double)grid * 256 * 16 * iters;
      printf("dfma           ctas/sm=%d  %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM @1965MHz)\n", cps, ms, 2 * total / ms * 1e-9,
             total / sms / (ms * 1e-3 * 1.965e9));
    }
  }
  // mixed (same warp): NACC MMAs (8*4096 FMA) + NMUL warp-DMULs (32 lanes each) per iteration
  {
    int grid = sms * 2;
    float m0 = time_ms([&] { k_mixed<0><<<grid, 256>>>(out, iters, 1.0); });
    float m8 = time_ms([&] { k_mixed<8><<<grid, 256>>>(out, iters, 1.0); });
    float m32 = time_ms([&] { k_mixed<32><<<grid, 256>>>(out, iters, 1.0); });
    float m64 = time_ms([&] { k_mixed<64><<<grid, 256>>>(out, iters, 1.0); });
    printf("mixed same-warp: 8xm16n8k16 + {0,8,32,64} DMUL/iter: %.3f %.3f %.3f %.3f ms\n", m0, m8, m32, m64);
    printf("  if shared pipe expect +%.1f%% +%.1f%% +%.1f%%\n", 100.0 * 8 * 32 / (8 * 4096), 100.0 * 32 * 32 / (8 * 4096), 100.0 * 64 * 32 / (8 * 4096));
  }
  {
    int grid = sms * 2;
    float s0 = time_ms([&] { k_split<<<grid, 256>>>(out, iters, 1.0, 0); });
    float s1 = time_ms([&] { k_split<<<grid, 256>>>(out, iters, 1.0, 1); });
    float s4 = time_ms([&] { k_split<<<grid, 256>>>(out, iters, 1.0, 4); });
    float s16 = time_ms([&] { k_split<<<grid, 256>>>(out, iters, 1.0, 16); });
    printf("split warps (4 MMA warps + 4 DMUL warps x{0,1,4,16}x16 DMUL/iter): %.3f %.3f %.3f %.3f ms\n", s0, s1, s4, s16);
  }
  return 0;
}
