// Row- and column-partitioned Khatri-Rao mat-vec (reference: tensors/khatri_rao_matrix.py:156-167, RowColKhatriRaoMatrix.__mul__,
// "mvKRrowcol" of the GP-GRIEF paper):
//   y[i] = sum_j x[j] * prod_t G_t[i, j],   G_t = R_t C_t   (R_t: rows x m_t, C_t: m_t x cols; C_t = K_t C_t when the caller has a K)
// The reference materialises `n_rows_at_once` rows of the product at a time (get_rows: one small GEMM per factor, then the
// Hadamard product, then a dot with x).  Here nothing is materialised: a CTA owns 64 rows, its R rows sit in shared memory, the
// threads sweep the columns (lane = column, so every load of C_t[g][j] is coalesced and shared through L1 by the four warps that work
// on different rows of the same columns), each thread carries 16 rows in registers: the per-factor inner products (sum_t m_t FMAs
// per element), the running Hadamard product and the x-weighted row sums.  A factor whose R_t is a SELECTION matrix
// (tensors/selection_matrix.py:55-107: one 1 per row) is passed as an index vector and costs one gathered load instead of m_t FMAs.
// Reduction order is fixed (lanes by butterfly, then the column phases in sequence): results are run-to-run identical.
// Bound: FP64 pipe (rows * cols * sum m_t FMAs) for dense R; L2 -> SM reads of C (cols * sum m_t * 8 B per 64 rows) for gathers.
#include <algorithm>

#include "plan.h"

namespace grief {

constexpr int kKrMaxFactors = 32;
constexpr int kKrRowsPerThread = 16;
constexpr int kKrRowGroups = 4;                 // 64 rows per CTA
constexpr int kKrPhases = 2;                    // column phases (warps that share a row group)
constexpr int kKrThreads = 32 * kKrRowGroups * kKrPhases;

struct KrParams {
  int d;
  int m[kKrMaxFactors];
  int r_off[kKrMaxFactors];                     // offset (doubles) of factor t inside a row's shared-memory record; -1: gather factor
  const double* R[kKrMaxFactors];               // rows x m_t row-major, or nullptr
  const int* ridx[kKrMaxFactors];               // rows, or nullptr
  const double* C[kKrMaxFactors];               // m_t x cols row-major
  int64_t rows, cols;
  int rec;                                      // doubles per row record (sum of m_t over dense factors, odd)
  const double* x;
  double* y;
};

__global__ void __launch_bounds__(kKrThreads) k_rowcol_kr_matvec(const KrParams P) {
  extern __shared__ double sm[];
  constexpr int RPT = kKrRowsPerThread, CTA_ROWS = RPT * kKrRowGroups;
  double* sR = sm;                                                     // [CTA_ROWS][rec]
  int* sIdx = reinterpret_cast<int*>(sm + (size_t)CTA_ROWS * P.rec);   // [d][CTA_ROWS]
  double* sRed = reinterpret_cast<double*>(sIdx + (size_t)P.d * CTA_ROWS + ((P.d * CTA_ROWS) & 1));   // [phases][CTA_ROWS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = warp % kKrRowGroups, phase = warp / kKrRowGroups;
  const int64_t row0 = (int64_t)blockIdx.x * CTA_ROWS;
  for (int t = 0; t < P.d; ++t) {
    if (P.r_off[t] >= 0) {
      const int mt = P.m[t];
      for (int e = threadIdx.x; e < CTA_ROWS * mt; e += kKrThreads) {
        const int r = e / mt, g = e - r * mt;
        sR[(size_t)r * P.rec + P.r_off[t] + g] = (row0 + r < P.rows) ? P.R[t][(row0 + r) * mt + g] : 0.0;
      }
    } else {
      for (int r = threadIdx.x; r < CTA_ROWS; r += kKrThreads) sIdx[t * CTA_ROWS + r] = (row0 + r < P.rows) ? P.ridx[t][row0 + r] : 0;
    }
  }
  __syncthreads();
  double acc[RPT];
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r] = 0.0;
  const double* myR = sR + (size_t)(rg * RPT) * P.rec;
  const int* myIdx = sIdx + rg * RPT;
  for (int64_t j = phase * 32 + lane; j < P.cols; j += 32 * kKrPhases) {
    double prod[RPT];
    const double xj = P.x[j];
#pragma unroll
    for (int r = 0; r < RPT; ++r) prod[r] = xj;
    for (int t = 0; t < P.d; ++t) {
      const double* Ct = P.C[t] + j;
      if (P.r_off[t] < 0) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) prod[r] *= __ldg(Ct + (size_t)myIdx[t * CTA_ROWS + r] * P.cols);
      } else {
        double tmp[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) tmp[r] = 0.0;
        const double* Rt = myR + P.r_off[t];
        for (int g = 0; g < P.m[t]; ++g) {
          const double c = __ldg(Ct + (size_t)g * P.cols);
#pragma unroll
          for (int r = 0; r < RPT; ++r) tmp[r] = fma(Rt[(size_t)r * P.rec + g], c, tmp[r]);
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) prod[r] *= tmp[r];
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r] += prod[r];
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const double s = warp_sum(acc[r]);
    if (lane == 0) sRed[phase * CTA_ROWS + rg * RPT + r] = s;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < CTA_ROWS; r += kKrThreads) {
    double s = 0.0;
#pragma unroll
    for (int ph = 0; ph < kKrPhases; ++ph) s += sRed[ph * CTA_ROWS + r];
    if (row0 + r < P.rows) P.y[row0 + r] = s;
  }
}

int launch_rowcol_kr_matvec(int d, const int32_t* m, const double* const* R, const int32_t* const* ridx, const double* const* C, int64_t rows,
                            int64_t cols, const double* x, double* y, cudaStream_t stream) {
  GRIEF_REQUIRE(d >= 1 && d <= kKrMaxFactors, "rowcol_kr_matvec: d=%d outside [1,%d]", d, kKrMaxFactors);
  GRIEF_REQUIRE(rows >= 0 && cols >= 0, "rowcol_kr_matvec: rows=%lld cols=%lld", (long long)rows, (long long)cols);
  if (rows == 0) return GRIEF_OK;
  KrParams P;
  P.d = d; P.rows = rows; P.cols = cols; P.x = x; P.y = y;
  int rec = 0;
  for (int t = 0; t < d; ++t) {
    GRIEF_REQUIRE(m[t] >= 1, "rowcol_kr_matvec: factor %d has inner size %d", t, m[t]);
    GRIEF_REQUIRE(C[t] != nullptr && ((R && R[t]) != (ridx && ridx[t])), "rowcol_kr_matvec: factor %d needs C and exactly one of R / ridx", t);
    P.m[t] = m[t];
    P.C[t] = C[t];
    const bool gather = ridx && ridx[t];
    P.R[t] = gather ? nullptr : R[t];
    P.ridx[t] = gather ? ridx[t] : nullptr;
    P.r_off[t] = gather ? -1 : rec;
    if (!gather) rec += m[t];
  }
  for (int t = d; t < kKrMaxFactors; ++t) { P.m[t] = 0; P.r_off[t] = -1; P.R[t] = nullptr; P.ridx[t] = nullptr; P.C[t] = nullptr; }
  rec = std::max(rec, 1) | 1;                   // odd stride: the 16 rows of a thread start in different banks
  P.rec = rec;
  constexpr int CTA_ROWS = kKrRowsPerThread * kKrRowGroups;
  const size_t smem = (size_t)CTA_ROWS * rec * sizeof(double) + ((size_t)d * CTA_ROWS + 1) * sizeof(int) + (size_t)kKrPhases * CTA_ROWS * sizeof(double) + 16;
  if (smem > 220 * 1024)
    return fail(GRIEF_ERR_UNSUPPORTED, "rowcol_kr_matvec: %d inner entries per row exceed shared memory (64 rows x %d doubles)", rec, rec);
  GRIEF_CUDA(cudaFuncSetAttribute(k_rowcol_kr_matvec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned blocks = (unsigned)((rows + CTA_ROWS - 1) / CTA_ROWS);
  k_rowcol_kr_matvec<<<blocks, kKrThreads, smem, stream>>>(P);
  GRIEF_CUDA(cudaGetLastError());
  return GRIEF_OK;
}

}  // namespace grief
