// grief_plan: the device-resident description of one GRIEF basis (one value of the kernel
// hyper-parameters): per-dimension grids, kernel parameters and scaled eigenvectors, plus the
// "group table" layout that every row kernel shares.
//
// Basis column j of Phi is a product over input dimensions (reference: tensors/tensors.py:116-124,
// kern/grief_kernel.py:104).  The FP64 pipe of a B200 SM is shared by DMMA, DFMA and DMUL
// (profiles/r01_fp64_pipes_microbench.txt), so every multiply spent building a Phi element is
// taken from the Gram MMA.  The plan therefore partitions the dimensions into G contiguous groups
// and enumerates, per group, the DISTINCT index sub-tuples that occur among the p columns; the
// prepass kernel evaluates one table entry per distinct sub-tuple per data row, and the MMA
// kernels build a Phi element with G gathers and G-1 multiplies instead of d gathers and d-1.
#pragma once
#include <cstdint>
#include <vector>
#include "common.cuh"

namespace grief {

// KERN_HOST: the kernel of this dimension is evaluated by the caller (K_xu columns uploaded per call, grief_build_tables_kxu)
enum KernelId : int { KERN_RBF = 0, KERN_EXPONENTIAL = 1, KERN_MATERN32 = 2, KERN_MATERN52 = 3, KERN_HOST = 4 };

struct DimDesc {          // one input dimension, device-visible POD
  int m;                  // grid points
  int u;                  // unique selected eigen-indices
  int kernel;             // KernelId
  int grid_off;           // offset into grid[]      (sum of m over previous dims)
  int q_off;              // offset into qs[]        (sum of m*u over previous dims)
  int f_off;              // offset into the per-row factor scratch (sum of u over previous dims)
  double variance;
  double lengthscale;
};

struct GradDesc;
void grad_desc_destroy(GradDesc* gd);

// Arithmetic / staging options of the two O(n p^2) products.  Every plan carries its own copy (taken from the calling thread's
// defaults when the plan is created, changed with grief_plan_set_option): nothing here is process-global.
struct PlanOpts {
  int gemm_mode = 1;                       // 0: FP64 DMMA GEMM (k_gemm_nt), 1: INT8 tensor-core emulation (k_ozaki)
  int cluster = 1;                         // INT8 mode, CTA pairs (cta_group::2, 256 x 128 tiles): 0 never, 1 (default) where both tile
                                           // counts are >= 16, 2 wherever there are two row tiles
  int digits_gram = 6;                     // INT8 digits per operand of A = Phi^T Phi (46-bit operands + the diagonal pair, ozaki.cu)
  int digits_z = 4;                        // INT8 digits per operand of Zp = Phi P^-1 in the gradient pass (30-bit; the rank-one part is FP64)
  int digits_var = 6;                      // INT8 digits per operand of Z = Phi B in grief_quadform_rows (predictive variance)
  size_t slab_budget = (size_t)5 << 28;    // 1.25 GiB of Phi^T (as FP64) staged per pass-1 slab, rounded down to whole builder waves
                                           // (148 x 128 rows): 37888 rows at p = 4096, so one K split
                                           // (~12600 rows x p x digits) passes through L2 in pieces that the tiles in flight share
};
PlanOpts& default_plan_opts();             // thread-local defaults for plans created on this thread

struct Plan {
  // ---- host copies ----
  int d = 0, p = 0, p_pad = 0;          // p_pad = p rounded up to kTileN
  int n_groups = 0;                     // G
  int width = 0;                        // table entries per row incl. the 2 constant slots
  int stride = 0;                       // row stride in doubles (odd, >= width)
  int sum_m = 0, sum_u = 0;
  int max_group_dims = 0;
  int n_host_dims = 0;                  // dimensions with KERN_HOST
  std::vector<DimDesc> dims;
  std::vector<int> group_begin;         // G+1 dimension boundaries
  std::vector<int> group_slot0;         // first table slot of each group
  std::vector<int> group_size;          // distinct sub-tuples per group
  std::vector<uint16_t> col_slot_h;     // p_pad x G slot index per column and group
  // ---- device buffers (all carved out of d_pool) ----
  void* d_pool = nullptr;
  DimDesc* d_dims = nullptr;
  double* d_grid = nullptr;             // concatenated grids
  double* d_qs = nullptr;               // concatenated m_i x u_i scaled eigenvectors (row-major [g][k])
  uint8_t* d_slot_k = nullptr;          // width x max_group_dims: unique-index of every dim of the slot's group
  int* d_slot_group = nullptr;          // width: group of the slot (-1 for the constant slots)
  int* d_group_begin = nullptr;         // G+1
  uint16_t* d_col_slot = nullptr;       // p_pad x G, external column order (row kernels)
  // Tile builders (k_gram, k_zgemm) walk the columns in LEXICOGRAPHIC slot order (smallest group = major key):
  // consecutive columns then share their leading slots and only the factors from `level` on are re-gathered.
  uint16_t* d_sorted_slot = nullptr;    // p_pad x G: slots of sorted column c, listed in key order
  uint8_t* d_sorted_level = nullptr;    // p_pad: first key position where sorted column c differs from c-1 (G: identical)
  uint32_t* d_sorted_pack = nullptr;    // p_pad x pack_words: the key-order slots of sorted column c, one byte each, 4 per word (MSB first)
  int pack_words = 1;                   // (G + 3) / 4
  uint32_t* d_gram_order = nullptr;     // n_gram_order entries (pair row << 16 | column tile): launch order of the Gram's lower pair tiles
  int n_gram_order = 0;
  int* d_perm = nullptr;                // p_pad: external column of sorted column c (-1 for padding columns)
  std::vector<int> perm_h;
  int device = 0;
  PlanOpts opts;
  int* d_err = nullptr;                 // device error flag of this plan's INT8 launches (see OzOpts::err)
  GradDesc* grad = nullptr;             // set by grief_grad_setup
  ~Plan();
};

// Builds the plan.  uinv is p x d (row-major): uinv[j*d+i] indexes dimension i's unique list.
// qs holds, per dimension, an m_i x u_i row-major matrix.  Returns GRIEF_OK or an error code.
int plan_create(Plan** out, int d, const int32_t* m, const int32_t* kernel_id, const double* variance,
                const double* lengthscale, const double* grid_concat, const int32_t* u,
                const double* qs_concat, int p, const int32_t* uinv, int width_cap);

// TMA-fed FP64 DMMA GEMM (dense.cu):  C = beta*C + alpha * A * B^T,  A (M x K) and B (N x K) row-major, K contiguous.
struct GemmOpts {
  bool lower_only = false;   // skip tiles strictly above the block diagonal
  bool store_t = false;      // write C^T
  bool tri_k = false;        // K loop starts at the block row (operands upper block triangular in K)
  int splits = 1;            // split K over gridDim.z; split z accumulates into C + z * c_split_stride
  int64_t c_split_stride = 0;
  int rows_a = 0, rows_b = 0;   // rows of A / B present in memory (0: all M / N); the rest reads as zeros
};
int gemm_nt_ex(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N, int K, double alpha,
               double beta, const GemmOpts& opts, cudaStream_t stream, int* launches);

// FP64 GEMM on the INT8 tensor cores (ozaki.cu).  Digit planes: [digits][rows_alloc][K rounded up to 32] int8; buffers are always
// sized for kOzMaxDigits planes, `digits` of them are written and read.
constexpr int kOzMinDigits = 3, kOzMaxDigits = 7;
struct OzOpts {
  int digits = kOzMaxDigits;   // balanced 8-bit digits per operand: 8 * digits - 2 bits + sign below the row maximum
  int cluster = 0;             // CTA pairs with tcgen05.mma.cta_group::2 on 256 x 128 tiles: 0 never, 1 for >= 16 x 16 tiles, 2 always
  int store_t = 0;             // 1: C is written transposed (element (m, n) at C[n * ldc + m])
  int diag_pair = 0;           // 1: both operands are the same matrix (Gram): keep the pair a = b = digits / 2 of group g = digits
  const uint32_t* tile_order = nullptr;   // lower_only with CTA pairs: (pair row << 16 | column tile) per cluster, n_tile_order entries
  int n_tile_order = 0;
  int* err = nullptr;          // device int: 1-3 barrier time-out in k_ozaki, 4 non-finite operand value
};
size_t ozaki_plane_bytes(int64_t rows, int K);
int ozaki_slice(const double* X, int64_t ld, int rows, int K, int* exps, int exps_len, int8_t* planes, int digits, int* err, cudaStream_t stream);
int ozaki_gemm(const int8_t* pa, int64_t rows_a_alloc, const int* ea, int M, const int8_t* pb, int64_t rows_b_alloc, const int* eb, int N,
               int K, double* C, int64_t ldc, bool accumulate, bool lower_only, int splits, int64_t c_split_stride, const OzOpts& opt,
               cudaStream_t stream, int* launches);
int ozaki_check(int* err, cudaStream_t stream);

}  // namespace grief
