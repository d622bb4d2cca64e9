// G1: top-p selection over the Kronecker product of per-dimension grid eigenvalues.
//
// Reference being replaced: KronMatrix.find_extremum_eigs(mode='largest', log_expand=True, sort=True)
// (tensors/kron_matrix.py:369-446) with log_kron (linalg.py:74-89).  The m^d candidate products are
// never enumerated: like the reference, a d-step beam keeps the p best partial products (exact for
// positive eigenvalues -- any prefix outside the p best prefixes is dominated by p full products).
// Candidate values are formed exactly as NumPy forms them -- one IEEE double add of the running log
// value and the host-computed log eigenvalue (`a.reshape(-1,1) + b.reshape(1,-1)`, candidate index
// = prev * m_i + j) -- so the selected SET is bit-identical to the reference whenever the p-th and
// (p+1)-th candidate differ (no tie at the boundary); ties are broken towards the smaller
// candidate index (NumPy's introselect order is implementation-defined there).
//
// One CTA of 1024 threads runs the whole beam: per step an 8-pass MSD radix select (8-bit digits,
// warp-aggregated shared-memory histogram) finds the p-th largest key, an ordered block scan
// compacts the survivors and records (parent, choice) back-pointers; a bitonic sort of the final
// p values in shared memory produces the descending order, and the index tuples are rebuilt by
// walking the back-pointers.
#include "plan.h"

namespace grief {

constexpr int kTopkThreads = 1024;

__device__ __forceinline__ uint64_t order_key(double v) {
  const uint64_t u = (uint64_t)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

struct TopkParams {
  const double* raw0;      // raw eigenvalues of factor 0 (selection key of the first step, kron_matrix.py:407)
  const double* logeig;    // concatenated log eigenvalues of all factors
  const int* m;            // factor sizes
  const int* off;          // offsets into logeig
  int d, p;
  double* vals_a;          // p scratch
  double* vals_b;          // p scratch
  uint16_t* parent;        // d x p
  uint8_t* choice;         // d x p
  int32_t* idx_out;        // p x d
  double* loglam_out;      // p
  int* n_out;              // number of eigenvalues actually returned (min(p, prod m))
};

// value of candidate c at step `step` (step 0: raw eigenvalue of factor 0)
__device__ __forceinline__ double cand_value(const TopkParams& P, int step, const double* cur, int mi, int c,
                                             const double* le) {
  if (step == 0) return P.raw0[c];
  const int prev = c / mi;
  return cur[prev] + le[c - prev * mi];
}

__global__ void __launch_bounds__(kTopkThreads, 1) k_topk(const TopkParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_krem;
  __shared__ int s_warp_gt[32], s_warp_eq[32];
  __shared__ int s_base_gt, s_base_eq, s_base_out;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  double* cur = P.vals_a;
  double* nxt = P.vals_b;
  int n_cur = 1;
  for (int step = 0; step < P.d; ++step) {
    const int mi = P.m[step];
    const double* le = P.logeig + P.off[step];
    const int C = (step == 0) ? mi : n_cur * mi;
    uint16_t* par = P.parent + (size_t)step * P.p;
    uint8_t* cho = P.choice + (size_t)step * P.p;
    if (C <= P.p) {                                    // kron_matrix.py:397-398 -- keep everything, in order
      for (int c = tid; c < C; c += kTopkThreads) {
        const int prev = (step == 0) ? 0 : c / mi;
        const int j = (step == 0) ? c : c - prev * mi;
        nxt[c] = (step == 0) ? le[c] : cur[prev] + le[j];
        par[c] = (uint16_t)prev;
        cho[c] = (uint8_t)j;
      }
      n_cur = C;
    } else {
      // ---- radix select: key of the p-th largest candidate ----
      if (tid == 0) { s_prefix = 0ull; s_krem = P.p; }
      for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int c0 = 0; c0 < C; c0 += kTopkThreads) {
          const int c = c0 + tid;
          bool act = false;
          unsigned int bin = 0;
          if (c < C) {
            const uint64_t key = order_key(cand_value(P, step, cur, mi, c, le));
            act = (pass == 0) || ((key >> (shift + 8)) == prefix);
            bin = (unsigned int)((key >> shift) & 255ull);
          }
          // warp-aggregated histogram update
          const unsigned int amask = __ballot_sync(0xffffffffu, act);
          if (act) {
            const unsigned int peers = __match_any_sync(amask, bin);
            if (lane == (__ffs(peers) - 1)) atomicAdd(&hist[bin], (unsigned int)__popc(peers));
          }
        }
        __syncthreads();
        if (tid == 0) {
          int krem = s_krem;
          int cum = 0, b = 255;
          for (; b > 0; --b) {
            if (cum + (int)hist[b] >= krem) break;
            cum += (int)hist[b];
          }
          s_krem = krem - cum;                         // rank inside bin b
          s_prefix = (prefix << 8) | (unsigned long long)b;
        }
        __syncthreads();
      }
      const uint64_t tau = s_prefix;                   // exact key of the p-th largest candidate
      const int take_eq = s_krem;                      // how many candidates equal to tau are kept
      if (tid == 0) { s_base_gt = 0; s_base_eq = 0; s_base_out = 0; }
      __syncthreads();
      // ---- ordered compaction (candidate order), ties at tau resolved towards small candidate index ----
      for (int c0 = 0; c0 < C; c0 += kTopkThreads) {
        const int c = c0 + tid;
        double v = 0.0;
        bool gt = false, eq = false;
        if (c < C) {
          v = cand_value(P, step, cur, mi, c, le);
          const uint64_t key = order_key(v);
          gt = key > tau;
          eq = key == tau;
        }
        const unsigned int bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) { s_warp_gt[warp] = __popc(bg); s_warp_eq[warp] = __popc(be); }
        __syncthreads();
        int pre_gt = s_base_gt, pre_eq = s_base_eq;
        for (int w2 = 0; w2 < warp; ++w2) { pre_gt += s_warp_gt[w2]; pre_eq += s_warp_eq[w2]; }
        const unsigned int lower = (1u << lane) - 1u;
        const int my_gt = pre_gt + __popc(bg & lower);          // # greater-than before me (global)
        const int my_eq = pre_eq + __popc(be & lower);          // # equal before me (global)
        const bool keep = gt || (eq && my_eq < take_eq);
        if (keep) {
          const int pos = my_gt + min(my_eq, take_eq);          // kept elements before me
          const int prev = (step == 0) ? 0 : c / mi;
          const int j = (step == 0) ? c : c - prev * mi;
          nxt[pos] = (step == 0) ? le[c] : v;                    // first step: log taken after selection (:410)
          par[pos] = (uint16_t)prev;
          cho[pos] = (uint8_t)j;
        }
        __syncthreads();
        if (tid == 0) {
          int tg = 0, te = 0;
          for (int w2 = 0; w2 < (kTopkThreads >> 5); ++w2) { tg += s_warp_gt[w2]; te += s_warp_eq[w2]; }
          s_base_gt += tg;
          s_base_eq += te;
        }
        __syncthreads();
      }
      n_cur = P.p;
    }
    __syncthreads();
    double* t = cur; cur = nxt; nxt = t;
    __threadfence_block();
  }

  // ---- sort the final beam: value descending, beam position ascending on ties ----
  const int n_fin = n_cur;
  int N = 1;
  while (N < n_fin) N <<= 1;
  double* skey = reinterpret_cast<double*>(smem_raw);
  int* sidx = reinterpret_cast<int*>(smem_raw + (size_t)N * sizeof(double));
  for (int e = tid; e < N; e += kTopkThreads) {
    skey[e] = (e < n_fin) ? cur[e] : -__longlong_as_double(0x7ff0000000000000ll) * 1.0;  // -inf padding
    sidx[e] = e;
  }
  __syncthreads();
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = tid; e < N; e += kTopkThreads) {
        const int partner = e ^ j;
        if (partner > e) {
          const bool desc_block = ((e & k) == 0);      // first half of each 2k block sorted "before" order
          const double ka = skey[e], kb = skey[partner];
          const int ia = sidx[e], ib = sidx[partner];
          // "a before b" in final order: larger value first, smaller position on ties; padding (idx>=n_fin) last
          const bool a_before_b = (ka > kb) || (ka == kb && ia < ib);
          const bool swap = desc_block ? !a_before_b : a_before_b;
          if (swap) { skey[e] = kb; skey[partner] = ka; sidx[e] = ib; sidx[partner] = ia; }
        }
      }
      __syncthreads();
    }
  }
  // ---- rebuild index tuples from the back-pointers ----
  for (int q = tid; q < n_fin; q += kTopkThreads) {
    int pos = sidx[q];
    P.loglam_out[q] = skey[q];
    for (int step = P.d - 1; step >= 0; --step) {
      P.idx_out[(size_t)q * P.d + step] = (int32_t)P.choice[(size_t)step * P.p + pos];
      pos = (int)P.parent[(size_t)step * P.p + pos];
    }
  }
  if (tid == 0) *P.n_out = n_fin;
}

size_t topk_scratch_bytes(int d, int p) {
  size_t b = 0;
  b += 2 * (size_t)p * sizeof(double);                 // vals_a, vals_b
  b += (size_t)d * p * sizeof(uint16_t);               // parent
  b += (size_t)d * p;                                  // choice
  b += (size_t)kMaxDims * kMaxGrid * sizeof(double) * 2;  // raw0 + logeig upload area
  b += 2 * kMaxDims * sizeof(int) + sizeof(int) + 1024;
  return b;
}

// Host pointers in, device pointers out (idx_dev: p x d int32, loglam_dev: p doubles).
int launch_topk(int d, const int32_t* m_host, const double* raw0_host, const double* logeig_host, int p,
                int32_t* idx_dev, double* loglam_dev, int* n_out_host, cudaStream_t stream, int* launches) {
  GRIEF_REQUIRE(d >= 1 && d <= kMaxDims, "topk: d=%d outside [1,%d]", d, kMaxDims);
  GRIEF_REQUIRE(p >= 1 && p <= 16384, "topk: p=%d outside [1,16384]", p);
  std::vector<int> off(d), mm(d);
  int tot = 0;
  double total = 1.0;
  for (int i = 0; i < d; ++i) {
    GRIEF_REQUIRE(m_host[i] >= 1 && m_host[i] <= 255, "topk: m[%d]=%d outside [1,255]", i, m_host[i]);
    off[i] = tot;
    mm[i] = m_host[i];
    tot += m_host[i];
    total *= m_host[i];
  }
  GRIEF_REQUIRE(total >= p, "topk: n_eigs=%d exceeds the %g eigenvalues of the grid", p, total);
  const size_t bytes_vals = 2 * (size_t)p * sizeof(double);
  const size_t bytes_par = ((size_t)d * p * sizeof(uint16_t) + 15) / 16 * 16;
  const size_t bytes_cho = ((size_t)d * p + 15) / 16 * 16;
  const size_t bytes_eig = ((size_t)(tot + m_host[0]) * sizeof(double) + 15) / 16 * 16;
  const size_t bytes_int = ((size_t)(2 * d + 1) * sizeof(int) + 15) / 16 * 16;
  // grow-only scratch kept between calls (one evaluation = one call; a stream-ordered alloc/free pair per call showed
  // erratic 10-800 ms host stalls after long GPU phases)
  // (per thread AND per device: a scratch pointer is only valid on the device it was allocated on)
  constexpr int kMaxDev = 64;
  static thread_local char* scratch_dev[kMaxDev] = {};
  static thread_local size_t scratch_bytes_dev[kMaxDev] = {};
  int dev = 0;
  GRIEF_CUDA(cudaGetDevice(&dev));
  GRIEF_REQUIRE(dev >= 0 && dev < kMaxDev, "topk: device ordinal %d", dev);
  char*& scratch = scratch_dev[dev];
  size_t& scratch_bytes = scratch_bytes_dev[dev];
  const size_t need = bytes_vals + bytes_par + bytes_cho + bytes_eig + bytes_int;
  if (need > scratch_bytes) {
    if (scratch) cudaFree(scratch);
    scratch = nullptr;
    scratch_bytes = 0;
    GRIEF_CUDA(cudaMalloc(reinterpret_cast<void**>(&scratch), need));
    scratch_bytes = need;
  }
  char* q = scratch;
  TopkParams P;
  P.vals_a = reinterpret_cast<double*>(q); P.vals_b = P.vals_a + p; q += bytes_vals;
  P.parent = reinterpret_cast<uint16_t*>(q); q += bytes_par;
  P.choice = reinterpret_cast<uint8_t*>(q); q += bytes_cho;
  double* d_eig = reinterpret_cast<double*>(q); q += bytes_eig;
  int* d_int = reinterpret_cast<int*>(q);
  std::vector<double> heig(tot + m_host[0]);
  for (int i = 0; i < tot; ++i) heig[i] = logeig_host[i];
  for (int i = 0; i < m_host[0]; ++i) heig[tot + i] = raw0_host[i];
  std::vector<int> hint(2 * d + 1, 0);
  for (int i = 0; i < d; ++i) { hint[i] = mm[i]; hint[d + i] = off[i]; }
  cudaError_t e1 = cudaMemcpyAsync(d_eig, heig.data(), heig.size() * sizeof(double), cudaMemcpyHostToDevice, stream);
  cudaError_t e2 = cudaMemcpyAsync(d_int, hint.data(), hint.size() * sizeof(int), cudaMemcpyHostToDevice, stream);
  if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(GRIEF_ERR_CUDA, "topk: upload failed");
  P.logeig = d_eig; P.raw0 = d_eig + tot; P.m = d_int; P.off = d_int + d; P.n_out = d_int + 2 * d;
  P.d = d; P.p = p; P.idx_out = idx_dev; P.loglam_out = loglam_dev;
  int N = 1;
  while (N < p) N <<= 1;
  const size_t smem = (size_t)N * (sizeof(double) + sizeof(int));
  cudaError_t e3 = cudaFuncSetAttribute(k_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e3 == cudaSuccess) {
    prof_begin(PROF_TOPK, stream);
    k_topk<<<1, kTopkThreads, smem, stream>>>(P);
    prof_end(PROF_TOPK, stream);
    e3 = cudaGetLastError();
  }
  int n_out = 0;
  if (e3 == cudaSuccess) e3 = cudaMemcpyAsync(&n_out, P.n_out, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(stream);   // host staging vectors die at return
  if (e3 != cudaSuccess) return fail(GRIEF_ERR_CUDA, "topk: %s", cudaGetErrorString(e3));
  if (n_out_host) *n_out_host = n_out;
  if (launches) *launches += 1;
  return GRIEF_OK;
}

}  // namespace grief
