// Host-side construction of a grief::Plan (see plan.h) and the error-string plumbing.
#include "plan.h"

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <numeric>
#include <vector>

namespace grief {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

PlanOpts& default_plan_opts() {
  static thread_local PlanOpts opts;
  return opts;
}

// ---- per-kernel timing ----
struct ProfRec { int slot; cudaEvent_t e0, e1; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> g_prof_recs;
static thread_local std::vector<cudaEvent_t> g_prof_open(PROF_COUNT, nullptr);

void prof_enable(bool on) { g_prof_on = on; }
void prof_begin(int slot, cudaStream_t stream) {
  if (!g_prof_on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  g_prof_open[slot] = e;
}
void prof_end(int slot, cudaStream_t stream) {
  if (!g_prof_on || !g_prof_open[slot]) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, stream);
  g_prof_recs.push_back({slot, g_prof_open[slot], e});
  g_prof_open[slot] = nullptr;
}
// sums (ms) and launch counts per slot since the last read; synchronises on the recorded events
void prof_read(double* ms, int* count) {
  for (int i = 0; i < PROF_COUNT; ++i) { ms[i] = 0.0; count[i] = 0; }
  for (auto& r : g_prof_recs) {
    cudaEventSynchronize(r.e1);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms[r.slot] += t; count[r.slot] += 1; }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof_recs.clear();
}

Plan::~Plan() {
  if (grad) grad_desc_destroy(grad);
  cudaFree(d_pool);          // every d_* array of the plan lives in this one allocation
}

static inline uint64_t mix64(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  return h;
}

// All device arrays of a plan are carved out of ONE allocation and filled by ONE host-to-device copy: a plan is rebuilt at every
// new hyper-parameter value, and a dozen cudaMalloc + cudaMemcpy pairs cost ~6 ms of serial time per evaluation on every rank.
struct PoolBuilder {
  std::vector<unsigned char> host;
  struct Fix { void** dst; size_t off; };
  std::vector<Fix> fixes;
  template <typename T>
  void add(T** dst, const std::vector<T>& src) {
    const size_t off = (host.size() + 255) / 256 * 256;
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    host.resize(off + bytes, 0);
    if (!src.empty()) std::memcpy(host.data() + off, src.data(), src.size() * sizeof(T));
    fixes.push_back({reinterpret_cast<void**>(dst), off});
  }
  int commit(void** pool) {
    GRIEF_CUDA(cudaMalloc(pool, std::max<size_t>(host.size(), 256)));
    GRIEF_CUDA(cudaMemcpy(*pool, host.data(), host.size(), cudaMemcpyHostToDevice));
    for (auto& f : fixes) *f.dst = static_cast<unsigned char*>(*pool) + f.off;
    return GRIEF_OK;
  }
};

int plan_create(Plan** out, int d, const int32_t* m, const int32_t* kernel_id, const double* variance,
                const double* lengthscale, const double* grid_concat, const int32_t* u,
                const double* qs_concat, int p, const int32_t* uinv, int width_cap) {
  GRIEF_REQUIRE(out != nullptr, "plan_create: out is null");
  GRIEF_REQUIRE(d >= 1 && d <= kMaxDims, "plan_create: d=%d outside [1,%d]", d, kMaxDims);
  GRIEF_REQUIRE(p >= 1 && p <= 65535 - kTileN, "plan_create: p=%d outside [1,%d]", p, 65535 - kTileN);
  if (width_cap <= 0 || width_cap > kTableCap) width_cap = kTableCap;
  Plan* pl = new Plan();
  pl->d = d;
  pl->p = p;
  pl->p_pad = (p + kTileN - 1) / kTileN * kTileN;
  cudaGetDevice(&pl->device);
  pl->opts = default_plan_opts();
  pl->dims.resize(d);
  int goff = 0, qoff = 0, foff = 0;
  for (int i = 0; i < d; ++i) {
    DimDesc& dd = pl->dims[i];
    if (m[i] < 1 || m[i] > kMaxGrid || u[i] < 1 || u[i] > m[i] || kernel_id[i] < 0 || kernel_id[i] > KERN_HOST) {
      delete pl;
      return fail(GRIEF_ERR_BAD_ARG, "plan_create: dimension %d has m=%d u=%d kernel=%d (need 1<=u<=m<=%d)", i, m[i],
                  u[i], kernel_id[i], kMaxGrid);
    }
    dd.m = m[i];
    dd.u = u[i];
    dd.kernel = kernel_id[i];
    if (kernel_id[i] == KERN_HOST) pl->n_host_dims += 1;
    dd.grid_off = goff;
    dd.q_off = qoff;
    dd.f_off = foff;
    dd.variance = variance[i];
    dd.lengthscale = lengthscale[i];
    goff += m[i];
    qoff += m[i] * u[i];
    foff += u[i];
  }
  pl->sum_m = goff;
  pl->sum_u = foff;
  for (int64_t j = 0; j < (int64_t)p * d; ++j) {
    int i = (int)(j % d);
    if (uinv[j] < 0 || uinv[j] >= u[i]) {
      delete pl;
      return fail(GRIEF_ERR_BAD_ARG, "plan_create: uinv[%lld]=%d outside [0,%d)", (long long)j, uinv[j], u[i]);
    }
  }

  // ---- distinct-sub-tuple counts of every contiguous dimension range [a,b) (hashed) ----
  const int INF = 1 << 29;
  std::vector<int> cost((size_t)(d + 1) * (d + 1), INF);
  {
    std::vector<uint64_t> h(p), tmp(p);
    for (int a = 0; a < d; ++a) {
      std::fill(h.begin(), h.end(), 0x1234567ull + a);
      for (int b = a + 1; b <= d; ++b) {
        for (int j = 0; j < p; ++j) h[j] = mix64(h[j], (uint64_t)uinv[(size_t)j * d + (b - 1)] + 1);
        tmp = h;
        std::sort(tmp.begin(), tmp.end());
        int cnt = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
        cost[(size_t)a * (d + 1) + b] = cnt;
        if (cnt + 2 > width_cap) break;  // monotone in b: longer ranges cannot fit either
      }
    }
  }
  // ---- smallest number of groups whose tables fit the cap; among those the narrowest ----
  int bestG = -1;
  std::vector<int> cuts;
  {
    const int GM = std::min(kMaxGroups, d);
    std::vector<std::vector<int>> best(GM + 1, std::vector<int>(d + 1, INF)), arg(GM + 1, std::vector<int>(d + 1, -1));
    best[0][0] = 0;
    for (int g = 1; g <= GM && bestG < 0; ++g) {
      for (int b = 1; b <= d; ++b)
        for (int a = g - 1; a < b; ++a) {
          if (best[g - 1][a] >= INF || cost[(size_t)a * (d + 1) + b] >= INF) continue;
          int c = best[g - 1][a] + cost[(size_t)a * (d + 1) + b];
          if (c < best[g][b]) { best[g][b] = c; arg[g][b] = a; }
        }
      if (best[g][d] + 2 <= width_cap) {
        bestG = g;
        cuts.assign(g + 1, 0);
        int b = d;
        for (int gg = g; gg >= 1; --gg) { cuts[gg] = b; b = arg[gg][b]; }
        cuts[0] = 0;
      }
    }
  }
  if (bestG < 0) {
    delete pl;
    return fail(GRIEF_ERR_UNSUPPORTED,
                "plan_create: no partition of the %d dimensions into <=%d groups fits a table of %d entries per row",
                d, kMaxGroups, width_cap);
  }
  const int G = bestG;
  pl->n_groups = G;
  pl->group_begin = cuts;
  pl->group_slot0.assign(G, 0);
  pl->group_size.assign(G, 0);
  pl->max_group_dims = 1;
  for (int g = 0; g < G; ++g) pl->max_group_dims = std::max(pl->max_group_dims, cuts[g + 1] - cuts[g]);

  // ---- exact enumeration of the distinct sub-tuples of each chosen group ----
  pl->col_slot_h.assign((size_t)pl->p_pad * G, 0);
  std::vector<uint8_t> slot_k;   // filled below (width x max_group_dims)
  std::vector<int> slot_group;
  int slot = 2;                  // slot 0 = constant 1.0, slot 1 = constant 0.0
  std::vector<std::vector<uint8_t>> slot_rows;
  slot_rows.push_back(std::vector<uint8_t>(pl->max_group_dims, 0));
  slot_rows.push_back(std::vector<uint8_t>(pl->max_group_dims, 0));
  slot_group.push_back(-1);
  slot_group.push_back(-1);
  std::vector<int> order(p);
  for (int g = 0; g < G; ++g) {
    const int a = cuts[g], b = cuts[g + 1];
    std::iota(order.begin(), order.end(), 0);
    auto less = [&](int x, int y) {
      const int32_t* px = uinv + (size_t)x * d;
      const int32_t* py = uinv + (size_t)y * d;
      for (int i = a; i < b; ++i)
        if (px[i] != py[i]) return px[i] < py[i];
      return false;
    };
    std::stable_sort(order.begin(), order.end(), less);
    pl->group_slot0[g] = slot;
    for (int r = 0; r < p; ++r) {
      const int j = order[r];
      if (r == 0 || less(order[r - 1], j)) {
        std::vector<uint8_t> row(pl->max_group_dims, 0);
        for (int i = a; i < b; ++i) row[i - a] = (uint8_t)uinv[(size_t)j * d + i];
        slot_rows.push_back(row);
        slot_group.push_back(g);
        ++slot;
      }
      pl->col_slot_h[(size_t)j * G + g] = (uint16_t)(slot - 1);
    }
    pl->group_size[g] = slot - pl->group_slot0[g];
  }
  pl->width = slot;
  pl->stride = (slot % 2 == 1) ? slot : slot + 1;  // odd stride: conflict-free lane<->row gathers
  // padding columns j >= p: zero slot for group 0, one slot for the rest  => Phi[:, j] = 0
  for (int j = p; j < pl->p_pad; ++j) {
    pl->col_slot_h[(size_t)j * G + 0] = 1;
    for (int g = 1; g < G; ++g) pl->col_slot_h[(size_t)j * G + g] = 0;
  }
  slot_k.resize((size_t)pl->width * pl->max_group_dims);
  for (int s = 0; s < pl->width; ++s)
    std::memcpy(&slot_k[(size_t)s * pl->max_group_dims], slot_rows[s].data(), pl->max_group_dims);

  // ---- sorted column order for the tile builders ----
  std::vector<uint16_t> sorted_slot((size_t)pl->p_pad * G, 0);
  std::vector<uint8_t> sorted_level(pl->p_pad, 0);
  pl->perm_h.assign(pl->p_pad, -1);
  {
    std::vector<int> key_order(G);                       // key position -> group (fewest distinct sub-tuples first)
    std::iota(key_order.begin(), key_order.end(), 0);
    std::stable_sort(key_order.begin(), key_order.end(), [&](int a, int b) { return pl->group_size[a] < pl->group_size[b]; });
    std::vector<int> cols(p);
    std::iota(cols.begin(), cols.end(), 0);
    auto key = [&](int j, int kpos) { return pl->col_slot_h[(size_t)j * G + key_order[kpos]]; };
    std::stable_sort(cols.begin(), cols.end(), [&](int x, int y) {
      for (int k = 0; k < G; ++k)
        if (key(x, k) != key(y, k)) return key(x, k) < key(y, k);
      return false;
    });
    for (int c = 0; c < pl->p_pad; ++c) {
      const int j = c < p ? cols[c] : c;                 // padding columns keep their (all-constant) slots
      if (c < p) pl->perm_h[c] = j;
      for (int k = 0; k < G; ++k) sorted_slot[(size_t)c * G + k] = pl->col_slot_h[(size_t)j * G + key_order[k]];
      int lv = 0;
      if (c > 0) {
        lv = G;
        for (int k = 0; k < G; ++k)
          if (sorted_slot[(size_t)c * G + k] != sorted_slot[(size_t)(c - 1) * G + k]) { lv = k; break; }
      }
      sorted_level[c] = (uint8_t)lv;
    }
  }
  // packed form for the lane = row kernels: word w of sorted column c holds the slots of key positions 4w .. 4w+3, most significant
  // byte first (a slot index is < 256: kTableCap); the first differing key position of two consecutive columns is then a
  // count-leading-zeros of the XOR of their words
  static_assert(kTableCap + 20 < 256, "slot indices must fit one byte");
  pl->pack_words = (G + 3) / 4;
  std::vector<uint32_t> sorted_pack((size_t)pl->p_pad * pl->pack_words, 0);
  for (int c = 0; c < pl->p_pad; ++c)
    for (int k = 0; k < G; ++k)
      sorted_pack[(size_t)c * pl->pack_words + k / 4] |= (uint32_t)(sorted_slot[(size_t)c * G + k] & 0xFF) << (8 * (3 - k % 4));

  // ---- launch order of the Gram's lower-triangle tiles when CTA pairs (256 x 128) compute them ----
  // (pair row pm = tile rows 2 pm, 2 pm + 1; column tile bn <= 2 pm + 1).  One K split of the digit planes (p_pad columns x ~10^4 rows
  // x digits) is larger than L2, so what matters is how many distinct panels the ~74 pair tiles in flight touch: bands of 8 pair rows
  // are walked in blocks of 9 column tiles (<= 72 pair tiles, 16 + 9 panels) instead of whole columns (32 + ~5 panels each time).
  std::vector<uint32_t> gram_order;
  {
    const int tiles = pl->p_pad / kTileN, prow = (tiles + 1) / 2;
    constexpr int R = 8, CW = 9;
    for (int b0 = 0; b0 < prow; b0 += R) {
      const int b1 = std::min(prow, b0 + R);
      const int bn_max = std::min(tiles - 1, 2 * (b1 - 1) + 1);
      for (int c0 = 0; c0 <= bn_max; c0 += CW)
        for (int pm = b0; pm < b1; ++pm)
          for (int bn = c0; bn < std::min(c0 + CW, bn_max + 1); ++bn)
            if (bn <= std::min(tiles - 1, 2 * pm + 1)) gram_order.push_back(((uint32_t)pm << 16) | (uint32_t)bn);
    }
    pl->n_gram_order = (int)gram_order.size();
  }

  // ---- upload ----
  std::vector<double> grid(grid_concat, grid_concat + pl->sum_m);
  std::vector<double> qs(qs_concat, qs_concat + qoff);
  PoolBuilder pool;
  pool.add(&pl->d_dims, pl->dims);
  pool.add(&pl->d_grid, grid);
  pool.add(&pl->d_qs, qs);
  pool.add(&pl->d_slot_k, slot_k);
  pool.add(&pl->d_slot_group, slot_group);
  pool.add(&pl->d_group_begin, pl->group_begin);
  pool.add(&pl->d_col_slot, pl->col_slot_h);
  pool.add(&pl->d_sorted_slot, sorted_slot);
  pool.add(&pl->d_sorted_level, sorted_level);
  pool.add(&pl->d_sorted_pack, sorted_pack);
  pool.add(&pl->d_gram_order, gram_order);
  pool.add(&pl->d_perm, pl->perm_h);
  pool.add(&pl->d_err, std::vector<int>(1, 0));
  int rc = pool.commit(&pl->d_pool);
  if (rc != GRIEF_OK) {
    delete pl;
    return rc;
  }
  *out = pl;
  return GRIEF_OK;
}

}  // namespace grief
