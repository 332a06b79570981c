// den_pack.cu -- host-side packer of a shared denominator graph (see den_blob.h).
//
// Input: the reference's CSR tensors for ONE graph (produced by
// /root/reference/openfst_binding/src/fstext.cc:19-117 or by pychain_b200.graph.Fst).
// Output: the blob consumed by den_shared.cu.  Pure host C++ (no CUDA calls).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "den_blob.h"

namespace chainb200 {
namespace {

struct Arc {
  uint32_t ab;  // a | b << 16
  float w;
};

struct Unit {            // a row, or one piece of a split row
  int32_t row_code;      // >= 0 row id; <= -2 sub-row slot
  std::vector<Arc> arcs;
};

struct ViewBuild {
  int n_rows = 0, n_slices = 0, n_split = 0, n_sub = 0;
  std::vector<int32_t> rows, width, split;  // rows: per slice while building, re-emitted per warp below
  std::vector<int64_t> warp_info;            // [DEN_NW][3]
  std::vector<uint32_t> rec_ab;
  std::vector<float> rec_w;
  double predicted = 0, ideal = 0;
};

// Bank geometry of the gathers for a group width G: a vector load of G floats per lane is served in
// phases of P = 32/G lanes (G=4: quarter-warps of 8 lanes and 16-byte elements, G=2: half-warps of 16
// lanes and 8-byte elements, G=1: the whole warp and 4-byte elements), and within a phase element e
// lives in bank group e mod P.  Wavefronts of one phase = the largest number of DISTINCT elements
// that fall into the same bank group (equal addresses are a broadcast).
static int phase_wavefronts(const uint32_t *idx, int n, int P) {
  int worst = 1;
  const uint32_t mask = (uint32_t)P - 1u;
  for (int i = 0; i < n; i++) {
    int cnt = 1;
    bool first = true;
    for (int j = 0; j < n; j++) {
      if (j == i || (idx[j] & mask) != (idx[i] & mask)) continue;
      if (idx[j] == idx[i]) { if (j < i) first = false; continue; }   // broadcast partner
      // count distinct elements only once: skip j if an equal element appears earlier
      bool dup = false;
      for (int l = 0; l < j; l++) dup |= (idx[l] == idx[j]);
      if (!dup) cnt++;
    }
    if (first) worst = std::max(worst, cnt);
  }
  return worst;
}

// ---- deterministic RNG for the packer's local search -------------------------------------
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
  uint32_t next() {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (uint32_t)(s >> 32);
  }
  uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
  double unit() { return next() * (1.0 / 4294967296.0); }
};

// Search effort of the packer: 0 = plain degree-sorted SELL, >= 1 = conflict-aware.  The graph is
// packed once per training run, so the default spends ~1 s on a 40k-arc graph and scales back
// for very large graphs.
static int pack_effort(int K) {
  const char *e = getenv("CHAIN_B200_PACK_EFFORT");
  if (e) return std::max(0, atoi(e));
  return K <= 120000 ? 4 : (K <= 400000 ? 2 : 1);
}

// A. Regroup the (length-sorted) units so that every run of P consecutive units -- the rows one
// phase of a vector load processes side by side -- has flat histograms of bank groups (index mod P)
// for both gathered operands and distinct bank groups for its output rows.  Flat histograms are what
// makes a conflict-free column order exist.
static void balance_groups(const std::vector<Unit> &units, std::vector<int> &order, int P) {
  const int n = (int)order.size();
  const uint32_t mask = (uint32_t)P - 1u;
  std::vector<std::array<uint16_t, 64>> hist(units.size());
  for (size_t u = 0; u < units.size(); u++) {
    hist[u].fill(0);
    for (const Arc &a : units[u].arcs) { hist[u][a.ab & mask]++; hist[u][32 + ((a.ab >> 16) & mask)]++; }
  }
  std::vector<int> out;
  out.reserve(n);
  std::vector<char> used(n, 0);
  int head = 0;
  const int WIN = 160;
  while ((int)out.size() < n) {
    while (head < n && used[head]) head++;
    const int first = order[head];
    used[head] = 1;
    out.push_back(first);
    long sum[64];
    for (int c = 0; c < 64; c++) sum[c] = hist[first][c];
    const size_t len0 = units[first].arcs.size();
    // bank groups already taken by the output rows of this group (the epilogue stores out[row][g])
    auto out_group = [&](int u) { const int c = units[u].row_code; return (uint32_t)(c >= 0 ? (c & 0xfffff) : -2 - c) & mask; };
    unsigned out_used = 1u << out_group(first);
    for (int pick = 1; pick < P && (int)out.size() < n; pick++) {
      int best = -1;
      long best_cost = 0;
      int seen = 0;
      for (int j = head + 1; j < n && seen < WIN; j++) {
        if (used[j]) continue;
        const int u = order[j];
        // keep slices tight: prefer equal length; allow one shorter once the equal ones run out
        if (units[u].arcs.size() + 1 < len0 && seen >= 8) break;
        seen++;
        long cost = 0;
        for (int c = 0; c < 64; c++) { const long t = sum[c] + hist[u][c]; cost += t * t; }
        cost = cost * 4 + (long)(len0 - units[u].arcs.size()) * 64;
        if ((out_used >> out_group(u)) & 1u) cost += 24;
        if (best < 0 || cost < best_cost) { best = j; best_cost = cost; }
      }
      if (best < 0) break;
      used[best] = 1;
      out.push_back(order[best]);
      out_used |= 1u << out_group(order[best]);
      for (int c = 0; c < 64; c++) sum[c] += hist[order[best]][c];
    }
  }
  order.swap(out);
}

// B. Column order inside one phase group (P lanes x W columns): simulated annealing over swaps of
// two entries of one lane, minimising the modelled wavefronts sum_k (max multiplicity of the
// a-bank-groups + max multiplicity of the b-bank-groups), ties broken towards flat counts.
static void anneal_group(const Unit *const *lanes, int P, int W, int *arr /* [P][W] arc index or -1 */, Rng &rng,
                         int effort) {
  const uint32_t mask = (uint32_t)P - 1u;
  std::vector<uint8_t> c1((size_t)W * P, 0), c2((size_t)W * P, 0);
  auto col_of = [&](int i, int k, int &x, int &y) {
    const int j = arr[i * W + k];
    if (j < 0) { x = y = -1; return; }
    const uint32_t ab = lanes[i]->arcs[j].ab;
    x = ab & mask; y = (ab >> 16) & mask;
  };
  for (int i = 0; i < P; i++)
    for (int k = 0; k < W; k++) {
      int x, y;
      col_of(i, k, x, y);
      if (x >= 0) { c1[k * P + x]++; c2[k * P + y]++; }
    }
  auto col_cost = [&](int k) {
    int m1 = 0, m2 = 0, sq = 0;
    for (int c = 0; c < P; c++) {
      const int u = c1[k * P + c], v = c2[k * P + c];
      m1 = std::max(m1, u); m2 = std::max(m2, v);
      sq += u * u + v * v;
    }
    return 24 * (m1 + m2) + sq;
  };
  std::vector<int> cc(W);
  long total = 0;
  for (int k = 0; k < W; k++) { cc[k] = col_cost(k); total += cc[k]; }
  std::vector<int> best(arr, arr + P * W);
  long best_total = total;
  const int moves = effort * 28 * P * W + 400;
  double T = 10.0;
  const double decay = pow(0.05 / T, 1.0 / moves);
  for (int it = 0; it < moves; it++, T *= decay) {
    const int i = rng.below(P), k1 = rng.below(W);
    int k2 = rng.below(W - 1);
    if (k2 >= k1) k2++;
    const int j1 = arr[i * W + k1], j2 = arr[i * W + k2];
    if (j1 < 0 && j2 < 0) continue;
    int x1, y1, x2, y2;
    col_of(i, k1, x1, y1);
    col_of(i, k2, x2, y2);
    auto mv = [&](int k, int x, int y, int d) {
      if (x >= 0) { c1[k * P + x] += d; c2[k * P + y] += d; }
    };
    mv(k1, x1, y1, -1); mv(k2, x2, y2, -1); mv(k1, x2, y2, +1); mv(k2, x1, y1, +1);
    const int n1 = col_cost(k1), n2 = col_cost(k2);
    const int delta = n1 + n2 - cc[k1] - cc[k2];
    if (delta <= 0 || rng.unit() < exp(-delta / T)) {
      arr[i * W + k1] = j2; arr[i * W + k2] = j1;
      cc[k1] = n1; cc[k2] = n2;
      total += delta;
      if (total < best_total) { best_total = total; std::copy(arr, arr + P * W, best.begin()); }
    } else {
      mv(k1, x2, y2, -1); mv(k2, x1, y1, -1); mv(k1, x1, y1, +1); mv(k2, x2, y2, +1);
    }
  }
  std::copy(best.begin(), best.end(), arr);
}

// Order the arcs of the 32 units of a slice into columns.  `cols[k*32+lane]` = index into
// lanes[lane]->arcs or -1 (idle).
static void arrange_slice(const std::vector<const Unit *> &lanes, int width, std::vector<int> &cols, Rng &rng,
                          int effort, int P) {
  cols.assign((size_t)width * 32, -1);
  for (int l = 0; l < 32; l++) {
    if (!lanes[l]) continue;
    const int n = (int)lanes[l]->arcs.size();
    for (int k = 0; k < n; k++) cols[(size_t)k * 32 + l] = k;
  }
  if (effort <= 0 || width < 2) return;
  static const Unit empty_unit{-1, {}};
  std::vector<int> arr((size_t)P * width);
  for (int q = 0; q < 32 / P; q++) {
    const Unit *ql[32];
    bool any = false;
    for (int i = 0; i < P; i++) {
      ql[i] = lanes[q * P + i] ? lanes[q * P + i] : &empty_unit;
      any |= !ql[i]->arcs.empty();
      for (int k = 0; k < width; k++) arr[(size_t)i * width + k] = cols[(size_t)k * 32 + q * P + i];
    }
    if (!any) continue;
    anneal_group(ql, P, width, arr.data(), rng, effort);
    for (int i = 0; i < P; i++)
      for (int k = 0; k < width; k++) cols[(size_t)k * 32 + q * P + i] = arr[(size_t)i * width + k];
  }
}

// The arcs of one view: arc k belongs to row row_of[k] and multiplies A[ab & 0xffff] (and B[ab >> 16]) by w.
struct ArcSet {
  std::vector<int32_t> row_of;
  std::vector<uint32_t> ab;
  std::vector<float> w;
  void push(int32_t row, uint32_t ab_, float w_) { row_of.push_back(row); ab.push_back(ab_); w.push_back(w_); }
  int size() const { return (int)row_of.size(); }
};

// skip_empty: rows without arcs get no unit (views whose epilogue need not run for every row);
// row_flags: optional per-row bits OR-ed into the row code (kRowExtraFlag).
static void build_view(int n_rows, const ArcSet &arcs, int n_a, int n_b, bool structure_only, bool log_domain, int P,
                       bool skip_empty, const std::vector<int32_t> *row_flags, int effort_K, bool one_operand,
                       const float *row_aux, bool embed, ViewBuild &v) {
  v.n_rows = n_rows;
  const int K = arcs.size();
  const std::vector<int32_t> &row_of = arcs.row_of;
  const std::vector<uint32_t> &ab = arcs.ab;
  const float *w = arcs.w.data();
  const int effort = pack_effort(effort_K);
  Rng rng(0xC0FFEEull + (uint64_t)n_rows * 7919u + (uint64_t)K);
  // rows -> arc lists, arc order inside a row = CSR order (the reference's summation order)
  std::vector<int32_t> deg(n_rows, 0);
  for (int k = 0; k < K; k++) deg[row_of[k]]++;
  std::vector<std::vector<Arc>> by_row(n_rows);
  for (int r = 0; r < n_rows; r++) by_row[r].reserve(deg[r]);
  for (int k = 0; k < K; k++) by_row[row_of[k]].push_back(Arc{ab[k], w[k]});
  auto code_of = [&](int r) { return row_flags ? (r | (*row_flags)[r]) : r; };

  // split heavy rows so that no lane carries more than `cap` arcs
  int rows_used = n_rows;
  if (skip_empty) { rows_used = 0; for (int r = 0; r < n_rows; r++) rows_used += deg[r] > 0; }
  const double mean = (double)K / std::max(1, rows_used);
  int cap = std::max(32, 4 * (int)(mean + 0.999));
  cap += cap & 1;
  std::vector<Unit> units;
  units.reserve(n_rows + 64);
  v.split.clear();
  int n_sub = 0;
  for (int r = 0; r < n_rows; r++) {
    const int d = deg[r];
    if (d == 0 && skip_empty) continue;
    if (d <= cap) {
      units.push_back(Unit{code_of(r), std::move(by_row[r])});
      continue;
    }
    const int pieces = (d + cap - 1) / cap;
    v.split.push_back(code_of(r)); v.split.push_back(n_sub); v.split.push_back(pieces);
    int pos = 0;
    for (int pc = 0; pc < pieces; pc++) {
      const int len = d / pieces + (pc < d % pieces ? 1 : 0);
      Unit u;
      u.row_code = -2 - (n_sub++);
      u.arcs.assign(by_row[r].begin() + pos, by_row[r].begin() + pos + len);
      pos += len;
      units.push_back(std::move(u));
    }
  }
  v.n_split = (int)v.split.size() / 3;
  v.n_sub = n_sub;

  // longest first, 32 units per slice
  std::vector<int> order(units.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(),
                   [&](int x, int y) { return units[x].arcs.size() > units[y].arcs.size(); });
  if (effort >= 1) balance_groups(units, order, P);
  const int n_slices = ((int)units.size() + 31) / 32;   // 0 for an empty view: every warp then runs 0 pairs
  v.n_slices = n_slices;
  v.rows.assign((size_t)n_slices * 32, -1);
  v.width.assign(n_slices, 0);
  std::vector<long> slice_cost(n_slices, 0);  // modelled shared-memory wavefronts of each slice
  std::vector<std::vector<uint32_t>> s_ab(n_slices);
  std::vector<std::vector<float>> s_w(n_slices);
  const uint32_t na = (uint32_t)n_a, nb = (uint32_t)n_b;
  for (int s = 0; s < n_slices; s++) {
    std::vector<const Unit *> lanes(32, nullptr);
    int width = 0;
    for (int l = 0; l < 32; l++) {
      const size_t i = (size_t)s * 32 + l;
      if (i >= units.size()) break;
      lanes[l] = &units[order[i]];
      v.rows[i] = lanes[l]->row_code;
      width = std::max(width, (int)lanes[l]->arcs.size());
    }
    // embed: the LAST slot of every lane is not an arc but the lane's row code (see den_blob.h): width = arcs + 1,
    // rounded up to even (slots are consumed in pairs).  Otherwise width = arcs rounded up to even, >= 2.
    const int ncode = embed ? 1 : 0;
    width = embed ? ((width + 2) & ~1) : std::max(2, width + (width & 1));
    std::vector<int> cols;
    arrange_slice(lanes, width - ncode, cols, rng, structure_only ? 0 : effort, P);
    v.width[s] = width;
    s_ab[s].assign((size_t)width * 32, 0u);
    s_w[s].assign((size_t)width * 32, log_domain ? -INFINITY : 0.f);
    for (int l = 0; embed && l < 32; l++) {
      const size_t i = (size_t)s * 32 + l;
      const int32_t code = (i < units.size()) ? lanes[l]->row_code : -1;
      float aux = (row_aux && code >= 0) ? fabsf(row_aux[code & 0xfffff]) : 0.f;
      uint32_t mark;
      if (log_domain) {
        mark = kCodeMarkLog;
      } else {
        aux = -aux;                                   // sign bit = "this slot is the row code" (arc weights are >= 0)
        if (!std::signbit(aux)) aux = -0.0f;
        memcpy(&mark, &aux, 4);
      }
      s_ab[s][(size_t)(width - 1) * 32 + l] = (uint32_t)code;
      memcpy(&s_w[s][(size_t)(width - 1) * 32 + l], &mark, 4);
    }
    for (int k = 0; k < width - ncode; k++) {
      uint32_t ia[32], ib[32];
      for (int l = 0; l < 32; l++) {
        const int j = cols[(size_t)k * 32 + l];
        uint32_t rec;
        float wt = log_domain ? -INFINITY : 0.f;
        if (j >= 0) {
          rec = lanes[l]->arcs[j].ab;
          wt = lanes[l]->arcs[j].w;
        } else {  // idle lane: weight 0, placed below
          rec = 0xffffffffu;
        }
        s_ab[s][(size_t)k * 32 + l] = rec;
        s_w[s][(size_t)k * 32 + l] = wt;
        ia[l] = rec & 0xffffu;
        ib[l] = rec >> 16;
      }
      for (int q = 0; q < 32 / P; q++) {
        // idle lanes read element `g` of a bank group g nobody else uses in this phase of the column
        const uint32_t mask = (uint32_t)P - 1u;
        unsigned used_a = 0, used_b = 0;
        for (int i = 0; i < P; i++)
          if (s_ab[s][(size_t)k * 32 + q * P + i] != 0xffffffffu) { used_a |= 1u << (ia[q * P + i] & mask); used_b |= 1u << (ib[q * P + i] & mask); }
        for (int i = 0; i < P; i++) {
          const size_t pos = (size_t)k * 32 + q * P + i;
          if (s_ab[s][pos] != 0xffffffffu) continue;
          uint32_t ga = 0, gb = 0;
          while (ga < mask && ((used_a >> ga) & 1u)) ga++;
          while (gb < mask && ((used_b >> gb) & 1u)) gb++;
          used_a |= 1u << ga; used_b |= 1u << gb;
          if (ga >= na) ga = 0;
          if (gb >= nb) gb = 0;
          s_ab[s][pos] = ga | (gb << 16);
          ia[q * P + i] = ga; ib[q * P + i] = gb;
        }
        const int wf = phase_wavefronts(ia + P * q, P, P) + (one_operand ? 0 : phase_wavefronts(ib + P * q, P, P));
        v.predicted += wf;
        slice_cost[s] += wf;
        v.ideal += one_operand ? 1 : 2;
      }
    }
  }

  // slices -> warps: longest-processing-time first on the modelled cost of each slice (its
  // shared-memory wavefronts + a per-slice epilogue constant), so the warps of a CTA reach the
  // barrier that ends a pass together
  std::vector<int> by_cost(n_slices);
  std::iota(by_cost.begin(), by_cost.end(), 0);
  auto cost_of = [&](int s) { return structure_only ? (long)v.width[s] * 10 + 12 : slice_cost[s] + 12; };
  std::stable_sort(by_cost.begin(), by_cost.end(), [&](int x, int y) { return cost_of(x) > cost_of(y); });
  std::vector<long> load(DEN_NW, 0);
  std::vector<std::vector<int32_t>> per(DEN_NW);
  for (int s : by_cost) {
    int best = 0;
    for (int wv = 1; wv < DEN_NW; wv++)
      if (load[wv] < load[best]) best = wv;
    per[best].push_back(s);
    load[best] += cost_of(s);
  }
  v.warp_info.assign((size_t)DEN_NW * 3, 0);
  v.rec_ab.clear();
  v.rec_w.clear();
  std::vector<int32_t> rows_seq;
  for (int wv = 0; wv < DEN_NW; wv++) {
    const size_t rec0 = v.rec_ab.size();
    v.warp_info[3 * wv + 0] = (int64_t)rec0;
    v.warp_info[3 * wv + 2] = embed ? (int64_t)per[wv].size() : (int64_t)rows_seq.size();  // slices / row-code offset
    for (int s : per[wv]) {
      if (!embed) rows_seq.insert(rows_seq.end(), v.rows.begin() + (size_t)s * 32, v.rows.begin() + (size_t)s * 32 + 32);
      // records are consumed in pairs: the two records of a lane's pair are ADJACENT (16 bytes, one 128-bit
      // load per lane and pair, 512 contiguous bytes per warp): [pair][lane][2]
      const int width = v.width[s];
      for (int j = 0; j < width / 2; j++)
        for (int l = 0; l < 32; l++)
          for (int h = 0; h < 2; h++) {
            v.rec_ab.push_back(s_ab[s][(size_t)(2 * j + h) * 32 + l]);
            v.rec_w.push_back(s_w[s][(size_t)(2 * j + h) * 32 + l]);
          }
      if (!embed) {  // end-of-slice flag on the second slot of every lane's last pair
        const size_t last = v.rec_w.size() - 64;
        for (int l = 0; l < 32; l++) {
          const size_t q = last + 2 * l + 1;
          if (log_domain) {
            v.rec_ab[q] |= 0x80000000u;
          } else {
            v.rec_w[q] = -fabsf(v.rec_w[q]);
            if (!std::signbit(v.rec_w[q])) v.rec_w[q] = -0.0f;
          }
        }
      }
    }
    if (!embed) rows_seq.insert(rows_seq.end(), 3 * 32, -1);  // the row prefetch runs slices ahead
    size_t pairs = (v.rec_ab.size() - rec0) / 64;
    const size_t run = pairs;  // the kernel runs the last (partial) double bank through a predicated tail
    v.warp_info[3 * wv + 1] = (int64_t)run;
    v.rec_ab.insert(v.rec_ab.end(), (size_t)3 * kFifoPairs * 64, 0u);  // register-bank run-out
    v.rec_w.insert(v.rec_w.end(), (size_t)3 * kFifoPairs * 64, log_domain ? -INFINITY : 0.f);
  }
  if (embed) v.rows.clear();   // the row codes travel inside the record streams
  else v.rows.swap(rows_seq);
}

}  // namespace

int64_t pack_den_blob(const int32_t *ft, const int32_t *fi, const float *fp, const int32_t *bt, const int32_t *bi,
                      const float *bp, const float *leaky, const float *init, const float *fin, int S, int K, int D,
                      bool log_domain, int group_width, void *out, int64_t out_bytes) {
  const int P = (group_width == 1) ? 32 : (group_width == 2 ? 16 : 8);  // lanes per phase of the vector gathers
  if (S <= 0 || K <= 0 || D <= 0 || !ft || !fi || !bt || !bi) return CHAIN_B200_ERR_INVALID_ARG;
  if (S > 65535 || D > 65535) return CHAIN_B200_ERR_UNSUPPORTED;
  if (log_domain && (S > 32767 || D > 32767)) return CHAIN_B200_ERR_UNSUPPORTED;  // bit 31 of the index word is the flag
  // validate the CSR: contiguous coverage, indices in range, rows grouped as claimed
  int kf = 0, kb = 0;
  for (int s = 0; s < S; s++) {
    if (fi[2 * s] != kf || fi[2 * s + 1] < kf || fi[2 * s + 1] > K) return CHAIN_B200_ERR_INVALID_ARG;
    kf = fi[2 * s + 1];
    if (bi[2 * s] != kb || bi[2 * s + 1] < kb || bi[2 * s + 1] > K) return CHAIN_B200_ERR_INVALID_ARG;
    kb = bi[2 * s + 1];
  }
  if (kf != kb || kf <= 0) return CHAIN_B200_ERR_INVALID_ARG;
  const int Kv = kf;  // real arcs (entries beyond are batch padding)
  for (int s = 0; s < S; s++) {
    for (int k = bi[2 * s]; k < bi[2 * s + 1]; k++)
      if (bt[3 * k + 1] != s) return CHAIN_B200_ERR_INVALID_ARG;
    for (int k = fi[2 * s]; k < fi[2 * s + 1]; k++)
      if (ft[3 * k] != s) return CHAIN_B200_ERR_INVALID_ARG;
  }
  if (fp && bp && !log_domain)
    for (int k = 0; k < Kv; k++)
      if (!(fp[k] >= 0.f) || !(bp[k] >= 0.f)) return CHAIN_B200_ERR_INVALID_ARG;  // probabilities (sign bit is a flag)
  for (int k = 0; k < Kv; k++) {
    const int32_t bs = bt[3 * k], bd = bt[3 * k + 1], bn = bt[3 * k + 2];
    const int32_t fs = ft[3 * k], fd = ft[3 * k + 1], fn = ft[3 * k + 2];
    if (bs < 0 || bs >= S || bd < 0 || bd >= S || bn < 0 || bn >= D) return CHAIN_B200_ERR_INVALID_ARG;
    if (fs < 0 || fs >= S || fd < 0 || fd >= S || fn < 0 || fn >= D) return CHAIN_B200_ERR_INVALID_ARG;
  }
  const int S_pad = (S + 3) / 4 * 4, D_pad = (D + 3) / 4 * 4;
  const bool so = (out == nullptr);
  bool embed = false;   // stream format (den_blob.h), decided below from how thin the slices of views 0 and 1 are
  bool tried_resident = false, no_resident = log_domain;   // (the log-domain kernel streams; it has no resident mode)

  // ---- (dst, pdf) classes: the arcs of one destination that share a pdf (see den_blob.h) ----------------
  // candidates in backward-CSR order (arcs grouped by destination), members keep their CSR order
  struct Cls { int32_t dst, pdf; std::vector<int32_t> arcs; };
  std::vector<Cls> cls;
  ViewBuild vb[kNumViews];
  int C = 0, C_pad = 0, R_pad = S_pad;
  const int G = (group_width == 1 || group_width == 2) ? group_width : 4;
  const long chip_floats = (227L * 1024 - 1024) / 4 / G;   // per utterance of a group (den_smem_bytes + static)
  long c_cap = 65532L - S_pad;
  for (int attempt = 0;; attempt++) {
  cls.clear();
  std::vector<int32_t> cls_of_bwd(Kv, -1);
  {
    int min_class = 4;
    if (const char *e = getenv("CHAIN_B200_MIN_CLASS")) min_class = std::max(2, atoi(e));
    // room on chip next to 2 state rows + 3 frame rows (+ an allowance for split-row slots) of G utterances
    long budget = chip_floats - 2L * S_pad - 3L * D_pad - 64;
    long c_max = std::max(0L, std::min(budget / 2, c_cap)) / 4 * 4;
    if (const char *e = getenv("CHAIN_B200_MAX_CLASSES")) c_max = std::min(c_max, (long)std::max(0, atoi(e)) / 4 * 4);
    if (log_domain) c_max = 0;
    std::vector<Cls> cand;
    if (c_max > 0) {
      std::vector<int32_t> idx;
      for (int s = 0; s < S; s++) {
        idx.clear();
        for (int k = bi[2 * s]; k < bi[2 * s + 1]; k++) idx.push_back(k);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return bt[3 * x + 2] < bt[3 * y + 2]; });
        for (size_t i = 0; i < idx.size();) {
          size_t j = i;
          while (j < idx.size() && bt[3 * idx[j] + 2] == bt[3 * idx[i] + 2]) j++;
          if ((int)(j - i) >= min_class) {
            Cls c{s, bt[3 * idx[i] + 2], {}};
            c.arcs.assign(idx.begin() + i, idx.begin() + j);
            cand.push_back(std::move(c));
          }
          i = j;
        }
      }
      std::stable_sort(cand.begin(), cand.end(), [](const Cls &x, const Cls &y) { return x.arcs.size() > y.arcs.size(); });
      if ((long)cand.size() > c_max) cand.resize((size_t)c_max);
    }
    cls.swap(cand);
    for (size_t c = 0; c < cls.size(); c++)
      for (int k : cls[c].arcs) cls_of_bwd[k] = (int32_t)c;
  }
  C = (int)cls.size(); C_pad = (C + 3) / 4 * 4; R_pad = S_pad + C_pad;
  // class of a forward-CSR arc: by (dst, pdf)
  std::vector<int32_t> cls_of_fwd(Kv, -1);
  if (C > 0) {
    std::vector<std::pair<uint64_t, int32_t>> keys(C);
    for (int c = 0; c < C; c++) keys[c] = {((uint64_t)cls[c].dst << 32) | (uint32_t)cls[c].pdf, c};
    std::sort(keys.begin(), keys.end());
    for (int k = 0; k < Kv; k++) {
      const uint64_t key = ((uint64_t)ft[3 * k + 1] << 32) | (uint32_t)ft[3 * k + 2];
      auto it = std::lower_bound(keys.begin(), keys.end(), std::make_pair(key, (int32_t)-1));
      if (it != keys.end() && it->first == key) cls_of_fwd[k] = it->second;
    }
  }
  auto W = [&](const float *p, int k) { return p ? p[k] : 0.f; };
  const float one = log_domain ? 0.f : 1.f;
  ArcSet a0, a1, a2, a3, a4;
  std::vector<int32_t> flags1(S, 0);
  for (int k = 0; k < Kv; k++) {                       // backward-CSR order: rows = destinations
    const int32_t src = bt[3 * k], dst = bt[3 * k + 1], pdf = bt[3 * k + 2], c = cls_of_bwd[k];
    if (c < 0) {
      a0.push(dst, (uint32_t)src | ((uint32_t)pdf << 16), W(bp, k));
    } else {
      if (cls[c].arcs.front() == k) a0.push(dst, (uint32_t)(S_pad + c) | ((uint32_t)pdf << 16), one);  // the class as ONE arc
      a3.push(c, (uint32_t)src, W(bp, k));             // its members: one operand each
    }
  }
  for (int k = 0; k < Kv; k++) {                       // forward-CSR order: rows = sources / pdfs
    const int32_t src = ft[3 * k], dst = ft[3 * k + 1], pdf = ft[3 * k + 2], c = cls_of_fwd[k];
    if (c < 0) {
      a1.push(src, (uint32_t)dst | ((uint32_t)pdf << 16), W(fp, k));
      a2.push(pdf, (uint32_t)src | ((uint32_t)dst << 16), W(fp, k));
    } else {
      a4.push(src, (uint32_t)(S_pad + c), W(fp, k));
      flags1[src] = kRowExtraFlag;
    }
  }
  for (int c = 0; c < C; c++) a2.push(cls[c].pdf, (uint32_t)(S_pad + c) | ((uint32_t)cls[c].dst << 16), one);
  // Stream format (den_blob.h).  Embedded row codes cost one slot per slice but (i) slices of one arc pair or less --
  // chain graphs, whose states keep a self loop and one class arc, are made of them -- leave a separate row-code
  // stream, read two slices ahead, no time to cover its L2 latency, and (ii) the code stream's registers are what the
  // G = 4 kernel is short of.  Measured on B200, ms per call, embedded / compact: cfg4 biphone 5.85 / 6.97, cfg4 skew
  // 8.93 / 9.66, cfg4 uniform 8.00 / 8.63, cfg5 (5 arcs per state: 3 pairs per slice) 1.67 / 1.59 streamed from L2 --
  // but 1.50 embedded and RESIDENT in shared memory.  So: compact only for mid-size graphs of fat-free, SHORT slices,
  // where the extra slot is a sixth of the stream and the streams do not fit on chip.
  {
    auto slice_stats = [&](int n_rows, const ArcSet &as, double &thin_frac, double &pairs_per_slice) {
      std::vector<int> deg(n_rows, 0);
      for (int r : as.row_of) deg[r]++;
      std::sort(deg.begin(), deg.end(), std::greater<int>());
      int thin = 0, n_sl = 0;
      long pairs = 0;
      for (int i = 0; i < n_rows; i += 32, n_sl++) { thin += deg[i] <= 2; pairs += (deg[i] + 1) / 2; }
      thin_frac = n_sl ? (double)thin / n_sl : 0.0;
      pairs_per_slice = n_sl ? (double)pairs / n_sl : 0.0;
    };
    double t0, p0, t1, p1;
    slice_stats(S, a0, t0, p0);
    slice_stats(S, a1, t1, p1);
    // small graphs keep their slot streams resident in shared memory (den_shared.cu), which takes the embedded format
    const int Gw = (group_width == 1 || group_width == 2) ? group_width : 4;
    const double resident = 3.0 * Kv * 8 * 1.3 + 4.0 * Gw * (2.0 * R_pad + 3.0 * D_pad + 64);
    const bool by_shape = std::max(t0, t1) > 0.25 || 0.5 * (p0 + p1) >= 3.5;
    tried_resident = !by_shape && !no_resident && resident <= 240.0 * 1024;   // verified after packing, see below
    embed = by_shape || tried_resident;
    if (const char *e = getenv("CHAIN_B200_EMBED_CODES")) embed = atoi(e) != 0;
  }
  for (int i = 0; i < kNumViews; i++) vb[i] = ViewBuild();
  // n_a / n_b bound what an IDLE lane may gather (weight 0): only entries that always hold finite data, i.e. the
  // S real states / D real pdfs -- never the row padding or the class extension (0 * garbage = NaN)
  build_view(S, a0, S, D, so, log_domain, P, false, nullptr, Kv, false, nullptr, embed, vb[0]);
  build_view(S, a1, S, D, so, log_domain, P, false, C > 0 ? &flags1 : nullptr, Kv, false, log_domain ? nullptr : leaky, embed, vb[1]);
  build_view(D, a2, S, S, so, log_domain, P, false, nullptr, Kv, false, nullptr, embed, vb[2]);
  build_view(std::max(C, 1), a3, S, 1, so, log_domain, P, true, nullptr, Kv, true, nullptr, embed, vb[3]);
  build_view(S, a4, S, 1, so, log_domain, P, true, nullptr, Kv, true, nullptr, embed, vb[4]);
  // split-row slots share the chip with the class sums: give classes back until everything fits
  int max_sub = 0;
  for (int i = 0; i < kNumViews; i++) max_sub = std::max(max_sub, vb[i].n_sub);
  const long need = 2L * R_pad + 3L * D_pad + max_sub;
  if (tried_resident) {   // embedded only for the sake of residency: does it really fit (den_shared.cu's launcher test)?
    size_t slot_bytes = 0;
    for (int i = 0; i < kNumViews; i++)
      for (int wv = 0; wv < DEN_NW; wv++) slot_bytes += (size_t)vb[i].warp_info[3 * wv + 1] * 512;
    if ((size_t)need * 4 * G + 128 + slot_bytes + 16 > (size_t)227 * 1024) {
      no_resident = true;   // pack again, compact
      continue;
    }
  }
  if (C == 0 || need <= chip_floats || attempt >= 4) break;
  c_cap = std::max(0L, (long)C - ((need - chip_floats + 1) / 2 + 4));
  }

  BlobHeader h;
  memset(&h, 0, sizeof(h));
  h.magic = kBlobMagic; h.version = kBlobVersion;
  h.S = S; h.K = Kv; h.D = D; h.nwarps = DEN_NW; h.log_domain = log_domain ? 1 : 0;
  h.opt_G = 32 / P;
  h.S_pad = S_pad; h.D_pad = D_pad;
  h.C = C; h.C_pad = C_pad; h.R_pad = R_pad;
  h.embed_codes = embed ? 1 : 0;
  size_t off = align_up(sizeof(BlobHeader));
  for (int i = 0; i < kNumViews; i++) {
    BlobView &bv = h.view[i];
    const ViewBuild &b = vb[i];
    bv.n_rows = b.n_rows; bv.n_slices = b.n_slices; bv.n_split = b.n_split; bv.n_sub = b.n_sub;
    bv.n_recs = (int64_t)b.rec_ab.size();
    h.max_sub = std::max(h.max_sub, b.n_sub);
    h.predicted_wavefronts[i] = b.predicted;
    h.ideal_wavefronts[i] = b.ideal;
    bv.off_warp_info = (int64_t)off; off = align_up(off + b.warp_info.size() * 8);
    bv.off_rows = (int64_t)off; off = align_up(off + b.rows.size() * 4);
    bv.off_recs = (int64_t)off; off = align_up(off + b.rec_ab.size() * 8 + 8);
    bv.off_split = (int64_t)off; off = align_up(off + b.split.size() * 4 + 4);
  }
  h.off_leaky = (int64_t)off; off = align_up(off + (size_t)h.S_pad * 4);
  h.off_init = (int64_t)off; off = align_up(off + (size_t)h.S_pad * 4);
  h.off_final = (int64_t)off; off = align_up(off + (size_t)h.S_pad * 4);
  h.off_cls_dst = (int64_t)off; off = align_up(off + (size_t)C_pad * 4 + 4);
  h.off_cls_pdf = (int64_t)off; off = align_up(off + (size_t)C_pad * 4 + 4);
  h.total_bytes = (int64_t)off;
  if (!out) return h.total_bytes;
  if (out_bytes < h.total_bytes || !fp || !bp || (!leaky && !log_domain) || !init || !fin) return CHAIN_B200_ERR_INVALID_ARG;
  char *base = (char *)out;
  memset(base, 0, (size_t)h.total_bytes);
  memcpy(base, &h, sizeof(h));
  for (int c = 0; c < C; c++) {
    ((int32_t *)(base + h.off_cls_dst))[c] = cls[c].dst;
    ((int32_t *)(base + h.off_cls_pdf))[c] = cls[c].pdf;
  }
  for (int i = 0; i < kNumViews; i++) {
    const BlobView &bv = h.view[i];
    const ViewBuild &b = vb[i];
    memcpy(base + bv.off_warp_info, b.warp_info.data(), b.warp_info.size() * 8);
    memcpy(base + bv.off_rows, b.rows.data(), b.rows.size() * 4);
    memcpy(base + bv.off_split, b.split.data(), b.split.size() * 4);
    uint32_t *rec = (uint32_t *)(base + bv.off_recs);
    for (size_t j = 0; j < b.rec_ab.size(); j++) {
      rec[2 * j] = b.rec_ab[j];
      memcpy(&rec[2 * j + 1], &b.rec_w[j], 4);
    }
  }
  if (leaky) memcpy(base + h.off_leaky, leaky, (size_t)S * 4);
  memcpy(base + h.off_init, init, (size_t)S * 4);
  memcpy(base + h.off_final, fin, (size_t)S * 4);
  return h.total_bytes;
}

}  // namespace chainb200

using namespace chainb200;

extern "C" {

int64_t chain_b200_den_prepare_bytes(const int32_t *ft, const int32_t *fi, const int32_t *bt, const int32_t *bi, int S,
                                     int K, int D, unsigned flags) {
  return pack_den_blob(ft, fi, nullptr, bt, bi, nullptr, nullptr, nullptr, nullptr, S, K, D,
                       (flags & CHAIN_B200_LOG_DOMAIN) != 0, (int)((flags >> 8) & 7u), nullptr, 0);
}

int chain_b200_den_prepare(const int32_t *ft, const int32_t *fi, const float *fp, const int32_t *bt,
                           const int32_t *bi, const float *bp, const float *leaky, const float *init,
                           const float *fin, int S, int K, int D, unsigned flags, void *blob, int64_t blob_bytes) {
  if (!blob) return CHAIN_B200_ERR_INVALID_ARG;
  const int64_t r = pack_den_blob(ft, fi, fp, bt, bi, bp, leaky, init, fin, S, K, D,
                                  (flags & CHAIN_B200_LOG_DOMAIN) != 0, (int)((flags >> 8) & 7u), blob, blob_bytes);
  return r < 0 ? (int)r : CHAIN_B200_OK;
}

}  // extern "C"
