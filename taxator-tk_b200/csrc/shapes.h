// Length-bucket "shapes" for the bit-vector edit-distance kernel: a pattern of `mwords` 32-row words
// is mapped onto L lanes x W words per lane (x strips if it does not fit one pass).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define TRPA_HD __host__ __device__ __forceinline__
#else
#define TRPA_HD inline
#endif

namespace trpa {

constexpr int kNumW = 8;
constexpr int kNumL = 6;
constexpr int kNumShapes = kNumW * kNumL * 2;  // x2: HASN

TRPA_HD int shape_W(int widx) {
  switch (widx) {
    case 0: return 1;
    case 1: return 2;
    case 2: return 4;
    case 3: return 8;
    case 4: return 12;
    case 5: return 16;
    case 6: return 20;
    default: return 24;
  }
}

TRPA_HD int shape_id(int widx, int lidx, int hasn) { return (hasn * kNumL + lidx) * kNumW + widx; }
TRPA_HD int shape_widx(int id) { return id % kNumW; }
TRPA_HD int shape_lidx(int id) { return (id / kNumW) % kNumL; }
TRPA_HD int shape_hasn(int id) { return id / (kNumW * kNumL); }

// Cheapest padded capacity L*W >= mwords under cost = cap * (W+1)/W (the +1 models the per-column
// boundary/broadcast work that is amortised over the W words of a lane); patterns beyond 32*24
// words run in strips on a full warp.
// lmin_idx > 0 forces at least 2^lmin_idx lanes per pair: used when a round has too few pairs to fill
// the GPU, where more lanes per pair shorten the critical path (latency) at the price of padding.
TRPA_HD int choose_shape(uint32_t mwords, int hasn, int lmin_idx = 0) {
  if (mwords == 0) mwords = 1;
  int best_w = -1, best_l = -1;
  uint64_t best_cost = ~0ull;
  for (int l = lmin_idx; l < kNumL; ++l) {
    for (int w = 0; w < kNumW; ++w) {
      const uint32_t W = (uint32_t)shape_W(w);
      const uint64_t cap = (uint64_t)W << l;
      if (cap < mwords) continue;
      const uint64_t cost = cap * 840u * (W + 1) / W;
      if (cost < best_cost || (cost == best_cost && w > best_w)) { best_cost = cost; best_w = w; best_l = l; }
    }
  }
  if (best_w >= 0) return shape_id(best_w, best_l, hasn);
  for (int w = 0; w < kNumW; ++w) {
    const uint32_t W = (uint32_t)shape_W(w);
    const uint64_t per = (uint64_t)W * 32u;
    const uint64_t cap = ((mwords + per - 1) / per) * per;
    const uint64_t cost = cap * 840u * (W + 1) / W;
    if (cost < best_cost || (cost == best_cost && w > best_w)) { best_cost = cost; best_w = w; }
  }
  return shape_id(best_w, kNumL - 1, hasn);
}

// Minimum lanes per pair (as log2) so that n_pairs pairs still run concurrently on `lanes` lanes.
TRPA_HD int lmin_for(uint32_t n_pairs, uint32_t lanes) {
  if (n_pairs == 0) return 0;
  uint32_t per = lanes / n_pairs;
  int l = 0;
  while (l + 1 < kNumL && (2u << l) <= per) ++l;
  return l;
}

// ---------------------------------------------------------------------------------------------
// Banded kernel (myers3.cuh): geometry + cost model used to pick (W, L) per pair.
// a = half band width below the main diagonal, ua = delta + a (above); see myers3.cuh.
struct BandGeom {
  uint32_t m, n;
  uint32_t a0, a1;      // half band width at the first / last pattern row (a1 < a0: wedge)
  uint32_t row0;        // wedge: the band keeps its full width up to this pattern row
  uint32_t taper_q;     // ((a0 - a1) << 32) / (m - row0) (slope < 1 column per row); 0 = plain band
  uint32_t k;           // effective threshold; 0xffffffff when the band covers the whole matrix
  TRPA_HD uint32_t a_at(uint32_t row) const {   // branch-free: taper_q == 0 gives a0
    const uint32_t rr = row < m ? row : m;
    const uint32_t r = rr > row0 ? rr - row0 : 0u;
    return a0 - (uint32_t)(((uint64_t)r * taper_q) >> 32);
  }
  TRPA_HD uint32_t b0(uint32_t s, uint32_t R) const {
    const uint32_t x = s * R, a = a_at(x);
    return x > a ? (x - a) >> 5 : 0u;
  }
  TRPA_HD uint32_t b1(uint32_t s, uint32_t R) const {
    const uint32_t hi = (s + 1u) * R - 1u + (n - m) + a_at(s * R);
    return (hi < n - 1u ? hi : n - 1u) >> 5;
  }
  // Extra group steps between round r and r + 1 of the rotating schedule (myers3.cuh): an upper bound of
  //   max over the strips st of round r with a successor st + L < S of  b1(st) - b0(st + L) + 1 - L
  // (no lane may start a strip before it finished the previous one; a larger gap only idles).  Two bounds,
  // both from b0/b1 being non-decreasing in st and a_at non-increasing: the unclipped closed form at the
  // round's first strip (exact away from the matrix borders), and last-b1 minus first-b0 (exact for L = 1).
  // All lanes of a group and the planner evaluate exactly this function.
  TRPA_HD uint32_t round_gap(uint32_t r, uint32_t W, uint32_t L, uint32_t S) const {
    const uint32_t R = 32u * W, s0 = r * L;
    if (s0 + L >= S) return 1u;
    const uint32_t l_hi = (S - L - 1u - s0) < (L - 1u) ? (S - L - 1u - s0) : (L - 1u);
    const int cf = (int)(((n - m) + a_at(s0 * R) - 1u) >> 5) + (int)((a_at((s0 + L) * R) + 31u) >> 5) - (int)((L - 1u) * (W + 1u));
    const int b2 = (int)b1(s0 + l_hi, R) - (int)b0(s0 + L, R) + 1 - (int)L;
    const int g = cf < b2 ? cf : b2;
    return g > 1 ? (uint32_t)g : 1u;
  }
  TRPA_HD bool wedge() const { return taper_q != 0; }
};

// threshold k -> band (m <= n).  The half width is at least 32 so that consecutive strips always
// share a block; k = n is always sufficient (d <= n); patterns shorter than 64 are never banded.
// wedge (0 = none) = a1 | (x0 / 64) << 20: requested half width a1 at the last row and the number of
// rows x0 after the first a0 = (k - delta) / 2 ones for which the band keeps its full width; from row
// a0 + x0 on it narrows linearly to a1 (a wedge is accepted only with the kernel's certificate).
constexpr uint32_t kWedgeA1Bits = 20;
TRPA_HD uint32_t wedge_pack(uint32_t a1, uint32_t x0) {
  const uint32_t xq = (x0 + 63u) >> 6;
  if (a1 >= (1u << kWedgeA1Bits) || xq >= (1u << (32 - kWedgeA1Bits))) return 0u;
  return a1 | (xq << kWedgeA1Bits);
}
TRPA_HD BandGeom band_from_k(uint32_t m, uint32_t n, uint32_t k, bool force_full = false, uint32_t wedge = 0) {
  BandGeom g;
  const uint32_t delta = n - m;
  g.m = m; g.n = n; g.taper_q = 0; g.row0 = 0;
  if (k < delta + 64u) k = delta + 64u;
  if (k > n) k = n;
  if (force_full || k < delta + 64u) { g.a0 = g.a1 = m; g.k = 0xffffffffu; return g; }
  g.a0 = g.a1 = (k - delta) >> 1; g.k = k;
  if (wedge && m >= 256u) {
    uint32_t a1 = wedge & ((1u << kWedgeA1Bits) - 1u);
    const uint32_t row0 = g.a0 + ((wedge >> kWedgeA1Bits) << 6);
    if (a1 < 32u) a1 = 32u;
    if (a1 + 16u < g.a0 && row0 + 64u < m && g.a0 - a1 < m - row0) {   // slope < 1: b0 / b1 stay monotonic
      g.a1 = a1; g.row0 = row0;
      g.taper_q = (uint32_t)((((uint64_t)(g.a0 - a1)) << 32) / (m - row0));
      if (!g.taper_q) { g.a1 = g.a0; g.row0 = 0; }
    }
  }
  return g;
}

// Group steps the rotating schedule needs for one attempt (long schedules are sampled: the gap is
// piecewise linear in the round index).
TRPA_HD uint64_t band_steps(const BandGeom& g, int W, int L) {
  const uint32_t mwords = (g.m + 31u) >> 5, nblk = (g.n + 31u) >> 5;
  const uint32_t S = (mwords + W - 1) / W;
  uint64_t off = 0;
  if (S > (uint32_t)L) {
    const uint32_t rounds = (S - 1u) / (uint32_t)L;          // rounds that have a successor round
    const uint32_t stride = rounds > 5u ? (rounds + 4u) / 5u : 1u;   // 5 samples: the planner is instruction bound on this loop
    uint64_t acc = 0;
    uint32_t n = 0;
    for (uint32_t r = 0; r < rounds; r += stride, ++n) acc += (uint64_t)L + (uint64_t)g.round_gap(r, (uint32_t)W, (uint32_t)L, S);
    // (32-bit division whenever the product fits: the planner is instruction bound, profiles/r2_plan_decide_ncu.md)
    off = (acc < 0x10000u && rounds < 0x10000u) ? (uint64_t)((uint32_t)acc * rounds / n) : acc * rounds / n;
  }
  return off + ((S - 1u) % (uint32_t)L) + nblk;
}

// Planner parameters (trpa_set_tuning hooks; results never depend on them).
//  hint_mul64 / hint_add: band threshold from a distance ESTIMATE = hint * hint_mul64 / 64 + hint_add
//  cost model, alu-pipe instructions of one lane-step: 32 columns x (word10 / 10 per word + col10 / 10 per
//  column) + step (boundary hand-over, schedule); strip set-up (equality table, geometry): setup + setup_w * W
struct PlanParams {
  uint32_t hint_mul64 = 72, hint_add = 32;
  uint32_t tail_log2 = 22;   // latency penalty time + time^2 / 2^tail_log2 (>= 9)
  uint32_t word10 = 100, col10 = 35, step = 200, setup = 250, setup_w = 25;   // SASS + ncu region counts, profiles/r03_myers3_ncu.md
  // wedge (plan_kernel): the band narrows at wedge_s8 / 8 of the slowest prefix mismatch rate once wedge_e0 errors
  // are expected; k0 gets wedge_cushion extra.  More aggressive = fewer cells, more failed certificates (re-runs):
  // measured on C2, s = 6/8: 283.7 k seg/s, 3 re-runs per step; 8/8: 296.2 k, 1074 re-runs (gpurun_out r2_07); C4 neutral.
  // wedge_max_k: beyond ~2048 the certificates fail too often (C4 with 4096: 40 k re-runs per step, -13 %).
  uint32_t wedge_s8 = 8, wedge_e0 = 30, wedge_cushion = 16, wedge_max_k = 2048;
  // the same for pairs whose threshold comes from a distance ESTIMATE (hint) instead of the mismatch profile: errors are
  // assumed to accumulate uniformly at hint / m; 0 = no wedge for such pairs
  uint32_t wedge_hint_s8 = 0, wedge_hint_max_k = 16384;
};
TRPA_HD uint64_t band_step_ops(int W, const PlanParams& pp) { return 32ull * ((uint64_t)pp.word10 * (uint64_t)W + pp.col10) / 10ull + pp.step; }
TRPA_HD uint64_t band_setup_ops(int W, const PlanParams& pp) { return pp.setup + (uint64_t)pp.setup_w * (uint64_t)W; }

// duration classes inside a shape bucket (longest-processing-time-first order of the persistent launches)
constexpr int kNumCls = 24;
TRPA_HD uint32_t duration_class(uint64_t time) {
  // l = floor(log2(time)) (0 for time <= 1)
#ifdef __CUDA_ARCH__
  const uint32_t l = time > 1u ? 63u - (uint32_t)__clzll((long long)time) : 0u;
#else
  uint32_t l = 0;
  while (time > 1u) { time >>= 1; ++l; }
#endif
  return l < 8u ? 0u : (l - 8u < (uint32_t)kNumCls ? l - 8u : (uint32_t)kNumCls - 1u);
}

struct ShapeCost { uint64_t time, cost; };   // time: alu instructions on the critical lane; cost = time * L
TRPA_HD ShapeCost band_shape_cost(const BandGeom& g, int widx, int lidx, const PlanParams& pp) {
  const int W = shape_W(widx), L = 1 << lidx;
  const uint32_t mwords = (g.m + 31u) >> 5;
  const uint32_t S = (mwords + W - 1) / W;
  ShapeCost c;
  c.time = band_steps(g, W, L) * band_step_ops(W, pp) + (uint64_t)((S + L - 1) / L) * band_setup_ops(W, pp);
  c.cost = c.time * (uint64_t)L;
  return c;
}

}  // namespace trpa
