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

}  // namespace trpa
