// CPU check of the band / wedge geometry shared by the planner and the edit-distance kernel
// (taxator-tk_b200/csrc/shapes.h): BandGeom::round_gap must be an upper bound of the exact requirement
// (max over the strips of a round of b1(st) - b0(st + L) + 1 - L), tight away from the matrix borders,
// and the rotating schedule built from it must never ask a lane to work on two strips at once nor let
// a strip run ahead of the strip above it.  Exit code 0 = all good.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <random>
#include <vector>
#include "shapes.h"
using namespace trpa;

int main() {
  std::mt19937_64 rng(12345);
  auto rnd = [&](uint64_t lo, uint64_t hi) { return (uint32_t)(lo + rng() % (hi - lo + 1)); };
  long checked = 0, loose = 0, rounds_total = 0, nontrivial = 0;
  for (int it = 0; it < 60000; ++it) {
    const uint32_t m = rnd(64, it % 7 == 0 ? 60000 : 9000);
    const uint32_t n = m + (it % 3 == 0 ? 0u : rnd(0, m / 4 + 40));
    const uint32_t k = rnd(0, n + 100);
    uint32_t wedge = 0;
    if (it % 2) wedge = wedge_pack(rnd(0, 600), rnd(0, m));
    const BandGeom g = band_from_k(m, n, k, it % 97 == 0, wedge);
    const int Ws[] = {1, 2, 4, 8, 12, 16, 20, 24};
    const uint32_t W = (uint32_t)Ws[rnd(0, 7)], L = 1u << rnd(0, 5), R = 32u * W;
    const uint32_t mwords = (m + 31u) >> 5, S = (mwords + W - 1) / W, nblk = (n + 31u) >> 5;
    // monotonic geometry, strips overlap their neighbours by at least one block
    for (uint32_t s = 0; s + 1 < S; ++s) {
      if (g.b0(s + 1, R) < g.b0(s, R) || g.b1(s + 1, R) < g.b1(s, R)) { printf("not monotonic m=%u n=%u k=%u s=%u\n", m, n, k, s); return 1; }
      if (g.b0(s + 1, R) > g.b1(s, R)) { printf("strips do not share a block m=%u n=%u k=%u s=%u\n", m, n, k, s); return 1; }
    }
    if (g.b1(S - 1, R) != nblk - 1 || g.b0(0, R) != 0) { printf("band misses a matrix corner m=%u n=%u k=%u\n", m, n, k); return 1; }
    // schedule: T0(s) = offset_r + s % L, offset_{r+1} = offset_r + L + gap_r
    std::vector<long> T0(S);
    long off = 0;
    for (uint32_t r = 0; r * L < S; ++r) {
      for (uint32_t l = 0; l < L && r * L + l < S; ++l) T0[r * L + l] = off + l;
      const uint32_t gap = g.round_gap(r, W, L, S);
      int exact = 1;
      for (uint32_t l = 0; l < L; ++l) {
        const uint32_t st = r * L + l;
        if (st + L >= S) break;
        const int term = (int)g.b1(st, R) - (int)g.b0(st + L, R) + 1 - (int)L;
        if (term > exact) exact = term;
      }
      if ((int)gap < exact) { printf("gap too small m=%u n=%u k=%u W=%u L=%u r=%u: %u < %d\n", m, n, k, W, L, r, gap, exact); return 1; }
      if ((int)gap > exact) ++loose;
      if (exact > 1) ++nontrivial;
      ++rounds_total;
      off += L + gap;
    }
    for (uint32_t s = 0; s < S; ++s) {
      if (s >= 1 && T0[s] - T0[s - 1] < 1) { printf("strip runs ahead m=%u n=%u s=%u\n", m, n, s); return 1; }
      if (s + L < S && T0[s] + (long)g.b1(s, R) >= T0[s + L] + (long)g.b0(s + L, R)) { printf("lane busy m=%u n=%u k=%u W=%u L=%u s=%u\n", m, n, k, W, L, s); return 1; }
    }
    ++checked;
  }
  printf("ok: %ld geometries, %ld rounds (%ld with a gap > 1), %ld gaps above the exact requirement\n", checked, rounds_total, nontrivial, loose);
  return 0;
}
