// Edit-distance kernel: the same integer as -globalAlignmentScore(.., MyersBitVector()) (hh:150), but
// only the cells inside an Ukkonen band are computed (exact: see below), and the pattern strips of
// a pair ROTATE over the lanes of its group so that a band of any width keeps every lane busy.
//
//  * pair (m <= n, delta = n - m), threshold k >= delta: every alignment of cost <= k stays inside
//    the diagonals [-a, delta + a], a = (k - delta) / 2.  Strip s (32*W pattern rows) therefore
//    only needs the text blocks b0(s)..b1(s) (32 columns each) that intersect that band.
//  * outside the computed region the DP is replaced by upper bounds: a strip starts its first block
//    with VP = all ones (vertical deltas +1) and, for blocks the strip above never computed, takes
//    horizontal deltas +1 as its top boundary.  Hence the computed value v >= d always, and v == d
//    whenever v <= k (then an optimal path lies inside the band and is computed exactly).  v > k
//    proves d > k: the group widens the band (k <- min(v, 3k); v itself is an upper bound of d) and
//    runs the pair again.  k = n is always sufficient (d <= n), so the loop ends.
//  * schedule: a group of L lanes owns the pair; strip s runs on lane s % L in round r = s / L; at
//    group step t it works on block t - T0(s), T0 = offset_r + (s % L), so a strip is always exactly
//    one block behind the strip above it: lanes 1..L-1 take their top boundary (HP/HN shift-out
//    bits, running bottom-row value; the carry into a strip's adder IS the HN bit, see myers3_column) from the lane above by shuffle; lane 0 takes it
//    from a per-group scratch line in L2 that lane L-1 wrote at least one step earlier (prefetched
//    one step ahead).  offset_{r+1} - offset_r = L + gap_r with gap_r just large enough that no lane
//    has to start a strip before it finished the previous one.
//  * the score is assembled from deltas: corner(s) (value above the strip at its first column,
//    handed down the same way) + the top-boundary deltas of the last strip + popc(VP) - popc(VN) of
//    the last column.
//  * wedge (PairDesc.aux = half width at the last pattern row, 0 = plain band): errors accumulate along
//    an alignment, so towards the end of the matrix a cell far from the diagonal cannot lie on a path of
//    cost <= k any more (its value plus the remaining diagonal distance exceeds k).  The planner
//    therefore lets the band narrow linearly.  Unlike the plain band a wedge is not exact by
//    construction; the kernel CERTIFIES it: a path that leaves the computed region does so through a
//    cell of a strip's bottom row left of the next strip's first column, through a strip's last
//    column, or along row 0 / column 0, and then still has to come back to the final diagonal, so it
//    costs at least cert = min over those cells of (computed value + remaining diagonal distance).
//    cert >= v proves that no such path beats v; otherwise the pair is run again with the plain band.
// scripts/band_model.py is an executable model of exactly this schedule (checked against a plain DP).
#pragma once
#include "bitvec.cuh"
#include "shapes.h"

namespace trpa {

__device__ __forceinline__ uint4 ldcg_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stcg_u4(uint4* p, uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// PairDesc.pad as written by plan_kernel: bits 0..7 shape id, bits 8..31 initial threshold k0
// (kPadKFull = no band).
constexpr u32 kPadKFull = 0xffffffu;

// Equality-mask table in shared memory: entry(sym, wq) of a lane = the masks of pattern words 4wq..4wq+3 of
// its strip for text symbol sym (one uint4; the 32 lanes of a warp are contiguous: conflict-free LDS.128).
//  * pairs without N: the table of a warp is 2 KB aligned and laid out [wq][sym][lane], so the address of a
//    column's row is ONE LOP3 -- base | (codes >> shift) & 0x600 -- on the text's 2-bit column codes
//    (common.cuh nt_codes, written by the staging kernel) and wq is an immediate offset of the load;
//  * pairs with N (5 symbols, rare): [sym][wq][lane], symbol taken from the three bit-planes.
template <int W, bool HASN>
struct Myers3Cfg {
  static constexpr int WQ = (W + 3) / 4;
  static constexpr int NSYM = HASN ? 5 : 4;
  static constexpr int kWarps = 4;
  static constexpr u32 kWarpBytes = HASN ? (u32)(NSYM * WQ * 512) : (u32)(WQ * 2048);
  static constexpr size_t kSmemBytes = (size_t)kWarps * kWarpBytes + (HASN ? 0 : 2048);
  __device__ static constexpr u32 entry(int sym, int wq) { return HASN ? (u32)((sym * WQ + wq) * 512) : (u32)(wq * 2048 + sym * 512); }
};

__device__ __forceinline__ uint4 lds_u4(u32 addr, u32 imm) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr + imm));
  return v;
}
__device__ __forceinline__ void sts_u4(u32 addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One DP column for the W words of a lane .  The carry into the strip's adder is the
// HN bit handed down from the row above: the exact carry of the strip above only matters where
// Eq = VN = 0 in bit 0, and there it equals that bit (Myers' block formulation: hin < 0 <=> Eq |= 1),
// so no carry word travels between strips.
template <int W>
__device__ __forceinline__ void myers3_column(u32 (&VP)[W], u32 (&VN)[W], const u32 (&Eq)[W], u32 hpc, u32 hnc,
                                              u32& hpOut, u32& hnOut) {
  u32 T[W], S[W];
#pragma unroll
  for (int w = 0; w < W; ++w) T[w] = Eq[w] & VP[w];
  add_words<W>(S, VP, T, hnc);
  u32 hpPrev = hpc, hnPrev = hnc;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const u32 X = Eq[w] | VN[w];
    const u32 D0 = lop3<0xBE>(S[w], VP[w], X);       // (S ^ VP) | X
    const u32 HN = VP[w] & D0;
    const u32 HP = lop3<0xF1>(VN[w], VP[w], D0);     // VN | ~(VP | D0)
    const u32 Xh = __funnelshift_l(hpPrev, HP, 1);
    const u32 HNs = __funnelshift_l(hnPrev, HN, 1);
    VN[w] = Xh & D0;
    VP[w] = lop3<0xF1>(HNs, Xh, D0);                 // HNs | ~(Xh | D0)
    hpPrev = HP;
    hnPrev = HN;
  }
  hpOut = __funnelshift_l(hpPrev, hpOut, 1);
  hnOut = __funnelshift_l(hnPrev, hnOut, 1);
}

// stats[0] += executed 32x32-cell word-blocks, stats[1] += band retries, stats[2] += pairs,
// stats[3] += wedges whose certificate failed (re-run with the plain band)
template <int W, bool HASN>
__global__ void __launch_bounds__(128)
myers3_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
              const uint2* __restrict__ planes, const u32* __restrict__ nplane, const uint2* __restrict__ codes,
              int* __restrict__ out, int L, uint4* __restrict__ scratch, u32 scratch_stride,
              const uint2* __restrict__ bucket, u32* __restrict__ cursor, unsigned long long* __restrict__ stats,
              int force_full) {
  typedef Myers3Cfg<W, HASN> Cfg;
  constexpr int WQ = Cfg::WQ;
  constexpr int NSYM = Cfg::NSYM;
  constexpr u32 R = 32u * W;
  extern __shared__ uint4 eqtab[];
  if (bucket) {
    const uint2 bk = *bucket;
    pairs += bk.x;
    count = bk.y;
  }
  const u32 lane = threadIdx.x & 31;
  const u32 warp_in_cta = threadIdx.x >> 5;
  const u32 G = 32 / L;
  const u32 g = lane / L;
  const u32 sl = lane - g * L;
  const bool in_group = g < G;   // L need not divide 32: leftover lanes idle
  const u32 gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u) << (g * L));
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 slot = warp_gid * G + g;
  u32 sbase = (u32)__cvta_generic_to_shared(eqtab);
  if (!HASN) sbase = (sbase + 2047u) & ~2047u;
  const u32 mybase = sbase + warp_in_cta * Cfg::kWarpBytes + lane * 16u;   // + Cfg::entry(sym, wq)
  uint4* my_scratch = scratch + (size_t)slot * scratch_stride;

  bool active = false, exhausted = !in_group, need_setup = false;
  // pair (uniform inside a group)
  u32 m = 0, n = 0, pw = 0, tw = 0, oidx = 0, mwords = 0, S = 0, k = 0, t = 0;
  BandGeom bg;   // band / wedge of the current attempt
  bg.m = bg.n = 1; bg.a0 = bg.a1 = 0; bg.taper_q = 0; bg.k = 0;
  // lane
  u32 s = 0, r = 0, T0 = 0, sb0 = 0, sb1 = 0, pb1 = 0, nb0 = 0;
  int corner = 0, botacc = 0, tsum = 0, vfinal = 0, cert = 0x7fffffff;
  u32 VP[W], VN[W];
  u32 hpOut = 0, hnOut = 0;
  uint4 pre = make_uint4(0u, 0u, 0u, 0u);
  bool have_pre = false;
  u32 nblocks = 0, nretry = 0, npairs = 0, nwfail = 0;

  auto gb0 = [&](u32 st) -> u32 { return bg.b0(st, R); };
  auto gb1 = [&](u32 st) -> u32 { return bg.b1(st, R); };
  auto set_band = [&](u32 kk, u32 a1_req) {
    bg = band_from_k(m, n, kk, force_full != 0, a1_req);
    k = bg.k;
  };
  auto start_attempt = [&]() {
    t = 0; s = sl; r = 0; T0 = sl;
    need_setup = s < S;
    hpOut = 0; hnOut = 0; botacc = 0;
    cert = 0x7fffffff;
  };
  auto setup_strip = [&]() {
    sb0 = gb0(s); sb1 = gb1(s); pb1 = s ? gb1(s - 1u) : 0u;
    nb0 = s + 1u < S ? gb0(s + 1u) : 0u;
    if (bg.wedge()) {   // ways out of the region along the matrix borders (row 0, column 0)
      const int delta = (int)(n - m);
      const u32 nblk = (n + 31u) >> 5;
      if (s == 0 && sb1 + 1u < nblk) cert = min(cert, 64 * (int)(sb1 + 1u) - delta);
      if (s > 0 && sb0 > 0 && gb0(s - 1u) == 0) cert = min(cert, 2 * (int)(s * R + 1u) + delta);
    }
    const u32 kbase = s * W;
#pragma unroll
    for (int q = 0; q < WQ; ++q) {
      u32 e[NSYM][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = 4 * q + j;
        uint2 p = make_uint2(0u, 0u);
        u32 pn = 0;
        if (w < W && kbase + w < mwords) {
          p = planes[pw + kbase + w];
          if (HASN) pn = nplane[pw + kbase + w];
        }
        e[0][j] = ~(p.x | p.y | pn);   // A
        e[1][j] = p.x & ~p.y;          // C
        e[2][j] = ~p.x & p.y;          // G
        e[3][j] = p.x & p.y;           // T
        if (HASN) e[NSYM - 1][j] = pn; // N matches N
      }
#pragma unroll
      for (int sym = 0; sym < NSYM; ++sym) sts_u4(mybase + Cfg::entry(sym, q), make_uint4(e[sym][0], e[sym][1], e[sym][2], e[sym][3]));
    }
#pragma unroll
    for (int w = 0; w < W; ++w) { VP[w] = 0xffffffffu; VN[w] = 0u; }
    tsum = 0;
    have_pre = false;
  };
  // extra steps between round r and r+1 so that no lane starts a strip before finishing the previous one
  auto round_gap = [&](u32 rr) -> u32 { return bg.round_gap(rr, (u32)W, (u32)L, S); };

  for (;;) {
    if (!active && !exhausted) {  // uniform inside a group
      u32 idx = 0;
      if (sl == 0) idx = atomicAdd(cursor, 1u);
      idx = __shfl_sync(gmask, idx, g * L);
      if (idx >= count) exhausted = true;
      else {
        const PairDesc pd = pairs[idx];
        const SeqDesc A = seqs[pd.a], B = seqs[pd.b];
        if (A.len < B.len) { m = A.len; pw = A.woff; n = B.len; tw = B.woff; }
        else               { m = B.len; pw = B.woff; n = A.len; tw = A.woff; }
        oidx = pd.out;
        mwords = (m + 31) >> 5;
        if (sl == 0) ++npairs;
        if (m == 0) {
          if (sl == 0) out[oidx] = (int)n;   // empty pattern: n insertions (m <= n)
        } else {
          S = (mwords + W - 1) / W;
          const u32 k0 = pd.pad >> 8;
          set_band(k0 >= kPadKFull ? 0xffffffffu : k0, pd.aux);
          start_attempt();
          active = true;
        }
      }
    }
    if (__all_sync(0xffffffffu, !active && exhausted)) break;

    // boundary handed down by the lane above (its outputs of the previous step)
    const u32 src = sl == 0 ? lane : lane - 1;
    const u32 hpIn = __shfl_sync(0xffffffffu, hpOut, src);
    const u32 hnIn = __shfl_sync(0xffffffffu, hnOut, src);
    const int botIn = __shfl_sync(0xffffffffu, botacc, src);

    bool do_block = false, islast = false, fin_lane = false;
    u32 hpc = 0, hnc = 0, run = 32, b = 0;
    if (active && s < S) {
      if (need_setup) { setup_strip(); need_setup = false; }
      const int bi = (int)t - (int)T0;
      if (bi >= (int)sb0) {
        do_block = true;
        b = (u32)bi;
        int boti = 0;
        if (s > 0 && b <= pb1) {  // the strip above computed this block
          if (sl == 0) {
            const uint4 q = have_pre ? pre : ldcg_u4(my_scratch + b);
            hpc = q.x; hnc = q.y; boti = (int)q.w;
          } else { hpc = hpIn; hnc = hnIn; boti = botIn; }
        } else { hpc = 0xffffffffu; hnc = 0u; }   // row 0, or upper bound (+1 deltas)
        if (b == sb0) {
          corner = s ? boti - (__popc(hpc) - __popc(hnc)) : 0;
          botacc = corner + (int)R;
        }
        have_pre = false;
        if (sl == 0 && s > 0 && b + 1u <= (sb1 < pb1 ? sb1 : pb1)) { pre = ldcg_u4(my_scratch + b + 1u); have_pre = true; }
        islast = s + 1u == S;
        if (islast) {
          const u32 ncols = min(32u, n - 32u * b);
          run = ncols;
          const u32 vm = 0xffffffffu << (32u - ncols);
          tsum += __popc(hpc & vm) - __popc(hnc & vm);
        }
      }
    }
    const bool any_partial = __any_sync(0xffffffffu, do_block && run < 32u);
    if (do_block) {
      hpOut = 0; hnOut = 0;
      // one column: row = shared address of the lane's entry (sym, 0) for the column's text symbol
      auto column = [&](u32 row) {
        u32 Eq[W];
#pragma unroll
        for (int q = 0; q < WQ; ++q) {
          const uint4 e = lds_u4(row, Cfg::entry(0, q));
          if (4 * q + 0 < W) Eq[4 * q + 0] = e.x;
          if (4 * q + 1 < W) Eq[4 * q + 1] = e.y;
          if (4 * q + 2 < W) Eq[4 * q + 2] = e.z;
          if (4 * q + 3 < W) Eq[4 * q + 3] = e.w;
        }
        myers3_column<W>(VP, VN, Eq, hpc, hnc, hpOut, hnOut);
        hpc <<= 1; hnc <<= 1;
      };
      if (!HASN) {
        const uint2 cx = codes[tw + b];   // 2-bit column codes, first column in the low bits
        if (!any_partial) {
          u32 x = cx.x;
#pragma unroll 1
          for (int g4 = 0; g4 < 8; ++g4) {
            if (g4 == 4) x = cx.y;
#pragma unroll
            for (int j = 0; j < 4; ++j) column(lop3<0xF8>(mybase, x << (9 - 2 * j), 0x600u));   // base | sym << 9
            x >>= 8;
          }
        } else {
          // some lane of the warp is in the partial final block of its last strip: everybody takes the
          // per-column predicated loop for this one step instead of serialising two loops
#pragma unroll 1
          for (u32 c = 0; c < 32u; ++c)
            if (c < run) column(mybase + ((((c < 16u ? cx.x : cx.y) >> (2u * (c & 15u))) & 3u) << 9));
        }
      } else {
        const uint2 tx = planes[tw + b];
        u32 t0 = __brev(tx.x), t1 = __brev(tx.y), tN = __brev(nplane[tw + b]);
        auto column_n = [&]() {
          const u32 sym = (tN >> 31) ? (u32)(NSYM - 1) : (t0 >> 31) + 2u * (t1 >> 31);
          column(mybase + sym * Cfg::entry(1, 0));
          t0 <<= 1; t1 <<= 1; tN <<= 1;
        };
        if (!any_partial) {
#pragma unroll 4
          for (int c = 0; c < 32; ++c) column_n();
        } else {
#pragma unroll 1
          for (u32 c = 0; c < 32u; ++c)
            if (c < run) column_n();
        }
      }
      ++nblocks;
      botacc += __popc(hpOut) - __popc(hnOut);
      if (sl == (u32)L - 1u && s + 1u < S) stcg_u4(my_scratch + b, make_uint4(hpOut, hnOut, 0u, (u32)botacc));
      if (bg.wedge() && s + 1u < S) {   // certificate: cells through which a path can leave the region
        const int delta = (int)(n - m);
        if (b < nb0) cert = min(cert, botacc + (int)((s + 1u) * R - 32u * (b + 1u)) + delta);
        if (b == sb1 && 32u * (b + 1u) < n) cert = min(cert, botacc + (int)(32u * (b + 1u) - (s + 1u) * R) - delta);
      }
      if (b == sb1) {  // strip finished
        if (islast) {
          int acc = corner + tsum;
#pragma unroll
          for (int w = 0; w < W; ++w) {
            const u32 kw = s * W + w;
            u32 valid = 0;
            if (kw < mwords) {
              const u32 rem = m - 32u * kw;
              valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
            }
            acc += __popc(VP[w] & valid) - __popc(VN[w] & valid);
          }
          vfinal = acc;
          fin_lane = true;
        }
        if (s + (u32)L < S) T0 += (u32)L + round_gap(r);
        ++r;
        s += (u32)L;
        need_setup = s < S;
      }
    }
    if (active) ++t;
    __syncwarp();   // orders this step's scratch writes before later steps' reads
    const u32 fin = __ballot_sync(0xffffffffu, fin_lane) & gmask;
    if (in_group && fin) {  // uniform inside a group
      const int v = __shfl_sync(gmask, vfinal, __ffs(fin) - 1);
      bool ok = (u32)v <= k;
      if (bg.wedge()) {
        int cm = cert;
        for (int o = L >> 1; o > 0; o >>= 1) cm = min(cm, __shfl_xor_sync(gmask, cm, o));
        if (ok && cm < v) { ok = false; if (sl == 0) ++nwfail; }
      }
      if (ok) {
        if (sl == 0) out[oidx] = v;
        active = false;
      } else {
        // plain band from here on: v <= k means only the certificate failed (the band of k is then
        // exact); otherwise d > k is proven and v is an upper bound of d
        const u32 k3 = k > 0x55555555u ? 0xffffffffu : 3u * k;
        set_band((u32)v <= k ? k : ((u32)v < k3 ? (u32)v : k3), 0u);
        start_attempt();
        if (sl == 0) ++nretry;
      }
    }
  }
  if (stats) {
    unsigned long long wb = (unsigned long long)nblocks * W;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wb += __shfl_xor_sync(0xffffffffu, wb, o);
      nretry += __shfl_xor_sync(0xffffffffu, nretry, o);
      nwfail += __shfl_xor_sync(0xffffffffu, nwfail, o);
      npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
    }
    if (lane == 0) {
      atomicAdd(&stats[0], wb);
      if (nretry) atomicAdd(&stats[1], (unsigned long long)nretry);
      atomicAdd(&stats[2], (unsigned long long)npairs);
      if (nwfail) atomicAdd(&stats[3], (unsigned long long)nwfail);
    }
  }
}

}  // namespace trpa
