// Second-generation edit-distance kernel (same integer as myers.cuh / hh:150):
//  * persistent sub-warp groups with DYNAMIC REFILL: every group of L lanes pulls the next pair of
//    its shape bucket from a device-side cursor as soon as it finishes one, so pairs of different
//    text length in one warp no longer wait for the longest (mixed-length batches, tails);
//  * the per-character equality masks come from a per-lane table in SHARED MEMORY (built once per
//    pair and strip: 4 (5 with N) masks per pattern word, one conflict-free LDS.128 per 4 words and
//    column) instead of 2 LOP3 per word -- the LOP3/IADD3.X/SHF "alu" pipe is the bottleneck
//    (profiles/r01_myers_ncu.md), the LSU pipe is idle;
//  * the column update is written as explicit 3-input LOP3s (7 per word).
// Per 32-cell word-step: 7 LOP3 + 1 IADD3.X + 2 SHF on the alu pipe (+ per-column boundary work).
#pragma once
#include "myers.cuh"

namespace trpa {

template <int LUT>
__device__ __forceinline__ u32 lop3(u32 a, u32 b, u32 c) {
  u32 d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return d;
}

// One DP column for the W words of a lane; Eq[] already holds the equality masks of the column's
// character.  Boundary conventions as in myers_column.
template <int W>
__device__ __forceinline__ void myers2_column(u32 (&VP)[W], u32 (&VN)[W], const u32 (&Eq)[W], u32 hpc, u32 hnc,
                                              u32 cc, u32& hpOut, u32& hnOut, u32& cOut) {
  u32 T[W], S[W];
#pragma unroll
  for (int w = 0; w < W; ++w) T[w] = Eq[w] & VP[w];
  const u32 co = add_words<W>(S, VP, T, cc);
  cOut = cOut + cOut + co;
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

template <int W, bool HASN>
struct Myers2Cfg {
  static constexpr int WQ = (W + 3) / 4;
  static constexpr int NSYM = HASN ? 5 : 4;
  static constexpr int kWarps = 4;
  static constexpr size_t kSmemBytes = (size_t)kWarps * NSYM * WQ * 32 * sizeof(uint4);
};

// pairs/bucket as in myers_kernel; cursor: zero-initialised device counter of this launch.
// scratch: 3*scratch_stride words per persistent group slot (gridDim.x*4 warps * 32/L groups).
template <int W, bool HASN>
__global__ void __launch_bounds__(128)
myers2_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
              const uint2* __restrict__ planes, const u32* __restrict__ nplane, int* __restrict__ out, int L,
              u32* __restrict__ scratch, u32 scratch_stride, const uint2* __restrict__ bucket, u32* __restrict__ cursor) {
  typedef Myers2Cfg<W, HASN> Cfg;
  constexpr int WQ = Cfg::WQ;
  constexpr int NSYM = Cfg::NSYM;
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
  const u32 gmask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u) << (g * L));
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 slot = warp_gid * G + g;
  uint4* my = eqtab + (size_t)warp_in_cta * NSYM * WQ * 32 + lane;   // entry(sym, wq) = my[(sym*WQ + wq)*32]
  u32* my_scratch = scratch ? scratch + (size_t)slot * 3 * scratch_stride : nullptr;

  bool active = false, exhausted = false;
  u32 m = 0, n = 0, pw = 0, tw = 0, oidx = 0, mwords = 0, nblk = 0, nstrips = 0, s = 0, t = 0, kbase = 0;
  int score = 0;
  u32 VP[W], VN[W];
  u32 hpOut = 0, hnOut = 0, cOut = 0;

  auto setup_strip = [&]() {
    kbase = (s * L + sl) * W;
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
      for (int sym = 0; sym < NSYM; ++sym) my[(sym * WQ + q) * 32] = make_uint4(e[sym][0], e[sym][1], e[sym][2], e[sym][3]);
    }
#pragma unroll
    for (int w = 0; w < W; ++w) { VP[w] = 0xffffffffu; VN[w] = 0u; }
    hpOut = 0; hnOut = 0; cOut = 0;
    t = 0;
  };

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
        nblk = (n + 31) >> 5;
        nstrips = (mwords + (u32)L * W - 1) / ((u32)L * W);
        if (mwords == 0 || nblk == 0) {
          if (sl == 0) out[oidx] = (int)n;   // empty pattern: n insertions (m <= n)
        } else {
          s = 0; score = 0;
          setup_strip();
          active = true;
        }
      }
    }
    if (__all_sync(0xffffffffu, !active && exhausted)) break;

    u32 hpIn = __shfl_up_sync(0xffffffffu, hpOut, 1, L);
    u32 hnIn = __shfl_up_sync(0xffffffffu, hnOut, 1, L);
    u32 cIn = __shfl_up_sync(0xffffffffu, cOut, 1, L);
    if (active) {
      const int blk = (int)t - (int)sl;
      if (blk >= 0 && blk < (int)nblk) {
        const u32 ncols = min(32u, n - 32u * (u32)blk);
        if (sl == 0) {
          if (s == 0) { hpIn = 0xffffffffu; hnIn = 0u; cIn = 0u; }
          else {
            hpIn = my_scratch[3 * blk + 0];
            hnIn = my_scratch[3 * blk + 1];
            cIn = my_scratch[3 * blk + 2];
          }
        }
        const uint2 tx = planes[tw + blk];
        u32 t0 = __brev(tx.x), t1 = __brev(tx.y), tN = 0;
        if (HASN) tN = __brev(nplane[tw + blk]);
        u32 hpc = hpIn << (32 - ncols), hnc = hnIn << (32 - ncols), cc = cIn << (32 - ncols);
        hpOut = 0; hnOut = 0; cOut = 0;
        auto column = [&]() {
          u32 sym = (t0 >> 31) + 2u * (t1 >> 31);
          if (HASN) sym = (tN >> 31) ? (u32)(NSYM - 1) : sym;
          const uint4* row = my + sym * (WQ * 32);
          u32 Eq[W];
#pragma unroll
          for (int q = 0; q < WQ; ++q) {
            const uint4 e = row[q * 32];
            if (4 * q + 0 < W) Eq[4 * q + 0] = e.x;
            if (4 * q + 1 < W) Eq[4 * q + 1] = e.y;
            if (4 * q + 2 < W) Eq[4 * q + 2] = e.z;
            if (4 * q + 3 < W) Eq[4 * q + 3] = e.w;
          }
          myers2_column<W>(VP, VN, Eq, hpc, hnc, cc, hpOut, hnOut, cOut);
          t0 <<= 1; t1 <<= 1; tN <<= 1; hpc <<= 1; hnc <<= 1; cc <<= 1;
        };
        if (ncols == 32) {
#pragma unroll 4
          for (int c = 0; c < 32; ++c) column();
        } else {
#pragma unroll 1
          for (u32 c = 0; c < ncols; ++c) column();
        }
        if (sl == (u32)L - 1 && s + 1 < nstrips) {
          my_scratch[3 * blk + 0] = hpOut;
          my_scratch[3 * blk + 1] = hnOut;
          my_scratch[3 * blk + 2] = cOut;
        }
      }
      ++t;
      if (t == nblk + (u32)L - 1) {  // strip finished (t is uniform inside the group)
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const u32 k = kbase + w;
          u32 valid = 0;
          if (k < mwords) {
            const u32 rem = m - 32u * k;
            valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
          }
          score += __popc(VP[w] & valid) - __popc(VN[w] & valid);
        }
        ++s;
        if (s == nstrips) {
          int acc = score;
          for (int o = L >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
          if (sl == 0) out[oidx] = acc + (int)n;
          active = false;
        } else {
          __syncwarp(gmask);  // strip boundary written by lane L-1 is read by lane 0 in the next strip
          setup_strip();
        }
      }
    }
  }
}

}  // namespace trpa
