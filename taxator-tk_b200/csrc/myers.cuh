// Batched global unit-cost edit distance (Levenshtein) over the 5-letter alphabet A,C,G,T,N
// (N==N is a match) -- the integer the reference obtains from
//   -seqan::globalAlignmentScore(short, long, MyersBitVector())
// at core/src/taxonpredictionmodelsequence.hh:150
// (recurrence: core/includes-external/seqan/align/global_alignment_myers_impl.h:62-197).
//
// B200 design (not a translation of the CPU loop):
//  * one pair per sub-warp group of L lanes (L = 1..32, power of two); each lane keeps W 32-bit
//    words of the vertical delta vectors VP/VN *and* the pattern's two bit-planes in registers,
//    so the whole DP column state of a pair (L*W*32 rows) is register resident;
//  * the column-to-column carry chain (adder carry + HP/HN shift-out bits) is broken into an
//    anti-diagonal wavefront: lane l works on text block (t - l) at step t and hands the three
//    boundary bit-vectors of 32 columns to lane l+1 with three __shfl_up_sync per 32 columns;
//  * the per-word adder carry uses the hardware carry flag (add.cc/addc.cc chains);
//  * the score is not tracked per column: D[m][n] = n + popc(VP_final) - popc(VN_final);
//  * patterns longer than L*W*32 rows are processed in horizontal strips; the strip boundary
//    (3 words per 32 text columns) goes through a per-group scratch line in HBM/L2.
// Equality masks are not looked up: Eq = (b0 xnor c0) & (b1 xnor c1) from the bit-planes
// (2 LOP3), with an extra N-plane term only in the HASN instantiation.
#pragma once
#include "common.cuh"

namespace trpa {

// sum[i] = a[i] + b[i] + carry chain; carry-in = (cin + K) overflow, carry-out returned as 0/1.
// K = 0x80000000 when cin carries the flag in bit 31, K = 0xffffffff when cin is 0/1.
template <int N>
struct AddChain;

template <>
struct AddChain<1> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, %3;\n\t"
        "addc.cc.u32 %0, %4, %5;\n\t"
        "addc.u32 %1, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]));
    return co;
  }
};
template <>
struct AddChain<2> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %3, %4;\n\t"
        "addc.cc.u32 %0, %5, %6;\n\t"
        "addc.cc.u32 %1, %7, %8;\n\t"
        "addc.u32 %2, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(s[1]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]), "r"(a[1]), "r"(b[1]));
    return co;
  }
};
template <>
struct AddChain<3> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %4, %5;\n\t"
        "addc.cc.u32 %0, %6, %7;\n\t"
        "addc.cc.u32 %1, %8, %9;\n\t"
        "addc.cc.u32 %2, %10, %11;\n\t"
        "addc.u32 %3, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]), "r"(a[1]), "r"(b[1]), "r"(a[2]), "r"(b[2]));
    return co;
  }
};
template <>
struct AddChain<4> {
  static __device__ __forceinline__ u32 run(u32* s, const u32* a, const u32* b, u32 cin, u32 K) {
    u32 co;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %5, %6;\n\t"
        "addc.cc.u32 %0, %7, %8;\n\t"
        "addc.cc.u32 %1, %9, %10;\n\t"
        "addc.cc.u32 %2, %11, %12;\n\t"
        "addc.cc.u32 %3, %13, %14;\n\t"
        "addc.u32 %4, 0, 0;\n\t}"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(co)
        : "r"(cin), "r"(K), "r"(a[0]), "r"(b[0]), "r"(a[1]), "r"(b[1]), "r"(a[2]), "r"(b[2]),
          "r"(a[3]), "r"(b[3]));
    return co;
  }
};

// chain over W words in chunks of <= 4
template <int W>
__device__ __forceinline__ u32 add_words(u32 (&s)[W], const u32 (&a)[W], const u32 (&b)[W], u32 cin_msb) {
  u32 c = cin_msb;
  u32 K = 0x80000000u;
#pragma unroll
  for (int w0 = 0; w0 < W; w0 += 4) {
    if (W - w0 >= 4) c = AddChain<4>::run(&s[w0], &a[w0], &b[w0], c, K);
    else if (W - w0 == 3) c = AddChain<3>::run(&s[w0], &a[w0], &b[w0], c, K);
    else if (W - w0 == 2) c = AddChain<2>::run(&s[w0], &a[w0], &b[w0], c, K);
    else c = AddChain<1>::run(&s[w0], &a[w0], &b[w0], c, K);
    K = 0xffffffffu;
  }
  return c;  // 0/1
}

// One DP column for the W words a lane owns.
//  c0m/c1m/cNm : text character broadcast to full-width masks (low bit, high bit, is-N)
//  hpc/hnc/cc  : boundary from the rows above, flag in bit 31
//  hpOut/hnOut/cOut : boundary to the rows below, shifted in at bit 0 (column c ends at bit ncols-1-c)
template <int W, bool HASN>
__device__ __forceinline__ void myers_column(u32 (&VP)[W], u32 (&VN)[W], const u32 (&B0)[W],
                                             const u32 (&B1)[W], const u32 (&BN)[W], u32 c0m, u32 c1m,
                                             u32 cNm, u32 hpc, u32 hnc, u32 cc, u32& hpOut, u32& hnOut,
                                             u32& cOut) {
  u32 Eq[W], T[W], S[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    u32 eq = ~(B0[w] ^ c0m) & ~(B1[w] ^ c1m);
    if (HASN) eq = (cNm & BN[w]) | (~cNm & ~BN[w] & eq);
    Eq[w] = eq;
    T[w] = eq & VP[w];  // == (Eq|VN)&VP because VP&VN == 0
  }
  u32 co = add_words<W>(S, VP, T, cc);
  cOut = cOut + cOut + co;
  u32 hpPrev = hpc, hnPrev = hnc;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    u32 X = Eq[w] | VN[w];
    u32 D0 = (S[w] ^ VP[w]) | X;
    u32 HN = VP[w] & D0;
    u32 HP = VN[w] | ~(VP[w] | D0);
    u32 Xh = __funnelshift_l(hpPrev, HP, 1);
    u32 HNs = __funnelshift_l(hnPrev, HN, 1);
    VN[w] = Xh & D0;
    VP[w] = HNs | ~(Xh | D0);
    hpPrev = HP;
    hnPrev = HN;
  }
  hpOut = __funnelshift_l(hpPrev, hpOut, 1);
  hnOut = __funnelshift_l(hnPrev, hnOut, 1);
}

// pairs[first .. first+count) all use the same (L, W) shape.  L lanes per pair.
// scratch: 3*scratch_stride words per sub-warp group slot (only used when a pair needs >1 strip).
// bucket (nullable): device-side {start, count} of this shape inside `pairs` (written by the
// bucketing kernels of the round); then `count` is only the host's upper bound used for the grid.
template <int W, bool HASN>
__global__ void __launch_bounds__(128)
myers_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
             const uint2* __restrict__ planes, const u32* __restrict__ nplane, int* __restrict__ out,
             int L, u32* __restrict__ scratch, u32 scratch_stride, const uint2* __restrict__ bucket) {
  if (bucket) {
    const uint2 bk = *bucket;
    pairs += bk.x;
    count = bk.y;
    if ((blockIdx.x * blockDim.x >> 5) * (32 / L) >= count) return;
  }
  const u32 lane = threadIdx.x & 31;
  const u32 G = 32 / L;                 // pairs per warp
  const u32 g = lane / L;               // group in warp
  const u32 sl = lane - g * L;          // lane in group
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 slot = warp_gid * G + g;
  const bool have = slot < count;

  u32 m = 0, n = 0, pw = 0, tw = 0, oidx = 0;
  if (have) {
    PairDesc pd = pairs[slot];
    SeqDesc A = seqs[pd.a], B = seqs[pd.b];
    // pattern = shorter (hh:142-147); distance is symmetric so ties do not matter
    if (A.len < B.len) { m = A.len; pw = A.woff; n = B.len; tw = B.woff; }
    else               { m = B.len; pw = B.woff; n = A.len; tw = A.woff; }
    oidx = pd.out;
  }
  const u32 mwords = (m + 31) >> 5;
  const u32 nblk = (n + 31) >> 5;
  const u32 rows_per_strip = (u32)L * W;                    // words
  const u32 nstrips = (mwords + rows_per_strip - 1) / rows_per_strip;

  // warp-uniform loop bounds
  u32 maxblk = nblk, maxstrips = nstrips;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    maxblk = max(maxblk, __shfl_xor_sync(0xffffffffu, maxblk, o));
    maxstrips = max(maxstrips, __shfl_xor_sync(0xffffffffu, maxstrips, o));
  }
  const u32 maxsteps = maxblk + L - 1;
  u32* my_scratch = scratch ? scratch + (size_t)slot * 3 * scratch_stride : nullptr;

  int score = (int)n;  // D[0][n]; only sl==0 keeps it

  for (u32 s = 0; s < maxstrips; ++s) {
    const bool strip_on = s < nstrips;
    u32 VP[W], VN[W], B0[W], B1[W], BN[W];
    const u32 kbase = (s * L + sl) * W;
#pragma unroll
    for (int w = 0; w < W; ++w) {
      const u32 k = kbase + w;
      uint2 p = make_uint2(0u, 0u);
      u32 pn = 0;
      if (strip_on && k < mwords) {
        p = planes[pw + k];
        if (HASN) pn = nplane[pw + k];
      }
      B0[w] = p.x; B1[w] = p.y; BN[w] = pn;
      VP[w] = 0xffffffffu; VN[w] = 0u;
    }
    u32 hpOut = 0, hnOut = 0, cOut = 0;
    const bool last_lane_writes = (sl == (u32)L - 1) && (s + 1 < nstrips);

    for (u32 t = 0; t < maxsteps; ++t) {
      u32 hpIn = __shfl_up_sync(0xffffffffu, hpOut, 1, L);
      u32 hnIn = __shfl_up_sync(0xffffffffu, hnOut, 1, L);
      u32 cIn = __shfl_up_sync(0xffffffffu, cOut, 1, L);
      const int blk = (int)t - (int)sl;
      const bool on = strip_on && blk >= 0 && blk < (int)nblk;
      if (on) {
        const u32 ncols = min(32u, n - 32u * (u32)blk);
        if (sl == 0) {
          if (s == 0) { hpIn = 0xffffffffu; hnIn = 0u; cIn = 0u; }
          else {
            hpIn = my_scratch[3 * blk + 0];
            hnIn = my_scratch[3 * blk + 1];
            cIn = my_scratch[3 * blk + 2];
          }
        }
        uint2 tx = planes[tw + blk];
        u32 t0 = __brev(tx.x), t1 = __brev(tx.y), tN = 0;
        if (HASN) tN = __brev(nplane[tw + blk]);
        // boundary words hold column c at bit ncols-1-c; bring column 0 to bit 31
        u32 hpc = hpIn << (32 - ncols), hnc = hnIn << (32 - ncols), cc = cIn << (32 - ncols);
        if (ncols == 32) { hpc = hpIn; hnc = hnIn; cc = cIn; }
        hpOut = 0; hnOut = 0; cOut = 0;
        if (ncols == 32) {
#pragma unroll 4
          for (int c = 0; c < 32; ++c) {
            u32 c0m = (u32)((int)t0 >> 31), c1m = (u32)((int)t1 >> 31), cNm = (u32)((int)tN >> 31);
            myers_column<W, HASN>(VP, VN, B0, B1, BN, c0m, c1m, cNm, hpc, hnc, cc, hpOut, hnOut, cOut);
            t0 <<= 1; t1 <<= 1; tN <<= 1; hpc <<= 1; hnc <<= 1; cc <<= 1;
          }
        } else {
#pragma unroll 1
          for (u32 c = 0; c < ncols; ++c) {
            u32 c0m = (u32)((int)t0 >> 31), c1m = (u32)((int)t1 >> 31), cNm = (u32)((int)tN >> 31);
            myers_column<W, HASN>(VP, VN, B0, B1, BN, c0m, c1m, cNm, hpc, hnc, cc, hpOut, hnOut, cOut);
            t0 <<= 1; t1 <<= 1; tN <<= 1; hpc <<= 1; hnc <<= 1; cc <<= 1;
          }
        }
        if (last_lane_writes) {
          my_scratch[3 * blk + 0] = hpOut;
          my_scratch[3 * blk + 1] = hnOut;
          my_scratch[3 * blk + 2] = cOut;
        }
      }
    }
    // vertical deltas of the last column for the rows this lane owns
    if (strip_on) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const u32 k = kbase + w;
        u32 valid = 0;
        if (k < mwords) {
          const u32 rem = m - 32u * k;  // >= 1
          valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
        }
        score += __popc(VP[w] & valid) - __popc(VN[w] & valid);
      }
    }
    __syncwarp();  // strip boundary written by lane L-1 is read by lane 0 next strip
  }
  // reduce over the group; every lane started from n, so subtract the duplicates
  int acc = score - (int)n;
  for (int o = L >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (have && sl == 0) out[oidx] = acc + (int)n;
}

}  // namespace trpa
