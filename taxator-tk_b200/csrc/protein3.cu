// Protein kernel for all pairs of up to 1000 x 1000 residues: the result of getAlignmentProtein
// (core/src/taxonpredictionmodelsequence.hh:173-242: BLOSUM62, linear gap -1, SeqAn tie order diagonal >= vertical >=
// horizontal, dp_formula_linear.h:62-105 / dp_formula.h:152-163, traced alignment length hh:215-228).
//
// Each DP cell is ONE packed 32-bit integer
//     P = score * 2^13 + priority * 2^11 + #gap columns on the traced path
// so that a single signed max3 picks the best score and, on ties, SeqAn's priority (diagonal 2, vertical 1, horizontal
// 0); the priority field is cleared before the value is stored.  The traced length follows from the gap count:
// len = (|A| + |B| + #gaps) / 2.  Substitution scores come from a per-lane "query profile" in shared memory,
// profile[b][c/4][lane][c%4] = 2*BLOSUM62(a_c, b) + 9 (uint8; bank == lane for every access: one conflict-free LDS.U8 per
// cell), so the diagonal candidate is one multiply-add: D = profile * 2^12 + diag (= score + sub, priority 2, plus a
// constant 2^15 that is carried as a per-row bias i*2^15 on every stored value, so the table stays unsigned and needs no
// sign extension).  Valid while the gap count fits 11 bits: |A| + |B| <= 2047, i.e. both sequences <= 1000 residues
// (longer pairs: the 32-bit-per-field kernel of protein.cu).  Per cell: 1 LDS.U8 + 1 IMAD + 2 VIADDMNMX + 1 LOP3.
//
// Shape of the wavefront (round 2; round 1 ran one warp or half a warp per pair, two rows per step):
//   * EIGHT lanes per pair (four pairs per warp) for |A| <= 320, up to 40 columns per lane: the ramp of the wavefront is 7
//     steps instead of 15 or 31; 16 lanes for 321-640 columns, a whole warp for 641-1000;
//   * R = 4 rows per step, walked in SKEWED order -- cells (0,k), (1,k-1), (2,k-2), (3,k-3) are adjacent in the
//     instruction stream -- so every lane carries four independent max chains and the per-step bookkeeping (shuffles,
//     predicates, row addresses) is paid once per R * C cells (160 at C = 40);
//   * the profile of a pair is built from 16-byte rows of the (symmetric) substitution table with byte transposes
//     (8 PRMT + 4 STS.32 per 16 entries) and keeps rows only for the residues that occur in the stores;
//   * every third column computes both additions as IMAD on the fma pipe and takes one VIMNMX3 (2 instead of 3 alu-pipe
//     instructions, one more issue slot): the alu pipe and the issue slots are loaded evenly.
#include "common.cuh"
#include "launch.h"
#include "blosum62_table.h"
#include <cstdlib>

namespace trpa {

namespace {

constexpr int kQMaxLen = 1000;    // the gap count must fit 11 bits: |A| + |B| <= 2047
// a group of LANES lanes takes the pairs with lo(LANES) < |A| <= hi(LANES): 8 lanes up to 320 columns (40 per lane),
// 16 lanes up to 640, 32 lanes up to 1000
__host__ __device__ constexpr int q_lo(int lanes) { return lanes == 8 ? 0 : (lanes == 16 ? 320 : 640); }
__host__ __device__ constexpr int q_hi(int lanes) { return lanes == 8 ? 320 : (lanes == 16 ? 640 : kQMaxLen); }
constexpr int kQWarps = 1;        // warps per CTA (shared memory decides how many pairs are resident: finest granularity)

constexpr int SH = 13;
constexpr int PRIO_MASK = 3 << 11;
constexpr int ROWBIAS = 8 << (SH - 1);                        // what the +8 of the unsigned profile adds per row
constexpr int CV = -(1 << SH) + (1 << 11) + 1 + ROWBIAS;      // vertical: score-1, priority 1, +1 gap, next row
constexpr int CH = -(1 << SH) + 1;                            // horizontal: score-1, priority 0, +1 gap
constexpr int BND = -(1 << SH) + 1;                           // boundary(k) = k * BND: score -k after k gap columns

// e[r][b] = 2 * BLOSUM62(r, b) + 9 (the unsigned profile entry); rows / columns 27..31 are 0 = the
// profile of a padding column
__device__ __align__(16) uint8_t c_prof_e[32][32];   // read once per CTA with two coalesced 16-byte loads per lane
bool g_loaded3[16] = {false};

cudaError_t ensure_table3() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 16 && g_loaded3[dev]) return cudaSuccess;
  uint8_t h[32][32];
  for (int r = 0; r < 32; ++r)
    for (int b = 0; b < 32; ++b) h[r][b] = (r < 27 && b < 27) ? (uint8_t)(2 * TRPA_BLOSUM62[r][b] + 9) : (uint8_t)0;
  cudaError_t e = cudaMemcpyToSymbol(c_prof_e, h, sizeof(h));
  if (e == cudaSuccess && dev >= 0 && dev < 16) g_loaded3[dev] = true;
  return e;
}

// columns per lane the kernel is instantiated for: what `cols` columns per lane are rounded up to (the host sizes the
// shared-memory profile with the same function)
__host__ __device__ constexpr int q_round_cols(int lanes, int cols) {
  return lanes == 8 ? (cols <= 8 ? 8 : cols <= 16 ? 16 : cols <= 24 ? 24 : cols <= 32 ? 32 : cols <= 36 ? 36 : cols <= 38 ? 38 : 40)
       : lanes == 16 ? (cols <= 24 ? 24 : cols <= 32 ? 32 : cols <= 36 ? 36 : 40)
                     : (cols <= 24 ? 24 : cols <= 28 ? 28 : 32);
}
constexpr bool q_round_cols_covers() {   // every admissible column count is rounded UP to an instantiated one
  for (int lanes = 8; lanes <= 32; lanes *= 2)
    for (int n = q_lo(lanes) + 1; n <= q_hi(lanes); ++n) {
      const int cols = (n + lanes - 1) / lanes, r = q_round_cols(lanes, cols);
      if (r < cols || r > 40 || (r + 3) / 4 > 10) return false;
    }
  return true;
}
static_assert(q_round_cols_covers(), "protein3: column rounding does not cover the lane width's range");
template <int LANES>
__device__ __forceinline__ bool q_takes(int n, int m) { return n > q_lo(LANES) && n <= q_hi(LANES) && m > 0 && m <= kQMaxLen; }

__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }

__device__ __forceinline__ uint4 lds128(u32 addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(u32 addr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// One pair on a group of LANES lanes, C columns per lane, RIGHT aligned (the first LANES * C - n columns of the low
// lanes are padding: profile 0 (= substitution -4.5) for every residue and row-0 value 0, so neither the diagonal nor the
// horizontal candidate can win there and the vertical candidate reproduces the left boundary column value (-i, i gaps)
// in every padding cell -- the first real column sees exactly the boundary it needs; no per-cell predicates).  All
// groups of a warp run the same number of steps (the longest B among them decides).
template <int LANES, int C, int R, int ALT>
__device__ __forceinline__ void protein3_run(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int m,
                                             bool mine, u32 prof_sa, u32 t2_sa, u32 lp, int steps, u32 mask, int one,
                                             int2* __restrict__ out2, u32 oidx) {
  constexpr int CQ = (C + 3) / 4;   // the profile keeps whole column quads; columns >= C of the last quad stay unused
  const int pad = LANES * C - n;                  // leading padding columns
  const int v1 = (int)lp * C - pad;               // index of this lane's first column (may be < 0)

  // ---- profile[bb][q][lane][k] = e(a[v1 + 4q + k], bb): 16-byte rows of the symmetric table, transposed bytewise
  // (fully unrolled: the residue loads of all quads are issued up front)
#pragma unroll
  for (int q = 0; q < CQ; ++q) {
    u32 x[4][7];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v = v1 + 4 * q + k;
      const u32 r = (4 * q + k < C && v >= 0 && mine) ? (u32)a[v] : 27u;
      const uint4 lo = lds128(t2_sa + r * 32u), hi = lds128(t2_sa + r * 32u + 16u);
      x[k][0] = lo.x; x[k][1] = lo.y; x[k][2] = lo.z; x[k][3] = lo.w;
      x[k][4] = hi.x; x[k][5] = hi.y; x[k][6] = hi.z;
    }
    const u32 dst = prof_sa + (u32)q * 128u;
    // row of residue bb = its rank among the residues of the mask (uniform: the mask is a kernel parameter)
    auto put = [&](int bb, u32 w) {
      if ((mask >> bb) & 1u) sts32(dst + (u32)__popc(mask & ((1u << bb) - 1u)) * (u32)(CQ * 128), w);
    };
#pragma unroll
    for (int g = 0; g < 7; ++g) {
      const u32 p01 = __byte_perm(x[0][g], x[1][g], 0x5140), p23 = __byte_perm(x[2][g], x[3][g], 0x5140);
      const u32 q01 = __byte_perm(x[0][g], x[1][g], 0x7362), q23 = __byte_perm(x[2][g], x[3][g], 0x7362);
      put(4 * g + 0, __byte_perm(p01, p23, 0x5410));
      put(4 * g + 1, __byte_perm(p01, p23, 0x7632));
      put(4 * g + 2, __byte_perm(q01, q23, 0x5410));
      if (4 * g + 3 < 27) put(4 * g + 3, __byte_perm(q01, q23, 0x7632));
    }
  }
  int up[C];
#pragma unroll
  for (int c = 0; c < C; ++c) up[c] = (v1 + c >= 0) ? (v1 + c + 1) * BND : 0;   // row 0; padding: cell(0, 0)
  __syncwarp();

  // Residues of the rows of block j (rows R(j-1)+1 .. Rj), fetched one step ahead, as profile rows (rank of the residue
  // in the mask).  Row indices are clamped into [1, m]; a residue outside the mask is remembered and trapped on at the
  // end (no branch on a value that has only just been requested from memory).
  u32 bad = 0;
  auto fetch_rows = [&](int j, u32 (&br)[R]) {
    const int jj = j < 1 ? 1 : j;
    const int i0n = R * (jj - 1) + 1;
    u32 x[R];
    if (i0n + R - 1 <= m) {
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = b[i0n - 1 + r];
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) { const int i = i0n + r > m ? m : i0n + r; x[r] = m > 0 ? (u32)b[i - 1] : 0u; }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      bad |= ~mask >> x[r];
      br[r] = (u32)__popc(mask & ((1u << x[r]) - 1u));
    }
  };
  int last[R];
#pragma unroll
  for (int r = 0; r < R; ++r) last[r] = 0;
  int res = 0;
  int dcarry = v1 > 0 ? v1 * BND : 0;             // cell(0, first column - 1)
  u32 bcur[R];
  fetch_rows(1 - (int)lp, bcur);
  for (int t = 1; t <= steps; ++t) {
    int recv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) recv[r] = __shfl_up_sync(0xffffffffu, last[r], 1, LANES);
    const int j = t - (int)lp;                    // this lane's row block at step t
    const int i0 = R * (j - 1) + 1;
    u32 bnext[R];
    fetch_rows(j + 1, bnext);
    if (j >= 1 && i0 <= m) {
      int left[R], dg[R];
      u32 prow[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        prow[r] = prof_sa + bcur[r] * (u32)(CQ * 128);
        left[r] = lp == 0 ? (i0 + r) * (BND + ROWBIAS) : recv[r];   // column 0: boundary(i) + the row bias
      }
      dg[0] = dcarry;                             // diagonal input of a row's first cell = left boundary of the row above
#pragma unroll
      for (int r = 1; r < R; ++r) dg[r] = left[r - 1];
      dcarry = left[R - 1];
      // skewed walk: R independent chains next to each other in the instruction stream
#pragma unroll
      for (int k = 0; k < C + R - 1; ++k) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int c = k - r;
          if (c >= 0 && c < C) {
            int e;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(e) : "r"(prow[r] + (u32)((c >> 2) * 128 + (c & 3))));
            const int D = e * (1 << (SH - 1)) + dg[r];   // score + sub, priority 2, gaps unchanged, row bias + 1
            int cell;
            if (ALT > 0 && c % ALT == ALT - 1) {
              // every ALT-th column: both additions on the fma pipe (IMAD with a multiplier ptxas cannot fold) and one
              // three-way maximum -- 2 instead of 3 alu-pipe instructions for this cell, one more issue slot
              int V, H;
              asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(V) : "r"(up[c]), "r"(one), "r"(CV));
              asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(H) : "r"(left[r]), "r"(one), "r"(CH));
              cell = __vimax3_s32(D, V, H) & ~PRIO_MASK;
            } else {
              const int V = up[c] + CV;
              const int H = left[r] + CH;
              cell = max3i(D, V, H) & ~PRIO_MASK;
            }
            dg[r] = up[c];
            up[c] = cell;
            left[r] = cell;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        last[r] = left[r];
        if (lp == LANES - 1 && i0 + r == m) res = left[r];
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) bcur[r] = bnext[r];
  }
  if (mine && (bad & 1u)) __trap();   // a residue the stores were said not to contain: fail loudly
  if (lp == LANES - 1 && mine) {
    const int P = res - m * ROWBIAS;        // remove the per-row bias of the last row
    const int score = P >> SH;              // arithmetic shift: floor, low fields are non-negative
    const int gaps = P & 0x7ff;
    out2[oidx] = make_int2(score, (n + m - gaps) / 2);   // |A| + |B| = 2*#diag + #gaps
  }
}

template <int LANES, int R, int ALT>
__global__ void __launch_bounds__(32 * kQWarps, 1)
protein3_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
                const uint8_t* __restrict__ residues, int2* __restrict__ out2, u32 cq_cap, u32 mask, int one) {
  // cq_cap: column quads per lane the shared-memory profile is sized for (the launch's longest sequence decides);
  // mask: residue ordinals that may occur (one profile row each)
  extern __shared__ __align__(16) unsigned char smem3[];
  unsigned char* t2 = smem3;                                      // e[32][32]
  unsigned char* prof_all = smem3 + 1024;                         // [warp][bb][q][lane][4]
  constexpr u32 G = 32u / LANES;   // pairs per warp
  const u32 lane = threadIdx.x & 31, lp = lane & (u32)(LANES - 1);
  const u32 warp_in_cta = threadIdx.x >> 5;
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 pidx = warp_gid * G + lane / (u32)LANES;
  int n = 0, m = 0;
  const uint8_t* a = residues;
  const uint8_t* b = residues;
  u32 oidx = 0;
  bool mine = false;   // this group's pair is handled here
  if (pidx < count) {
    const PairDesc pd = pairs[pidx];
    const SeqDesc A = seqs[pd.a], B = seqs[pd.b];
    if (q_takes<LANES>((int)A.len, (int)B.len)) {
      mine = true; n = (int)A.len; m = (int)B.len; oidx = pd.out;
      a = residues + A.woff; b = residues + B.woff;
    } else if (LANES == 8 && lp == 0 && (A.len == 0 || B.len == 0) && A.len <= (u32)kQMaxLen && B.len <= (u32)kQMaxLen) {
      out2[pd.out] = make_int2(-(int)(A.len + B.len), 0);   // an empty side: all gaps, nothing traced diagonally
    }
  }
  // (a CTA is one warp: a launch that walks a pair list without pairs of its column range costs a few loads per CTA)
  static_assert(kQWarps == 1, "the early exit below is per warp");
  if (!__any_sync(0xffffffffu, mine)) return;
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
    reinterpret_cast<uint4*>(t2)[i] = reinterpret_cast<const uint4*>(&c_prof_e[0][0])[i];
  __syncthreads();
  // all groups use the lane width the widest pair needs and run as many steps as the longest one
  int nmax = n, mmax = m;
#pragma unroll
  for (int o = LANES; o < 32; o <<= 1) {
    nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
    mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
  }
  const int cols = q_round_cols(LANES, (nmax + LANES - 1) / LANES);
  if ((cols + 3) / 4 > (int)cq_cap) __trap();   // the launcher sized the profile from a wrong max_len: fail loudly
  const int steps = (mmax + R - 1) / R + (LANES - 1);
  const u32 prof_sa = (u32)__cvta_generic_to_shared(prof_all) + (warp_in_cta * (u32)__popc(mask) * cq_cap * 32u + lane) * 4u;
  const u32 t2_sa = (u32)__cvta_generic_to_shared(t2);
#define TRPA_P3_RUN(CC) protein3_run<LANES, CC, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx)
  if (LANES == 8) {
    if (cols == 8) TRPA_P3_RUN(8);
    else if (cols == 16) TRPA_P3_RUN(16);
    else if (cols == 24) TRPA_P3_RUN(24);
    else if (cols == 32) TRPA_P3_RUN(32);
    else if (cols == 36) TRPA_P3_RUN(36);
    else if (cols == 38) TRPA_P3_RUN(38);
    else TRPA_P3_RUN(40);
  } else if (LANES == 16) {   // 321 .. 640 columns: 21 .. 40 per lane
    if (cols == 24) TRPA_P3_RUN(24);
    else if (cols == 32) TRPA_P3_RUN(32);
    else if (cols == 36) TRPA_P3_RUN(36);
    else TRPA_P3_RUN(40);
  } else {                    // 641 .. 1000 columns: 21 .. 32 per lane
    if (cols == 24) TRPA_P3_RUN(24);
    else if (cols == 28) TRPA_P3_RUN(28);
    else TRPA_P3_RUN(32);
  }
#undef TRPA_P3_RUN
}

}  // namespace

bool protein3_takes(u32 n, u32 m) { return n > 0 && m > 0 && n <= (u32)kQMaxLen && m <= (u32)kQMaxLen; }

// Cell mix of the build: four rows per step, every 3rd column in the fma-heavy form.  Measured on C3 against the other
// mixes (gpurun_out/r2_34, r2_35; alignment time per step): plain cells 34.0 ms, every 4th column 33.2, every 3rd 33.0,
// every 2nd 33.5; two rows per step with plain cells 34.8 ms.
namespace {
constexpr int kRows = 4, kAlt = 3;

template <int LANES>
cudaError_t launch_lanes(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues, int2* out2,
                         u32 max_len, u32 mask, cudaStream_t stream) {
  const u32 nrows = (u32)__builtin_popcount(mask);
  const u32 longest = (max_len == 0 || max_len > (u32)q_hi(LANES)) ? (u32)q_hi(LANES) : max_len;
  // profile capacity: the column count the kernel rounds the launch's widest pair up to, in whole quads
  const u32 cq = ((u32)q_round_cols(LANES, (int)((longest + LANES - 1) / LANES)) + 3u) / 4u;
  const size_t smem = 1024 + (size_t)kQWarps * nrows * cq * 128;
  auto kern = protein3_kernel<LANES, kRows, kAlt>;
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr = attr_set[dev & 15];   // the attributes are per device
  if (!attr) {
    const int cap = (int)(1024 + kQWarps * 27 * 10 * 128);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  constexpr u32 G = 32u / LANES;
  const u32 warps = (count + G - 1u) / G;
  const u32 blocks = (warps + kQWarps - 1) / kQWarps;
  kern<<<blocks, 32 * kQWarps, smem, stream>>>(pairs, count, seqs, residues, out2, cq, mask, 1);
  return cudaGetLastError();
}
}  // namespace

// max_len: longest staged sequence of the launch (0 = unknown) -- decides which lane widths are launched at all and
// sizes the shared-memory profile together with aa_mask (residue ordinals that may occur, 0 = all 27): 20 residues x
// 40 columns per lane = 25.6 KB per warp, eight warps per SM; all 27: 34.6 KB, six warps.  Every launch walks the whole
// pair list and takes the pairs of its column range.
cudaError_t launch_protein3(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues, int2* out2,
                            u32 max_len, u32 aa_mask, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  cudaError_t e = ensure_table3();
  if (e != cudaSuccess) return e;
  u32 mask = aa_mask & 0x7ffffffu;
  if (mask == 0) mask = 0x7ffffffu;
  e = launch_lanes<8>(pairs, count, seqs, residues, out2, max_len, mask, stream);
  if (e != cudaSuccess) return e;
  if (max_len == 0 || max_len > (u32)q_lo(16)) {
    e = launch_lanes<16>(pairs, count, seqs, residues, out2, max_len, mask, stream);
    if (e != cudaSuccess) return e;
  }
  if (max_len == 0 || max_len > (u32)q_lo(32)) {
    e = launch_lanes<32>(pairs, count, seqs, residues, out2, max_len, mask, stream);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace trpa
