// Third-generation protein kernel for the short pairs (|A| <= 320 columns, |B| <= 1000 rows) that make up the
// protein workloads: same result as protein2.cu / getAlignmentProtein (core/src/taxonpredictionmodelsequence.hh:173-242:
// BLOSUM62, linear gap -1, SeqAn tie order diagonal >= vertical >= horizontal, traced alignment length), same packed
// 32-bit cell (score * 2^13 + priority * 2^11 + #gap columns of the traced path, see protein2.cu) and the same per-lane
// query profile in shared memory (one conflict-free LDS.U8 per cell).  What changes is the shape of the wavefront:
//   * EIGHT lanes per pair (four pairs per warp), up to 40 columns per lane: the ramp of the wavefront is 7 steps
//     instead of 15 (half-warp kernel) or 31 (one warp per pair);
//   * R rows per step (2 or 4), walked in SKEWED order -- cells (0,k), (1,k-1), (2,k-2), ... are adjacent in the
//     instruction stream -- so every lane carries R independent max chains and the per-step bookkeeping (shuffles,
//     predicates, row addresses) is paid once per R * C cells (160 at R = 4, C = 40; the half-warp kernel: 40);
//   * the profile of a pair is built from 16-byte rows of the (symmetric) substitution table with byte transposes
//     (8 PRMT + 4 STS.32 per 16 entries) instead of one table look-up per entry.
#include "common.cuh"
#include "launch.h"
#include "blosum62_table.h"
#include <cstdlib>

namespace trpa {

namespace {

constexpr int kQLanes = 8;        // lanes per pair
constexpr int kQMaxCols = 320;    // 8 lanes x 40 columns
constexpr int kQMaxRows = 1000;   // the gap count must fit 11 bits: |A| + |B| <= 2047
constexpr int kQWarps = 1;        // warps per CTA (shared memory decides how many pairs are resident: finest granularity)

constexpr int SH = 13;
constexpr int PRIO_MASK = 3 << 11;
constexpr int ROWBIAS = 8 << (SH - 1);                        // what the +8 of the unsigned profile adds per row
constexpr int CV = -(1 << SH) + (1 << 11) + 1 + ROWBIAS;      // vertical: score-1, priority 1, +1 gap, next row
constexpr int CH = -(1 << SH) + 1;                            // horizontal: score-1, priority 0, +1 gap
constexpr int BND = -(1 << SH) + 1;                           // boundary(k) = k * BND: score -k after k gap columns

// e[r][b] = 2 * BLOSUM62(r, b) + 9 (the unsigned profile entry of protein2.cu); rows / columns 27..31 are 0 = the
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

__device__ __forceinline__ bool q_takes(int n, int m) { return n > 0 && m > 0 && n <= kQMaxCols && m <= kQMaxRows; }

__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }

__device__ __forceinline__ uint4 lds128(u32 addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(u32 addr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// One pair on the 8 lanes of a quarter warp, C columns per lane, RIGHT aligned (the first 8 * C - n columns of the low
// lanes are padding: profile 0, row-0 value 0 -- the vertical candidate reproduces the left boundary column there, see
// protein2.cu).  All four quarters of a warp run the same number of steps (the longest B of the four decides).
template <int C, int R, int ALT>
__device__ __forceinline__ void protein3_run(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int m,
                                             bool mine, u32 prof_sa, u32 t2_sa, u32 lp, int steps, u32 mask, int one,
                                             int2* __restrict__ out2, u32 oidx) {
  constexpr int CQ = (C + 3) / 4;   // the profile keeps whole column quads; columns >= C of the last quad stay unused
  const int pad = kQLanes * C - n;                // leading padding columns
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
    for (int r = 0; r < R; ++r) recv[r] = __shfl_up_sync(0xffffffffu, last[r], 1, kQLanes);
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
        if (lp == kQLanes - 1 && i0 + r == m) res = left[r];
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) bcur[r] = bnext[r];
  }
  if (mine && (bad & 1u)) __trap();   // a residue the stores were said not to contain: fail loudly
  if (lp == kQLanes - 1 && mine) {
    const int P = res - m * ROWBIAS;        // remove the per-row bias of the last row
    const int score = P >> SH;              // arithmetic shift: floor, low fields are non-negative
    const int gaps = P & 0x7ff;
    out2[oidx] = make_int2(score, (n + m - gaps) / 2);   // |A| + |B| = 2*#diag + #gaps
  }
}

template <int R, int ALT>
__global__ void __launch_bounds__(32 * kQWarps, 1)
protein3_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
                const uint8_t* __restrict__ residues, int2* __restrict__ out2, u32 cq_cap, u32 mask, int one) {
  // cq_cap: column quads per lane the shared-memory profile is sized for (the launch's longest sequence decides);
  // mask: residue ordinals that may occur (one profile row each)
  extern __shared__ __align__(16) unsigned char smem3[];
  unsigned char* t2 = smem3;                                      // e[32][32]
  unsigned char* prof_all = smem3 + 1024;                         // [warp][bb][q][lane][4]
  for (int i = threadIdx.x; i < 64; i += blockDim.x)
    reinterpret_cast<uint4*>(t2)[i] = reinterpret_cast<const uint4*>(&c_prof_e[0][0])[i];
  __syncthreads();
  const u32 lane = threadIdx.x & 31, lp = lane & 7u;
  const u32 warp_in_cta = threadIdx.x >> 5;
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 pidx = warp_gid * 4u + (lane >> 3);
  int n = 0, m = 0;
  const uint8_t* a = residues;
  const uint8_t* b = residues;
  u32 oidx = 0;
  bool mine = false;   // this quarter's pair is handled here
  if (pidx < count) {
    const PairDesc pd = pairs[pidx];
    const SeqDesc A = seqs[pd.a], B = seqs[pd.b];
    if (q_takes((int)A.len, (int)B.len)) {
      mine = true; n = (int)A.len; m = (int)B.len; oidx = pd.out;
      a = residues + A.woff; b = residues + B.woff;
    }
  }
  if (!__any_sync(0xffffffffu, mine)) return;
  // all quarters use the lane width the widest pair needs and run as many steps as the longest one
  int nmax = n, mmax = m;
  nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8)); nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
  mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, 8)); mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, 16));
  const int cols = (nmax + kQLanes - 1) / kQLanes;
  if ((cols + 3) / 4 > (int)cq_cap) __trap();   // the launcher sized the profile from a wrong max_len: fail loudly
  const int steps = (mmax + R - 1) / R + (kQLanes - 1);
  const u32 prof_sa = (u32)__cvta_generic_to_shared(prof_all) + (warp_in_cta * (u32)__popc(mask) * cq_cap * 32u + lane) * 4u;
  const u32 t2_sa = (u32)__cvta_generic_to_shared(t2);
  if (cols <= 8) protein3_run<8, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx);
  else if (cols <= 16) protein3_run<16, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx);
  else if (cols <= 24) protein3_run<24, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx);
  else if (cols <= 32) protein3_run<32, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx);
  else if (cols <= 36) protein3_run<36, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx);
  else if (cols <= 38) protein3_run<38, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx);
  else protein3_run<40, R, ALT>(a, b, n, m, mine, prof_sa, t2_sa, lp, steps, mask, one, out2, oidx);
}

// 0 (default): four rows per step, every 3rd column in the fma-heavy form; A/B hooks: TRPA_PROTEIN_ROWS=2 (1: two rows,
// plain cells), TRPA_PROTEIN_ALT=4 / 0 / 2 (2 / 3 / 4: four rows with every 4th / no / every 2nd column fma-heavy).
// Measured on C3 (gpurun_out/r2_34): plain 34.0 ms of alignment per step, every 4th 33.2, every 3rd 33.0.
int protein3_variant() {
  static int v = -1;
  if (v < 0) {
    v = 0;
    if (const char* e = getenv("TRPA_PROTEIN_ROWS")) if (e[0] == '2') v = 1;
    if (const char* e = getenv("TRPA_PROTEIN_ALT")) { if (e[0] == '4') v = 2; else if (e[0] == '0') v = 3; else if (e[0] == '2') v = 4; }
  }
  return v;
}

}  // namespace

bool protein3_takes(u32 n, u32 m) { return n > 0 && m > 0 && n <= (u32)kQMaxCols && m <= (u32)kQMaxRows; }

// max_len: longest staged sequence of the launch (0 = unknown) -- sizes the shared-memory profile together with
// aa_mask (residue ordinals that may occur, 0 = all 27): 20 residues x 40 columns per lane = 25.6 KB per warp, eight
// warps = 32 pairs per SM; all 27: 34.6 KB, six warps
cudaError_t launch_protein3(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues, int2* out2,
                            u32 max_len, u32 aa_mask, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  cudaError_t e = ensure_table3();
  if (e != cudaSuccess) return e;
  u32 mask = aa_mask & 0x7ffffffu;
  if (mask == 0) mask = 0x7ffffffu;
  const u32 nrows = (u32)__builtin_popcount(mask);
  const u32 longest = (max_len == 0 || max_len > (u32)kQMaxCols) ? (u32)kQMaxCols : max_len;
  const u32 cols = (longest + kQLanes - 1) / kQLanes;
  const u32 cq = cols <= 8 ? 2u : (cols <= 16 ? 4u : (cols <= 24 ? 6u : (cols <= 32 ? 8u : (cols <= 36 ? 9u : 10u))));
  const size_t smem = 1024 + (size_t)kQWarps * nrows * cq * 128;
  typedef void (*Kern)(const PairDesc*, u32, const SeqDesc*, const uint8_t*, int2*, u32, u32, int);
  static const Kern kerns[5] = {protein3_kernel<4, 3>, protein3_kernel<2, 0>, protein3_kernel<4, 4>, protein3_kernel<4, 0>, protein3_kernel<4, 2>};
  const int variant = protein3_variant();
  const Kern kern = kerns[variant];
  static bool attr_set[16][5] = {{false}};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr = attr_set[dev & 15][variant];   // the attributes are per device
  if (!attr) {
    const int cap = (int)(1024 + kQWarps * 27 * 10 * 128);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const u32 warps = (count + 3u) / 4u;
  const u32 blocks = (warps + kQWarps - 1) / kQWarps;
  kern<<<blocks, 32 * kQWarps, smem, stream>>>(pairs, count, seqs, residues, out2, cq, mask, 1);
  return cudaGetLastError();
}

}  // namespace trpa
