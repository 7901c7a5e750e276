// Second-generation protein kernel: same result as protein.cu / getAlignmentProtein
// (core/src/taxonpredictionmodelsequence.hh:173-242: BLOSUM62, linear gap -1, SeqAn tie order
// diagonal >= vertical >= horizontal, traced alignment length), with the per-cell work cut to
//   1 LDS.U8 + 1 multiply-add + 2 fused add-max (VIADDMNMX) + 1 LOP.
// Each DP cell is ONE packed 32-bit integer
//     P = score * 2^13 + priority * 2^11 + #gap columns on the traced path
// so that a single signed max3 picks the best score and, on ties, SeqAn's priority (diagonal 2,
// vertical 1, horizontal 0); the priority field is cleared before the value is stored.  The traced
// length follows from the gap count: len = (|A| + |B| + #gaps) / 2.  Substitution scores come from a
// per-lane "query profile" in shared memory, profile[b][c/4][lane][c%4] = 2*BLOSUM62(a_c, b) + 9 (uint8;
// bank == lane for every access, built with one STS.32 per 4 columns), so the diagonal candidate is one
// multiply-add: D = profile * 2^12 + diag  (= score + sub, priority 2, plus a constant 2^15 that is
// carried as a per-row bias i*2^15 on every stored value, so the table stays unsigned and needs no
// sign extension).  Valid while the gap count fits 11 bits: |A| + |B| <= 2047, i.e. both sequences
// <= 1000 residues (longer pairs use protein.cu).
#include "common.cuh"
#include "launch.h"
#include "blosum62_table.h"
#include <cstdlib>

namespace trpa {

constexpr int kC2Max = 16;
constexpr int kP2MaxLen = 1000;
constexpr int kP2Warps = 4;
constexpr int kCQ = kC2Max / 4;
constexpr size_t kP2ProfBytes = (size_t)kP2Warps * 27 * kCQ * 32 * 4;
constexpr size_t kP2SmemBytes = kP2ProfBytes + 27 * 32;

__constant__ signed char c_blosum_p2[27][32];
static bool g_loaded2[16] = {false};

static cudaError_t ensure_table2() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 16 && g_loaded2[dev]) return cudaSuccess;
  cudaError_t e = cudaMemcpyToSymbol(c_blosum_p2, TRPA_BLOSUM62, sizeof(TRPA_BLOSUM62));
  if (e == cudaSuccess && dev >= 0 && dev < 16) g_loaded2[dev] = true;
  return e;
}

__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }

// half-warp variant (protein2h_kernel below): pairs with |A| <= 320 columns
constexpr int kHMaxCols = 320;
constexpr int kHCQ = 5;   // profile capacity of the half-warp kernel: 20 columns per lane

// true: the pair is computed by protein2h_kernel (and skipped by protein2_kernel)
__device__ __forceinline__ bool p2h_takes(int n, int m) { return n > 0 && m > 0 && n <= kHMaxCols && m <= kP2MaxLen; }

constexpr int SH = 13;
constexpr int PRIO_MASK = 3 << 11;
constexpr int ROWBIAS = 8 << (SH - 1);                        // what the +8 of the unsigned profile adds per row
constexpr int CV = -(1 << SH) + (1 << 11) + 1 + ROWBIAS;      // vertical: score-1, priority 1, +1 gap, next row
constexpr int CH = -(1 << SH) + 1;                            // horizontal: score-1, priority 0, +1 gap
__device__ __forceinline__ int p2_boundary(int k) { return -k * (1 << SH) + k; }  // score -k after k gap columns

// One column strip of <= 32*C columns, C columns per lane, RIGHT aligned: the strip's last column is
// lane 31's last column and the first (32*C - ns) columns of the low lanes are padding.  A padding
// column has profile 0 (= substitution -4.5) for every residue and row-0 value 0, so neither the diagonal nor the
// horizontal candidate can win there and the vertical candidate reproduces the left boundary column
// value (-i, i gaps) in every padding cell -- the first real column sees exactly the boundary it
// needs.  No per-cell predicates, the wavefront always spans all 32 lanes.  Only the FIRST strip of a
// pair may be partial (the caller right-aligns the whole sequence); later strips are full.
template <int C>
__device__ __forceinline__ int protein2_strip(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int m,
                                              int s0, int ns, bool first_strip, bool last_strip, unsigned char* prof,
                                              const signed char* tbl, int* my_scratch, u32 lane) {
  constexpr int CQ = (C + 3) / 4;
  const int pad = 32 * C - ns;                    // leading padding columns of the strip
  const int v1 = (int)lane * C - pad;             // strip-relative index of this lane's first column (may be < 0)
  int up[C];
  int ac[CQ * 4];
  __syncwarp();
#pragma unroll
  for (int c = 0; c < CQ * 4; ++c) {
    const int v = v1 + c;
    ac[c] = (c < C && v >= 0) ? (int)a[s0 + v] : -1;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int j = s0 + v1 + c + 1;                // 1-based global column; <= s0 for padding
    up[c] = (v1 + c >= 0) ? p2_boundary(j) : 0;   // padding (first strip only): cell(0, 0)
  }
  for (int bb = 0; bb < 27; ++bb) {
#pragma unroll
    for (int q = 0; q < CQ; ++q) {
      u32 w = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = ac[4 * q + k];
        const int e = r >= 0 ? 2 * tbl[bb * 32 + r] + 9 : 0;   // BLOSUM62 is symmetric: [bb][r] keeps the lanes of a warp on distinct banks (or the same word)
        w |= (u32)(uint8_t)e << (8 * k);
      }
      *reinterpret_cast<u32*>(prof + (bb * CQ + q) * 128) = w;
    }
  }
  __syncwarp();
  // diagonal input of this lane's first column at row 1 = cell(0, first column - 1)
  int diag0;
  {
    const int v = v1 - 1;                         // strip-relative column left of my first one
    const int j = s0 + v + 1;
    diag0 = (v >= 0) ? p2_boundary(j) : p2_boundary(s0);
  }
  // TWO rows per wavefront step: lane l works on rows 2(t-l)-1 and 2(t-l) at step t.  Both row bodies are
  // straight-line code in one basic block, so the scheduler interleaves row i+1's first cells with row i's
  // last ones (two independent max chains instead of one: the fixed-latency "wait" stall was the largest,
  // profiles/r02_protein2_ncu.md) and the per-step bookkeeping (shuffles, predicates, addresses) is paid
  // once per two rows.  A row past the end (m odd) is computed on clamped inputs and never read.
  int last0 = 0, last1 = 0, res = 0;
  const u32 prof_sa = (u32)__cvta_generic_to_shared(prof);
  const int steps = ((m + 1) >> 1) + 31;
  // one LDS.U8 per cell (bank == lane, conflict free): the LSU pipe has room, the alu pipe does not -- a 32-bit
  // load per four cells costs a PRMT byte extract per cell on the alu pipe
  auto row = [&](u32 prow, int left, int diag) -> int {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      int e;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(e) : "r"(prow + (u32)((c >> 2) * 128 + (c & 3))));   // ptxas folds the offset
      const int D = e * (1 << (SH - 1)) + diag;   // score + sub, priority 2, gaps unchanged, row bias + 1
      const int V = up[c] + CV;
      const int H = left + CH;
      const int cell = max3i(D, V, H) & ~PRIO_MASK;
      diag = up[c];
      up[c] = cell;
      left = cell;
    }
    return left;
  };
  for (int t = 1; t <= steps; ++t) {
    const int recv0 = __shfl_up_sync(0xffffffffu, last0, 1);
    const int recv1 = __shfl_up_sync(0xffffffffu, last1, 1);
    const int i0 = 2 * (t - (int)lane) - 1, i1 = i0 + 1;
    if (i0 >= 1 && i0 <= m) {
      int left0, left1;
      if (lane == 0) {
        const int j1 = i1 <= m ? i1 : m;
        left0 = first_strip ? (p2_boundary(i0) + i0 * ROWBIAS) : my_scratch[i0];
        left1 = first_strip ? (p2_boundary(i1) + i1 * ROWBIAS) : my_scratch[j1];
      } else { left0 = recv0; left1 = recv1; }
      const u32 prow0 = prof_sa + (u32)b[i0 - 1] * (u32)(CQ * 128);
      const u32 prow1 = prof_sa + (u32)b[(i1 <= m ? i1 : m) - 1] * (u32)(CQ * 128);
      const int l0 = row(prow0, left0, diag0);      // row i0: diagonal input = left boundary of row i0 - 1
      const int l1 = row(prow1, left1, left0);      // row i1: diagonal input = left boundary of row i0
      diag0 = left1;
      last0 = l0; last1 = l1;
      if (lane == 31) {
        if (!last_strip) { my_scratch[i0] = l0; if (i1 <= m) my_scratch[i1] = l1; }
        else if (i0 == m) res = l0;
        else if (i1 == m) res = l1;
      }
    }
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, res, 31);
}

__global__ void __launch_bounds__(128)
protein2_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
                const uint8_t* __restrict__ residues, int2* __restrict__ out2, int2* __restrict__ scratch,
                u32 scratch_stride, u32 cq_rt, int skip_cols) {
  // cq_rt: column quads per lane the shared-memory profile is sized for (the launch's longest sequence decides:
  // 300-aa batches need 3 instead of 4 quads and fit a fifth CTA per SM)
  extern __shared__ __align__(16) signed char prof_all[];
  signed char* tbl = prof_all + (size_t)kP2Warps * 27 * cq_rt * 128;   // BLOSUM62 [a][b]
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) tbl[i] = c_blosum_p2[i >> 5][i & 31];
  __syncthreads();
  const u32 lane = threadIdx.x & 31;
  const u32 warp_in_cta = threadIdx.x >> 5;
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp_gid >= count) return;  // whole warp
  const PairDesc pd = pairs[warp_gid];
  const SeqDesc A = seqs[pd.a], B = seqs[pd.b];
  const int n = (int)A.len;  // columns (H)
  const int m = (int)B.len;  // rows (V)
  if (n > kP2MaxLen || m > kP2MaxLen) return;  // left to protein_kernel
  if (skip_cols == 1 && p2h_takes(n, m)) return;    // done by protein2h_kernel
  if (skip_cols == 2 && n > 0 && m > 0) return;     // done by protein3_kernel (all pairs of up to 1000 x 1000)
  const uint8_t* a = residues + A.woff;
  const uint8_t* b = residues + B.woff;
  if (n == 0 || m == 0) {
    if (lane == 0) out2[pd.out] = make_int2(-(n + m), 0);
    return;
  }
  unsigned char* prof = reinterpret_cast<unsigned char*>(prof_all) + ((size_t)warp_in_cta * 27 * cq_rt * 32 + lane) * 4;  // [b][c/4][lane][c%4]
  int* my_scratch = reinterpret_cast<int*>(scratch + (size_t)warp_gid * scratch_stride);
  int res = 0;
  int ns = ((n - 1) % (32 * kC2Max)) + 1;  // the first strip takes the remainder, later strips are full
  for (int s0 = 0; s0 < n; s0 += ns, ns = 32 * kC2Max) {
    const int Cneed = (ns + 31) >> 5;
    if (Cneed > 4 * (int)cq_rt) __trap();   // the launcher sized the profile from a wrong max_len: fail loudly
    const bool first = s0 == 0, lastS = s0 + ns == n;
    if (Cneed <= 4) res = protein2_strip<4>(a, b, m, s0, ns, first, lastS, prof, tbl, my_scratch, lane);
    else if (Cneed <= 8) res = protein2_strip<8>(a, b, m, s0, ns, first, lastS, prof, tbl, my_scratch, lane);
    else if (Cneed <= 10) res = protein2_strip<10>(a, b, m, s0, ns, first, lastS, prof, tbl, my_scratch, lane);
    else if (Cneed <= 12) res = protein2_strip<12>(a, b, m, s0, ns, first, lastS, prof, tbl, my_scratch, lane);
    else res = protein2_strip<16>(a, b, m, s0, ns, first, lastS, prof, tbl, my_scratch, lane);
  }
  if (lane == 0) {
    const int P = res - m * ROWBIAS;        // remove the per-row bias of the last row
    const int score = P >> SH;              // arithmetic shift: floor, low fields are non-negative
    const int gaps = P & 0x7ff;
    const int ndiag = (n + m - gaps) / 2;   // |A| + |B| = 2*#diag + #gaps
    out2[pd.out] = make_int2(score, ndiag);
  }
}

// ---- half-warp variant: TWO pairs per warp, 16 lanes and up to 20 columns per lane each (|A| <= 320, one strip).
// For the 300-aa pairs of C3 the wavefront ramp shrinks from 31 to 15 steps and the per-step bookkeeping is
// shared by 40 instead of 20 cells.  Same packed cells, same per-lane profile layout (bank == warp lane), same
// two rows per step as protein2_strip; pairs that do not fit are left to protein2_kernel / protein_kernel.

template <int C>
__device__ __forceinline__ void protein2h_run(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int n, int m,
                                              bool mine, unsigned char* prof, const signed char* tbl, u32 lp,
                                              int2* __restrict__ out2, u32 oidx) {
  constexpr int CQ = (C + 3) / 4;
  const int pad = 16 * C - n;                     // leading padding columns (right-aligned)
  const int v1 = (int)lp * C - pad;               // index of this lane's first column (may be < 0)
  int up[C];
  int ac[CQ * 4];
#pragma unroll
  for (int c = 0; c < CQ * 4; ++c) {
    const int v = v1 + c;
    ac[c] = (c < C && v >= 0 && mine) ? (int)a[v] : -1;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) up[c] = (v1 + c >= 0) ? p2_boundary(v1 + c + 1) : 0;
  for (int bb = 0; bb < 27; ++bb) {
#pragma unroll
    for (int q = 0; q < CQ; ++q) {
      u32 w = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = ac[4 * q + k];
        const int e = r >= 0 ? 2 * tbl[bb * 32 + r] + 9 : 0;   // BLOSUM62 is symmetric: [bb][r] keeps the lanes of a warp on distinct banks (or the same word)
        w |= (u32)(uint8_t)e << (8 * k);
      }
      *reinterpret_cast<u32*>(prof + (bb * CQ + q) * 128) = w;
    }
  }
  __syncwarp();
  int diag0 = (v1 - 1 >= 0) ? p2_boundary(v1) : p2_boundary(0);   // cell(0, first column - 1)
  const u32 prof_sa = (u32)__cvta_generic_to_shared(prof);
  auto row = [&](u32 prow, int left, int diag) -> int {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      int e;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(e) : "r"(prow + (u32)((c >> 2) * 128 + (c & 3))));
      const int D = e * (1 << (SH - 1)) + diag;
      const int V = up[c] + CV;
      const int H = left + CH;
      const int cell = max3i(D, V, H) & ~PRIO_MASK;
      diag = up[c];
      up[c] = cell;
      left = cell;
    }
    return left;
  };
  const int mo = __shfl_xor_sync(0xffffffffu, m, 16);
  const int steps = (((m > mo ? m : mo) + 1) >> 1) + 15;
  int last0 = 0, last1 = 0, res = 0;
  for (int t = 1; t <= steps; ++t) {
    const int recv0 = __shfl_up_sync(0xffffffffu, last0, 1, 16);
    const int recv1 = __shfl_up_sync(0xffffffffu, last1, 1, 16);
    const int i0 = 2 * (t - (int)lp) - 1, i1 = i0 + 1;
    if (i0 >= 1 && i0 <= m) {
      int left0, left1;
      if (lp == 0) { left0 = p2_boundary(i0) + i0 * ROWBIAS; left1 = p2_boundary(i1) + i1 * ROWBIAS; }
      else { left0 = recv0; left1 = recv1; }
      const u32 prow0 = prof_sa + (u32)b[i0 - 1] * (u32)(CQ * 128);
      const u32 prow1 = prof_sa + (u32)b[(i1 <= m ? i1 : m) - 1] * (u32)(CQ * 128);
      const int l0 = row(prow0, left0, diag0);
      const int l1 = row(prow1, left1, left0);
      diag0 = left1;
      last0 = l0; last1 = l1;
      if (lp == 15) { if (i0 == m) res = l0; else if (i1 == m) res = l1; }
    }
  }
  if (lp == 15 && mine) {
    const int P = res - m * ROWBIAS;
    const int score = P >> SH;
    const int gaps = P & 0x7ff;
    out2[oidx] = make_int2(score, (n + m - gaps) / 2);
  }
}

__global__ void __launch_bounds__(128)
protein2h_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
                 const uint8_t* __restrict__ residues, int2* __restrict__ out2) {
  extern __shared__ __align__(16) signed char prof_all[];
  signed char* tbl = prof_all + (size_t)kP2Warps * 27 * kHCQ * 128;   // BLOSUM62 [a][b]
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) tbl[i] = c_blosum_p2[i >> 5][i & 31];
  __syncthreads();
  const u32 lane = threadIdx.x & 31, lp = lane & 15u;
  const u32 warp_in_cta = threadIdx.x >> 5;
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 pidx = warp_gid * 2u + (lane >> 4);
  int n = 0, m = 0;
  const uint8_t* a = residues;
  const uint8_t* b = residues;
  u32 oidx = 0;
  bool mine = false;   // this half's pair is handled here
  if (pidx < count) {
    const PairDesc pd = pairs[pidx];
    const SeqDesc A = seqs[pd.a], B = seqs[pd.b];
    if (p2h_takes((int)A.len, (int)B.len)) {
      mine = true; n = (int)A.len; m = (int)B.len; oidx = pd.out;
      a = residues + A.woff; b = residues + B.woff;
    }
  }
  if (!__any_sync(0xffffffffu, mine)) return;
  // both halves use the lane width (columns per lane) the longer of the two pairs needs
  const int no = __shfl_xor_sync(0xffffffffu, n, 16);
  const int cols = ((n > no ? n : no) + 15) >> 4;
  unsigned char* prof = reinterpret_cast<unsigned char*>(prof_all) + ((size_t)warp_in_cta * 27 * kHCQ * 32 + lane) * 4;  // [b][c/4][lane][c%4]
  if (cols <= 8) protein2h_run<8>(a, b, n, m, mine, prof, tbl, lp, out2, oidx);
  else if (cols <= 12) protein2h_run<12>(a, b, n, m, mine, prof, tbl, lp, out2, oidx);
  else if (cols <= 16) protein2h_run<16>(a, b, n, m, mine, prof, tbl, lp, out2, oidx);
  else protein2h_run<20>(a, b, n, m, mine, prof, tbl, lp, out2, oidx);
}

static bool protein_half_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TRPA_PROTEIN_HALF"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// protein3.cu: eight lanes per pair, the default for the pairs the half-warp kernel would take (A/B hook:
// TRPA_PROTEIN_QUARTER=0 runs the half-warp kernel instead)
cudaError_t launch_protein3(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues, int2* out2,
                            u32 max_len, u32 aa_mask, cudaStream_t stream);
static bool protein_quarter_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TRPA_PROTEIN_QUARTER"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

static cudaError_t launch_h(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues, int2* out2,
                            cudaStream_t stream) {
  const size_t smem = (size_t)kP2Warps * 27 * kHCQ * 128 + 27 * 32;
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr = attr_set[dev & 15];   // function attributes are per device
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(protein2h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const u32 warps = (count + 1u) / 2u;
  const u32 blocks = (warps + kP2Warps - 1) / kP2Warps;
  protein2h_kernel<<<blocks, 32 * kP2Warps, smem, stream>>>(pairs, count, seqs, residues, out2);
  return cudaGetLastError();
}

cudaError_t launch_protein2(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues,
                            int2* out2, int2* scratch, u32 scratch_stride, u32 max_len, u32 aa_mask, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  cudaError_t e = ensure_table2();
  if (e != cudaSuccess) return e;
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr = attr_set[dev & 15];   // function attributes are per device
  if (!attr) {
    e = cudaFuncSetAttribute(protein2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kP2SmemBytes);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const bool half = protein_half_enabled();
  const bool quarter = half && protein_quarter_enabled();
  if (half) {
    e = quarter ? launch_protein3(pairs, count, seqs, residues, out2, max_len, aa_mask, stream)
                : launch_h(pairs, count, seqs, residues, out2, stream);
    if (e != cudaSuccess) return e;
  }
  const u32 blocks = (count + kP2Warps - 1) / kP2Warps;
  // profile capacity: columns per lane of the longest strip any pair of this launch can have
  const u32 longest = max_len ? (max_len > (u32)kP2MaxLen ? (u32)kP2MaxLen : max_len) : (u32)kP2MaxLen;
  u32 cols = (longest + 31u) / 32u;
  if (cols > (u32)kC2Max) cols = kC2Max;
  // the strip templates round the columns per lane up to 4, 8, 10, 12 or 16
  const u32 cq = cols <= 4 ? 1u : (cols <= 8 ? 2u : (cols <= 12 ? 3u : 4u));
  const size_t smem = (size_t)kP2Warps * 27 * cq * 128 + 27 * 32;
  // with protein3 only the empty pairs are left for this kernel (pairs beyond 1000 residues: protein.cu)
  protein2_kernel<<<blocks, 32 * kP2Warps, smem, stream>>>(pairs, count, seqs, residues, out2, scratch, scratch_stride, cq,
                                                           quarter ? 2 : (half ? 1 : 0));
  return cudaGetLastError();
}

int protein2_max_len() { return kP2MaxLen; }

}  // namespace trpa
