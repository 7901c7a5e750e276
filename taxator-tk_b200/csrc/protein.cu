// Batched protein-space global alignment: Needleman-Wunsch, BLOSUM62, linear gap -1 per column, with
// the length of SeqAn's single traced path -- what the reference computes with
//   seqan::globalAlignment(align, Blosum62(), LinearGaps()) + row walk
// at core/src/taxonpredictionmodelsequence.hh:199-227 (recurrence dp_formula_linear.h:62-105, tie
// order diagonal >= vertical >= horizontal via dp_formula.h:152-163, traceback
// dp_traceback_impl.h:379-418).  No trace matrix: every cell carries (score, #diagonal steps of the
// path SeqAn's traceback would take), so traced length = |A| + |B| - #diag.
//
// One warp per pair, anti-diagonal wavefront: lane l owns C consecutive columns (H = sequence A) in
// registers and walks rows (V = sequence B) skewed by l; the right boundary of its block travels to
// lane l+1 by __shfl_up_sync.  A longer than 32*CMAX columns are processed in column strips with
// the strip's last column kept in an HBM/L2 scratch line.  BLOSUM62 sits in shared memory.
#include <cstdlib>

#include "common.cuh"
#include "launch.h"
#include "blosum62_table.h"

namespace trpa {

constexpr int kCMax = 16;

__constant__ signed char c_blosum_p[27][32];
static bool g_loaded[16] = {false};

static cudaError_t ensure_table() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 16 && g_loaded[dev]) return cudaSuccess;
  cudaError_t e = cudaMemcpyToSymbol(c_blosum_p, TRPA_BLOSUM62, sizeof(TRPA_BLOSUM62));
  if (e == cudaSuccess && dev >= 0 && dev < 16) g_loaded[dev] = true;
  return e;
}

__global__ void __launch_bounds__(128)
protein_kernel(const PairDesc* __restrict__ pairs, u32 count, const SeqDesc* __restrict__ seqs,
               const uint8_t* __restrict__ residues, int2* __restrict__ out2, int2* __restrict__ scratch,
               u32 scratch_stride, int skip_below) {
  __shared__ signed char tbl[27 * 32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) tbl[i] = c_blosum_p[i >> 5][i & 31];
  __syncthreads();

  const u32 lane = threadIdx.x & 31;
  const u32 warp_gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp_gid >= count) return;  // whole warp
  const PairDesc pd = pairs[warp_gid];
  const SeqDesc A = seqs[pd.a], B = seqs[pd.b];
  const uint8_t* a = residues + A.woff;
  const uint8_t* b = residues + B.woff;
  const int n = (int)A.len;  // columns (H)
  const int m = (int)B.len;  // rows (V)
  if (n <= skip_below && m <= skip_below) return;  // handled by protein3_kernel
  int2* my_scratch = scratch + (size_t)warp_gid * scratch_stride;

  int res_s = 0, res_nd = 0;
  if (n == 0 || m == 0) {
    if (lane == 0) out2[pd.out] = make_int2(-(n + m), 0);
    return;
  }

  for (int s0 = 0; s0 < n; s0 += 32 * kCMax) {
    const int ns = min(n - s0, 32 * kCMax);
    const int C = (ns + 31) >> 5;                 // columns per lane in this strip
    const int j1 = s0 + (int)lane * C + 1;        // first (1-based) column of this lane
    const int myC = max(0, min(C, s0 + ns - (j1 - 1)));
    const bool last_strip = s0 + ns == n;
    const int owner = (ns - 1) / C;               // lane holding the strip's last column
    int ac[kCMax], us[kCMax], und[kCMax];
#pragma unroll
    for (int c = 0; c < kCMax; ++c) {
      ac[c] = (c < myC) ? a[j1 - 1 + c] : 0;
      us[c] = -(j1 + c);   // row 0
      und[c] = 0;
    }
    int prev_s = -(j1 - 1), prev_nd = 0;   // cell(0, j1-1): diagonal for the first column at row 1
    int last_s = 0, last_nd = 0;           // my right boundary at the row just processed
    const int steps = m + owner;
    for (int t = 1; t <= steps; ++t) {
      int rs = __shfl_up_sync(0xffffffffu, last_s, 1);
      int rnd = __shfl_up_sync(0xffffffffu, last_nd, 1);
      const int i = t - (int)lane;  // row (1-based)
      if (i >= 1 && i <= m && myC > 0) {
        int ls, lnd;
        if (lane == 0) {
          if (s0 == 0) { ls = -i; lnd = 0; }
          else { int2 v = my_scratch[i]; ls = v.x; lnd = v.y; }
        } else { ls = rs; lnd = rnd; }
        const int left_in_s = ls, left_in_nd = lnd;
        int ds = prev_s, dnd = prev_nd;
        const signed char* trow = tbl + 32 * (int)b[i - 1];
#pragma unroll
        for (int c = 0; c < kCMax; ++c) {
          if (c < myC) {
            const int sub = trow[ac[c]];
            int bs = ds + sub, bnd = dnd + 1;                     // diagonal first
            const int vs = us[c] - 1;                             // vertical (gap in H)
            if (vs > bs) { bs = vs; bnd = und[c]; }
            const int hs = ls - 1;                                // horizontal (gap in V)
            if (hs > bs) { bs = hs; bnd = lnd; }
            ds = us[c]; dnd = und[c];
            us[c] = bs; und[c] = bnd;
            ls = bs; lnd = bnd;
          }
        }
        prev_s = left_in_s; prev_nd = left_in_nd;
        last_s = ls; last_nd = lnd;
        if ((int)lane == owner) {
          if (!last_strip) my_scratch[i] = make_int2(ls, lnd);
          else if (i == m) { res_s = ls; res_nd = lnd; }
        }
      }
    }
    __syncwarp();
    if (last_strip) {
      res_s = __shfl_sync(0xffffffffu, res_s, owner);
      res_nd = __shfl_sync(0xffffffffu, res_nd, owner);
    }
  }
  if (lane == 0) out2[pd.out] = make_int2(res_s, res_nd);
}

// protein3.cu: the packed-cell kernel for every pair of up to 1000 x 1000 residues (and the empty pairs)
cudaError_t launch_protein3(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues, int2* out2,
                            u32 max_len, u32 aa_mask, cudaStream_t stream);
constexpr int kPackedMaxLen = 1000;

// max_len: longest staged sequence of the launch (decides whether the 32-bit fallback kernel runs)
cudaError_t launch_protein(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint8_t* residues,
                           int2* out2, int2* scratch, u32 scratch_stride, u32 max_len, u32 aa_mask, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  cudaError_t e = ensure_table();
  if (e != cudaSuccess) return e;
  const u32 blocks = (count + 3) / 4;
  e = launch_protein3(pairs, count, seqs, residues, out2, max_len, aa_mask, stream);
  if (e != cudaSuccess) return e;
  if ((int)max_len > kPackedMaxLen) {
    protein_kernel<<<blocks, 128, 0, stream>>>(pairs, count, seqs, residues, out2, scratch, scratch_stride, kPackedMaxLen);
    return cudaGetLastError();
  }
  return cudaSuccess;
}

}  // namespace trpa
