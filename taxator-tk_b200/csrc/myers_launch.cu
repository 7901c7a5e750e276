// Per-round planning (threshold + shape of every pair), instantiations and shape dispatch of the banded
// bit-vector edit-distance kernel (myers3.cuh).
#include <algorithm>
#include <cstdlib>

#include "myers3.cuh"
#include "shapes.h"
#include "launch.h"

namespace trpa {

static int g_num_sms = 0;

// ------------------------------------------------------------------------------------------------
// Banded path (myers3.cuh): per-round planning + launch.

// One warp per pair: (1) upper bound of the distance from the ungapped (Hamming) alignment of the
// pattern against the text prefix (+ the length difference), combined with the state machine's hint
// (PairDesc.pad on entry, 0 = none) -> initial threshold k0; (2) shape: the lanes evaluate the 48
// (W, L) candidates in parallel and keep the one minimising  cost * pairs-per-lane + time  (lane-time
// when pairs are plentiful, latency when they are scarce).  pad <- shape | k0 << 8; histogram.
__global__ void plan_kernel(PairDesc* __restrict__ pairs, u32 n_pairs, const SeqDesc* __restrict__ descs,
                            const uint2* __restrict__ planes, const u32* __restrict__ nplane,
                            u32* __restrict__ hist, u32 lanes_total, int band, int force_shape, int wedge, const PlanParams pp) {
  __shared__ u32 plan_bins[4][32];   // blockDim.x == 128
  // shape / duration-class counts of this CTA's pairs; added to the global histogram once per CTA (one pair = two
  // atomics on a handful of global addresses otherwise: 8.5 M per C2 step)
  __shared__ u32 sh_hist[kNumShapes * (1 + kNumCls)];
  for (u32 i = threadIdx.x; i < (u32)(kNumShapes * (1 + kNumCls)); i += blockDim.x) sh_hist[i] = 0u;
  __syncthreads();
  const u32 lane = threadIdx.x & 31;
  const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 p = warp; p < n_pairs; p += nwarps) {
    const PairDesc pd = pairs[p];
    const SeqDesc A = descs[pd.a], B = descs[pd.b];
    u32 m, n, pw, tw;
    if (A.len < B.len) { m = A.len; pw = A.woff; n = B.len; tw = B.woff; }
    else               { m = B.len; pw = B.woff; n = A.len; tw = A.woff; }
    const int hasn = (int)((A.flags | B.flags) & 1u);
    u32 k0 = kPadKFull, a1 = 0;
    if (band && m >= 64u) {
      // lane l ends up with the mismatches of the l-th contiguous 1/32 of the pattern, so that the prefix sums
      // over the lanes give the mismatch profile along the sequence.  The words are READ interleaved
      // (coalesced) and binned through a 32-entry shared-memory histogram per warp.
      const u32 mwords = (m + 31) >> 5;
      const u32 per = (mwords + 31u) >> 5;
      u32* bins = plan_bins[threadIdx.x >> 5];
      bins[lane] = 0u;
      __syncwarp();
      // (the words of the NEXT iteration are requested before this iteration's are used: the loop was the kernel's
      // largest stall, 36 % of the samples on the first use of a word just loaded)
      const uint2 zero2 = make_uint2(0u, 0u);
      uint2 x = lane < mwords ? planes[pw + lane] : zero2, y = lane < mwords ? planes[tw + lane] : zero2;
      u32 xn = (hasn && lane < mwords) ? nplane[pw + lane] : 0u, yn = (hasn && lane < mwords) ? nplane[tw + lane] : 0u;
      for (u32 w = lane; w < mwords; w += 32u) {
        const u32 wn = w + 32u;
        const bool more = wn < mwords;
        const uint2 x2 = more ? planes[pw + wn] : zero2, y2 = more ? planes[tw + wn] : zero2;
        const u32 xn2 = (hasn && more) ? nplane[pw + wn] : 0u, yn2 = (hasn && more) ? nplane[tw + wn] : 0u;
        u32 mm = (x.x ^ y.x) | (x.y ^ y.y) | (xn ^ yn);
        if (w == mwords - 1 && (m & 31u)) mm &= (1u << (m & 31u)) - 1u;
        const u32 c = __popc(mm);
        if (c) atomicAdd(&bins[w / per], c);
        x = x2; y = y2; xn = xn2; yn = yn2;
      }
      __syncwarp();
      const u32 h = bins[lane];
      __syncwarp();
      u32 cum = h;   // inclusive prefix sum over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, cum, o);
        if ((int)lane >= o) cum += v;
      }
      const u32 total = __shfl_sync(0xffffffffu, cum, 31);
      u32 ub = total + (n - m);                   // a valid alignment: d <= ub
      const u32 hint = pd.pad;
      bool from_hint = false;
      u32 hint_value = 0;
      if (hint) {
        // an estimate gets a margin (the kernel verifies and widens); a true bound (kHintIsBound) is used as is
        const u32 hv = hint & ~kHintIsBound;
        const u32 hk = (hint & kHintIsBound) ? hv : (u32)min((uint64_t)hv * pp.hint_mul64 / 64u + pp.hint_add, (uint64_t)kPadKFull);
        if (hk < ub) { ub = hk; from_hint = true; hint_value = hv; }
      }
      if (band > 1) ub = (u32)band;               // test hook: forced initial threshold
      k0 = ub < kPadKFull ? ub : kPadKFull - 1u;
      // wedge: slowest average mismatch rate over any prefix (>= 1/8 of the pattern) as a cautious
      // estimate of how fast errors accumulate along the alignment; 3/4 of it narrows the band
      // (measured, scripts/wedge_probe.py: pays from ~8 % divergence and 2 kb on; with half widths beyond
      // ~1000 the cheapest paths to far-off cells dodge most mismatches and the certificate fails)
      const bool pays = m >= 2048u && total * 3u < m && total * 12u >= m && total + (n - m) <= pp.wedge_max_k;
      // threshold from an estimate (noisy long reads: the mismatch profile of an ungapped overlay says nothing once
      // indels shift the diagonal): assume the estimated errors accumulate uniformly along the pattern
      const bool pays_hint = from_hint && pp.wedge_hint_s8 && m >= 2048u && (uint64_t)hint_value * 3u < m &&
                             (uint64_t)hint_value * 16u >= m && ub <= pp.wedge_hint_max_k;
      if (wedge && band == 1 && (wedge == 2 ? (m >= 256u && !from_hint) : (from_hint ? pays_hint : pays))) {
        const u32 pos = min(m, (lane + 1u) * per * 32u);      // pattern rows covered up to this lane
        // rate in mismatches per 2^16 rows, rounded down
        u32 rate = pos * 8u >= m ? (u32)(((uint64_t)cum << 16) / pos) : 0xffffffffu;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rate = min(rate, __shfl_xor_sync(0xffffffffu, rate, o));
        const u32 s8 = from_hint ? pp.wedge_hint_s8 : pp.wedge_s8;
        if (from_hint) rate = (u32)(((uint64_t)hint_value << 16) / m);
        // A cell at diagonal offset a of row `row` has seen the errors of its first (row - a) columns
        // only, and the count fluctuates: keep the full width until about 30 errors are expected
        // (x0 rows after the first a0), then  2 a + s rate (row - a - x0) = k - delta  with s = 3/4
        //   ->  a(m) = (K - s rate (m - x0)) / (2 - s rate)
        // (the exit cells of the first strips have seen almost no errors yet: their certificate term is
        // about k0 itself, so k0 gets a small cushion over the distance bound)
        const u32 delta = n - m;
        const u32 k0w = k0 + pp.wedge_cushion;
        const u32 a0 = k0w > delta ? (k0w - delta) >> 1 : 0u;
        const uint64_t K = k0w > delta ? k0w - delta : 0u;
        uint64_t x0l = rate ? (((uint64_t)pp.wedge_e0 << 16) + rate - 1u) / rate : (uint64_t)m;
        if (x0l > m) x0l = m;
        const u32 x0 = (u32)x0l;
        const uint64_t sr = ((uint64_t)rate * s8) >> 3;                   // s * rate, per 2^16 rows (s = 3/4 by default)
        const uint64_t e_end = m > x0 ? (sr * (m - x0)) >> 16 : 0u;
        a1 = K > e_end ? (u32)(((K - e_end) << 16) / ((2ull << 16) - sr)) : 0u;
        if (a1 < 32u) a1 = 32u;
        if (a0 + x0 + 64u >= m) a1 = a0;          // nothing left to narrow
        const u32 packed = wedge_pack(a1, x0);
        a1 = a1 + 16u >= a0 ? 0u : packed;      // a1 now holds the packed wedge request (0 = none)
        if (a1 && k0w < kPadKFull) k0 = k0w;
        if (wedge == 2) a1 = wedge_pack(32u, 0u); // test hook: narrow at once to the minimum, no cushion
      }
    }
    const BandGeom g = band_from_k(m ? m : 1u, n ? n : 1u, k0 >= kPadKFull ? 0xffffffffu : k0, !band, a1);
    uint64_t key = ~0ull, mytime = 0;
    int best = 0;
    // 32 candidates, one per lane: W in {2, 4, 8, 12, 16} x L in {1 .. 32}, + W in {20, 24} at L = 32
    // (long patterns with wide bands); W = 1 never wins (per-column overhead)
    {
      int w, l;
      if (lane < 30) { w = 1 + (int)(lane % 5u); l = (int)(lane / 5u); }
      else { w = 6 + (int)(lane - 30u); l = kNumL - 1; }
      const ShapeCost sc = band_shape_cost(g, w, l, pp);
      // lane-time when pairs are plentiful, latency when they are scarce; the latency term is convex (time +
      // time^2 / 2^pp.tail_log2): a pair that runs for a large part of a round IS the round's tail, short pairs hide
      // behind the others (measured with the longest-first bucket order)
      const uint64_t t2 = (sc.time >> 8) * (sc.time >> (pp.tail_log2 - 8u));
      key = sc.cost * n_pairs + (sc.time + t2) * lanes_total;
      best = l * kNumW + w;
      mytime = sc.time;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t ok = __shfl_xor_sync(0xffffffffu, key, o);
      const int ob = __shfl_xor_sync(0xffffffffu, best, o);
      const uint64_t ot = __shfl_xor_sync(0xffffffffu, mytime, o);
      if (ok < key || (ok == key && ob < best)) { key = ok; best = ob; mytime = ot; }
    }
    if (force_shape >= 0) best = force_shape;
    if (lane == 0) {
      const int shape = shape_id(best % kNumW, best / kNumW, hasn);
      pairs[p].pad = (k0 << 8) | (u32)shape;
      pairs[p].aux = g.wedge() ? a1 : 0u;
      const u32 cls = duration_class(mytime);
      pairs[p].cls = cls;
      atomicAdd(&sh_hist[shape], 1u);
      atomicAdd(&sh_hist[kNumShapes + shape * kNumCls + cls], 1u);
    }
  }
  __syncthreads();
  for (u32 i = threadIdx.x; i < (u32)(kNumShapes * (1 + kNumCls)); i += blockDim.x) {
    const u32 v = sh_hist[i];
    if (v) atomicAdd(&hist[i < (u32)kNumShapes ? i : 3u * kNumShapes + (i - kNumShapes)], v);
  }
}

cudaError_t launch_plan(PairDesc* pairs, u32 n_pairs, const SeqDesc* descs, const uint2* planes, const u32* nplane,
                        u32* hist, u32 lanes_total, int band, int force_shape, int wedge, const PlanParams& pp, cudaStream_t stream) {
  if (n_pairs == 0) return cudaSuccess;
  const u32 blocks = std::min<u32>((n_pairs + 3) / 4, 148u * 16u);
  plan_kernel<<<blocks, 128, 0, stream>>>(pairs, n_pairs, descs, planes, nplane, hist, lanes_total, band, force_shape, wedge, pp);
  return cudaGetLastError();
}

template <int W, bool HASN>
static cudaError_t launch_one3(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                               const u32* nplane, const uint2* codes, int* out, int L, uint4* scratch, u32 scratch_stride,
                               u32* cursor, unsigned long long* stats, int force_full, u32* slots_out,
                               cudaStream_t stream) {
  typedef Myers3Cfg<W, HASN> Cfg;
  static int occ = 0;
  if (occ == 0) {
    cudaError_t e = cudaFuncSetAttribute(myers3_kernel<W, HASN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, myers3_kernel<W, HASN>, 128, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (g_num_sms <= 0) g_num_sms = 148;
    }
  }
  const u32 G = 32 / L;
  const u32 warps = (count + G - 1) / G;
  u32 blocks = (warps + 3) / 4;
  const u32 resident = (u32)g_num_sms * (u32)occ;
  if (blocks > resident) blocks = resident;
  if (slots_out) { *slots_out = blocks * 4 * G; return cudaSuccess; }
  myers3_kernel<W, HASN><<<blocks, 128, Cfg::kSmemBytes, stream>>>(pairs, count, seqs, planes, nplane, codes, out, L, scratch,
                                                                  scratch_stride, nullptr, cursor, stats, force_full);
  return cudaGetLastError();
}

// slots_out != nullptr: only report the number of group slots (scratch lines) the launch would use
cudaError_t launch_myers3(int shape, const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                          const u32* nplane, const uint2* codes, int* out, uint4* scratch, u32 scratch_stride, u32* cursor,
                          unsigned long long* stats, int force_full, u32* slots_out, cudaStream_t stream) {
  if (count == 0) { if (slots_out) *slots_out = 0; return cudaSuccess; }
  const int L = 1 << shape_lidx(shape);
  const bool hasn = shape_hasn(shape) != 0;
#define TRPA_CASE3(I, WV)                                                                                          \
  case I:                                                                                                          \
    return hasn ? launch_one3<WV, true>(pairs, count, seqs, planes, nplane, codes, out, L, scratch, scratch_stride, cursor, \
                                        stats, force_full, slots_out, stream)                                      \
                : launch_one3<WV, false>(pairs, count, seqs, planes, nplane, codes, out, L, scratch, scratch_stride, cursor, \
                                         stats, force_full, slots_out, stream);
  switch (shape_widx(shape)) {
    TRPA_CASE3(0, 1) TRPA_CASE3(1, 2) TRPA_CASE3(2, 4) TRPA_CASE3(3, 8) TRPA_CASE3(4, 12) TRPA_CASE3(5, 16) TRPA_CASE3(6, 20)
    default:
      return hasn ? launch_one3<24, true>(pairs, count, seqs, planes, nplane, codes, out, L, scratch, scratch_stride, cursor,
                                          stats, force_full, slots_out, stream)
                  : launch_one3<24, false>(pairs, count, seqs, planes, nplane, codes, out, L, scratch, scratch_stride, cursor,
                                           stats, force_full, slots_out, stream);
  }
#undef TRPA_CASE3
}

}  // namespace trpa
