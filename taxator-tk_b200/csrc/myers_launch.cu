// Instantiations + shape dispatch of the bit-vector edit-distance kernels (myers2.cuh: persistent,
// shared-memory equality tables; myers.cuh: first-generation kernel kept for A/B runs with
// TRPA_MYERS_V1=1).
#include <cstdlib>

#include "myers2.cuh"
#include "shapes.h"
#include "launch.h"

namespace trpa {

static int g_num_sms = 0;
static bool use_v1() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TRPA_MYERS_V1"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

template <int W, bool HASN>
static cudaError_t launch_one(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                              const u32* nplane, int* out, int L, u32* scratch, u32 scratch_stride,
                              const uint2* bucket, u32* cursor, cudaStream_t stream) {
  const u32 G = 32 / L;
  const u32 warps = (count + G - 1) / G;
  u32 blocks = (warps + 3) / 4;  // 4 warps per CTA
  if (use_v1() || cursor == nullptr) {
    myers_kernel<W, HASN><<<blocks, 128, 0, stream>>>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket);
    return cudaGetLastError();
  }
  typedef Myers2Cfg<W, HASN> Cfg;
  static int occ = 0;  // resident CTAs per SM of this instantiation
  if (occ == 0) {
    cudaError_t e = cudaFuncSetAttribute(myers2_kernel<W, HASN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, myers2_kernel<W, HASN>, 128, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (g_num_sms <= 0) g_num_sms = 148;
    }
  }
  const u32 resident = (u32)g_num_sms * (u32)occ;  // persistent grid: one wave
  if (blocks > resident) blocks = resident;
  myers2_kernel<W, HASN><<<blocks, 128, Cfg::kSmemBytes, stream>>>(pairs, count, seqs, planes, nplane, out, L, scratch,
                                                                  scratch_stride, bucket, cursor);
  return cudaGetLastError();
}

template <bool HASN>
static cudaError_t launch_w(int widx, const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                            const u32* nplane, int* out, int L, u32* scratch, u32 scratch_stride,
                            const uint2* bucket, u32* cursor, cudaStream_t stream) {
#define TRPA_CASE(I, WV) \
  case I: return launch_one<WV, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, cursor, stream);
  switch (widx) {
    TRPA_CASE(0, 1) TRPA_CASE(1, 2) TRPA_CASE(2, 4) TRPA_CASE(3, 8) TRPA_CASE(4, 12) TRPA_CASE(5, 16) TRPA_CASE(6, 20)
    default: return launch_one<24, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, cursor, stream);
  }
#undef TRPA_CASE
}

// number of persistent group slots a launch of `count` pairs of this shape can use (scratch sizing)
u32 myers_group_slots(int shape, u32 count) {
  const int L = 1 << shape_lidx(shape);
  const u32 G = 32 / L;
  const u32 warps = (count + G - 1) / G;
  const u32 blocks = (warps + 3) / 4;
  return blocks * 4 * G;  // upper bound (v1 uses one slot per pair, v2 at most one per resident group)
}

cudaError_t launch_myers(int shape, const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                         const u32* nplane, int* out, u32* scratch, u32 scratch_stride, const uint2* bucket,
                         u32* cursor, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const int L = 1 << shape_lidx(shape);
  const int widx = shape_widx(shape);
  if (shape_hasn(shape))
    return launch_w<true>(widx, pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, cursor, stream);
  return launch_w<false>(widx, pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, cursor, stream);
}

}  // namespace trpa
