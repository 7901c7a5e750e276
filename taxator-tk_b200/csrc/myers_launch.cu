// Instantiations + shape dispatch of the bit-vector edit-distance kernel (see myers.cuh).
#include "myers.cuh"
#include "shapes.h"
#include "launch.h"

namespace trpa {

template <int W, bool HASN>
static cudaError_t launch_one(const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                              const u32* nplane, int* out, int L, u32* scratch, u32 scratch_stride,
                              const uint2* bucket, cudaStream_t stream) {
  const u32 G = 32 / L;
  const u32 warps = (count + G - 1) / G;
  const u32 blocks = (warps + 3) / 4;  // 4 warps per CTA
  myers_kernel<W, HASN><<<blocks, 128, 0, stream>>>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket);
  return cudaGetLastError();
}

template <bool HASN>
static cudaError_t launch_w(int widx, const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                            const u32* nplane, int* out, int L, u32* scratch, u32 scratch_stride,
                            const uint2* bucket, cudaStream_t stream) {
  switch (widx) {
    case 0: return launch_one<1, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
    case 1: return launch_one<2, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
    case 2: return launch_one<4, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
    case 3: return launch_one<8, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
    case 4: return launch_one<12, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
    case 5: return launch_one<16, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
    case 6: return launch_one<20, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
    default: return launch_one<24, HASN>(pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
  }
}

cudaError_t launch_myers(int shape, const PairDesc* pairs, u32 count, const SeqDesc* seqs, const uint2* planes,
                         const u32* nplane, int* out, u32* scratch, u32 scratch_stride, const uint2* bucket,
                         cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  const int L = 1 << shape_lidx(shape);
  const int widx = shape_widx(shape);
  if (shape_hasn(shape))
    return launch_w<true>(widx, pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
  return launch_w<false>(widx, pairs, count, seqs, planes, nplane, out, L, scratch, scratch_stride, bucket, stream);
}

}  // namespace trpa
