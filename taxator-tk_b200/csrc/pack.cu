// Sequence store packing and segment staging kernels (HBM-bound byte/bit work).
//
// Store layout in HBM (replaces RandomInmemorySeqStoreRO / RandomIndexedSeqstoreRO,
// core/src/sequencestorage.hh:56-140,318-457):
//   NT: every sequence starts on a 32-base word boundary; word k of a sequence holds bases
//       32k..32k+31 as two bit-planes (uint2: .x low bit, .y high bit of A=0,C=1,G=2,T=3) plus an
//       "is N" plane.  char -> Dna5 as SeqAn does it: A,C,G,T,U (either case) -> 0..3, else N.
//   AA: 5-bit SeqAn AminoAcid ordinals, 6 per 32-bit word, every sequence word aligned.
// Staged segments (what the alignment kernels read) use the same NT plane layout / one byte per
// residue for AA, cut out per getSequence(id,start,stop,left_ext,right_ext) semantics
// (core/src/taxonpredictionmodelsequence.hh:856-880): clip to the sequence, reverse-complement
// when rstart > rstop (nucleotide only).
#include "common.cuh"
#include "launch.h"
#include "blosum62_table.h"

namespace trpa {

__constant__ signed char c_blosum[27][32];
static bool g_blosum_loaded[16] = {false};

cudaError_t ensure_blosum_constant(int device) {
  if (device >= 0 && device < 16 && g_blosum_loaded[device]) return cudaSuccess;
  cudaError_t e = cudaMemcpyToSymbol(c_blosum, TRPA_BLOSUM62, sizeof(TRPA_BLOSUM62));
  if (e == cudaSuccess && device >= 0 && device < 16) g_blosum_loaded[device] = true;
  return e;
}

__device__ __forceinline__ u32 dna5_code(u32 c) {
  c &= 0xdfu;  // fold case (also maps some punctuation onto letters' slots, excluded below)
  // only true letters reach a match: 'A'..'Z' are 0x41..0x5a in both cases after folding
  u32 code = 4;
  if (c == 'A') code = 0;
  else if (c == 'C') code = 1;
  else if (c == 'G') code = 2;
  else if (c == 'T' || c == 'U') code = 3;
  return code;
}

__device__ __forceinline__ u32 aa_code(u32 c) {
  // SeqAn order: A B C D E F G H I J K L M N O P Q R S T U V W Y Z X *   (X=25, '*'=26)
  if (c >= 'a' && c <= 'z') c -= 32;
  if (c == '*') return 26;
  if (c < 'A' || c > 'Z') return 25;
  if (c <= 'W') return c - 'A';
  if (c == 'X') return 25;
  return c - 'A' - 1;  // Y -> 23, Z -> 24
}

// one warp per output word
__global__ void pack_nt_kernel(const uint8_t* __restrict__ chars, const u64* __restrict__ off,
                               const SeqDesc* __restrict__ seqs, u32 n_seq, u64 total_words,
                               uint2* __restrict__ planes, u32* __restrict__ nplane, u32* __restrict__ seq_flags) {
  const u32 lane = threadIdx.x & 31;
  const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 gw = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < total_words; gw += warps) {
    // sequence containing word gw: last s with seqs[s].woff <= gw
    u32 lo = 0, hi = n_seq;
    while (hi - lo > 1) {
      u32 mid = (lo + hi) >> 1;
      if ((u64)seqs[mid].woff <= gw) lo = mid; else hi = mid;
    }
    const SeqDesc sd = seqs[lo];
    const u64 base = (gw - sd.woff) * 32 + lane;
    u32 code = 0;
    bool valid = base < sd.len;
    if (valid) code = dna5_code(chars[off[lo] + base]);
    const u32 isn = valid && code == 4;
    const u32 b0 = __ballot_sync(0xffffffffu, valid && !isn && (code & 1));
    const u32 b1 = __ballot_sync(0xffffffffu, valid && !isn && (code & 2));
    const u32 bn = __ballot_sync(0xffffffffu, isn);
    if (lane == 0) {
      planes[gw] = make_uint2(b0, b1);
      nplane[gw] = bn;
      if (bn && seq_flags) atomicOr(&seq_flags[lo], 1u);
    }
  }
}

cudaError_t launch_pack_nt(const uint8_t* chars, const u64* off, const SeqDesc* seqs, u32 n_seq, u64 total_words,
                           uint2* planes, u32* nplane, u32* seq_flags, cudaStream_t stream) {
  if (total_words == 0) return cudaSuccess;
  u64 blocks = (total_words + 7) / 8;  // 8 warps per CTA
  if (blocks > 148 * 64) blocks = 148 * 64;
  pack_nt_kernel<<<(u32)blocks, 256, 0, stream>>>(chars, off, seqs, n_seq, total_words, planes, nplane, seq_flags);
  return cudaGetLastError();
}

// column codes (common.cuh nt_codes) for a whole plane array (low-level pair API; the batch path gets
// them from the staging kernel)
__global__ void codes_from_planes_kernel(const uint2* __restrict__ planes, uint2* __restrict__ codes, u64 n_words) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    const uint2 p = planes[w];
    codes[w] = nt_codes(p.x, p.y);
  }
}

cudaError_t launch_codes_from_planes(const uint2* planes, uint2* codes, u64 n_words, cudaStream_t stream) {
  if (n_words == 0) return cudaSuccess;
  u64 blocks = (n_words + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  codes_from_planes_kernel<<<(u32)blocks, 256, 0, stream>>>(planes, codes, n_words);
  return cudaGetLastError();
}

// ASCII -> one ordinal byte per residue at the same offsets
__global__ void aa_codes_kernel(const uint8_t* __restrict__ chars, uint8_t* __restrict__ out, u64 n) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint8_t)aa_code(chars[i]);
}

cudaError_t launch_aa_codes(const uint8_t* chars, uint8_t* out, u64 n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  u64 blocks = (n + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  aa_codes_kernel<<<(u32)blocks, 256, 0, stream>>>(chars, out, n);
  return cudaGetLastError();
}

// 5-bit packing of an ordinal byte stream: word w of sequence s holds residues 6w..6w+5
__global__ void pack_aa_kernel(const uint8_t* __restrict__ chars, const u64* __restrict__ off,
                               const u64* __restrict__ woff, const u32* __restrict__ len, u32 n_seq,
                               u64 total_words, u32* __restrict__ packed) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 gw = (u64)blockIdx.x * blockDim.x + threadIdx.x; gw < total_words; gw += stride) {
    u32 lo = 0, hi = n_seq;
    while (hi - lo > 1) {
      u32 mid = (lo + hi) >> 1;
      if (woff[mid] <= gw) lo = mid; else hi = mid;
    }
    const u64 r0 = (gw - woff[lo]) * 6;
    u32 w = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (r0 + k < len[lo]) w |= aa_code(chars[off[lo] + r0 + k]) << (5 * k);
    packed[gw] = w;
  }
}

// which of the 27 residue ordinals occur in a packed store (bit r of *mask; unused fields of a sequence's last word
// read as ordinal 0) -- sizes the query profile of the protein kernel
__global__ void aa_mask_kernel(const u32* __restrict__ packed, u64 n_words, u32* __restrict__ mask) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  u32 m = 0;
  for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    const u32 x = packed[w];
#pragma unroll
    for (int k = 0; k < 6; ++k) m |= 1u << ((x >> (5 * k)) & 31u);
  }
  m = __reduce_or_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicOr(mask, m);
}

cudaError_t launch_aa_mask(const u32* packed, u64 n_words, u32* mask, cudaStream_t stream) {
  if (n_words == 0) return cudaSuccess;
  u64 blocks = (n_words + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  aa_mask_kernel<<<(u32)blocks, 256, 0, stream>>>(packed, n_words, mask);
  return cudaGetLastError();
}

cudaError_t launch_pack_aa(const uint8_t* chars, const u64* off, const u64* woff, const u32* len, u32 n_seq,
                           u64 total_words, u32* packed, cudaStream_t stream) {
  if (total_words == 0) return cudaSuccess;
  u64 blocks = (total_words + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_aa_kernel<<<(u32)blocks, 256, 0, stream>>>(chars, off, woff, len, n_seq, total_words, packed);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ staging
// 32 consecutive bits of a plane starting at (possibly negative) bit position s relative to word
// pointer p; bits below 0 read as 0.  The store arrays carry one pad word at the end.
__device__ __forceinline__ u32 window_x(const uint2* p, long long s) {
  if (s < 0) { const u32 sh = (u32)(-s); return sh >= 32 ? 0u : (p[0].x << sh); }
  const u64 w = (u64)s >> 5; const u32 sh = (u32)s & 31;
  return sh ? __funnelshift_r(p[w].x, p[w + 1].x, sh) : p[w].x;
}
__device__ __forceinline__ u32 window_y(const uint2* p, long long s) {
  if (s < 0) { const u32 sh = (u32)(-s); return sh >= 32 ? 0u : (p[0].y << sh); }
  const u64 w = (u64)s >> 5; const u32 sh = (u32)s & 31;
  return sh ? __funnelshift_r(p[w].y, p[w + 1].y, sh) : p[w].y;
}
__device__ __forceinline__ u32 window_n(const u32* p, long long s) {
  if (s < 0) { const u32 sh = (u32)(-s); return sh >= 32 ? 0u : (p[0] << sh); }
  const u64 w = (u64)s >> 5; const u32 sh = (u32)s & 31;
  return sh ? __funnelshift_r(p[w], p[w + 1], sh) : p[w];
}

// one warp per request
__global__ void stage_nt_kernel(const StageReq* __restrict__ reqs, u32 n_req,
                                const uint2* __restrict__ q_planes, const u32* __restrict__ q_n, const u64* __restrict__ q_woff,
                                const uint2* __restrict__ r_planes, const u32* __restrict__ r_n, const u64* __restrict__ r_woff,
                                SeqDesc* __restrict__ descs, uint2* __restrict__ out_planes, u32* __restrict__ out_n,
                                uint2* __restrict__ out_codes) {
  const u32 lane = threadIdx.x & 31;
  const u32 warps = (gridDim.x * blockDim.x) >> 5;
  for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_req; r += warps) {
    const StageReq rq = reqs[r];
    const SeqDesc d = descs[rq.desc];
    const uint2* sp = rq.store ? r_planes + r_woff[rq.seq] : q_planes + q_woff[rq.seq];
    const u32* sn = rq.store ? r_n + r_woff[rq.seq] : q_n + q_woff[rq.seq];
    const u32 nwords = (d.len + 31) >> 5;
    u32 anyn = 0;
    // 32 consecutive bits of the three planes starting at bit position s of the source (one 8-byte and one
    // 4-byte load per source word instead of one load per plane; bits below 0 read as 0)
    auto window3 = [&](long long s, u32& w0, u32& w1, u32& wn) {
      if (s < 0) {
        const u32 sh = (u32)(-s);
        if (sh >= 32) { w0 = w1 = wn = 0u; return; }
        const uint2 a = sp[0];
        w0 = a.x << sh; w1 = a.y << sh; wn = sn[0] << sh;
        return;
      }
      const u64 w = (u64)s >> 5; const u32 sh = (u32)s & 31;
      const uint2 a = sp[w];
      const u32 na = sn[w];
      if (!sh) { w0 = a.x; w1 = a.y; wn = na; return; }
      const uint2 b = sp[w + 1];
      const u32 nb = sn[w + 1];
      w0 = __funnelshift_r(a.x, b.x, sh); w1 = __funnelshift_r(a.y, b.y, sh); wn = __funnelshift_r(na, nb, sh);
    };
#pragma unroll 2
    for (u32 k = lane; k < nwords; k += 32) {
      const u32 rem = d.len - 32 * k;
      const u32 valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
      u32 p0, p1, pn;
      if (!rq.rev) {
        window3((long long)rq.begin + 32ll * k, p0, p1, pn);
      } else {
        // out base j = complement(src[begin + len - 1 - j])
        window3((long long)rq.begin + (long long)d.len - 32ll * k - 32ll, p0, p1, pn);
        p0 = ~__brev(p0); p1 = ~__brev(p1); pn = __brev(pn);
      }
      pn &= valid;
      p0 &= valid & ~pn; p1 &= valid & ~pn;
      out_planes[(u64)d.woff + k] = make_uint2(p0, p1);
      out_n[(u64)d.woff + k] = pn;
      if (out_codes) out_codes[(u64)d.woff + k] = nt_codes(p0, p1);
      anyn |= pn;
    }
    anyn = __reduce_or_sync(0xffffffffu, anyn);
    if (lane == 0) descs[rq.desc].flags = anyn ? 1u : 0u;
  }
}

cudaError_t launch_stage_nt(const StageReq* reqs, u32 n_req, const uint2* q_planes, const u32* q_n, const u64* q_woff,
                            const uint2* r_planes, const u32* r_n, const u64* r_woff, SeqDesc* descs,
                            uint2* out_planes, u32* out_n, uint2* out_codes, cudaStream_t stream) {
  if (n_req == 0) return cudaSuccess;
  u32 blocks = (n_req + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stage_nt_kernel<<<blocks, 256, 0, stream>>>(reqs, n_req, q_planes, q_n, q_woff, r_planes, r_n, r_woff, descs,
                                              out_planes, out_n, out_codes);
  return cudaGetLastError();
}

// AA staging: unpack 5-bit residues to bytes; SeqDesc.pad receives the self score
// sum_i BLOSUM62(a_i,a_i) (== NW(X,X), hh:190-191; equality pinned in tests).  One warp per request.
__global__ void stage_aa_kernel(const StageReq* __restrict__ reqs, u32 n_req, const u32* __restrict__ q_packed,
                                const u64* __restrict__ q_woff, const u32* __restrict__ r_packed,
                                const u64* __restrict__ r_woff, SeqDesc* __restrict__ descs,
                                uint8_t* __restrict__ out) {
  const u32 lane = threadIdx.x & 31;
  const u32 warps = (gridDim.x * blockDim.x) >> 5;
  for (u32 r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_req; r += warps) {
    const StageReq rq = reqs[r];
    const SeqDesc d = descs[rq.desc];
    const u32* sp = rq.store ? r_packed + r_woff[rq.seq] : q_packed + q_woff[rq.seq];
    int self = 0;
    // BLOSUM62(x, x) of ordinal `lane` sits in lane `lane`: one shuffle per residue instead of a constant-memory
    // load whose address differs from lane to lane (serialised by the constant cache)
    const int diag = lane < 27 ? (int)c_blosum[lane][lane] : 0;
    const u32 rounds = (d.len + 31u) >> 5;
    for (u32 it = 0; it < rounds; ++it) {
      const u32 k = it * 32u + lane;
      u32 code = 31u;
      if (k < d.len) {
        const u32 idx = rq.begin + k;
        code = (sp[idx / 6] >> (5 * (idx % 6))) & 31u;
        out[(u64)d.woff + k] = (uint8_t)code;
      }
      self += __shfl_sync(0xffffffffu, diag, code);   // lanes past the end read lane 31: 0
    }
    self = __reduce_add_sync(0xffffffffu, self);
    if (lane == 0) { descs[rq.desc].pad = (u32)self; descs[rq.desc].flags = 0; }
  }
}

cudaError_t launch_stage_aa(const StageReq* reqs, u32 n_req, const u32* q_packed, const u64* q_woff,
                            const u32* r_packed, const u64* r_woff, SeqDesc* descs, uint8_t* out, cudaStream_t stream) {
  if (n_req == 0) return cudaSuccess;
  u32 blocks = (n_req + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stage_aa_kernel<<<blocks, 256, 0, stream>>>(reqs, n_req, q_packed, q_woff, r_packed, r_woff, descs, out);
  return cudaGetLastError();
}

// self scores for a table of byte sequences (low-level API); one warp per sequence
__global__ void selfscore_kernel(SeqDesc* __restrict__ descs, u32 n, const uint8_t* __restrict__ residues) {
  const u32 lane = threadIdx.x & 31;
  const u32 warps = (gridDim.x * blockDim.x) >> 5;
  for (u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const SeqDesc d = descs[s];
    int self = 0;
    const int diag = lane < 27 ? (int)c_blosum[lane][lane] : 0;   // BLOSUM62(x, x) of ordinal `lane` (see stage_aa_kernel)
    for (u32 it = 0; it * 32u < d.len; ++it) {
      const u32 k = it * 32u + lane;
      const u32 c = k < d.len ? (u32)residues[(u64)d.woff + k] : 31u;
      self += __shfl_sync(0xffffffffu, diag, c);
    }
    self = __reduce_add_sync(0xffffffffu, self);
    if (lane == 0) descs[s].pad = (u32)self;
  }
}

cudaError_t launch_selfscore(SeqDesc* descs, u32 n, const uint8_t* residues, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  u32 blocks = (n + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  selfscore_kernel<<<blocks, 256, 0, stream>>>(descs, n, residues);
  return cudaGetLastError();
}

// staged NT planes -> ordinal bytes (only for trpa_fetch_segments, the byte-for-byte fetch test)
__global__ void unstage_nt_kernel(const SeqDesc* __restrict__ descs, u32 n, const uint2* __restrict__ planes,
                                  const u32* __restrict__ nplane, const u64* __restrict__ out_off,
                                  uint8_t* __restrict__ out) {
  const u32 lane = threadIdx.x & 31;
  const u32 warps = (gridDim.x * blockDim.x) >> 5;
  for (u32 s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n; s += warps) {
    const SeqDesc d = descs[s];
    for (u32 j = lane; j < d.len; j += 32) {
      const uint2 p = planes[(u64)d.woff + (j >> 5)];
      const u32 nn = nplane[(u64)d.woff + (j >> 5)];
      const u32 b = j & 31;
      u32 code = ((p.x >> b) & 1u) | (((p.y >> b) & 1u) << 1);
      if ((nn >> b) & 1u) code = 4;
      out[out_off[s] + j] = (uint8_t)code;
    }
  }
}

cudaError_t launch_unstage_nt(const SeqDesc* descs, u32 n, const uint2* planes, const u32* nplane,
                              const u64* out_off, uint8_t* out, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  u32 blocks = (n + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  unstage_nt_kernel<<<blocks, 256, 0, stream>>>(descs, n, planes, nplane, out_off, out);
  return cudaGetLastError();
}

}  // namespace trpa
