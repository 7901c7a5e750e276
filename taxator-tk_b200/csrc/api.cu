// C ABI of the B200-native RPA hot path (include/taxator_rpa_b200.h) + the per-round driver:
//   decide kernel (one thread per query segment, machine.h) -> stage kernel (pack.cu)
//   -> shape bucketing -> edit-distance / protein kernels (myers3.cuh / protein3.cu) -> decide ...
// There is no CPU fallback: every compute entry point needs a CUDA device.
// Compile with --fmad=false (decision arithmetic must match the reference's IEEE float/double ops).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "launch.h"
#include "machine.h"
#include "hostprep.h"

namespace trpa {

static thread_local std::string g_error;
static std::atomic<int> g_first_device{-1};   // device of the first context of this process (trpa_create)
void set_error(const std::string& s) { g_error = s; }
int alu_probe_ops_per_iter();

#define CK(expr)                                                                        \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + " " + #expr + ": " + cudaGetErrorString(_e)); \
      return TRPA_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

// host-side parallel loop over [0, n) in contiguous blocks (batch preparation is memory bound and
// would otherwise cost more than the H2D copy of a 5 M-candidate batch)
// `work`: items the whole loop touches (a table of 50 k segments x 30 records is worth 16 threads, 50 k bare items
// are not)
template <class F>
static void parallel_blocks(size_t n, size_t work, const F& f) {
  unsigned T = std::thread::hardware_concurrency();
  if (T > 16) T = 16;
  if (T < 2 || n < 64 || work < 262144) { f(0, n); return; }
  std::vector<std::thread> th;
  const size_t per = (n + T - 1) / T;
  for (unsigned t = 0; t < T; ++t) {
    const size_t b = t * per, e = std::min(n, b + per);
    if (b >= e) break;
    th.emplace_back([&f, b, e]() { f(b, e); });
  }
  for (auto& x : th) x.join();
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  int ensure(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 16;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc ") + std::to_string(want * sizeof(T)) + " B: " + cudaGetErrorString(e)); cudaGetLastError(); return TRPA_ERR_NOMEM; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// a DevBuf local to one call: released on every return path
template <class T>
struct ScopedBuf : DevBuf<T> {
  ~ScopedBuf() { this->release(); }
};

struct Store {
  int alphabet = -1;
  u32 n_seq = 0;
  u64 n_words = 0;
  DevBuf<uint2> planes;
  DevBuf<u32> nplane;
  DevBuf<u32> packed;  // AA
  DevBuf<u64> woff;
  DevBuf<u32> len;
  u32 max_len = 0;
  u32 aa_mask = 0;          // AA: residue ordinals that occur in the store (profile rows of the protein kernel)
  std::vector<u32> h_len;   // host copy of the sequence lengths (upload validation)
  void release() { planes.release(); nplane.release(); packed.release(); woff.release(); len.release(); n_seq = 0; alphabet = -1; aa_mask = 0; h_len.clear(); }
};

struct EventPair { cudaEvent_t a, b; int kind; };

// One sub-batch pipeline: its own stream, round queues' counters, shape histogram / work cursors,
// kernel scratch and timing events.  A batch is cut into chunks of segments; the pipes of a context
// work on different chunks at the same time, so that the tail of one chunk's alignment kernels, its
// small decide / plan kernels and its host round trips overlap another chunk's alignments.
struct Pipe {
  static constexpr int kAux = 6;   // shape buckets of a round run concurrently on these streams
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool ready = false;
  DevBuf<u32> d_counters;
  DevBuf<u32> d_hist;         // kNumShapes counts + scatter cursors + kernel work cursors, then (banded path)
                              // kNumShapes x kNumCls duration-class counts and their scatter cursors
  DevBuf<uint2> d_buckets;    // kNumShapes {start,count}
  DevBuf<uint4> scratch3;     // banded kernel: wrap-around strip boundaries, one line per group slot
  DevBuf<unsigned long long> d_plan;  // [1..3] {word-blocks, retries, pairs} of myers3
  DevBuf<int2> scratch_aa;
  cudaStream_t aux[kAux] = {};
  cudaEvent_t aux_done[kAux] = {};
  cudaEvent_t fork_ev = nullptr;
  u32* h_counters = nullptr;  // pinned: kNumCounters round counters + kNumShapes shape histogram
  std::vector<EventPair> ev_pool;
  size_t ev_used = 0;
  // run state of trpa_batch_run
  int state = 0;
  u32 sb = 0, se = 0, n_pairs = 0;
  PairDesc* pairs = nullptr;
  PairDesc* sorted = nullptr;
  Batch B;

  int init(cudaStream_t user_stream);
  void release();
};

}  // namespace trpa

using namespace trpa;

struct trpa_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  float exclude_factor = 0.5f, toppercent = 0.05f;
  u64 arena_bytes = 0;
  // taxonomy
  DevBuf<u32> t_parent, t_left, t_right;
  DevBuf<uint8_t> t_depth;
  u32 n_nodes = 0, root = 0, max_depth = 0;
  Store store[2];
  // batch (whole batch resident)
  std::vector<u32> chunk_begin;  // segment index boundaries
  std::vector<u64> chunk_qoff;   // start of the chunk's slice of the pair / stage queues
  int run_pipes = 1;             // pipes the uploaded batch was planned for
  u32 n_segs = 0, n_cands = 0;
  u32 max_stage_len = 0;
  u32 avg_stage_len = 0;         // average length of the segments the batch can stage (look-ahead budget)
  bool batch_ready = false;
  DevBuf<trpa_segment> d_segs;
  DevBuf<trpa_candidate> d_cands;      // SortFilter order
  DevBuf<trpa_candidate> d_cands_raw;  // as uploaded
  DevBuf<trpa_result> d_results;
  DevBuf<SegState> d_state;
  DevBuf<float> d_qd, d_qsim, d_bf_d;
  DevBuf<uint8_t> d_cflags;
  DevBuf<u32> d_og_i, d_bf_node, d_tag;
  int lookahead = -1;         // -1: automatic (fill idle capacity), >= 0: fixed budget
  DevBuf<int32_t> d_og_d;
  DevBuf<int32_t> d_res;      // NT: 1 int per slot; AA: 2 ints per slot
  DevBuf<SeqDesc> d_descs;
  DevBuf<PairDesc> d_pairs, d_pairs_sorted;
  DevBuf<StageReq> d_stage;
  DevBuf<uint2> arena_planes;
  DevBuf<uint2> arena_codes;     // column codes of the staged words (common.cuh nt_codes)
  // tables of trpa_predict_lca_batch (kept between calls: no allocation in the steady state)
  DevBuf<trpa_segment> l_segs; DevBuf<trpa_candidate> l_cands; DevBuf<double> l_ev; DevBuf<uint8_t> l_un; DevBuf<trpa_result> l_out;
  DevBuf<u32> arena_n;
  DevBuf<uint8_t> arena_aa;
  u64 arena_units = 0;
  // alignment trace (verbose log): on/off, device buffer, entries recorded by the last run
  int trace_on = 0;
  DevBuf<trpa_trace_entry> d_trace;
  u64 trace_used = 0;
  bool trace_overflow = false;
  int band = 1;               // 1: Ukkonen band (exact), 0: full DP matrix
  int wedge = 1;              // 1: let the band narrow along the matrix where a certificate can prove it (exact)
  // pipes: independent sub-batch pipelines whose rounds overlap on the GPU (pipe[0] runs on `stream`)
  static constexpr int kMaxPipes = 4;
  Pipe pipe[kMaxPipes];
  int n_pipes = 1;   // measured on C2: overlapping chunks gains nothing (the alignment kernels already saturate the GPU)
  u32 band_k0 = 0;            // test hook: forced initial threshold (0 = planned), exercises the retry loop
  u32 la_cap = 0;             // pairs per round the automatic look-ahead aims for; 0 = scaled with the chunk: 5.5 per
                              // segment within [150 k, 750 k] (measured, DESIGN.md section 6: C2 prefers 550 k for 100 k
                              // segments, the 250 k-segment C4 shard 750 k, the 50 k-segment protein batch 150-300 k)
  u32 la_max = 32;            // largest automatic look-ahead budget per segment and round
  int force_shape = -1;       // tuning hook: (lidx * kNumW + widx) forced for every pair, -1 = planner
  PlanParams plan;            // hint margin + cost model of the shape planner (tuning hooks)
  u32 plan_lanes = 0;         // tuning hook: weight of a pair's latency against the summed lane-time in the shape planner (0 = automatic, see bucket_pairs3)
  int num_sms = 148;
  // profiling
  trpa_profile prof;
};

namespace trpa {

// ------------------------------------------------------------------------------------ kernels
// per_warp: segments (= working lanes, spread evenly) per warp.  The threads of a warp follow different paths through
// predict()'s loops and serialise each other's chains of dependent loads (ncu: 1.96 active threads per executed warp
// instruction with 32 segments per warp); fewer segments per warp trade idle lanes for independent instruction streams.
__global__ void __launch_bounds__(64, 12) decide_kernel(Batch B, u32 seg_begin, u32 seg_end, u32 per_warp) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 stride = 32u / per_warp, lane = t & 31u;
  if (lane % stride) return;
  const u32 s = seg_begin + (t >> 5) * per_warp + lane / stride;
  if (s >= seg_end) return;
  if (B.st[s].phase == PH_DONE) return;
  SegState st = B.st[s];   // one copy in, one copy out instead of a global access per field
  Machine M(B, s, st);
  M.advance();
  B.st[s] = st;
  if (st.phase != PH_DONE) atomicAdd(&B.counters[CN_ACTIVE], 1u);
}

__global__ void init_state_kernel(SegState* st, u32 n) {
  const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) st[s].phase = PH_INIT;
}

// banded path: buckets by shape and, inside a shape, by duration class, longest first (the persistent groups pull
// pairs in this order: longest-processing-time-first keeps the tail of a launch short)
constexpr u32 kHistWords = 3u * kNumShapes + 2u * kNumShapes * kNumCls;
__global__ void scan3_kernel(u32* hist, uint2* buckets) {
  // exclusive prefix sum of the shape counts, all threads at once (it was one thread's loop of kNumShapes dependent
  // steps: 8.5 us per round, which a 1 k-segment batch feels)
  static_assert(kNumShapes <= 128, "one thread per shape in a 128-thread block");
  __shared__ u32 start[kNumShapes];
  __shared__ u32 warp_tot[4];
  const int s = threadIdx.x;
  const u32 mine = s < kNumShapes ? hist[s] : 0u;
  u32 incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((s & 31) >= o) incl += v;
  }
  if ((s & 31) == 31) warp_tot[s >> 5] = incl;
  __syncthreads();
  u32 base = 0;
  for (int w = 0; w < (s >> 5); ++w) base += warp_tot[w];
  if (s < kNumShapes) {
    start[s] = base + incl - mine;
    buckets[s] = make_uint2(start[s], mine);
  }
  __syncthreads();
  if (s < kNumShapes) {
    const u32* cnt = hist + 3 * kNumShapes + s * kNumCls;
    u32* cur = hist + 3 * kNumShapes + kNumShapes * kNumCls + s * kNumCls;
    u32 acc = start[s];
    for (int c = kNumCls - 1; c >= 0; --c) { cur[c] = acc; acc += cnt[c]; }
  }
}
__global__ void scatter3_kernel(const PairDesc* pairs, u32 n, u32* hist, PairDesc* sorted) {
  u32* cur = hist + 3 * kNumShapes + kNumShapes * kNumCls;
  for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const PairDesc p = pairs[k];
    const u32 cls = p.cls < (u32)kNumCls ? p.cls : (u32)kNumCls - 1u;
    const u32 key = (p.pad & 0xffu) * kNumCls + cls;   // pad: shape | k0 << 8
    // one atomic per distinct (shape, class) of the warp: most pairs of a round share a handful of keys
    const u32 active = __activemask();
    const u32 same = __match_any_sync(active, key);
    const u32 leader = __ffs(same) - 1u, lane = threadIdx.x & 31u;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(&cur[key], (u32)__popc(same));
    base = __shfl_sync(same, base, leader);
    sorted[base + __popc(same & ((1u << lane) - 1u))] = p;
  }
}
// SortFilter on the device (core/src/alignmentsfilter.hh:171-190: stable sort, descending (score,
// identities)): one warp per segment ranks every record against all others of its segment; the
// original position breaks ties, which is exactly what a stable sort does.  raw -> sorted table.
__global__ void sort_cands_kernel(const trpa_segment* __restrict__ segs, u32 n_segs,
                                  const trpa_candidate* __restrict__ raw, trpa_candidate* __restrict__ out) {
  const u32 lane = threadIdx.x & 31;
  const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 s = warp; s < n_segs; s += nwarps) {
    const trpa_segment sg = segs[s];
    const trpa_candidate* in = raw + sg.cand_begin;
    for (u32 i = lane; i < sg.cand_count; i += 32) {
      const trpa_candidate x = in[i];
      u32 rank = 0;
      for (u32 j = 0; j < sg.cand_count; ++j) {
        const float sj = in[j].score;
        const u32 ij = in[j].identities;
        const bool before = sj > x.score || (sj == x.score && (ij > x.identities || (ij == x.identities && j < i)));
        rank += before ? 1u : 0u;
      }
      out[sg.cand_begin + rank] = x;
    }
  }
}

__global__ void lca_kernel(Taxonomy T, const u32* a, const u32* b, u32 n, u32* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tx_lca(T, a[i], b[i]);
}

// ------------------------------------------------------------------------------------ helpers
static int begin_event(Pipe& P, int kind) {
  if (P.ev_used == P.ev_pool.size()) {
    EventPair e; e.kind = kind;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return -1;
    P.ev_pool.push_back(e);
  }
  P.ev_pool[P.ev_used].kind = kind;
  cudaEventRecord(P.ev_pool[P.ev_used].a, P.stream);
  return (int)P.ev_used++;
}
static void end_event(Pipe& P, int id) {
  if (id >= 0) cudaEventRecord(P.ev_pool[id].b, P.stream);
}
enum { EV_MYERS = 0, EV_PROTEIN, EV_STAGE, EV_DECIDE, EV_OTHER };
// call after a stream sync
static void harvest_events(trpa_ctx* c, Pipe& P) {
  for (size_t i = 0; i < P.ev_used; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, P.ev_pool[i].a, P.ev_pool[i].b) != cudaSuccess) { cudaGetLastError(); continue; }
    switch (P.ev_pool[i].kind) {
      case EV_MYERS: c->prof.ms_edit_distance += ms; break;
      case EV_PROTEIN: c->prof.ms_protein += ms; break;
      case EV_STAGE: c->prof.ms_stage += ms; break;
      case EV_DECIDE: c->prof.ms_decide += ms; break;
      default: c->prof.ms_other += ms; break;
    }
  }
  P.ev_used = 0;
}

static int use_device(trpa_ctx* c) {
  CK(cudaSetDevice(c->device));
  return 0;
}

int Pipe::init(cudaStream_t user_stream) {
  if (ready) return 0;
  if (user_stream) { stream = user_stream; own_stream = false; }
  else {
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
  }
  CK(cudaMallocHost(&h_counters, sizeof(u32) * (kNumCounters + 2 * kNumShapes)));
  if (d_counters.ensure(kNumCounters) || d_hist.ensure(kHistWords) || d_buckets.ensure(kNumShapes) || d_plan.ensure(8)) return TRPA_ERR_NOMEM;
  CK(cudaMemsetAsync(d_plan.p, 0, 8 * sizeof(unsigned long long), stream));
  CK(cudaStreamSynchronize(stream));
  ready = true;
  return 0;
}
void Pipe::release() {
  if (stream) cudaStreamSynchronize(stream);
  d_counters.release(); d_hist.release(); d_buckets.release(); scratch3.release();
  d_plan.release(); scratch_aa.release();
  for (auto& e : ev_pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  ev_pool.clear();
  if (fork_ev) {
    cudaEventDestroy(fork_ev);
    for (int i = 0; i < kAux; ++i) { cudaStreamDestroy(aux[i]); cudaEventDestroy(aux_done[i]); }
    fork_ev = nullptr;
  }
  if (h_counters) cudaFreeHost(h_counters);
  h_counters = nullptr;
  if (own_stream && stream) cudaStreamDestroy(stream);
  stream = nullptr;
  ready = false;
}

static Taxonomy dev_tax(trpa_ctx* c) { return Taxonomy{c->t_parent.p, c->t_left.p, c->t_right.p, c->t_depth.p, c->root}; }

// ---- plan (threshold + shape per pair), counting sort by shape, one launch per shape
static int bucket_pairs3(trpa_ctx* c, Pipe& P, PairDesc* pairs, u32 n_pairs, const SeqDesc* descs, const uint2* planes,
                         const u32* nplane, PairDesc* sorted, u32* h_hist) {
  CK(cudaMemsetAsync(P.d_hist.p, 0, sizeof(u32) * kHistWords, P.stream));
  const u32 blocks = std::min<u32>((n_pairs + 255) / 256, 148 * 8);
  // weight of a pair's own latency against the summed lane-time = the resident lanes; the planner makes the latency
  // term convex (long pairs are the tail of a round, see plan_kernel)
  const u32 lat_weight = c->plan_lanes ? c->plan_lanes : (u32)c->num_sms * 16u * 32u;
  CK(launch_plan(pairs, n_pairs, descs, planes, nplane, P.d_hist.p, lat_weight,
                 c->band ? (c->band_k0 ? (int)c->band_k0 : 1) : 0, c->force_shape, c->wedge, c->plan, P.stream));
  scan3_kernel<<<1, 128, 0, P.stream>>>(P.d_hist.p, P.d_buckets.p);
  scatter3_kernel<<<blocks, 256, 0, P.stream>>>(pairs, n_pairs, P.d_hist.p, sorted);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h_hist, P.d_hist.p, sizeof(u32) * kNumShapes, cudaMemcpyDeviceToHost, P.stream));
  return 0;   // asynchronous: synchronise P.stream before reading h_hist
}

static int launch_myers_shapes3(trpa_ctx* c, Pipe& P, const u32* h_hist, const PairDesc* sorted, const SeqDesc* descs,
                                const uint2* planes, const u32* nplane, const uint2* codes, int* out, u32 max_len) {
  const u32 stride = (max_len + 31) / 32 + 1;
  CK(cudaMemsetAsync(P.d_hist.p + 2 * kNumShapes, 0, sizeof(u32) * kNumShapes, P.stream));
  // every shape bucket is one persistent launch; the buckets run concurrently (own stream, own
  // scratch region) so that the tail of one bucket overlaps the others
  struct Job { int shape; u32 start, cnt, slots; size_t scr; };
  Job jobs[kNumShapes];
  int nj = 0;
  u32 start = 0;
  size_t scr_total = 0;
  for (int shape = 0; shape < kNumShapes; ++shape) {
    const u32 cnt = h_hist[shape];
    if (!cnt) continue;
    u32 slots = 0;
    CK(launch_myers3(shape, nullptr, cnt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, &slots, P.stream));
    jobs[nj++] = Job{shape, start, cnt, slots, scr_total};
    scr_total += (size_t)slots * stride;
    start += cnt;
  }
  if (P.scratch3.ensure(scr_total + 1)) return TRPA_ERR_NOMEM;
  static const bool debug = getenv("TRPA_DEBUG") != nullptr;
  if (debug) {
    fprintf(stderr, "[trpa] myers3 round:");
    for (int j = 0; j < nj; ++j)
      fprintf(stderr, " W%dxL%d%s:%u", shape_W(shape_widx(jobs[j].shape)), 1 << shape_lidx(jobs[j].shape),
              shape_hasn(jobs[j].shape) ? "N" : "", jobs[j].cnt);
    fprintf(stderr, "\n");
  }
  // SMALLEST buckets first: the big bucket's persistent launch takes every CTA slot of the GPU, a small latency-bound
  // bucket (1-180 CTAs, 0.6-0.9 ms on its own) launched behind it only starts in its tail and ends the round late;
  // launched first it runs beside the big one (C2: 331.8 -> 328.5 ms per step, C4 neutral; TRPA_SHAPE_ORDER=desc: A/B hook)
  static const bool asc = [] { const char* e = getenv("TRPA_SHAPE_ORDER"); return !(e && e[0] == 'd'); }();
  if (asc) std::sort(jobs, jobs + nj, [](const Job& x, const Job& y) { return x.cnt < y.cnt; });
  else std::sort(jobs, jobs + nj, [](const Job& x, const Job& y) { return x.cnt > y.cnt; });
  const bool fork = nj > 1;
  if (fork) {
    if (!P.fork_ev) {
      CK(cudaEventCreateWithFlags(&P.fork_ev, cudaEventDisableTiming));
      for (int i = 0; i < Pipe::kAux; ++i) {
        CK(cudaStreamCreateWithFlags(&P.aux[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&P.aux_done[i], cudaEventDisableTiming));
      }
    }
    CK(cudaEventRecord(P.fork_ev, P.stream));
  }
  int used = 0;
  for (int j = 0; j < nj; ++j) {
    cudaStream_t st = P.stream;
    if (fork && j > 0) {
      const int a = (j - 1) % Pipe::kAux;
      st = P.aux[a];
      if (a >= used) { CK(cudaStreamWaitEvent(st, P.fork_ev, 0)); used = a + 1; }
    }
    CK(launch_myers3(jobs[j].shape, sorted + jobs[j].start, jobs[j].cnt, descs, planes, nplane, codes, out,
                     P.scratch3.p + jobs[j].scr, stride, P.d_hist.p + 2 * kNumShapes + jobs[j].shape, P.d_plan.p + 1,
                     c->band ? 0 : 1, nullptr, st));
    c->prof.launches_edit_distance++;
  }
  for (int a = 0; a < used; ++a) {
    CK(cudaEventRecord(P.aux_done[a], P.aux[a]));
    CK(cudaStreamWaitEvent(P.stream, P.aux_done[a], 0));
  }
  if (debug) {   // per-round device time and executed cells (debug only: adds a sync)
    static cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    unsigned long long h0[3], h1[3];
    cudaStreamSynchronize(P.stream);
    cudaMemcpy(h1, P.d_plan.p + 1, sizeof(h1), cudaMemcpyDeviceToHost);
    static unsigned long long last[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) { h0[i] = h1[i] >= last[i] ? h1[i] - last[i] : h1[i]; last[i] = h1[i]; }
    static auto t_last = std::chrono::steady_clock::now();
    const auto t_now = std::chrono::steady_clock::now();
    fprintf(stderr, "[trpa]   executed %.3g cells, %llu retries, %llu pairs, %.3f ms since previous round\n",
            (double)h0[0] * 1024.0, h0[1], h0[2], std::chrono::duration<double, std::milli>(t_now - t_last).count());
    t_last = t_now;
  }
  return 0;
}

// {word-blocks, retries, pairs} of the banded kernel since the last call -> profile
static int harvest_band_stats(trpa_ctx* c, Pipe& P) {
  if (!P.d_plan.p) return 0;
  unsigned long long h[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(h, P.d_plan.p + 1, sizeof(h), cudaMemcpyDeviceToHost, P.stream));
  CK(cudaStreamSynchronize(P.stream));
  CK(cudaMemsetAsync(P.d_plan.p + 1, 0, sizeof(h), P.stream));
  c->prof.cells_edit_distance += h[0] * 1024ull;
  c->prof.band_retries += h[1];
  c->prof.wedge_failures += h[3];
  return 0;
}

}  // namespace trpa

// =========================================================================================== ABI
extern "C" {

int trpa_abi_version(void) { return TRPA_ABI_VERSION; }
const char* trpa_last_error(void) { return g_error.c_str(); }

trpa_ctx* trpa_create(int device, void* cuda_stream) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0 || device >= ndev) { set_error("bad device index"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
  trpa_ctx* c = new trpa_ctx();
  c->device = device;
  { int none = -1; g_first_device.compare_exchange_strong(none, device); }
  memset(&c->prof, 0, sizeof(c->prof));
  if (c->pipe[0].init((cudaStream_t)cuda_stream)) { delete c; return nullptr; }
  c->stream = c->pipe[0].stream;
  c->own_stream = false;   // owned by pipe[0]
  if (cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->num_sms <= 0) c->num_sms = 148;
  if (const char* e = getenv("TRPA_BAND")) c->band = e[0] != '0';
  if (const char* e = getenv("TRPA_WEDGE")) c->wedge = e[0] != '0';
  if (const char* e = getenv("TRPA_PIPES")) c->n_pipes = std::max(1, std::min((int)trpa_ctx::kMaxPipes, atoi(e)));
  return c;
}

void trpa_destroy(trpa_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->t_parent.release(); c->t_left.release(); c->t_right.release(); c->t_depth.release();
  c->store[0].release(); c->store[1].release();
  c->d_segs.release(); c->d_cands.release(); c->d_cands_raw.release(); c->d_results.release(); c->d_state.release();
  c->d_qd.release(); c->d_qsim.release(); c->d_bf_d.release(); c->d_cflags.release(); c->d_og_i.release(); c->d_tag.release();
  c->d_bf_node.release(); c->d_og_d.release(); c->d_res.release(); c->d_descs.release(); c->d_pairs.release();
  c->d_pairs_sorted.release(); c->d_stage.release(); c->d_trace.release();
  c->arena_planes.release(); c->arena_codes.release(); c->arena_n.release();
  c->l_segs.release(); c->l_cands.release(); c->l_ev.release(); c->l_un.release(); c->l_out.release(); c->arena_aa.release();
  for (int i = trpa_ctx::kMaxPipes - 1; i >= 0; --i) c->pipe[i].release();
  delete c;
}

int trpa_set_params(trpa_ctx* c, float exclude_factor, float toppercent) {
  if (!c) { set_error("null ctx"); return TRPA_ERR_ARG; }
  c->exclude_factor = exclude_factor; c->toppercent = toppercent;
  return 0;
}
int trpa_set_arena_bytes(trpa_ctx* c, uint64_t bytes) {
  if (!c) { set_error("null ctx"); return TRPA_ERR_ARG; }
  c->arena_bytes = bytes;
  return 0;
}
int trpa_set_lookahead(trpa_ctx* c, int k) {
  if (!c) { set_error("null ctx"); return TRPA_ERR_ARG; }
  c->lookahead = k < 0 ? -1 : k;
  return 0;
}
int trpa_set_band(trpa_ctx* c, int on) {
  if (!c) { set_error("null ctx"); return TRPA_ERR_ARG; }
  c->band = on ? 1 : 0;
  return 0;
}
int trpa_set_tuning(trpa_ctx* c, const char* key, int64_t value) {
  if (!c || !key) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  const std::string k(key);
  if (k == "band_k0") c->band_k0 = value < 0 ? 0u : (u32)std::min<int64_t>(value, 0xfffffe);
  else if (k == "hint_mul64") c->plan.hint_mul64 = (u32)std::max<int64_t>(1, std::min<int64_t>(value, 1 << 16));
  else if (k == "tail_log2") c->plan.tail_log2 = (u32)std::max<int64_t>(9, std::min<int64_t>(value, 40));
  else if (k == "hint_add") c->plan.hint_add = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20));
  else if (k == "cost_word10") c->plan.word10 = (u32)std::max<int64_t>(1, std::min<int64_t>(value, 1 << 16));
  else if (k == "cost_col10") c->plan.col10 = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 16));
  else if (k == "cost_step") c->plan.step = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20));
  else if (k == "cost_setup") c->plan.setup = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20));
  else if (k == "cost_setup_w") c->plan.setup_w = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 16));
  else if (k == "wedge_max_k") c->plan.wedge_max_k = (u32)std::max<int64_t>(64, std::min<int64_t>(value, 1 << 22));
  else if (k == "wedge_hint_s8") c->plan.wedge_hint_s8 = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 15));
  else if (k == "wedge_hint_max_k") c->plan.wedge_hint_max_k = (u32)std::max<int64_t>(64, std::min<int64_t>(value, 1 << 22));
  else if (k == "wedge_s8") c->plan.wedge_s8 = (u32)std::max<int64_t>(1, std::min<int64_t>(value, 15));
  else if (k == "wedge_e0") c->plan.wedge_e0 = (u32)std::max<int64_t>(1, std::min<int64_t>(value, 1 << 16));
  else if (k == "wedge_cushion") c->plan.wedge_cushion = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 16));
  else if (k == "plan_lanes") c->plan_lanes = value < 0 ? 0u : (u32)std::min<int64_t>(value, 1 << 30);
  else if (k == "wedge") c->wedge = value < 0 ? 0 : (value > 2 ? 2 : (int)value);   // 2: test hook, see plan_kernel
  else if (k == "la_cap") c->la_cap = (u32)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 30));
  else if (k == "la_max") c->la_max = (u32)std::max<int64_t>(0, std::min<int64_t>(64, value));
  else if (k == "force_shape") c->force_shape = value < 0 || value >= kNumW * kNumL ? -1 : (int)value;
  else if (k == "pipes") c->n_pipes = (int)std::max<int64_t>(1, std::min<int64_t>(trpa_ctx::kMaxPipes, value));
  else { set_error("unknown tuning key: " + k); return TRPA_ERR_ARG; }
  return 0;
}
int trpa_profile_reset(trpa_ctx* c) { if (!c) return TRPA_ERR_ARG; memset(&c->prof, 0, sizeof(c->prof)); return 0; }
int trpa_profile_get(trpa_ctx* c, trpa_profile* out) { if (!c || !out) return TRPA_ERR_ARG; *out = c->prof; return 0; }

int trpa_load_taxonomy(trpa_ctx* c, const uint32_t* parent, const uint32_t* left, const uint32_t* right,
                       const uint8_t* depth, uint32_t n_nodes, uint32_t root) {
  if (!c || !parent || !left || !right || !depth || n_nodes == 0 || root >= n_nodes) { set_error("bad taxonomy arguments"); return TRPA_ERR_ARG; }
  for (u32 i = 0; i < n_nodes; ++i) {
    if (parent[i] >= n_nodes) { set_error("taxonomy: parent index out of range"); return TRPA_ERR_ARG; }
    if (depth[i] >= 64) { set_error("taxonomy: depth >= 64 not supported"); return TRPA_ERR_ARG; }
  }
  if (parent[root] != root) { set_error("taxonomy: parent[root] must be root"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  if (c->t_parent.ensure(n_nodes) || c->t_left.ensure(n_nodes) || c->t_right.ensure(n_nodes) || c->t_depth.ensure(n_nodes)) return TRPA_ERR_NOMEM;
  CK(cudaMemcpyAsync(c->t_parent.p, parent, 4ull * n_nodes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->t_left.p, left, 4ull * n_nodes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->t_right.p, right, 4ull * n_nodes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->t_depth.p, depth, 1ull * n_nodes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->n_nodes = n_nodes; c->root = root;
  c->max_depth = 0;
  for (u32 i = 0; i < n_nodes; ++i) c->max_depth = std::max<u32>(c->max_depth, depth[i]);
  return 0;
}

// residue ordinals present in a packed AA store (one pass at load time)
static int aa_store_mask(trpa_ctx* c, Store& S, u64 n_words) {
  ScopedBuf<u32> d_mask;
  if (d_mask.ensure(1)) return TRPA_ERR_NOMEM;
  CK(cudaMemsetAsync(d_mask.p, 0, sizeof(u32), c->stream));
  CK(launch_aa_mask(S.packed.p, n_words, d_mask.p, c->stream));
  u32 h = 0;
  CK(cudaMemcpyAsync(&h, d_mask.p, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  S.aa_mask = (h | 1u) & 0x7ffffffu;
  return 0;
}

int trpa_load_store(trpa_ctx* c, int store, int alphabet, const char* chars, const uint64_t* off, const uint32_t* len,
                    uint32_t n_seq) {
  if (!c || store < 0 || store > 1 || (alphabet != TRPA_ALPHA_NT && alphabet != TRPA_ALPHA_AA) || (n_seq && (!chars || !off || !len))) {
    set_error("bad store arguments"); return TRPA_ERR_ARG;
  }
  if (use_device(c)) return TRPA_ERR_CUDA;
  Store& S = c->store[store];
  S.release();
  std::vector<u64> woff(n_seq + 1);
  u64 words = 0, nchars = 0; u32 maxlen = 0;
  const u32 per = alphabet == TRPA_ALPHA_NT ? 32 : 6;
  for (u32 i = 0; i < n_seq; ++i) {
    woff[i] = words;
    words += ((u64)len[i] + per - 1) / per;
    nchars = std::max<u64>(nchars, off[i] + len[i]);
    maxlen = std::max(maxlen, len[i]);
  }
  woff[n_seq] = words;
  if (alphabet == TRPA_ALPHA_NT && words >= 0xffffffffull) {
    // SeqDesc.woff of the packing descriptor is 32 bit; stores beyond 137 Gbases need a wider layout
    set_error("nucleotide store larger than 2^32 words is not supported yet"); return TRPA_ERR_ARG;
  }
  ScopedBuf<uint8_t> d_chars; ScopedBuf<u64> d_off;
  if (d_chars.ensure(nchars + 1) || d_off.ensure(n_seq + 1) || S.woff.ensure(n_seq + 1) || S.len.ensure(n_seq + 1)) return TRPA_ERR_NOMEM;
  CK(cudaMemcpyAsync(d_chars.p, chars, nchars, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_off.p, off, 8ull * n_seq, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(S.woff.p, woff.data(), 8ull * (n_seq + 1), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(S.len.p, len, 4ull * n_seq, cudaMemcpyHostToDevice, c->stream));
  if (alphabet == TRPA_ALPHA_NT) {
    std::vector<SeqDesc> sd(n_seq);
    for (u32 i = 0; i < n_seq; ++i) sd[i] = SeqDesc{(u32)woff[i], len[i], 0, 0};
    ScopedBuf<SeqDesc> d_sd;
    if (d_sd.ensure(n_seq + 1) || S.planes.ensure(words + 2) || S.nplane.ensure(words + 2)) return TRPA_ERR_NOMEM;
    CK(cudaMemsetAsync(S.planes.p, 0, (words + 2) * sizeof(uint2), c->stream));
    CK(cudaMemsetAsync(S.nplane.p, 0, (words + 2) * sizeof(u32), c->stream));
    CK(cudaMemcpyAsync(d_sd.p, sd.data(), sizeof(SeqDesc) * n_seq, cudaMemcpyHostToDevice, c->stream));
    CK(launch_pack_nt(d_chars.p, d_off.p, d_sd.p, n_seq, words, S.planes.p, S.nplane.p, nullptr, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    d_sd.release();
  } else {
    if (S.packed.ensure(words + 2)) return TRPA_ERR_NOMEM;
    CK(cudaMemsetAsync(S.packed.p, 0, (words + 2) * sizeof(u32), c->stream));
    CK(launch_pack_aa(d_chars.p, d_off.p, S.woff.p, S.len.p, n_seq, words, S.packed.p, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int rc = aa_store_mask(c, S, words);
    if (rc) return rc;
  }
  d_chars.release(); d_off.release();
  S.alphabet = alphabet; S.n_seq = n_seq; S.n_words = words; S.max_len = maxlen;
  S.h_len.assign(len, len + n_seq);
  c->batch_ready = false;
  return 0;
}

// ---- packed stores: what trpa_load_store built, out of and back into HBM without the ASCII detour
int trpa_store_info(trpa_ctx* c, int store, int* alphabet, uint32_t* n_seq, uint64_t* n_words) {
  if (!c || store < 0 || store > 1 || !alphabet || !n_seq || !n_words) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  const Store& S = c->store[store];
  if (S.alphabet < 0) { set_error("store not loaded"); return TRPA_ERR_STATE; }
  *alphabet = S.alphabet; *n_seq = S.n_seq; *n_words = S.n_words;
  return 0;
}

int trpa_export_store(trpa_ctx* c, int store, uint64_t* woff, uint32_t* len, void* payload) {
  if (!c || store < 0 || store > 1 || !woff || !len || !payload) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  const Store& S = c->store[store];
  if (S.alphabet < 0) { set_error("store not loaded"); return TRPA_ERR_STATE; }
  CK(cudaMemcpyAsync(woff, S.woff.p, 8ull * (S.n_seq + 1), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(len, S.len.p, 4ull * S.n_seq, cudaMemcpyDeviceToHost, c->stream));
  uint8_t* out = (uint8_t*)payload;
  if (S.alphabet == TRPA_ALPHA_NT) {
    CK(cudaMemcpyAsync(out, S.planes.p, 8ull * S.n_words, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(out + 8ull * S.n_words, S.nplane.p, 4ull * S.n_words, cudaMemcpyDeviceToHost, c->stream));
  } else {
    CK(cudaMemcpyAsync(out, S.packed.p, 4ull * S.n_words, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int trpa_load_store_packed(trpa_ctx* c, int store, int alphabet, const uint64_t* woff, const uint32_t* len,
                           uint32_t n_seq, const void* payload, uint64_t n_words) {
  if (!c || store < 0 || store > 1 || (alphabet != TRPA_ALPHA_NT && alphabet != TRPA_ALPHA_AA) ||
      (n_seq && (!woff || !len)) || (n_words && !payload)) {
    set_error("bad store arguments"); return TRPA_ERR_ARG;
  }
  // the word offsets must describe word-aligned, non-overlapping sequences in order
  const u64 per = alphabet == TRPA_ALPHA_NT ? 32 : 6;
  u64 expect = 0; u32 maxlen = 0;
  for (u32 i = 0; i < n_seq; ++i) {
    if (woff[i] != expect) { set_error("packed store: word offsets are not contiguous"); return TRPA_ERR_ARG; }
    expect += ((u64)len[i] + per - 1) / per;
    maxlen = std::max(maxlen, len[i]);
  }
  if (expect != n_words || (n_seq && woff[n_seq] != n_words)) { set_error("packed store: word count does not match the lengths"); return TRPA_ERR_ARG; }
  if (alphabet == TRPA_ALPHA_NT && n_words >= 0xffffffffull) { set_error("nucleotide store larger than 2^32 words is not supported yet"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  Store& S = c->store[store];
  S.release();
  if (S.woff.ensure(n_seq + 1) || S.len.ensure(n_seq + 1)) return TRPA_ERR_NOMEM;
  const u64 zero = 0;
  CK(cudaMemcpyAsync(S.woff.p, n_seq ? woff : &zero, 8ull * (n_seq + 1), cudaMemcpyHostToDevice, c->stream));
  if (n_seq) CK(cudaMemcpyAsync(S.len.p, len, 4ull * n_seq, cudaMemcpyHostToDevice, c->stream));
  const uint8_t* in = (const uint8_t*)payload;
  if (alphabet == TRPA_ALPHA_NT) {
    if (S.planes.ensure(n_words + 2) || S.nplane.ensure(n_words + 2)) return TRPA_ERR_NOMEM;
    CK(cudaMemsetAsync(S.planes.p + n_words, 0, 2 * sizeof(uint2), c->stream));
    CK(cudaMemsetAsync(S.nplane.p + n_words, 0, 2 * sizeof(u32), c->stream));
    if (n_words) {
      CK(cudaMemcpyAsync(S.planes.p, in, 8ull * n_words, cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpyAsync(S.nplane.p, in + 8ull * n_words, 4ull * n_words, cudaMemcpyHostToDevice, c->stream));
    }
  } else {
    if (S.packed.ensure(n_words + 2)) return TRPA_ERR_NOMEM;
    CK(cudaMemsetAsync(S.packed.p + n_words, 0, 2 * sizeof(u32), c->stream));
    if (n_words) CK(cudaMemcpyAsync(S.packed.p, in, 4ull * n_words, cudaMemcpyHostToDevice, c->stream));
    const int rc = aa_store_mask(c, S, n_words);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(c->stream));
  S.alphabet = alphabet; S.n_seq = n_seq; S.n_words = n_words; S.max_len = maxlen;
  S.h_len.assign(len, len + n_seq);
  c->batch_ready = false;
  return 0;
}

// ------------------------------------------------------------------------------ the hot path
int trpa_batch_upload(trpa_ctx* c, const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                      uint32_t n_cands) {
  if (!c || (n_segs && !segs) || (n_cands && !cands)) { set_error("bad batch arguments"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  if (c->n_nodes == 0) { set_error("taxonomy not loaded"); return TRPA_ERR_STATE; }
  const Store& Q = c->store[TRPA_STORE_QUERY];
  const Store& R = c->store[TRPA_STORE_REF];
  if (Q.alphabet < 0 || R.alphabet < 0 || Q.alphabet != R.alphabet) { set_error("query/reference stores not loaded or of different alphabets"); return TRPA_ERR_STATE; }
  if ((u64)n_segs + n_cands >= 0xfffffff0ull) { set_error("batch too large"); return TRPA_ERR_ARG; }
  const bool protein = Q.alphabet == TRPA_ALPHA_AA;
  c->batch_ready = false;   // a failed upload must not leave the previous batch's plan runnable over new tables
  // validate the segment table (the reference throws SequenceNotFound / TaxonNotFound at parse time).
  // Contract (include/taxator_rpa_b200.h): record sets in ascending order and disjoint (unused candidates between
  // two sets are fine) -- the per-candidate work arrays and the round queues are laid out by cand_begin.
  {
    u64 expect = 0;
    for (u32 s = 0; s < n_segs; ++s) {
      if (segs[s].cand_begin < expect || (u64)segs[s].cand_begin + segs[s].cand_count > n_cands) {
        set_error("segment table: candidate ranges must be ascending, disjoint and inside the candidate table");
        return TRPA_ERR_ARG;
      }
      expect = (u64)segs[s].cand_begin + segs[s].cand_count;
      if (segs[s].cand_count && segs[s].query_seq >= Q.n_seq) { set_error("segment query ordinal out of range"); return TRPA_ERR_ARG; }
    }
  }
  // The raw tables go to the device straight from the caller's buffers (a DMA when they are pinned)
  // while the host makes its one validation pass over them; SortFilter then runs on the device.
  if (c->d_segs.ensure(n_segs + 1) || c->d_cands.ensure(n_cands + 1) || c->d_cands_raw.ensure(n_cands + 1)) return TRPA_ERR_NOMEM;
  // from here on the stream reads the caller's buffers: every return synchronises first
  struct SyncGuard { cudaStream_t st; ~SyncGuard() { cudaStreamSynchronize(st); } } sync_guard{c->stream};
  CK(cudaMemcpyAsync(c->d_segs.p, segs, sizeof(trpa_segment) * n_segs, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_cands_raw.p, cands, sizeof(trpa_candidate) * n_cands, cudaMemcpyHostToDevice, c->stream));
  if (n_segs) {
    const u32 blocks = std::min<u32>((n_segs + 7) / 8, (u32)c->num_sms * 8u);
    sort_cands_kernel<<<blocks, 256, 0, c->stream>>>(c->d_segs.p, n_segs, c->d_cands_raw.p, c->d_cands.p);
  }
  std::vector<u64> bound(n_segs);
  std::vector<u32> maxspan(n_segs, 0);
  {
    // A segment whose best score is negative (or NaN) realigns nothing in pass 0 and trips
    // assert(!qgroup.empty()) in the reference (hh:563); reject it instead of guessing.
    const float factor = 1. - c->toppercent;
    const u32 nseq = R.n_seq, nnodes = c->n_nodes;
    const u32* qlen = Q.h_len.data();
    // first bad record in table order wins (one slot per worker block, reduced afterwards)
    struct Bad { size_t seg; int code; };
    std::mutex bad_mutex;
    Bad first_bad{~(size_t)0, 0};
    parallel_blocks(n_segs, (size_t)n_segs + n_cands, [&](size_t b, size_t e) {
      int code = 0;
      size_t at = 0;
      for (size_t s = b; s < e && !code; ++s) {
        const trpa_segment sg = segs[s];
        const trpa_candidate* rc = cands + sg.cand_begin;
        u32 ms = 0, qmin = 0xffffffffu;
        float best = sg.cand_count ? rc[0].score : 0.f;
        for (u32 k = 0; k < sg.cand_count; ++k) {
          const trpa_candidate& x = rc[k];
          if (x.ref_seq >= nseq) code = 1;
          else if (x.node >= nnodes) code = 2;
          else if (x.qstart > x.qstop || x.qstart == 0) code = 3;
          else if (x.rstart == 0 || x.rstop == 0) code = 4;
          if (code) break;
          best = std::max(best, x.score);
          qmin = std::min(qmin, x.qstart);
          const u64 span = (x.rstart <= x.rstop ? (u64)x.rstop - x.rstart : (u64)x.rstart - x.rstop) + 1;
          ms = (u32)std::min<u64>(0xffffffffull, std::max<u64>(ms, span));
        }
        if (!code && sg.cand_count >= 2 && !(best >= factor * best)) code = 5;
        // the reference fetches the query range of every n >= 2 record set (hh:415) and its in-memory query
        // store throws SequenceRangeError when the range starts beyond the sequence (sequencestorage.hh:118)
        if (!code && sg.cand_count >= 2 && qmin > qlen[sg.query_seq]) code = 6;
        if (code) { at = s; break; }
        maxspan[s] = ms;
        bound[s] = segment_arena_bound(sg, cands, protein);
      }
      if (code) {
        std::lock_guard<std::mutex> lock(bad_mutex);
        if (at < first_bad.seg) first_bad = Bad{at, code};
      }
    });
    static const char* const kMsg[7] = {nullptr, "candidate reference ordinal out of range", "candidate taxon node out of range",
      "candidate query range invalid (qstart must be >= 1 and <= qstop)", "candidate reference coordinates are 1-based",
      "segment with negative best alignment score: undefined in the reference (hh:563)",
      "segment query range starts beyond the query sequence (SequenceRangeError, sequencestorage.hh:118)"};
    if (first_bad.code) {
      cudaGetLastError();
      set_error(std::string(kMsg[first_bad.code]) + " (segment " + std::to_string(first_bad.seg) + ")");
      return TRPA_ERR_ARG;
    }
  }
  c->n_segs = n_segs; c->n_cands = n_cands;

  // arena + chunk plan
  u64 total_bound = 0, max_bound = 0; u32 max_len = 0;
  for (u32 s = 0; s < n_segs; ++s) {
    total_bound += bound[s];
    max_bound = std::max(max_bound, bound[s]);
    max_len = std::max(max_len, maxspan[s]);
  }
  const u64 unit_bytes = protein ? 1 : 20;   // planes 8 + column codes 8 + N plane 4 per 32 bases
  const u64 have_units = protein ? (u64)c->arena_aa.cap : std::min<u64>(std::min<u64>(c->arena_planes.cap, c->arena_codes.cap), c->arena_n.cap);
  u64 units_cap;
  if (!c->arena_bytes && have_units >= total_bound + 16) {
    units_cap = have_units - 16;   // the arena of an earlier batch is large enough: one chunk, no query
  } else {
    u64 arena_bytes = c->arena_bytes;
    if (!arena_bytes) {
      size_t free_b = 0, total_b = 0;
      CK(cudaMemGetInfo(&free_b, &total_b));
      const u64 per_cand = 36 + 4 + 4 + 4 + 4 + 1 + 4 + 4 + 4 + 8 + 16 + 32 + 20 + 16;  // rough per-candidate work bytes
      const u64 fixed = (u64)n_cands * per_cand + (u64)n_segs * (sizeof(SegState) + sizeof(trpa_result) + 96);
      const u64 avail = (free_b > fixed ? free_b - fixed : 0) + have_units * unit_bytes;
      arena_bytes = avail / 2;
    }
    units_cap = std::min<u64>(arena_bytes / unit_bytes, 0xfffffff0ull);
  }
  // extensions can add at most the query range; sequences are clipped to the store anyway
  max_len = std::min<u64>((u64)max_len + Q.max_len, std::max(R.max_len, Q.max_len));
  c->max_stage_len = max_len;
  c->avg_stage_len = (u32)std::min<u64>(0xffffffffull, (protein ? total_bound : total_bound * 32ull) / std::max<u64>(1, (u64)n_cands + n_segs));
  // chunks of segments: every pipe works on one chunk at a time inside its own arena region
  const int K = (c->n_pipes > 1 && n_segs >= 8192u) ? c->n_pipes : 1;
  u64 cap_pipe = units_cap / K;
  cap_pipe = std::min(cap_pipe, (total_bound + K - 1) / K + max_bound);
  cap_pipe = std::max<u64>(cap_pipe, 1);
  if (max_bound > cap_pipe) { set_error("staging arena too small for the largest segment; raise trpa_set_arena_bytes"); return TRPA_ERR_NOMEM; }
  const u64 n_chunks_min = std::max<u64>(K, (total_bound + cap_pipe - 1) / cap_pipe);
  const u64 target = std::min<u64>(cap_pipe, (total_bound + n_chunks_min - 1) / n_chunks_min + 1);
  c->chunk_begin.clear();
  c->chunk_qoff.clear();
  c->chunk_begin.push_back(0);
  c->chunk_qoff.push_back(0);
  u64 acc = 0, qacc = 0;
  for (u32 s = 0; s < n_segs; ++s) {
    if (acc && (acc + bound[s] > cap_pipe || acc >= target)) { c->chunk_begin.push_back(s); c->chunk_qoff.push_back(qacc); acc = 0; }
    acc += bound[s];
    qacc += (u64)segs[s].cand_count + 1u;
  }
  c->chunk_begin.push_back(n_segs);
  c->chunk_qoff.push_back(qacc);
  c->run_pipes = K;
  c->arena_units = cap_pipe;      // per pipe
  units_cap = cap_pipe * K;

  const size_t nslots = (size_t)n_cands + n_segs;
  if (c->d_results.ensure(n_segs + 1) ||
      c->d_state.ensure(n_segs + 1) || c->d_qd.ensure(n_cands + 1) || c->d_qsim.ensure(n_cands + 1) ||
      c->d_bf_d.ensure(nslots + 1) || c->d_bf_node.ensure(nslots + 1) || c->d_cflags.ensure(n_cands + 1) ||
      c->d_og_i.ensure(n_cands + 1) || c->d_tag.ensure(n_cands + 1) || c->d_og_d.ensure(n_cands + 1) || c->d_res.ensure(2 * nslots + 2) ||
      c->d_descs.ensure(nslots + 1) || c->d_pairs.ensure(nslots + 1) || c->d_pairs_sorted.ensure(nslots + 1) ||
      c->d_stage.ensure(nslots + 1))
    return TRPA_ERR_NOMEM;
  if (protein) { if (c->arena_aa.ensure(units_cap + 16)) return TRPA_ERR_NOMEM; }
  else { if (c->arena_planes.ensure(units_cap + 2) || c->arena_codes.ensure(units_cap + 2) || c->arena_n.ensure(units_cap + 2)) return TRPA_ERR_NOMEM; }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->batch_ready = true;
  return 0;
}

namespace trpa {

enum PipeState { PS_IDLE = 0, PS_WAIT_DECIDE, PS_WAIT_PLAN, PS_DONE };

// decide kernel of the pipe's chunk + read-back of the round counters (asynchronous)
static int enqueue_decide(trpa_ctx* c, Pipe& P) {
  const int ev = begin_event(P, EV_DECIDE);
  {
    // Segments per warp: as few as still fit the chunk into ONE wave of warps (24 resident warps per SM at the
    // kernel's register count).  Measured (gpurun_out/r2_54): C1, 1 k segments: 32 per warp 1.04 ms of decide per step,
    // 1 per warp 0.50 ms; C3 / C2 (50 k / 100 k segments) want all 32 lanes (8 per warp: +8 % / +12 % decide time).
    static const int forced = [] { const char* e = getenv("TRPA_DECIDE_PER_WARP"); const int v = e ? atoi(e) : 0;
                                   return (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ? v : 0; }();
    const u32 n = P.se - P.sb, wave = (u32)c->num_sms * 24u;
    u32 per_warp = 1;
    while (per_warp < 32u && (n + per_warp - 1) / per_warp > wave) per_warp *= 2u;
    if (forced) per_warp = (u32)forced;
    const u32 warps = (n + per_warp - 1) / per_warp;
    decide_kernel<<<(warps + 1) / 2, 64, 0, P.stream>>>(P.B, P.sb, P.se, per_warp);
  }
  CK(cudaGetLastError());
  end_event(P, ev);
  c->prof.launches_decide++;
  CK(cudaMemcpyAsync(P.h_counters, P.d_counters.p, sizeof(u32) * kNumCounters, cudaMemcpyDeviceToHost, P.stream));
  P.state = PS_WAIT_DECIDE;
  return 0;
}

// hand the next chunk of the batch to pipe P (or retire it)
static int start_chunk(trpa_ctx* c, Pipe& P, int pipe_index, size_t& next_chunk, const Batch& base) {
  while (next_chunk + 1 < c->chunk_begin.size() && c->chunk_begin[next_chunk] == c->chunk_begin[next_chunk + 1]) ++next_chunk;
  if (next_chunk + 1 >= c->chunk_begin.size()) { P.state = PS_DONE; return 0; }
  const size_t ch = next_chunk++;
  P.sb = c->chunk_begin[ch]; P.se = c->chunk_begin[ch + 1];
  P.B = base;
  P.B.pairs = c->d_pairs.p + c->chunk_qoff[ch];
  P.B.stage = c->d_stage.p + c->chunk_qoff[ch];
  P.B.counters = P.d_counters.p;
  P.B.arena_base = (u32)(c->arena_units * (u64)pipe_index);
  P.B.arena_capacity = (u32)c->arena_units;
  P.B.spec_k = 0;
  if (c->trace_on && c->d_trace.p) {   // the chunk appends behind what earlier chunks of this run recorded
    P.B.trace = c->d_trace.p + c->trace_used;
    P.B.trace_capacity = (u32)std::min<u64>(0xffffffffull, c->d_trace.cap - c->trace_used);
  }
  P.pairs = P.B.pairs;
  P.sorted = c->d_pairs_sorted.p + c->chunk_qoff[ch];
  CK(cudaMemsetAsync(P.d_counters.p, 0, sizeof(u32) * kNumCounters, P.stream));
  init_state_kernel<<<(P.se - P.sb + 255) / 256, 256, 0, P.stream>>>(c->d_state.p + P.sb, P.se - P.sb);
  CK(cudaGetLastError());
  return enqueue_decide(c, P);
}

// advance pipe P by one host step: wait for what it has in flight, then enqueue its next phase
static int step_pipe(trpa_ctx* c, Pipe& P, int pipe_index, size_t& next_chunk, const Batch& base) {
  const Store& Q = c->store[TRPA_STORE_QUERY];
  const Store& R = c->store[TRPA_STORE_REF];
  const bool protein = Q.alphabet == TRPA_ALPHA_AA;
  CK(cudaStreamSynchronize(P.stream));
  harvest_events(c, P);
  if (P.state == PS_WAIT_DECIDE) {
    const u32 n_pairs = P.h_counters[CN_PAIRS], n_stage = P.h_counters[CN_STAGE], n_active = P.h_counters[CN_ACTIVE];
    // look-ahead budget of the NEXT decide round: aim for about half a pair per resident lane and round
    // (a banded pair needs only a few lanes; more pairs per round = cheaper shapes, fewer dependent rounds)
    if (c->lookahead >= 0) P.B.spec_k = (u32)c->lookahead;
    else {
      const u32 chunk_segs = P.se - P.sb;
      // long pairs fill the GPU with fewer of them and a wasted look-ahead alignment costs more: beyond 12 kb average
      // the budget shrinks in proportion (200 k x 10-50 kb reads: 750 k pairs per round cost 3.5 %)
      // (protein: every alignment is a full matrix, an unused look-ahead result costs as much as a used one: 3 pairs
      // per segment, measured on C3 with 150 k against 275 k pairs per round: +0.9 %, gpurun_out/r2_35)
      u64 cap_scaled = std::min<u64>(750000u, std::max<u64>(150000u, protein ? (u64)chunk_segs * 3u : (u64)chunk_segs * 11u / 2u));
      if (c->avg_stage_len > 12000u) cap_scaled = std::max<u64>(150000u, cap_scaled * 12000u / c->avg_stage_len);
      const u32 cap_auto = (u32)cap_scaled;
      const u32 cap = c->la_cap ? c->la_cap / (u32)c->run_pipes : cap_auto;
      P.B.spec_k = n_active ? std::min<u32>(c->la_max, cap / n_active > 0 ? cap / n_active - 1 : 0) : 0;
    }
    if (P.h_counters[CN_OVERFLOW]) { set_error("internal: staging arena overflow"); return TRPA_ERR_STATE; }
    if (n_pairs == 0) {
      if (n_active) { set_error("internal: segments active without pending alignments"); return TRPA_ERR_STATE; }
      // algorithmic staging traffic of the chunk: packed store bits read + staged bits written
      // (nt: 3 planes x 4 B per 32-base word read, those + 8 B of column codes written; aa: 5 bit read +
      // 1 byte written per residue)
      const u64 units = P.h_counters[CN_ARENA];
      c->prof.bytes_stage += protein ? (units * 13) / 8 : units * 32;
      if (!protein) { const int rc = harvest_band_stats(c, P); if (rc) return rc; }
      if (P.B.trace) {
        const u64 got = P.h_counters[CN_TRACE];
        if (got > P.B.trace_capacity) c->trace_overflow = true;
        c->trace_used += std::min<u64>(got, P.B.trace_capacity);
      }
      return start_chunk(c, P, pipe_index, next_chunk, base);
    }
    P.n_pairs = n_pairs;
    c->prof.rounds++;
    c->prof.pairs += n_pairs;
    // --- stage
    int ev = begin_event(P, EV_STAGE);
    if (protein)
      CK(launch_stage_aa(P.B.stage, n_stage, Q.packed.p, Q.woff.p, R.packed.p, R.woff.p, c->d_descs.p, c->arena_aa.p, P.stream));
    else
      CK(launch_stage_nt(P.B.stage, n_stage, Q.planes.p, Q.nplane.p, Q.woff.p, R.planes.p, R.nplane.p, R.woff.p,
                         c->d_descs.p, c->arena_planes.p, c->arena_n.p, c->arena_codes.p, P.stream));
    end_event(P, ev);
    if (n_stage) c->prof.launches_stage++;
    if (protein) {
      int2* scr = nullptr; u32 stride = 0;
      if (c->max_stage_len > 512) {
        stride = c->max_stage_len + 2;
        if (P.scratch_aa.ensure((size_t)n_pairs * stride)) return TRPA_ERR_NOMEM;
        scr = P.scratch_aa.p;
      }
      ev = begin_event(P, EV_PROTEIN);
      CK(launch_protein(P.pairs, n_pairs, c->d_descs.p, c->arena_aa.p, (int2*)c->d_res.p, scr, stride,
                        c->max_stage_len, c->store[0].aa_mask | c->store[1].aa_mask, P.stream));
      end_event(P, ev);
      c->prof.launches_protein++;
      CK(cudaMemsetAsync(P.d_counters.p + CN_PAIRS, 0, sizeof(u32) * 3, P.stream));
      return enqueue_decide(c, P);
    }
    // --- plan: threshold + shape of every pair, counting sort by shape, histogram to the host
    ev = begin_event(P, EV_OTHER);
    u32* h_hist = P.h_counters + kNumCounters;
    const int rc = bucket_pairs3(c, P, P.pairs, n_pairs, c->d_descs.p, c->arena_planes.p, c->arena_n.p, P.sorted, h_hist);
    if (rc) return rc;
    end_event(P, ev);
    c->prof.launches_other += 3;
    P.state = PS_WAIT_PLAN;
    return 0;
  }
  if (P.state == PS_WAIT_PLAN) {
    // --- align: one persistent launch per non-empty shape
    const u32* h_hist = P.h_counters + kNumCounters;
    const int ev = begin_event(P, EV_MYERS);
    const int rc = launch_myers_shapes3(c, P, h_hist, P.sorted, c->d_descs.p, c->arena_planes.p, c->arena_n.p, c->arena_codes.p,
                                        c->d_res.p, c->max_stage_len);
    if (rc) return rc;
    end_event(P, ev);
    // reset the per-round counters, keep the arena cursor
    CK(cudaMemsetAsync(P.d_counters.p + CN_PAIRS, 0, sizeof(u32) * 3, P.stream));
    return enqueue_decide(c, P);
  }
  return 0;
}

}  // namespace trpa

int trpa_batch_run(trpa_ctx* c) {
  if (!c) { set_error("null ctx"); return TRPA_ERR_ARG; }
  if (!c->batch_ready) { set_error("no batch uploaded"); return TRPA_ERR_STATE; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  const Store& Q = c->store[TRPA_STORE_QUERY];
  const Store& R = c->store[TRPA_STORE_REF];
  const bool protein = Q.alphabet == TRPA_ALPHA_AA;
  const u32 n_segs = c->n_segs, n_cands = c->n_cands;
  if (n_segs == 0) return 0;

  Batch B;
  B.segs = c->d_segs.p; B.cands = c->d_cands.p; B.n_segs = n_segs; B.n_cands = n_cands;
  B.q_len = Q.len.p; B.r_len = R.len.p;
  B.tax = dev_tax(c);
  B.protein = protein ? 1 : 0;
  B.exclude_factor = c->exclude_factor;
  B.reeval_bandwidth_factor = 1. - c->toppercent;  // taxonpredictionmodelsequence.hh:334
  B.st = c->d_state.p; B.qd = c->d_qd.p; B.qsim = c->d_qsim.p; B.cflags = c->d_cflags.p;
  B.tag = c->d_tag.p; B.spec_k = 0;
  B.og_i = c->d_og_i.p; B.og_d = c->d_og_d.p; B.bf_d = c->d_bf_d.p; B.bf_node = c->d_bf_node.p;
  B.res_nt = c->d_res.p; B.res_aa = c->d_res.p;
  B.descs = c->d_descs.p; B.arena_capacity = (u32)c->arena_units; B.arena_base = 0;
  B.pairs = c->d_pairs.p; B.stage = c->d_stage.p; B.counters = nullptr; B.results = c->d_results.p;
  B.trace = nullptr; B.trace_capacity = 0;
  c->trace_used = 0; c->trace_overflow = false;
  if (c->trace_on) {
    // every record can be realigned against the query and against each anchor; 8 entries per record + 64 per
    // segment covers the inputs the log is used on (an overflow is reported, never silent)
    if (c->d_trace.ensure((size_t)n_cands * 8 + (size_t)n_segs * 64 + 1024)) return TRPA_ERR_NOMEM;
  }
  if (protein) CK(ensure_blosum_constant(c->device));

  // the chunks of the batch are worked on by run_pipes pipes at the same time; the host advances
  // them round-robin, so while it waits for one pipe the others have work queued on the GPU
  const int K = c->trace_on ? 1 : c->run_pipes;   // the trace is appended chunk after chunk
  for (int k = 0; k < K; ++k) { const int rc = c->pipe[k].init(nullptr); if (rc) return rc; }
  if (K > 1) {   // the other pipes start after whatever is queued on the context's stream
    Pipe& P0 = c->pipe[0];
    if (!P0.fork_ev) {
      CK(cudaEventCreateWithFlags(&P0.fork_ev, cudaEventDisableTiming));
      for (int i = 0; i < Pipe::kAux; ++i) {
        CK(cudaStreamCreateWithFlags(&P0.aux[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&P0.aux_done[i], cudaEventDisableTiming));
      }
    }
    CK(cudaEventRecord(P0.fork_ev, P0.stream));
    for (int k = 1; k < K; ++k) CK(cudaStreamWaitEvent(c->pipe[k].stream, P0.fork_ev, 0));
  }
  size_t next_chunk = 0;
  for (int k = 0; k < K; ++k) { const int rc = start_chunk(c, c->pipe[k], k, next_chunk, B); if (rc) return rc; }
  for (;;) {
    bool any = false;
    for (int k = 0; k < K; ++k) {
      Pipe& P = c->pipe[k];
      if (P.state == PS_DONE) continue;
      any = true;
      const int rc = step_pipe(c, P, k, next_chunk, B);
      if (rc) {
        for (int q = 0; q < K; ++q) cudaStreamSynchronize(c->pipe[q].stream);
        return rc;
      }
    }
    if (!any) break;
  }
  return 0;
}

int trpa_batch_download(trpa_ctx* c, trpa_result* out) {
  if (!c || !out) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (!c->batch_ready) { set_error("no batch uploaded"); return TRPA_ERR_STATE; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  CK(cudaMemcpyAsync(out, c->d_results.p, sizeof(trpa_result) * c->n_segs, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- page-locked host buffers
static std::mutex g_host_mutex;
static std::vector<void*> g_host_plain;   // allocations that fell back to ordinary memory (no CUDA device)

void* trpa_host_alloc(uint64_t bytes) {
  if (bytes == 0) bytes = 1;
  void* p = nullptr;
  // allocate under a device this process already uses: a thread that never touched CUDA (the ingest thread of the
  // CLI) would otherwise create a context on device 0, which the caller may not own
  static const bool no_pinned = getenv("TRPA_NO_PINNED") != nullptr;   // A/B switch
  const int dev = no_pinned ? -1 : g_first_device.load();
  if (dev >= 0 && cudaSetDevice(dev) == cudaSuccess && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) return p;
  cudaGetLastError();
  p = malloc(bytes);
  if (p) { std::lock_guard<std::mutex> l(g_host_mutex); g_host_plain.push_back(p); }
  return p;
}
void trpa_host_free(void* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> l(g_host_mutex);
    auto it = std::find(g_host_plain.begin(), g_host_plain.end(), p);
    if (it != g_host_plain.end()) { g_host_plain.erase(it); free(p); return; }
  }
  cudaFreeHost(p);
}

int trpa_set_trace(trpa_ctx* c, int on) {
  if (!c) { set_error("null ctx"); return TRPA_ERR_ARG; }
  c->trace_on = on ? 1 : 0;
  return 0;
}

int trpa_batch_trace(trpa_ctx* c, trpa_trace_entry* out, uint64_t cap, uint64_t* n) {
  if (!c || !n || (cap && !out)) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (!c->batch_ready) { set_error("no batch uploaded"); return TRPA_ERR_STATE; }
  if (c->trace_overflow) { set_error("alignment trace overflow: run the log on smaller batches"); return TRPA_ERR_NOMEM; }
  *n = c->trace_used;
  if (cap == 0 || c->trace_used == 0) return 0;
  if (cap < c->trace_used) { set_error("trace buffer too small"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  CK(cudaMemcpyAsync(out, c->d_trace.p, sizeof(trpa_trace_entry) * c->trace_used, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // device threads append concurrently: order by segment, keeping each segment's own (program) order
  std::stable_sort(out, out + c->trace_used, [](const trpa_trace_entry& x, const trpa_trace_entry& y) { return x.seg < y.seg; });
  return 0;
}

int trpa_batch_results_dev(trpa_ctx* c, void** dev_ptr, uint32_t* n_segs) {
  if (!c || !dev_ptr || !n_segs) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (!c->batch_ready) { set_error("no batch uploaded"); return TRPA_ERR_STATE; }
  *dev_ptr = c->d_results.p; *n_segs = c->n_segs;
  return 0;
}

int trpa_predict_batch(trpa_ctx* c, const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands,
                       uint32_t n_cands, trpa_result* out) {
  static const bool timing = getenv("TRPA_DEBUG_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = trpa_batch_upload(c, segs, n_segs, cands, n_cands);
  if (rc) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  rc = trpa_batch_run(c);
  if (rc) return rc;
  const auto t2 = std::chrono::steady_clock::now();
  rc = trpa_batch_download(c, out);
  if (timing) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[trpa] predict_batch: upload %.3f ms, run %.3f ms, download %.3f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3));
  }
  return rc;
}

// ----------------------------------------------------------------------- lower-level entry points
static int upload_table(trpa_ctx* c, const char* chars, const uint64_t* off, const uint32_t* len, uint32_t n_seq,
                        ScopedBuf<uint8_t>& d_chars, ScopedBuf<u64>& d_off, u64* nchars_out) {
  u64 nchars = 0;
  for (u32 i = 0; i < n_seq; ++i) nchars = std::max<u64>(nchars, off[i] + len[i]);
  if (d_chars.ensure(nchars + 1) || d_off.ensure(n_seq + 1)) return TRPA_ERR_NOMEM;
  CK(cudaMemcpyAsync(d_chars.p, chars, nchars, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_off.p, off, 8ull * n_seq, cudaMemcpyHostToDevice, c->stream));
  *nchars_out = nchars;
  return 0;
}

int trpa_edit_distance_batch(trpa_ctx* c, const char* chars, const uint64_t* off, const uint32_t* len, uint32_t n_seq,
                             const uint32_t* pair_a, const uint32_t* pair_b, uint32_t n_pairs, int32_t* out_dist,
                             int repeat, double* kernel_ms) {
  if (!c || !chars || !off || !len || !pair_a || !pair_b || !out_dist) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  if (n_pairs == 0) return 0;
  for (u32 k = 0; k < n_pairs; ++k)
    if (pair_a[k] >= n_seq || pair_b[k] >= n_seq) { set_error("pair index out of range"); return TRPA_ERR_ARG; }
  ScopedBuf<uint8_t> d_chars; ScopedBuf<u64> d_off; u64 nchars = 0;
  int rc = upload_table(c, chars, off, len, n_seq, d_chars, d_off, &nchars);
  if (rc) return rc;
  std::vector<SeqDesc> sd(n_seq);
  u64 words = 0; u32 max_len = 0;
  for (u32 i = 0; i < n_seq; ++i) { sd[i] = SeqDesc{(u32)words, len[i], 0, 0}; words += ((u64)len[i] + 31) / 32; max_len = std::max(max_len, len[i]); }
  if (words >= 0xffffffffull) { set_error("table too large"); return TRPA_ERR_ARG; }
  ScopedBuf<SeqDesc> d_sd; ScopedBuf<uint2> planes, codes; ScopedBuf<u32> nplane; ScopedBuf<PairDesc> d_pairs, d_sorted; ScopedBuf<int32_t> d_out;
  ScopedBuf<u32> d_cnt;
  if (d_sd.ensure(n_seq + 1) || planes.ensure(words + 2) || codes.ensure(words + 2) || nplane.ensure(words + 2) || d_pairs.ensure(n_pairs) ||
      d_sorted.ensure(n_pairs) || d_out.ensure(n_pairs) || d_cnt.ensure(kNumCounters))
    return TRPA_ERR_NOMEM;
  CK(cudaMemcpyAsync(d_sd.p, sd.data(), sizeof(SeqDesc) * n_seq, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(planes.p, 0, (words + 2) * sizeof(uint2), c->stream));
  CK(cudaMemsetAsync(nplane.p, 0, (words + 2) * sizeof(u32), c->stream));
  // flags live inside the descriptor table (3rd u32 of each 16-byte entry): pass a strided view
  ScopedBuf<u32> d_flags;
  if (d_flags.ensure(n_seq + 1)) return TRPA_ERR_NOMEM;
  CK(cudaMemsetAsync(d_flags.p, 0, sizeof(u32) * (n_seq + 1), c->stream));
  CK(launch_pack_nt(d_chars.p, d_off.p, d_sd.p, n_seq, words, planes.p, nplane.p, d_flags.p, c->stream));
  CK(launch_codes_from_planes(planes.p, codes.p, words + 2, c->stream));
  std::vector<u32> h_flags(n_seq);
  CK(cudaMemcpyAsync(h_flags.data(), d_flags.p, sizeof(u32) * n_seq, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (u32 i = 0; i < n_seq; ++i) sd[i].flags = h_flags[i];
  CK(cudaMemcpyAsync(d_sd.p, sd.data(), sizeof(SeqDesc) * n_seq, cudaMemcpyHostToDevice, c->stream));
  std::vector<PairDesc> hp(n_pairs);
  for (u32 k = 0; k < n_pairs; ++k) hp[k] = PairDesc{pair_a[k], pair_b[k], k, 0, 0};
  CK(cudaMemcpyAsync(d_pairs.p, hp.data(), sizeof(PairDesc) * n_pairs, cudaMemcpyHostToDevice, c->stream));
  std::vector<u32> h_hist(kNumShapes, 0);
  Pipe& P = c->pipe[0];
  rc = bucket_pairs3(c, P, d_pairs.p, n_pairs, d_sd.p, planes.p, nplane.p, d_sorted.p, P.h_counters + kNumCounters);
  if (rc) return rc;
  CK(cudaStreamSynchronize(P.stream));
  memcpy(h_hist.data(), P.h_counters + kNumCounters, sizeof(u32) * kNumShapes);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  if (repeat < 1) repeat = 1;
  // one untimed pass when timing is requested
  if (kernel_ms && repeat > 1) {
    rc = launch_myers_shapes3(c, P, h_hist.data(), d_sorted.p, d_sd.p, planes.p, nplane.p, codes.p, d_out.p, max_len);
    if (rc) return rc;
  }
  CK(cudaEventRecord(e0, c->stream));
  for (int r = 0; r < repeat; ++r) {
    rc = launch_myers_shapes3(c, P, h_hist.data(), d_sorted.p, d_sd.p, planes.p, nplane.p, codes.p, d_out.p, max_len);
    if (rc) return rc;
  }
  CK(cudaEventRecord(e1, c->stream));
  CK(cudaMemcpyAsync(out_dist, d_out.p, sizeof(int32_t) * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (kernel_ms) *kernel_ms = ms / repeat;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  rc = harvest_band_stats(c, P); if (rc) return rc;
  d_chars.release(); d_off.release(); d_sd.release(); planes.release(); codes.release(); nplane.release(); d_pairs.release();
  d_sorted.release(); d_out.release(); d_cnt.release(); d_flags.release();
  return 0;
}

int trpa_protein_align_batch(trpa_ctx* c, const char* chars, const uint64_t* off, const uint32_t* len, uint32_t n_seq,
                             const uint32_t* pair_a, const uint32_t* pair_b, uint32_t n_pairs, int32_t* out3,
                             int repeat, double* kernel_ms) {
  if (!c || !chars || !off || !len || !pair_a || !pair_b || !out3) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  if (n_pairs == 0) return 0;
  for (u32 k = 0; k < n_pairs; ++k)
    if (pair_a[k] >= n_seq || pair_b[k] >= n_seq) { set_error("pair index out of range"); return TRPA_ERR_ARG; }
  ScopedBuf<uint8_t> d_chars, d_codes; ScopedBuf<u64> d_off; u64 nchars = 0;
  int rc = upload_table(c, chars, off, len, n_seq, d_chars, d_off, &nchars);
  if (rc) return rc;
  if (nchars >= 0xffffffffull) { set_error("table too large"); return TRPA_ERR_ARG; }
  std::vector<SeqDesc> sd(n_seq);
  u32 max_len = 0;
  for (u32 i = 0; i < n_seq; ++i) { sd[i] = SeqDesc{(u32)off[i], len[i], 0, 0}; max_len = std::max(max_len, len[i]); }
  ScopedBuf<SeqDesc> d_sd; ScopedBuf<PairDesc> d_pairs; ScopedBuf<int2> d_out;
  if (d_codes.ensure(nchars + 1) || d_sd.ensure(n_seq + 1) || d_pairs.ensure(n_pairs) || d_out.ensure(n_pairs)) return TRPA_ERR_NOMEM;
  CK(ensure_blosum_constant(c->device));
  CK(launch_aa_codes(d_chars.p, d_codes.p, nchars, c->stream));
  CK(cudaMemcpyAsync(d_sd.p, sd.data(), sizeof(SeqDesc) * n_seq, cudaMemcpyHostToDevice, c->stream));
  CK(launch_selfscore(d_sd.p, n_seq, d_codes.p, c->stream));
  std::vector<PairDesc> hp(n_pairs);
  for (u32 k = 0; k < n_pairs; ++k) hp[k] = PairDesc{pair_a[k], pair_b[k], k, 0, 0};
  CK(cudaMemcpyAsync(d_pairs.p, hp.data(), sizeof(PairDesc) * n_pairs, cudaMemcpyHostToDevice, c->stream));
  int2* scr = nullptr; u32 stride = 0;
  if (max_len > 512) {
    stride = max_len + 2;
    if (c->pipe[0].scratch_aa.ensure((size_t)n_pairs * stride)) return TRPA_ERR_NOMEM;
    scr = c->pipe[0].scratch_aa.p;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  if (repeat < 1) repeat = 1;
  if (kernel_ms && repeat > 1) CK(launch_protein(d_pairs.p, n_pairs, d_sd.p, d_codes.p, d_out.p, scr, stride, max_len, 0u, c->stream));
  CK(cudaEventRecord(e0, c->stream));
  for (int r = 0; r < repeat; ++r) CK(launch_protein(d_pairs.p, n_pairs, d_sd.p, d_codes.p, d_out.p, scr, stride, max_len, 0u, c->stream));
  CK(cudaEventRecord(e1, c->stream));
  std::vector<int2> ho(n_pairs);
  CK(cudaMemcpyAsync(ho.data(), d_out.p, sizeof(int2) * n_pairs, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(sd.data(), d_sd.p, sizeof(SeqDesc) * n_seq, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (kernel_ms) *kernel_ms = ms / repeat;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  for (u32 k = 0; k < n_pairs; ++k) {
    out3[3 * k + 0] = ho[k].x;
    out3[3 * k + 1] = (int)sd[pair_a[k]].pad + (int)sd[pair_b[k]].pad;
    out3[3 * k + 2] = (int)len[pair_a[k]] + (int)len[pair_b[k]] - ho[k].y;
  }
  d_chars.release(); d_codes.release(); d_off.release(); d_sd.release(); d_pairs.release(); d_out.release();
  return 0;
}

int trpa_fetch_segments(trpa_ctx* c, const uint32_t* ref_seq, const uint32_t* start, const uint32_t* stop,
                        const uint32_t* left_ext, const uint32_t* right_ext, uint32_t n, uint8_t* out_codes,
                        uint64_t out_capacity, uint64_t* out_off, uint32_t* out_len) {
  if (!c || !ref_seq || !start || !stop || !left_ext || !right_ext || !out_codes || !out_off || !out_len) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  const Store& R = c->store[TRPA_STORE_REF];
  if (R.alphabet < 0) { set_error("reference store not loaded"); return TRPA_ERR_STATE; }
  const bool protein = R.alphabet == TRPA_ALPHA_AA;
  std::vector<u32> rlen(R.n_seq);
  CK(cudaMemcpyAsync(rlen.data(), R.len.p, 4ull * R.n_seq, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  std::vector<SeqDesc> sd(n); std::vector<StageReq> rq(n);
  u64 units = 0, total = 0;
  for (u32 i = 0; i < n; ++i) {
    if (ref_seq[i] >= R.n_seq || start[i] == 0 || stop[i] == 0) { set_error("bad fetch request"); return TRPA_ERR_ARG; }
    // same arithmetic as Machine::stage_candidate / emit_stage (hh:870-880)
    u64 ns, ne; u32 rev = 0;
    if (start[i] <= stop[i]) { ns = left_ext[i] < start[i] ? start[i] - left_ext[i] : 1; ne = (u64)stop[i] + right_ext[i]; }
    else { ns = right_ext[i] < stop[i] ? stop[i] - right_ext[i] : 1; ne = (u64)start[i] + left_ext[i]; rev = protein ? 0 : 1; }
    const u64 L = rlen[ref_seq[i]];
    if (ne > L) ne = L;
    u64 b = ns - 1; if (b > L) b = L;
    u64 e = ne > b ? ne : b; if (e > L) e = L;
    const u32 len = (u32)(e - b);
    sd[i] = SeqDesc{(u32)units, len, 0, 0};
    rq[i] = StageReq{i, 1, ref_seq[i], (u32)b, rev};
    units += protein ? ((len + 3u) & ~3u) : ((len + 31u) >> 5);
    out_off[i] = total; out_len[i] = len; total += len;
  }
  if (total > out_capacity) { set_error("output buffer too small"); return TRPA_ERR_ARG; }
  if (units >= 0xffffffffull) { set_error("fetch too large"); return TRPA_ERR_ARG; }
  ScopedBuf<SeqDesc> d_sd; ScopedBuf<StageReq> d_rq; ScopedBuf<uint8_t> d_out; ScopedBuf<u64> d_ooff;
  if (d_sd.ensure(n + 1) || d_rq.ensure(n + 1) || d_out.ensure(total + 16) || d_ooff.ensure(n + 1)) return TRPA_ERR_NOMEM;
  CK(cudaMemcpyAsync(d_sd.p, sd.data(), sizeof(SeqDesc) * n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_rq.p, rq.data(), sizeof(StageReq) * n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_ooff.p, out_off, 8ull * n, cudaMemcpyHostToDevice, c->stream));
  if (protein) {
    ScopedBuf<uint8_t> stg;
    if (stg.ensure(units + 16)) return TRPA_ERR_NOMEM;
    CK(ensure_blosum_constant(c->device));
    CK(launch_stage_aa(d_rq.p, n, nullptr, nullptr, R.packed.p, R.woff.p, d_sd.p, stg.p, c->stream));
    for (u32 i = 0; i < n; ++i)
      if (out_len[i]) CK(cudaMemcpyAsync(d_out.p + out_off[i], stg.p + sd[i].woff, out_len[i], cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(out_codes, d_out.p, total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    stg.release();
  } else {
    ScopedBuf<uint2> pl; ScopedBuf<u32> pn;
    if (pl.ensure(units + 2) || pn.ensure(units + 2)) return TRPA_ERR_NOMEM;
    CK(launch_stage_nt(d_rq.p, n, nullptr, nullptr, nullptr, R.planes.p, R.nplane.p, R.woff.p, d_sd.p, pl.p, pn.p, nullptr, c->stream));
    CK(launch_unstage_nt(d_sd.p, n, pl.p, pn.p, d_ooff.p, d_out.p, c->stream));
    CK(cudaMemcpyAsync(out_codes, d_out.p, total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    pl.release(); pn.release();
  }
  d_sd.release(); d_rq.release(); d_out.release(); d_ooff.release();
  return 0;
}

int trpa_predict_lca_batch(trpa_ctx* c, const trpa_lca_params* pp, const trpa_segment* segs, uint32_t n_segs,
                           const trpa_candidate* cands, uint32_t n_cands, const double* evalue,
                           const uint8_t* node_unclassified, trpa_result* out, int repeat, double* kernel_ms) {
  if (!c || !pp || (n_segs && (!segs || !out)) || (n_cands && !cands)) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (pp->model > TRPA_MODEL_NBEST_LCA) { set_error("unknown placement model"); return TRPA_ERR_ARG; }
  if (c->n_nodes == 0) { set_error("taxonomy not loaded"); return TRPA_ERR_STATE; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  if (kernel_ms) *kernel_ms = 0.0;
  if (n_segs == 0) return 0;
  // one validation pass over the tables (the same contract as trpa_batch_upload: contiguous record sets)
  u64 expect = 0;
  for (u32 s = 0; s < n_segs; ++s) {
    if (segs[s].cand_begin != expect || (u64)segs[s].cand_begin + segs[s].cand_count > n_cands) {
      set_error("segment table: candidate ranges must be contiguous and inside the candidate table");
      return TRPA_ERR_ARG;
    }
    expect += segs[s].cand_count;
  }
  for (u32 i = 0; i < n_cands; ++i)
    if (cands[i].node >= c->n_nodes) { set_error("candidate node out of range"); return TRPA_ERR_ARG; }
  DevBuf<trpa_segment>& d_segs = c->l_segs; DevBuf<trpa_candidate>& d_cands = c->l_cands; DevBuf<double>& d_ev = c->l_ev;
  DevBuf<uint8_t>& d_un = c->l_un; DevBuf<trpa_result>& d_out = c->l_out;
  if (d_segs.ensure(n_segs) || d_cands.ensure(n_cands + 1) || d_out.ensure(n_segs) || (evalue && d_ev.ensure(n_cands + 1)) ||
      (node_unclassified && d_un.ensure(c->n_nodes)))
    return TRPA_ERR_NOMEM;
  CK(cudaMemcpyAsync(d_segs.p, segs, sizeof(trpa_segment) * n_segs, cudaMemcpyHostToDevice, c->stream));
  if (n_cands) CK(cudaMemcpyAsync(d_cands.p, cands, sizeof(trpa_candidate) * n_cands, cudaMemcpyHostToDevice, c->stream));
  if (evalue && n_cands) CK(cudaMemcpyAsync(d_ev.p, evalue, sizeof(double) * n_cands, cudaMemcpyHostToDevice, c->stream));
  if (node_unclassified) CK(cudaMemcpyAsync(d_un.p, node_unclassified, c->n_nodes, cudaMemcpyHostToDevice, c->stream));
  if (repeat < 1) repeat = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const Taxonomy tax = dev_tax(c);
  if (kernel_ms && repeat > 1)   // one untimed pass when timing is requested
    CK(launch_lca_models(d_segs.p, n_segs, d_cands.p, evalue ? d_ev.p : nullptr, node_unclassified ? d_un.p : nullptr, tax, *pp, d_out.p, c->stream));
  CK(cudaEventRecord(e0, c->stream));
  for (int r = 0; r < repeat; ++r)
    CK(launch_lca_models(d_segs.p, n_segs, d_cands.p, evalue ? d_ev.p : nullptr, node_unclassified ? d_un.p : nullptr, tax, *pp, d_out.p, c->stream));
  CK(cudaEventRecord(e1, c->stream));
  CK(cudaMemcpyAsync(out, d_out.p, sizeof(trpa_result) * n_segs, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (kernel_ms) *kernel_ms = ms / repeat;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  c->prof.launches_other += (u64)repeat;
  return 0;
}

// ------------------------------------------------------------------------------------------ binner
int trpa_bin_batch(trpa_ctx* c, const trpa_bin_params* pp, const trpa_bin_record* records, uint32_t n_records,
                   const uint32_t* supports, uint32_t n_supports, const uint32_t* group_begin, uint32_t n_groups,
                   const uint8_t* rank_of_node, const float* pid_per_rank, trpa_bin_result* out, trpa_bin_stats* stats) {
  if (!c || !pp || (n_records && (!records || !supports)) || !group_begin || (n_groups && !out)) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (c->n_nodes == 0) { set_error("taxonomy not loaded"); return TRPA_ERR_STATE; }
  if (pp->n_ranks && (!rank_of_node || !pid_per_rank)) { set_error("identity constraints need rank_of_node and pid_per_rank"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  // validation: groups contiguous and in order; ranges are ancestor paths; supports slices inside the table
  if (group_begin[0] != 0 || group_begin[n_groups] != n_records) { set_error("group table must cover the record table"); return TRPA_ERR_ARG; }
  for (u32 g = 0; g < n_groups; ++g) if (group_begin[g] > group_begin[g + 1]) { set_error("group table not ascending"); return TRPA_ERR_ARG; }
  if (n_records) {
    std::vector<uint8_t> h_depth(c->n_nodes);
    std::vector<u32> h_parent(c->n_nodes);
    CK(cudaMemcpyAsync(h_depth.data(), c->t_depth.p, c->n_nodes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_parent.data(), c->t_parent.p, 4ull * c->n_nodes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (u32 r = 0; r < n_records; ++r) {
      const trpa_bin_record& b = records[r];
      if (b.lower_node >= c->n_nodes || b.upper_node >= c->n_nodes || h_depth[b.lower_node] < h_depth[b.upper_node]) { set_error("prediction record: bad node range"); return TRPA_ERR_ARG; }
      u32 x = b.lower_node;
      while (h_depth[x] > h_depth[b.upper_node]) x = h_parent[x];
      if (x != b.upper_node) { set_error("prediction record: upper node is not an ancestor of the lower node"); return TRPA_ERR_ARG; }
      if ((u64)b.support_begin + (h_depth[b.lower_node] - h_depth[b.upper_node] + 1u) > n_supports) { set_error("prediction record: support slice out of range"); return TRPA_ERR_ARG; }
    }
  }
  const u32 D1 = c->max_depth + 1;
  ScopedBuf<trpa_bin_record> d_rec; ScopedBuf<u32> d_sup, d_gb, d_nodes3, d_lower, d_cur, d_majn, d_pathn, d_stats;
  ScopedBuf<float> d_majs, d_pid; ScopedBuf<uint8_t> d_alive, d_state, d_pathb, d_rank; ScopedBuf<unsigned short> d_tot, d_pathd, d_patht;
  ScopedBuf<trpa_bin_result> d_out;
  if (d_rec.ensure(n_records + 1) || d_sup.ensure(n_supports + 1) || d_gb.ensure(n_groups + 1) || d_nodes3.ensure(3ull * c->n_nodes) ||
      d_lower.ensure(n_records + 1) || d_cur.ensure(n_records + 1) || d_majn.ensure(n_records + 1) || d_majs.ensure(n_records + 1) ||
      d_alive.ensure(n_records + 1) || d_state.ensure(n_records + 1) || d_tot.ensure((size_t)(n_records + 1) * D1) ||
      d_pathn.ensure((size_t)(n_groups + 1) * D1) || d_pathd.ensure((size_t)(n_groups + 1) * D1) || d_patht.ensure((size_t)(n_groups + 1) * D1) ||
      d_pathb.ensure((size_t)(n_groups + 1) * D1) || d_out.ensure(n_groups + 1) || d_stats.ensure(4) ||
      (pp->n_ranks && (d_rank.ensure(c->n_nodes) || d_pid.ensure(pp->n_ranks))))
    return TRPA_ERR_NOMEM;
  struct SyncGuard { cudaStream_t st; ~SyncGuard() { cudaStreamSynchronize(st); } } sync_guard{c->stream};
  if (n_records) CK(cudaMemcpyAsync(d_rec.p, records, sizeof(trpa_bin_record) * n_records, cudaMemcpyHostToDevice, c->stream));
  if (n_supports) CK(cudaMemcpyAsync(d_sup.p, supports, 4ull * n_supports, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_gb.p, group_begin, 4ull * (n_groups + 1), cudaMemcpyHostToDevice, c->stream));
  if (pp->n_ranks) {
    for (u32 i = 0; i < c->n_nodes; ++i) if (rank_of_node[i] >= pp->n_ranks) { set_error("rank_of_node out of range"); return TRPA_ERR_ARG; }
    CK(cudaMemcpyAsync(d_rank.p, rank_of_node, c->n_nodes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_pid.p, pid_per_rank, sizeof(float) * pp->n_ranks, cudaMemcpyHostToDevice, c->stream));
  }
  BinScratch S;
  S.node_support = d_nodes3.p; S.node_seen = d_nodes3.p + c->n_nodes; S.node_pruned = d_nodes3.p + 2ull * c->n_nodes;
  S.lower = d_lower.p; S.alive = d_alive.p; S.state = d_state.p; S.curnode = d_cur.p; S.maj_node = d_majn.p; S.maj_sum = d_majs.p;
  S.tot = d_tot.p; S.path_node = d_pathn.p; S.path_direct = d_pathd.p; S.path_total = d_patht.p; S.path_branch = d_pathb.p;
  CK(launch_binner(d_rec.p, n_records, d_sup.p, d_gb.p, n_groups, dev_tax(c), c->n_nodes, c->max_depth, *pp,
                   pp->n_ranks ? d_rank.p : nullptr, pp->n_ranks ? d_pid.p : nullptr, S, d_out.p, d_stats.p, c->stream));
  c->prof.launches_other += 4;
  u32 h_stats[4] = {0, 0, 0, 0};
  u32 h_root = 0;
  if (n_groups) CK(cudaMemcpyAsync(out, d_out.p, sizeof(trpa_bin_result) * n_groups, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(h_stats, d_stats.p, sizeof(h_stats), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&h_root, S.node_support + c->root, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (stats) { stats->nested_taxa = h_stats[0]; stats->pruned_taxa = h_stats[1]; stats->min_support_found = h_stats[2]; stats->root_support = h_root; }
  return 0;
}

// Host-only helper (no GPU work): where to cut a segment table into `world` contiguous shards of about equal
// DP work.  weight(segment) = 1 + sum over its records of span^2 (span = reference range of the record): the
// number of realignments grows with the record count and each costs about |A| * |B| ~ span^2 cells.
int trpa_shard_bounds(const trpa_segment* segs, uint32_t n_segs, const trpa_candidate* cands, uint32_t n_cands,
                      uint32_t world, uint32_t* bounds) {
  if (!bounds || world == 0 || (n_segs && !segs) || (n_cands && !cands)) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  std::vector<u64> w(n_segs);
  unsigned __int128 total = 0;
  for (u32 s = 0; s < n_segs; ++s) {
    if ((u64)segs[s].cand_begin + segs[s].cand_count > n_cands) { set_error("segment candidate range out of bounds"); return TRPA_ERR_ARG; }
    u64 acc = 1;
    const trpa_candidate* rc = cands + segs[s].cand_begin;
    for (u32 k = 0; k < segs[s].cand_count; ++k) {
      const u64 span = (rc[k].rstart <= rc[k].rstop ? (u64)rc[k].rstop - rc[k].rstart : (u64)rc[k].rstart - rc[k].rstop) + 1;
      acc += span * span;   // span < 2^32: no overflow of the square; the sum saturates far beyond any real table
    }
    w[s] = acc;
    total += acc;
  }
  bounds[0] = 0;
  u32 r = 1;
  unsigned __int128 acc = 0;
  for (u32 s = 0; s < n_segs && r < world; ++s) {
    acc += w[s];
    while (r < world && acc * world >= total * r) bounds[r++] = s + 1;
  }
  while (r <= world) bounds[r++] = n_segs;
  return 0;
}

int trpa_lca_batch(trpa_ctx* c, const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* out) {
  if (!c || !a || !b || !out) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (c->n_nodes == 0) { set_error("taxonomy not loaded"); return TRPA_ERR_STATE; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  for (u32 i = 0; i < n; ++i) if (a[i] >= c->n_nodes || b[i] >= c->n_nodes) { set_error("node out of range"); return TRPA_ERR_ARG; }
  if (n == 0) return 0;
  ScopedBuf<u32> da, db, dout;
  if (da.ensure(n) || db.ensure(n) || dout.ensure(n)) return TRPA_ERR_NOMEM;
  CK(cudaMemcpyAsync(da.p, a, 4ull * n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(db.p, b, 4ull * n, cudaMemcpyHostToDevice, c->stream));
  lca_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(dev_tax(c), da.p, db.p, n, dout.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, dout.p, 4ull * n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  da.release(); db.release(); dout.release();
  return 0;
}

int trpa_int_alu_peak(trpa_ctx* c, double* lane_ops_per_s) {
  if (!c || !lane_ops_per_s) { set_error("bad arguments"); return TRPA_ERR_ARG; }
  if (use_device(c)) return TRPA_ERR_CUDA;
  ScopedBuf<u32> sink;
  if (sink.ensure(4)) return TRPA_ERR_NOMEM;
  int blocks = 0, threads = 0;
  const int iters = 20000;
  CK(launch_alu_probe(sink.p, 2000, c->stream, &blocks, &threads));  // warm-up
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0, c->stream));
    CK(launch_alu_probe(sink.p, iters, c->stream, &blocks, &threads));
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double ops = (double)blocks * threads * (double)iters * alu_probe_ops_per_iter();
    best = std::max(best, ops / (ms * 1e-3));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  sink.release();
  *lane_ops_per_s = best;
  return 0;
}

}  // extern "C"
