// eesen_b200/csrc/decode.cu -- one-best WFST token passing for a batch of utterances (SURVEY.md 8f row N3,
// BASELINE config 5: latgen-faster over TLG.fst, 64 utterances in parallel).
//
// Replaces the search core of the reference's LatticeFasterDecoder (src/decoder/lattice-faster-decoder.cc):
//   InitDecoding :53-71, Decode :77-97, GetCutoff :594-658 (beam, max_active, min_active), ProcessEmitting :660-752,
//   ProcessNonemitting :756-816, ComputeFinalCosts :531-577, with the acoustic scores of DecodableMatrixScaled
//   (src/decoder/decodable-matrix.h:54-56) read straight from the packed posterior matrix.
// First slice: the best path (words = non-zero olabels, and its cost); lattices are out of scope.
//
// The reference walks one utterance at a time through hash lists of heap-allocated tokens and OpenFst arc iterators.
// Here the graph is a flat CSR in HBM (emitting arcs of a state first, then its epsilon arcs; 16 bytes per arc) and
// every utterance owns dense per-state arrays, which 180 GB of HBM make affordable (a 3 M-state graph x 64
// utterances = 1.5 GB for the 8-byte cost/back-pointer cells):
//   cell[u][state]   u64: (order-preserving bits of the best cost so far << 32) | id of the arc that achieved it.
//                    One atomicMin per relaxation does FindOrAddToken (:146-167) and keeps the back-pointer;
//                    ties go to the smaller arc id (the reference: to whichever arc its hash order met first).
//   slot[u][state]   index of the state's token in the frame being built (assigned by whoever touches a state first).
// Per frame: expand the surviving tokens over their emitting arcs (thread per token), close over epsilon arcs to a
// fix point (warp per state: the arcs of high-degree states -- the LM hub -- are strided over the lanes), then turn
// the touched cells into the frame's token records and reset them.  The order-dependent online tightening of
// next_cutoff (:684-700, :727-729) is not reproduced: it only drops tokens more than `beam` above the new frame's
// best, which neither the next frame (:716) nor the closure (:775) would expand.
#include "common.cuh"
#include "kernels.h"

namespace eb {

namespace {

constexpr unsigned long long kCellInf = ~0ull;
constexpr uint32_t kNoArc = 0xffffffffu;

__device__ __forceinline__ uint32_t ord_bits(float c) {   // monotone float -> uint
  const uint32_t b = __float_as_uint(c);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct DecodeState {
  // graph
  const int *row, *eps, *ilabel, *olabel, *nextstate, *arc_from;
  const float *weight, *final_cost;
  int num_states;
  // per utterance (u) arrays
  unsigned long long *cell;   // [S][num_states]
  int *slot_cur, *slot_nxt;   // [S][num_states]
  int *touched;               // [S][frame_cap]   states of the frame being built, index = slot
  int *wl_a, *wl_b;           // [S][wl_cap]      epsilon-closure work lists
  int *n_touched, *n_wl_a, *n_wl_b;   // [S]
  uint32_t *frame_best;       // [S] ordered bits of the best token cost of the frame just finished
  uint32_t *build_best;       // [S] same for the frame under construction (before the closure)
  float *cutoff;              // [S] GetCutoff of the frame being expanded (beam, max_active, min_active)
  double *offset_sum;         // [S]
  int *err;                   // bit 0: frame_cap overflow, bit 1: work-list overflow, bit 2: token store overflow
  // token store, per utterance [S][tok_cap]
  int *tok_state, *tok_prev, *tok_olabel;
  float *tok_cost;
  int frame_cap, wl_cap, tok_cap;
};

// first touch of a state in the frame being built: give it a slot
__device__ __forceinline__ void first_touch(const DecodeState &d, int u, int ns) {
  const int p = atomicAdd(&d.n_touched[u], 1);
  if (p < d.frame_cap) {
    d.touched[(size_t)u * d.frame_cap + p] = ns;
    d.slot_nxt[(size_t)u * d.num_states + ns] = p;
  } else {
    atomicOr(d.err, 1);
  }
}

__global__ void decode_init_kernel(DecodeState d, int S, int start) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= S) return;
  d.cell[(size_t)u * d.num_states + start] = ((unsigned long long)ord_bits(0.f) << 32) | kNoArc;
  d.touched[(size_t)u * d.frame_cap] = start;
  d.slot_nxt[(size_t)u * d.num_states + start] = 0;
  d.n_touched[u] = 1;
  d.build_best[u] = ord_bits(0.f);
  d.offset_sum[u] = 0.0;
  d.n_wl_a[u] = 0; d.n_wl_b[u] = 0;
}

// GetCutoff :594-658 -- one CTA per utterance.  best + beam, tightened to the cost of the max_active-th best token when
// more than max_active tokens are alive (:626-631), loosened to the min_active-th best when the beam would keep fewer
// (:632-650).  The k-th smallest cost is found EXACTLY (the reference uses std::nth_element): a most-significant-byte-
// first radix select over the order-preserving bit patterns of the costs, four 256-bin histogram passes in shared memory.
__device__ uint32_t kth_smallest_bits(const float *cost, int n, int k, unsigned *hist, uint32_t *sh_prefix, int *sh_k) {
  // (all threads of the CTA call this; returns the ordered bits of the k-th smallest, 0-based)
  uint32_t prefix = 0, mask = 0;
  for (int pass = 3; pass >= 0; pass--) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint32_t key = ord_bits(cost[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int kk = k, b = 0;
      for (; b < 255; b++) {
        if (kk < (int)hist[b]) break;
        kk -= (int)hist[b];
      }
      *sh_prefix = prefix | ((uint32_t)b << shift);
      *sh_k = kk;
    }
    __syncthreads();
    prefix = *sh_prefix;
    k = *sh_k;
    mask |= 255u << shift;
    __syncthreads();
  }
  return prefix;
}

__global__ void decode_cutoff_kernel(DecodeState d, int S, int t, const int *frames, const int *f0v, const int *nv,
                                     float beam, int max_active, int min_active) {
  __shared__ unsigned hist[256];
  __shared__ uint32_t sh_prefix;
  __shared__ int sh_k;
  const int u = blockIdx.x;
  if (t >= frames[u]) return;
  const int n = nv[u];
  const float best = ord_float(d.frame_best[u]);
  float cut = best + beam;                                                       // :604-609 / :646
  const float *cost = d.tok_cost + (size_t)u * d.tok_cap + f0v[u];
  float max_c = INFINITY, min_c = INFINITY;
  if (n > max_active) max_c = ord_float(kth_smallest_bits(cost, n, max_active, hist, &sh_prefix, &sh_k));     // :626-631
  if (n > min_active) min_c = min_active == 0 ? best : ord_float(kth_smallest_bits(cost, n, min_active, hist, &sh_prefix, &sh_k));   // :636-644
  if (max_c < cut) cut = max_c;                                                  // :632-635
  else if (min_c > cut) cut = min_c;                                             // :645-650
  if (threadIdx.x == 0) d.cutoff[u] = cut;
}

// ProcessEmitting :660-752 -- one thread per token of frame t (tokens [f0, f0 + n) of utterance u)
__global__ void decode_expand_kernel(DecodeState d, int S, int t, const int *frames, const int *f0v, const int *nv,
                                     const float *loglikes, int ld, float scale, float beam, int limits) {
  const int u = blockIdx.y;
  if (t >= frames[u]) return;
  const int n = nv[u], f0 = f0v[u];
  const float best = ord_float(d.frame_best[u]);
  const float cutoff = limits ? d.cutoff[u] : best + beam, cost_offset = -best;   // GetCutoff (decode_cutoff_kernel) / :604-609, :689
  if (blockIdx.x == 0 && threadIdx.x == 0) d.offset_sum[u] += (double)cost_offset;
  const float *ll = loglikes + ((size_t)t * S + u) * ld;
  unsigned long long *cell = d.cell + (size_t)u * d.num_states;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const size_t ti = (size_t)u * d.tok_cap + f0 + i;
    const float cur_cost = d.tok_cost[ti];
    if (!(cur_cost <= cutoff)) continue;                                       // :716
    const int st = d.tok_state[ti];
    for (int a = d.row[st]; a < d.eps[st]; a++) {
      const float ac_cost = cost_offset - scale * __ldg(ll + d.ilabel[a] - 1);   // :722-723, decodable-matrix.h:54-56
      const float tot = cur_cost + ac_cost + d.weight[a];                      // :724-726, left to right
      const int ns = d.nextstate[a];
      const unsigned long long key = ((unsigned long long)ord_bits(tot) << 32) | (uint32_t)a;
      const unsigned long long old = atomicMin(&cell[ns], key);
      if (old == kCellInf) first_touch(d, u, ns);
    }
  }
}

// start of ProcessNonemitting :763-775: every token of the frame under construction goes on the work list,
// best cost -> closure cutoff
__global__ void decode_closure_begin_kernel(DecodeState d, int S, int t, const int *frames) {
  const int u = blockIdx.y;
  if (t >= frames[u]) return;   // no frame t+1 to build for this utterance
  const int n = min(d.n_touched[u], d.frame_cap);
  const unsigned long long *cell = d.cell + (size_t)u * d.num_states;
  uint32_t mn = 0xffffffffu;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int st = d.touched[(size_t)u * d.frame_cap + i];
    if (i < d.wl_cap) d.wl_a[(size_t)u * d.wl_cap + i] = st; else atomicOr(d.err, 2);
    mn = min(mn, (uint32_t)(cell[st] >> 32));
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  if ((threadIdx.x & 31) == 0 && mn != 0xffffffffu) atomicMin(&d.build_best[u], mn);
  if (blockIdx.x == 0 && threadIdx.x == 0) { d.n_wl_a[u] = min(n, d.wl_cap); d.n_wl_b[u] = 0; }
}

// one round of the closure: a warp per work-list state, lanes over its epsilon arcs (:777-815)
__global__ void decode_closure_round_kernel(DecodeState d, int S, int t, const int *frames, const int *wl_in,
                                            const int *n_in, int *wl_out, int *n_out, float beam) {
  const int u = blockIdx.y;
  if (t >= frames[u]) return;   // no frame t+1 to build for this utterance
  const int n = n_in[u];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const float cutoff = ord_float(d.build_best[u]) + beam;                      // :775
  unsigned long long *cell = d.cell + (size_t)u * d.num_states;
  for (int i = warp; i < n; i += nwarps) {
    const int st = wl_in[(size_t)u * d.wl_cap + i];
    const float cur_cost = ord_float((uint32_t)(cell[st] >> 32));
    if (cur_cost > cutoff) continue;                                           // :782-783
    for (int a = d.eps[st] + lane; a < d.row[st + 1]; a += 32) {
      const float tot = cur_cost + d.weight[a];                                // :797-798
      if (tot < cutoff) {                                                      // :799
        const int ns = d.nextstate[a];
        const unsigned long long key = ((unsigned long long)ord_bits(tot) << 32) | (uint32_t)a;
        const unsigned long long old = atomicMin(&cell[ns], key);
        if (key < old) {                                                       // "changed" :803-811
          if (old == kCellInf) first_touch(d, u, ns);
          const int p = atomicAdd(&n_out[u], 1);
          if (p < d.wl_cap) wl_out[(size_t)u * d.wl_cap + p] = ns; else atomicOr(d.err, 2);
        }
      }
    }
  }
}

// the touched cells become the token records of frame t+1 (tokens [f1, f1 + n)); cells are reset for the next frame
__global__ void decode_finalize_kernel(DecodeState d, int S, int t, const int *frames, const int *f0v, const int *f1v) {
  const int u = blockIdx.y;
  if (t >= frames[u]) return;   // no frame t+1 to build for this utterance
  const int n = min(d.n_touched[u], d.frame_cap);
  const int f0 = f0v[u], f1 = f1v[u];
  unsigned long long *cell = d.cell + (size_t)u * d.num_states;
  const int *slot_cur = d.slot_cur + (size_t)u * d.num_states, *slot_nxt = d.slot_nxt + (size_t)u * d.num_states;
  uint32_t mn = 0xffffffffu;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    if (f1 + p >= d.tok_cap) { atomicOr(d.err, 4); break; }
    const int st = d.touched[(size_t)u * d.frame_cap + p];
    const unsigned long long key = cell[st];
    cell[st] = kCellInf;
    const uint32_t arc = (uint32_t)key;
    const size_t ti = (size_t)u * d.tok_cap + f1 + p;
    d.tok_state[ti] = st;
    d.tok_cost[ti] = ord_float((uint32_t)(key >> 32));
    if (arc == kNoArc) { d.tok_prev[ti] = -1; d.tok_olabel[ti] = 0; }
    else {
      const int src = d.arc_from[arc];
      d.tok_prev[ti] = d.ilabel[arc] != 0 ? f0 + slot_cur[src] : f1 + slot_nxt[src];
      d.tok_olabel[ti] = d.olabel[arc];
    }
    mn = min(mn, (uint32_t)(key >> 32));
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  if ((threadIdx.x & 31) == 0 && mn != 0xffffffffu) atomicMin(&d.frame_best[u], mn);
}

// ComputeFinalCosts :531-577 over the tokens of the last frame, then the back-trace (one thread per utterance)
__global__ void decode_backtrace_kernel(DecodeState d, int S, const int *f_last, const int *n_last, int *out_labels,
                                        int max_out, int *out_len, float *out_cost) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= S) return;
  const size_t base = (size_t)u * d.tok_cap;
  float best_final = INFINITY, best_plain = INFINITY;
  int tf = -1, tp = -1;
  for (int i = 0; i < n_last[u]; i++) {
    const int ti = f_last[u] + i;
    const float c = d.tok_cost[base + ti];
    if (c < best_plain) { best_plain = c; tp = ti; }
    const float f = d.final_cost[d.tok_state[base + ti]];
    if (f != INFINITY && c + f < best_final) { best_final = c + f; tf = ti; }
  }
  int bt = tf >= 0 ? tf : tp;
  if (bt < 0) { out_len[u] = -1; out_cost[u] = INFINITY; return; }
  const float path_cost = tf >= 0 ? best_final : best_plain;
  int cnt = 0;
  for (int k = bt; k >= 0; k = d.tok_prev[base + k]) if (d.tok_olabel[base + k] != 0) cnt++;
  int w = cnt;
  for (int k = bt; k >= 0; k = d.tok_prev[base + k])
    if (d.tok_olabel[base + k] != 0) { w--; if (w < max_out) out_labels[(size_t)u * max_out + w] = d.tok_olabel[base + k]; }
  out_len[u] = min(cnt, max_out);
  out_cost[u] = (float)((double)path_cost - d.offset_sum[u]);   // the per-frame offsets were added to every path
}

__global__ void fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_u32_kernel(uint32_t *p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

size_t decode_workspace_bytes(int S, int num_states, int frame_cap, int wl_cap, int tok_cap) {
  size_t b = 0;
  b += (size_t)S * num_states * 8;                 // cell
  b += (size_t)S * num_states * 4 * 2;             // slot maps
  b += (size_t)S * frame_cap * 4;                  // touched
  b += (size_t)S * wl_cap * 4 * 2;                 // work lists
  b += (size_t)S * 4 * 16 + 1024;                  // counters, bests, offsets, cutoffs, err
  b += (size_t)S * tok_cap * 16;                   // token store
  return b + 4096;
}

// Everything device-side; `frames` etc. are small host arrays.  Returns the error bits of DecodeState::err in *err_bits.
cudaError_t decode_best_path(cudaStream_t st, int num_sms, const DecodeGraph &g, int S, int T, const int *h_frames,
                             const float *d_loglikes, int ld, float scale, float beam, int max_active, int min_active,
                             void *ws, int frame_cap, int wl_cap, int tok_cap, int *d_out_labels, int max_out,
                             int *d_out_len, float *d_out_cost, int *err_bits, long *closure_rounds) {
  if (S <= 0) return cudaSuccess;
  // carve the workspace
  char *p = (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  auto take = [&](size_t bytes) { void *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  DecodeState d;
  d.row = g.row; d.eps = g.eps; d.ilabel = g.ilabel; d.olabel = g.olabel; d.nextstate = g.nextstate; d.arc_from = g.arc_from;
  d.weight = g.weight; d.final_cost = g.final_cost; d.num_states = g.num_states;
  d.cell = (unsigned long long *)take((size_t)S * g.num_states * 8);
  d.slot_cur = (int *)take((size_t)S * g.num_states * 4);
  d.slot_nxt = (int *)take((size_t)S * g.num_states * 4);
  d.touched = (int *)take((size_t)S * frame_cap * 4);
  d.wl_a = (int *)take((size_t)S * wl_cap * 4);
  d.wl_b = (int *)take((size_t)S * wl_cap * 4);
  d.n_touched = (int *)take((size_t)S * 4); d.n_wl_a = (int *)take((size_t)S * 4); d.n_wl_b = (int *)take((size_t)S * 4);
  d.frame_best = (uint32_t *)take((size_t)S * 4); d.build_best = (uint32_t *)take((size_t)S * 4);
  d.offset_sum = (double *)take((size_t)S * 8);
  d.cutoff = (float *)take((size_t)S * 4);
  d.err = (int *)take(4);
  int *d_frames = (int *)take((size_t)S * 4), *d_f0 = (int *)take((size_t)S * 4), *d_f1 = (int *)take((size_t)S * 4),
      *d_n = (int *)take((size_t)S * 4);
  d.tok_state = (int *)take((size_t)S * tok_cap * 4); d.tok_prev = (int *)take((size_t)S * tok_cap * 4);
  d.tok_olabel = (int *)take((size_t)S * tok_cap * 4); d.tok_cost = (float *)take((size_t)S * tok_cap * 4);
  d.frame_cap = frame_cap; d.wl_cap = wl_cap; d.tok_cap = tok_cap;

  cudaError_t e;
  fill_u64_kernel<<<4 * num_sms, 256, 0, st>>>(d.cell, (size_t)S * g.num_states, kCellInf);
  if ((e = cudaMemsetAsync(d.err, 0, 4, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(d_frames, h_frames, sizeof(int) * S, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  decode_init_kernel<<<(S + 127) / 128, 128, 0, st>>>(d, S, g.start);

  // host mirrors of the per-utterance frame bookkeeping
  int *h_f0 = new int[4 * S], *h_f1 = h_f0 + S, *h_n = h_f0 + 2 * S, *h_cnt = h_f0 + 3 * S;
  for (int u = 0; u < S; u++) { h_f0[u] = 0; h_f1[u] = 0; h_n[u] = 0; }
  long rounds = 0;
  const dim3 blk(256);
  auto grid_for = [&](int n_max, int per_thread_items) {
    int b = (n_max * per_thread_items + 255) / 256;
    if (b < 1) b = 1;
    if (b > 2 * num_sms) b = 2 * num_sms;
    return dim3(b, S);
  };
  // frame 0 = the start token + its closure (InitDecoding); frame t+1 is built from frame t
  for (int t = -1; t < T; t++) {
    if (t >= 0) {
      // tokens of frame t live at [h_f0, h_f0 + h_n); frame t+1 goes to [h_f1, ...)
      if ((e = cudaMemcpyAsync(d_f0, h_f0, sizeof(int) * S, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync(d_n, h_n, sizeof(int) * S, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
      if ((e = cudaMemsetAsync(d.n_touched, 0, sizeof(int) * S, st)) != cudaSuccess) break;
      int n_max = 0;
      for (int u = 0; u < S; u++) if (t < h_frames[u] && h_n[u] > n_max) n_max = h_n[u];
      const int limits = !(max_active == 2147483647 && min_active == 0);
      if (limits) decode_cutoff_kernel<<<S, 1024, 0, st>>>(d, S, t, d_frames, d_f0, d_n, beam, max_active, min_active);
      decode_expand_kernel<<<grid_for(n_max, 1), blk, 0, st>>>(d, S, t, d_frames, d_f0, d_n, d_loglikes, ld, scale, beam, limits);
    }
    fill_u32_kernel<<<1, 64, 0, st>>>(d.build_best, (size_t)S, 0xffffffffu);
    // closure of the frame under construction (frame index t + 1; utterances with frames > t take part)
    decode_closure_begin_kernel<<<dim3(2 * num_sms, S), blk, 0, st>>>(d, S, t, d_frames);
    int *wl_in = d.wl_a, *n_in = d.n_wl_a, *wl_out = d.wl_b, *n_out = d.n_wl_b;
    while (true) {
      decode_closure_round_kernel<<<dim3(2 * num_sms, S), blk, 0, st>>>(d, S, t, d_frames, wl_in, n_in, wl_out, n_out, beam);
      rounds++;
      if ((e = cudaMemcpyAsync(h_cnt, n_out, sizeof(int) * S, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
      bool any = false;
      for (int u = 0; u < S; u++) if (t < h_frames[u] && h_cnt[u] > 0) any = true;   // (t = -1: every utterance)
      if (!any) break;
      if ((e = cudaMemsetAsync(n_in, 0, sizeof(int) * S, st)) != cudaSuccess) break;
      int *tw = wl_in; wl_in = wl_out; wl_out = tw;
      int *tn = n_in; n_in = n_out; n_out = tn;
    }
    if (e != cudaSuccess) break;
    // finalize frame t+1
    if ((e = cudaMemcpyAsync(h_cnt, d.n_touched, sizeof(int) * S, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
    for (int u = 0; u < S; u++) {
      if (t >= h_frames[u]) continue;          // this utterance is finished: its last frame stays where it is
      h_f1[u] = h_f0[u] + h_n[u];
    }
    if ((e = cudaMemcpyAsync(d_f0, h_f0, sizeof(int) * S, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(d_f1, h_f1, sizeof(int) * S, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    fill_u32_kernel<<<1, 64, 0, st>>>(d.frame_best, (size_t)S, 0xffffffffu);
    decode_finalize_kernel<<<dim3(2 * num_sms, S), blk, 0, st>>>(d, S, t, d_frames, d_f0, d_f1);
    for (int u = 0; u < S; u++) {
      if (t >= h_frames[u]) continue;
      h_f0[u] = h_f1[u];
      h_n[u] = h_cnt[u] < frame_cap ? h_cnt[u] : frame_cap;
    }
    // the maps swap roles: what was "next" is the current frame's state -> slot map now
    int *ts = d.slot_cur; d.slot_cur = d.slot_nxt; d.slot_nxt = ts;
  }
  if (e == cudaSuccess) {
    e = cudaMemcpyAsync(d_f0, h_f0, sizeof(int) * S, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_n, h_n, sizeof(int) * S, cudaMemcpyHostToDevice, st);
    decode_backtrace_kernel<<<(S + 63) / 64, 64, 0, st>>>(d, S, d_f0, d_n, d_out_labels, max_out, d_out_len, d_out_cost);
    if (e == cudaSuccess) e = cudaMemcpyAsync(err_bits, d.err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  if (closure_rounds) *closure_rounds = rounds;
  delete[] h_f0;
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace eb
