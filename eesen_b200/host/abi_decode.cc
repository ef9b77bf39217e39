// eesen_b200/host/abi_decode.cc -- C ABI of the batched one-best WFST search (include/eesen_b200.h, "decoding").
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/eesen_b200.h"
#include "context.h"

struct eesen_b200_graph {
  eb::DecodeGraph g;
  std::vector<void *> bufs;
  int device = 0;
};

extern "C" {

int eesen_b200_graph_create(eesen_b200_ctx *ctx, int num_states, int num_arcs, int start, const int *row, const int *eps,
                            const int *ilabel, const int *olabel, const float *weight, const int *nextstate,
                            const float *final_cost, eesen_b200_graph **out) {
  if (!ctx || !out || num_states < 1 || num_arcs < 0 || start < 0 || start >= num_states || !row || !eps || !final_cost ||
      (num_arcs > 0 && (!ilabel || !olabel || !weight || !nextstate)))
    return EESEN_B200_EINVAL;
  *out = nullptr;
  // structure checks: monotone CSR, emitting arcs first (ilabel >= 1), then epsilon-input arcs (ilabel == 0)
  if (row[0] != 0 || row[num_states] != num_arcs) return ctx->fail(EESEN_B200_ESHAPE, "graph: row[] is not a CSR over the arcs");
  std::vector<int> from((size_t)num_arcs);
  for (int s = 0; s < num_states; s++) {
    if (row[s] > row[s + 1] || eps[s] < row[s] || eps[s] > row[s + 1])
      return ctx->fail(EESEN_B200_ESHAPE, "graph: bad arc range at state " + std::to_string(s));
    for (int a = row[s]; a < row[s + 1]; a++) {
      from[a] = s;
      if ((a < eps[s]) != (ilabel[a] != 0) || ilabel[a] < 0 || nextstate[a] < 0 || nextstate[a] >= num_states)
        return ctx->fail(EESEN_B200_ESHAPE, "graph: arc " + std::to_string(a) + " is out of order or out of range");
    }
  }
  eesen_b200_graph *g = new eesen_b200_graph();
  g->device = ctx->device;
  auto up = [&](const void *src, size_t bytes) -> void * {
    void *d = nullptr;
    if (cudaMalloc(&d, bytes ? bytes : 4) != cudaSuccess) return nullptr;
    g->bufs.push_back(d);
    if (bytes && cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    return d;
  };
  g->g.num_states = num_states; g->g.num_arcs = num_arcs; g->g.start = start;
  g->g.row = (const int *)up(row, sizeof(int) * (num_states + 1));
  g->g.eps = (const int *)up(eps, sizeof(int) * num_states);
  g->g.ilabel = (const int *)up(ilabel, sizeof(int) * num_arcs);
  g->g.olabel = (const int *)up(olabel, sizeof(int) * num_arcs);
  g->g.nextstate = (const int *)up(nextstate, sizeof(int) * num_arcs);
  g->g.arc_from = (const int *)up(from.data(), sizeof(int) * num_arcs);
  g->g.weight = (const float *)up(weight, sizeof(float) * num_arcs);
  g->g.final_cost = (const float *)up(final_cost, sizeof(float) * num_states);
  if (!g->g.row || !g->g.eps || !g->g.ilabel || !g->g.olabel || !g->g.nextstate || !g->g.arc_from || !g->g.weight ||
      !g->g.final_cost) {
    eesen_b200_graph_free(g);
    return ctx->fail((int)cudaErrorMemoryAllocation, "graph: device allocation / upload failed");
  }
  *out = g;
  return 0;
}

void eesen_b200_graph_free(eesen_b200_graph *g) {
  if (!g) return;
  cudaSetDevice(g->device);
  for (void *p : g->bufs) cudaFree(p);
  delete g;
}

int eesen_b200_decode_best_path(eesen_b200_ctx *ctx, const eesen_b200_graph *g, int S, int T, const int *frames,
                                const float *d_loglikes, int ld, int K, float acoustic_scale, float beam,
                                int max_active, int min_active, int frame_cap, int tok_cap, int *out_labels,
                                int max_out, int *out_len, float *out_cost, double *stats) {
  if (!ctx || !g || S < 1 || T < 0 || !frames || !d_loglikes || ld < K || K < 1 || !out_labels || !out_len || !out_cost ||
      max_out < 1 || frame_cap < 1 || tok_cap < frame_cap || !(beam > 0.f))
    return EESEN_B200_EINVAL;
  if (max_active < 1 || min_active < 0 || min_active > max_active)
    return ctx->fail(EESEN_B200_EINVAL, "decode: need 1 <= max_active, 0 <= min_active <= max_active (no limit: 2147483647 / 0)");
  for (int s = 0; s < S; s++)
    if (frames[s] < 0 || frames[s] > T) return ctx->fail(EESEN_B200_EINVAL, "decode: frames[] out of range");
  ctx->join_side();
  const int wl_cap = 4 * frame_cap;
  void *ws = nullptr;
  // outputs live behind the workspace
  const size_t need = eb::decode_workspace_bytes(S, g->g.num_states, frame_cap, wl_cap, tok_cap);
  const size_t out_bytes = (size_t)S * max_out * 4 + (size_t)S * 8 + 1024;
  int rc = ctx->reserve(ctx->decode_ws, need + out_bytes, &ws);
  if (rc) return rc;
  char *ob = (char *)ws + ((need + 255) & ~(size_t)255);
  int *d_labels = (int *)ob;
  int *d_len = (int *)(ob + (((size_t)S * max_out * 4 + 255) & ~(size_t)255));
  float *d_cost = (float *)(d_len + S);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, ctx->stream);
  int err_bits = 0;
  long rounds = 0;
  cudaError_t ce = eb::decode_best_path(ctx->stream, ctx->num_sms, g->g, S, T, frames, d_loglikes, ld, acoustic_scale, beam, max_active,
                                        min_active, ws, frame_cap, wl_cap, tok_cap, d_labels, max_out, d_len, d_cost, &err_bits, &rounds);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches += ((max_active == 2147483647 && min_active == 0) ? 4 : 5) * (long)(T + 1) + rounds;
  if ((rc = ctx->check(ce, "decode_best_path"))) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
  cudaError_t c2 = cudaMemcpyAsync(out_labels, d_labels, (size_t)S * max_out * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (c2 == cudaSuccess) c2 = cudaMemcpyAsync(out_len, d_len, (size_t)S * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (c2 == cudaSuccess) c2 = cudaMemcpyAsync(out_cost, d_cost, (size_t)S * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (c2 == cudaSuccess) c2 = cudaStreamSynchronize(ctx->stream);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if ((rc = ctx->check(c2, "decode: result copy"))) return rc;
  if (stats) { stats[0] = (double)rounds; stats[1] = (double)ms; }
  if (err_bits & 1) return ctx->fail(EESEN_B200_ESHAPE, "decode: more than frame_cap tokens in one frame (raise frame_cap or lower the beam)");
  if (err_bits & 2) return ctx->fail(EESEN_B200_ESHAPE, "decode: epsilon-closure work list overflow (raise frame_cap)");
  if (err_bits & 4) return ctx->fail(EESEN_B200_ESHAPE, "decode: more than tok_cap tokens in one utterance (raise tok_cap)");
  return 0;
}

}  // extern "C"
