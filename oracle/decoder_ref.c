/* oracle/decoder_ref.c -- TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's WFST token-passing search
 * (BASELINE config 5), one-best output.
 *
 * PARITY UNPINNED against the reference binary: `latgen-faster` needs OpenFst 1.4.1 (un-vendored, tools/Makefile:8)
 * which is neither in /root/reference nor in this image, and the reference ships no decoder test or golden output
 * (src/decoder has no *-test.cc).  This restatement follows the reference's own sources line by line and is pinned
 * instead on hand-built graphs against exhaustive path enumeration (tests/test_decoder.py).
 *
 * Follows (all paths relative to /root/reference/src):
 *   decoder/lattice-faster-decoder.cc:53-71    InitDecoding: start token at cost 0, then ProcessNonemitting
 *   decoder/lattice-faster-decoder.cc:77-97    Decode: per frame ProcessEmitting, ProcessNonemitting
 *   decoder/lattice-faster-decoder.cc:594-658  GetCutoff: best + beam, max_active / min_active through nth_element
 *   decoder/lattice-faster-decoder.cc:660-752  ProcessEmitting: cost_offset = -best, tot = cur + (offset - loglike) + graph,
 *                                              tokens with tot_cost <= cur_cutoff are expanded
 *   decoder/lattice-faster-decoder.cc:756-816  ProcessNonemitting: cutoff = best + beam, epsilon arcs relaxed to a fix point,
 *                                              new cost must be < cutoff
 *   decoder/lattice-faster-decoder.cc:146-167  FindOrAddToken: a state keeps the smaller cost (strict >: first wins a tie)
 *   decoder/lattice-faster-decoder.cc:531-577  ComputeFinalCosts: final weights if any surviving token is final
 *   decoder/decodable-matrix.h:54-56           LogLikelihood(frame, tid) = scale * likes(frame, tid - 1)
 *   decoderbin/latgen-faster.cc:96-126         one utterance at a time; words = non-zero olabels of the best path
 * Not restated: the order-dependent "online" tightening of next_cutoff inside ProcessEmitting (:684-700,:727-729).
 * It only ever drops tokens that are more than `beam` above the best token of the NEW frame; those are not
 * expanded at the next frame anyway (tot_cost <= cur_cutoff, :716) nor by the epsilon closure (cutoff = best + beam,
 * :775), so the one-best path and its cost do not depend on it.  Lattice generation (forward links, PruneActiveTokens,
 * determinisation) is out of this slice.
 *
 * Graph format (no OpenFst): CSR over states; the arcs of a state are stored emitting arcs first, then epsilon arcs:
 *   row[s] .. eps[s]   emitting arcs (ilabel >= 1, the 1-based CTC token id)
 *   eps[s] .. row[s+1] epsilon-input arcs (ilabel == 0)
 *   arc: ilabel, olabel (0 = none), weight (graph cost, -log), nextstate;  final[s]: final cost, +inf = not final. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int state;
  float cost;
  int prev;    /* index of the predecessor token in the token store, -1 for the start token */
  int olabel;  /* output label of the arc that created (last improved) the token */
} Tok;

static int cmp_float(const void *a, const void *b) {
  float x = *(const float *)a, y = *(const float *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* k-th smallest (0-based) of v[0..n): what std::nth_element leaves at position k */
static float kth_smallest(const float *v, int n, int k) {
  float *t = (float *)malloc(sizeof(float) * (size_t)n);
  memcpy(t, v, sizeof(float) * (size_t)n);
  qsort(t, (size_t)n, sizeof(float), cmp_float);
  float r = t[k];
  free(t);
  return r;
}

/* Returns the number of output labels written (<= max_out), -1 if no token survived.
 * loglikes: [T x K] row-major (what net-output-extract writes); *total_cost = path cost incl. the acoustic offsets
 * added back and the final weight (the cost ShortestPath would report on the raw lattice). */
int oracle_decode_best_path(int num_states, int start, const int *row, const int *eps, const int *ilabel,
                            const int *olabel, const float *weight, const int *nextstate, const float *final_cost,
                            int T, int K, const float *loglikes, float acoustic_scale, float beam, int max_active,
                            int min_active, int *out_labels, int max_out, float *total_cost, int *frames_decoded,
                            long *arcs_expanded) {
  (void)K;
  size_t cap = 1 << 16, ntok = 0;
  Tok *toks = (Tok *)malloc(sizeof(Tok) * cap);
  int *slot = (int *)malloc(sizeof(int) * (size_t)num_states);   /* state -> token of the frame being built */
  int *cur = NULL, ncur = 0;                                      /* token indices of the current frame */
  int *queue = (int *)malloc(sizeof(int) * 16), qcap = 16;
  double offset_sum = 0.0;
  long expanded = 0;
  for (int s = 0; s < num_states; s++) slot[s] = -1;

#define PUSH_TOK(S, C, P, O)                                              \
  do {                                                                    \
    if (ntok == cap) { cap *= 2; toks = (Tok *)realloc(toks, sizeof(Tok) * cap); } \
    toks[ntok].state = (S); toks[ntok].cost = (C); toks[ntok].prev = (P); toks[ntok].olabel = (O); \
    ntok++;                                                               \
  } while (0)

  /* frame list under construction */
  int *nxt = (int *)malloc(sizeof(int) * 16), nnxt = 0, nxtcap = 16;
#define ADD_NEXT(TI)                                                      \
  do {                                                                    \
    if (nnxt == nxtcap) { nxtcap *= 2; nxt = (int *)realloc(nxt, sizeof(int) * (size_t)nxtcap); } \
    nxt[nnxt++] = (TI);                                                   \
  } while (0)

  /* epsilon closure over the tokens in nxt[] (ProcessNonemitting :756-816) */
#define CLOSURE()                                                         \
  do {                                                                    \
    float best_cost = INFINITY;                                           \
    int qn = 0;                                                           \
    for (int i = 0; i < nnxt; i++) {                                      \
      if (qn == qcap) { qcap *= 2; queue = (int *)realloc(queue, sizeof(int) * (size_t)qcap); } \
      queue[qn++] = toks[nxt[i]].state;                                   \
      if (toks[nxt[i]].cost < best_cost) best_cost = toks[nxt[i]].cost;   \
    }                                                                     \
    const float cutoff = best_cost + beam;                                \
    while (qn > 0) {                                                      \
      const int st = queue[--qn];                                         \
      const int ti = slot[st];                                            \
      const float cur_cost = toks[ti].cost;                               \
      if (cur_cost > cutoff) continue;                                    \
      for (int a = eps[st]; a < row[st + 1]; a++) {                       \
        const float tot = cur_cost + weight[a];                           \
        expanded++;                                                       \
        if (tot < cutoff) {                                               \
          const int ns = nextstate[a];                                    \
          int changed = 0;                                                \
          if (slot[ns] < 0) {                                             \
            PUSH_TOK(ns, tot, ti, olabel[a]);                             \
            slot[ns] = (int)ntok - 1;                                     \
            ADD_NEXT((int)ntok - 1);                                      \
            changed = 1;                                                  \
          } else if (toks[slot[ns]].cost > tot) {                         \
            toks[slot[ns]].cost = tot; toks[slot[ns]].prev = ti; toks[slot[ns]].olabel = olabel[a]; \
            changed = 1;                                                  \
          }                                                               \
          if (changed) {                                                  \
            if (qn == qcap) { qcap *= 2; queue = (int *)realloc(queue, sizeof(int) * (size_t)qcap); } \
            queue[qn++] = ns;                                             \
          }                                                               \
        }                                                                 \
      }                                                                   \
    }                                                                     \
  } while (0)

  /* InitDecoding */
  PUSH_TOK(start, 0.0f, -1, 0);
  slot[start] = 0;
  ADD_NEXT(0);
  CLOSURE();

  int t = 0;
  for (; t < T; t++) {
    /* the tokens built so far become the current frame; the state map is cleared (toks_.Clear()) */
    free(cur);
    cur = nxt; ncur = nnxt;
    nxt = (int *)malloc(sizeof(int) * 16); nnxt = 0; nxtcap = 16;
    for (int i = 0; i < ncur; i++) slot[toks[cur[i]].state] = -1;
    if (ncur == 0) break;
    /* GetCutoff :594-658 */
    float best = INFINITY;
    for (int i = 0; i < ncur; i++) if (toks[cur[i]].cost < best) best = toks[cur[i]].cost;
    float cur_cutoff = best + beam;
    if (!(max_active == 2147483647 && min_active == 0)) {
      float *tmp = (float *)malloc(sizeof(float) * (size_t)ncur);
      for (int i = 0; i < ncur; i++) tmp[i] = toks[cur[i]].cost;
      float max_c = INFINITY, min_c = INFINITY;
      if (ncur > max_active) max_c = kth_smallest(tmp, ncur, max_active);                 /* :626-631 */
      /* :632-642: nth_element within the max_active smallest values -- the same value as over all of them */
      if (ncur > min_active) min_c = min_active == 0 ? best : kth_smallest(tmp, ncur, min_active);
      free(tmp);
      if (max_c < cur_cutoff) cur_cutoff = max_c;
      else if (min_c > cur_cutoff) cur_cutoff = min_c;
    }
    const float cost_offset = -best;                                  /* :689 */
    offset_sum += (double)cost_offset;
    const float *ll = loglikes + (size_t)t * K;
    for (int i = 0; i < ncur; i++) {
      const int ti = cur[i];
      if (!(toks[ti].cost <= cur_cutoff)) continue;                   /* :716 */
      const int st = toks[ti].state;
      for (int a = row[st]; a < eps[st]; a++) {
        const float ac_cost = cost_offset - acoustic_scale * ll[ilabel[a] - 1];   /* :722-723, decodable-matrix.h:54-56 */
        const float graph_cost = weight[a], cur_cost = toks[ti].cost;
        const float tot = cur_cost + ac_cost + graph_cost;            /* :724-726, left to right */
        expanded++;
        const int ns = nextstate[a];
        if (slot[ns] < 0) {
          PUSH_TOK(ns, tot, ti, olabel[a]);
          slot[ns] = (int)ntok - 1;
          ADD_NEXT((int)ntok - 1);
        } else if (toks[slot[ns]].cost > tot) {
          toks[slot[ns]].cost = tot; toks[slot[ns]].prev = ti; toks[slot[ns]].olabel = olabel[a];
        }
      }
    }
    CLOSURE();
  }
  if (frames_decoded) *frames_decoded = t;
  if (arcs_expanded) *arcs_expanded = expanded;

  /* ComputeFinalCosts :531-577 + best path */
  int best_tok = -1;
  float best_final = INFINITY, best_plain = INFINITY;
  int best_plain_tok = -1;
  for (int i = 0; i < nnxt; i++) {
    const Tok *k = &toks[nxt[i]];
    if (k->cost < best_plain) { best_plain = k->cost; best_plain_tok = nxt[i]; }
    const float f = final_cost[k->state];
    if (f != INFINITY && k->cost + f < best_final) { best_final = k->cost + f; best_tok = nxt[i]; }
  }
  float path_cost;
  if (best_tok >= 0) path_cost = best_final;
  else { best_tok = best_plain_tok; path_cost = best_plain; }
  int n = -1;
  if (best_tok >= 0) {
    /* olabels from the end back to the start, then reversed */
    int cnt = 0;
    for (int k = best_tok; k >= 0; k = toks[k].prev) if (toks[k].olabel != 0) cnt++;
    n = cnt;
    int w = cnt;
    for (int k = best_tok; k >= 0; k = toks[k].prev)
      if (toks[k].olabel != 0) { w--; if (w < max_out) out_labels[w] = toks[k].olabel; }
    if (n > max_out) n = max_out;
    if (total_cost) *total_cost = (float)((double)path_cost - offset_sum);   /* the offsets were ADDED to every path */
  }
  free(toks); free(slot); free(cur); free(nxt); free(queue);
  return n;
}
