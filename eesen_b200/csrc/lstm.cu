// eesen_b200/csrc/lstm.cu -- persistent recurrent kernels for the utterance-parallel BiLSTM.
//
// Replaces the per-timestep launch chains of the reference
//   forward : BiLstmParallel::PropagateFncVanillaPass{Forward,Backward}  bilstm-parallel-layer.h:112-149,166-205
//             (per step: cublasSgemm + 3x add_mat_diag_vec + 3x sigmoid + 2x tanh + 3x add_mat_dot_mat,
//              + per-row memsets for the padding of the backward direction, :201-204)
//   backward: BackpropagateFncVanillaPass{Forward,Backward}               :450-499,541-590
//             and the bias / peephole reductions after the loop           :507-510,598-601
// with ONE cooperative launch per layer and pass that runs all T steps of BOTH directions.
//
// Decomposition (B200: 148 SMs, 227 KB smem/CTA): CTA(dir, group, slice) owns a slice of
// 8*NCT cells for a group of 8*NUT utterances of one direction and keeps its 4*8*NCT rows of
// the recurrent matrix Wm resident in shared memory, pre-arranged in mma fragment order, for
// the whole sequence.  Utterance groups are independent recurrences; the CTAs of one
// (dir, group) exchange h/m through global memory (L2-resident) and synchronise with a
// monotonic release/acquire step counter -- no grid-wide barrier, no host round trip.
//   forward : every CTA needs the full m_{t-1} of its group  (all-gather, [8*NUT x C] per step)
//   backward: every CTA multiplies ITS d(gates) by ITS Wm rows into a partial d_m for all C
//             cells; the owner of a cell slice sums the `slices` partials in fixed order
//             (deterministic reduce-scatter, same [8*NUT x C] volume per step as forward).
// Tensor path: mma.sync m16n8k8 TF32 (gate rows x 8 utterances x k), "3xTF32" split by default
// so the recurrence is fp32-faithful.  Peepholes, activations, cell update, padding mask and
// (backward) bias/peephole gradient accumulation are fused into the step.
#include "common.cuh"
#include "kernels.h"

#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

namespace eb {

#ifdef EB_LSTM_TIMING
// debug build only (make TIMING=1): per-phase clock64 deltas of thread 0 of CTA (0,0,0)
__device__ long long g_lstm_timing[2][16];
#define EB_T0() long long tk_ = clock64()
#define EB_TICK(kernel, i)                                                          \
  do {                                                                              \
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {        \
      long long n_ = clock64();                                                     \
      g_lstm_timing[kernel][i] += n_ - tk_;                                         \
      tk_ = n_;                                                                     \
    }                                                                               \
  } while (0)
#else
#define EB_T0()
#define EB_TICK(kernel, i)
#endif

namespace {

__device__ __forceinline__ uint32_t cvt_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}

// Inter-CTA exchange between the CTAs of one (direction, utterance group), "LL" style: every
// exchanged float travels with the step tag in one 8-byte relaxed store (common.cuh:st_tagged);
// consumers spin on the data words themselves until the tag of the step they need shows up.
// No release fence on the producer, no flag round trip, no poll barrier on the consumer.  Two
// parity buffers suffice: a CTA can only publish step n+2 after it has consumed every word of
// step n+1, which in turn were produced by CTAs that had consumed all of step n.
// Buffers are zeroed before each launch (tags of a launch are step+1 >= 1).

// One k-step of the (gate-rows x utterances) product for one 16-row fragment A (fp32 in smem,
// fragment order) against B (2 regs): acc += A*B.
template <int PREC>
__device__ __forceinline__ void mma_step(float (&acc)[4], float (&accc)[4], const float4 &A, float b0, float b1) {
  if (PREC == 0) {
    uint32_t ah[4], al[4], bh[2], bl[2];
    split_tf32(A.x, ah[0], al[0]); split_tf32(A.y, ah[1], al[1]);
    split_tf32(A.z, ah[2], al[2]); split_tf32(A.w, ah[3], al[3]);
    split_tf32(b0, bh[0], bl[0]); split_tf32(b1, bh[1], bl[1]);
    mma_tf32(accc, al, bh);
    mma_tf32(accc, ah, bl);
    mma_tf32(acc, ah, bh);
  } else {
    uint32_t a[4] = {cvt_tf32(A.x), cvt_tf32(A.y), cvt_tf32(A.z), cvt_tf32(A.w)};
    uint32_t b[2] = {cvt_tf32(b0), cvt_tf32(b1)};
    mma_tf32(acc, a, b);
  }
}

// ------------------------------------------------------------------------------------ forward
template <int NUT, int NCT, int KSPLIT, int PREC, int DROP>
__global__ void __launch_bounds__(32 * NUT * NCT * KSPLIT, 1)
lstm_fwd_kernel(LstmFwdArgs a, int groups, unsigned expected_per_step) {
  constexpr int NTASK = NUT * NCT;
  constexpr int NTHREADS = 32 * NTASK * KSPLIT;
  extern __shared__ __align__(16) float smem[];
  const int C = a.C, S = a.S, T = a.T;
  const int KS = C / 8;
  const int SST = C + 4;  // staging row stride
  float *Wsm = smem;                                   // [NCT][2][KS][32][4]
  float *stg = Wsm + (size_t)NCT * KS * 256;           // [8*NUT][C+4]
  float *scr = stg + (size_t)8 * NUT * SST;            // [NTASK][KSPLIT-1][32][8]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int task = warp % NTASK, ksid = warp / NTASK;
  const int ut = task / NCT, ct = task % NCT;
  const int slice = blockIdx.x, group = blockIdx.y, dir = blockIdx.z;
  const LstmDirParams P = a.p[dir];
  uint2 *xbuf = reinterpret_cast<uint2 *>(a.xbuf);   // [2 parity][2 dir][groups][8*NUT][C] tagged words

  // ---- resident recurrent weights, fragment order
  for (int idx = tid; idx < NCT * KS * 256; idx += NTHREADS) {
    int e = idx & 3, ln = (idx >> 2) & 31, rest = idx >> 7;
    int ks = rest % KS, mt = (rest / KS) & 1, c_ = rest / (2 * KS);
    int gg = ln >> 2, tt = ln & 3;
    int gate = mt * 2 + (e & 1);
    int k = ks * 8 + tt + ((e & 2) ? 4 : 0);
    int cellw = (slice * NCT + c_) * 8 + gg;
    Wsm[idx] = cellw < C ? P.wm[((size_t)gate * C + cellw) * P.ldwm + k] : 0.f;
  }

  const int cell = (slice * NCT + ct) * 8 + g;  // this lane's cell (finalising warps)
  const bool cell_ok = cell < C;
  const int kper = (KS + KSPLIT - 1) / KSPLIT;
  const int kb = ksid * kper, ke = min(KS, kb + kper);

  // finalising-warp state
  float cprev[2] = {0.f, 0.f};
  float pre[4][2];
  float rmk[2] = {1.f, 1.f};   // recurrent dropout mask of this lane's (cell, utterance) pairs at the current step
  float ppi = 0.f, ppf = 0.f, ppo = 0.f;
  int lenu[2] = {0, 0};
  int uidx[2];
  const int s0 = a.s_begin, s1 = a.s_begin + a.s_count;   // utterance window of this launch (rows still index all S)
  uidx[0] = s0 + (group * NUT + ut) * 8 + 2 * tg;
  uidx[1] = uidx[0] + 1;
  const bool fin = (ksid == 0);
  if (fin && cell_ok) {
    ppi = P.pi[cell]; ppf = P.pf[cell]; ppo = P.po[cell];
  }
  if (fin) {
#pragma unroll
    for (int e = 0; e < 2; e++) lenu[e] = (a.len && uidx[e] < s1) ? a.len[uidx[e]] : 0;   // NULL: uni-directional (no masking)
  }
  auto load_pre = [&](int t) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      bool ok = cell_ok && uidx[e] < s1;
      const float *row = a.G + ((size_t)t * S + (ok ? uidx[e] : 0)) * a.ldg + (size_t)dir * 4 * C + (ok ? cell : 0);
#pragma unroll
      for (int q = 0; q < 4; q++) pre[q][e] = ok ? __ldcs(row + (size_t)q * C) : 0.f;
      if (DROP != 0)
        rmk[e] = ok ? __ldg(a.rmask + (size_t)(a.rmask_per_step ? (size_t)t * S + uidx[e] : (size_t)uidx[e]) * a.ldr +
                            (size_t)dir * C + cell)
                    : 0.f;
    }
  };
  if (fin) load_pre(dir == 0 ? 0 : T - 1);
  __syncthreads();
  // The first k-slice of this warp's weights lives in registers for the whole sequence (8 registers): 10 % less
  // shared-memory traffic per step and tensor work that can start before the first LDS of a step returns
  // (tests/micro/lstm_inner.cu "1 of 10 slices resident": 3879 -> 3662 cycles for the product phase).
  const bool warp_has_cells = (slice * NCT + ct) * 8 < C;
  float4 RA0 = make_float4(0.f, 0.f, 0.f, 0.f), RA1 = RA0;
  if (warp_has_cells && kb < ke) {
    RA0 = reinterpret_cast<const float4 *>(Wsm)[((size_t)(ct * 2 + 0) * KS + kb) * 32 + lane];
    RA1 = reinterpret_cast<const float4 *>(Wsm)[((size_t)(ct * 2 + 1) * KS + kb) * 32 + lane];
  }

  for (int step = 0; step < T; step++) {
    const int t = dir == 0 ? step : T - 1 - step;
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[i][c] = 0.f;

    EB_T0();
    if (step > 0) {
      const int tp = dir == 0 ? t - 1 : t + 1;
      EB_TICK(0, 0);
      EB_TICK(0, 1);
      // stage m_{t-1} of the whole group (L2 -> smem): spin on the tagged words of step-1
      (void)tp;
      const int c2n = C / 2;
      const uint4 *xr = reinterpret_cast<const uint4 *>(xbuf + ((size_t)((step - 1) & 1) * 2 * groups + (size_t)dir * groups + group) * (8 * NUT) * C);
      const unsigned want = (unsigned)step;   // tag of the data produced at step-1
      for (int v = tid; v < 8 * NUT * c2n; v += NTHREADS) {
        int u = v / c2n, c2 = v % c2n;
        int s = s0 + group * NUT * 8 + u;
        float2 val = make_float2(0.f, 0.f);
        if (s < s1) {
          uint4 q;
          do { q = ld_tagged2(xr + v); } while (q.y != want || q.w != want);
          val = make_float2(__uint_as_float(q.x), __uint_as_float(q.z));
        }
        *reinterpret_cast<float2 *>(stg + (size_t)u * SST + c2 * 2) = val;
      }
      EB_TICK(0, 2);
      __syncthreads();
      EB_TICK(0, 3);
      if ((slice * NCT + ct) * 8 < C) {
        float accc[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int c = 0; c < 4; c++) accc[i][c] = 0.f;
        const float4 *W0 = reinterpret_cast<const float4 *>(Wsm) + ((size_t)(ct * 2 + 0) * KS) * 32 + lane;
        const float4 *W1 = reinterpret_cast<const float4 *>(Wsm) + ((size_t)(ct * 2 + 1) * KS) * 32 + lane;
        const float *brow = stg + (size_t)(ut * 8 + g) * SST + tg;
        if (kb < ke) {
          const float b0 = brow[kb * 8], b1 = brow[kb * 8 + 4];
          mma_step<PREC>(acc[0], accc[0], RA0, b0, b1);
          mma_step<PREC>(acc[1], accc[1], RA1, b0, b1);
        }
        // (an explicit register double buffer for the next k-slice was measured slower here -- 11.3 vs
        //  10.7 ms per C2 step -- while the same change helps the backward kernel; tests/micro/lstm_inner.cu
        //  has the isolated loop: shared-memory side 2148 clk, tensor side 2944 clk, this schedule 3879)
#pragma unroll 2
        for (int ks = kb + 1; ks < ke; ks++) {
          float4 A0 = W0[(size_t)ks * 32], A1 = W1[(size_t)ks * 32];
          float b0 = brow[ks * 8], b1 = brow[ks * 8 + 4];
          mma_step<PREC>(acc[0], accc[0], A0, b0, b1);
          mma_step<PREC>(acc[1], accc[1], A1, b0, b1);
        }
        if (PREC == 0) {
#pragma unroll
          for (int i = 0; i < 2; i++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][c] += accc[i][c];
        }
      }
      EB_TICK(0, 4);
      if (KSPLIT > 1) {
        if (!fin) {
          float4 *dst = reinterpret_cast<float4 *>(scr) + ((size_t)(task * (KSPLIT - 1) + ksid - 1) * 32 + lane) * 2;
          dst[0] = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
          dst[1] = make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]);
        }
        __syncthreads();
        EB_TICK(0, 5);
        if (fin) {
#pragma unroll
          for (int k = 0; k < KSPLIT - 1; k++) {
            const float4 *src = reinterpret_cast<const float4 *>(scr) + ((size_t)(task * (KSPLIT - 1) + k) * 32 + lane) * 2;
            float4 v0 = src[0], v1 = src[1];
            acc[0][0] += v0.x; acc[0][1] += v0.y; acc[0][2] += v0.z; acc[0][3] += v0.w;
            acc[1][0] += v1.x; acc[1][1] += v1.y; acc[1][2] += v1.z; acc[1][3] += v1.w;
          }
        }
      }
      EB_TICK(0, 9);
    }

    if (fin) {
      // acc[0] = {g(u0), g(u1), i(u0), i(u1)}, acc[1] = {f(u0), f(u1), o(u0), o(u1)} for cell `cell`
      // The two utterances of a lane are evaluated branch-free side by side (invalid slots carry zeros):
      // their exp/rcp chains -- four dependent SFU round trips each -- interleave instead of running
      // one after the other.
      float sg[2], si[2], sf[2], so[2], sc[2], sm[2];
      bool valid[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        valid[e] = cell_ok && uidx[e] < s1;
        float yg = pre[0][e] + acc[0][e];
        float yi = pre[1][e] + acc[0][2 + e] + cprev[e] * ppi;   // :127
        float yf = pre[2][e] + acc[1][e] + cprev[e] * ppf;       // :129
        float gi = sigmoidf_(yi), gf = sigmoidf_(yf), gg = tanhf_(yg);   // :131-133
        float c;
        if (DROP == 0) {
          c = gg * gi + cprev[e] * gf;                            // :136-137
        } else if (DROP == 1) {
          c = rmk[e] * (gg * gi) + cprev[e] * gf;                 // :267-274 no-mem-loss: only the new content is dropped
        } else {
          c = rmk[e] * (gg * gi + cprev[e] * gf);                 // :276-277 RNNdrop: the whole cell
        }
        float h = tanhf_(c);                                      // :140
        float go = sigmoidf_(pre[3][e] + acc[1][2 + e] + c * ppo);  // :143-144
        float m = h * go;                                         // :147
        const bool keep = valid[e] && !(dir == 1 && t >= lenu[e]);   // :201-204 (backward cells only)
        sg[e] = keep ? gg : 0.f; si[e] = keep ? gi : 0.f; sf[e] = keep ? gf : 0.f; so[e] = keep ? go : 0.f;
        sc[e] = keep ? c : 0.f; sm[e] = keep ? m : 0.f;
        cprev[e] = sc[e];
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int u = uidx[e];
        if (valid[e]) {
          const float m = sm[e];
          // only m is on the inter-CTA critical path: publish it first (tagged word, no fence) ...
          if (step + 1 < T)
            st_tagged(xbuf + (((size_t)(step & 1) * 2 * groups + (size_t)dir * groups + group) * (8 * NUT) + (ut * 8 + 2 * tg + e)) * C + cell,
                      m, (unsigned)step + 1u);
          __stcs(a.out + ((size_t)t * S + u) * a.ldo + (size_t)dir * C + cell, m);
        }
      }
      EB_TICK(0, 6);
      EB_TICK(0, 7);
      // ... the saved state for the backward pass is written behind the release
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int u = uidx[e];
        if (valid[e]) {
          float *grow = a.G + ((size_t)t * S + u) * a.ldg + (size_t)dir * 4 * C + cell;
          __stcs(grow, sg[e]); __stcs(grow + (size_t)C, si[e]);
          __stcs(grow + (size_t)2 * C, sf[e]); __stcs(grow + (size_t)3 * C, so[e]);
          __stcs(a.cell + ((size_t)t * S + u) * a.ldc + (size_t)dir * C + cell, sc[e]);
        }
      }
      if (step + 1 < T) load_pre(dir == 0 ? t + 1 : t - 1);
      EB_TICK(0, 8);
    }
  }
}

// ------------------------------------------------------------------------------------ backward
template <int NUT, int NCT, int NWARPS, int PREC, int DROP>
__global__ void __launch_bounds__(32 * NWARPS, 1)
lstm_bwd_kernel(LstmBwdArgs a, int groups, int slices, unsigned expected_per_step) {
  constexpr int NTHREADS = 32 * NWARPS;
  constexpr int KSB = 4 * NCT;            // k-steps over this CTA's own 32*NCT gate rows
  constexpr int DST = 32 * NCT + 4;       // Dsm row stride
  constexpr int ITEMS = 64 * NUT * NCT;   // (utterance, cell) pairs owned by the CTA
  static_assert(ITEMS <= NTHREADS, "one thread per (utt, cell) item");
  extern __shared__ __align__(16) float smem[];
  const int C = a.C, S = a.S, T = a.T;
  const int MT = (C + 15) / 16;
  const int CP = MT * 16;
  float *Wsm = smem;                              // [MT][KSB][32][4]
  float *Dsm = Wsm + (size_t)MT * KSB * 128;      // [8*NUT][DST]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int slice = blockIdx.x, group = blockIdx.y, dir = blockIdx.z;
  const LstmDirParams P = a.p[dir];
  uint2 *pbuf = reinterpret_cast<uint2 *>(a.pbuf);   // tagged partial d_m words

  for (int idx = tid; idx < MT * KSB * 128; idx += NTHREADS) {
    int e = idx & 3, ln = (idx >> 2) & 31, rest = idx >> 7;
    int ks = rest % KSB, mt = rest / KSB;
    int gg = ln >> 2, tt = ln & 3;
    int rl = ks * 8 + tt + ((e & 2) ? 4 : 0);     // local gate row (K index)
    int j = mt * 16 + gg + ((e & 1) ? 8 : 0);     // output cell (M index)
    int gate = rl / (8 * NCT), rem = rl % (8 * NCT);
    int cellw = slice * NCT * 8 + rem;
    Wsm[idx] = (cellw < C && j < C) ? P.wm[((size_t)gate * C + cellw) * P.ldwm + j] : 0.f;
  }

  // per-item state
  const bool is_item = tid < ITEMS;
  const int cl = tid % (8 * NCT), ul = tid / (8 * NCT);
  const int cell = slice * NCT * 8 + cl;
  const int u = a.s_begin + group * NUT * 8 + ul;
  const bool ok = is_item && cell < C && u < a.s_begin + a.s_count;
  float ppi = 0.f, ppf = 0.f, ppo = 0.f;
  if (ok) { ppi = P.pi[cell]; ppf = P.pf[cell]; ppo = P.po[cell]; }
  float dc_next = 0.f, f_next = 0.f, di_next = 0.f, df_next = 0.f;
  float dcm_next = 0.f, vr = 1.f;   // recurrent dropout: masked cell gradient of the next step, this step's mask
  float sb[4] = {0.f, 0.f, 0.f, 0.f}, spi = 0.f, spf = 0.f, spo = 0.f;
  float vg = 0.f, vi = 0.f, vf = 0.f, vo = 0.f, vc = 0.f, vcp = 0.f, vd = 0.f;
  const int tstep = dir == 0 ? -1 : 1;  // time order of the BPTT sweep
  auto prefetch = [&](int t) {
    if (!ok) return;
    const float *grow = a.G + ((size_t)t * S + u) * a.ldg + (size_t)dir * 4 * C + cell;
    vg = __ldcs(grow); vi = __ldcs(grow + (size_t)C); vf = __ldcs(grow + (size_t)2 * C); vo = __ldcs(grow + (size_t)3 * C);
    vc = __ldcs(a.cell + ((size_t)t * S + u) * a.ldc + (size_t)dir * C + cell);
    int tp = t + tstep;  // the step the forward pass took before t, i.e. c_{prev}
    vcp = (tp >= 0 && tp < T) ? __ldcs(a.cell + ((size_t)tp * S + u) * a.ldc + (size_t)dir * C + cell) : 0.f;
    vd = __ldcs(a.dout + ((size_t)t * S + u) * a.ldd + (size_t)dir * C + cell);
    if (DROP != 0)
      vr = __ldg(a.rmask + (size_t)(a.rmask_per_step ? (size_t)t * S + u : (size_t)u) * a.ldr + (size_t)dir * C + cell);
  };
  prefetch(dir == 0 ? T - 1 : 0);
  __syncthreads();

  const size_t pstride_slice = (size_t)8 * NUT * CP;
  for (int step = 0; step < T; step++) {
    const int t = dir == 0 ? T - 1 - step : step;
    EB_T0();
    EB_TICK(1, 0);
    EB_TICK(1, 1);
    if (is_item) {
      float dm = vd;
      // (hoisting the d_m-independent factors -- tanh(c) and the gate derivatives -- above the spin-wait was
      //  measured slower, 11.08 vs 10.78 ms per C2 step: more values live across the wait, a spill)
      if (step > 0) {
        if (ok) {
          const uint2 *pb = pbuf + ((((size_t)((step - 1) & 1) * 2 + dir) * groups + group) * slices) * pstride_slice +
                            (size_t)ul * CP + cell;
          const unsigned want = (unsigned)step;
          float s_ = 0.f;
          for (int sl0 = 0; sl0 < slices; sl0 += 8) {
            uint2 q[8];
            bool all;
            do {   // issue the (up to) 8 loads together, retry until every tag has arrived
              all = true;
#pragma unroll
              for (int i = 0; i < 8; i++) {
                if (sl0 + i < slices) {
                  q[i] = ld_tagged(pb + (size_t)(sl0 + i) * pstride_slice);
                  all = all && (q[i].y == want);
                }
              }
            } while (!all);
#pragma unroll
            for (int i = 0; i < 8; i++)
              if (sl0 + i < slices) s_ += __uint_as_float(q[i].x);      // fixed order: deterministic
          }
          dm += s_;                                                       // :470 / :561
        }
      }
      float dg = 0.f, di = 0.f, df = 0.f, dO = 0.f;
      if (ok) {
        float h = tanhf_(vc);
        float dh = dm * vo * (1.f - h * h);                               // :473-474
        dO = dm * h * vo * (1.f - vo);                                    // :477-478
        float dc, dcm;
        if (DROP == 0) {
          dc = dh + dc_next * f_next + di_next * ppi + df_next * ppf + dO * ppo;   // :481-485
          dcm = dc;
        } else {
          // :697-711: RNNdrop carries the MASKED cell gradient through the forget gate, no-mem-loss the plain one
          dc = dh + di_next * ppi + df_next * ppf + dO * ppo + (DROP == 2 ? dcm_next : dc_next) * f_next;
          dcm = dc * vr;
        }
        df = (DROP == 2 ? dcm : dc) * vcp * vf * (1.f - vf);              // :488-489 / :715-719
        di = dcm * vg * vi * (1.f - vi);                                  // :492-493 / :723
        dg = dcm * vi * (1.f - vg * vg);                                  // :496-497 / :727
        dc_next = dc; dcm_next = dcm; f_next = vf; di_next = di; df_next = df;
        float *drow = a.DG + ((size_t)t * S + u) * a.lddg + (size_t)dir * 4 * C + cell;
        drow[0] = dg; drow[(size_t)C] = di; drow[(size_t)2 * C] = df; drow[(size_t)3 * C] = dO;
        sb[0] += dg; sb[1] += di; sb[2] += df; sb[3] += dO;               // :507 / :598
        spi += di * vcp; spf += df * vcp; spo += dO * vc;                 // :508-510 / :599-601
      }
      float *drow_s = Dsm + (size_t)ul * DST + cl;
      drow_s[0] = dg; drow_s[8 * NCT] = di; drow_s[16 * NCT] = df; drow_s[24 * NCT] = dO;
      if (step + 1 < T) prefetch(t + tstep);
    }
    if (step + 1 == T) break;  // the last step's recurrent contribution is never consumed
    EB_TICK(1, 2);
    __syncthreads();
    EB_TICK(1, 3);
    // partial d_m for ALL cells from this CTA's gate rows: P[j, utt] = sum_r Wm[r, j] * D[utt, r]
    for (int wt = warp; wt < MT * NUT; wt += NWARPS) {
      const int mt = wt % MT, ut = wt / MT;
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, accc[4] = {0.f, 0.f, 0.f, 0.f};
      const float4 *W = reinterpret_cast<const float4 *>(Wsm) + ((size_t)mt * KSB) * 32 + lane;
      const float *brow = Dsm + (size_t)(ut * 8 + g) * DST + tg;
      // operands two k-slices ahead of the tensor instructions (see the forward kernel)
      float4 An[2] = {W[0], W[32]};
      float bn[2][2] = {{brow[0], brow[4]}, {brow[8], brow[12]}};
#pragma unroll 4
      for (int ks = 0; ks < KSB; ks++) {
        const float4 A = An[ks & 1];
        const float b0 = bn[ks & 1][0], b1 = bn[ks & 1][1];
        if (ks + 2 < KSB) {
          An[ks & 1] = W[(size_t)(ks + 2) * 32];
          bn[ks & 1][0] = brow[(ks + 2) * 8]; bn[ks & 1][1] = brow[(ks + 2) * 8 + 4];
        }
        mma_step<PREC>(acc, accc, A, b0, b1);
      }
      if (PREC == 0) {
#pragma unroll
        for (int c = 0; c < 4; c++) acc[c] += accc[c];
      }
      uint2 *pb = pbuf + (((((size_t)(step & 1) * 2 + dir) * groups + group) * slices + slice) * 8 * NUT + ut * 8) * CP;
      const int j = mt * 16 + g;
      const unsigned tagw = (unsigned)step + 1u;
      st_tagged(pb + (size_t)(2 * tg) * CP + j, acc[0], tagw);
      st_tagged(pb + (size_t)(2 * tg + 1) * CP + j, acc[1], tagw);
      st_tagged(pb + (size_t)(2 * tg) * CP + j + 8, acc[2], tagw);
      st_tagged(pb + (size_t)(2 * tg + 1) * CP + j + 8, acc[3], tagw);
    }
    EB_TICK(1, 4);
    EB_TICK(1, 5);
  }

  // ---- bias / peephole gradient partial sums of this (dir, group): reduce over the CTA's utterances
  __syncthreads();
  float *red = smem;  // reuse (weights no longer needed)
  if (is_item) {
    float v[7] = {sb[0], sb[1], sb[2], sb[3], spi, spf, spo};
#pragma unroll
    for (int q = 0; q < 7; q++) red[(size_t)q * ITEMS + tid] = v[q];
  }
  __syncthreads();
  for (int idx = tid; idx < 7 * 8 * NCT; idx += NTHREADS) {
    int q = idx / (8 * NCT), c_ = idx % (8 * NCT);
    int cellw = slice * NCT * 8 + c_;
    if (cellw < C) {
      float s_ = 0.f;
      for (int uu = 0; uu < 8 * NUT; uu++) s_ += red[(size_t)q * ITEMS + uu * 8 * NCT + c_];
      a.gsum[(((size_t)dir * groups + group) * 7 + q) * C + cellw] = s_;
    }
  }
}

__global__ void gsum_reduce_kernel(int C, int groups, int nchunks, size_t chunk_stride, const float *gsum, float *db,
                                   float *dpi, float *dpf, float *dpo, int dir) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 7 * C) return;
  int q = i / C, c = i % C;
  float s = 0.f;
  for (int ch = 0; ch < nchunks; ch++)
    for (int gr = 0; gr < groups; gr++) s += gsum[ch * chunk_stride + (((size_t)dir * groups + gr) * 7 + q) * C + c];
  if (q < 4) db[(size_t)q * C + c] = s;
  else if (q == 4) dpi[c] = s;
  else if (q == 5) dpf[c] = s;
  else dpo[c] = s;
}

struct Cfg { int nut, nct, ksplit; };
constexpr Cfg kCfgs[] = {{1, 1, 4}, {1, 2, 4}, {1, 4, 4}, {1, 5, 4}, {2, 2, 4}, {2, 4, 2}, {4, 2, 2}};

size_t fwd_smem(const Cfg &c, int C) {
  return sizeof(float) * ((size_t)c.nct * (C / 8) * 256 + (size_t)8 * c.nut * (C + 4) +
                          (size_t)c.nut * c.nct * (c.ksplit - 1) * 256);
}
size_t bwd_smem(const Cfg &c, int C) {
  size_t mt = (C + 15) / 16;
  size_t w = mt * 4 * c.nct * 128 + (size_t)8 * c.nut * (32 * c.nct + 4);
  size_t red = (size_t)7 * 64 * c.nut * c.nct;
  return sizeof(float) * (w > red ? w : red);
}

template <int NUT, int NCT, int KSPLIT>
cudaError_t launch_fwd(cudaStream_t st, const LstmPlan &pl, const LstmFwdArgs &a) {
  unsigned expected = (unsigned)(pl.slices * NCT * NUT);  // finalising warps per (dir, group) and step
  dim3 grid(pl.slices, pl.groups, pl.ndir), block(pl.threads);
  int groups = pl.groups;
  LstmFwdArgs args = a;
  void *kargs[] = {&args, &groups, &expected};
  const void *fn;
  if (a.drop == 0)
    fn = a.precision == 0 ? (const void *)lstm_fwd_kernel<NUT, NCT, KSPLIT, 0, 0> : (const void *)lstm_fwd_kernel<NUT, NCT, KSPLIT, 1, 0>;
  else if (a.drop == 1)
    fn = a.precision == 0 ? (const void *)lstm_fwd_kernel<NUT, NCT, KSPLIT, 0, 1> : (const void *)lstm_fwd_kernel<NUT, NCT, KSPLIT, 1, 1>;
  else
    fn = a.precision == 0 ? (const void *)lstm_fwd_kernel<NUT, NCT, KSPLIT, 0, 2> : (const void *)lstm_fwd_kernel<NUT, NCT, KSPLIT, 1, 2>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd);
  if (e != cudaSuccess) return e;
  return cudaLaunchCooperativeKernel(fn, grid, block, kargs, pl.smem_fwd, st);
}

template <int NUT, int NCT, int KSPLIT>
cudaError_t launch_bwd(cudaStream_t st, const LstmPlan &pl, const LstmBwdArgs &a) {
  constexpr int NWARPS = NUT * NCT * KSPLIT;
  unsigned expected = (unsigned)(NWARPS * pl.slices);
  dim3 grid(pl.slices, pl.groups, pl.ndir), block(pl.threads);
  int groups = pl.groups, slices = pl.slices;
  LstmBwdArgs args = a;
  void *kargs[] = {&args, &groups, &slices, &expected};
  const void *fn;
  if (a.drop == 0)
    fn = a.precision == 0 ? (const void *)lstm_bwd_kernel<NUT, NCT, NWARPS, 0, 0> : (const void *)lstm_bwd_kernel<NUT, NCT, NWARPS, 1, 0>;
  else if (a.drop == 1)
    fn = a.precision == 0 ? (const void *)lstm_bwd_kernel<NUT, NCT, NWARPS, 0, 1> : (const void *)lstm_bwd_kernel<NUT, NCT, NWARPS, 1, 1>;
  else
    fn = a.precision == 0 ? (const void *)lstm_bwd_kernel<NUT, NCT, NWARPS, 0, 2> : (const void *)lstm_bwd_kernel<NUT, NCT, NWARPS, 1, 2>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bwd);
  if (e != cudaSuccess) return e;
  return cudaLaunchCooperativeKernel(fn, grid, block, kargs, pl.smem_bwd, st);
}

}  // namespace

int lstm_debug_timing(long long *out32, int reset) {
#ifdef EB_LSTM_TIMING
  if (out32) {
    cudaMemcpyFromSymbol(out32, g_lstm_timing, sizeof(long long) * 32);
    long long tc[32];
    if (lstm_tc_debug_timing(tc, reset))
      for (int i = 0; i < 32; i++) out32[i] += tc[i];   // only one engine runs per build-and-measure session
  }
  if (reset) {
    long long z[32] = {0};
    cudaMemcpyToSymbol(g_lstm_timing, z, sizeof(z));
  }
  return 1;
#else
  (void)out32; (void)reset;
  return 0;
#endif
}

// Engine per pass, read at every plan (not cached: the tests switch engines inside one process).
//   EESEN_B200_LSTM_ENGINE unset or "tc" : tcgen05 kernels (lstm_tc.cu) where the shape allows (cells % 64 == 0,
//                                  <= 384), else the warp-level kernels below.  Measured on C2 (ms per step, 4 layers):
//                                  forward 6.1 vs 10.2, backward 10.7 vs 10.8 -- and the tcgen05 grids occupy 80 SMs
//                                  instead of 128, which leaves room for the weight-gradient GEMMs of the side stream
//                                  (whole step 23.3 ms with both passes on tcgen05, 25.5 ms with the warp-level backward)
//   "legacy"     : both passes here;  "tcfwd" : tcgen05 forward + warp-level backward
static int engine_for_pass(int pass) {
  const char *e = getenv("EESEN_B200_LSTM_ENGINE");
  if (e && strcmp(e, "legacy") == 0) return 0;
  if (e && strcmp(e, "tcfwd") == 0) return pass == 0 ? 1 : 0;
  return 1;
}

LstmPlan lstm_plan(int S, int C, int num_sms, size_t max_smem, int ndir, int pass) {
  if (engine_for_pass(pass) == 1) {
    LstmPlan tc = lstm_tc_plan(S, C, num_sms, max_smem, ndir);
    if (tc.valid) return tc;
  }
  LstmPlan best;
  best.valid = 0;
  best.engine = 0;
  best.cluster = 0;
  best.ndir = ndir;
  long best_work = -1;
  if (C % 8 != 0 || C <= 0 || S <= 0 || ndir < 1 || ndir > 2) return best;
  for (const Cfg &c : kCfgs) {
    int groups = (S + 8 * c.nut - 1) / (8 * c.nut);
    int slices = (C + 8 * c.nct - 1) / (8 * c.nct);
    long ctas = (long)ndir * groups * slices;
    size_t sf = fwd_smem(c, C), sb = bwd_smem(c, C);
    if (ctas > num_sms || sf > max_smem || sb > max_smem) continue;
    long work = (long)c.nut * c.nct * 16 + (4 - c.ksplit);  // per-CTA MMA work, then prefer deeper K split
    if (best_work < 0 || work < best_work) {
      best_work = work;
      best.nut = c.nut; best.nct = c.nct; best.ksplit = c.ksplit;
      best.groups = groups; best.slices = slices;
      best.threads = 32 * c.nut * c.nct * c.ksplit;
      best.smem_fwd = sf; best.smem_bwd = sb;
      size_t cp = (size_t)((C + 15) / 16) * 16;
      best.pbuf_floats = (size_t)2 * 2 * 2 * groups * slices * 8 * c.nut * cp;   // 8-byte tagged words
      best.xbuf_bytes = (size_t)2 * 2 * groups * 8 * c.nut * C * 8;
      best.gsum_floats = (size_t)2 * groups * 7 * C;
      best.valid = 1;
    }
  }
  return best;
}

#define EB_DISPATCH(FN, ...)                                                             \
  do {                                                                                   \
    const int key = plan.nut * 100 + plan.nct * 10 + plan.ksplit;                        \
    switch (key) {                                                                       \
      case 114: return FN<1, 1, 4>(__VA_ARGS__);                                         \
      case 124: return FN<1, 2, 4>(__VA_ARGS__);                                         \
      case 144: return FN<1, 4, 4>(__VA_ARGS__);                                         \
      case 154: return FN<1, 5, 4>(__VA_ARGS__);                                         \
      case 224: return FN<2, 2, 4>(__VA_ARGS__);                                         \
      case 242: return FN<2, 4, 2>(__VA_ARGS__);                                         \
      case 422: return FN<4, 2, 2>(__VA_ARGS__);                                         \
      default: return cudaErrorInvalidConfiguration;                                     \
    }                                                                                    \
  } while (0)

cudaError_t lstm_forward(cudaStream_t st, const LstmPlan &plan, const LstmFwdArgs &a) {
  if (!plan.valid) return cudaErrorInvalidConfiguration;
  if (plan.engine == 1) return lstm_tc_forward(st, plan, a);
  if (a.drop < 0 || a.drop > 2 || (a.drop != 0 && !a.rmask)) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(a.xbuf, 0, plan.xbuf_bytes, st);
  if (e != cudaSuccess) return e;
  EB_DISPATCH(launch_fwd, st, plan, a);
}

cudaError_t lstm_backward(cudaStream_t st, const LstmPlan &plan, const LstmBwdArgs &a) {
  if (!plan.valid) return cudaErrorInvalidConfiguration;
  if (plan.engine == 1) return lstm_tc_backward(st, plan, a);
  if (a.drop < 0 || a.drop > 2 || (a.drop != 0 && !a.rmask)) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(a.pbuf, 0, plan.pbuf_floats * sizeof(float), st);
  if (e != cudaSuccess) return e;
  EB_DISPATCH(launch_bwd, st, plan, a);
}

cudaError_t lstm_reduce_gsum(cudaStream_t st, const LstmPlan &plan, int C, const float *gsum, int nchunks, float *db,
                             float *dpi, float *dpf, float *dpo, int dir) {
  int n = 7 * C;
  gsum_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(C, plan.groups, nchunks, plan.gsum_floats, gsum, db, dpi, dpf, dpo, dir);
  return cudaGetLastError();
}

}  // namespace eb
