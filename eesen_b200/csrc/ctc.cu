// eesen_b200/csrc/ctc.cu -- row softmax (+log, +argmax) and the fused CTC forward-backward.
//
// Replaces, for one packed minibatch:
//   Softmax::PropagateFnc -> _softmax_reduce            softmax-layer.h:44-47, cuda-kernels.cu:744-808
//   CuMatrixBase::FindRowMaxId (ErrorRateMSeq)          cuda-matrix.cc:1038-1095, ctc-loss.cc:238-239
//   Ctc::EvalParallel                                    ctc-loss.cc:101-168:
//       ApplyLog, 2T launches of _compute_ctc_{alpha,beta}_multiple_sequence (cuda-kernels.cu:1369-1408,
//       1484-1544) each with label-matrix uploads, 2S scalar D2H reads for pzx (ctc-loss.cc:146-153),
//       _compute_ctc_error_multiple_sequence (cuda-kernels.cu:1605-1627) and the softmax back-prop
//       (MulElements/AddColSumMat/MulRowsVec/AddMat, ctc-loss.cc:160-168)
// with two launches.  ctc_kernel: one CTA per utterance; warp 0 runs the whole alpha recursion and
// warp 1 the whole beta recursion CONCURRENTLY, each holding the 2|l|+1 lattice positions in
// registers (R per lane) and exchanging only the chunk-boundary values with warp shuffles -- no
// block barrier per time step; then all warps turn alpha+beta into per-class occupancies and write
// diff = y*Z - occ (= y - occ up to rounding; rows t >= T_s are zeroed as in the reference).
// Log-domain arithmetic keeps the reference's log(0) = -1e30 sentinel (ctc-utils.h:36): in fp32
// -1e30 + x == -1e30 for any emission x, so the sentinel propagates without branches.
#include "common.cuh"
#include "kernels.h"

namespace eb {

namespace {

// softmax over K columns, one warp per row
__global__ void softmax_rows_kernel(int N, int K, const float *__restrict__ logits, int ld,
                                    float *__restrict__ probs, int ldp, int *__restrict__ argmax) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float *x = logits + (size_t)row * ld;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    float v = x[k];
    if (v > mx) { mx = v; mi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  float sum = 0.f;
  for (int k = lane; k < K; k += 32) sum += expf(x[k] - mx);
  sum = warp_sum(sum);
  float *p = probs + (size_t)row * ldp;
  for (int k = lane; k < ldp; k += 32) p[k] = k < K ? expf(x[k] - mx) / sum : 0.f;  // cuda-kernels.cu:778-806
  if (argmax && lane == 0) argmax[row] = mi;
}

// the output side of net-output-extract (src/netbin/net-output-extract.cc:100-110) in one pass:
//   y = log(y) if apply_log (CuMatrixBase::ApplyLog, cuda-kernels.cu:221-227)
//   y -= prior_scale * log_prior[col]   (ClassPrior::SubtractOnLogpost -> AddVecToRows, class-prior.cc:78-90)
__global__ void loglik_rows_kernel(long n, int K, float *__restrict__ y, int ld, int apply_log,
                                   const float *__restrict__ log_prior, float prior_scale) {
  long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long r = i / K;
    int k = (int)(i - r * K);
    float v = y[r * ld + k];
    if (apply_log) v = logf(v);
    if (log_prior) v += -prior_scale * log_prior[k];
    y[r * ld + k] = v;
  }
}

// first index of the row maximum (CPU branch of FindRowMaxId, cuda-matrix.cc:1077-1093)
__global__ void row_argmax_kernel(int N, int K, const float *__restrict__ x, int ld, int *__restrict__ argmax) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  float mx = -INFINITY;
  int mi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    float v = x[(size_t)row * ld + k];
    if (v > mx) { mx = v; mi = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
  }
  if (lane == 0) argmax[row] = mi;
}

// The lattice is kept in the log2 domain (alpha2 = alpha / ln 2): every log-sum-exp is then bare SFU
// work -- ex2.approx / lg2.approx, no range reduction, no denormal fix-ups -- and only pzx is
// converted back to natural log.  Sentinel log(0) = -1e30 (ApplyLog ctc-loss.cc:133 + ctc-utils.h:36)
// survives unchanged: -1e30 + x == -1e30 in fp32 and ex2(-1e30) == 0.
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float logprob(float y) { return fmaxf(lg2_approx(y), kLogZero); }   // log2 y

// log-sum-exp on the SFU: the sum of exponentials lies in [1, 3], where __logf is accurate to
// 2^-21 absolute -- far below one fp32 ulp of the log-domain values themselves (|alpha| ~ 10^2..10^3)
__device__ __forceinline__ float lse2(float a, float b) {   // log2(2^a + 2^b)
  float m = fmaxf(a, b);
  return m + lg2_approx(ex2_approx(a - m) + ex2_approx(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  return m + lg2_approx(ex2_approx(a - m) + ex2_approx(b - m) + ex2_approx(c - m));
}

template <int R>
__global__ void __launch_bounds__(256, 1)
ctc_kernel(int T, int S, int K, int max_lab, const int *__restrict__ len, const int *__restrict__ labels,
           const int *__restrict__ lab_len, const float *__restrict__ probs, int ldp,
           float *__restrict__ pzx_out, float *__restrict__ diff, int ldd, float *__restrict__ ws) {
  constexpr int PF = R <= 4 ? 8 : (R <= 8 ? 4 : 2);   // emission prefetch distance (time steps)
  constexpr int LP = 32 * R;           // padded lattice width
  // per-warp class occupancies, accumulated as 24.40 fixed point: integer addition is associative, so
  // the shared-memory atomics below give the same bits whatever order the lanes of a pass are served in
  extern __shared__ unsigned long long occ_sm[];    // [nwarps][K]
  __shared__ float pzx_sm;
  const int s = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int Ts = min(len[s], T);
  const int nl = lab_len[s];
  const int L = 2 * nl + 1;
  const int *lab_s = labels + (size_t)s * max_lab;
  auto lab_at = [&](int i) {   // class of label i; ids outside [0, K) (a corrupt targets archive) read as blank
    const int c = lab_s[i];    // instead of indexing out of bounds -- the host rejects such utterances (net.cc)
    return (unsigned)c < (unsigned)K ? c : 0;
  };
  float *alpha = ws + (size_t)s * T * LP;
  float *beta = ws + (size_t)S * T * LP + (size_t)s * T * LP;

  if (warp < 2 && Ts > 0) {
    // lattice positions of this lane: j = lane*R + r
    int cls[R];
    bool skip[R];  // alpha: j-2 transition allowed ; beta: j+2 transition allowed
#pragma unroll
    for (int r = 0; r < R; r++) {
      int j = lane * R + r;
      cls[r] = j >= L ? -1 : ((j & 1) ? lab_at(j >> 1) : 0);
      if (warp == 0) skip[r] = (j & 1) && j >= 3 && j < L && lab_s[j >> 1] != lab_s[(j >> 1) - 1];
      else skip[r] = (j & 1) && j + 2 < L && lab_s[j >> 1] != lab_s[(j >> 1) + 1];
    }
    float cur[R], e[PF][R];
    const int dt = warp == 0 ? 1 : -1;
    const int t0 = warp == 0 ? 0 : Ts - 1;
    auto emis = [&](float (&dst)[R], int t) {
      const float *row = probs + ((size_t)t * S + s) * ldp;
#pragma unroll
      for (int r = 0; r < R; r++) dst[r] = cls[r] >= 0 ? __ldg(row + cls[r]) : 0.f;   // RAW posteriors: the log is taken
      // where the value is used, PF steps later -- taking it here makes the in-order warp wait for the load it just issued
    };
#pragma unroll
    for (int i = 0; i < PF; i++)
      if (i < Ts) emis(e[i], t0 + dt * i);
    float *dstbuf = warp == 0 ? alpha : beta;
    for (int n0 = 0; n0 < Ts; n0 += PF) {
#pragma unroll
      for (int i = 0; i < PF; i++) {
        const int n = n0 + i;
        if (n < Ts) {
          const int t = t0 + dt * n;
          if (n == 0) {
#pragma unroll
            for (int r = 0; r < R; r++) {
              int j = lane * R + r;
              bool init = warp == 0 ? (j < 2) : (j > L - 3);   // cuda-kernels.cu:1392-1394 / 1526-1528
              cur[r] = (cls[r] >= 0 && init) ? logprob(e[i][r]) : kLogZero;
            }
          } else {
            float nb1, nb2;  // neighbour chunk values: alpha <- previous lane's last two, beta <- next lane's first two
            if (warp == 0) {
              nb1 = __shfl_up_sync(0xffffffffu, cur[R - 1], 1);
              nb2 = R >= 2 ? __shfl_up_sync(0xffffffffu, cur[R >= 2 ? R - 2 : 0], 1)
                           : __shfl_up_sync(0xffffffffu, cur[0], 2);
              if (lane == 0) { nb1 = kLogZero; nb2 = kLogZero; }
              if (R == 1 && lane == 1) nb2 = kLogZero;
            } else {
              nb1 = __shfl_down_sync(0xffffffffu, cur[0], 1);
              nb2 = R >= 2 ? __shfl_down_sync(0xffffffffu, cur[R >= 2 ? 1 : 0], 1)
                           : __shfl_down_sync(0xffffffffu, cur[0], 2);
              if (lane == 31) { nb1 = kLogZero; nb2 = kLogZero; }
              if (R == 1 && lane == 30) nb2 = kLogZero;
            }
            float nxt[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
              float x0 = cur[r], x1, x2;
              if (warp == 0) {
                x1 = r >= 1 ? cur[r >= 1 ? r - 1 : 0] : nb1;
                x2 = r >= 2 ? cur[r >= 2 ? r - 2 : 0] : (r == 1 ? nb1 : nb2);
                if (R == 1) x2 = nb2;
              } else {
                x1 = r + 1 < R ? cur[r + 1 < R ? r + 1 : 0] : nb1;
                x2 = r + 2 < R ? cur[r + 2 < R ? r + 2 : 0] : (r + 1 < R ? nb1 : nb2);
                if (R == 1) x2 = nb2;
              }
              // branch-free: a forbidden j-2 / j+2 transition enters as log 0 (2^(-1e30 - m) == 0 exactly, so the sum is
              // the two-term one bit for bit); without the per-lane branch the R chains of a lane interleave
              float v = lse3(x0, x1, skip[r] ? x2 : kLogZero);
              v = logprob(e[i][r]) + v;              // AddAB(prob, LogAPlusB(..)) :1397-1406 / 1531-1541
              nxt[r] = (cls[r] >= 0 && v > kLogZero) ? v : kLogZero;
            }
#pragma unroll
            for (int r = 0; r < R; r++) cur[r] = nxt[r];
          }
          float *dst = dstbuf + (size_t)t * LP + lane * R;
          if (R % 4 == 0) {   // 16-byte row stores (the workspace rows are 128-float multiples)
#pragma unroll
            for (int r = 0; r < R; r += 4)
              *reinterpret_cast<float4 *>(dst + r) = make_float4(cur[r], cur[r + 1 < R ? r + 1 : 0], cur[r + 2 < R ? r + 2 : 0], cur[r + 3 < R ? r + 3 : 0]);
          } else {
#pragma unroll
            for (int r = 0; r < R; r++) dst[r] = cur[r];
          }
          if (n + PF < Ts) emis(e[i], t0 + dt * (n + PF));
        }
      }
    }
    if (warp == 0) {
      // pzx = logadd(alpha(T_s-1, L-1), alpha(T_s-1, L-2))   ctc-loss.cc:146-153
      float a1 = kLogZero, a2 = kLogZero;
#pragma unroll
      for (int r = 0; r < R; r++) {
        int j = lane * R + r;
        if (j == L - 1) a1 = cur[r];
        if (j == L - 2) a2 = cur[r];
      }
      a1 = warp_max(a1);
      a2 = warp_max(a2);
      if (lane == 0) {
        float p = lse2(a1, a2);
        pzx_sm = p;                                   // log2 domain for the occupancies below
        pzx_out[s] = p < 0.5f * kLogZero ? kLogZero : p * 0.6931471805599453f;   // natural log at the boundary
      }
    }
  } else if (warp == 0 && lane == 0) {
    pzx_sm = kLogZero;
    pzx_out[s] = kLogZero;
  }
  __syncthreads();

  // ---- occupancies -> gradient wrt the pre-softmax activations
  const float pzx = pzx_sm;
  unsigned long long *occ = occ_sm + (size_t)warp * K;
  constexpr float kFix = 1099511627776.f, kUnfix = 1.f / 1099511627776.f;   // 2^40
  for (int t = warp; t < T; t += nwarps) {
    float *drow = diff + ((size_t)t * S + s) * ldd;
    if (t >= Ts) {
      for (int k = lane; k < K; k += 32) drow[k] = 0.f;   // padded rows: ctc_err_ stays 0 (:1615)
      continue;
    }
    for (int k = lane; k < K; k += 32) occ[k] = 0ull;
    __syncwarp();
    const float *arow = alpha + (size_t)t * LP, *brow = beta + (size_t)t * LP;
    const float *yrow = probs + ((size_t)t * S + s) * ldp;
    float blank = 0.f;
    const float lb = logprob(yrow[0]);
    // all loads of the row first (alpha, beta, the labels' posteriors: up to R of each per lane), then the arithmetic --
    // one round trip to L2/HBM per row instead of one per 32 lattice positions
    float abv[R], yv[R];
    int cv[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int j = lane + 32 * r;
      abv[r] = 0.f; yv[r] = 1.f; cv[r] = 0;
      if (j < L) {
        abv[r] = arow[j] + brow[j];
        if (j & 1) { cv[r] = lab_at(j >> 1); yv[r] = yrow[cv[r]]; }
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int j = lane + 32 * r;
      if (j < L) {
        if (j & 1) {
          float gam = ex2_approx(abv[r] - pzx - logprob(yv[r]));    // exp(log(a*b) - pzx - 2 log y) * y  (:1624 then MulElements)
          atomicAdd(&occ[cv[r]], (unsigned long long)(fminf(gam, 1048576.f) * kFix));   // each term <= 1 up to rounding
        } else {
          blank += ex2_approx(abv[r] - pzx - lb);
        }
      }
    }
    blank = warp_sum(blank);
    __syncwarp();
    float z = 0.f;
    for (int k = lane; k < K; k += 32) z += (float)occ[k] * kUnfix + (k == 0 ? blank : 0.f);
    z = warp_sum(z);                              // = -rowsum(ctc_err .* y)  (ctc-loss.cc:161-162)
    for (int k = lane; k < K; k += 32)
      drow[k] = yrow[k] * z - ((float)occ[k] * kUnfix + (k == 0 ? blank : 0.f));   // :164-168
    __syncwarp();
  }
}

}  // namespace

cudaError_t softmax_rows(cudaStream_t st, int N, int K, const float *logits, int ld, float *probs, int ldp,
                         int *argmax) {
  if (N <= 0) return cudaSuccess;
  int rows_per_block = 8;
  softmax_rows_kernel<<<(N + rows_per_block - 1) / rows_per_block, rows_per_block * 32, 0, st>>>(
      N, K, logits, ld, probs, ldp, argmax);
  return cudaGetLastError();
}

cudaError_t loglik_rows(cudaStream_t st, int num_sms, int N, int K, float *y, int ld, int apply_log,
                        const float *log_prior, float prior_scale) {
  long n = (long)N * K;
  if (n <= 0 || (!apply_log && !log_prior)) return cudaSuccess;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * num_sms) blocks = 8 * num_sms;
  loglik_rows_kernel<<<blocks, 256, 0, st>>>(n, K, y, ld, apply_log, log_prior, prior_scale);
  return cudaGetLastError();
}

cudaError_t row_argmax(cudaStream_t st, int N, int K, const float *x, int ld, int *argmax) {
  if (N <= 0) return cudaSuccess;
  row_argmax_kernel<<<(N + 7) / 8, 256, 0, st>>>(N, K, x, ld, argmax);
  return cudaGetLastError();
}

static int ctc_R(int max_lab) {
  int L = 2 * max_lab + 1;
  int R = 1;
  while (32 * R < L) R *= 2;
  return R;
}

size_t ctc_workspace_floats(int T, int S, int max_lab) {
  return (size_t)2 * S * T * 32 * ctc_R(max_lab);
}

cudaError_t ctc_eval(cudaStream_t st, int T, int S, int K, int max_lab, const int *len, const int *labels,
                     const int *lab_len, const float *probs, int ldp, float *pzx, float *diff, int ldd,
                     float *ws) {
  if (S <= 0 || T <= 0) return cudaSuccess;
  int R = ctc_R(max_lab);
  // one occupancy row of K 8-byte cells per warp: as many warps (2..8) as the opt-in shared memory holds,
  // so that character / BPE output layers (K in the thousands) launch too -- the reference takes any K
  int dev = 0, max_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int nwarps = 8;
  while (nwarps > 2 && sizeof(unsigned long long) * nwarps * (size_t)K > (size_t)max_optin - 1024) nwarps--;
  size_t smem = sizeof(unsigned long long) * nwarps * (size_t)K;
  if (smem > (size_t)max_optin - 1024) return cudaErrorInvalidValue;   // K > ~14 000 classes
#define EB_CTC(RR)                                                                                        \
  case RR:                                                                                                \
    if (smem > 48 * 1024) {                                                                               \
      cudaError_t ae = cudaFuncSetAttribute(ctc_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (ae != cudaSuccess) return ae;                                                                   \
    }                                                                                                     \
    ctc_kernel<RR><<<S, 32 * nwarps, smem, st>>>(T, S, K, max_lab, len, labels, lab_len, probs, ldp, pzx, diff, \
                                                 ldd, ws);                                                \
    break;
  switch (R) {
    EB_CTC(1) EB_CTC(2) EB_CTC(4) EB_CTC(8) EB_CTC(16) EB_CTC(32)
    default: return cudaErrorInvalidValue;  // more than 511 labels per utterance
  }
#undef EB_CTC
  return cudaGetLastError();
}

}  // namespace eb
