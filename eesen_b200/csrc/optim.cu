// eesen_b200/csrc/optim.cu -- fused momentum + clip + SGD over the contiguous parameter arena,
// and the column-sum used for the affine bias gradient.
//
// Replaces, per parameter tensor and step (12 tensors per BiLSTM layer + 2 for the affine layer):
//   AddMatMat/AddRowSumMat/AddDiagMatMat(..., beta = momentum)   bilstm-parallel-layer.h:504-510,595-601
//   ApplyFloor(-max_grad) + ApplyCeiling(max_grad) in place       bilstm-layer.h:848-862, affine-trans-layer.h:186-189
//   AddMat(-lr*coef, corr)                                        bilstm-layer.h:865-883, affine-trans-layer.h:191-195
// (36+ launches per layer in the reference) with ONE launch over the whole arena.  The raw
// gradient `grad` is kept separate from the momentum-carrying `corr` so that the data-parallel
// all-reduce acts on the raw sum (SURVEY.md section 8e ordering constraint):
//   corr = grad + momentum * corr ; corr = clamp(corr, +-max_grad) ; w -= lr * corr
#include "common.cuh"
#include "kernels.h"

namespace eb {

namespace {

// element update shared by the vector and the scalar path.
//   MODE 0  SGD      : w -= lr * corr                                   (bilstm-layer.h:865-883)
//   MODE 1  Adagrad  : accu += corr^2                                   (trainable-layer.h:65-78)
//   MODE 2  RMSProp  : accu  = rho*accu + one_minus_rho*corr^2          (trainable-layer.h:80-96)
//           then       w -= lr * corr / sqrt(accu + eps)                (trainable-layer.h:98-114, bilstm-layer.h:885-955)
// corr = clamp(grad + momentum*corr) in every mode (the adaptive rules act on the momentum buffer too).
template <int MODE>
__device__ __forceinline__ void opt_elem(float &w, float &c, float &a, float g, float momentum, float lr, float max_grad,
                                         float eps, float rho, float omr) {
  c = g + momentum * c;
  if (max_grad > 0.f) c = fminf(fmaxf(c, -max_grad), max_grad);
  if (MODE == 0) {
    w -= lr * c;
  } else {
    a = MODE == 1 ? a + c * c : rho * a + omr * (c * c);
    w -= lr * (1.f / sqrtf(a + eps)) * c;
  }
}

template <int MODE>
__global__ void opt_kernel(float *__restrict__ w, float *__restrict__ corr, float *__restrict__ accu,
                           const float *__restrict__ grad, float momentum, float eps, float rho, float omr,
                           const SgdSegment *__restrict__ segs, int nseg, long total) {
  long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (segs[mid].offset <= i) lo = mid; else hi = mid - 1;
    }
    const SgdSegment sg = segs[lo];
    float a = MODE ? accu[i] : 0.f, c = corr[i], ww = w[i];
    opt_elem<MODE>(ww, c, a, grad[i], momentum, sg.lr, sg.max_grad, eps, rho, omr);
    corr[i] = c;
    w[i] = ww;
    if (MODE) accu[i] = a;
  }
}

__global__ void sgd_kernel(float *__restrict__ w, float *__restrict__ corr, const float *__restrict__ grad,
                           float momentum, const SgdSegment *__restrict__ segs, int nseg, long total) {
  long stride = (long)gridDim.x * blockDim.x * 4;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += stride) {
    // locate the segment of element i (segments are sorted, few dozen entries)
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (segs[mid].offset <= i) lo = mid; else hi = mid - 1;
    }
    SgdSegment sg = segs[lo];
    long seg_end = sg.offset + sg.count;
    if (i + 4 <= seg_end && i + 4 <= total) {
      float4 g4 = *reinterpret_cast<const float4 *>(grad + i);
      float4 c4 = *reinterpret_cast<float4 *>(corr + i);
      float4 w4 = *reinterpret_cast<float4 *>(w + i);
      float c[4] = {g4.x + momentum * c4.x, g4.y + momentum * c4.y, g4.z + momentum * c4.z, g4.w + momentum * c4.w};
      if (sg.max_grad > 0.f) {
#pragma unroll
        for (int q = 0; q < 4; q++) c[q] = fminf(fmaxf(c[q], -sg.max_grad), sg.max_grad);
      }
      *reinterpret_cast<float4 *>(corr + i) = make_float4(c[0], c[1], c[2], c[3]);
      *reinterpret_cast<float4 *>(w + i) =
          make_float4(w4.x - sg.lr * c[0], w4.y - sg.lr * c[1], w4.z - sg.lr * c[2], w4.w - sg.lr * c[3]);
    } else {
      for (long k = i; k < i + 4 && k < total; k++) {
        while (lo + 1 < nseg && segs[lo + 1].offset <= k) lo++;
        SgdSegment s2 = segs[lo];
        float c = grad[k] + momentum * corr[k];
        if (s2.max_grad > 0.f) c = fminf(fmaxf(c, -s2.max_grad), s2.max_grad);
        corr[k] = c;
        w[k] -= s2.lr * c;
      }
    }
  }
}

// out = a (.) b elementwise over [N x cols] (CuMatrixBase::MulElements: the forward-dropout mask on the layer output
// and on out_diff, bilstm-parallel-layer.h:414, :893)
__global__ void mul_elements_kernel(long n, int cols, const float *__restrict__ a, int lda, const float *__restrict__ b,
                                    int ldb, float *__restrict__ out, int ldo) {
  long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long r = i / cols;
    int c = (int)(i - r * cols);
    out[r * ldo + c] = a[r * lda + c] * b[r * ldb + c];
  }
}

// Counter-based uniform: splitmix64 finaliser of (seed, stream, index) -> 24 random bits -> u in (0,1).
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long stream, unsigned long long idx) {
  unsigned long long z = seed + 0x9e3779b97f4a7c15ULL * (idx + 1ULL) + 0xbf58476d1ce4e5b9ULL * (stream + 1ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  z = z ^ (z >> 31);
  return ((float)(unsigned)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);
}

// Scaled dropout mask, on the device (the reference draws it on the CPU and copies T*S x 2C floats per layer and
// step to the GPU, bilstm-parallel-layer.h:46-94): mask = Heaviside(u - p) / (1 - p).  per_col = 1: one draw per
// column, repeated in every row (MatrixBase::SetRandUniformCol, cpucompute/matrix.cc:952-965).
__global__ void dropout_mask_kernel(long n, int cols, float *__restrict__ mask, int ld, float p, int per_col,
                                    unsigned long long seed, unsigned long long stream) {
  const float scale = 1.0f / (1.0f - p);
  long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long r = i / cols;
    int c = (int)(i - r * cols);
    float u = uniform01(seed, stream, per_col ? (unsigned long long)c : (unsigned long long)i);
    mask[r * ld + c] = (u - p > 0.f) ? scale : 0.f;
  }
}

// partial column sums: block b sums rows b, b+gridDim.x*8, ... ; ws[b][K]
__global__ void col_sum_partial_kernel(int N, int K, const float *__restrict__ x, int ld, float *__restrict__ ws) {
  __shared__ float red[8][33];
  int col = blockIdx.y * 32 + threadIdx.x;
  float s = 0.f;
  if (col < K)
    for (long r = (long)blockIdx.x * 8 + threadIdx.y; r < N; r += (long)gridDim.x * 8) s += x[r * ld + col];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < K) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; q++) t += red[q][threadIdx.x];
    ws[(size_t)blockIdx.x * K + col] = t;
  }
}
__global__ void col_sum_final_kernel(int K, int nblocks, const float *__restrict__ ws, float *__restrict__ out) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= K) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; b++) s += ws[(size_t)b * K + col];
  out[col] = s;
}

// bit 0: a NaN, bit 1: an Inf somewhere in x[0, n)  (Net::Check, reference net.cc:461-468 / CheckNanInf
// utils-functions.h:118 look at the SUM on the host; element tests cannot be fooled by +inf + -inf or overflow)
__global__ void check_finite_kernel(const float *__restrict__ x, long n, int *__restrict__ flags) {
  int f = 0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    if (v != v) f |= 1;
    else if (fabsf(v) == INFINITY) f |= 2;
  }
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

}  // namespace

cudaError_t check_finite(cudaStream_t st, int num_sms, const float *x, long n, int *d_flags) {
  cudaError_t e = cudaMemsetAsync(d_flags, 0, sizeof(int), st);
  if (e != cudaSuccess || n <= 0) return e;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 8 * num_sms) blocks = 8 * num_sms;
  check_finite_kernel<<<blocks, 256, 0, st>>>(x, n, d_flags);
  return cudaGetLastError();
}

cudaError_t sgd_momentum_clip(cudaStream_t st, int num_sms, float *w, float *corr, const float *grad,
                              float momentum, const SgdSegment *d_segs, int nseg, long total) {
  if (total <= 0) return cudaSuccess;
  long vec = (total + 3) / 4;
  int blocks = (int)((vec + 255) / 256);
  if (blocks > 8 * num_sms) blocks = 8 * num_sms;
  sgd_kernel<<<blocks, 256, 0, st>>>(w, corr, grad, momentum, d_segs, nseg, total);
  return cudaGetLastError();
}

cudaError_t optimizer_update(cudaStream_t st, int num_sms, int mode, float *w, float *corr, float *accu,
                             const float *grad, float momentum, float eps, float rho, float one_minus_rho,
                             const SgdSegment *d_segs, int nseg, long total) {
  if (mode == 0) return sgd_momentum_clip(st, num_sms, w, corr, grad, momentum, d_segs, nseg, total);
  if (total <= 0) return cudaSuccess;
  if (!accu || (mode != 1 && mode != 2)) return cudaErrorInvalidValue;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 16 * num_sms) blocks = 16 * num_sms;
  if (mode == 1) opt_kernel<1><<<blocks, 256, 0, st>>>(w, corr, accu, grad, momentum, eps, rho, one_minus_rho, d_segs, nseg, total);
  else opt_kernel<2><<<blocks, 256, 0, st>>>(w, corr, accu, grad, momentum, eps, rho, one_minus_rho, d_segs, nseg, total);
  return cudaGetLastError();
}

cudaError_t mul_elements(cudaStream_t st, int num_sms, int N, int cols, const float *a, int lda, const float *b, int ldb,
                         float *out, int ldo) {
  long n = (long)N * cols;
  if (n <= 0) return cudaSuccess;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 16 * num_sms) blocks = 16 * num_sms;
  mul_elements_kernel<<<blocks, 256, 0, st>>>(n, cols, a, lda, b, ldb, out, ldo);
  return cudaGetLastError();
}

cudaError_t dropout_mask(cudaStream_t st, int num_sms, int rows, int cols, float *mask, int ld, float p, int per_col,
                         unsigned long long seed, unsigned long long stream) {
  long n = (long)rows * cols;
  if (n <= 0) return cudaSuccess;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 16 * num_sms) blocks = 16 * num_sms;
  dropout_mask_kernel<<<blocks, 256, 0, st>>>(n, cols, mask, ld, p, per_col, seed, stream);
  return cudaGetLastError();
}

size_t col_sum_ws_floats(int K, int num_sms) { return (size_t)2 * num_sms * K; }

cudaError_t col_sum(cudaStream_t st, int num_sms, int N, int K, const float *x, int ld, float *out, float *ws) {
  int nb = 2 * num_sms;
  if (nb > (N + 7) / 8) nb = (N + 7) / 8;
  if (nb < 1) nb = 1;
  dim3 grid(nb, (K + 31) / 32), block(32, 8);
  col_sum_partial_kernel<<<grid, block, 0, st>>>(N, K, x, ld, ws);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  col_sum_final_kernel<<<(K + 127) / 128, 128, 0, st>>>(K, nb, ws, out);
  return cudaGetLastError();
}

}  // namespace eb
