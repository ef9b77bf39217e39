// eesen_b200/csrc/gemm.cu -- dense fp32-storage GEMM for the input-side contractions of the
// BiLSTM/affine layers:  C = alpha * op(A) * op(B) + beta * C  (+ bias row), strided-batched.
//
// Replaces CuMatrixBase::AddMatMat -> cublasSgemm (reference gpucompute/cuda-matrix.cc:603-639,
// cublas-wrappers.h:28-30) at these call sites of the hot path:
//   bilstm-parallel-layer.h:109,163 (X*Wx^T, NT)      :502,593 (DGIFO*Wx, NN)
//   bilstm-parallel-layer.h:505-506,596-597 (DGIFO^T*X, DGIFO^T*M, TN)
//   affine-trans-layer.h:165 (NT), :171 (NN), :182 (TN)
// and folds AddVecToRows (bias, :110,164 / affine :163) into the epilogue.
//
// Tensor-core path: warp-level mma.sync m16n8k8 TF32 with an fp32 accumulator.  Precision
// modes: 0 = "3xTF32" (hi/lo split of both operands, 3 MMAs; fp32-faithful, error ~2^-21),
// 1 = single-pass TF32, 2 = BF16x1 (operands rounded to bf16, m16n8k16).
// Tiles: 128x128x32 (or 128x64x32 for narrow N), 256 threads, 3-stage cp.async pipeline,
// padded shared tiles (conflict-free fragment reads for every layout).  Split-K with a
// deterministic workspace reduction for the long-K weight-gradient products.
#include "common.cuh"
#include "kernels.h"

namespace eb {

namespace {

constexpr int BK = 32;
constexpr int STAGES = 3;
constexpr int THREADS = 256;

template <int BM, int BN>
struct GemmSmem {
  // op(A) tile: TA=0 -> [BM][BK+4] (k contiguous) ; TA=1 -> [BK][BM+8] (m contiguous)
  // op(B) tile: TB=1 -> [BN][BK+4] (k contiguous) ; TB=0 -> [BK][BN+8] (n contiguous)
  static constexpr int A_ELEMS = (BM * (BK + 4) > BK * (BM + 8)) ? BM * (BK + 4) : BK * (BM + 8);
  static constexpr int B_ELEMS = (BN * (BK + 4) > BK * (BN + 8)) ? BN * (BK + 4) : BK * (BN + 8);
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr int BYTES = STAGES * STAGE_ELEMS * 4;
};

struct GemmArgs {
  int M, N, K;
  const float *A; int lda; long strideA;
  const float *B; int ldb; long strideB;
  float *C; int ldc; long strideC;
  const float *bias; long strideBias;   // optional [N] row added to every output row
  float alpha, beta;
  int splits;             // split-K factor (1 = direct epilogue)
  int k_per_split;        // multiple of BK
  float *ws;              // [batch][splits][M*N] partials when splits > 1
};

// Loads one BMxBK (or BKxBM) operand tile with 16-byte cp.async; rows/cols outside the matrix
// and the K tail are zero-filled via the src-size operand.
template <int ROWS, int TRANS>
__device__ __forceinline__ void load_tile(float *s, const float *g, int ld, int row0, int k0,
                                          int nrows, int kend, int tid) {
  if (TRANS == 0) {
    // global [row][k], k contiguous -> smem [ROWS][BK+4]
    constexpr int VPR = BK / 4;  // float4 per row
    for (int v = tid; v < ROWS * VPR; v += THREADS) {
      int r = v / VPR, c = (v % VPR) * 4;
      int gr = row0 + r, gk = k0 + c;
      int bytes = 0;
      const float *src = g;
      if (gr < nrows && gk < kend) {
        bytes = min(4, kend - gk) * 4;
        src = g + (long)gr * ld + gk;
      }
      cp_async16(s + r * (BK + 4) + c, src, bytes);
    }
  } else {
    // global [k][row], row contiguous -> smem [BK][ROWS+8]
    constexpr int VPR = ROWS / 4;
    for (int v = tid; v < BK * VPR; v += THREADS) {
      int k = v / VPR, c = (v % VPR) * 4;
      int gk = k0 + k, gr = row0 + c;
      int bytes = 0;
      const float *src = g;
      if (gk < kend && gr < nrows) {
        bytes = min(4, nrows - gr) * 4;
        src = g + (long)gk * ld + gr;
      }
      cp_async16(s + k * (ROWS + 8) + c, src, bytes);
    }
  }
}

template <int ROWS, int TRANS>
__device__ __forceinline__ float lds_elem(const float *s, int r, int k) {
  return TRANS == 0 ? s[r * (BK + 4) + k] : s[k * (ROWS + 8) + r];
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// TA: 0 -> A is [M x K] row-major, 1 -> A is stored [K x M] (op(A) = A^T)
// TB: 0 -> B is [K x N] row-major, 1 -> B is stored [N x K] (op(B) = B^T)
template <int BM, int BN, int WM, int WN, int TA, int TB, int PREC>
__global__ void __launch_bounds__(THREADS, (BN == 128 ? 1 : 2))
gemm_kernel(GemmArgs p) {
  extern __shared__ __align__(16) float smem[];
  using SM = GemmSmem<BM, BN>;
  constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static_assert(WARPS_M * WARPS_N * 32 == THREADS, "warp layout");
  constexpr int MT = WM / 16, NT = WN / 8;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int wm = (warp / WARPS_N) * WM, wn = (warp % WARPS_N) * WN;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z % p.splits;

  const float *A = p.A + batch * p.strideA;
  const float *B = p.B + batch * p.strideB;
  const int kbeg = split * p.k_per_split;
  const int kend = min(p.K, kbeg + p.k_per_split);
  const int ktiles = (kend - kbeg + BK - 1) / BK;

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[i][j][c] = 0.f;

  auto issue = [&](int kt) {
    float *sa = smem + (kt % STAGES) * SM::STAGE_ELEMS;
    float *sb = sa + SM::A_ELEMS;
    load_tile<BM, TA>(sa, A, p.lda, m0, kbeg + kt * BK, p.M, kend, tid);
    load_tile<BN, TB ? 0 : 1>(sb, B, p.ldb, n0, kbeg + kt * BK, p.N, kend, tid);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < ktiles) issue(s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (kt + STAGES - 1 < ktiles) issue(kt + STAGES - 1);
    cp_async_commit();

    const float *sa = smem + (kt % STAGES) * SM::STAGE_ELEMS;
    const float *sb = sa + SM::A_ELEMS;

    if (PREC == 2) {
      // BF16: two k8 halves form one m16n8k16
#pragma unroll
      for (int kk = 0; kk < BK; kk += 16) {
        uint32_t af[MT][4], bf[NT][2];
#pragma unroll
        for (int i = 0; i < MT; i++) {
          int r = wm + i * 16 + g;
          af[i][0] = pack_bf16(lds_elem<BM, TA>(sa, r, kk + 2 * tg), lds_elem<BM, TA>(sa, r, kk + 2 * tg + 1));
          af[i][1] = pack_bf16(lds_elem<BM, TA>(sa, r + 8, kk + 2 * tg), lds_elem<BM, TA>(sa, r + 8, kk + 2 * tg + 1));
          af[i][2] = pack_bf16(lds_elem<BM, TA>(sa, r, kk + 2 * tg + 8), lds_elem<BM, TA>(sa, r, kk + 2 * tg + 9));
          af[i][3] = pack_bf16(lds_elem<BM, TA>(sa, r + 8, kk + 2 * tg + 8), lds_elem<BM, TA>(sa, r + 8, kk + 2 * tg + 9));
        }
#pragma unroll
        for (int j = 0; j < NT; j++) {
          int c = wn + j * 8 + g;
          bf[j][0] = pack_bf16(lds_elem<BN, TB ? 0 : 1>(sb, c, kk + 2 * tg), lds_elem<BN, TB ? 0 : 1>(sb, c, kk + 2 * tg + 1));
          bf[j][1] = pack_bf16(lds_elem<BN, TB ? 0 : 1>(sb, c, kk + 2 * tg + 8), lds_elem<BN, TB ? 0 : 1>(sb, c, kk + 2 * tg + 9));
        }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
          for (int j = 0; j < NT; j++) mma_bf16(acc[i][j], af[i], bf[j]);
      }
    } else {
#pragma unroll
      for (int kk = 0; kk < BK; kk += 8) {
        uint32_t ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
        for (int i = 0; i < MT; i++) {
          int r = wm + i * 16 + g;
          float v0 = lds_elem<BM, TA>(sa, r, kk + tg), v1 = lds_elem<BM, TA>(sa, r + 8, kk + tg);
          float v2 = lds_elem<BM, TA>(sa, r, kk + tg + 4), v3 = lds_elem<BM, TA>(sa, r + 8, kk + tg + 4);
          if (PREC == 0) {
            split_tf32(v0, ah[i][0], al[i][0]); split_tf32(v1, ah[i][1], al[i][1]);
            split_tf32(v2, ah[i][2], al[i][2]); split_tf32(v3, ah[i][3], al[i][3]);
          } else {
            ah[i][0] = f2u(v0); ah[i][1] = f2u(v1); ah[i][2] = f2u(v2); ah[i][3] = f2u(v3);
          }
        }
#pragma unroll
        for (int j = 0; j < NT; j++) {
          int c = wn + j * 8 + g;
          float v0 = lds_elem<BN, TB ? 0 : 1>(sb, c, kk + tg), v1 = lds_elem<BN, TB ? 0 : 1>(sb, c, kk + tg + 4);
          if (PREC == 0) {
            split_tf32(v0, bh[j][0], bl[j][0]); split_tf32(v1, bh[j][1], bl[j][1]);
          } else {
            bh[j][0] = f2u(v0); bh[j][1] = f2u(v1);
          }
        }
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
          for (int j = 0; j < NT; j++) {
            if (PREC == 0) {
              mma_tf32(acc[i][j], al[i], bh[j]);  // small terms first
              mma_tf32(acc[i][j], ah[i], bl[j]);
            }
            mma_tf32(acc[i][j], ah[i], bh[j]);
          }
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue
  if (p.splits > 1) {
    float *W = p.ws + ((long)blockIdx.z) * ((long)p.M * p.N);
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          int r = m0 + wm + i * 16 + g + h * 8;
          int c = n0 + wn + j * 8 + 2 * tg;
          if (r < p.M) {
            if (c < p.N) W[(long)r * p.N + c] = acc[i][j][2 * h];
            if (c + 1 < p.N) W[(long)r * p.N + c + 1] = acc[i][j][2 * h + 1];
          }
        }
    return;
  }
  float *Cb = p.C + batch * p.strideC;
  const float *bias = p.bias ? p.bias + batch * p.strideBias : nullptr;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NT; j++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int r = m0 + wm + i * 16 + g + h * 8;
        int c = n0 + wn + j * 8 + 2 * tg;
        if (r >= p.M) continue;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          if (c + e < p.N) {
            float v = p.alpha * acc[i][j][2 * h + e];
            if (bias) v += bias[c + e];
            float *dst = Cb + (long)r * p.ldc + c + e;
            if (p.beta != 0.f) v += p.beta * (*dst);
            *dst = v;
          }
        }
      }
}

// C = alpha * sum_s ws[s] + beta * C (+bias): deterministic split-K reduction
__global__ void splitk_reduce_kernel(GemmArgs p) {
  long n = (long)p.M * p.N;
  int batch = blockIdx.y;
  const float *W = p.ws + (long)batch * p.splits * n;
  float *Cb = p.C + batch * p.strideC;
  const float *bias = p.bias ? p.bias + batch * p.strideBias : nullptr;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < p.splits; k++) s += W[k * n + i];
    int r = (int)(i / p.N), c = (int)(i % p.N);
    float v = p.alpha * s;
    if (bias) v += bias[c];
    float *dst = Cb + (long)r * p.ldc + c;
    if (p.beta != 0.f) v += p.beta * (*dst);
    *dst = v;
  }
}

template <int BM, int BN, int WM, int WN, int TA, int TB, int PREC>
cudaError_t launch_cfg(const GemmArgs &p, int batch, cudaStream_t st) {
  using SM = GemmSmem<BM, BN>;
  auto kern = gemm_kernel<BM, BN, WM, WN, TA, TB, PREC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, batch * p.splits);
  kern<<<grid, THREADS, SM::BYTES, st>>>(p);
  return cudaGetLastError();
}

template <int TA, int TB, int PREC>
cudaError_t launch_layout(const GemmArgs &p, int batch, cudaStream_t st) {
  if (p.N > 64) return launch_cfg<128, 128, 64, 32, TA, TB, PREC>(p, batch, st);
  return launch_cfg<128, 64, 32, 32, TA, TB, PREC>(p, batch, st);
}

template <int PREC>
cudaError_t launch_prec(int ta, int tb, const GemmArgs &p, int batch, cudaStream_t st) {
  if (ta == 0 && tb == 1) return launch_layout<0, 1, PREC>(p, batch, st);
  if (ta == 0 && tb == 0) return launch_layout<0, 0, PREC>(p, batch, st);
  if (ta == 1 && tb == 0) return launch_layout<1, 0, PREC>(p, batch, st);
  return cudaErrorInvalidValue;  // TT is not used by the path
}

}  // namespace

size_t gemm_workspace_bytes(int M, int N, int K, int batch, int num_sms) {
  (void)K;
  // worst case: enough splits to fill ~2 waves
  long tiles = (long)((M + 127) / 128) * ((N + 63) / 64) * batch;
  int splits = 1;
  if (tiles < num_sms) splits = (int)((2L * num_sms + tiles - 1) / tiles);
  if (splits > 64) splits = 64;
  return (size_t)splits * batch * (size_t)M * N * sizeof(float);
}

cudaError_t gemm(cudaStream_t st, int num_sms, int transA, int transB, int M, int N, int K, float alpha,
                 const float *A, int lda, long strideA, const float *B, int ldb, long strideB, float beta,
                 float *C, int ldc, long strideC, const float *bias, long strideBias, int batch,
                 int precision, float *ws, size_t ws_bytes) {
  if (M <= 0 || N <= 0 || batch <= 0) return cudaSuccess;
  if ((lda & 3) || (ldb & 3) || (((uintptr_t)A) & 15) || (((uintptr_t)B) & 15) || (strideA & 3) || (strideB & 3))
    return cudaErrorMisalignedAddress;
  GemmArgs p;
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.strideA = strideA;
  p.B = B; p.ldb = ldb; p.strideB = strideB;
  p.C = C; p.ldc = ldc; p.strideC = strideC;
  p.bias = bias; p.strideBias = strideBias;
  p.alpha = alpha; p.beta = beta;
  p.splits = 1; p.k_per_split = ((K + BK - 1) / BK) * BK; p.ws = ws;
  if (K <= 0) p.k_per_split = BK;
  // split-K only for long reductions that cannot fill the machine (the weight-gradient products)
  int bn = N > 64 ? 128 : 64;
  long tiles = (long)((M + 127) / 128) * ((N + bn - 1) / bn) * batch;
  if (tiles < num_sms && K >= 4096 && ws) {
    int splits = (int)((2L * num_sms + tiles - 1) / tiles);
    int max_by_k = K / 1024;
    if (splits > max_by_k) splits = max_by_k;
    if (splits > 64) splits = 64;
    while (splits > 1 && (size_t)splits * batch * (size_t)M * N * sizeof(float) > ws_bytes) splits--;
    if (splits > 1) {
      int kt = (K + BK - 1) / BK;
      int per = (kt + splits - 1) / splits;
      p.k_per_split = per * BK;
      p.splits = (kt + per - 1) / per;
    }
  }
  cudaError_t e;
  if (precision == 0) e = launch_prec<0>(transA, transB, p, batch, st);
  else if (precision == 1) e = launch_prec<1>(transA, transB, p, batch, st);
  else if (precision == 2) e = launch_prec<2>(transA, transB, p, batch, st);
  else return cudaErrorInvalidValue;
  if (e != cudaSuccess) return e;
  if (p.splits > 1) {
    long n = (long)M * N;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4 * num_sms) blocks = 4 * num_sms;
    splitk_reduce_kernel<<<dim3(blocks, batch), 256, 0, st>>>(p);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace eb
