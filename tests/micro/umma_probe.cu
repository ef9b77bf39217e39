// tests/micro/umma_probe.cu -- hardware probe for the tcgen05 recurrent-step design.
// Answers, on a B200:
//   1. kind::f16 with hand-filled K-major SWIZZLE_128B tiles (no TMA): is the layout right?
//   2. mixed operand formats (A = fp16, B = bf16) in one instruction: allowed by the hardware?
//   3. TMEM lane layout of an M=64 accumulator (rows -> lanes (r%16) + 32*(r/16))?
//   4. cost of one recurrent step's MMA chain: K=320 (20 k-slices), A = resident weights [M x K]
//      (hi and lo' tiles), B = activations [N x K]; N = 32 (hi|lo') then N = 16 (hi) per k-slice.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

// byte offset of element (row r, k) in a K-major SWIZZLE_128B tile of 16-bit elements with ROWS rows
__host__ __device__ inline size_t sw128_off(int ROWS, int r, int k) {
  int kb = k / 64, kk = k % 64, chunk = kk / 8, within = kk % 8;
  return (size_t)kb * ROWS * 128 + (size_t)(r / 8) * 1024 + (size_t)(r % 8) * 128 + (size_t)((chunk ^ (r % 8)) * 16) +
         within * 2;
}

constexpr int K = 320;
constexpr int KB = K / 64;

// fmt: 0 fp16, 1 bf16
__device__ __forceinline__ uint16_t cvt16(float x, int fmt) {
  if (fmt == 0) return __half_as_ushort(__float2half_rn(x));
  return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}

// out[lane][col] raw accumulator dump (128 lanes x 64 cols)
template <int M>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const float *A, const float *B, int NB, int afmt, int bfmt, float *out, long long *cycles, int reps,
             int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;                          // [M x K] hi
  uint8_t *sA2 = sA + (size_t)M * K * 2;       // second A tile (same contents: timing only)
  uint8_t *sB = sA2 + (size_t)M * K * 2;       // [64 x K]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * K; i += 128) {
    int r = i / K, k = i % K;
    uint16_t v = cvt16(A[i], afmt);
    *reinterpret_cast<uint16_t *>(sA + sw128_off(M, r, k)) = v;
    *reinterpret_cast<uint16_t *>(sA2 + sw128_off(M, r, k)) = v;
  }
  for (int i = tid; i < 64 * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<uint16_t *>(sB + sw128_off(64, r, k)) = r < NB ? cvt16(B[i], bfmt) : 0;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sm)), "n"(128)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_sm;
  const int NH = NB == 48 ? 32 : NB / 2;
  const uint32_t idescN = (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(NB >> 3) << 17) |
                          ((uint32_t)(M >> 4) << 24);
  const uint32_t idescH = (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(NH >> 3) << 17) |
                          ((uint32_t)(M >> 4) << 24);
  uint32_t phase = 0;
  long long t0 = 0, t1 = 0;
  if (tid == 0) t0 = clock64();
  for (int rep = 0; rep < reps; rep++) {
    if (tid == 0) {
      for (int kb = 0; kb < KB; kb++) {
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint64_t ad = umma_desc(smem_u32(sA) + kb * (M * 128) + ks * 32, 16, 1024, 2);
          const uint64_t bd = umma_desc(smem_u32(sB) + kb * (64 * 128) + ks * 32, 16, 1024, 2);
          umma_f16(tmem_base, ad, bd, idescN, (kb | ks) != 0);
          if (mode >= 1) {   // second weight tile against the first half of the B rows -> accumulator at column 64
            const uint64_t ad2 = umma_desc(smem_u32(sA2) + kb * (M * 128) + ks * 32, 16, 1024, 2);
            umma_f16(tmem_base + 64, ad2, bd, idescH, (kb | ks) != 0);
          }
        }
      }
      umma_commit(&bar);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (mode == 2) {   // include a TMEM read + sync in the loop (epilogue pacing)
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(tmem_base + ((uint32_t)(warp * 32) << 16)));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      if (r[0] == 0x12345678u) out[0] = 1.f;
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
  }
  if (tid == 0) {
    t1 = clock64();
    cycles[0] = t1 - t0;
  }
  // dump 128 lanes x 128 columns
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; j++) out[(size_t)(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(128) : "memory");
  }
}

static float round16(float x, int fmt) {
  if (fmt == 0) return __half2float(__float2half_rn(x));
  return __bfloat162float(__float2bfloat16_rn(x));
}

template <int M>
static void run(int NB, int afmt, int bfmt, int reps, int mode, const char *what) {
  std::vector<float> A((size_t)M * K), B((size_t)64 * K);
  srand(7);
  for (auto &v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 0.4f;
  for (auto &v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
  float *dA, *dB, *dO;
  long long *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, 128 * 128 * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, 128 * 128 * 4));
  size_t smem = (size_t)M * K * 2 * 2 + 64 * K * 2 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<M><<<1, 128, smem>>>(dA, dB, NB, afmt, bfmt, dO, dC, reps, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-40s M=%d NB=%d afmt=%d bfmt=%d: LAUNCH/EXEC ERROR %s\n", what, M, NB, afmt, bfmt, cudaGetErrorString(e));
    exit(2);
  }
  std::vector<float> O(128 * 128);
  long long cyc;
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  // reference
  double maxerr = 0, maxref = 0;
  int bad = 0;
  for (int r = 0; r < M; r++) {
    int lane = M == 128 ? r : (r % 16) + 32 * (r / 16);
    for (int n = 0; n < NB; n++) {
      double s = 0;
      for (int k = 0; k < K; k++) s += (double)round16(A[(size_t)r * K + k], afmt) * (double)round16(B[(size_t)n * K + k], bfmt);
      double got = O[(size_t)lane * 128 + n];
      double err = fabs(got - s);
      if (err > maxerr) maxerr = err;
      if (fabs(s) > maxref) maxref = fabs(s);
      if (err > 1e-3) bad++;
      if (mode >= 1 && n < (NB == 48 ? 32 : NB / 2)) {
        double got2 = O[(size_t)lane * 128 + 64 + n];
        if (fabs(got2 - s) > 1e-3) bad++;
      }
    }
  }
  printf("%-40s M=%d NB=%d afmt=%d bfmt=%d mode=%d: max|err|=%.3e (max|ref|=%.2f) bad=%d  cycles/rep=%.1f\n", what, M, NB, afmt,
         bfmt, mode, maxerr, maxref, bad, (double)cyc / reps);
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
}


// ---------------------------------------------------------------------------------------------------------------
// 5. A operand resident in TMEM (tcgen05.mma "TS" form): weights [128 x K] fp16 written once with tcgen05.st
//    (lane = row, 2 halfs per 32-bit column, 8 columns per 16-wide k-slice), B from shared memory as before.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_ts_kernel(const float *A, const float *B, int NB, float *out, long long *cycles, int reps, int two) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sB = smem;   // [64 x K]
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * K; i += 128) {
    int r = i / K, k = i % K;
    *reinterpret_cast<uint16_t *>(sB + sw128_off(64, r, k)) = r < NB ? cvt16(B[i], 0) : 0;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sm)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_sm;
  // A -> TMEM columns [128, 128 + K/2) (and a second copy at [320, 320 + K/2) for the two-tile timing)
  {
    const int row = warp * 32 + lane;
    for (int ks = 0; ks < K / 16; ks++) {
      uint32_t w[8];
      for (int j = 0; j < 8; j++) {
        uint32_t lo = cvt16(A[(size_t)row * K + ks * 16 + 2 * j], 0), hi = cvt16(A[(size_t)row * K + ks * 16 + 2 * j + 1], 0);
        w[j] = lo | (hi << 16);
      }
      for (int cpy = 0; cpy < 2; cpy++) {
        const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((cpy ? 320 : 128) + ks * 8);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(ta), "r"(w[0]), "r"(w[1]),
                     "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t idescN = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idescH = (1u << 4) | ((uint32_t)((NB / 2) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t dB = umma_desc(smem_u32(sB), 16, 1024, 2);
  uint32_t phase = 0;
  long long t0 = 0;
  if (tid == 0) t0 = clock64();
  for (int rep = 0; rep < reps; rep++) {
    if (warp == 0 && two < 10) {
      uint32_t pe;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pe));
      if (pe) {
#pragma unroll
        for (int kb = 0; kb < KB; kb++) {
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const uint64_t bd = dB + (uint64_t)((kb * (64 * 128) + ks * 32) >> 4);
            umma_f16_ts(tmem_base, tmem_base + 128 + (kb * 4 + ks) * 8, bd, idescN, (kb | ks) != 0);
            if (two == 1) umma_f16_ts(tmem_base + 64, tmem_base + 320 + (kb * 4 + ks) * 8, bd, idescH, (kb | ks) != 0);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
    }
    if (warp == 1 && two >= 10) {   // K split over (two - 10) independent accumulator chains (X: N = NB at column c*NB; Y: N = NB/2 behind them)
      uint32_t pe;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pe));
      if (pe) {
        const int nch = two % 10, withy = two >= 30;
#pragma unroll
        for (int kb = 0; kb < KB; kb++) {
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const int s = kb * 4 + ks, c = s % nch;
            const uint64_t bd = dB + (uint64_t)((kb * (64 * 128) + ks * 32) >> 4);
            umma_f16_ts(tmem_base + 512 - 128 + c * 16, tmem_base + 128 + s * 8, bd, idescH, s >= nch);
            if (withy) umma_f16_ts(tmem_base + 64 + c * 16, tmem_base + 128 + s * 8, bd, idescH, s >= nch);
          }
        }
        umma_commit(&bar2);
      }
      __syncwarp();
    }
    if (warp == 1 && two == 2) {   // second accumulator chain issued by a second warp (own commit)
      uint32_t pe;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pe));
      if (pe) {
#pragma unroll
        for (int kb = 0; kb < KB; kb++) {
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const uint64_t bd = dB + (uint64_t)((kb * (64 * 128) + ks * 32) >> 4);
            umma_f16_ts(tmem_base + 64, tmem_base + 320 + (kb * 4 + ks) * 8, bd, idescH, (kb | ks) != 0);
          }
        }
        umma_commit(&bar2);
      }
      __syncwarp();
    }
    if (two < 10) mbar_wait(&bar, phase);
    if (two == 2 || two >= 10) mbar_wait(&bar2, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  if (tid == 0) cycles[0] = clock64() - t0;
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; j++) out[(size_t)(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

static void run_ts(int NB, int reps, int two, const char *what) {
  const int M = 128;
  std::vector<float> A((size_t)M * K), B((size_t)64 * K);
  srand(7);
  for (auto &v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 0.4f;
  for (auto &v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
  float *dA, *dB, *dO;
  long long *dC;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, 128 * 128 * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, 128 * 128 * 4));
  size_t smem = 64 * K * 2 + 2048;
  CK(cudaFuncSetAttribute(probe_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_ts_kernel<<<1, 128, smem>>>(dA, dB, NB, dO, dC, reps, two);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-40s: LAUNCH/EXEC ERROR %s\n", what, cudaGetErrorString(e)); exit(2); }
  std::vector<float> O(128 * 128);
  long long cyc;
  CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int r = 0; r < M; r++)
    for (int n = 0; n < NB; n++) {
      double s = 0;
      for (int k = 0; k < K; k++) s += (double)round16(A[(size_t)r * K + k], 0) * (double)round16(B[(size_t)n * K + k], 0);
      double err = fabs(O[(size_t)r * 128 + n] - s);
      if (err > maxerr) maxerr = err;
      if (err > 1e-3) bad++;
      if (two && n < NB / 2 && fabs(O[(size_t)r * 128 + 64 + n] - s) > 1e-3) bad++;
    }
  printf("%-44s NB=%d two=%d: max|err|=%.3e bad=%d  cycles/rep=%.1f\n", what, NB, two, maxerr, bad, (double)cyc / reps);
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
}


// ---------------------------------------------------------------------------------------------------------------
// 6. The backward step of lstm_tc_bwd_kernel at C = 320 in isolation: a stacked tile (8 MMAs, N = 32, one accumulator) and two
//    full tiles (8 x (N = 32 + N = 16) each), all operands A resident in TMEM, K = 128 (B tile [32 x 128]); one commit per
//    tile.  Reports when the elected thread is done issuing and when each tile's commit is observed (cycles from the
//    start of the issue, mean over reps).  order = 0: tile after tile (as the kernel), 1: all tiles k-slice by k-slice.
__global__ void __launch_bounds__(256, 1) probe_bwd_kernel(long long *cycles, int reps, int order, int noise, const float *gsrc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *sB = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // [32 x 128] fp16
  __shared__ uint64_t bar[3];
  __shared__ uint32_t tmem_base_sm;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 128; i += 256) *reinterpret_cast<uint16_t *>(sB + sw128_off(32, i / 128, i % 128)) = cvt16(0.01f * (i % 37), 0);
  float nacc = 0.f;
  if (tid == 0) {
    for (int i = 0; i < 3; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_sm)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tm = tmem_base_sm;
  {   // some fp16 pattern in the weight columns [160, 480)
    for (int c0 = 160; c0 < 480 && warp < 4; c0 += 8) {
      const uint32_t ta = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      const uint32_t w = 0x2c002c00u;   // 2 x fp16 0.0625
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(ta), "r"(w) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t idN = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idH = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t dB = umma_desc(smem_u32(sB), 16, 1024, 2);
  long long tissue = 0, tw[3] = {0, 0, 0};
  for (int rep = 0; rep < reps; rep++) {
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0) {
      uint32_t pe;
      asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pe));
      if (pe) {
        if (order == 0) {
#pragma unroll
          for (int ks = 0; ks < 8; ks++)
            umma_f16_ts(tm + 96, tm + 416 + ks * 8, dB + (uint64_t)(((ks >> 2) * 4096 + (ks & 3) * 32) >> 4), idN, ks != 0);
          umma_commit(&bar[0]);
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
              const uint64_t bd = dB + (uint64_t)(((ks >> 2) * 4096 + (ks & 3) * 32) >> 4);
              umma_f16_ts(tm + mt * 48, tm + 160 + mt * 128 + ks * 8, bd, idN, ks != 0);
              umma_f16_ts(tm + mt * 48 + 32, tm + 160 + mt * 128 + 64 + ks * 8, bd, idH, ks != 0);
            }
            umma_commit(&bar[1 + mt]);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {
            const uint64_t bd = dB + (uint64_t)(((ks >> 2) * 4096 + (ks & 3) * 32) >> 4);
            umma_f16_ts(tm + 96, tm + 416 + ks * 8, bd, idN, ks != 0);
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
              umma_f16_ts(tm + mt * 48, tm + 160 + mt * 128 + ks * 8, bd, idN, ks != 0);
              umma_f16_ts(tm + mt * 48 + 32, tm + 160 + mt * 128 + 64 + ks * 8, bd, idH, ks != 0);
            }
          }
          for (int i = 0; i < 3; i++) umma_commit(&bar[i]);
        }
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    tissue += t1 - t0;
    if (warp >= 4 && noise) {
      // warps 4-7 keep the SM busy while the tensor pipe works (until the last commit): 1 = tcgen05.ld of the accumulator
      // columns, 2 = shared-memory stores, 3 = global loads, 4 = FMA chains
      uint32_t done = 0;
      int guard = 0;
      while (!done && guard++ < 100000) {
        if (noise == 1) {
          uint32_t r[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(tm + ((uint32_t)((warp & 3) * 32) << 16) + 96u));
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          nacc += __uint_as_float(r[0]);
        } else if (noise == 2) {
          for (int k = 0; k < 16; k++) *reinterpret_cast<volatile uint16_t *>(sB + 9216 + ((tid * 66 + k * 130) & 4095)) = (uint16_t)k;
        } else if (noise == 3) {
          for (int k = 0; k < 8; k++) nacc += __ldcs(gsrc + ((size_t)(rep * 8 + k) * 65536 + tid * 32) % (1u << 24));
        } else {
          for (int k = 0; k < 64; k++) nacc = fmaf(nacc, 1.0001f, 0.5f);
        }
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(&bar[2])), "r"((uint32_t)(rep & 1)) : "memory");
      }
    }
    for (int i = 0; i < 3; i++) {
      mbar_wait(&bar[i], (uint32_t)(rep & 1));
      tw[i] += clock64() - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  if (tid == 0) { cycles[0] = tissue; cycles[1] = tw[0]; cycles[2] = tw[1]; cycles[3] = tw[2]; }
  if (tid == 64) { cycles[4] = tissue; cycles[5] = tw[0]; cycles[6] = tw[1]; cycles[7] = tw[2]; }
  if (nacc == 12345.678f) cycles[8] = 1;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "n"(512) : "memory");
  }
}
static void run_bwd(int order, int noise = 0) {
  long long *dC, h[8];
  float *gsrc;
  CK(cudaMalloc(&dC, 128));
  CK(cudaMalloc(&gsrc, (size_t)(1u << 24) * 4 + 4096));
  CK(cudaMemset(gsrc, 0, (size_t)(1u << 24) * 4 + 4096));
  const int reps = 200;
  CK(cudaFuncSetAttribute(probe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  probe_bwd_kernel<<<1, 256, 16384>>>(dC, reps, order, noise, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("bwd step probe: LAUNCH/EXEC ERROR %s\n", cudaGetErrorString(e)); exit(2); }
  CK(cudaMemcpy(h, dC, 64, cudaMemcpyDeviceToHost));
  printf("[noise %d] bwd step (8 stacked + 2 x 16 MMAs), %s: issuing warp: issue done %.0f, commits seen at %.0f / %.0f / %.0f;  a waiting warp: %.0f / %.0f / %.0f cycles\n",
         noise, order ? "k-slice interleaved" : "tile after tile    ", (double)h[0] / reps, (double)h[1] / reps, (double)h[2] / reps, (double)h[3] / reps,
         (double)h[5] / reps, (double)h[6] / reps, (double)h[7] / reps);
  cudaFree(dC); cudaFree(gsrc);
}

int main() {
  run_bwd(0);
  run_bwd(1);
  for (int nz = 1; nz <= 4; nz++) run_bwd(0, nz);
  run_ts(32, 1, 0, "TS form (A in TMEM), single pass");
  run_ts(32, 1, 1, "TS form, two accumulators");
  run_ts(32, 200, 0, "timing TS: 20 MMA N=32");
  run_ts(32, 200, 1, "timing TS: 20x(N=32 + N=16)");
  run_ts(64, 200, 1, "timing TS: 20x(N=64 + N=32)");
  run_ts(32, 1, 2, "TS form, two accumulators, two issuing warps");
  run_ts(32, 200, 2, "timing TS: 20x N=32 | 20x N=16, 2 issuers");
  run_ts(16, 200, 0, "timing TS: 20 MMA N=16");
  // (results of the K-split timings are not checked: the accumulators sit in other columns)
  run_ts(32, 200, 11, "timing TS: 20 MMA N=16, 1 chain (ref)");
  run_ts(32, 200, 12, "timing TS: 20 MMA N=16, 2 chains");
  run_ts(32, 200, 14, "timing TS: 20 MMA N=16, 4 chains");
  run_ts(32, 200, 18, "timing TS: 20 MMA N=16, 8 chains");
  run_ts(32, 200, 32, "timing TS: 40 MMA N=16, 2+2 chains");
  run_ts(32, 200, 34, "timing TS: 40 MMA N=16, 4+4 chains");
  run_ts(64, 200, 0, "timing TS: 20 MMA N=64");
  run<128>(32, 0, 0, 1, 0, "fp16 x fp16, single pass");
  run<128>(32, 1, 1, 1, 0, "bf16 x bf16, single pass");
  run<64>(32, 0, 0, 1, 0, "M=64 lane layout");
  run<64>(64, 0, 0, 1, 1, "M=64, two accumulators");
  run<128>(32, 0, 0, 1, 1, "two weight tiles (hi: N=32, lo: N=16)");
  // timing: 200 dependent step-chains of 20 k-slices
  run<128>(16, 0, 0, 200, 0, "timing: 20 MMA N=16");
  run<128>(32, 0, 0, 200, 0, "timing: 20 MMA N=32");
  run<128>(64, 0, 0, 200, 0, "timing: 20 MMA N=64");
  run<128>(32, 0, 0, 200, 1, "timing: 20x(N=32 + N=16), 2 A tiles");
  run<128>(32, 0, 0, 200, 2, "timing: same + tcgen05.ld + barrier");
  run<128>(64, 0, 0, 200, 1, "timing: 20x(N=64 + N=32)");
  run<64>(16, 0, 0, 200, 0, "timing: M=64 20 MMA N=16");
  run<64>(32, 0, 0, 200, 0, "timing: M=64 20 MMA N=32");
  run<64>(64, 0, 0, 200, 0, "timing: M=64 20 MMA N=64");
  run<64>(64, 0, 0, 200, 1, "timing: M=64 20x(N=64 + N=32)");
  run<64>(32, 0, 0, 200, 1, "timing: M=64 20x(N=32 + N=16)");
  // (A = fp16 with B = bf16 in one instruction: "illegal instruction" on B200 -- measured, removed)
  return 0;
}
