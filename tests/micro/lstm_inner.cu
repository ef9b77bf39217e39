// Microbenchmark (test infrastructure): the inner product loop of lstm_fwd_kernel in isolation -- 20 warps,
// each 2 M-tiles x 10 k-slices of resident fp32 weights in shared memory (fragment order) against an 8-column
// activation tile, 3xTF32 -- in several instruction-schedule variants, to see what limits the phase.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o eesen_b200/bin/lstm_inner tests/micro/lstm_inner.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float u2f(uint32_t x) { return __uint_as_float(x); }
__device__ __forceinline__ void split(float x, uint32_t &hi, uint32_t &lo) { hi = f2u(x) & 0xffffe000u; lo = f2u(x - u2f(hi)); }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void step3(float (&acc)[4], float (&accc)[4], const float4 &A, float b0, float b1) {
  uint32_t ah[4], al[4], bh[2], bl[2];
  split(A.x, ah[0], al[0]); split(A.y, ah[1], al[1]); split(A.z, ah[2], al[2]); split(A.w, ah[3], al[3]);
  split(b0, bh[0], bl[0]); split(b1, bh[1], bl[1]);
  mma_tf32(accc, al, bh); mma_tf32(accc, ah, bl); mma_tf32(acc, ah, bh);
}
__device__ __forceinline__ void step1(float (&acc)[4], const float4 &A, float b0, float b1) {
  uint32_t a[4] = {f2u(A.x), f2u(A.y), f2u(A.z), f2u(A.w)}, b[2] = {f2u(b0), f2u(b1)};
  mma_tf32(acc, a, b);
}
// two k-slices: 2 tf32 (hi*hi, raw operands: the low 13 bits are ignored by the instruction? NO for mma.sync -> mask)
// + 2 bf16 k16 corrections (lo*hi, hi*lo) with the k relabelling of DESIGN.md
__device__ __forceinline__ void step_mixed(float (&acc)[4], float (&accc)[4], const float4 &A, const float4 &An, float b0, float b1,
                                           float b0n, float b1n) {
  uint32_t ah[4], al[4], ahn[4], aln[4], bh[2], bl[2], bhn[2], bln[2];
  split(A.x, ah[0], al[0]); split(A.y, ah[1], al[1]); split(A.z, ah[2], al[2]); split(A.w, ah[3], al[3]);
  split(An.x, ahn[0], aln[0]); split(An.y, ahn[1], aln[1]); split(An.z, ahn[2], aln[2]); split(An.w, ahn[3], aln[3]);
  split(b0, bh[0], bl[0]); split(b1, bh[1], bl[1]); split(b0n, bhn[0], bln[0]); split(b1n, bhn[1], bln[1]);
  mma_tf32(acc, ah, bh); mma_tf32(acc, ahn, bhn);
  uint32_t alo[4] = {pack_bf16(u2f(al[0]), u2f(al[2])), pack_bf16(u2f(al[1]), u2f(al[3])), pack_bf16(u2f(aln[0]), u2f(aln[2])),
                     pack_bf16(u2f(aln[1]), u2f(aln[3]))};
  uint32_t ahi[4] = {pack_bf16(A.x, A.z), pack_bf16(A.y, A.w), pack_bf16(An.x, An.z), pack_bf16(An.y, An.w)};
  uint32_t bhi[2] = {pack_bf16(b0, b1), pack_bf16(b0n, b1n)};
  uint32_t blo[2] = {pack_bf16(u2f(bl[0]), u2f(bl[1])), pack_bf16(u2f(bln[0]), u2f(bln[1]))};
  mma_bf16(accc, alo, bhi); mma_bf16(accc, ahi, blo);
}

constexpr int NCT = 5, KSPLIT = 4, KS = 40, C = 320, SST = C + 4;

template <int V>
__global__ void __launch_bounds__(640, 1) inner_kernel(int iters, float *out, long long *cycles) {
  extern __shared__ __align__(16) float smem[];
  float *Wsm = smem;                       // [NCT][2][KS][32][4]
  float *stg = Wsm + NCT * KS * 256;       // [8][SST]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
  const int ct = warp % NCT, ksid = warp / NCT;
  for (int i = tid; i < NCT * KS * 256; i += 640) Wsm[i] = 0.001f * (float)((i * 37) % 201 - 100);
  for (int i = tid; i < 8 * SST; i += 640) stg[i] = 0.01f * (float)((i * 11) % 97 - 48);
  __syncthreads();
  const int kb = ksid * 10, ke = kb + 10;
  const float4 *W0 = reinterpret_cast<const float4 *>(Wsm) + ((size_t)(ct * 2 + 0) * KS) * 32 + lane;
  const float4 *W1 = reinterpret_cast<const float4 *>(Wsm) + ((size_t)(ct * 2 + 1) * KS) * 32 + lane;
  const float *brow = stg + (size_t)g * SST + tg;
  float tot = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, accc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    if (V == 0) {          // the kernel's loop (unroll 2, loads at the top of each slice)
#pragma unroll 2
      for (int ks = kb; ks < ke; ks++) {
        float4 A0 = W0[ks * 32], A1 = W1[ks * 32];
        float b0 = brow[ks * 8], b1 = brow[ks * 8 + 4];
        step3(acc[0], accc[0], A0, b0, b1); step3(acc[1], accc[1], A1, b0, b1);
      }
    } else if (V == 1) {   // all operands of the warp's 10 slices loaded up front (register resident), then the MMAs
      float4 A0[10], A1[10]; float b0[10], b1[10];
#pragma unroll
      for (int q = 0; q < 10; q++) { A0[q] = W0[(kb + q) * 32]; A1[q] = W1[(kb + q) * 32]; b0[q] = brow[(kb + q) * 8]; b1[q] = brow[(kb + q) * 8 + 4]; }
#pragma unroll
      for (int q = 0; q < 10; q++) { step3(acc[0], accc[0], A0[q], b0[q], b1[q]); step3(acc[1], accc[1], A1[q], b0[q], b1[q]); }
    } else if (V == 2) {   // no shared-memory loads at all (operands fixed in registers): tensor + split ALU only
      float4 A0 = W0[kb * 32], A1 = W1[kb * 32]; float b0 = brow[kb * 8], b1 = brow[kb * 8 + 4];
#pragma unroll 2
      for (int ks = kb; ks < ke; ks++) {
        step3(acc[0], accc[0], A0, b0, b1); step3(acc[1], accc[1], A1, b0, b1);
        A0.x += 1e-6f; A1.y += 1e-6f;
      }
    } else if (V == 3) {   // loads only (no MMAs): shared-memory side alone
#pragma unroll 2
      for (int ks = kb; ks < ke; ks++) {
        float4 A0 = W0[ks * 32], A1 = W1[ks * 32];
        float b0 = brow[ks * 8], b1 = brow[ks * 8 + 4];
        acc[0][0] += A0.x + A0.y + A0.z + A0.w + b0; acc[1][0] += A1.x + A1.y + A1.z + A1.w + b1;
      }
    } else if (V == 4) {   // one term (tf32 mode)
#pragma unroll 2
      for (int ks = kb; ks < ke; ks++) {
        float4 A0 = W0[ks * 32], A1 = W1[ks * 32];
        float b0 = brow[ks * 8], b1 = brow[ks * 8 + 4];
        step1(acc[0], A0, b0, b1); step1(acc[1], A1, b0, b1);
      }
    } else if (V == 5) {   // mixed: tf32 hi*hi + bf16 k16 corrections, two slices at a time
#pragma unroll 1
      for (int ks = kb; ks < ke; ks += 2) {
        float4 A0 = W0[ks * 32], A1 = W1[ks * 32], A0n = W0[(ks + 1) * 32], A1n = W1[(ks + 1) * 32];
        float b0 = brow[ks * 8], b1 = brow[ks * 8 + 4], b0n = brow[ks * 8 + 8], b1n = brow[ks * 8 + 12];
        step_mixed(acc[0], accc[0], A0, A0n, b0, b1, b0n, b1n); step_mixed(acc[1], accc[1], A1, A1n, b0, b1, b0n, b1n);
      }
    } else if (V == 6) {   // 3 terms, MMAs only with pre-split operands in registers (no ALU, no loads): pure tensor issue
      uint32_t ah[4] = {1, 2, 3, 4}, al[4] = {5, 6, 7, 8}, bh[2] = {9, 10}, bl[2] = {11, 12};
#pragma unroll 2
      for (int ks = kb; ks < ke; ks++) {
        mma_tf32(accc[0], al, bh); mma_tf32(accc[0], ah, bl); mma_tf32(acc[0], ah, bh);
        mma_tf32(accc[1], al, bh); mma_tf32(accc[1], ah, bl); mma_tf32(acc[1], ah, bh);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int c = 0; c < 4; c++) tot += acc[i][c] + accc[i][c];
    __syncthreads();   // the kernel has block barriers around this phase too
  }
  long long t1 = clock64();
  out[blockIdx.x * 640 + tid] = tot;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

// Variant family B: NW warps per CTA (K split over NW/5 warps), the first R k-slices of every warp's weights
// stay in registers across steps (loaded once), the rest comes from shared memory every step.
template <int NW, int R>
__global__ void __launch_bounds__(32 * NW, 1) resident_kernel(int iters, float *out, long long *cycles) {
  extern __shared__ __align__(16) float smem[];
  float *Wsm = smem;
  float *stg = Wsm + NCT * KS * 256;
  constexpr int NT = 32 * NW, KSPL = NW / NCT, PER = KS / KSPL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
  const int ct = warp % NCT, ksid = warp / NCT;
  for (int i = tid; i < NCT * KS * 256; i += NT) Wsm[i] = 0.001f * (float)((i * 37) % 201 - 100);
  for (int i = tid; i < 8 * SST; i += NT) stg[i] = 0.01f * (float)((i * 11) % 97 - 48);
  __syncthreads();
  const int kb = ksid * PER;
  const float4 *W0 = reinterpret_cast<const float4 *>(Wsm) + ((size_t)(ct * 2 + 0) * KS) * 32 + lane;
  const float4 *W1 = reinterpret_cast<const float4 *>(Wsm) + ((size_t)(ct * 2 + 1) * KS) * 32 + lane;
  const float *brow = stg + (size_t)g * SST + tg;
  float4 R0[R > 0 ? R : 1], R1[R > 0 ? R : 1];
#pragma unroll
  for (int q = 0; q < R; q++) { R0[q] = W0[(kb + q) * 32]; R1[q] = W1[(kb + q) * 32]; }
  float tot = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, accc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
    for (int q = 0; q < R; q++) {
      float b0 = brow[(kb + q) * 8], b1 = brow[(kb + q) * 8 + 4];
      step3(acc[0], accc[0], R0[q], b0, b1); step3(acc[1], accc[1], R1[q], b0, b1);
    }
#pragma unroll 2
    for (int ks = kb + R; ks < kb + PER; ks++) {
      float4 A0 = W0[ks * 32], A1 = W1[ks * 32];
      float b0 = brow[ks * 8], b1 = brow[ks * 8 + 4];
      step3(acc[0], accc[0], A0, b0, b1); step3(acc[1], accc[1], A1, b0, b1);
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int c = 0; c < 4; c++) tot += acc[i][c] + accc[i][c];
    __syncthreads();
  }
  long long t1 = clock64();
  out[blockIdx.x * NT + tid] = tot;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NW, int R>
void run_resident(const char *name) {
  int iters = 500, blocks = 148;
  size_t smem = sizeof(float) * (NCT * KS * 256 + 8 * SST);
  float *out; long long *cyc;
  cudaMalloc(&out, sizeof(float) * blocks * 32 * NW);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  cudaFuncSetAttribute(resident_kernel<NW, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  resident_kernel<NW, R><<<blocks, 32 * NW, smem>>>(5, out, cyc);
  resident_kernel<NW, R><<<blocks, 32 * NW, smem>>>(iters, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
  printf("%-62s %7.0f clk per step  (%s)\n", name, c / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

template <int V>
void run(const char *name) {
  int iters = 500, blocks = 148;
  size_t smem = sizeof(float) * (NCT * KS * 256 + 8 * SST);
  float *out; long long *cyc;
  cudaMalloc(&out, sizeof(float) * blocks * 640);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  cudaFuncSetAttribute(inner_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  inner_kernel<V><<<blocks, 640, smem>>>(5, out, cyc);
  inner_kernel<V><<<blocks, 640, smem>>>(iters, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
  printf("%-62s %7.0f clk per step  (%s)\n", name, c / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("V0 kernel loop: LDS at slice top, 3xTF32");
  run<1>("V1 all 10 slices loaded first, then 60 MMAs");
  run<2>("V2 no LDS: split ALU + 60 MMAs");
  run<3>("V3 LDS only");
  run<4>("V4 one term (tf32 mode) with LDS");
  run<5>("V5 mixed tf32 + bf16 corrections with LDS");
  run<6>("V6 60 MMAs only (pre-split registers)");
  run_resident<20, 0>("B 20 warps, 0 slices resident");
  run_resident<20, 1>("B 20 warps, 1 of 10 slices resident");
  run_resident<20, 2>("B 20 warps, 2 of 10 slices resident");
  run_resident<10, 0>("B 10 warps, 0 of 20 slices resident");
  run_resident<10, 4>("B 10 warps, 4 of 20 slices resident");
  run_resident<10, 8>("B 10 warps, 8 of 20 slices resident");
  run_resident<10, 12>("B 10 warps, 12 of 20 slices resident");
  run_resident<5, 0>("B 5 warps, 0 of 40 slices resident");
  run_resident<5, 16>("B 5 warps, 16 of 40 slices resident");
  run_resident<5, 24>("B 5 warps, 24 of 40 slices resident");
  return 0;
}
