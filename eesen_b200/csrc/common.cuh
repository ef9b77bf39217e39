// eesen_b200/csrc/common.cuh -- device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eb {

constexpr float kLogZero = -1e30f;  // reference sentinel, gpucompute/ctc-utils.h:36

__device__ __forceinline__ uint32_t f2u(float x) { return __float_as_uint(x); }
__device__ __forceinline__ float u2f(uint32_t x) { return __uint_as_float(x); }

// TF32 split of an fp32 value: hi keeps the top 19 bits (what the tensor core reads),
// lo = x - hi is exact in fp32.  a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b  ("3xTF32").
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = f2u(x) & 0xffffe000u;
  lo = f2u(x - u2f(hi));
}

// D(16x8,f32) += A(16x8,tf32,row) * B(8x8,tf32,col)   -- legacy warp-level tensor path
// (SASS HMMA.1688.F32.TF32).  Used only where the per-step tile is too small/latency-bound
// for tcgen05 (recurrent kernels).
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Activations on the SFU: one MUFU.EX2 + one MUFU.RCP each (no IEEE-division slow path on the
// serial per-timestep critical path).  The reference evaluates 1/(1+exp(-x)) and
// (e^{2x}-1)/(e^{2x}+1) with double-literal promotion (cuda-kernels.cu:693,718-723); these agree
// to ~2e-7 absolute (ex2.approx 2 ulp, rcp.approx 1 ulp), exact limits at +-inf.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoidf_(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanhf_(float x) {
  // 1 - 2/(e^{2x}+1)
  return fmaf(-2.0f, rcp_approx(ex2_approx(2.8853900817779268f * x) + 1.0f), 1.0f);
}

// release/acquire flag primitives for the inter-CTA step flags of the recurrent kernels
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin with relaxed loads (no L1 invalidate per poll), acquire once at the end
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// "LL" exchange words: a 32-bit payload and a 32-bit step tag travel in ONE 8-byte store, so a
// consumer that sees the expected tag has the payload -- no fence, no separate flag (the idea of
// NCCL's LL protocol).  Loads/stores are relaxed at gpu scope (served by L2, never by a stale L1 line).
__device__ __forceinline__ void st_tagged(uint2 *p, float v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};\n" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_tagged2(const uint4 *p) {
  uint4 q;
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
  return q;
}
__device__ __forceinline__ uint2 ld_tagged(const uint2 *p) {
  uint2 q;
  asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];\n" : "=r"(q.x), "=r"(q.y) : "l"(p) : "memory");
  return q;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace eb
