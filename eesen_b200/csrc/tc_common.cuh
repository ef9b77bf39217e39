// eesen_b200/csrc/tc_common.cuh -- tcgen05 / TMEM / TMA / mbarrier primitives (inline PTX) shared by the
// dense GEMM (gemm_tc.cu) and the recurrent kernels (lstm_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout).
//   K-major operand : LayoutType::SWIZZLE_128B (2): rows of 128 B, 8-row atoms, SBO = 1024 B between atoms
//   MN-major tf32   : LayoutType::SWIZZLE_128B_BASE32B (1) -- the only MN-major layout for 32-bit operands
//                     (cutlass sm100_common.inl:92): rows of 128 B = 32 elements along M/N, 4-k-row atoms
//                     (Swizzle<2,5,2>), SBO = 512 B between k atoms, LBO between 32-wide M/N groups
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);             // start address  [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   // leading byte offset [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;   // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (Blackwell)
  d |= (uint64_t)layout << 61;                         // LayoutType
  return d;
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}


// ---- CTA-pair ("2-SM", cta_group::2) variants: two CTAs of a cluster, ranks 2k / 2k+1, run ONE M = 256 MMA; each
// stages its own 128 rows of A and HALF of B's N rows.  TMA of either CTA reports to the LEADER's (even rank) mbarrier:
// bit 24 of a shared::cluster address selects the CTA of the pair (the same convention as CUTLASS' Sm100MmaPeerBitMask).
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in both CTAs of the pair when the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_2sm(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ---- distributed shared memory (thread-block cluster): address of the same shared-memory offset in CTA `rank` of the
// cluster, and a 16-byte asynchronous store there that reports its bytes to an mbarrier of the destination CTA
// (complete_tx): the consumer sleeps on its own mbarrier until all the bytes it armed (expect_tx) have landed.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async16(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace eb
