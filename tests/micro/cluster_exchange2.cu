// Microbenchmark (test infrastructure): per-step cost of the recurrent exchange of lstm_tc_fwd_kernel at its real
// footprint (one CTA per SM, 256 threads, 32 cells x 16 utterances per CTA, `SL` CTAs per (direction, group)):
//   (a) through L2 with 4-byte tagged words (what lstm_tc.cu does today): 2 tagged stores per thread, then every CTA
//       polls the whole [16 x C] block of its group with 16-byte relaxed loads;
//   (b) through distributed shared memory inside a thread-block cluster of SL CTAs: every CTA pushes its 2 KB slab
//       (hi | lo' halves, 16-byte chunks) into the receive tile of all SL CTAs with st.async + mbarrier complete_tx.
// Prints cycles per dependent step (pure ping-pong: no compute between steps) and checks the payload.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o eesen_b200/bin/cluster_exchange2 tests/micro/cluster_exchange2.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int THREADS = 256, UG = 16, CS = 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async16(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ---- (b) DSMEM.  Receive tile per parity: [SL producers][2 halves][16 utts][32 cells] fp16 = SL * 2 KB.
// Thread mapping of the push: warp w owns utterances 2w, 2w+1 (as the gate phase of the kernel): 2 utts x 2 halves x 4
// chunks of 16 B = 16 chunks per destination; lane l pushes chunk (l & 15) to destinations (l >> 4), (l >> 4) + 2, ...
__global__ void __launch_bounds__(THREADS, 1) dsmem_kernel(int iters, int SL, int delay, unsigned *out, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  const uint32_t tile = smem_u32(smem), slab = 2048u, tile_bytes = (uint32_t)SL * slab;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int par = it & 1;
    if (tid == 0) mbar_expect_tx(&bar[par], tile_bytes);
    // payload: (it, rank, warp, chunk) -- checked by the consumer
    const int ch = lane & 15;
    const uint4 v = make_uint4((uint32_t)it, rank, (uint32_t)warp, (uint32_t)ch);
    const uint32_t off = (uint32_t)par * tile_bytes + rank * slab + (uint32_t)(warp * 16 + ch) * 16u;
    for (int d = lane >> 4; d < SL; d += 2) st_async16(mapa(tile + off, (uint32_t)d), v, mapa(smem_u32(&bar[par]), (uint32_t)d));
    mbar_wait(&bar[par], (uint32_t)((it >> 1) & 1));
    // consume: every thread reads one 16-byte chunk per producer it is responsible for (checks the step number)
    for (int i = tid; i < SL * 128; i += THREADS) {
      const uint4 q = *reinterpret_cast<const uint4 *>(smem + (size_t)par * tile_bytes + (size_t)i * 16);
      acc += (q.x == (uint32_t)it) ? 1u : 1000000u;
    }
    if (delay) { long long t = clock64(); while (clock64() - t < delay) {} }
  }
  long long t1 = clock64();
  cluster_sync_all();
  out[blockIdx.x * THREADS + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}


// ---- (c) DSMEM bulk copies: the slab is staged locally (generic stores), then ONE elected thread per destination issues a
// 2 KB cp.async.bulk shared::cta -> shared::cluster with complete_tx on the destination's mbarrier
__device__ __forceinline__ void bulk_s2s(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(rbar)
               : "memory");
}
__global__ void __launch_bounds__(THREADS, 1) bulk_kernel(int iters, int SL, int delay, unsigned *out, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(rank));
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  const uint32_t tile = smem_u32(smem), slab = 2048u, tile_bytes = (uint32_t)SL * slab;
  uint8_t *stage = smem + 2 * tile_bytes;   // [2 parity][2 KB] local staging
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int par = it & 1;
    if (tid == 0) mbar_expect_tx(&bar[par], tile_bytes);
    // every thread writes its 8 bytes of the slab (it, rank, ...), then the proxy fence and a CTA barrier
    *reinterpret_cast<uint2 *>(stage + par * slab + tid * 8) = make_uint2((uint32_t)it, rank);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if (tid < SL)
      bulk_s2s(mapa(tile + (uint32_t)par * tile_bytes + rank * slab, (uint32_t)tid), smem_u32(stage + par * slab), slab,
               mapa(smem_u32(&bar[par]), (uint32_t)tid));
    mbar_wait(&bar[par], (uint32_t)((it >> 1) & 1));
    for (int i = tid; i < SL * 128; i += THREADS) {
      const uint4 q = *reinterpret_cast<const uint4 *>(smem + (size_t)par * tile_bytes + (size_t)i * 16);
      acc += (q.x == (uint32_t)it) ? 1u : 1000000u;
    }
    if (delay) { long long t = clock64(); while (clock64() - t < delay) {} }
  }
  long long t1 = clock64();
  cluster_sync_all();
  out[blockIdx.x * THREADS + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  (void)lane; (void)warp;
}

// ---- (a) L2 tagged words, as lstm_tc_fwd_kernel: xbuf [2 parity][groups][16][C] words
__device__ __forceinline__ void st_word(uint32_t *p, uint32_t v) { asm volatile("st.relaxed.gpu.global.b32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 ld_word4(const uint4 *p) {
  uint4 q;
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
  return q;
}
__global__ void __launch_bounds__(THREADS, 1) l2_kernel(int iters, int SL, int delay, uint32_t *xbuf, unsigned *out, long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = blockIdx.x / SL, rank = blockIdx.x % SL, C = SL * CS, c8n = C / 8;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int par = it & 1;
    const uint32_t tag = (uint32_t)((it >> 1) & 1) << 16;
    uint32_t *xb = xbuf + ((size_t)par * gridDim.x / SL + group) * UG * C;
    for (int e = 0; e < 2; e++) st_word(xb + (size_t)(2 * warp + e) * C + rank * CS + lane, ((uint32_t)it & 0xfffeu) | tag | 0x3c00u << 0);
    const uint4 *xr = reinterpret_cast<const uint4 *>(xb);
    for (int i = 0; i < 4; i++) {
      const int v = tid + i * THREADS;
      if (v >= UG * c8n) break;
      uint4 q0, q1;
      do {
        q0 = ld_word4(xr + (size_t)v * 2); q1 = ld_word4(xr + (size_t)v * 2 + 1);
      } while ((((q0.x ^ tag) | (q0.y ^ tag) | (q0.z ^ tag) | (q0.w ^ tag) | (q1.x ^ tag) | (q1.y ^ tag) | (q1.z ^ tag) | (q1.w ^ tag)) & 0x10000u) != 0u);
      *reinterpret_cast<uint4 *>(smem + (size_t)v * 32) = q0;
      *reinterpret_cast<uint4 *>(smem + (size_t)v * 32 + 16) = q1;
      acc += q0.x & 1u;
    }
    __syncthreads();
    if (delay) { long long t = clock64(); while (clock64() - t < delay) {} }
  }
  long long t1 = clock64();
  out[blockIdx.x * THREADS + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  const size_t smem = 180 * 1024;
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(l2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 4000;
  unsigned *out; long long *cyc; uint32_t *xbuf;
  cudaMalloc(&out, sizeof(unsigned) * 148 * THREADS); cudaMalloc(&cyc, sizeof(long long) * 148);
  cudaMalloc(&xbuf, 4 * 2 * 16 * UG * 512);
  struct Cfg { int SL, nclusters; };
  const Cfg cfgs[] = {{10, 8}, {8, 8}, {4, 8}, {16, 8}, {12, 8}, {10, 4}, {5, 16}};
  for (const Cfg &c : cfgs) {
    const int blocks = c.SL * c.nclusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.SL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxc = -1;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&maxc, dsmem_kernel, &cfg);
    printf("== %d clusters of %d CTAs (256 threads, 180 KB): cudaOccupancyMaxActiveClusters = %d (%s)\n", c.nclusters, c.SL, maxc,
           cudaGetErrorString(qe));
    if (qe != cudaSuccess || maxc < c.nclusters) { cudaGetLastError(); printf("   (not co-resident: skipped)\n"); }
    for (int delay : {0, 2000}) {
      if (qe == cudaSuccess && maxc >= 1) {
        int SL = c.SL;
        cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel, iters, SL, delay, out, cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        unsigned ho[148 * THREADS]; cudaMemcpy(ho, out, sizeof(unsigned) * blocks * THREADS, cudaMemcpyDeviceToHost);
        double s = 0; long long mx = 0; for (int i = 0; i < blocks; i++) { s += h[i]; if (h[i] > mx) mx = h[i]; }
        unsigned long long tot = 0; for (int i = 0; i < blocks * THREADS; i++) tot += ho[i];
        printf("   DSMEM st.async + mbarrier, delay %4d: %.0f clk per step (max CTA %.0f)  payload %s  (%s / %s)\n", delay, s / blocks / iters,
               (double)mx / iters, tot == (unsigned long long)blocks * c.SL * 128 * iters ? "ok" : "WRONG", cudaGetErrorString(e), cudaGetErrorString(e2));
      }
      if (qe == cudaSuccess && maxc >= 1) {
        int SL = c.SL;
        cudaError_t e = cudaLaunchKernelEx(&cfg, bulk_kernel, iters, SL, delay, out, cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        unsigned ho[148 * THREADS]; cudaMemcpy(ho, out, sizeof(unsigned) * blocks * THREADS, cudaMemcpyDeviceToHost);
        double s = 0; long long mx = 0; for (int i = 0; i < blocks; i++) { s += h[i]; if (h[i] > mx) mx = h[i]; }
        unsigned long long tot = 0; for (int i = 0; i < blocks * THREADS; i++) tot += ho[i];
        printf("   DSMEM bulk copies (2 KB each), delay %4d: %.0f clk per step (max CTA %.0f)  payload %s  (%s / %s)\n", delay, s / blocks / iters,
               (double)mx / iters, tot == (unsigned long long)blocks * c.SL * 128 * iters ? "ok" : "WRONG", cudaGetErrorString(e), cudaGetErrorString(e2));
      }
      {
        cudaMemset(xbuf, 0xff, 4 * 2 * 16 * UG * 512);
        int SL = c.SL, it = iters;
        void *args[] = {&it, &SL, &delay, &xbuf, &out, &cyc};
        cudaError_t e = cudaLaunchCooperativeKernel((void *)l2_kernel, dim3(blocks), dim3(THREADS), args, smem, 0);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        double s = 0; long long mx = 0; for (int i = 0; i < blocks; i++) { s += h[i]; if (h[i] > mx) mx = h[i]; }
        printf("   L2 tagged words,           delay %4d: %.0f clk per step (max CTA %.0f)  (%s / %s)\n", delay, s / blocks / iters, (double)mx / iters,
               cudaGetErrorString(e), cudaGetErrorString(e2));
      }
    }
  }
  return 0;
}
