// Microbenchmark (test infrastructure): feasibility and latency of exchanging the per-step recurrent state of one
// (direction, utterance group) through distributed shared memory inside an 8-CTA thread-block cluster, at the
// resource footprint of lstm_fwd_kernel on C2 (640 threads, ~210 KB dynamic shared memory per CTA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o eesen_b200/bin/cluster_exchange tests/micro/cluster_exchange.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int CLUSTER = 8, THREADS = 640, CELLS = 40, UTTS = 8, C = 320;

// every CTA owns CELLS cells; per step it writes its [UTTS x CELLS] block of m into the staging buffer
// [UTTS x C] of all CLUSTER CTAs (its own included), then the cluster synchronises.
__global__ void __launch_bounds__(THREADS, 1) exchange_kernel(int iters, float *out, long long *cycles) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  float *stg = smem;   // [2 parity][UTTS][C]
  const int tid = threadIdx.x;
  float *peer[CLUSTER];
#pragma unroll
  for (int r = 0; r < CLUSTER; r++) peer[r] = cluster.map_shared_rank(stg, r);
  float acc = 0.f;
  cluster.sync();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    float *base_off = nullptr; (void)base_off;
    const int par = it & 1;
    if (tid < UTTS * CELLS) {
      const int u = tid / CELLS, c = tid % CELLS;
      const float v = (float)(it + rank) + 0.001f * tid;
#pragma unroll
      for (int r = 0; r < CLUSTER; r++) peer[r][(par * UTTS + u) * C + rank * CELLS + c] = v;
    }
    cluster.sync();                       // barrier.cluster.arrive.release + wait.acquire
    for (int i = tid; i < UTTS * C; i += THREADS) acc += stg[par * UTTS * C + i];
  }
  long long t1 = clock64();
  out[blockIdx.x * THREADS + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

// same data movement through global memory (L2) with a grid-wide generation counter per cluster: the baseline idea
__global__ void __launch_bounds__(THREADS, 1) l2_kernel(int iters, float *xbuf, unsigned *flags, float *out, long long *cycles) {
  extern __shared__ __align__(16) float smem[];
  const int cl = blockIdx.x / CLUSTER, rank = blockIdx.x % CLUSTER, tid = threadIdx.x;
  float *xb = xbuf + (size_t)cl * 2 * UTTS * C;
  unsigned *flag = flags + cl * 32;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int par = it & 1;
    if (tid < UTTS * CELLS) {
      const int u = tid / CELLS, c = tid % CELLS;
      __stcg(&xb[(par * UTTS + u) * C + rank * CELLS + c], (float)(it + rank) + 0.001f * tid);
    }
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicAdd(flag, 1u); while (*((volatile unsigned *)flag) < (unsigned)(CLUSTER * (it + 1))) {} }
    __syncthreads();
    for (int i = tid; i < UTTS * C; i += THREADS) acc += __ldcg(&xb[par * UTTS * C + i]);
  }
  long long t1 = clock64();
  out[blockIdx.x * THREADS + tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  (void)smem;
}

int main() {
  const size_t smem = 210 * 1024;
  cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(exchange_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(l2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int nclusters : {16, 18}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * CLUSTER); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxc = -1;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&maxc, exchange_kernel, &cfg);
    printf("grid of %d clusters of %d CTAs (640 threads, 210 KB): cudaOccupancyMaxActiveClusters = %d (%s)\n", nclusters, CLUSTER, maxc,
           cudaGetErrorString(qe));
  }
  int iters = 2000, nclusters = 16, blocks = nclusters * CLUSTER;
  float *out, *xbuf; long long *cyc; unsigned *flags;
  cudaMalloc(&out, sizeof(float) * blocks * THREADS); cudaMalloc(&cyc, sizeof(long long) * blocks);
  cudaMalloc(&xbuf, sizeof(float) * nclusters * 2 * UTTS * C); cudaMalloc(&flags, 4 * 32 * nclusters);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, exchange_kernel, iters, out, cyc);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[512]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
    printf("DSMEM exchange + cluster.sync, 16 clusters x 8 CTAs: %.0f clk per step (%s / %s)\n", c / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
  }
  {
    cudaMemset(flags, 0, 4 * 32 * nclusters);
    void *args[] = {&iters, &xbuf, &flags, &out, &cyc};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)l2_kernel, dim3(blocks), dim3(THREADS), args, smem, 0);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[512]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
    printf("L2 exchange + counter barrier,   16 groups   x 8 CTAs: %.0f clk per step (%s / %s)\n", c / iters, cudaGetErrorString(e), cudaGetErrorString(e2));
  }
  return 0;
}
