// Microbenchmark (test infrastructure): issue rate of the warp-level tensor instructions the recurrent kernels
// can use on sm_100a -- mma.sync m16n8k8 tf32 vs m16n8k16 bf16 -- from registers only (no memory traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o eesen_b200/bin/hmma_rate tests/micro/hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int NACC>
__global__ void rate_kernel(int iters, float *out, long long *cycles) {
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  unsigned a0 = threadIdx.x * 0x9e3779b9u, a1 = a0 ^ 0x1234567u, a2 = a0 + 77u, a3 = a1 + 99u;
  unsigned b0 = a0 * 3u, b1 = a1 * 5u;
  a0 &= 0x3f7fffffu; a1 &= 0x3f7fffffu; a2 &= 0x3f7fffffu; a3 &= 0x3f7fffffu; b0 &= 0x3f7fffffu; b1 &= 0x3f7fffffu;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      if (KIND == 0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND, int NACC>
void run(const char *name, int threads, int macs_per_instr) {
  int iters = 2000, blocks = 148;
  float *out; long long *cyc;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  rate_kernel<KIND, NACC><<<blocks, threads>>>(10, out, cyc);
  rate_kernel<KIND, NACC><<<blocks, threads>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < blocks; i++) c += h[i]; c /= blocks;
  double instr = (double)iters * NACC * (threads / 32);
  printf("%-28s warps/SM %2d chains %d : %.2f clk per warp-instr per SM, %.0f MAC/clk/SM\n", name, threads / 32, NACC,
         c / instr, instr * macs_per_instr / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 8>("m16n8k8  tf32", 640, 1024);
  run<1, 8>("m16n8k16 bf16", 640, 2048);
  run<0, 8>("m16n8k8  tf32", 128, 1024);
  run<1, 8>("m16n8k16 bf16", 128, 2048);
  run<0, 2>("m16n8k8  tf32", 640, 1024);
  run<1, 2>("m16n8k16 bf16", 640, 2048);
  run<0, 1>("m16n8k8  tf32 (dependent)", 32, 1024);
  run<1, 1>("m16n8k16 bf16 (dependent)", 32, 2048);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
