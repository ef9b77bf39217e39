// eesen_b200/host/context.h -- per-process device context: replaces the reference's singleton
// CuDevice (src/gpucompute/cuda-device.{h,cc}) with {device, one compute stream, grow-only
// workspace arenas, optional NCCL communicator}.  No per-call device synchronisation
// (the reference syncs after every kernel, cuda-common.h:37-44).
#ifndef EESEN_B200_HOST_CONTEXT_H_
#define EESEN_B200_HOST_CONTEXT_H_

#include <cuda_runtime.h>

#include <string>

#include "../csrc/kernels.h"

struct eesen_b200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 0;
  size_t max_smem = 0;
  int gemm_prec = 0, rec_prec = 0;
  std::string err;
  long launches = 0;

  struct Buf {
    void *p = nullptr;
    size_t bytes = 0;
  };
  Buf gemm_ws, lstm_pbuf, lstm_gsum, lstm_flags, ctc_ws, colsum_ws, seg_buf;

  // NCCL (dlopen'ed on demand)
  void *nccl_lib = nullptr;
  void *nccl_comm = nullptr;
  int rank = 0, nranks = 1;

  int fail(int code, const std::string &msg) {
    err = msg;
    return code;
  }
  int check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return (int)e;
  }
  // grow-only scratch; stream-ordered reuse is safe because everything runs on `stream`
  int reserve(Buf &b, size_t bytes, void **out) {
    if (bytes > b.bytes) {
      if (b.p) {
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) return check(e, "cudaStreamSynchronize");
        cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
      }
      size_t want = bytes + bytes / 8 + 256;
      cudaError_t e = cudaMalloc(&b.p, want);
      if (e != cudaSuccess) return check(e, "cudaMalloc(workspace)");
      b.bytes = want;
    }
    *out = b.p;
    return 0;
  }
};

#endif
