// oracle/shim/b200_wrap.cc -- link-time re-pointing of the reference's call sites at the C ABI (TEST
// INFRASTRUCTURE: proves that INTEGRATION.md's binding compiles and runs against the UNMODIFIED reference).
//
// The reference objects are built from /root/reference as they are; `ld --wrap=<symbol>` makes every call of
//   Ctc::EvalParallel              (src/net/ctc-loss.cc:101-194; caller src/netbin/train-ctc-parallel.cc:199)
//   CuMatrixBase<float>::AddMatMat (src/gpucompute/cuda-matrix.cc:603-639 -> cublasSgemm; callers
//                                   src/net/affine-trans-layer.h:161-219)
// from another object land in the functions below, which forward to eesen_b200_ctc_eval / eesen_b200_gemm.
// BiLstmParallel::{PropagateFnc,BackpropagateFnc} are re-pointed by the shadow header net/bilstm-parallel-layer.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>

#include "eesen_b200.h"
#define private public     // the registries of Ctc (obj_progress_, sequences_num_, ...) are private members,
#define protected public   // CuMatrixBase::Data() is protected (cuda-matrix.h:296-311)
#include "gpucompute/cuda-matrix.h"
#include "net/ctc-loss.h"
#undef private
#undef protected

namespace eesen {

eesen_b200_ctx *B200Context() {
  static eesen_b200_ctx *ctx = NULL;
  if (!ctx) {
    int dev = 0;
    cudaGetDevice(&dev);   // the device CuDevice::SelectGpuId picked (train-ctc-parallel.cc:106)
    int rc = eesen_b200_create(&ctx, dev);
    if (rc) KALDI_ERR << "eesen_b200_create failed: " << eesen_b200_last_error(NULL);
  }
  return ctx;
}

void B200Check(int rc, const char *what) {
  if (rc) KALDI_ERR << what << " failed (" << rc << "): " << eesen_b200_last_error(B200Context());
}

}  // namespace eesen

using namespace eesen;

extern "C" {

// ---- Ctc::EvalParallel(frame_num_utt, net_out, label, diff)
void __wrap__ZN5eesen3Ctc12EvalParallelERKSt6vectorIiSaIiEERKNS_12CuMatrixBaseIfEERS1_IS3_SaIS3_EEPNS_8CuMatrixIfEE(
    Ctc *self, const std::vector<int32> &frame_num_utt, const CuMatrixBase<BaseFloat> &net_out,
    std::vector<std::vector<int32> > &label, CuMatrix<BaseFloat> *diff) {
  diff->Resize(net_out.NumRows(), net_out.NumCols());
  const int32 S = frame_num_utt.size(), rows = net_out.NumRows();
  KALDI_ASSERT(S > 0 && rows % S == 0);
  const int32 T = rows / S, K = net_out.NumCols();
  int32 max_lab = 1;
  for (int32 s = 0; s < S; s++) max_lab = std::max<int32>(max_lab, label[s].size());
  // [len S][lablen S][labels S x max_lab] in one upload (the reference uploads the expanded label matrix on
  // each of its 2T kernel launches, cuda-matrix.cc:882-883,948-950)
  std::vector<int32> h((size_t)2 * S + (size_t)S * max_lab, 0);
  for (int32 s = 0; s < S; s++) {
    h[s] = frame_num_utt[s];
    h[S + s] = label[s].size();
    for (size_t l = 0; l < label[s].size(); l++) h[2 * S + (size_t)s * max_lab + l] = label[s][l];
  }
  int32 *d = NULL;
  float *d_pzx = NULL;
  cudaMalloc((void **)&d, sizeof(int32) * h.size());
  cudaMalloc((void **)&d_pzx, sizeof(float) * S);
  cudaMemcpy(d, h.data(), sizeof(int32) * h.size(), cudaMemcpyHostToDevice);
  cudaDeviceSynchronize();
  B200Check(eesen_b200_ctc_eval(B200Context(), T, S, K, max_lab, d, d + 2 * S, d + S, net_out.Data(), net_out.Stride(),
                                d_pzx, diff->Data(), diff->Stride()), "eesen_b200_ctc_eval");
  B200Check(eesen_b200_synchronize(B200Context()), "eesen_b200_synchronize");
  std::vector<float> pzx(S);
  cudaMemcpy(pzx.data(), d_pzx, sizeof(float) * S, cudaMemcpyDeviceToHost);
  cudaFree(d);
  cudaFree(d_pzx);
  // registries and progressive report, as ctc-loss.cc:170-192
  double sum = 0.0;
  for (int32 s = 0; s < S; s++) sum += pzx[s];
  self->obj_progress_ += sum;
  self->sequences_progress_ += S;
  self->sequences_num_ += S;
  for (int32 s = 0; s < S; s++) {
    self->frames_progress_ += frame_num_utt[s];
    self->frames_ += frame_num_utt[s];
  }
  if (self->sequences_progress_ >= self->report_step_) {
    KALDI_VLOG(1) << "After " << self->sequences_num_ << " sequences (" << self->frames_ / (100.0 * 3600) << "Hr): "
                  << "Obj(log[Pzx]) = " << self->obj_progress_ / self->sequences_progress_ << "   TokenAcc = "
                  << 100.0 * (1.0 - self->error_num_progress_ / self->ref_num_progress_) << "%";
    self->sequences_progress_ = 0;
    self->frames_progress_ = 0;
    self->obj_progress_ = 0.0;
    self->error_num_progress_ = 0;
    self->ref_num_progress_ = 0;
  }
}

// ---- CuMatrixBase<float>::AddMatMat(alpha, A, transA, B, transB, beta)
void __real__ZN5eesen12CuMatrixBaseIfE9AddMatMatEfRKS1_NS_19MatrixTransposeTypeES3_S4_f(
    CuMatrixBase<float> *self, float alpha, const CuMatrixBase<float> &A, MatrixTransposeType transA,
    const CuMatrixBase<float> &B, MatrixTransposeType transB, float beta);

void __wrap__ZN5eesen12CuMatrixBaseIfE9AddMatMatEfRKS1_NS_19MatrixTransposeTypeES3_S4_f(
    CuMatrixBase<float> *self, float alpha, const CuMatrixBase<float> &A, MatrixTransposeType transA,
    const CuMatrixBase<float> &B, MatrixTransposeType transB, float beta) {
  const int ta = transA == kTrans, tb = transB == kTrans;
  if ((ta && tb) || self->NumRows() == 0 || self->NumCols() == 0) {   // (T,T) is not on the path: the reference's own
    __real__ZN5eesen12CuMatrixBaseIfE9AddMatMatEfRKS1_NS_19MatrixTransposeTypeES3_S4_f(self, alpha, A, transA, B, transB, beta);
    return;
  }
  const int M = self->NumRows(), N = self->NumCols(), K = ta ? A.NumRows() : A.NumCols();
  KALDI_ASSERT((ta ? A.NumCols() : A.NumRows()) == M && (tb ? B.NumRows() : B.NumCols()) == N &&
               (tb ? B.NumCols() : B.NumRows()) == K);
  cudaDeviceSynchronize();
  B200Check(eesen_b200_gemm(B200Context(), ta, tb, M, N, K, alpha, A.Data(), A.Stride(), B.Data(), B.Stride(), beta,
                            self->Data(), self->Stride()), "eesen_b200_gemm");
  B200Check(eesen_b200_synchronize(B200Context()), "eesen_b200_synchronize");
}

}  // extern "C"
