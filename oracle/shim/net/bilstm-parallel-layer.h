// oracle/shim/net/bilstm-parallel-layer.h -- the reference-side binding of INTEGRATION.md section 2, compiled.
//
// TEST INFRASTRUCTURE (oracle/): this header SHADOWS the reference's src/net/bilstm-parallel-layer.h on the include
// path when the reference's own, unmodified src/net/layer.cc is compiled for `ref_train_ctc_parallel_b200`
// (oracle/Makefile).  It declares the same class -- same name, marker <BiLstmParallel> / type l_BiLstm_Parallel
// (layer.cc:37-106), same base class BiLstm with its parameters, ReadData/WriteData and Update (bilstm-layer.h) --
// and re-points the two virtuals of the hot path at the C ABI of libeesen_b200.so:
//   PropagateFnc      (reference bilstm-parallel-layer.h:379-420)  -> eesen_b200_bilstm_forward
//   BackpropagateFnc  (reference :881-913, :422-602)               -> eesen_b200_bilstm_backward
// The reference's CuMatrix weights are bound in place through their pitch (ldwx / ldwm of the ABI structs); the
// raw gradient sums come back into scratch matrices and are folded into the reference's momentum-carrying
// *_corr_ buffers exactly as its AddMatMat(..., beta = momentum) calls do (:504-510,596-601):
// corr = momentum * corr + grad.  BiLstm::Update (bilstm-layer.h:846-956) then runs unchanged.
// Covers the vanilla path (no dropout), which is what train-ctc-parallel uses with the stock recipes.
#ifndef EESEN_BILSTM_PARALLEL_LAYER_H_
#define EESEN_BILSTM_PARALLEL_LAYER_H_

#include <cuda_runtime.h>

#include "eesen_b200.h"
#include "net/bilstm-layer.h"
#include "net/layer.h"
#include "net/trainable-layer.h"

namespace eesen {

// one library context per process, on the device the reference's CuDevice selected
eesen_b200_ctx *B200Context();
void B200Check(int rc, const char *what);

// CuMatrixBase::Data() is protected in the reference (cuda-matrix.h:296-311) -- a maintainer would add a public
// accessor; this unmodified-reference build reaches it through a member-less derived class.
struct B200MatrixAccess : public CuMatrixBase<BaseFloat> {
  static const BaseFloat *Ptr(const CuMatrixBase<BaseFloat> &m) { return static_cast<const B200MatrixAccess &>(m).Data(); }
  static BaseFloat *Ptr(CuMatrixBase<BaseFloat> *m) { return static_cast<B200MatrixAccess *>(m)->Data(); }
};

class B200DeviceBuffer {
 public:
  B200DeviceBuffer() : p_(NULL), bytes_(0) {}
  ~B200DeviceBuffer() { if (p_) cudaFree(p_); }
  B200DeviceBuffer(const B200DeviceBuffer &) : p_(NULL), bytes_(0) {}   // Layer::Copy(): scratch is not shared
  B200DeviceBuffer &operator=(const B200DeviceBuffer &) { return *this; }
  void *Reserve(size_t bytes) {
    if (bytes > bytes_) {
      if (p_) cudaFree(p_);
      if (cudaMalloc(&p_, bytes) != cudaSuccess) KALDI_ERR << "cudaMalloc of " << bytes << " bytes failed";
      bytes_ = bytes;
    }
    return p_;
  }
 private:
  void *p_;
  size_t bytes_;
};

class BiLstmParallel : public BiLstm {
 public:
  BiLstmParallel(int32 input_dim, int32 output_dim) : BiLstm(input_dim, output_dim) {}
  ~BiLstmParallel() {}

  Layer *Copy() const { return new BiLstmParallel(*this); }
  LayerType GetType() const { return l_BiLstm_Parallel; }
  LayerType GetTypeNonParal() const { return l_BiLstm; }

  void SetSeqLengths(std::vector<int> &sequence_lengths) { sequence_lengths_ = sequence_lengths; }

  void PropagateFnc(const CuMatrixBase<BaseFloat> &in, CuMatrixBase<BaseFloat> *out) {
    const int32 S = sequence_lengths_.size();
    KALDI_ASSERT(S > 0 && in.NumRows() % S == 0);
    const int32 T = in.NumRows() / S, C = cell_dim_;
    CheckVanilla();
    int *d_len = static_cast<int *>(len_.Reserve(sizeof(int) * S));
    cudaMemcpy(d_len, sequence_lengths_.data(), sizeof(int) * S, cudaMemcpyHostToDevice);
    float *gates = static_cast<float *>(gates_.Reserve(sizeof(float) * (size_t)T * S * 8 * C));
    float *cell = static_cast<float *>(cell_.Reserve(sizeof(float) * (size_t)T * S * 2 * C));
    eesen_b200_bilstm_params p;
    Params(&p);
    cudaDeviceSynchronize();   // the reference works on the legacy default stream, the library on its own
    B200Check(eesen_b200_bilstm_forward(B200Context(), T, S, input_dim_, C, d_len, B200MatrixAccess::Ptr(in), in.Stride(), &p, gates,
                                        cell, B200MatrixAccess::Ptr(out), out->Stride()), "eesen_b200_bilstm_forward");
    B200Check(eesen_b200_synchronize(B200Context()), "eesen_b200_synchronize");
  }

  void BackpropagateFnc(const CuMatrixBase<BaseFloat> &in, const CuMatrixBase<BaseFloat> &out,
                        const CuMatrixBase<BaseFloat> &out_diff, CuMatrixBase<BaseFloat> *in_diff) {
    if (!in_train) KALDI_ERR << "Can't backpropagate in test mode";
    const int32 S = sequence_lengths_.size();
    const int32 T = in.NumRows() / S, C = cell_dim_;
    float *dgates = static_cast<float *>(dgates_.Reserve(sizeof(float) * (size_t)T * S * 8 * C));
    // raw gradient sums of this minibatch (same shapes, hence the same pitch, as the weights)
    g_wx_fw_.Resize(4 * C, input_dim_, kUndefined); g_wx_bw_.Resize(4 * C, input_dim_, kUndefined);
    g_wm_fw_.Resize(4 * C, C, kUndefined); g_wm_bw_.Resize(4 * C, C, kUndefined);
    g_b_fw_.Resize(4 * C, kUndefined); g_b_bw_.Resize(4 * C, kUndefined);
    g_pi_fw_.Resize(C, kUndefined); g_pf_fw_.Resize(C, kUndefined); g_po_fw_.Resize(C, kUndefined);
    g_pi_bw_.Resize(C, kUndefined); g_pf_bw_.Resize(C, kUndefined); g_po_bw_.Resize(C, kUndefined);
    eesen_b200_bilstm_params p;
    Params(&p);
    eesen_b200_bilstm_grads g;
    g.wx[0] = B200MatrixAccess::Ptr(&g_wx_fw_); g.wx[1] = B200MatrixAccess::Ptr(&g_wx_bw_); g.wm[0] = B200MatrixAccess::Ptr(&g_wm_fw_); g.wm[1] = B200MatrixAccess::Ptr(&g_wm_bw_);
    g.bias[0] = g_b_fw_.Data(); g.bias[1] = g_b_bw_.Data();
    g.pi[0] = g_pi_fw_.Data(); g.pi[1] = g_pi_bw_.Data(); g.pf[0] = g_pf_fw_.Data(); g.pf[1] = g_pf_bw_.Data();
    g.po[0] = g_po_fw_.Data(); g.po[1] = g_po_bw_.Data();
    KALDI_ASSERT(g_wx_fw_.Stride() == g_wx_bw_.Stride() && g_wm_fw_.Stride() == g_wm_bw_.Stride());
    g.ldwx = g_wx_fw_.Stride(); g.ldwm = g_wm_fw_.Stride();
    cudaDeviceSynchronize();
    B200Check(eesen_b200_bilstm_backward(B200Context(), T, S, input_dim_, C, B200MatrixAccess::Ptr(in), in.Stride(), &p,
                                         static_cast<const float *>(gates_.Reserve(0)),
                                         static_cast<const float *>(cell_.Reserve(0)), B200MatrixAccess::Ptr(out), out.Stride(),
                                         B200MatrixAccess::Ptr(out_diff), out_diff.Stride(), dgates, B200MatrixAccess::Ptr(in_diff), in_diff->Stride(),
                                         &g), "eesen_b200_bilstm_backward");
    B200Check(eesen_b200_synchronize(B200Context()), "eesen_b200_synchronize");
    // corr = momentum * corr + grad   (reference :504-510, :596-601 do it through AddMatMat's beta)
    const BaseFloat mmt = opts_.momentum;
    Fold(&wei_gifo_x_fw_corr_, g_wx_fw_, mmt); Fold(&wei_gifo_x_bw_corr_, g_wx_bw_, mmt);
    Fold(&wei_gifo_m_fw_corr_, g_wm_fw_, mmt); Fold(&wei_gifo_m_bw_corr_, g_wm_bw_, mmt);
    Fold(&bias_fw_corr_, g_b_fw_, mmt); Fold(&bias_bw_corr_, g_b_bw_, mmt);
    Fold(&phole_i_c_fw_corr_, g_pi_fw_, mmt); Fold(&phole_i_c_bw_corr_, g_pi_bw_, mmt);
    Fold(&phole_f_c_fw_corr_, g_pf_fw_, mmt); Fold(&phole_f_c_bw_corr_, g_pf_bw_, mmt);
    Fold(&phole_o_c_fw_corr_, g_po_fw_, mmt); Fold(&phole_o_c_bw_corr_, g_po_bw_, mmt);
  }

 private:
  void CheckVanilla() const {
    if (in_train && (forward_dropout > 0.0 || ((rnndrop || no_mem_loss_dropout) && recurrent_dropout > 0.0)))
      KALDI_ERR << "the compiled reference-side binding covers the vanilla (no dropout) path";
  }
  void Params(eesen_b200_bilstm_params *p) const {
    p->wx[0] = B200MatrixAccess::Ptr(wei_gifo_x_fw_); p->wx[1] = B200MatrixAccess::Ptr(wei_gifo_x_bw_);
    p->wm[0] = B200MatrixAccess::Ptr(wei_gifo_m_fw_); p->wm[1] = B200MatrixAccess::Ptr(wei_gifo_m_bw_);
    p->bias[0] = bias_fw_.Data(); p->bias[1] = bias_bw_.Data();
    p->pi[0] = phole_i_c_fw_.Data(); p->pi[1] = phole_i_c_bw_.Data();
    p->pf[0] = phole_f_c_fw_.Data(); p->pf[1] = phole_f_c_bw_.Data();
    p->po[0] = phole_o_c_fw_.Data(); p->po[1] = phole_o_c_bw_.Data();
    KALDI_ASSERT(wei_gifo_x_fw_.Stride() == wei_gifo_x_bw_.Stride() && wei_gifo_m_fw_.Stride() == wei_gifo_m_bw_.Stride());
    p->ldwx = wei_gifo_x_fw_.Stride();   // cudaMallocPitch'ed (cuda-matrix.cc:46-79): bound without repacking
    p->ldwm = wei_gifo_m_fw_.Stride();
  }
  static void Fold(CuMatrix<BaseFloat> *corr, const CuMatrix<BaseFloat> &grad, BaseFloat mmt) {
    corr->Scale(mmt);
    corr->AddMat(1.0, grad);
  }
  static void Fold(CuVector<BaseFloat> *corr, const CuVector<BaseFloat> &grad, BaseFloat mmt) {
    corr->Scale(mmt);
    corr->AddVec(1.0, grad);
  }

  std::vector<int> sequence_lengths_;   // reference :918
  B200DeviceBuffer len_, gates_, cell_, dgates_;
  CuMatrix<BaseFloat> g_wx_fw_, g_wx_bw_, g_wm_fw_, g_wm_bw_;
  CuVector<BaseFloat> g_b_fw_, g_b_bw_, g_pi_fw_, g_pf_fw_, g_po_fw_, g_pi_bw_, g_pf_bw_, g_po_bw_;
};

}  // namespace eesen

#endif
