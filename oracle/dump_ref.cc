// oracle/dump_ref.cc -- TEST INFRASTRUCTURE ONLY.
//
// A small driver (our own code) that links against the UNMODIFIED reference
// objects compiled in place from /root/reference/src (see oracle/Makefile) and
// runs the reference's own public API on one packed minibatch:
//     Net::Read -> SetTrainOptions -> SetSeqLengths -> Propagate ->
//     Ctc::EvalParallel -> Ctc::ErrorRateMSeq -> Net::Backpropagate (incl. Update)
// exactly as src/netbin/train-ctc-parallel.cc:195-207 does, then dumps every
// tensor the parity tests compare against (as .npy files):
//     out_l<i>.npy        Net::propagate_buf_[i]          (layer outputs, i=0 is the input)
//     net_out.npy         softmax probabilities [T*S x K]
//     alpha.npy/beta.npy  Ctc::alpha_/beta_ (GPU build only; zeros on CPU)
//     pzx.npy             per-utterance log p(z|x)  (recomputed as ctc-loss.cc:146-153)
//     obj_diff.npy        d(-log p)/d(logits) [T*S x K]
//     in_diff.npy         gradient wrt the network input
//     grad_<layer>_<name>.npy   the *_corr_ buffers after Update (momentum + clip applied)
//     <outdir>/model_out  the updated model (Net::Write, binary)
// Built twice: ref_dump_cpu (cpucompute; CTC is a no-op there, so --diff-in must
// supply obj_diff) and ref_dump_gpu (-DHAVE_CUDA, the reference's own CUDA kernels
// on the B200: the authoritative oracle).
//
// Access to private/protected members uses the "#define private public" shim on the
// reference headers only (standard headers are included first).

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <limits>
#include <list>
#include <utility>
#include <stdexcept>
#include <iomanip>
#include <typeinfo>

#define private public
#define protected public
#include "net/train-opts.h"
#include "net/net.h"
#include "net/layer.h"
#include "net/bilstm-layer.h"
#include "net/bilstm-parallel-layer.h"
#include "net/lstm-layer.h"
#include "net/lstm-parallel-layer.h"
#include "net/affine-trans-layer.h"
#include "net/ctc-loss.h"
#undef private
#undef protected
#include "base/kaldi-common.h"
#include "base/timer.h"
#include "gpucompute/cuda-device.h"

using namespace eesen;

static void write_npy(const std::string &path, const void *data, const char *descr,
                      size_t elsize, const std::vector<size_t> &shape) {
  std::ostringstream hdr;
  hdr << "{'descr': '" << descr << "', 'fortran_order': False, 'shape': (";
  size_t n = 1;
  for (size_t i = 0; i < shape.size(); i++) { hdr << shape[i] << ","; n *= shape[i]; }
  hdr << "), }";
  std::string h = hdr.str();
  size_t total = 10 + h.size() + 1;
  size_t pad = (64 - total % 64) % 64;
  h.append(pad, ' ');
  h.push_back('\n');
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) { perror(path.c_str()); exit(2); }
  const unsigned char magic[8] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0};
  fwrite(magic, 1, 8, f);
  unsigned short hl = (unsigned short)h.size();
  fwrite(&hl, 2, 1, f);
  fwrite(h.data(), 1, h.size(), f);
  fwrite(data, elsize, n, f);
  fclose(f);
}

static void dump_mat(const std::string &path, const CuMatrixBase<BaseFloat> &m) {
  Matrix<BaseFloat> h(m.NumRows(), m.NumCols());
  m.CopyToMat(&h);
  std::vector<float> flat((size_t)h.NumRows() * h.NumCols());
  for (int r = 0; r < h.NumRows(); r++)
    memcpy(&flat[(size_t)r * h.NumCols()], h.RowData(r), sizeof(float) * h.NumCols());
  write_npy(path, flat.data(), "<f4", 4, {(size_t)h.NumRows(), (size_t)h.NumCols()});
}

static void dump_vec(const std::string &path, const CuVectorBase<BaseFloat> &v) {
  Vector<BaseFloat> h(v.Dim());
  v.CopyToVec(&h);
  write_npy(path, h.Data(), "<f4", 4, {(size_t)h.Dim()});
}

struct Batch {
  int S, T, I;
  std::vector<int> frames, label_len;
  std::vector<std::vector<int> > labels;
  Matrix<BaseFloat> feats;  // [T*S x I], row = t*S+s
};

// batch file: int32 magic 0x45534E42, S, T, I; int32 frames[S]; int32 label_len[S];
// int32 labels (concatenated); float32 feats[T*S*I] (time-major interleaved, zero padded)
static void read_batch(const char *path, Batch *b) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  int hdr[4];
  if (fread(hdr, 4, 4, f) != 4 || hdr[0] != 0x45534E42) { fprintf(stderr, "bad batch file\n"); exit(2); }
  b->S = hdr[1]; b->T = hdr[2]; b->I = hdr[3];
  b->frames.resize(b->S); b->label_len.resize(b->S); b->labels.resize(b->S);
  if (fread(b->frames.data(), 4, b->S, f) != (size_t)b->S) exit(2);
  if (fread(b->label_len.data(), 4, b->S, f) != (size_t)b->S) exit(2);
  for (int s = 0; s < b->S; s++) {
    b->labels[s].resize(b->label_len[s]);
    if (b->label_len[s] && fread(b->labels[s].data(), 4, b->label_len[s], f) != (size_t)b->label_len[s]) exit(2);
  }
  b->feats.Resize(b->T * b->S, b->I);
  std::vector<float> row(b->I);
  for (int r = 0; r < b->T * b->S; r++) {
    if (fread(row.data(), 4, b->I, f) != (size_t)b->I) exit(2);
    memcpy(b->feats.RowData(r), row.data(), 4 * b->I);
  }
  fclose(f);
}

static void read_npy_f32(const char *path, Matrix<BaseFloat> *m, int rows, int cols) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  unsigned char pre[10];
  if (fread(pre, 1, 10, f) != 10) exit(2);
  int hl = pre[8] | (pre[9] << 8);
  fseek(f, 10 + hl, SEEK_SET);
  m->Resize(rows, cols);
  std::vector<float> row(cols);
  for (int r = 0; r < rows; r++) {
    if (fread(row.data(), 4, cols, f) != (size_t)cols) { fprintf(stderr, "short diff file\n"); exit(2); }
    memcpy(m->RowData(r), row.data(), 4 * cols);
  }
  fclose(f);
}

int main(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s <model-in> <batch.bin> <outdir> [--lr x] [--momentum m] [--steps n]"
                    " [--diff-in obj_diff.npy] [--time-only] [--no-dump-layers]"
                    " [--opt SGD|Adagrad|RMSProp] [--eps e] [--rho r]\n", argv[0]);
    return 1;
  }
  std::string model = argv[1], batch_path = argv[2], outdir = argv[3];
  float lr = 0.0f, momentum = 0.0f;
  int steps = 1;
  std::string diff_in, opt = "SGD";
  float eps = -1.f, rho = -1.f;
  bool time_only = false, dump_layers = true;
  for (int i = 4; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--lr") lr = atof(argv[++i]);
    else if (a == "--momentum") momentum = atof(argv[++i]);
    else if (a == "--steps") steps = atoi(argv[++i]);
    else if (a == "--diff-in") diff_in = argv[++i];
    else if (a == "--opt") opt = argv[++i];
    else if (a == "--eps") eps = atof(argv[++i]);
    else if (a == "--rho") rho = atof(argv[++i]);
    else if (a == "--time-only") time_only = true;
    else if (a == "--no-dump-layers") dump_layers = false;
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 1; }
  }
  try {
#if HAVE_CUDA == 1
    CuDevice::Instantiate().SelectGpuId("yes");
    CuDevice::Instantiate().DisableCaching();
#endif
    Batch b;
    read_batch(batch_path.c_str(), &b);

    Net net;
    net.Read(model);
    NetTrainOptions opts;
    opts.learn_rate = lr;
    opts.momentum = momentum;
    if (eps >= 0.f) opts.adagrad_epsilon = eps;
    if (rho >= 0.f) opts.rmsprop_rho = rho;   // rmsprop_one_minus_rho stays 0.1, as through the driver's option parser
    net.SetTrainOptions(opts);
    net.SetUpdateAlgorithm(opt);
    net.SetTrainMode();

    Ctc ctc;
    ctc.SetReportStep(1000000);
    CuMatrix<BaseFloat> net_out, obj_diff, in_diff;
    std::string no_out_file = "";

    Timer timer;
    std::vector<double> step_seconds;
    for (int step = 0; step < steps; step++) {
      Timer st;
      net.SetSeqLengths(b.frames);
      net.Propagate(CuMatrix<BaseFloat>(b.feats), &net_out);
      ctc.EvalParallel(b.frames, net_out, b.labels, &obj_diff);
      ctc.ErrorRateMSeq(b.frames, net_out, b.labels, no_out_file);
      if (!diff_in.empty()) {
        Matrix<BaseFloat> d;
        read_npy_f32(diff_in.c_str(), &d, net_out.NumRows(), net_out.NumCols());
        obj_diff.Resize(d.NumRows(), d.NumCols());
        obj_diff.CopyFromMat(d);
      }
      if (step == steps - 1 && !time_only) {
        // forward-side dumps are taken before the update of the LAST step
        if (dump_layers)
          for (size_t i = 0; i < net.propagate_buf_.size(); i++) {
            std::ostringstream p; p << outdir << "/out_l" << i << ".npy";
            dump_mat(p.str(), net.propagate_buf_[i]);
          }
        // dropout masks drawn by this Propagate (scaled 0 | 1/(1-p)); shapes as the reference holds them:
        // forward [T*S x 2C]; recurrent fw/bw [(T+2)*S x C] (step) or [S x C] (sequence)
        for (int l = 0; l < net.NumLayers(); l++) {
          if (net.layers_[l]->GetType() != Layer::l_BiLstm_Parallel) continue;
          BiLstm *bl = dynamic_cast<BiLstm*>(net.layers_[l]);
          std::ostringstream pre; pre << outdir << "/mask_" << l << "_";
          if (bl->forward_drop_mask_.NumRows() > 0) dump_mat(pre.str() + "fwd.npy", bl->forward_drop_mask_);
          if (bl->recurrent_drop_mask_fw_.NumRows() > 0) {
            dump_mat(pre.str() + "rec_fw.npy", bl->recurrent_drop_mask_fw_);
            dump_mat(pre.str() + "rec_bw.npy", bl->recurrent_drop_mask_bw_);
          }
        }
        dump_mat(outdir + "/net_out.npy", net_out);
        dump_mat(outdir + "/obj_diff.npy", obj_diff);
        if (ctc.alpha_.NumRows() > 0) {
          dump_mat(outdir + "/alpha.npy", ctc.alpha_);
          dump_mat(outdir + "/beta.npy", ctc.beta_);
          // pzx as in ctc-loss.cc:146-153
          Matrix<BaseFloat> a(ctc.alpha_.NumRows(), ctc.alpha_.NumCols());
          ctc.alpha_.CopyToMat(&a);
          std::vector<float> pzx(b.S);
          for (int s = 0; s < b.S; s++) {
            int L = 2 * b.labels[s].size() + 1, fn = b.frames[s];
            float t1 = a((fn - 1) * b.S + s, L - 1);
            float t2 = L >= 2 ? a((fn - 1) * b.S + s, L - 2) : -1e30f;
            float e = (t2 - t1) <= -1e30f ? 0.f : expf(t2 - t1);
            pzx[s] = t1 + logf(1 + e);
          }
          write_npy(outdir + "/pzx.npy", pzx.data(), "<f4", 4, {(size_t)b.S});
        }
      }
      net.Backpropagate(obj_diff, &in_diff);
      step_seconds.push_back(st.Elapsed());
    }
    double elapsed = timer.Elapsed();
    long valid = 0;
    for (int s = 0; s < b.S; s++) valid += b.frames[s];
    fprintf(stdout, "{\"steps\": %d, \"seconds\": %.6f, \"valid_frames_per_step\": %ld, "
                    "\"padded_frames_per_step\": %d, \"valid_fps\": %.3f, \"token_err\": %.1f, \"ref_tokens\": %d, "
                    "\"step_seconds\": [",
            steps, elapsed, valid, b.T * b.S, valid * steps / elapsed,
            ctc.NumErrorTokens(), ctc.NumRefTokens());
    for (size_t i = 0; i < step_seconds.size(); i++) fprintf(stdout, "%s%.6f", i ? ", " : "", step_seconds[i]);
    fprintf(stdout, "]}\n");
    if (time_only) return 0;

    dump_mat(outdir + "/in_diff.npy", in_diff);
    for (int l = 0; l < net.NumLayers(); l++) {
      std::ostringstream pre; pre << outdir << "/grad_" << l << "_";
      Layer *ly = net.layers_[l];
      if (ly->GetType() == Layer::l_BiLstm_Parallel || ly->GetType() == Layer::l_BiLstm) {
        BiLstm *bl = dynamic_cast<BiLstm*>(ly);
        dump_mat(pre.str() + "wx_fw.npy", bl->wei_gifo_x_fw_corr_);
        dump_mat(pre.str() + "wm_fw.npy", bl->wei_gifo_m_fw_corr_);
        dump_vec(pre.str() + "b_fw.npy", bl->bias_fw_corr_);
        dump_vec(pre.str() + "pi_fw.npy", bl->phole_i_c_fw_corr_);
        dump_vec(pre.str() + "pf_fw.npy", bl->phole_f_c_fw_corr_);
        dump_vec(pre.str() + "po_fw.npy", bl->phole_o_c_fw_corr_);
        dump_mat(pre.str() + "wx_bw.npy", bl->wei_gifo_x_bw_corr_);
        dump_mat(pre.str() + "wm_bw.npy", bl->wei_gifo_m_bw_corr_);
        dump_vec(pre.str() + "b_bw.npy", bl->bias_bw_corr_);
        dump_vec(pre.str() + "pi_bw.npy", bl->phole_i_c_bw_corr_);
        dump_vec(pre.str() + "pf_bw.npy", bl->phole_f_c_bw_corr_);
        dump_vec(pre.str() + "po_bw.npy", bl->phole_o_c_bw_corr_);
      } else if (ly->GetType() == Layer::l_Lstm_Parallel || ly->GetType() == Layer::l_Lstm) {
        Lstm *ul = dynamic_cast<Lstm*>(ly);
        dump_mat(pre.str() + "wx.npy", ul->wei_gifo_x_corr_);
        dump_mat(pre.str() + "wm.npy", ul->wei_gifo_m_corr_);
        dump_vec(pre.str() + "b.npy", ul->bias_corr_);
        dump_vec(pre.str() + "pi.npy", ul->phole_i_c_corr_);
        dump_vec(pre.str() + "pf.npy", ul->phole_f_c_corr_);
        dump_vec(pre.str() + "po.npy", ul->phole_o_c_corr_);
      } else if (ly->GetType() == Layer::l_Affine_Transform) {
        AffineTransform *af = dynamic_cast<AffineTransform*>(ly);
        dump_mat(pre.str() + "w.npy", af->linearity_corr_);
        dump_vec(pre.str() + "b.npy", af->bias_corr_);
      }
    }
    net.Write(outdir + "/model_out", true);
    return 0;
  } catch (const std::exception &e) {
    std::cerr << e.what();
    return 3;
  }
}
