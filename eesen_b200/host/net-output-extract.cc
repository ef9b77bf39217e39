// eesen_b200/host/net-output-extract.cc -- the forward-only tool of the reference
// (src/netbin/net-output-extract.cc:30-135): same usage line, options and output archive, i.e.
//   for every utterance:  Net::Feedforward -> [ApplyLog] -> [ClassPrior::SubtractOnLogpost] -> write
// What changes (B200): utterances are not pushed through one at a time.  Up to --num-sequence
// consecutive utterances (within --frame-limit padded frames) are packed time-major like a
// training minibatch and run through the same recurrent kernels in one pass; the rows of every
// utterance are written back under its key in the input order.  The results of valid frames do
// not depend on the packing (the backward direction is masked per utterance, the forward direction
// only sees padding AFTER the valid frames), so the archive equals the reference's utterance by
// utterance.  --num-sequence=1 reproduces the reference's call pattern exactly.
#include <algorithm>
#include <chrono>
#include <cstring>

#include "net.h"
#include "options.h"

using namespace eesen;

static const char *kUsage =
    "Perform a forward pass through the network for classification/feature extraction.\n"
    "\n"
    "Usage:  net-output-extract [options] <model-in> <feature-rspecifier> <feature-wspecifier>\n"
    "e.g.: \n"
    "net-output-extract net ark:features.ark ark:output.ark\n"
    "Options: --apply-log --class-frame-counts --prior-scale --prior-cutoff --blank-scale --use-gpu (ignored)\n"
    "         --num-sequence=32 --frame-limit=200000 --gemm-precision=fp32x3|tf32|bf16 --recurrent-precision=fp32x3|tf32\n";

static int PrecFromString(const std::string &s) {
  if (s == "fp32x3" || s == "0") return 0;
  if (s == "tf32" || s == "1") return 1;
  if (s == "bf16" || s == "2") return 2;
  KALDI_ERR << "unknown precision " << s;
  return 0;
}

int main(int argc, char *argv[]) {
  try {
    Options po;
    po.Parse(argc, argv);
    if (po.args.size() != 3) {
      std::cerr << kUsage;
      return 1;
    }
    ClassPriorOptions prior_opts;
    prior_opts.class_frame_counts = po.Str("class-frame-counts", "");
    prior_opts.prior_scale = po.Num("prior-scale", prior_opts.prior_scale);
    prior_opts.prior_cutoff = po.Num("prior-cutoff", prior_opts.prior_cutoff);
    prior_opts.blank_scale = po.Num("blank-scale", prior_opts.blank_scale);
    const bool apply_log = po.Bool("apply-log", false);
    const int32 num_sequence = std::max(1, (int32)po.Num("num-sequence", 32));
    const double frame_limit = po.Num("frame-limit", 200000);
    std::string model_filename = po.args[0], feature_rspecifier = po.args[1], feature_wspecifier = po.args[2];

    eesen_b200_ctx *ctx = NULL;
    if (eesen_b200_create(&ctx, -1)) KALDI_ERR << "eesen_b200_create failed: " << eesen_b200_last_error(NULL);
    CheckAbi(ctx, eesen_b200_set_precision(ctx, PrecFromString(po.Str("gemm-precision", "fp32x3")),
                                           PrecFromString(po.Str("recurrent-precision", "fp32x3"))),
             "eesen_b200_set_precision");
    int32 num_done = 0;
    int64 tot_t = 0;
    auto t0 = std::chrono::steady_clock::now();
    {
      Net net(ctx);
      net.Read(model_filename);
      net.SetTestMode();
      ClassPrior class_prior(ctx, prior_opts);

      SequentialBaseFloatMatrixReader feature_reader(feature_rspecifier);
      BaseFloatMatrixWriter feature_writer(feature_wspecifier);
      CuMatrix<BaseFloat> feats, net_out;
      HostMatrix packed, out_host, utt_out;
      std::vector<std::string> keys;
      std::vector<HostMatrix> utts;

      while (!feature_reader.Done()) {
        // gather the next batch: consecutive utterances, padded size within the frame limit
        keys.clear(); utts.clear();
        int32 max_len = 0;
        while (!feature_reader.Done() && (int32)utts.size() < num_sequence) {
          const HostMatrix &m = feature_reader.Value();
          int32 new_max = std::max(max_len, m.rows);
          if (!utts.empty() && (double)new_max * (utts.size() + 1) > frame_limit) break;
          if (m.rows == 0) { KALDI_WARN << "Empty feature matrix for " << feature_reader.Key(); feature_reader.Next(); continue; }
          keys.push_back(feature_reader.Key());
          utts.push_back(m);
          max_len = new_max;
          feature_reader.Next();
        }
        if (utts.empty()) continue;
        const int32 S = utts.size(), T = max_len, I = utts[0].cols;
        std::vector<int> frames(S);
        packed.Resize(T * S, I);
        for (int32 s = 0; s < S; s++) {
          if (utts[s].cols != I) KALDI_ERR << "Feature dimension changes at " << keys[s];
          frames[s] = utts[s].rows;
          for (int32 t = 0; t < frames[s]; t++) memcpy(packed.Row(t * S + s), utts[s].Row(t), sizeof(float) * I);
        }
        net.SetSeqLengths(frames);
        feats.Resize(T * S, I, kUndefined);
        feats.CopyFromHost(packed.data.data(), I);
        net.Feedforward(feats, &net_out);
        if (apply_log) net_out.ApplyLog();
        if (prior_opts.class_frame_counts != "") class_prior.SubtractOnLogpost(&net_out);
        const int32 K = net_out.NumCols();
        out_host.Resize(T * S, K);
        net_out.CopyToHost(out_host.data.data(), K);
        for (int32 s = 0; s < S; s++) {
          utt_out.Resize(frames[s], K);
          for (int32 t = 0; t < frames[s]; t++) memcpy(utt_out.Row(t), out_host.Row(t * S + s), sizeof(float) * K);
          feature_writer.Write(keys[s], utt_out);
          num_done++;
          tot_t += frames[s];
        }
      }
    }
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    KALDI_LOG << "Done " << num_done << " files in " << el / 60 << "min, (fps " << tot_t / el << ")";
    eesen_b200_destroy(ctx);
    if (num_done == 0) return -1;
    return 0;
  } catch (const std::exception &e) {
    std::cerr << e.what();
    return -1;
  }
}
