// eesen_b200/host/train-ctc-parallel.cc -- the training driver, host logic as in the reference
// (src/netbin/train-ctc-parallel.cc:30-264): same usage line, options, batching/padding rule,
// per-batch call sequence on Net/Ctc and log lines (the recipes grep "TOKEN_ACCURACY" and the fps
// line, asr_egs/wsj/steps/train_ctc_parallel.sh:146,158).  What changes (north_star):
//   * no CuDevice::SelectGpuId / per-call sync: one eesen_b200 context per process;
//   * --num-jobs/--job-id select NCCL ranks: every step all-reduces the gradient over NVLink
//     instead of averaging model files every --utts-per-avg utterances
//     (src/net/communicator.h:39-119).  The NCCL unique id is exchanged through the file
//     <model-out>.ncclid written by job 1 (the same shared-filesystem rendezvous the reference uses).
#include <unistd.h>

#include <cuda_runtime.h>

#include <chrono>
#include <cstring>
#include <fstream>

#include "minibatch.h"
#include "net.h"
#include "options.h"

using namespace eesen;

static const char *kUsage =
    "Perform one iteration of CTC training by SGD.\n"
    "The updates are done per-utterance and by processing multiple utterances in parallel.\n"
    "\n"
    "Usage: train-ctc-parallel [options] <feature-rspecifier> <labels-rspecifier> <model-in> [<model-out>]\n"
    "e.g.: \n"
    "train-ctc-parallel scp:feature.scp ark:labels.ark nnet.init nnet.iter1\n"
    "Options: --learn-rate --momentum --binary --cross-validate --num-sequence --frame-limit --report-step\n"
    "         --num-jobs --job-id --opt-algorithm=SGD|Adagrad|RMSProp --adagrad-epsilon --rms-prop-rho --sequence-out-file --verbose --dropout-seed\n"
    "         --gemm-precision=fp32x3|tf32|bf16 --recurrent-precision=fp32x3|tf32\n";

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
    g_verbose_level = (int)po.Num("verbose", 0);

    NetTrainOptions trn_opts;
    trn_opts.learn_rate = po.Num("learn-rate", trn_opts.learn_rate);
    trn_opts.momentum = po.Num("momentum", trn_opts.momentum);
    trn_opts.adagrad_epsilon = po.Num("adagrad-epsilon", trn_opts.adagrad_epsilon);
    // as in the reference, rmsprop_one_minus_rho keeps its default 0.1: train-opts.h:50 derives it
    // inside Register(), i.e. before --rms-prop-rho is parsed
    trn_opts.rmsprop_rho = po.Num("rms-prop-rho", trn_opts.rmsprop_rho);
    bool binary = po.Bool("binary", true), crossvalidate = po.Bool("cross-validate", false);
    std::string sequence_out_file = po.Str("sequence-out-file", "");
    int32 num_sequence = (int32)po.Num("num-sequence", 5);
    double frame_limit = po.Num("frame-limit", 100000);
    int32 report_step = (int32)po.Num("report-step", 100);
    int32 num_jobs = (int32)po.Num("num-jobs", 1), job_id = (int32)po.Num("job-id", 1);
    std::string opt = po.Str("opt-algorithm", "SGD");

    if ((int)po.args.size() != 4 - (crossvalidate ? 1 : 0)) {
      std::cerr << kUsage;
      return 1;
    }
    std::string feature_rspecifier = po.args[0], targets_rspecifier = po.args[1], model_filename = po.args[2];
    std::string target_model_filename = crossvalidate ? "" : po.args[3];

    eesen_b200_ctx *ctx = NULL;
    int rc = eesen_b200_create(&ctx, num_jobs > 1 ? job_id - 1 : -1);
    if (rc) KALDI_ERR << "eesen_b200_create failed: " << eesen_b200_last_error(NULL);
    CheckAbi(ctx, eesen_b200_set_precision(ctx, PrecFromString(po.Str("gemm-precision", "fp32x3")),
                                           PrecFromString(po.Str("recurrent-precision", "fp32x3"))),
             "eesen_b200_set_precision");
    if (num_jobs > 1) {
      // Rendezvous through a file next to the model, as the reference's jobs meet through files
      // (net/communicator.h:57-71).  A file left behind by an earlier run on the same model path must never be
      // read as this run's id: job 1 stamps the id with its own start time, the other jobs accept only a file
      // whose stamp is not older than a minute before their own start (the jobs of one run are started together
      // by the recipe), and job 1 removes the file once every rank has joined the communicator.
      std::string idfile = (crossvalidate ? model_filename + ".cv" : target_model_filename) + ".ncclid";
      char id[128];
      const long long now = (long long)std::chrono::duration_cast<std::chrono::seconds>(
                                std::chrono::system_clock::now().time_since_epoch()).count();
      if (job_id == 1) {
        std::remove(idfile.c_str());
        if (eesen_b200_nccl_unique_id(id)) KALDI_ERR << "ncclGetUniqueId failed";
        std::string tmp = idfile + ".tmp";
        { std::ofstream f(tmp.c_str(), std::ios::binary); f.write(id, 128); f.write((const char *)&now, sizeof(now)); }
        std::rename(tmp.c_str(), idfile.c_str());
      } else {
        for (int tries = 0;; tries++) {
          std::ifstream f(idfile.c_str(), std::ios::binary);
          long long stamp = 0;
          if (f.is_open() && f.read(id, 128) && f.read((char *)&stamp, sizeof(stamp)) && stamp >= now - 60) break;
          if (tries > 200000) KALDI_ERR << "timed out waiting for " << idfile;
          usleep(300);
        }
      }
      CheckAbi(ctx, eesen_b200_nccl_init(ctx, job_id - 1, num_jobs, id), "eesen_b200_nccl_init");
      // ncclCommInitRank returns only after all ranks have joined: nobody needs the file any more
      if (job_id == 1) std::remove(idfile.c_str());
    }

    {
      Net net(ctx);
      net.Read(model_filename);
      net.SetTrainOptions(trn_opts);
      net.SetUpdateAlgorithm(opt);
      if (po.Has("dropout-seed")) net.SetDropoutSeed((uint64_t)po.Num("dropout-seed", 0));   // default: random per run
      if (crossvalidate) net.SetTestMode(); else net.SetTrainMode();

      Ctc ctc(ctx);
      ctc.SetReportStep(report_step);
      CuMatrix<BaseFloat> feats_dev, net_out, obj_diff;

      KALDI_LOG << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << " STARTED";
      if (sequence_out_file.length()) std::remove(sequence_out_file.c_str());
      const auto t_start = std::chrono::steady_clock::now();

      // producer thread: archives -> packed, pinned minibatches (minibatch.h); this loop only feeds the GPU
      int device_index = 0;
      cudaGetDevice(&device_index);
      MinibatchAssembler batches(feature_rspecifier, targets_rspecifier, net.InputDim(), num_sequence, frame_limit,
                                 device_index, net.OutputDim());
      int64 total_frames = 0;
      int32 num_done = 0, steps_since_check = 0;
      cudaStream_t stream = (cudaStream_t)eesen_b200_stream(ctx);
      cudaEvent_t h2d_done;
      cudaEventCreateWithFlags(&h2d_done, cudaEventDisableTiming);
      bool h2d_pending = false;
      while (true) {
        // the pinned slot of the previous minibatch goes back to the producer inside Next():
        // its (asynchronous) upload must have left the host buffer by then
        if (h2d_pending) { cudaEventSynchronize(h2d_done); h2d_pending = false; }
        const Minibatch *mb = batches.Next();
        if (!mb) {
          // out of data: in a multi-rank training run keep voting "idle" in the per-step all-reduce
          // until every rank is done (ranks may hold different numbers of minibatches)
          if (num_jobs > 1 && !crossvalidate && net.BackpropagateShared(NULL) > 0) continue;
          break;
        }
        std::vector<int> frame_num_utt = mb->frames;
        std::vector<std::vector<int> > labels_utt = mb->labels;
        feats_dev.Resize(mb->T * mb->S, mb->dim, kUndefined);
        feats_dev.CopyFromHost(mb->feats, mb->dim);                                   // one async H2D from pinned memory
        cudaEventRecord(h2d_done, stream);
        h2d_pending = true;
        net.SetSeqLengths(frame_num_utt);                                             // reference :195
        net.Propagate(feats_dev, &net_out);                                           // :198
        ctc.EvalParallel(frame_num_utt, net_out, labels_utt, &obj_diff);              // :199
        ctc.ErrorRateMSeq(frame_num_utt, net_out, labels_utt, sequence_out_file);     // :202
        if (!crossvalidate) {                                                         // :207, gradient all-reduce inside
          if (num_jobs > 1) net.BackpropagateShared(&obj_diff);
          else net.Backpropagate(obj_diff, NULL);
        }
        num_done += mb->S;
        total_frames += mb->padded_frames();
        // a diverged run is caught at the next report instead of at the final Net::Write (net.cc:448-468);
        // every rank holds the same parameters, so every rank stops at the same step
        if (!crossvalidate && report_step > 0 && ++steps_since_check >= report_step) {
          net.Check();
          steps_since_check = 0;
        }
      }

      cudaEventDestroy(h2d_done);
      const std::string report = ctc.Report();   // drains the stream and folds the last statistics
      MinibatchAssembler::Counters skipped = batches.counters();
      if (!crossvalidate) {
        KALDI_LOG << net.Info();
        KALDI_LOG << net.InfoGradient();
        if (num_jobs == 1 || job_id == 1) net.Write(target_model_filename, binary);
      }
      const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
      // as the reference prints it (train-ctc-parallel.cc:247-251): utterances above --frame-limit are warned about
      // and skipped but not counted here; "other errors" stays 0 on this path too (a bad feature dimension aborts)
      KALDI_LOG << "Done " << num_done << " files, " << skipped.no_targets << " with no targets, "
                << skipped.bad_dim + skipped.bad_labels << " with other errors. [" << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << ", "
                << elapsed / 60 << " min, fps" << total_frames / elapsed << "]";
      KALDI_LOG << report;
    }
    eesen_b200_destroy(ctx);
    return 0;
  } catch (const std::exception &e) {
    std::cerr << e.what();
    return -1;
  }
}
