// eesen_b200/host/minibatch.h -- host input pipeline of the training driver.
//
// Behaviour is the reference's batching rule (src/netbin/train-ctc-parallel.cc:146-193): consecutive
// utterances with targets, at most --num-sequence per minibatch, the padded size
// (longest utterance x utterance count) never above --frame-limit, over-long utterances skipped,
// rows packed time-major interleaved (row t*S+s) and zero padded.  The design is not: a producer
// thread reads archives and packs the NEXT minibatch into pinned memory while the GPU works on the
// current one (two slots, handed over through a condition variable), so archive I/O, packing and
// the pageable->pinned copy leave the step's critical path.
#ifndef EESEN_B200_HOST_MINIBATCH_H_
#define EESEN_B200_HOST_MINIBATCH_H_

#include <cuda_runtime.h>

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "base.h"

namespace eesen {

struct Minibatch {
  float *feats = nullptr;      // pinned, [T*S x dim], row = t*S + s
  size_t capacity = 0;         // floats
  int32 T = 0, S = 0, dim = 0;
  std::vector<int> frames;                    // valid frames per utterance
  std::vector<std::vector<int> > labels;      // targets per utterance
  std::vector<std::string> keys;
  int64 padded_frames() const { return (int64)T * S; }
};

class MinibatchAssembler {
 public:
  struct Counters {
    int32 no_targets = 0, too_long = 0, bad_dim = 0, bad_labels = 0;
  };

  MinibatchAssembler(const std::string &feature_rspecifier, const std::string &targets_rspecifier, int32 feat_dim,
                     int32 num_sequence, double frame_limit, int device = 0, int32 num_classes = 0)
      : feats_(feature_rspecifier), targets_(targets_rspecifier), dim_(feat_dim), num_sequence_(num_sequence),
        frame_limit_(frame_limit), device_(device), num_classes_(num_classes) {
    worker_ = std::thread(&MinibatchAssembler::Run, this);
  }

  ~MinibatchAssembler() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (worker_.joinable()) worker_.join();
    for (int i = 0; i < 2; i++)
      if (slot_[i].feats) cudaFreeHost(slot_[i].feats);
  }

  // Blocks until the next minibatch is packed; the returned slot stays valid until the following call.
  // Returns NULL at the end of the feature archive.  Errors of the producer surface here.
  const Minibatch *Next() {
    std::unique_lock<std::mutex> lk(mu_);
    if (held_ >= 0) {            // give the previously returned slot back to the producer
      state_[held_] = kEmpty;
      held_ = -1;
      cv_.notify_all();
    }
    cv_.wait(lk, [&] { return state_[read_] == kFull || (done_ && state_[read_] != kFull) || !error_.empty(); });
    if (!error_.empty()) KALDI_ERR << error_;
    if (state_[read_] != kFull) return NULL;
    held_ = read_;
    read_ ^= 1;
    return &slot_[held_];
  }

  Counters counters() {
    std::unique_lock<std::mutex> lk(mu_);
    return counters_;
  }

 private:
  enum State { kEmpty, kFull };

  void Run() {
    try {
      cudaSetDevice(device_);   // pinned allocations belong to this rank's device context
      int w = 0;
      std::vector<HostMatrix> utts;
      while (true) {
        {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return stop_ || state_[w] == kEmpty; });
          if (stop_) return;
        }
        Minibatch &mb = slot_[w];
        if (!Gather(&mb, &utts)) break;
        Pack(&mb, utts);
        {
          std::unique_lock<std::mutex> lk(mu_);
          state_[w] = kFull;
        }
        cv_.notify_all();
        w ^= 1;
      }
    } catch (const std::exception &e) {
      std::unique_lock<std::mutex> lk(mu_);
      error_ = e.what();
    }
    {
      std::unique_lock<std::mutex> lk(mu_);
      done_ = true;
    }
    cv_.notify_all();
  }

  // Collect the utterances of one minibatch.  An utterance that would push the padded size over the
  // limit is kept for the next minibatch (the reference breaks out of its reader loop without advancing).
  bool Gather(Minibatch *mb, std::vector<HostMatrix> *utts) {
    utts->clear();
    mb->frames.clear(); mb->labels.clear(); mb->keys.clear();
    int32 longest = 0;
    Counters local;
    while (!feats_.Done() && (int32)utts->size() < num_sequence_) {
      const std::string key = feats_.Key();
      const HostMatrix &m = feats_.Value();
      if (!targets_.HasKey(key)) {
        KALDI_WARN << key << ", missing targets";
        local.no_targets++;
      } else if (m.rows > frame_limit_) {
        KALDI_WARN << key << ", has too many frames; ignoring: " << m.rows << " > " << frame_limit_;
        local.too_long++;
      } else if (m.cols != dim_) {
        KALDI_ERR << key << ": feature dim " << m.cols << " does not match the network input " << dim_;
      } else if (!LabelsUsable(key, targets_.Value(key))) {
        local.bad_labels++;
      } else {
        int32 cand = std::max<int32>(longest, m.rows);
        if ((double)cand * (utts->size() + 1) > frame_limit_) break;   // does not fit: starts the next minibatch
        longest = cand;
        utts->push_back(m);
        mb->frames.push_back(m.rows);
        mb->labels.push_back(targets_.Value(key));
        mb->keys.push_back(key);
      }
      feats_.Next();
    }
    {
      std::unique_lock<std::mutex> lk(mu_);
      counters_.no_targets += local.no_targets;
      counters_.too_long += local.too_long;
      counters_.bad_labels += local.bad_labels;
    }
    mb->S = utts->size();
    mb->T = longest;
    mb->dim = dim_;
    return mb->S > 0;
  }

  // The reference has no limit on the label sequence and never looks at the ids (ctc-loss.cc:116-129); here the
  // CTC lattice of one utterance lives in the registers of one warp (<= 511 labels) and an id outside
  // [0, num_classes) would index past the posterior row.  Such utterances are skipped with a warning (and
  // counted) instead of aborting the run -- in a multi-rank job an abort would strand the other ranks in
  // the all-reduce.
  bool LabelsUsable(const std::string &key, const std::vector<int> &lab) const {
    if (lab.size() > 511) {
      KALDI_WARN << key << ", has too many labels; ignoring: " << lab.size() << " > 511";
      return false;
    }
    if (num_classes_ > 0)
      for (size_t i = 0; i < lab.size(); i++)
        if (lab[i] < 0 || lab[i] >= num_classes_) {
          KALDI_WARN << key << ", label " << lab[i] << " outside [0, " << num_classes_ << "); ignoring the utterance";
          return false;
        }
    return true;
  }

  void Pack(Minibatch *mb, const std::vector<HostMatrix> &utts) {
    size_t need = (size_t)mb->T * mb->S * mb->dim;
    if (need > mb->capacity) {
      if (mb->feats) cudaFreeHost(mb->feats);
      size_t want = need + need / 4;
      if (cudaMallocHost((void **)&mb->feats, sizeof(float) * want) != cudaSuccess)
        KALDI_ERR << "cudaMallocHost(" << want * sizeof(float) << " bytes) failed";
      mb->capacity = want;
    }
    memset(mb->feats, 0, sizeof(float) * need);
    const size_t row_bytes = sizeof(float) * mb->dim;
    for (int32 s = 0; s < mb->S; s++)
      for (int32 t = 0; t < utts[s].rows; t++)
        memcpy(mb->feats + ((size_t)t * mb->S + s) * mb->dim, utts[s].Row(t), row_bytes);
  }

  SequentialBaseFloatMatrixReader feats_;
  RandomAccessInt32VectorReader targets_;
  int32 dim_, num_sequence_;
  double frame_limit_;
  int device_;
  int32 num_classes_ = 0;
  Minibatch slot_[2];
  State state_[2] = {kEmpty, kEmpty};
  int read_ = 0, held_ = -1;
  bool done_ = false, stop_ = false;
  std::string error_;
  Counters counters_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::thread worker_;
};

}  // namespace eesen
#endif
