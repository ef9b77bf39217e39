// eesen_b200/host/base.h -- minimal Kaldi-style host infrastructure for the hot path:
// logging/asserts (reference src/base/kaldi-error.h:91-111), binary/text token I/O
// (src/base/io-funcs-inl.h:32-60, io-funcs.cc), matrix/vector on-disk formats
// (src/cpucompute/matrix.cc:968-1010, vector.cc), ark/scp readers for the training driver
// (src/util/kaldi-table-inl.h, kaldi-holder-inl.h).  Own implementation; formats byte-compatible.
#ifndef EESEN_B200_HOST_BASE_H_
#define EESEN_B200_HOST_BASE_H_

#include <cstdint>
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace eesen {

typedef float BaseFloat;
typedef int32_t int32;
typedef int64_t int64;

extern int g_verbose_level;  // --verbose

// KALDI_ERR throws std::runtime_error when the temporary dies (kaldi-error.cc behaviour)
class MessageLogger {
 public:
  enum Kind { kError, kWarning, kLog, kVlog };
  MessageLogger(Kind kind, const char *func, const char *file, int line);
  ~MessageLogger() noexcept(false);
  std::ostream &stream() { return ss_; }

 private:
  Kind kind_;
  std::ostringstream ss_;
};

#define KALDI_ERR ::eesen::MessageLogger(::eesen::MessageLogger::kError, __func__, __FILE__, __LINE__).stream()
#define KALDI_WARN ::eesen::MessageLogger(::eesen::MessageLogger::kWarning, __func__, __FILE__, __LINE__).stream()
#define KALDI_LOG ::eesen::MessageLogger(::eesen::MessageLogger::kLog, __func__, __FILE__, __LINE__).stream()
#define KALDI_VLOG(v) \
  if ((v) <= ::eesen::g_verbose_level) ::eesen::MessageLogger(::eesen::MessageLogger::kVlog, __func__, __FILE__, __LINE__).stream()
#define KALDI_ASSERT(cond) \
  do { if (!(cond)) KALDI_ERR << "Assertion failed: " #cond; } while (0)

// ---- token / basic-type I/O
void WriteToken(std::ostream &os, bool binary, const std::string &token);
void ReadToken(std::istream &is, bool binary, std::string *token);
void ExpectToken(std::istream &is, bool binary, const std::string &token);
int Peek(std::istream &is, bool binary);
void WriteBasicType(std::ostream &os, bool binary, int32 v);
void WriteBasicType(std::ostream &os, bool binary, float v);
void WriteBasicType(std::ostream &os, bool binary, bool v);
void ReadBasicType(std::istream &is, bool binary, int32 *v);
void ReadBasicType(std::istream &is, bool binary, float *v);
void ReadBasicType(std::istream &is, bool binary, bool *v);
bool InitKaldiInputStream(std::istream &is, bool *binary);  // consumes the "\0B" header if present

// ---- host matrix (row-major, dense) -- enough for model/feature I/O and batching
struct HostMatrix {
  int32 rows = 0, cols = 0;
  std::vector<float> data;
  void Resize(int32 r, int32 c) { rows = r; cols = c; data.assign((size_t)r * c, 0.f); }
  float *Row(int32 r) { return data.data() + (size_t)r * cols; }
  const float *Row(int32 r) const { return data.data() + (size_t)r * cols; }
  void Read(std::istream &is, bool binary);   // "FM" / "DM" / text "[ ... ]"
  void Write(std::ostream &os, bool binary) const;
};
struct HostVector {
  std::vector<float> data;
  void Read(std::istream &is, bool binary);   // "FV" / "DV" / text "[ ... ]"
  void Write(std::ostream &os, bool binary) const;
};

// ---- table readers used by train-ctc-parallel (rspecifiers: ark:file, ark,t:file, scp:file, "ark:cmd |")
class SequentialBaseFloatMatrixReader {
 public:
  explicit SequentialBaseFloatMatrixReader(const std::string &rspecifier);
  ~SequentialBaseFloatMatrixReader();
  bool Done() const { return done_; }
  void Next();
  const std::string &Key() const { return key_; }
  const HostMatrix &Value() const { return value_; }

 private:
  void ReadOne();
  std::string kind_;
  FILE *pipe_ = nullptr;
  std::istream *is_ = nullptr;
  bool owns_ = false;
  std::vector<std::pair<std::string, std::string> > scp_;
  size_t scp_pos_ = 0;
  bool done_ = false;
  std::string key_;
  HostMatrix value_;
};

class RandomAccessInt32VectorReader {
 public:
  explicit RandomAccessInt32VectorReader(const std::string &rspecifier);
  bool HasKey(const std::string &key) const;
  const std::vector<int32> &Value(const std::string &key) const;

 private:
  std::map<std::string, std::vector<int32> > items_;
};

// wspecifiers: ark:file, ark,t:file, ark:- (stdout), "ark:| cmd", ark,scp:file.ark,file.scp
// (the archive entry format of util/kaldi-holder-inl.h: "<key> " + "\0B" + FM matrix, or text)
class BaseFloatMatrixWriter {
 public:
  explicit BaseFloatMatrixWriter(const std::string &wspecifier);
  ~BaseFloatMatrixWriter();
  void Write(const std::string &key, const HostMatrix &value);

 private:
  bool binary_ = true;
  FILE *pipe_ = nullptr;
  std::ostream *os_ = nullptr, *scp_ = nullptr;
  bool owns_ = false;
  std::string ark_name_;
};

std::istream *OpenInput(const std::string &rxfilename, FILE **pipe_out, bool *owns);

}  // namespace eesen
#endif
