// eesen_b200/host/base.cc -- see base.h.  Own implementation of the Kaldi on-disk conventions
// the hot path touches; formats follow reference src/base/io-funcs.cc, io-funcs-inl.h,
// src/cpucompute/matrix.cc:968-1010 (Write) / :1040-1160 (Read), src/util/kaldi-table-inl.h.
#include "base.h"

#include <cmath>
#include <cstring>
#include <ext/stdio_filebuf.h>
#include <fstream>

namespace eesen {

int g_verbose_level = 0;

MessageLogger::MessageLogger(Kind kind, const char *func, const char *file, int line) : kind_(kind) {
  const char *base = strrchr(file, '/');
  const char *tag = kind == kError ? "ERROR" : kind == kWarning ? "WARNING" : kind == kLog ? "LOG" : "VLOG";
  ss_ << tag << " (" << func << "():" << (base ? base + 1 : file) << ":" << line << ") ";
}

MessageLogger::~MessageLogger() noexcept(false) {
  if (kind_ == kError) throw std::runtime_error(ss_.str());
  std::cerr << ss_.str() << std::endl;
}

void WriteToken(std::ostream &os, bool, const std::string &token) { os << token << " "; }

void ReadToken(std::istream &is, bool binary, std::string *token) {
  if (!binary) is >> std::ws;
  is >> *token;
  if (is.fail()) KALDI_ERR << "ReadToken, failed to read token at file position " << is.tellg();
  if (!isspace(is.peek())) KALDI_ERR << "ReadToken, expected space after token, saw instead " << (char)is.peek();
  is.get();
}

void ExpectToken(std::istream &is, bool binary, const std::string &token) {
  std::string got;
  ReadToken(is, binary, &got);
  if (got != token) KALDI_ERR << "Expected token \"" << token << "\", got instead \"" << got << "\".";
}

int Peek(std::istream &is, bool binary) {
  if (!binary) is >> std::ws;
  return is.peek();
}

void WriteBasicType(std::ostream &os, bool binary, int32 v) {
  if (binary) { os.put((char)sizeof(v)); os.write(reinterpret_cast<const char *>(&v), sizeof(v)); }
  else os << v << " ";
}
void WriteBasicType(std::ostream &os, bool binary, float v) {
  if (binary) { os.put((char)sizeof(v)); os.write(reinterpret_cast<const char *>(&v), sizeof(v)); }
  else os << v << " ";
}
void WriteBasicType(std::ostream &os, bool binary, bool v) {
  os << (v ? "T" : "F");
  if (!binary) os << " ";
}
void ReadBasicType(std::istream &is, bool binary, int32 *v) {
  if (binary) {
    int sz = is.get();
    if (sz != (int)sizeof(*v)) KALDI_ERR << "ReadBasicType: expected int32 size marker, saw " << sz;
    is.read(reinterpret_cast<char *>(v), sizeof(*v));
  } else {
    is >> *v;
  }
  if (is.fail()) KALDI_ERR << "ReadBasicType<int32> failed";
}
void ReadBasicType(std::istream &is, bool binary, float *v) {
  if (binary) {
    int sz = is.peek();
    if (sz == (int)sizeof(float)) {
      is.get();
      is.read(reinterpret_cast<char *>(v), sizeof(float));
    } else if (sz == (int)sizeof(double)) {
      is.get();
      double d;
      is.read(reinterpret_cast<char *>(&d), sizeof(d));
      *v = (float)d;
    } else {
      KALDI_ERR << "ReadBasicType: expected float, saw " << sz;
    }
  } else {
    is >> *v;
  }
  if (is.fail()) KALDI_ERR << "ReadBasicType<float> failed";
}
void ReadBasicType(std::istream &is, bool binary, bool *v) {
  if (!binary) is >> std::ws;
  int c = is.peek();
  if (c == 'T') *v = true;
  else if (c == 'F') *v = false;
  else KALDI_ERR << "Read failure in ReadBasicType<bool>, next char is " << (char)c;
  is.get();
}

bool InitKaldiInputStream(std::istream &is, bool *binary) {
  if (is.peek() == '\0') {
    is.get();
    if (is.peek() != 'B') return false;
    is.get();
    *binary = true;
  } else {
    *binary = false;
  }
  return true;
}

static void ReadTextNumbers(std::istream &is, std::vector<std::vector<float> > *rows) {
  // " [" then rows separated by newlines / ";" then "]"
  std::string tok;
  is >> tok;
  if (tok != "[") KALDI_ERR << "Expected \"[\", got " << tok;
  rows->clear();
  std::vector<float> cur;
  while (true) {
    int c = is.peek();
    if (c == EOF) KALDI_ERR << "EOF while reading text matrix";
    if (c == ']') { is.get(); break; }
    if (c == '\n' || c == ';') {
      is.get();
      if (!cur.empty()) { rows->push_back(cur); cur.clear(); }
      continue;
    }
    if (isspace(c)) { is.get(); continue; }
    std::string num;
    is >> num;
    if (!num.empty() && num[num.size() - 1] == ']') {
      num.resize(num.size() - 1);
      if (!num.empty()) cur.push_back((float)atof(num.c_str()));
      break;
    }
    if (num == "inf" || num == "Inf") cur.push_back(HUGE_VALF);
    else if (num == "-inf" || num == "-Inf") cur.push_back(-HUGE_VALF);
    else cur.push_back((float)atof(num.c_str()));
  }
  if (!cur.empty()) rows->push_back(cur);
}

void HostMatrix::Read(std::istream &is, bool binary) {
  if (binary) {
    std::string tok;
    ReadToken(is, binary, &tok);
    int32 r, c;
    if (tok == "FM" || tok == "DM") {
      ReadBasicType(is, binary, &r);
      ReadBasicType(is, binary, &c);
      Resize(r, c);
      if (tok == "FM") {
        is.read(reinterpret_cast<char *>(data.data()), sizeof(float) * data.size());
      } else {
        std::vector<double> tmp((size_t)r * c);
        is.read(reinterpret_cast<char *>(tmp.data()), sizeof(double) * tmp.size());
        for (size_t i = 0; i < tmp.size(); i++) data[i] = (float)tmp[i];
      }
      if (is.fail()) KALDI_ERR << "Failed to read matrix data";
    } else {
      KALDI_ERR << "Expected token FM or DM, got " << tok;
    }
  } else {
    std::vector<std::vector<float> > rows_;
    ReadTextNumbers(is, &rows_);
    int32 r = rows_.size(), c = r ? rows_[0].size() : 0;
    Resize(r, c);
    for (int32 i = 0; i < r; i++) {
      if ((int32)rows_[i].size() != c) KALDI_ERR << "Inconsistent row length in text matrix";
      memcpy(Row(i), rows_[i].data(), sizeof(float) * c);
    }
  }
}

void HostMatrix::Write(std::ostream &os, bool binary) const {
  if (binary) {
    WriteToken(os, binary, "FM");
    WriteBasicType(os, binary, rows);
    WriteBasicType(os, binary, cols);
    os.write(reinterpret_cast<const char *>(data.data()), sizeof(float) * data.size());
  } else {
    if (cols == 0) { os << " [ ]\n"; return; }
    os << " [";
    for (int32 i = 0; i < rows; i++) {
      os << "\n  ";
      for (int32 j = 0; j < cols; j++) os << Row(i)[j] << " ";
    }
    os << "]\n";
  }
}

void HostVector::Read(std::istream &is, bool binary) {
  if (binary) {
    std::string tok;
    ReadToken(is, binary, &tok);
    int32 n;
    if (tok == "FV") {
      ReadBasicType(is, binary, &n);
      data.resize(n);
      is.read(reinterpret_cast<char *>(data.data()), sizeof(float) * n);
    } else if (tok == "DV") {
      ReadBasicType(is, binary, &n);
      std::vector<double> tmp(n);
      is.read(reinterpret_cast<char *>(tmp.data()), sizeof(double) * n);
      data.assign(tmp.begin(), tmp.end());
    } else {
      KALDI_ERR << "Expected token FV or DV, got " << tok;
    }
    if (is.fail()) KALDI_ERR << "Failed to read vector data";
  } else {
    std::vector<std::vector<float> > rows_;
    ReadTextNumbers(is, &rows_);
    data.clear();
    for (size_t i = 0; i < rows_.size(); i++) data.insert(data.end(), rows_[i].begin(), rows_[i].end());
  }
}

void HostVector::Write(std::ostream &os, bool binary) const {
  if (binary) {
    WriteToken(os, binary, "FV");
    WriteBasicType(os, binary, (int32)data.size());
    os.write(reinterpret_cast<const char *>(data.data()), sizeof(float) * data.size());
  } else {
    os << " [ ";
    for (size_t i = 0; i < data.size(); i++) os << data[i] << " ";
    os << "]\n";
  }
}

// ------------------------------------------------------------------------------------ tables
std::istream *OpenInput(const std::string &rx, FILE **pipe_out, bool *owns) {
  *pipe_out = nullptr;
  *owns = true;
  std::string f = rx;
  while (!f.empty() && isspace(f[f.size() - 1])) f.resize(f.size() - 1);
  if (f == "-" || f.empty()) { *owns = false; return &std::cin; }
  if (f[f.size() - 1] == '|') {
    std::string cmd = f.substr(0, f.size() - 1);
    FILE *p = popen(cmd.c_str(), "r");
    if (!p) KALDI_ERR << "Failed to open pipe: " << cmd;
    *pipe_out = p;
    auto *buf = new __gnu_cxx::stdio_filebuf<char>(p, std::ios::in | std::ios::binary);
    return new std::istream(buf);
  }
  auto *fs = new std::ifstream(f.c_str(), std::ios::in | std::ios::binary);
  if (!fs->is_open()) { delete fs; KALDI_ERR << "Failed to open input file " << f; }
  return fs;
}

static void SplitRspecifier(const std::string &rspec, std::string *kind, std::string *opts, std::string *rest) {
  size_t colon = rspec.find(':');
  if (colon == std::string::npos) KALDI_ERR << "Invalid rspecifier " << rspec;
  std::string head = rspec.substr(0, colon);
  *rest = rspec.substr(colon + 1);
  size_t comma = head.find(',');
  *kind = head.substr(0, comma);
  *opts = comma == std::string::npos ? "" : head.substr(comma + 1);
  if (*kind != "ark" && *kind != "scp") KALDI_ERR << "Unsupported rspecifier type in " << rspec;
}

SequentialBaseFloatMatrixReader::SequentialBaseFloatMatrixReader(const std::string &rspecifier) {
  std::string opts, rest;
  SplitRspecifier(rspecifier, &kind_, &opts, &rest);
  if (kind_ == "ark") {
    is_ = OpenInput(rest, &pipe_, &owns_);
  } else {
    FILE *p; bool own;
    std::istream *s = OpenInput(rest, &p, &own);
    std::string line;
    while (std::getline(*s, line)) {
      std::istringstream ls(line);
      std::string key, path;
      ls >> key;
      std::getline(ls, path);
      size_t b = path.find_first_not_of(" \t");
      if (key.empty() || b == std::string::npos) continue;
      scp_.push_back(std::make_pair(key, path.substr(b)));
    }
    if (own) delete s;
    if (p) pclose(p);
  }
  ReadOne();
}

SequentialBaseFloatMatrixReader::~SequentialBaseFloatMatrixReader() {
  if (owns_ && is_) delete is_;
  if (pipe_) pclose(pipe_);
}

void SequentialBaseFloatMatrixReader::Next() { ReadOne(); }

void SequentialBaseFloatMatrixReader::ReadOne() {
  if (kind_ == "ark") {
    *is_ >> std::ws;
    if (is_->peek() == EOF || !(*is_ >> key_)) { done_ = true; return; }
    is_->get();  // the space after the key
    bool binary;
    if (!InitKaldiInputStream(*is_, &binary)) KALDI_ERR << "Bad archive header for key " << key_;
    value_.Read(*is_, binary);
  } else {
    if (scp_pos_ >= scp_.size()) { done_ = true; return; }
    key_ = scp_[scp_pos_].first;
    std::string path = scp_[scp_pos_].second;
    scp_pos_++;
    long offset = -1;
    size_t colon = path.rfind(':');
    if (colon != std::string::npos && colon + 1 < path.size() && isdigit(path[colon + 1]) &&
        path.find_first_not_of("0123456789", colon + 1) == std::string::npos) {
      offset = atol(path.c_str() + colon + 1);
      path = path.substr(0, colon);
    }
    FILE *p; bool own;
    std::istream *s = OpenInput(path, &p, &own);
    if (offset >= 0) s->seekg(offset);
    bool binary;
    if (!InitKaldiInputStream(*s, &binary)) KALDI_ERR << "Bad header in " << path;
    value_.Read(*s, binary);
    if (own) delete s;
    if (p) pclose(p);
  }
}

BaseFloatMatrixWriter::BaseFloatMatrixWriter(const std::string &wspecifier) {
  size_t colon = wspecifier.find(':');
  if (colon == std::string::npos) KALDI_ERR << "Invalid wspecifier " << wspecifier;
  std::string head = wspecifier.substr(0, colon), rest = wspecifier.substr(colon + 1);
  bool want_scp = false, is_ark = false;
  std::istringstream hs(head);
  std::string opt;
  while (std::getline(hs, opt, ',')) {
    if (opt == "ark") is_ark = true;
    else if (opt == "scp") want_scp = true;
    else if (opt == "t") binary_ = false;
    else if (opt == "b" || opt == "f" || opt == "nf" || opt == "p") {}
    else KALDI_ERR << "Unsupported wspecifier option '" << opt << "' in " << wspecifier;
  }
  if (!is_ark) KALDI_ERR << "Only archive wspecifiers are supported: " << wspecifier;
  std::string ark = rest, scp;
  if (want_scp) {
    size_t comma = rest.find(',');
    if (comma == std::string::npos) KALDI_ERR << "ark,scp needs two file names: " << wspecifier;
    ark = rest.substr(0, comma);
    scp = rest.substr(comma + 1);
  }
  while (!ark.empty() && isspace(ark[0])) ark.erase(0, 1);
  ark_name_ = ark;
  if (ark == "-" || ark.empty()) {
    os_ = &std::cout;
  } else if (ark[0] == '|') {
    pipe_ = popen(ark.substr(1).c_str(), "w");
    if (!pipe_) KALDI_ERR << "Failed to open pipe: " << ark;
    os_ = new std::ostream(new __gnu_cxx::stdio_filebuf<char>(pipe_, std::ios::out | std::ios::binary));
    owns_ = true;
  } else {
    auto *fs = new std::ofstream(ark.c_str(), std::ios::out | std::ios::binary);
    if (!fs->is_open()) { delete fs; KALDI_ERR << "Failed to open " << ark << " for writing"; }
    os_ = fs;
    owns_ = true;
  }
  if (want_scp) {
    auto *fs = new std::ofstream(scp.c_str());
    if (!fs->is_open()) { delete fs; KALDI_ERR << "Failed to open " << scp << " for writing"; }
    scp_ = fs;
  }
}

BaseFloatMatrixWriter::~BaseFloatMatrixWriter() {
  if (os_) os_->flush();
  if (owns_ && os_) { std::streambuf *b = pipe_ ? os_->rdbuf() : nullptr; delete os_; delete b; }
  if (pipe_) pclose(pipe_);
  if (scp_) delete scp_;
}

void BaseFloatMatrixWriter::Write(const std::string &key, const HostMatrix &value) {
  *os_ << key << ' ';
  if (scp_) *scp_ << key << ' ' << ark_name_ << ':' << (long)os_->tellp() << "\n";
  if (binary_) { os_->put('\0'); os_->put('B'); }
  value.Write(*os_, binary_);
  if (os_->fail()) KALDI_ERR << "Write failure for key " << key;
}

RandomAccessInt32VectorReader::RandomAccessInt32VectorReader(const std::string &rspecifier) {
  std::string kind, opts, rest;
  SplitRspecifier(rspecifier, &kind, &opts, &rest);
  if (kind != "ark") KALDI_ERR << "Label rspecifier must be an archive: " << rspecifier;
  FILE *p; bool own;
  std::istream *s = OpenInput(rest, &p, &own);
  while (true) {
    *s >> std::ws;
    std::string key;
    if (s->peek() == EOF || !(*s >> key)) break;
    s->get();
    std::vector<int32> v;
    if (s->peek() == '\0') {  // binary: \0B \4 <size> then (\4 int32)*
      bool binary;
      InitKaldiInputStream(*s, &binary);
      int32 n;
      ReadBasicType(*s, true, &n);
      v.resize(n);
      for (int32 i = 0; i < n; i++) ReadBasicType(*s, true, &v[i]);
    } else {
      std::string line;
      std::getline(*s, line);
      std::istringstream ls(line);
      int32 x;
      while (ls >> x) v.push_back(x);
    }
    items_[key] = v;
  }
  if (own) delete s;
  if (p) pclose(p);
}

bool RandomAccessInt32VectorReader::HasKey(const std::string &key) const { return items_.count(key) > 0; }
const std::vector<int32> &RandomAccessInt32VectorReader::Value(const std::string &key) const {
  std::map<std::string, std::vector<int32> >::const_iterator it = items_.find(key);
  if (it == items_.end()) KALDI_ERR << "Value() called for a key that is not present: " << key;
  return it->second;
}

}  // namespace eesen
