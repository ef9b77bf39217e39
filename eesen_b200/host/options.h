// eesen_b200/host/options.h -- the "--name=value" command-line convention of the reference tools
// (src/util/parse-options.cc: options before/among positional arguments, --config=<file> with one
// "--name=value" per line and '#' comments, command line wins over the config file).
#ifndef EESEN_B200_HOST_OPTIONS_H_
#define EESEN_B200_HOST_OPTIONS_H_

#include <cstdlib>
#include <fstream>
#include <map>
#include <string>
#include <vector>

namespace eesen {

struct Options {
  std::map<std::string, std::string> kv;
  std::vector<std::string> args;
  bool Has(const std::string &k) const { return kv.count(k) > 0; }
  std::string Str(const std::string &k, const std::string &d) const { return Has(k) ? kv.at(k) : d; }
  double Num(const std::string &k, double d) const { return Has(k) ? atof(kv.at(k).c_str()) : d; }
  bool Bool(const std::string &k, bool d) const {
    if (!Has(k)) return d;
    const std::string &v = kv.at(k);
    return v == "" || v == "true" || v == "1" || v == "yes";
  }
  void Parse(int argc, char *argv[]) {
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i];
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        size_t eq = a.find('=');
        if (eq == std::string::npos) kv[a.substr(2)] = "";
        else kv[a.substr(2, eq - 2)] = a.substr(eq + 1);
      } else {
        args.push_back(a);
      }
    }
    if (!Has("config")) return;
    std::ifstream cf(kv["config"].c_str());
    std::string line;
    while (std::getline(cf, line)) {
      size_t h = line.find('#');
      if (h != std::string::npos) line.resize(h);
      size_t b = line.find("--");
      if (b == std::string::npos) continue;
      line = line.substr(b + 2);
      while (!line.empty() && isspace(line[line.size() - 1])) line.resize(line.size() - 1);
      size_t eq = line.find('=');
      std::string k = eq == std::string::npos ? line : line.substr(0, eq);
      if (!kv.count(k)) kv[k] = eq == std::string::npos ? "" : line.substr(eq + 1);
    }
  }
};

}  // namespace eesen
#endif
