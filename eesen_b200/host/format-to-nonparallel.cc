// eesen_b200/host/format-to-nonparallel.cc -- reference src/netbin/format-to-nonparallel.cc:24-76:
// rewrites the layer markers (<BiLstmParallel> -> <BiLstm>) so that decoding tools accept the model.
// Parameters are copied through the device arena unchanged (bit-exact).
#include "net.h"
#include "options.h"

using namespace eesen;

int main(int argc, char *argv[]) {
  try {
    Options po;
    po.Parse(argc, argv);
    if (po.args.size() != 2) {
      std::cerr << "Convert model format to the non-parallel version (<BiLstmParallel> -> <BiLstm>).\n"
                   "Usage:  format-to-nonparallel [--binary=true] <model-in> <model-out>\n";
      return 1;
    }
    eesen_b200_ctx *ctx = NULL;
    if (eesen_b200_create(&ctx, -1)) KALDI_ERR << "eesen_b200_create failed: " << eesen_b200_last_error(NULL);
    {
      Net net(ctx);
      net.Read(po.args[0]);
      net.WriteNonParal(po.args[1], po.Bool("binary", true));
    }
    KALDI_LOG << "Written model to " << po.args[1];
    eesen_b200_destroy(ctx);
    return 0;
  } catch (const std::exception &e) {
    std::cerr << e.what() << '\n';
    return -1;
  }
}
