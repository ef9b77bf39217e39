// eesen_b200/host/net-change-model.cc -- reference src/netbin/net-change-model.cc:24-103: sets the dropout
// options of every BiLSTM layer (all options must be given; unspecified ones default to off) and/or converts the
// model between binary and text.  Parameters pass through the device arena unchanged.
#include "net.h"
#include "options.h"

using namespace eesen;

int main(int argc, char *argv[]) {
  try {
    Options po;
    po.Parse(argc, argv);
    if (po.args.size() != 2) {
      std::cerr << "Change network model with specified options and possibly change binary/text format\n"
                   "Note: Model options must be fully specified, options will default to false otherwise\n"
                   "Usage:  net-change-model [options] <model-in> <model-out>\n"
                   "Options: --binary --forwarddrop --forwardstep --forwardseq --rnndrop --nmldrop --recurrentdrop\n"
                   "         --recurrentstep --recurrentseq --twiddleforward\n";
      return 1;
    }
    eesen_b200_ctx *ctx = NULL;
    if (eesen_b200_create(&ctx, -1)) KALDI_ERR << "eesen_b200_create failed: " << eesen_b200_last_error(NULL);
    {
      Net net(ctx);
      net.Read(po.args[0]);
      net.ChangeDropoutParameters(po.Num("forwarddrop", 0.0), po.Bool("forwardstep", false), po.Bool("forwardseq", false),
                                  po.Bool("rnndrop", false), po.Bool("nmldrop", false), po.Num("recurrentdrop", 0.0),
                                  po.Bool("recurrentstep", false), po.Bool("recurrentseq", false),
                                  po.Bool("twiddleforward", false));
      net.Write(po.args[1], po.Bool("binary", true));
    }
    KALDI_LOG << "Written model to " << po.args[1];
    eesen_b200_destroy(ctx);
    return 0;
  } catch (const std::exception &e) {
    std::cerr << e.what() << '\n';
    return -1;
  }
}
