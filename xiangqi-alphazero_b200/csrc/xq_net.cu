// placeholder -- filled in by the K3 milestone
#include "xq_ctx.h"
extern "C" void xq_net_free_(xq_ctx*) {}
