// placeholder -- filled in by the K2 milestone
#include "xq_ctx.h"
extern "C" void xq_mcts_free_(xq_ctx*) {}
