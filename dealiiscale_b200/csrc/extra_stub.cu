// temporary stubs
#include "ctx.h"
namespace msgpu {
int parabolic_setup(msgpu_ctx* ctx) { ctx->err = "parabolic not implemented"; return MSGPU_ERR_UNSUPPORTED; }
int parabolic_assemble(msgpu_ctx* ctx, double, double) { ctx->err = "parabolic not implemented"; return MSGPU_ERR_UNSUPPORTED; }
int compute_errors(msgpu_ctx* ctx, double*, double*) { ctx->err = "errors not implemented"; return MSGPU_ERR_UNSUPPORTED; }
int compute_residuals(msgpu_ctx* ctx, double*) { ctx->err = "residuals not implemented"; return MSGPU_ERR_UNSUPPORTED; }
int comm_allgather_coupling(msgpu_ctx* ctx, double*, double*) { ctx->err = "comm not implemented"; return MSGPU_ERR_UNSUPPORTED; }
void comm_destroy(msgpu_ctx*) {}
}
extern "C" {
int msgpu_comm_unique_id(char*) { return MSGPU_ERR_UNSUPPORTED; }
int msgpu_comm_init(msgpu_ctx*, const char*, int, int, int, int) { return MSGPU_ERR_UNSUPPORTED; }
int msgpu_broadcast_macro(msgpu_ctx*, double*, double*, int) { return MSGPU_ERR_UNSUPPORTED; }
}
