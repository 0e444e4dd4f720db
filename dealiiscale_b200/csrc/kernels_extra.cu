// Errors / residuals / parabolic driver pieces (filled in step by step).
#include "ctx.h"
namespace msgpu {
int parabolic_setup(msgpu_ctx* ctx) { ctx->err = "parabolic not implemented"; return MSGPU_ERR_UNSUPPORTED; }
int parabolic_assemble(msgpu_ctx* ctx, double, double) { ctx->err = "parabolic not implemented"; return MSGPU_ERR_UNSUPPORTED; }
int compute_errors(msgpu_ctx* ctx, double*, double*) { ctx->err = "errors not implemented"; return MSGPU_ERR_UNSUPPORTED; }
int compute_residuals(msgpu_ctx* ctx, double*) { ctx->err = "residuals not implemented"; return MSGPU_ERR_UNSUPPORTED; }
}
