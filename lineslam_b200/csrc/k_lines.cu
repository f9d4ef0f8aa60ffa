// k_lines.cu — per-line stages (under construction: first GPU run validates K1-K4 only)
#include "lsl_internal.h"
int lsl_launch_lines(lsl_ctx* ctx, int n, const float* d_depth, const double K[9], double dt) {
  LSL_CUDA(cudaMemsetAsync(ctx->wk.nlines, 0, sizeof(int32_t) * n, ctx->stream));
  return LSL_OK;
}
