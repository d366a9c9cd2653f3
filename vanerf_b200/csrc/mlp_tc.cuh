// Tensor-core (tcgen05 / TMEM / TMA) shading path — bf16 operands, fp32 accumulation.  Placeholder until the kernel
// lands: the bf16 precision is reported as unsupported instead of silently taking another path.
#pragma once
static int tc_pack_weights(vanerf_ctx*, const vanerf_weights*, void*) { return VANERF_OK; }
static int tc_shade_chunk(vanerf_ctx* ctx, const float*, long long, int, float*, float*, cudaStream_t) {
    snprintf(ctx->err, sizeof(ctx->err), "bf16 tensor-core path not built in this revision");
    return VANERF_ERR_UNSUPPORTED;
}
