extern "C" {
int hmc_estimate_counts(hmc_ctx *ctx, int, int, unsigned, int, const int32_t *, const int64_t *, const int8_t *, const double *, const int32_t *, double *, double *)
{ return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_estimate_counts: not built yet") : HMC_ERR_INVALID; }
int hmc_estimate_counts_device(hmc_ctx *ctx, int, int, unsigned, int, const int32_t *, const int64_t *, const int8_t *, const double *, const int32_t *, double **)
{ return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_estimate_counts_device: not built yet") : HMC_ERR_INVALID; }
}
