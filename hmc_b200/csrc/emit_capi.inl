// C-ABI of K8 (emit.cuh): hmc_get_likelihood, hmc_format_best.  Included by capi.cu.
extern "C" {

int hmc_get_likelihood(hmc_ctx *ctx, int n, const int8_t *haps, double *out)
{
    if (!ctx || !out) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_get_likelihood: bad argument") : HMC_ERR_INVALID;
    if (!ctx->have_model) return ctx->fail(HMC_ERR_INVALID, "hmc_get_likelihood before hmc_set_model");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int L = ctx->L;
    const signed char *src = nullptr;
    long long pair_stride = 2;
    if (haps) {
        if (n <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_get_likelihood: bad count");
        CK(ctx->d_cmp_infer.ensure((size_t) n * L));
        CK(cudaMemcpyAsync(ctx->d_cmp_infer.p, haps, (size_t) n * L, cudaMemcpyHostToDevice, st));
        src = ctx->d_cmp_infer.p;
    }
    else {      // both haplotypes of the best resolution of every individual of the last resolve call, where they lie
        if (ctx->r_count <= 0 || n != 2 * ctx->r_count) return ctx->fail(HMC_ERR_INVALID, "hmc_get_likelihood: no resolve results for %d haplotypes on the device", n);
        src = ctx->d_haps.p; pair_stride = 2ll * ctx->r_K;
    }
    CK(ctx->d_lk.ensure((size_t) n)); CK(ctx->d_model_err.ensure(1));
    CK(cudaMemsetAsync(ctx->d_model_err.p, 0, sizeof(int), st));
    likelihood_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, st>>>(ctx->dm, L, ctx->d_linfo.p, src, pair_stride, n, ctx->d_lk.p, ctx->d_model_err.p);
    CK(cudaGetLastError());
    int bad = 0;
    CK(cudaMemcpyAsync(out, ctx->d_lk.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&bad, ctx->d_model_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bad) return ctx->fail(HMC_ERR_INVALID, "hmc_get_likelihood: a haplotype holds a missing or out-of-range allele");
    return HMC_OK;
}

int hmc_format_best(hmc_ctx *ctx, const int32_t *symbols, const char *allele_type, char *out, int64_t capacity, int64_t *line_off, int64_t *total)
{
    if (!ctx || !symbols || !allele_type || !total) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_format_best: bad argument") : HMC_ERR_INVALID;
    if (ctx->r_count <= 0 || !ctx->d_raw.p) return ctx->fail(HMC_ERR_INVALID, "hmc_format_best: no resolve results on the device");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = ctx->r_count, L = ctx->L, A = ctx->Amax;
    // text widths of every allele of every locus, with the delimiter; the last entry of a locus is the missing allele
    std::vector<unsigned char> wid((size_t) L * (A + 1));
    for (int l = 0; l < L; ++l) {
        for (int k = 0; k <= A; ++k) {
            int w = 2;
            if (allele_type[l] != 'S') {
                const int v = k < A ? symbols[(size_t) l * A + k] : -1;
                char buf[16];
                w = snprintf(buf, sizeof buf, "%d ", v);
            }
            wid[(size_t) l * (A + 1) + k] = (unsigned char) w;
        }
    }
    CK(ctx->d_emit_sym.ensure((size_t) L * A)); CK(ctx->d_emit_wid.ensure(wid.size() + L)); CK(ctx->d_emit_off.ensure((size_t) 2 * n + 1));
    CK(cudaMemcpyAsync(ctx->d_emit_sym.p, symbols, sizeof(int) * (size_t) L * A, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_emit_wid.p, wid.data(), wid.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_emit_wid.p + wid.size(), allele_type, (size_t) L, cudaMemcpyHostToDevice, st));
    EmitParams e = {};
    e.n = n; e.L = L; e.Amax = A; e.K = ctx->r_K;
    e.haps = ctx->d_haps.p; e.raw = ctx->d_raw.p + (size_t) ctx->r_first * 2 * L; e.nres = ctx->d_nres.p;
    e.sym = ctx->d_emit_sym.p; e.wid = ctx->d_emit_wid.p; e.type = (const char *) (ctx->d_emit_wid.p + wid.size());
    e.off = ctx->d_emit_off.p;
    const unsigned grid = (unsigned) std::min<long long>(((long long) 2 * n * 32 + 255) / 256, (long long) ctx->sms * 16);
    CK(cudaMemsetAsync(e.off + 2 * n, 0, sizeof(long long), st));
    emit_len_kernel<<<grid, 256, 0, st>>>(e);
    CK(cudaGetLastError());
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, e.off, e.off, 2 * n + 1, st));
    CK(ctx->d_emit_tmp.ensure(tb));
    CK(cub::DeviceScan::ExclusiveSum(ctx->d_emit_tmp.p, tb, e.off, e.off, 2 * n + 1, st));
    long long tot = 0;
    CK(cudaMemcpyAsync(&tot, e.off + 2 * n, sizeof tot, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *total = tot;
    if (line_off) CK(cudaMemcpyAsync(line_off, e.off, sizeof(long long) * ((size_t) 2 * n + 1), cudaMemcpyDeviceToHost, st));
    if (out && capacity >= tot) {
        CK(ctx->d_emit_text.ensure((size_t) tot));
        e.text = ctx->d_emit_text.p;
        emit_text_kernel<<<grid, 256, 0, st>>>(e);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, e.text, (size_t) tot, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    if (out && capacity < tot) return ctx->fail(HMC_ERR_LIMIT, "hmc_format_best: the text takes %lld bytes, the buffer holds %lld", tot, (long long) capacity);
    return HMC_OK;
}

}  // extern "C"
