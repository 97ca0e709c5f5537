// Host driver of the level-synchronous forward path (forward_flat.cuh): the whole locus loop is enqueued on the
// context's stream without host synchronisation; sizes live in device memory.  Returns 1 when the wave must be re-run by
// the per-individual kernel (capacity or arena exceeded), 0 on success, < 0 on a CUDA error.

struct FlatPool {
    bool ready = false, dirty = true;
    int K = 0, n_cap = 0;
    long long hcap = 0, ccap = 0, scap = 0;
    DevBuf<unsigned char> mem;
    DevBuf<FlatTotals> tot;
    FlatParams base = {};
    void release() { mem.release(); tot.release(); ready = false; }
};

static int flat_build_pool(hmc_ctx *ctx, FlatPool &fp, int n, int K)
{
    if (fp.ready && fp.K >= K && fp.n_cap >= n) return HMC_OK;
    fp.release();
    const char *env = getenv("HMC_B200_FLAT_CCAP_LOG2");
    const int lg = env ? atoi(env) : 24;
    fp.ccap = 1ll << lg; fp.hcap = 2 * fp.ccap; fp.scap = fp.ccap / 2; fp.K = K; fp.n_cap = n;
    auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
    const size_t tiles = (size_t) (fp.ccap / FLAT_TILE + 2);
    size_t total = 0;
    auto take = [&](size_t bytes) { const size_t o = total; total += al(bytes); return o; };
    const size_t o_S = take(4 * (size_t) n), o_So = take(4 * (size_t) n), o_C = take(4 * (size_t) n), o_h = take(4 * (size_t) n), o_e = take(4 * (size_t) n), o_r = take(4 * (size_t) n);
    const size_t o_off = take(8 * (size_t) n), o_noff = take(8 * (size_t) n), o_cb = take(8 * (size_t) (n + 1)), o_hb = take(8 * (size_t) (n + 1)),
                 o_si = take(8 * (size_t) (n + 1)), o_so = take(8 * (size_t) (n + 1)), o_ab = take(8 * (size_t) (n + 1)), o_am = take(8 * (size_t) n);
    const size_t o_keys = take(8 * (size_t) fp.hcap), o_tpv = take(8 * (size_t) fp.hcap), o_ord = take(4 * (size_t) fp.hcap), o_cnt = take(4 * (size_t) fp.hcap);
    const size_t o_cand = take(4 * (size_t) fp.ccap), o_cw = take(4 * (size_t) fp.ccap), o_arr = take(4 * (size_t) fp.ccap), o_ex = take(8 * (size_t) fp.ccap);
    const size_t o_ts = take(8 * tiles), o_tp = take(8 * tiles);
    const size_t o_aoff = take(8 * (size_t) fp.scap), o_slot = take(4 * (size_t) fp.scap), o_stw = take(4 * (size_t) fp.scap);
    const size_t o_l0 = take(8 * (size_t) fp.scap * K), o_l1 = take(8 * (size_t) fp.scap * K), o_m0 = take(8 * (size_t) fp.scap), o_m1 = take(8 * (size_t) fp.scap);
    CK(fp.mem.ensure(total));
    CK(fp.tot.ensure(1));
    unsigned char *b = fp.mem.p;
    FlatParams &q = fp.base;
    q.S = (int *) (b + o_S); q.S_out = (int *) (b + o_So); q.C = (int *) (b + o_C); q.hsz = (int *) (b + o_h); q.expo = (int *) (b + o_e); q.rescale = (int *) (b + o_r);
    q.off = (long long *) (b + o_off); q.new_off = (long long *) (b + o_noff); q.cbase = (long long *) (b + o_cb); q.hbase = (long long *) (b + o_hb);
    q.sbase_in = (long long *) (b + o_si); q.sbase_out = (long long *) (b + o_so); q.abase = (long long *) (b + o_ab); q.amax = (u64 *) (b + o_am);
    q.keys = (u64 *) (b + o_keys); q.tpv = (double *) (b + o_tpv); q.ord = (u32 *) (b + o_ord); q.cnt = (u32 *) (b + o_cnt);
    q.cand = (int *) (b + o_cand); q.candw = (int *) (b + o_cw); q.arr = (int *) (b + o_arr); q.ex_local = (u64 *) (b + o_ex);
    q.tile_sum = (u64 *) (b + o_ts); q.tile_pre = (u64 *) (b + o_tp);
    q.aoff = (long long *) (b + o_aoff); q.slot_of = (int *) (b + o_slot); q.st_w = (int *) (b + o_stw);
    q.like[0] = (double *) (b + o_l0); q.like[1] = (double *) (b + o_l1); q.meta[0] = (u64 *) (b + o_m0); q.meta[1] = (u64 *) (b + o_m1);
    q.hcap = fp.hcap; q.ccap = fp.ccap; q.scap = fp.scap; q.tot = fp.tot.p;
    fp.ready = true; fp.dirty = true;
    return HMC_OK;
}

static int run_flat(hmc_ctx *ctx, FlatPool &fp, int first, int count, int K, unsigned flags)
{
    const int L = ctx->L, kmax = kmax_for(K);
    ctx->ecache.valid = false;
    int rc = ensure_arena(ctx);
    if (rc) return rc;
    rc = flat_build_pool(ctx, fp, count, K);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx->d_loci.ensure((size_t) count * (L + 1)));
    CK(ctx->d_wave.ensure(count)); CK(ctx->d_status.ensure(count)); CK(ctx->d_final_exp.ensure(count));
    CK(ctx->d_counters.ensure(24));
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 24 * sizeof(long long), st));
    ctx->h_wave.resize(count);
    for (int i = 0; i < count; ++i) ctx->h_wave[i] = first + i;
    CK(cudaMemcpyAsync(ctx->d_wave.p, ctx->h_wave.data(), sizeof(int) * count, cudaMemcpyHostToDevice, st));
    if (fp.dirty) {     // empty hash: keys = EMPTY, ord = +inf, cnt = 0 (kept clean by FE afterwards)
        CK(cudaMemsetAsync(fp.base.keys, 0xFF, 8 * (size_t) fp.hcap, st));
        CK(cudaMemsetAsync(fp.base.ord, 0xFF, 4 * (size_t) fp.hcap, st));
        CK(cudaMemsetAsync(fp.base.cnt, 0, 4 * (size_t) fp.hcap, st));
        fp.dirty = false;
    }
    CK(cudaMemsetAsync(fp.tot.p, 0, sizeof(FlatTotals), st));
    FlatParams p = fp.base;
    p.panel.N = ctx->N; p.panel.L = L; p.panel.Amax = ctx->Amax; p.panel.bits = ctx->bits; p.panel.words = ctx->words;
    p.panel.packed = ctx->d_packed.p; p.panel.linfo = ctx->d_linfo.p; p.panel.order = ctx->d_order.p;
    p.model = ctx->dm;
    p.arena = ctx->d_arena.p; p.arena_cap = (long long) ctx->d_arena.n;
    p.loci = ctx->d_loci.p; p.wave = ctx->d_wave.p; p.n_wave = count; p.K = K; p.scaled = (flags & HMC_FLAG_SCALED) ? 1 : 0;
    p.status = ctx->d_status.p; p.final_exp = ctx->d_final_exp.p; p.counters = ctx->d_counters.p;
    p.par = 0;
    const int grid = ctx->sms * 8;
    CK(cudaEventRecord(ctx->ev[0], st));
    flat_head_kernel<<<1, 1024, 0, st>>>(p);
    int launches = 1;
    for (int i = ctx->dm.head_len; i < L; ++i) {
        flat_f0_kernel<<<1, 1024, 0, st>>>(p);
        flat_fa_kernel<<<grid, 256, 0, st>>>(p);
        flat_fb_kernel<<<grid, 256, 0, st>>>(p);
        flat_fs_kernel<<<1, 1024, 0, st>>>(p);
        flat_f2_kernel<<<1, 1024, 0, st>>>(p);
        flat_fc_kernel<<<grid, 256, 0, st>>>(p);
        flat_fd_kernel<<<grid, 256, 0, st>>>(p);
        switch (kmax) {
            case 1: flat_fe_kernel<1><<<grid * 2, 128, 0, st>>>(p); break;
            case 4: flat_fe_kernel<4><<<grid * 2, 128, 0, st>>>(p); break;
            case 10: flat_fe_kernel<10><<<grid * 2, 128, 0, st>>>(p); break;
            case 16: flat_fe_kernel<16><<<grid * 2, 128, 0, st>>>(p); break;
            default: flat_fe_kernel<32><<<grid * 2, 128, 0, st>>>(p); break;
        }
        flat_f3_kernel<<<1, 1024, 0, st>>>(p);
        launches += 9;
        p.par ^= 1;
        if (p.scaled) { flat_rescale_kernel<<<grid, 256, 0, st>>>(p, p.par); ++launches; }
    }
    flat_final_alloc_kernel<<<1, 1024, 0, st>>>(p);
    flat_final_copy_kernel<<<grid, 256, 0, st>>>(p, p.par);
    launches += 2;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], st));
    TraceParams tp = {};
    tp.panel = p.panel; tp.model = p.model; tp.arena = ctx->d_arena.p; tp.loci = ctx->d_loci.p; tp.wave = ctx->d_wave.p;
    tp.status = ctx->d_status.p; tp.final_exp = ctx->d_final_exp.p; tp.n_wave = count; tp.K = K; tp.first = first;
    tp.pg = ctx->d_pg.p; tp.cov = ctx->d_cov.p; tp.prior = ctx->d_prior.p; tp.post = ctx->d_post.p; tp.exp2 = ctx->d_exp.p; tp.nres = ctx->d_nres.p;
    tp.haps = ctx->d_haps.p;
    CK(launch_traceback(kmax, tp, st));
    ++launches;
    CK(cudaEventRecord(ctx->ev[2], st));
    FlatTotals ht;
    CK(cudaMemcpyAsync(&ht, fp.tot.p, sizeof ht, cudaMemcpyDeviceToHost, st));
    std::vector<int> h_status(count);
    CK(cudaMemcpyAsync(h_status.data(), ctx->d_status.p, sizeof(int) * count, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ht.abort_code) {
        fp.dirty = true;
        if (getenv("HMC_B200_FLAT_STRICT")) return ctx->fail(HMC_ERR_NOMEM, "level-synchronous path aborted with code %d at locus %d", ht.abort_code, ht.locus);
        return 1;
    }
    for (int i = 0; i < count; ++i) if (h_status[i] == ST_NO_HEAD) return ctx->fail(HMC_ERR_NO_HEAD, "individual %d: Can not find matching pattern!", first + i);
    float ms = 0;
    memset(ctx->kernel_ms, 0, sizeof ctx->kernel_ms);
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->kernel_ms[0] = ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->kernel_ms[1] = ms;
    long long hc[8];
    CK(cudaMemcpy(hc, ctx->d_counters.p, sizeof hc, cudaMemcpyDeviceToHost));
    memset(ctx->counters, 0, sizeof ctx->counters);
    ctx->counters[HMC_CNT_STEPS] = hc[0]; ctx->counters[HMC_CNT_STATES] = hc[1]; ctx->counters[HMC_CNT_MAX_STATES] = hc[2];
    ctx->counters[HMC_CNT_ARRIVALS] = hc[3]; ctx->counters[HMC_CNT_WAVES] = 1; ctx->counters[HMC_CNT_LAUNCHES] = launches;
    ctx->last_K = K; ctx->last_estep = 0;
    return HMC_OK;
}
