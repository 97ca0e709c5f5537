// hmc_estimate_counts: K1 (E-step trellis, K = 1) -> K3 backward -> K4 estep_counts, per arena wave.

static int build_trie(hmc_ctx *ctx, EstepTrie &t, int Pc, const int32_t *start, const int64_t *al_off, const int8_t *alleles)
{
    const int L = ctx->L, A = ctx->Amax;
    // ForwardPatternTree::addPattern (PatternTree.cpp:188-212).  The sub-tries of different start loci share nothing, so they
    // are built by host threads (candidates bucketed by start, in list order: a later duplicate replaces the earlier one,
    // Tree.h:87-98) and then laid out one after the other; node i <= L is the root of start locus i.
    std::vector<int> bucket_off(L + 2, 0);
    for (int i = 0; i < Pc; ++i) {
        const long long len = al_off[i + 1] - al_off[i];
        if (start[i] < 0 || len <= 0 || start[i] + len > L) return ctx->fail(HMC_ERR_INVALID, "candidate pattern %d outside the panel", i);
        ++bucket_off[start[i] + 1];
    }
    for (int l = 0; l <= L; ++l) bucket_off[l + 1] += bucket_off[l];
    std::vector<int> by_start(Pc), cursor(bucket_off.begin(), bucket_off.end() - 1);
    for (int i = 0; i < Pc; ++i) by_start[cursor[start[i]]++] = i;
    struct Sub { std::vector<int> child, data; int max_depth = 0, bad = -1, bad_allele = 0; };
    std::vector<Sub> subs(L + 1);
    const int n_threads = (int) std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            const int st = next.fetch_add(1);
            if (st > L) break;
            Sub &sb = subs[st];
            sb.child.assign(A, -1); sb.data.assign(1, -1);          // local node 0 = the root of this start locus
            for (int q = bucket_off[st]; q < bucket_off[st + 1]; ++q) {
                const int i = by_start[q];
                const int len = (int) (al_off[i + 1] - al_off[i]);
                int node = 0;
                for (int k = 0; k < len; ++k) {
                    const int a = alleles[al_off[i] + k];
                    if (a < 0 || a >= A) { if (sb.bad < 0) { sb.bad = i; sb.bad_allele = a; } node = -1; break; }
                    int c = sb.child[(size_t) node * A + a];
                    if (c < 0) {
                        c = (int) sb.data.size();
                        sb.data.push_back(-1);
                        sb.child.insert(sb.child.end(), A, -1);
                        sb.child[(size_t) node * A + a] = c;
                    }
                    node = c;
                }
                if (node >= 0) sb.data[node] = i;
                sb.max_depth = std::max(sb.max_depth, len);
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (int q = 1; q < n_threads; ++q) pool.emplace_back(work);
        work();
        for (std::thread &th : pool) th.join();
    }
    int max_depth = 0;
    std::vector<long long> base(L + 2, 0);      // global id of local node k > 0 of start st: base[st] + k - 1
    base[0] = L + 1;
    for (int st = 0; st <= L; ++st) {
        if (subs[st].bad >= 0) return ctx->fail(HMC_ERR_INVALID, "candidate pattern %d holds allele index %d", subs[st].bad, subs[st].bad_allele);
        max_depth = std::max(max_depth, subs[st].max_depth);
        base[st + 1] = base[st] + (long long) subs[st].data.size() - 1;
    }
    if (base[L + 1] > 0x7FFFFFF0ll / std::max(A, 1)) return ctx->fail(HMC_ERR_LIMIT, "candidate trie too large");
    const size_t n_nodes = (size_t) base[L + 1];
    std::vector<int> child(n_nodes * A, -1), data(n_nodes, -1), root(L + 1);
    for (int i = 0; i <= L; ++i) root[i] = i;
    next = 0;
    auto place = [&]() {
        for (;;) {
            const int st = next.fetch_add(1);
            if (st > L) break;
            const Sub &sb = subs[st];
            const size_t nn = sb.data.size();
            for (size_t k = 0; k < nn; ++k) {
                const size_t gid = k == 0 ? (size_t) st : (size_t) (base[st] + (long long) k - 1);
                data[gid] = sb.data[k];
                for (int a = 0; a < A; ++a) { const int c = sb.child[k * A + a]; child[gid * A + a] = c < 0 ? -1 : (int) (base[st] + c - 1); }
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (int q = 1; q < n_threads; ++q) pool.emplace_back(place);
        place();
        for (std::thread &th : pool) th.join();
    }
    if (max_depth > ESTEP_MAX_DEPTH) return ctx->fail(HMC_ERR_LIMIT, "pattern length %d exceeds %d", max_depth, ESTEP_MAX_DEPTH);
    CK(t.d_root.ensure(root.size())); CK(t.d_child.ensure(child.size())); CK(t.d_data.ensure(data.size()));
    CK(cudaMemcpyAsync(t.d_root.p, root.data(), sizeof(int) * root.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(t.d_child.p, child.data(), sizeof(int) * child.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(t.d_data.p, data.data(), sizeof(int) * data.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    t.dev.n_nodes = (int) data.size(); t.dev.A = A; t.dev.max_depth = max_depth;
    t.dev.root = t.d_root.p; t.dev.child = t.d_child.p; t.dev.data = t.d_data.p;
    return HMC_OK;
}

static int estimate_counts_impl(hmc_ctx *ctx, int first, int count, unsigned flags, int Pc, const int32_t *start, const int64_t *al_off,
                                const int8_t *alleles, const double *pg, const int32_t *pg_exp)
{
    if (!ctx) return HMC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (Pc <= 0 || !start || !al_off || !alleles || !pg || count <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_estimate_counts: bad argument");
    EstepTrie &trie = ctx->trie;      // rebuilt per call; device buffers are reused
    int rc = build_trie(ctx, trie, Pc, start, al_off, alleles);
    if (rc) return rc;
    CK(ctx->d_counts.ensure((size_t) 2 * Pc));
    CK(cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(double) * 2 * Pc, ctx->stream));
    CK(ctx->d_pgin.ensure(count));
    CK(cudaMemcpyAsync(ctx->d_pgin.p, pg, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    const bool scaled = (flags & HMC_FLAG_SCALED) != 0;
    if (pg_exp) { CK(ctx->d_pgexp.ensure(count)); CK(cudaMemcpyAsync(ctx->d_pgexp.p, pg_exp, sizeof(int) * count, cudaMemcpyHostToDevice, ctx->stream)); }
    DevBuf<int> &d_bexp = ctx->d_bexp, &d_eflags = ctx->d_eflags;
    DevBuf<unsigned char> &d_escr = ctx->d_escr;
    DevBuf<EstepScratch> &d_edesc = ctx->d_edesc;
    CK(d_eflags.ensure(4));

    const bool reuse = ctx->ecache.valid && ctx->ecache.first == first && ctx->ecache.count == count && ctx->ecache.flags == flags &&
                       !getenv("HMC_B200_NO_ESTEP_CACHE");
    auto after = [&](const FwdParams &fp, int n_wave, const std::vector<int> &, int &launches) -> int {
        const int L = ctx->L;
        if (!reuse) {
        // K3
        BackParams bp = {};
        bp.panel = fp.panel; bp.model = fp.model; bp.arena = ctx->d_arena.p; bp.loci = ctx->d_loci.p; bp.wave = ctx->d_wave.p; bp.status = ctx->d_status.p;
        bp.n_wave = n_wave; bp.scaled = scaled ? 1 : 0; bp.bexp = nullptr;
        if (scaled) { CK(d_bexp.ensure((size_t) ctx->d_loci.n)); bp.bexp = d_bexp.p; }
        CK(cudaEventRecord(ctx->ev[3], ctx->stream));
        backward_kernel<<<(n_wave + 3) / 4, 128, 0, ctx->stream>>>(bp);
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[4], ctx->stream));
        ++launches;
        }
        // K4: scratch sized from the largest live-state set of this wave
        long long hc[8];
        CK(cudaMemcpyAsync(hc, ctx->d_counters.p, sizeof hc, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const int smax = (int) std::max<long long>(hc[2], 1);
        int occ = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, estep_kernel, ESTEP_WARPS * 32, 0));
        long long ecap = std::max<long long>(4096, 3LL * smax * 2);     // entry stack; grown x4 and the wave retried if a task overflows it
        for (int attempt = 0; attempt < 4; ++attempt) {
            auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
            const size_t b_acc = al((size_t) 3 * smax * 8), b_key = al((size_t) ecap * 4), b_w = al((size_t) ecap * 8);
            const size_t per = b_acc + b_key + b_w;
            size_t fr = 0, tot = 0;
            CK(cudaMemGetInfo(&fr, &tot));
            const size_t budget = std::max<size_t>((size_t) 2 << 30, std::min<size_t>((size_t) 40 << 30, (fr + d_escr.n) / 2));
            long long warps = std::min<long long>((long long) ctx->sms * occ * ESTEP_WARPS, std::max<long long>(ESTEP_WARPS, (long long) (budget / per)));
            warps = std::min<long long>(warps, ((long long) n_wave * L + ESTEP_WARPS - 1));
            const int blocks = (int) std::max<long long>(1, warps / ESTEP_WARPS);
            warps = (long long) blocks * ESTEP_WARPS;
            CK(d_escr.ensure(per * warps));
            CK(d_edesc.ensure(warps));
            CK(cudaMemsetAsync(d_escr.p, 0, per * warps, ctx->stream));
            std::vector<EstepScratch> h(warps);
            for (long long q = 0; q < warps; ++q) {
                unsigned char *b = d_escr.p + per * q;
                h[q].acc = (double *) b; b += b_acc;
                h[q].ekey = (int *) b; b += b_key;
                h[q].ew = (double *) b;
                h[q].smax = smax; h[q].ecap = (int) ecap;
            }
            CK(cudaMemcpyAsync(d_edesc.p, h.data(), sizeof(EstepScratch) * warps, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemsetAsync(d_eflags.p, 0, 4 * sizeof(int), ctx->stream));
            EstepParams ep = {};
            ep.panel = fp.panel; ep.model = fp.model; ep.trie = trie.dev; ep.arena = ctx->d_arena.p; ep.loci = ctx->d_loci.p; ep.wave = ctx->d_wave.p;
            ep.status = ctx->d_status.p; ep.aexp = scaled ? ctx->d_locus_exp.p : nullptr; ep.bexp = scaled ? d_bexp.p : nullptr; ep.final_exp = ctx->d_final_exp.p;
            ep.pg = ctx->d_pgin.p; ep.pg_exp = pg_exp ? ctx->d_pgexp.p : nullptr; ep.first = first; ep.n_wave = n_wave; ep.scratch = d_edesc.p;
            ep.counts = ctx->d_counts.p; ep.Pc = Pc; ep.queue = d_eflags.p; ep.error = d_eflags.p + 2;
            ep.work = (unsigned long long *) (ctx->d_counters.p + 7);
            // the counts of a failed attempt must not be kept: snapshot and restore
            DevBuf<double> &snap = ctx->d_counts_snap;
            CK(snap.ensure((size_t) 2 * Pc));
            CK(cudaMemcpyAsync(snap.p, ctx->d_counts.p, sizeof(double) * 2 * Pc, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaEventRecord(ctx->ev[5], ctx->stream));
            estep_kernel<<<blocks, ESTEP_WARPS * 32, 0, ctx->stream>>>(ep);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev[6], ctx->stream));
            ++launches;
            int eflags[4];
            CK(cudaMemcpyAsync(eflags, d_eflags.p, sizeof eflags, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            if (!reuse) { cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); if (attempt == 0) ctx->kernel_ms[2] += ms; }
            cudaEventElapsedTime(&ms, ctx->ev[5], ctx->ev[6]); ctx->kernel_ms[3] += ms;
            if (eflags[2] == 0) return HMC_OK;
            if (eflags[2] == 2) return ctx->fail(HMC_ERR_LIMIT, "candidate trie deeper than %d", ESTEP_MAX_DEPTH);
            CK(cudaMemcpyAsync(ctx->d_counts.p, snap.p, sizeof(double) * 2 * Pc, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_counters.p + 7, &hc[7], sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));     // entry counter of the failed attempt
            CK(cudaStreamSynchronize(ctx->stream));
            ecap *= 4;
        }
        return ctx->fail(HMC_ERR_LIMIT, "E-step entry stack overflow");
    };
    if (reuse) {
        // same individuals, same model, trellis still in the arena: only the candidate list changed (the later E-passes of an M-step)
        const std::vector<int> none;
        int launches = 0;
        memset(ctx->kernel_ms, 0, sizeof ctx->kernel_ms);
        CK(cudaMemsetAsync(ctx->d_counters.p + 7, 0, sizeof(long long), ctx->stream));
        rc = after(ctx->ecache.fp, ctx->ecache.n_wave, none, launches);
        if (rc) { ctx->ecache.valid = false; return rc; }
        long long entries = 0;
        CK(cudaMemcpy(&entries, ctx->d_counters.p + 7, sizeof entries, cudaMemcpyDeviceToHost));
        memcpy(ctx->counters, ctx->ecache.counters, sizeof ctx->counters);
        ctx->counters[HMC_CNT_ESTEP_ENTRIES] = entries; ctx->counters[HMC_CNT_LAUNCHES] = launches;
        return HMC_OK;
    }
    return run_waves(ctx, first, count, 1, flags, 1, after);
}

extern "C" {

int hmc_estimate_counts_device(hmc_ctx *ctx, int first, int count, unsigned flags, int Pc, const int32_t *start, const int64_t *al_off,
                               const int8_t *alleles, const double *pg, const int32_t *pg_exp, double **dev_counts)
{
    int rc = estimate_counts_impl(ctx, first, count, flags, Pc, start, al_off, alleles, pg, pg_exp);
    if (rc) return rc;
    if (dev_counts) *dev_counts = ctx->d_counts.p;
    return HMC_OK;
}

int hmc_estimate_counts(hmc_ctx *ctx, int first, int count, unsigned flags, int Pc, const int32_t *start, const int64_t *al_off,
                        const int8_t *alleles, const double *pg, const int32_t *pg_exp, double *out_freq, double *out_prefix)
{
    int rc = estimate_counts_impl(ctx, first, count, flags, Pc, start, al_off, alleles, pg, pg_exp);
    if (rc) return rc;
    if (out_freq) CK(cudaMemcpy(out_freq, ctx->d_counts.p, sizeof(double) * Pc, cudaMemcpyDeviceToHost));
    if (out_prefix) CK(cudaMemcpy(out_prefix, ctx->d_counts.p + Pc, sizeof(double) * Pc, cudaMemcpyDeviceToHost));
    return HMC_OK;
}

}  // extern "C"
