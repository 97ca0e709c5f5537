// Host driver of K5 (mine.cuh): level loop, capacity management, id assignment, result staging.

// per-level buffers grow geometrically: a reallocation per level would cost more than the kernels
template <typename T>
static cudaError_t grow(DevBuf<T> &b, size_t want)
{
    if (want <= b.n) return cudaSuccess;
    // headroom only while it is cheap: beyond 1 GiB the exact size is taken
    const size_t bytes = want * sizeof(T);
    return b.ensure(bytes > ((size_t) 1 << 30) ? want : std::max<size_t>(want + want / 2, (size_t) 1 << 20));
}

struct MinePool {
    // trie nodes (grown by doubling, contents preserved)
    size_t node_cap = 0;
    MineNodes nd = {};
    DevBuf<int> list_id[2], ext[2];
    DevBuf<double> list_fr[2];
    DevBuf<long long> need, need_scan, packed, packed_scan;
    DevBuf<unsigned char> cub_tmp;
    DevBuf<MineTotals> tot;
    DevBuf<signed char> a0, a1, samples_raw;
    DevBuf<double> w;
    DevBuf<int> min_len, max_len, rows;
    // results
    DevBuf<int> o_start, o_end, o_node;
    DevBuf<double> o_freq, o_prefix, o_tp;
    DevBuf<long long> o_len, o_aloff, o_P;
    DevBuf<signed char> o_al;
    DevBuf<hmc_pattern_rec> o_rec;
    void *pin = nullptr;            // pinned host copy of the records + allele strings (hmc_map_patterns)
    size_t pin_bytes = 0;
    bool pin_valid = false;
    bool handed_over = false;       // hmc_build_automaton took the list's device arrays for the model (no copy): fetch / map are over
    long long P = 0, n_alleles = 0;
    int panel_serial = -1;          // a0/a1 hold the transposed panel of this hmc_set_panel call
    float ms = 0;
    long long nodes = 0, levels = 0, max_list = 0;
    long long swept = 0, written = 0;   // match-list entries read by the sweeps (count + fill) / written by the fill sweeps

    template <typename T>
    static cudaError_t regrow(T *&p, size_t old_n, size_t new_n)
    {
        T *q = nullptr;
        cudaError_t e = cudaMalloc((void **) &q, new_n * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p && old_n) e = cudaMemcpy(q, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice);
        if (p) cudaFree(p);
        p = q;
        return e;
    }
    cudaError_t reserve_nodes(size_t want, size_t live)
    {
        if (want <= node_cap) return cudaSuccess;
        size_t cap = node_cap ? node_cap : (size_t) 1 << 20;
        while (cap < want) cap *= 2;
        cudaError_t e;
#define HMC_REGROW(f) if ((e = regrow(nd.f, live, cap)) != cudaSuccess) return e
        HMC_REGROW(parent); HMC_REGROW(start); HMC_REGROW(cnt); HMC_REGROW(cbase); HMC_REGROW(sub); HMC_REGROW(idb);
        HMC_REGROW(len); HMC_REGROW(flags); HMC_REGROW(al); HMC_REGROW(freq); HMC_REGROW(loff);
#undef HMC_REGROW
        node_cap = cap;
        return cudaSuccess;
    }
    void release()
    {
        cudaFree(nd.parent); cudaFree(nd.start); cudaFree(nd.cnt); cudaFree(nd.cbase); cudaFree(nd.sub); cudaFree(nd.idb);
        cudaFree(nd.len); cudaFree(nd.flags); cudaFree(nd.al); cudaFree(nd.freq); cudaFree(nd.loff);
        nd = {}; node_cap = 0;
        for (int i = 0; i < 2; ++i) { list_id[i].release(); list_fr[i].release(); ext[i].release(); }
        need.release(); need_scan.release(); packed.release(); packed_scan.release(); cub_tmp.release(); tot.release();
        a0.release(); a1.release(); samples_raw.release(); rows.release(); w.release(); min_len.release(); max_len.release();
        o_start.release(); o_end.release(); o_node.release(); o_freq.release(); o_prefix.release(); o_tp.release();
        o_len.release(); o_aloff.release(); o_P.release(); o_al.release(); o_rec.release();
        if (pin) cudaFreeHost(pin);
        pin = nullptr; pin_bytes = 0;
    }
};

static int mine_scan(hmc_ctx *ctx, MinePool &mp, const long long *in, long long *out, size_t n)
{
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (long long) n, ctx->stream));
    CK(grow(mp.cub_tmp, bytes + 16));
    CK(cub::DeviceScan::ExclusiveSum(mp.cub_tmp.p, bytes, in, out, (long long) n, ctx->stream));
    return HMC_OK;
}

static int run_mine(hmc_ctx *ctx, MinePool &mp, int n_samples, const int8_t *sample_haps, const int32_t *sample_rows, const double *sample_w, double total_weight,
                    double min_freq, const int32_t *min_len, const int32_t *max_len)
{
    const int L = ctx->L, N = ctx->N;
    const bool samples = n_samples > 0;
    const int n_items = samples ? n_samples : N;
    // the automaton build keeps rolling-hash powers for strings of up to AUT_MAX_LEN alleles (automaton.cuh); a pattern can be no longer
    // than the loci that are left, so only an effective limit above that is refused
    for (int l = 0; l < L; ++l)
        if (std::min(max_len[l], L - l) > AUT_MAX_LEN)
            return ctx->fail(HMC_ERR_LIMIT, "max pattern length %d at locus %d exceeds the device limit of %d", max_len[l], l, AUT_MAX_LEN);
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], st));
    // locus-major allele matrix
    const dim3 tb(32, 8);
    const dim3 tg((unsigned) ((n_items + 31) / 32), (unsigned) ((L + 31) / 32));
    if (samples) {
        CK(grow(mp.a0, (size_t) n_items * L)); CK(grow(mp.w, (size_t) n_items));
        CK(cudaMemcpyAsync(mp.w.p, sample_w, sizeof(double) * n_items, cudaMemcpyHostToDevice, st));
        if (sample_rows) {      // rows of the haplotype results the last resolve pass left on the device
            CK(grow(mp.rows, (size_t) n_items));
            CK(cudaMemcpyAsync(mp.rows.p, sample_rows, sizeof(int) * n_items, cudaMemcpyHostToDevice, st));
            mine_transpose_rows_kernel<<<tg, tb, 0, st>>>(ctx->d_haps.p, mp.rows.p, n_items, L, mp.a0.p);
            CK(cudaGetLastError());
            CK(cudaGetLastError());
        }
        else {
            CK(grow(mp.samples_raw, (size_t) n_items * L));
            CK(cudaMemcpyAsync(mp.samples_raw.p, sample_haps, (size_t) n_items * L, cudaMemcpyHostToDevice, st));
            mine_transpose_kernel<<<tg, tb, 0, st>>>(mp.samples_raw.p, n_items, L, 1, mp.a0.p, nullptr);
            CK(cudaGetLastError());
            CK(cudaGetLastError());
        }
        mp.panel_serial = -1;
    }
    else if (mp.panel_serial != ctx->panel_serial) {
        CK(mp.a0.ensure((size_t) N * L)); CK(mp.a1.ensure((size_t) N * L));
        mine_transpose_kernel<<<tg, tb, 0, st>>>(ctx->d_raw.p, N, L, 2, mp.a0.p, mp.a1.p);
        CK(cudaGetLastError());
        CK(cudaGetLastError());
        mp.panel_serial = ctx->panel_serial;
    }
    CK(mp.min_len.ensure(L)); CK(mp.max_len.ensure(L)); CK(mp.tot.ensure(1)); CK(mp.o_P.ensure(1));
    CK(cudaMemcpyAsync(mp.min_len.p, min_len, sizeof(int) * L, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(mp.max_len.p, max_len, sizeof(int) * L, cudaMemcpyHostToDevice, st));
    // The tries of different start loci are independent, and a level's match lists hold every (item, start) pair as often as the item
    // matches a candidate of that start (~100x at the deep levels of a long panel): the start loci are mined in batches of bounded
    // (item, start) pairs so that the lists of a level fit the device whatever the panel's size.  Batches go from the highest start
    // down -- the reference releases its candidate stack from the back (PatternManager.cpp:106-142), so their patterns are consecutive
    // id ranges and the batch results are appended.
    long long pair_budget = 20000000;
    {   // the lists of a deep level hold ~100 entries of 12 bytes per pair, twice (ping-pong): keep them within a quarter of what is free
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) pair_budget = std::min<long long>(pair_budget, std::max<long long>(1000000, (long long) (fr / 4 / 2400)));
    }
    if (const char *env = getenv("HMC_B200_MINE_PAIRS")) pair_budget = std::max(1ll, atoll(env));
    const int batch = (int) std::max<long long>(1, std::min<long long>(L, pair_budget / std::max(1, n_items)));
    CK(mp.reserve_nodes(std::max<size_t>((size_t) std::min(L, batch), (size_t) 1 << 23), 0));

    MineParams m = {};
    auto refresh = [&]() {
        m.n_items = n_items; m.L = L; m.Amax = ctx->Amax;
        m.a0 = mp.a0.p; m.a1 = mp.a1.p; m.w = mp.w.p; m.af = ctx->d_af.p; m.linfo = ctx->d_linfo.p;
        m.min_len = mp.min_len.p; m.max_len = mp.max_len.p;
        m.denom = samples ? total_weight : (double) N; m.min_freq = min_freq;
        m.nd = mp.nd;
        for (int i = 0; i < 2; ++i) { m.list_id[i] = mp.list_id[i].p; m.list_fr[i] = mp.list_fr[i].p; }
    };
    // output arrays grow batch by batch, keeping what earlier batches appended
    auto keep = [&](auto &buf, size_t live, size_t want) -> cudaError_t {
        if (want <= buf.n) return cudaSuccess;
        typedef typename std::remove_reference<decltype(*buf.p)>::type T;
        const size_t cap = want * sizeof(T) > ((size_t) 1 << 30) ? want : std::max<size_t>(want + want / 4, (size_t) 1 << 20);       // (no headroom once it is large)
        T *q = nullptr;
        cudaError_t e = cudaMalloc((void **) &q, cap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (buf.p && live) e = cudaMemcpyAsync(q, buf.p, live * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (buf.p) cudaFree(buf.p);
        buf.p = q; buf.n = cap;
        return e;
    };
    mp.max_list = 0; mp.swept = 0; mp.written = 0; mp.nodes = 0; mp.levels = 0;
    long long P_total = 0, al_total = 0;
    for (int s1 = L; s1 > 0; s1 -= batch) {
    const int s0 = std::max(0, s1 - batch), nroots = s1 - s0;
    {   // first guess for the match lists: every item once per start locus (exact for fully resolved samples)
        const size_t guess = std::min<size_t>((size_t) n_items * nroots * (samples ? 1 : 4) + 1024, (size_t) 1 << 30);
        for (int i = 0; i < 2; ++i) { CK(grow(mp.list_id[i], guess)); if (!samples) CK(grow(mp.list_fr[i], guess)); }
    }
    std::vector<long long> level_begin;
    long long node_end = 0;
    int par = 0, epar = 0, prev_ext = 0;
    long long c0 = 0, n = nroots;       // the level whose nodes were just created (level 0: the roots)
    long long prev_in = 0;              // entries in the lists of the current parents
    for (int d = 0;; ++d) {
        level_begin.push_back(c0);
        node_end = c0 + n;
        CK(grow(mp.need, n)); CK(grow(mp.need_scan, n)); CK(grow(mp.packed, n)); CK(grow(mp.packed_scan, n));
        CK(grow(mp.ext[epar ^ 1], n));
        refresh();
        m.s0 = s0; m.nroots = nroots;
        const unsigned gb = (unsigned) ((n + 255) / 256);
        if (d == 0) mine_roots_kernel<<<gb, 256, 0, st>>>(m, mp.need.p, mp.packed.p);
        else mine_decide_kernel<<<gb, 256, 0, st>>>(m, (int) c0, (int) n, mp.need.p, mp.packed.p);
        if (cudaError_t e_ = cudaGetLastError()) return ctx->fail(HMC_ERR_CUDA, "mine_roots / mine_decide launch (level %d, grid %u, n %lld): %s", d, gb, n, cudaGetErrorString(e_));
        int rc = mine_scan(ctx, mp, mp.need.p, mp.need_scan.p, n);
        if (rc) return rc;
        rc = mine_scan(ctx, mp, mp.packed.p, mp.packed_scan.p, n);
        if (rc) return rc;
        mine_link_kernel<<<gb, 256, 0, st>>>(m, (int) c0, (int) n, mp.need.p, mp.need_scan.p, mp.packed.p, mp.packed_scan.p, mp.ext[epar ^ 1].p, (int) node_end, mp.tot.p);
        CK(cudaGetLastError());
        MineTotals t;
        CK(cudaMemcpyAsync(&t, mp.tot.p, sizeof t, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        mp.max_list = std::max(mp.max_list, t.list_total);
        // lists of this level's extended nodes (their parents' lists are in `par`)
        if (d > 0 && t.list_total > 0) {
            CK(grow(mp.list_id[par ^ 1], (size_t) t.list_total));
            if (!samples) CK(grow(mp.list_fr[par ^ 1], (size_t) t.list_total));
            refresh();
            m.s0 = s0; m.nroots = nroots;
            m.ext = mp.ext[epar].p;
            const int n_par = prev_ext;
            const unsigned gw = (unsigned) std::min<long long>(((long long) n_par * 32 + 255) / 256, (long long) ctx->sms * 32);
            if (samples) mine_sweep_kernel<true, true><<<gw, 256, 0, st>>>(m, n_par, par);
            else mine_sweep_kernel<false, true><<<gw, 256, 0, st>>>(m, n_par, par);
            if (cudaError_t e_ = cudaGetLastError()) return ctx->fail(HMC_ERR_CUDA, "mine_sweep<FILL> launch (grid %u, %d parents): %s", gw, n_par, cudaGetErrorString(e_));
            mp.swept += prev_in; mp.written += t.list_total;
        }
        if (d > 0) par ^= 1;
        epar ^= 1;
        if (t.n_ext == 0 || t.n_child == 0) break;
        if (t.n_ext > 0x7FFFFFF0ll / std::max(1, ctx->Amax)) return ctx->fail(HMC_ERR_LIMIT, "pattern trie exceeds 2^31 nodes");     // (keeps the packed child count below 2^32)
        if (node_end + t.n_child > 0x7FFFFFF0ll) return ctx->fail(HMC_ERR_LIMIT, "pattern trie exceeds 2^31 nodes");
        CK(mp.reserve_nodes((size_t) (node_end + t.n_child), (size_t) node_end));
        refresh();
        m.s0 = s0; m.nroots = nroots;
        m.ext = mp.ext[epar].p;
        const unsigned gw = (unsigned) std::min<long long>((t.n_ext * 32 + 255) / 256, (long long) ctx->sms * 32);
        if (samples) mine_sweep_kernel<true, false><<<gw, 256, 0, st>>>(m, (int) t.n_ext, par);
        else mine_sweep_kernel<false, false><<<gw, 256, 0, st>>>(m, (int) t.n_ext, par);
        if (cudaError_t e_ = cudaGetLastError()) return ctx->fail(HMC_ERR_CUDA, "mine_sweep<COUNT> launch (grid %u, %lld parents): %s", gw, (long long) t.n_ext, cudaGetErrorString(e_));
        prev_ext = (int) t.n_ext;
        prev_in = d == 0 ? t.n_ext * (long long) n_items : t.list_total;
        mp.swept += prev_in;
        c0 = node_end; n = t.n_child;
    }
    level_begin.push_back(node_end);
    mp.nodes += node_end; mp.levels = std::max<long long>(mp.levels, (long long) level_begin.size() - 1);
    // ids: subtree sizes bottom-up, bases top-down
    refresh();
    m.s0 = s0; m.nroots = nroots;
    for (int d = (int) level_begin.size() - 2; d >= 0; --d) {
        const long long b = level_begin[d], cnt = level_begin[d + 1] - b;
        if (cnt > 0) mine_sub_kernel<<<(unsigned) ((cnt + 255) / 256), 256, 0, st>>>(m, (int) b, (int) cnt);
    }
    CK(mp.o_P.ensure(1));
    mine_root_ids_kernel<<<1, 32, 0, st>>>(m, mp.o_P.p);
    CK(cudaGetLastError());
    for (int d = 0; d + 1 < (int) level_begin.size(); ++d) {
        const long long b = level_begin[d], cnt = level_begin[d + 1] - b;
        if (cnt > 0) mine_ids_kernel<<<(unsigned) ((cnt + 255) / 256), 256, 0, st>>>(m, (int) b, (int) cnt);
    }
    long long P = 0;
    CK(cudaMemcpyAsync(&P, mp.o_P.p, sizeof P, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (P_total + P > 0x7FFFFFF0ll) return ctx->fail(HMC_ERR_LIMIT, "more than 2^31 patterns");
    if (P > 0) {        // this batch's patterns are ids [P_total, P_total + P)
        const size_t Pt = (size_t) P_total, Pn = (size_t) (P_total + P);
        CK(keep(mp.o_start, Pt, Pn)); CK(keep(mp.o_end, Pt, Pn)); CK(keep(mp.o_freq, Pt, Pn)); CK(keep(mp.o_prefix, Pt, Pn)); CK(keep(mp.o_tp, Pt, Pn));
        CK(keep(mp.o_aloff, Pt + 1, Pn + 1));
        CK(grow(mp.o_node, P)); CK(grow(mp.o_len, P));
        MineOut o = {mp.o_start.p + Pt, mp.o_end.p + Pt, mp.o_freq.p + Pt, mp.o_prefix.p + Pt, mp.o_tp.p + Pt, mp.o_len.p, mp.o_aloff.p + Pt, nullptr, mp.o_node.p};
        const long long cnt = node_end - nroots;
        if (cnt > 0) mine_emit_kernel<<<(unsigned) ((cnt + 255) / 256), 256, 0, st>>>(m, o, nroots, (int) cnt);
        int rc = mine_scan(ctx, mp, mp.o_len.p, mp.o_aloff.p + Pt, P);
        if (rc) return rc;
        if (al_total) mine_add_offset_kernel<<<(unsigned) ((P + 255) / 256), 256, 0, st>>>(mp.o_aloff.p + Pt, P, al_total);
        long long last_off = 0, last_len = 0;
        CK(cudaMemcpyAsync(&last_off, mp.o_aloff.p + Pt + (P - 1), 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&last_len, mp.o_len.p + (P - 1), 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const long long al_new = last_off + last_len;
        CK(keep(mp.o_al, (size_t) al_total, (size_t) al_new + 1));
        o.al = mp.o_al.p;
        mine_alleles_kernel<<<(unsigned) ((P + 255) / 256), 256, 0, st>>>(m, o, P);
        CK(cudaGetLastError());
        P_total += P; al_total = al_new;
    }
    }   // batches of start loci
    mp.P = P_total; mp.n_alleles = al_total; mp.pin_valid = false; mp.handed_over = false;
    if (P_total > 0) {
        MineOut o = {mp.o_start.p, mp.o_end.p, mp.o_freq.p, mp.o_prefix.p, mp.o_tp.p, nullptr, mp.o_aloff.p, mp.o_al.p, nullptr};
        CK(grow(mp.o_rec, (size_t) P_total));
        mine_records_kernel<<<(unsigned) ((P_total + 255) / 256), 256, 0, st>>>(o, mp.o_rec.p, P_total);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    cudaEventElapsedTime(&mp.ms, ctx->ev[0], ctx->ev[1]);
    {   // the match lists are scratch: keep them for the next search only while they are small next to the trellis arena
        size_t bytes = 0;
        for (int i = 0; i < 2; ++i) bytes += mp.list_id[i].n * sizeof(int) + mp.list_fr[i].n * sizeof(double);
        if (bytes > ((size_t) 4 << 30)) for (int i = 0; i < 2; ++i) { mp.list_id[i].release(); mp.list_fr[i].release(); }
        if (mp.node_cap * 43 > ((size_t) 2 << 30)) {     // the trie itself is not needed once the list is emitted
            cudaFree(mp.nd.parent); cudaFree(mp.nd.start); cudaFree(mp.nd.cnt); cudaFree(mp.nd.cbase); cudaFree(mp.nd.sub); cudaFree(mp.nd.idb);
            cudaFree(mp.nd.len); cudaFree(mp.nd.flags); cudaFree(mp.nd.al); cudaFree(mp.nd.freq); cudaFree(mp.nd.loff);
            mp.nd = {}; mp.node_cap = 0;
            mp.need.release(); mp.need_scan.release(); mp.packed.release(); mp.packed_scan.release(); mp.ext[0].release(); mp.ext[1].release();
            mp.o_node.release(); mp.o_len.release();
        }
    }
    return HMC_OK;
}

// ---------------------------------------------------------------------------------------------
// K6 (automaton.cuh): the automaton over the last mined pattern list, installed as the context's model without a host
// round trip (the device-side equivalent of PatternManager::initialize + hmc_set_model).
struct AutPool {
    DevBuf<u64> keys, bpow;
    DevBuf<int> vals, iota;
    DevBuf<unsigned char> flags, cub_tmp;
    DevBuf<int> n_sel;
    float ms = 0;
    void release() { keys.release(); bpow.release(); vals.release(); iota.release(); flags.release(); cub_tmp.release(); n_sel.release(); }
};

static int run_build_automaton(hmc_ctx *ctx, MinePool &mp, AutPool &ap, int head_len, int *out_n_heads)
{
    const long long P = mp.P;
    const int A = ctx->Amax, L = ctx->L;
    if (P <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_build_automaton: the mined pattern list is empty");
    if (P > 0x7FFFFFF0ll / std::max(1, A)) return ctx->fail(HMC_ERR_LIMIT, "too many patterns for 32-bit successor indexing");
    if (head_len < 1 || head_len > 8 || head_len > L) return ctx->fail(HMC_ERR_LIMIT, "head length %d outside [1,8]", head_len);
    cudaStream_t st = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], st));
    if (!ap.bpow.p) {
        std::vector<u64> pw(AUT_MAX_LEN + 1);
        pw[0] = 1;
        for (int t = 1; t <= AUT_MAX_LEN; ++t) pw[t] = pw[t - 1] * AUT_BASE;
        CK(ap.bpow.ensure(pw.size()));
        CK(cudaMemcpy(ap.bpow.p, pw.data(), pw.size() * sizeof(u64), cudaMemcpyHostToDevice));
    }
    size_t hsize = 1 << 16;
    while (hsize < (size_t) P * 2) hsize <<= 1;
    CK(grow(ap.keys, hsize)); CK(grow(ap.vals, hsize)); CK(grow(ap.flags, (size_t) P)); CK(ap.n_sel.ensure(1));
    CK(cudaMemsetAsync(ap.keys.p, 0xFF, hsize * sizeof(u64), st));
    // model arrays of the context
    if (mp.handed_over) return ctx->fail(HMC_ERR_INVALID, "hmc_build_automaton: this pattern list already became the model");
    ctx->have_model = false;
    ctx->ecache.valid = false;
    if ((size_t) P * 28 + (size_t) mp.n_alleles > ((size_t) 2 << 30)) {
        // large lists: their device arrays BECOME the model's (swap, no copy, nothing held twice): start, transition probability,
        // frequency, allele strings; fetch / map are over for this list
        mp.o_rec.release();
        std::swap(ctx->d_start, mp.o_start); std::swap(ctx->d_tp, mp.o_tp); std::swap(ctx->d_freq, mp.o_freq);
        std::swap(ctx->d_al_off, mp.o_aloff); std::swap(ctx->d_al, mp.o_al);
        mp.handed_over = true;
        mp.o_start.release(); mp.o_tp.release(); mp.o_freq.release(); mp.o_aloff.release(); mp.o_al.release();      // the previous model
    }
    else {      // small lists: copy, so that neither side reallocates from one EM iteration to the next
        CK(grow(ctx->d_start, (size_t) P)); CK(grow(ctx->d_tp, (size_t) P)); CK(grow(ctx->d_freq, (size_t) P));
        CK(grow(ctx->d_al_off, (size_t) P + 1)); CK(grow(ctx->d_al, (size_t) mp.n_alleles + 1));
        CK(cudaMemcpyAsync(ctx->d_start.p, mp.o_start.p, sizeof(int) * P, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_tp.p, mp.o_tp.p, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_freq.p, mp.o_freq.p, sizeof(double) * P, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_al_off.p, mp.o_aloff.p, sizeof(long long) * (P + 1), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_al.p, mp.o_al.p, (size_t) mp.n_alleles, cudaMemcpyDeviceToDevice, st));
    }
    CK(grow(ctx->d_succ16, (size_t) P * A)); CK(grow(ctx->d_succ, (size_t) P * A)); CK(grow(ctx->d_last, (size_t) P));
    AutParams a = {};
    a.P = P; a.L = L; a.Amax = A; a.start = ctx->d_start.p; a.end = mp.o_end.p; a.al_off = ctx->d_al_off.p; a.al = ctx->d_al.p; a.linfo = ctx->d_linfo.p;
    a.keys = ap.keys.p; a.vals = ap.vals.p; a.hmask = hsize - 1; a.succ = ctx->d_succ.p; a.last = ctx->d_last.p; a.bpow = ap.bpow.p;
    const unsigned gb = (unsigned) ((P + 255) / 256);
    aut_insert_kernel<<<gb, 256, 0, st>>>(a);
    aut_succ_kernel<<<gb, 256, 0, st>>>(a);
    // head list + lookup table
    int stride[8] = {0};
    long long size = 1;
    for (int j = head_len - 1; j >= 0; --j) { stride[j] = (int) size; size *= std::max(1, ctx->h_nall[j]); if (size > (1 << 20)) return ctx->fail(HMC_ERR_LIMIT, "head table too large"); }
    aut_head_flags_kernel<<<gb, 256, 0, st>>>(a, head_len, ap.flags.p);
    CK(grow(ctx->d_heads, (size_t) std::max<long long>(1, std::min<long long>(P, size))));
    CK(grow(ctx->d_head_lut, (size_t) size));
    {
        size_t bytes = 0;
        thrust::counting_iterator<int> ids(0);
        CK(cub::DeviceSelect::Flagged(nullptr, bytes, ids, ap.flags.p, ctx->d_heads.p, ap.n_sel.p, (int) P, st));
        CK(grow(ap.cub_tmp, bytes + 16));
        CK(cub::DeviceSelect::Flagged(ap.cub_tmp.p, bytes, ids, ap.flags.p, ctx->d_heads.p, ap.n_sel.p, (int) P, st));
    }
    int n_heads = 0;
    CK(cudaMemcpyAsync(&n_heads, ap.n_sel.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(ctx->d_head_lut.p, 0xFF, sizeof(int) * (size_t) size, st));
    CK(cudaStreamSynchronize(st));
    if (n_heads > size) return ctx->fail(HMC_ERR_INVALID, "more head patterns than head strings");
    HeadStride hs;
    for (int j = 0; j < 8; ++j) hs.s[j] = stride[j];
    if (n_heads > 0) aut_head_lut_kernel<<<(unsigned) ((n_heads + 255) / 256), 256, 0, st>>>(a, ctx->d_heads.p, n_heads, head_len, hs, ctx->d_head_lut.p);
    {
        const size_t n = (size_t) P * A;
        succ16_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(ctx->d_succ.p, ctx->d_tp.p, n, (int) P, ctx->d_succ16.p, nullptr);
    }
    CK(cudaEventRecord(ctx->ev[1], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    cudaEventElapsedTime(&ap.ms, ctx->ev[0], ctx->ev[1]);
    DevModel &m = ctx->dm;
    m.P = (int) P; m.head_len = head_len; m.n_heads = n_heads; m.Amax = A;
    m.succ = ctx->d_succ.p; m.succ16 = ctx->d_succ16.p; m.tp = ctx->d_tp.p; m.freq = ctx->d_freq.p; m.start = ctx->d_start.p; m.last = ctx->d_last.p;
    m.al_off = ctx->d_al_off.p; m.al = ctx->d_al.p; m.heads = ctx->d_heads.p; m.head_lut = ctx->d_head_lut.p;
    for (int j = 0; j < 8; ++j) m.head_stride[j] = stride[j];
    ctx->have_model = true;
    if (out_n_heads) *out_n_heads = n_heads;
    return HMC_OK;
}
