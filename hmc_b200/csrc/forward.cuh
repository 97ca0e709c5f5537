// K1 forward_resolve: the per-individual forward pass of HaploBuilder::resolve
// (HaploBuilder.cpp:35-86 with initHeadList :153-224, extendAll/extend/addHaploPair :226-261 and the
// HaploPair constructors / add, HaploPair.cpp:14-89).
//
// One CTA (NT threads, NT = 32 is one warp) owns one individual at a time and pulls individuals from
// a work queue (persistent grid).  Per locus step the lanes are spread over the candidate slots
// (predecessor state x ordered allele pair); the step is
//   A  successor-table gathers + canonicalise (id_a <= id_b) + insert into an open-addressing hash
//      on (id_a,id_b); the slot remembers the smallest candidate ordinal that hit it
//   B  walk the candidates in ordinal order; the one whose ordinal equals the slot's minimum creates
//      the state -> new states get exactly the reference's creation order (block prefix sums)
//   C  prefix-sum the per-state arrival counts          D  scatter arrivals into per-state segments
//   E  one lane per new state: arrivals in ordinal order -> alpha sum in the reference's order
//      (no FMA), K-best merge of the predecessors' link lists, het-locus rule, trellis write.
// The hash, candidate and segment arrays live in shared memory while the step fits (C <= CC slots);
// otherwise the same code runs on the CTA's global spill scratch.
#pragma once

#include "common.cuh"

namespace hmc {

template <int NT>
struct BlockScan {
    // exclusive prefix sum over the block in thread order; total returned to every thread
    __device__ static long long run(long long v, long long &total, long long *s_tmp)
    {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (NT == 32) { total = __shfl_sync(0xffffffffu, x, 31); return x - v; }
        __syncthreads();            // protect s_tmp from the previous use
        if (lane == 31) s_tmp[wid] = x;
        __syncthreads();
        long long base = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) { long long t = s_tmp[w]; if (w < wid) base += t; tot += t; }
        total = tot;
        return base + x - v;
    }
};

template <int NT>
__device__ __forceinline__ void block_sync()
{
    if (NT == 32) __syncwarp(); else __syncthreads();
}

// sorted-descending insert with static indexing (registers); slot KMAX-1 falls off the end.
template <int KMAX>
__device__ __forceinline__ void topk_insert(double (&bl)[KMAX], u32 (&bb)[KMAX], double v, u32 b)
{
#pragma unroll
    for (int j = KMAX - 1; j >= 1; --j) {
        const bool shift = v > bl[j - 1];
        const bool place = !shift && (v > bl[j]);
        if (shift) { bl[j] = bl[j - 1]; bb[j] = bb[j - 1]; }
        else if (place) { bl[j] = v; bb[j] = b; }
    }
    if (v > bl[0]) { bl[0] = v; bb[0] = b; }
}

struct StepView {       // pointers of one step; shared or global (generic addressing)
    u64 *keys;
    u32 *ord, *cnt;
    int *cand, *arr, *aoff, *slot_of;
    int hmask;
};

template <int NT, int CC, int KMAX>
__global__ void __launch_bounds__(NT) forward_kernel(FwdParams p)
{
    constexpr int HC = 2 * CC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_keys = (u64 *) smem_raw;                         // [HC]
    double *s_alpha = (double *) (s_keys + HC);             // [2][CC]
    u32 *s_ord = (u32 *) (s_alpha + 2 * CC);                // [HC]
    u32 *s_cnt = s_ord + HC;                                // [HC]
    int *s_cand = (int *) (s_cnt + HC);                     // [CC]
    int *s_arr = s_cand + CC;                               // [CC]
    int *s_aoff = s_arr + CC;                               // [CC+1]
    int *s_slot_of = s_aoff + CC + 1;                       // [CC]
    int *s_ida = s_slot_of + CC;                            // [2][CC]
    int *s_idb = s_ida + 2 * CC;                            // [2][CC]
    int *s_gx = s_idb + 2 * CC;                             // [MAX_GROUPS]
    int *s_gy = s_gx + MAX_GROUPS;
    int *s_gbase = s_gy + MAX_GROUPS;                       // [MAX_GROUPS+1]
    __shared__ long long s_ll[4];                           // work item, alloc result, chunk cursor/end
    __shared__ long long s_tmp[8];                          // scan scratch
    __shared__ int s_int[4];

    const int tid = threadIdx.x;
    const DevPanel &pn = p.panel;
    const DevModel &md = p.model;
    const int L = pn.L, A = md.Amax, K = p.K, head_len = md.head_len;
    const Scratch sc = p.scratch[blockIdx.x];
    double *const sc_like0 = sc.like[0], *const sc_like1 = sc.like[1];
    u64 *const sc_meta0 = sc.meta[0], *const sc_meta1 = sc.meta[1];
    u64 *const arena = p.arena.base;

    if (tid == 0) { s_ll[2] = 0; s_ll[3] = 0; }

    // chunked bump allocation (thread 0 only); returns -1 when the arena is exhausted
    auto alloc = [&](long long units) -> long long {
        constexpr long long CHUNK = 1 << 15;
        units = (units + 1) & ~1LL;
        if (s_ll[2] + units > s_ll[3]) {
            const long long grab = units > CHUNK ? units : CHUNK;
            const unsigned long long base = atomicAdd(p.arena.cursor, (unsigned long long) grab);
            if ((long long) (base + grab) > p.arena.capacity) { atomicExch(p.arena.exhausted, 1); return -1; }
            s_ll[2] = (long long) base; s_ll[3] = (long long) base + grab;
        }
        const long long off = s_ll[2];
        s_ll[2] += units;
        return off;
    };

    for (;;) {
        block_sync<NT>();
        if (tid == 0) s_int[0] = atomicAdd(p.queue, 1);
        block_sync<NT>();
        const int w = s_int[0];
        if (w >= p.n_wave) break;
        if (*(volatile int *) p.arena.exhausted) { if (tid == 0) p.status[w] = ST_NOMEM; continue; }
        const int ind = p.wave[w];
        const u32 *gp = pn.packed + (size_t) ind * pn.words;
        LocusRec *rec = p.loci + (size_t) w * (L + 1);
        for (int l = tid; l <= L; l += NT) { rec[l].count = 0; rec[l].fwd_slots = 0; rec[l].off = 0; rec[l].fwd_off = 0; }
        auto geno = [&](int l, int &g0, int &g1) {
            const int bit = l * 2 * pn.bits;
            const u32 wd = gp[bit >> 5] >> (bit & 31);
            const u32 m = (1u << pn.bits) - 1u;
            g0 = (int) (wd & m) - 1;
            g1 = (int) ((wd >> pn.bits) & m) - 1;
        };

        // ---------------- head states (initHeadList), serial: at most A^head_len pairs ----------------
        int S_in = 0, par = 0, status = ST_DONE, cur_exp = 0;
        long long cur_off = 0;
        if (tid == 0) {
            int nh = 0, bad = 0;
            int hg0[8], hg1[8];
            for (int j = 0; j < head_len; ++j) geno(j, hg0[j], hg1[j]);
            for (int hi = 0; hi < md.n_heads && !bad; ++hi) {
                const int head = md.heads[hi];
                const signed char *ha = md.al + md.al_off[head];
                bool match = true;
                for (int j = 0; j < head_len; ++j) {
                    const int a = ha[j];
                    if (!(hg0[j] < 0 || hg1[j] < 0 || hg0[j] == a || hg1[j] == a)) { match = false; break; }
                }
                if (!match) continue;
                // odometer over the complementary allele strings, locus 0 most significant
                int lo[8], hi_[8], curv[8];
                for (int j = 0; j < head_len; ++j) {
                    const int a = ha[j];
                    const bool both = hg0[j] < 0 && hg1[j] < 0, any = hg0[j] < 0 || hg1[j] < 0;
                    if (both || (any && (hg0[j] == a || hg1[j] == a))) { lo[j] = 0; hi_[j] = pn.nall[j] - 1; }
                    else if (!any && hg0[j] != hg1[j]) { lo[j] = hi_[j] = (a == hg0[j]) ? hg1[j] : hg0[j]; }
                    else { const int v = hg0[j] < 0 ? 0 : hg0[j]; lo[j] = hi_[j] = v; }   // missing -> wildcard lookup -> first child
                    curv[j] = lo[j];
                }
                for (;;) {
                    bool ok = true;
                    int code = 0;
                    for (int j = 0; j < head_len; ++j) {
                        if (lo[j] != hi_[j] && !((pn.amask[j] >> curv[j]) & 1u)) ok = false;
                        code += curv[j] * md.head_stride[j];
                    }
                    if (ok) {
                        const int hp = md.head_lut[code];
                        if (hp < 0) { bad = 1; break; }
                        if (hp >= head) {
                            if (nh < CC) { s_ida[nh] = head; s_idb[nh] = hp; }
                            if (nh < sc.scap && 2 * nh + 1 < sc.ccap) { sc.cand[2 * nh] = head; sc.cand[2 * nh + 1] = hp; }    // staging when nh >= CC
                            ++nh;
                        }
                    }
                    int j = head_len - 1;
                    while (j >= 0) { if (++curv[j] <= hi_[j]) break; curv[j] = lo[j]; --j; }
                    if (j < 0) break;
                }
            }
            long long off = -1;
            int st = ST_DONE;
            if (bad) st = ST_NO_HEAD;
            else if (nh > sc.scap || 2 * nh > sc.ccap) st = ST_TOO_BIG;
            else if (nh > 0) { off = alloc(block_layout(nh, K, p.estep).total); if (off < 0) st = ST_NOMEM; }
            s_int[1] = nh; s_int[2] = st; s_ll[1] = off;
        }
        block_sync<NT>();
        S_in = s_int[1]; status = s_int[2]; cur_off = s_ll[1];
        if (status != ST_DONE) { if (tid == 0) p.status[w] = status; continue; }
        {
            const BlockLayout bl = block_layout(S_in, K, p.estep);
            int *t_ida = (int *) (arena + cur_off + bl.ida), *t_idb = (int *) (arena + cur_off + bl.idb);
            double *t_alpha = (double *) (arena + cur_off + bl.alpha);
            unsigned char *t_ab = (unsigned char *) (arena + cur_off + bl.ab);
            u32 *t_bp = (u32 *) (arena + cur_off + bl.bp);
            double *t_tp = (double *) (arena + cur_off + bl.tp);
            for (int s = tid; s < S_in; s += NT) {
                const int a = s < CC ? s_ida[s] : sc.cand[2 * s], b = s < CC ? s_idb[s] : sc.cand[2 * s + 1];
                const double f = __dmul_rn(md.freq[a], md.freq[b]);
                const double al = (a == b) ? f : __dmul_rn(f, 2.0);
                if (s < CC) s_alpha[s] = al;
                t_ida[s] = a; t_idb[s] = b; t_alpha[s] = al;
                t_ab[s] = (unsigned char) (md.last[a] | (md.last[b] << 4));
                if (p.estep) t_tp[s] = f;
                else t_bp[(size_t) s * K] = BP_HEAD | (a == b ? BP_HOM : 0u);
                sc_like0[(size_t) s * K] = f;
                sc_meta0[s] = 1ull | ((a == b ? 1ull : 0ull) << 8);
            }
            if (tid == 0) { rec[head_len].off = cur_off; rec[head_len].count = S_in; }
        }
        par = 0;
        long long c_steps = 0, c_states = S_in, c_max = S_in, c_arr = 0, c_spill = 0;
        if (p.locus_exp && tid == 0) for (int l = 0; l <= head_len; ++l) p.locus_exp[(size_t) w * (L + 1) + l] = 0;

        // ---------------- locus steps ----------------
        for (int i = head_len; i < L && S_in > 0; ++i) {
            int g0, g1;
            geno(i, g0, g1);
            const int nA = pn.nall[i];
            const u32 am = pn.amask[i];
            // allele-pair groups in the reference's extendAll order (HaploBuilder.cpp:48-79)
            int ng = 0, C = 0;
            auto add_group = [&](int x, int y) {
                if (tid == 0) { s_gx[ng] = x; s_gy[ng] = y; s_gbase[ng] = C; }
                C += S_in * (x != y ? 2 : 1);
                ++ng;
            };
            if (g0 < 0 && g1 < 0) { for (int j = 0; j < nA; ++j) if ((am >> j) & 1u) for (int k = j; k < nA; ++k) if ((am >> k) & 1u) add_group(j, k); }
            else if (g0 < 0) { for (int j = 0; j < nA; ++j) if ((am >> j) & 1u) add_group(j, g1); }
            else if (g1 < 0) { for (int j = 0; j < nA; ++j) if ((am >> j) & 1u) add_group(j, g0); }
            else add_group(g0, g1);
            if (tid == 0) s_gbase[ng] = C;
            int hsize = 64;
            while (hsize < 2 * C) hsize <<= 1;
            const bool fast = (C <= CC);
            if (!fast && (C > sc.ccap || hsize > sc.hcap || C > sc.scap)) { status = ST_TOO_BIG; break; }
            StepView v;
            if (fast) { v.keys = s_keys; v.ord = s_ord; v.cnt = s_cnt; v.cand = s_cand; v.arr = s_arr; v.aoff = s_aoff; v.slot_of = s_slot_of; }
            else { v.keys = sc.keys; v.ord = sc.ord; v.cnt = sc.cnt; v.cand = sc.cand; v.arr = sc.arr; v.aoff = sc.aoff; v.slot_of = sc.slot_of; ++c_spill; }
            v.hmask = hsize - 1;
            for (int h = tid; h < hsize; h += NT) { v.keys[h] = KEY_EMPTY; v.ord[h] = 0xFFFFFFFFu; v.cnt[h] = 0; }

            // input state views
            const BlockLayout bli = block_layout(S_in, K, p.estep);
            const int *gi_ida = (const int *) (arena + cur_off + bli.ida), *gi_idb = (const int *) (arena + cur_off + bli.idb);
            const double *gi_alpha = (const double *) (arena + cur_off + bli.alpha);
            const bool in_cache = S_in <= CC;
            const int *ci_ida = s_ida + par * CC, *ci_idb = s_idb + par * CC;
            const double *ci_alpha = s_alpha + par * CC;
            const double *like_in = par ? sc_like1 : sc_like0;
            const u64 *meta_in = par ? sc_meta1 : sc_meta0;
            double *like_out = par ? sc_like0 : sc_like1;
            u64 *meta_out = par ? sc_meta0 : sc_meta1;

            // forward-link block of index i (E-step mode)
            long long fwd_off = 0;
            if (p.estep) {
                if (tid == 0) { s_ll[1] = alloc(((long long) C + 1) / 2); }
            }
            block_sync<NT>();       // hash cleared, groups visible, fwd allocation visible
            if (p.estep) {
                fwd_off = s_ll[1];
                if (fwd_off < 0) { status = ST_NOMEM; break; }
                if (tid == 0) { rec[i].fwd_off = fwd_off; rec[i].fwd_slots = C; }
            }
            int *fwd = (int *) (arena + fwd_off);

            auto decode = [&](int c, int &g, int &s, int &dir) {
                g = 0;
                while (g + 1 < ng && c >= s_gbase[g + 1]) ++g;
                const int r = c - s_gbase[g];
                if (s_gx[g] != s_gy[g]) { s = r >> 1; dir = r & 1; } else { s = r; dir = 0; }
            };

            // ---- pass A: gather successors, canonicalise, hash-insert ----
            for (int c = tid; c < C; c += NT) {
                int g, s, dir;
                decode(c, g, s, dir);
                const int x = dir ? s_gy[g] : s_gx[g], y = dir ? s_gx[g] : s_gy[g];
                const int pa = in_cache ? ci_ida[s] : gi_ida[s], pb = in_cache ? ci_idb[s] : gi_idb[s];
                const double al = in_cache ? ci_alpha[s] : gi_alpha[s];
                int enc = -1;
                if (al > 0.0) {
                    int hpa = __ldg(md.succ + (size_t) pa * A + x);
                    int hpb = __ldg(md.succ + (size_t) pb * A + y);
                    if (hpa >= 0 && hpb >= 0) {
                        u32 rev = 0;
                        if (hpa > hpb) { const int t = hpa; hpa = hpb; hpb = t; rev = 1; }
                        const u64 key = ((u64) (u32) hpa << 32) | (u32) hpb;
                        u32 h = mix64(key) & v.hmask;
                        for (;;) {
                            const u64 prev = atomicCAS(v.keys + h, KEY_EMPTY, key);
                            if (prev == KEY_EMPTY || prev == key) break;
                            h = (h + 1) & v.hmask;
                        }
                        atomicMin(v.ord + h, (u32) c);
                        atomicAdd(v.cnt + h, 1u);
                        enc = (int) (h | (rev << 30));
                    }
                }
                v.cand[c] = enc;
                if (p.estep && enc < 0) fwd[c] = -1;
            }
            block_sync<NT>();

            // ---- pass B+C: creators in ordinal order -> new state indices and arrival-segment offsets ----
            int S_out = 0, n_arr = 0;
            for (int base = 0; base < C; base += NT) {
                const int c = base + tid;
                int creator = 0, h = 0, n = 0;
                if (c < C) {
                    const int enc = v.cand[c];
                    if (enc >= 0) { h = enc & 0x3FFFFFFF; creator = (v.ord[h] == (u32) c); if (creator) n = (int) v.cnt[h]; }
                }
                // one scan for both: creators in the high half, arrivals in the low half (C < 2^24 slots)
                long long tot;
                const long long ex = BlockScan<NT>::run(((long long) creator << 32) | (long long) n, tot, s_tmp);
                if (creator) {
                    const int idx = S_out + (int) (ex >> 32);
                    if (idx < sc.scap) {
                        v.ord[h] = 0x80000000u | (u32) idx;     // never equals a later ordinal
                        v.slot_of[idx] = h;
                        v.aoff[idx] = n_arr + (int) (ex & 0xFFFFFFFFll);
                        v.cnt[h] = 0;
                    }
                }
                S_out += (int) (tot >> 32);
                n_arr += (int) (tot & 0xFFFFFFFFll);
            }
            if (S_out == 0) {   // every extension died: m_haplopairs[i+1] empty, HaploBuilder.cpp:80-82
                S_in = 0;
                break;
            }
            if (S_out > sc.scap || S_out >= MAX_STATES_PER_LOCUS) { status = ST_TOO_BIG; break; }
            if (tid == 0) { s_ll[1] = alloc(block_layout(S_out, K, p.estep).total); v.aoff[S_out] = n_arr; }
            c_arr += n_arr;
            block_sync<NT>();
            const long long new_off = s_ll[1];
            if (new_off < 0) { status = ST_NOMEM; break; }

            // ---- pass D: scatter arrivals (unordered inside a segment; sorted in E) ----
            for (int c = tid; c < C; c += NT) {
                const int enc = v.cand[c];
                if (enc >= 0) {
                    const int h = enc & 0x3FFFFFFF;
                    const u32 rev = ((u32) enc >> 30) & 1u;
                    const int idx = (int) (v.ord[h] & 0x7FFFFFFFu);
                    const int pos = (int) atomicAdd(v.cnt + h, 1u);
                    v.arr[v.aoff[idx] + pos] = (int) ((u32) c | (rev << 31));
                    if (p.estep) fwd[c] = (int) ((u32) idx | (rev << 31));
                }
            }
            block_sync<NT>();

            // ---- pass E: one lane per new state ----
            const BlockLayout blo = block_layout(S_out, K, p.estep);
            int *t_ida = (int *) (arena + new_off + blo.ida), *t_idb = (int *) (arena + new_off + blo.idb);
            double *t_alpha = (double *) (arena + new_off + blo.alpha);
            unsigned char *t_ab = (unsigned char *) (arena + new_off + blo.ab);
            u32 *t_bp = (u32 *) (arena + new_off + blo.bp);
            double *t_tp = (double *) (arena + new_off + blo.tp);
            int *co_ida = s_ida + (par ^ 1) * CC, *co_idb = s_idb + (par ^ 1) * CC;
            double *co_alpha = s_alpha + (par ^ 1) * CC;
            const bool out_cache = S_out <= CC;
            double amax = 0.0;
            for (int idx = tid; idx < S_out; idx += NT) {
                const int h = v.slot_of[idx];
                const u64 key = v.keys[h];
                const int hpa = (int) (key >> 32), hpb = (int) (key & 0xFFFFFFFFu);
                const int a0 = v.aoff[idx], n = v.aoff[idx + 1] - a0;
                for (int j = 1; j < n; ++j) {           // ordinal order inside the segment
                    const int e = v.arr[a0 + j];
                    int q = j - 1;
                    while (q >= 0 && (v.arr[a0 + q] & 0x7FFFFFFF) > (e & 0x7FFFFFFF)) { v.arr[a0 + q + 1] = v.arr[a0 + q]; --q; }
                    v.arr[a0 + q + 1] = e;
                }
                const double tpv = __dmul_rn(__ldg(md.tp + hpa), __ldg(md.tp + hpb));
                double alpha = 0.0;
                double bl[KMAX];
                u32 bb[KMAX];
                int nl_out = 0, xa = 0, ya = 0;
                for (int j = 0; j < n; ++j) {
                    const int e = v.arr[a0 + j];
                    const u32 rev = (u32) e >> 31;
                    int g, s, dir;
                    decode(e & 0x7FFFFFFF, g, s, dir);
                    const bool het = (s_gx[g] != s_gy[g]);
                    if (j == 0) { const int x = dir ? s_gy[g] : s_gx[g], y = dir ? s_gx[g] : s_gy[g]; xa = rev ? y : x; ya = rev ? x : y; }
                    // the predecessor's whole list in one batch of independent loads (sorted descending)
                    const u64 mt = meta_in[s];
                    const int nl = (int) (mt & 0xFF);
                    u32 hom = (u32) (mt >> 8);
                    double x[KMAX];
                    const double *lp = like_in + (size_t) s * K;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) x[k] = (k < nl) ? lp[k] : 0.0;
                    const double ain = in_cache ? ci_alpha[s] : gi_alpha[s];
                    const double term = __dmul_rn(ain, tpv);
                    alpha = (j == 0) ? term : __dadd_rn(alpha, term);
                    nl_out += nl;
                    // het-locus rule (HaploPair.cpp:53-60,77-84): links still flagged homozygous lose the flag, and are
                    // zeroed when this arrival is reversed
                    const u32 zero = (het && rev) ? hom : 0u;
                    if (het) hom = 0;
                    const u32 bpbase = (u32) s | (rev ? BP_REV : 0u);
                    if (j == 0 && zero == 0) {
                        // scaling keeps the order: the first arrival's list is the initial top-K as it stands
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) {
                            bl[k] = (k < nl) ? __dmul_rn(x[k], tpv) : -1.0;
                            bb[k] = bpbase | ((u32) k << BP_PRED_BITS) | (((hom >> k) & 1u) ? BP_HOM : 0u);
                        }
                    }
                    else {
                        if (j == 0) {
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) { bl[k] = -1.0; bb[k] = 0; }
                        }
                        // sorted input: stop at the first element that does not make the list (zeroed ones are retried)
                        int k = 0;
                        while (k < nl) {
                            double lv = __dmul_rn(x[0], tpv);
                            if ((zero >> k) & 1u) lv = 0.0;
                            if (lv > bl[KMAX - 1]) topk_insert<KMAX>(bl, bb, lv, bpbase | ((u32) k << BP_PRED_BITS) | (((hom >> k) & 1u) ? BP_HOM : 0u));
                            else if (zero == 0) break;
#pragma unroll
                            for (int q = 0; q + 1 < KMAX; ++q) x[q] = x[q + 1];
                            ++k;
                        }
                    }
                }
                if (nl_out > K) nl_out = K;
                u32 hom_out = 0;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (k < nl_out) {
                        like_out[(size_t) idx * K + k] = bl[k];
                        if (!p.estep) t_bp[(size_t) idx * K + k] = bb[k];
                        if (bb[k] & BP_HOM) hom_out |= 1u << k;
                    }
                }
                meta_out[idx] = (u64) nl_out | ((u64) hom_out << 8);
                t_ida[idx] = hpa; t_idb[idx] = hpb; t_alpha[idx] = alpha;
                t_ab[idx] = (unsigned char) (xa | (ya << 4));
                if (p.estep) t_tp[idx] = tpv;
                if (out_cache) { co_ida[idx] = hpa; co_idb[idx] = hpb; co_alpha[idx] = alpha; }
                amax = fmax(amax, alpha);
            }
            int new_exp = cur_exp;
            if (p.scaled) {     // exponent-only rescale: mantissas stay those of the unscaled product
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                if (NT > 32) {
                    __syncthreads();
                    if ((tid & 31) == 0) ((double *) s_keys)[tid >> 5] = amax;    // hash is dead by now
                    __syncthreads();
                    amax = 0.0;
                    for (int q = 0; q < NT / 32; ++q) amax = fmax(amax, ((double *) s_keys)[q]);
                }
                if (amax > 0.0 && amax < 0x1p-512) {
                    block_sync<NT>();
                    for (int idx = tid; idx < S_out; idx += NT) {
                        const double a2 = t_alpha[idx] * 0x1p+512;
                        t_alpha[idx] = a2;
                        if (out_cache) co_alpha[idx] = a2;
                        const int nl = (int) (meta_out[idx] & 0xFF);
                        for (int k = 0; k < nl; ++k) like_out[(size_t) idx * K + k] *= 0x1p+512;
                    }
                    new_exp += 512;
                }
            }
            if (tid == 0) {
                rec[i + 1].off = new_off; rec[i + 1].count = S_out;
                if (p.locus_exp) p.locus_exp[(size_t) w * (L + 1) + i + 1] = new_exp;
            }
            cur_exp = new_exp;
            ++c_steps; c_states += S_out; if (S_out > c_max) c_max = S_out;
            S_in = S_out; par ^= 1; cur_off = new_off;
            block_sync<NT>();
        }
        block_sync<NT>();
        if (status == ST_DONE && S_in > 0 && !p.estep) {
            // final link lists for the traceback kernel (selection over all final states' links)
            if (tid == 0) s_ll[1] = alloc((long long) S_in * K + S_in);
            block_sync<NT>();
            const long long fo = s_ll[1];
            if (fo < 0) status = ST_NOMEM;
            else {
                double *fl = (double *) (arena + fo);
                u64 *fm = arena + fo + (long long) S_in * K;
                for (int s = tid; s < S_in; s += NT) {
                    const u64 mt = (par ? sc_meta1 : sc_meta0)[s];
                    fm[s] = mt;
                    const int nl = (int) (mt & 0xFF);
                    for (int k = 0; k < nl; ++k) fl[(size_t) s * K + k] = (par ? sc_like1 : sc_like0)[(size_t) s * K + k];
                }
                if (tid == 0) { rec[0].off = fo; rec[0].count = S_in; }     // index 0 is unused by the trellis when head_len >= 1
            }
        }
        if (tid == 0) {
            p.status[w] = status;
            p.final_exp[w] = cur_exp;
            if (status == ST_DONE) {
                atomicAdd((unsigned long long *) p.counters + 0, (unsigned long long) c_steps);
                atomicAdd((unsigned long long *) p.counters + 1, (unsigned long long) c_states);
                atomicMax((long long *) p.counters + 2, c_max);
                atomicAdd((unsigned long long *) p.counters + 3, (unsigned long long) c_arr);
                atomicAdd((unsigned long long *) p.counters + 5, (unsigned long long) c_spill);
            }
        }
    }
}

template <int CC>
constexpr size_t forward_smem_bytes()
{
    return (size_t) (2 * CC) * 8 + (size_t) 2 * CC * 8 + (size_t) 2 * (2 * CC) * 4 + (size_t) (CC + CC + CC + 1 + CC + 2 * CC + 2 * CC) * 4 +
           (size_t) (3 * MAX_GROUPS + 1 + 32) * 4 + 16;
}

}  // namespace hmc
