// K1 forward_resolve, warp-streaming variant: the per-individual forward pass of HaploBuilder::resolve
// (HaploBuilder.cpp:35-86; initHeadList :153-224; extendAll/extend/addHaploPair :226-261; HaploPair
// constructors / add, HaploPair.cpp:14-89).
//
// One warp owns one individual (one-warp CTAs, persistent grid, work queue).  Per locus the warp
// streams over the candidate slots (predecessor state x ordered allele pair) IN THE REFERENCE'S ORDER,
// 32 at a time:
//   - one 128-bit gather per haplotype from the successor table {successor id, its transition
//     probability}; the gathers of the next 32 slots are issued before the current 32 are consumed;
//   - lanes with the same canonical pair (id_a <= id_b) are grouped with match.any; each group's leader
//     probes / inserts the pair in an open-addressing hash in shared memory; new pairs get their state
//     index from a ballot prefix, so states are created exactly in the reference's order;
//   - every candidate then folds itself into its state, group members in lane (= ordinal) order:
//     alpha += alpha_pred * tp with no FMA, and the predecessor's sorted K-best list is merged into the
//     state's list (het-locus rule applied), again in the reference's order.
// No pass over the candidates is repeated and no block-level barrier is needed.  Live states, their
// K-best lists and the hash live in shared memory up to CS / CL entries; anything beyond goes to the
// warp's global spill scratch through the same accessors.
#pragma once

#include "common.cuh"
#include "forward.cuh"

namespace hmc {

#ifdef HMC_PROFILE_PASSES
#define WMARK(slot) do { const long long t_ = clock64(); wcyc[slot] += t_ - wmark; wmark = t_; } while (0)
#else
#define WMARK(slot) do { } while (0)
#endif

template <int CS, int CL, int KMAX>
struct WarpSmem {
    static constexpr int HS = 2 * CS;
    u64 hkeys[HS];
    double alpha[2][CS];
    double tp[CS];
    double like[2][CL * KMAX];
    u64 meta[2][CL];
    int hidx[HS];
    int ida[2][CS], idb[2][CS];
    u32 bp[CL * KMAX];
    unsigned char ab[CS];
    long long chunk_cur, chunk_end;
};

template <int CS, int CL, int KMAX>
__global__ void __launch_bounds__(32, 8) forward_warp_kernel(FwdParams p)
{
    typedef WarpSmem<CS, CL, KMAX> Smem;
    constexpr int HS = Smem::HS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    constexpr u32 FULL = 0xffffffffu;

    const int lane = threadIdx.x;
    const u32 lt_mask = (1u << lane) - 1u;
    const DevPanel &pn = p.panel;
    const DevModel &md = p.model;
    const int L = pn.L, A = md.Amax, K = p.K, head_len = md.head_len;
    const Scratch sc = p.scratch[blockIdx.x];
    double *const sc_like0 = sc.like[0], *const sc_like1 = sc.like[1];
    u64 *const sc_meta0 = sc.meta[0], *const sc_meta1 = sc.meta[1];
    // spill-step output arrays (state index >= CS), reusing the block kernel's scratch
    int *const o_ida = sc.cand, *const o_idb = sc.arr;
    double *const o_alpha = sc.tpv, *const o_tp = sc.otp;
    u32 *const o_bp = sc.obp;
    unsigned char *const o_ab = (unsigned char *) sc.slot_of;
    u64 *const arena = p.arena.base;
    const int4 *const succ16 = reinterpret_cast<const int4 *>(md.succ16);

    if (lane == 0) { sm.chunk_cur = 0; sm.chunk_end = 0; }

    // chunked bump allocation by lane 0, result broadcast; -1 when the arena is exhausted
    auto alloc = [&](long long units) -> long long {
        long long off = 0;
        if (lane == 0) {
            constexpr long long CHUNK = 1 << 14;
            units = (units + 1) & ~1LL;
            if (sm.chunk_cur + units > sm.chunk_end) {
                const long long grab = units > CHUNK ? units : CHUNK;
                const unsigned long long base = atomicAdd(p.arena.cursor, (unsigned long long) grab);
                if ((long long) (base + grab) > p.arena.capacity) { atomicExch(p.arena.exhausted, 1); off = -1; }
                else { sm.chunk_cur = (long long) base; sm.chunk_end = (long long) base + grab; }
            }
            if (off == 0) { off = sm.chunk_cur; sm.chunk_cur += units; }
        }
        return __shfl_sync(FULL, off, 0);
    };

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.queue, 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= p.n_wave) break;
        if (*(volatile int *) p.arena.exhausted) { if (lane == 0) p.status[w] = ST_NOMEM; continue; }
        const long long t_begin = clock64();
        const int ind = p.wave[w];
        const u32 *gp = pn.packed + (size_t) ind * pn.words;
        LocusRec *rec = p.loci + (size_t) w * (L + 1);
        for (int l = lane; l <= L; l += 32) { rec[l].count = 0; rec[l].fwd_slots = 0; rec[l].off = 0; rec[l].fwd_off = 0; }
        const int gbits = pn.bits;
        const u32 gmask = (1u << gbits) - 1u;
        auto geno_word = [&](int l) -> u32 { const int bit = l * 2 * gbits; return gp[bit >> 5] >> (bit & 31); };

        // ---------------- head states (initHeadList), serial on lane 0 ----------------
        int S_in = 0, status = ST_DONE;
        {
            int nh = 0, st = ST_DONE;
            if (lane == 0) {
                int bad = 0;
                int hg0[8], hg1[8];
                for (int j = 0; j < head_len; ++j) { const u32 wd = geno_word(j); hg0[j] = (int) (wd & gmask) - 1; hg1[j] = (int) ((wd >> gbits) & gmask) - 1; }
                for (int hi = 0; hi < md.n_heads && !bad; ++hi) {
                    const int head = md.heads[hi];
                    const signed char *ha = md.al + md.al_off[head];
                    bool match = true;
                    for (int j = 0; j < head_len; ++j) {
                        const int a = ha[j];
                        if (!(hg0[j] < 0 || hg1[j] < 0 || hg0[j] == a || hg1[j] == a)) { match = false; break; }
                    }
                    if (!match) continue;
                    int lo[8], hi_[8], curv[8];
                    for (int j = 0; j < head_len; ++j) {
                        const int a = ha[j];
                        const bool both = hg0[j] < 0 && hg1[j] < 0, any = hg0[j] < 0 || hg1[j] < 0;
                        if (both || (any && (hg0[j] == a || hg1[j] == a))) { lo[j] = 0; hi_[j] = (int) (pn.linfo[j] & 0xFF) - 1; }
                        else if (!any && hg0[j] != hg1[j]) { lo[j] = hi_[j] = (a == hg0[j]) ? hg1[j] : hg0[j]; }
                        else { const int v = hg0[j] < 0 ? 0 : hg0[j]; lo[j] = hi_[j] = v; }   // missing -> wildcard lookup -> first child
                        curv[j] = lo[j];
                    }
                    for (;;) {
                        bool ok = true;
                        int code = 0;
                        for (int j = 0; j < head_len; ++j) {
                            if (lo[j] != hi_[j] && !(((pn.linfo[j] >> 8) >> curv[j]) & 1u)) ok = false;
                            code += curv[j] * md.head_stride[j];
                        }
                        if (ok) {
                            const int hp = md.head_lut[code];
                            if (hp < 0) { bad = 1; break; }
                            if (hp >= head) {
                                if (nh < CS) { sm.ida[0][nh] = head; sm.idb[0][nh] = hp; }
                                else if (nh < sc.scap) { o_ida[nh] = head; o_idb[nh] = hp; }
                                ++nh;
                            }
                        }
                        int j = head_len - 1;
                        while (j >= 0) { if (++curv[j] <= hi_[j]) break; curv[j] = lo[j]; --j; }
                        if (j < 0) break;
                    }
                }
                if (bad) st = ST_NO_HEAD;
                else if (nh > sc.scap) st = ST_TOO_BIG;
            }
            S_in = __shfl_sync(FULL, nh, 0);
            status = __shfl_sync(FULL, st, 0);
        }
        if (status != ST_DONE) { if (lane == 0) p.status[w] = status; continue; }
        long long cur_off = 0;
        if (S_in > 0) {
            cur_off = alloc(block_layout(S_in, K, p.estep).total);
            if (cur_off < 0) { if (lane == 0) p.status[w] = ST_NOMEM; continue; }
            const BlockLayout bl = block_layout(S_in, K, p.estep);
            int *t_ida = (int *) (arena + cur_off + bl.ida), *t_idb = (int *) (arena + cur_off + bl.idb);
            double *t_alpha = (double *) (arena + cur_off + bl.alpha);
            unsigned char *t_ab = (unsigned char *) (arena + cur_off + bl.ab);
            u32 *t_bp = (u32 *) (arena + cur_off + bl.bp);
            double *t_tp = (double *) (arena + cur_off + bl.tp);
            __syncwarp();
            for (int s = lane; s < S_in; s += 32) {
                const int a = s < CS ? sm.ida[0][s] : o_ida[s], b = s < CS ? sm.idb[0][s] : o_idb[s];
                const double f = __dmul_rn(md.freq[a], md.freq[b]);
                const double al = (a == b) ? f : __dmul_rn(f, 2.0);
                if (s < CS) sm.alpha[0][s] = al;
                t_ida[s] = a; t_idb[s] = b; t_alpha[s] = al;
                t_ab[s] = (unsigned char) (md.last[a] | (md.last[b] << 4));
                if (p.estep) t_tp[s] = f;
                else t_bp[(size_t) s * K] = BP_HEAD | (a == b ? BP_HOM : 0u);
                const u64 mt = 1ull | ((a == b ? 1ull : 0ull) << 8);
                if (s < CL) { sm.like[0][s * KMAX] = f; sm.meta[0][s] = mt; }
                else { sc_like0[(size_t) s * K] = f; sc_meta0[s] = mt; }
            }
            if (lane == 0) { rec[head_len].off = cur_off; rec[head_len].count = S_in; }
        }
#ifdef HMC_PROFILE_PASSES
        long long wcyc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long wmark = clock64();
#endif
        int par = 0, cur_exp = 0;
        long long c_steps = 0, c_states = S_in, c_max = S_in, c_arr = 0, c_spill = 0;
        if (p.locus_exp && lane == 0) for (int l = 0; l <= head_len; ++l) p.locus_exp[(size_t) w * (L + 1) + l] = 0;

        u32 nx_word = 0, nx_info = 0;
        if (head_len < L) { nx_word = geno_word(head_len); nx_info = pn.linfo[head_len]; }

        // ---------------- locus steps ----------------
        for (int i = head_len; i < L && S_in > 0; ++i) {
            __syncwarp();
            WMARK(7);
            const int g0 = (int) (nx_word & gmask) - 1, g1 = (int) ((nx_word >> gbits) & gmask) - 1;
            const int nA = (int) (nx_info & 0xFF);
            const u32 am = nx_info >> 8;
            if (i + 1 < L) { nx_word = geno_word(i + 1); nx_info = pn.linfo[i + 1]; }

            // Candidate slots in the reference's extendAll order (HaploBuilder.cpp:48-79): groups of allele pairs
            // (x, y); inside a group, per predecessor state, (x,y) then (y,x) when x != y.
            const int mode = (g0 >= 0 && g1 >= 0) ? 0 : (g0 < 0 && g1 < 0) ? 3 : (g0 < 0 ? 1 : 2);
            const int nlive = __popc(am & ((1u << nA) - 1u));
            int C;
            if (mode == 0) C = S_in * (g0 != g1 ? 2 : 1);
            else if (mode == 3) C = S_in * nlive * nlive;                   // j<k pairs twice, j==k once
            else { const int known = (mode == 1) ? g1 : g0; C = S_in * (2 * nlive - (((am >> known) & 1u) ? 1 : 0)); }
            // slot -> (pred state, x, y): closed forms per mode
            auto decode = [&](int c, int &s, int &x, int &y) {
                if (mode == 0) {
                    if (g0 != g1) { s = c >> 1; const int d = c & 1; x = d ? g1 : g0; y = d ? g0 : g1; }
                    else { s = c; x = y = g0; }
                    return;
                }
                // walk the groups (few: <= A(A+1)/2); each group spans S_in * ndir slots
                int rem = c;
                if (mode == 3) {
                    for (int j = 0; j < nA; ++j) { if (!((am >> j) & 1u)) continue;
                        for (int k = j; k < nA; ++k) { if (!((am >> k) & 1u)) continue;
                            const int span = S_in * (j != k ? 2 : 1);
                            if (rem < span) { if (j != k) { s = rem >> 1; const int d = rem & 1; x = d ? k : j; y = d ? j : k; } else { s = rem; x = y = j; } return; }
                            rem -= span;
                        } }
                }
                else {
                    const int known = (mode == 1) ? g1 : g0;
                    for (int j = 0; j < nA; ++j) { if (!((am >> j) & 1u)) continue;
                        const int span = S_in * (j != known ? 2 : 1);
                        if (rem < span) { if (j != known) { s = rem >> 1; const int d = rem & 1; x = d ? known : j; y = d ? j : known; } else { s = rem; x = y = j; } return; }
                        rem -= span;
                    }
                }
                s = 0; x = y = 0;
            };

            int hsize = 64;
            while (hsize < 2 * C) hsize <<= 1;
            const bool hfast = hsize <= HS;
            if (!hfast && (hsize > sc.hcap || C > sc.ccap)) { status = ST_TOO_BIG; break; }
            if (!hfast) ++c_spill;
            u64 *const hkeys = hfast ? sm.hkeys : sc.keys;
            int *const hidx = hfast ? sm.hidx : (int *) sc.ord;
            const u32 hmask = (u32) hsize - 1u;
            for (int h = lane; h < hsize; h += 32) hkeys[h] = KEY_EMPTY;

            // input views
            const BlockLayout bli = block_layout(S_in, K, p.estep);
            const int *gi_ida = (const int *) (arena + cur_off + bli.ida), *gi_idb = (const int *) (arena + cur_off + bli.idb);
            const double *gi_alpha = (const double *) (arena + cur_off + bli.alpha);
            const int *ci_ida = sm.ida[par], *ci_idb = sm.idb[par];
            const double *ci_alpha = sm.alpha[par];
            const double *ci_like = sm.like[par];
            const u64 *ci_meta = sm.meta[par];
            const double *like_in = par ? sc_like1 : sc_like0;
            const u64 *meta_in = par ? sc_meta1 : sc_meta0;
            int *co_ida = sm.ida[par ^ 1], *co_idb = sm.idb[par ^ 1];
            double *co_alpha = sm.alpha[par ^ 1];
            double *co_like = sm.like[par ^ 1];
            u64 *co_meta = sm.meta[par ^ 1];
            double *like_out = par ? sc_like0 : sc_like1;
            u64 *meta_out = par ? sc_meta0 : sc_meta1;

            long long fwd_off = 0;
            if (p.estep) {
                fwd_off = alloc(((long long) C + 1) / 2);
                if (fwd_off < 0) { status = ST_NOMEM; break; }
                if (lane == 0) { rec[i].fwd_off = fwd_off; rec[i].fwd_slots = C; }
            }
            int *fwd = (int *) (arena + fwd_off);
            __syncwarp();

            // prefetch registers for the next 32 slots
            int n_s = 0, n_x = 0, n_y = 0;
            bool n_live = false;
            double n_al = 0.0;
            int4 n_ea = make_int4(-1, 0, 0, 0), n_eb = make_int4(-1, 0, 0, 0);
            auto prefetch = [&](int c) {
                n_live = false;
                if (c < C) {
                    decode(c, n_s, n_x, n_y);
                    const int pa = n_s < CS ? ci_ida[n_s] : gi_ida[n_s], pb = n_s < CS ? ci_idb[n_s] : gi_idb[n_s];
                    n_al = n_s < CS ? ci_alpha[n_s] : gi_alpha[n_s];
                    if (n_al > 0.0) {       // HaploBuilder.cpp:237
                        n_live = true;
                        n_ea = __ldg(succ16 + (size_t) pa * A + n_x);
                        n_eb = __ldg(succ16 + (size_t) pb * A + n_y);
                    }
                }
            };
            WMARK(0);
            prefetch(lane);

            int S_out = 0, n_arr = 0;
            bool too_big = false;
            for (int base = 0; base < C; base += 32) {
                const int c = base + lane;
                const int s = n_s, x = n_x, y = n_y;
                const double ain = n_al;
                int hpa = n_ea.x, hpb = n_eb.x;
                const bool valid = n_live && hpa >= 0 && hpb >= 0;
                const double tv = __dmul_rn(__hiloint2double(n_ea.w, n_ea.z), __hiloint2double(n_eb.w, n_eb.z));
                if (__any_sync(FULL, valid) || true) WMARK(1);
                prefetch(c + 32);
                WMARK(2);

                u32 rev = 0;
                if (hpa > hpb) { const int t = hpa; hpa = hpb; hpb = t; rev = 1; }
                const u64 key = valid ? (((u64) (u32) hpa << 32) | (u32) hpb) : (0xFFFFFFFF00000000ull | (u32) lane);    // invalid lanes: unique dummy keys
                const u32 grp = __match_any_sync(FULL, key);
                const int leader = __ffs(grp) - 1;
                const int rank = __popc(grp & lt_mask);
                // leaders probe / insert
                int idx = -1;
                bool isnew = false;
                u32 h = 0;
                if (valid && lane == leader) {
                    h = mix64(key) & hmask;
                    for (;;) {
                        const u64 prev = atomicCAS(hkeys + h, KEY_EMPTY, key);
                        if (prev == KEY_EMPTY) { isnew = true; break; }
                        if (prev == key) break;
                        h = (h + 1) & hmask;
                    }
                }
                const u32 newmask = __ballot_sync(FULL, isnew);
                if (isnew) { idx = S_out + __popc(newmask & lt_mask); hidx[h] = idx; }
                __syncwarp();
                if (valid && lane == leader && !isnew) idx = hidx[h];
                S_out += __popc(newmask);
                idx = __shfl_sync(FULL, idx, leader);
                const bool grp_new = __shfl_sync(FULL, (int) isnew, leader) != 0;
                const u32 vmask = __ballot_sync(FULL, valid);
                n_arr += __popc(vmask);
                if (S_out > sc.scap || S_out >= MAX_STATES_PER_LOCUS) { too_big = true; break; }
                if (p.estep && c < C) fwd[c] = valid ? (int) ((u32) idx | (rev << 31)) : -1;

                WMARK(3);
                // fold the candidates into their states, group members in ordinal order
                const bool het = (x != y);
                int maxrank = valid ? rank : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) maxrank = max(maxrank, __shfl_xor_sync(FULL, maxrank, o));
                for (int r = 0; r <= maxrank; ++r) {
                    if (valid && rank == r) {
                        const bool first = grp_new && r == 0;
                        // predecessor's list (sorted descending)
                        const u64 mt = s < CL ? ci_meta[s] : meta_in[s];
                        const int nl = (int) (mt & 0xFF);
                        u32 hom = (u32) (mt >> 8);
                        double xl[KMAX];
                        if (s < CL) {
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) xl[k] = ci_like[s * KMAX + k];
                        }
                        else {
                            const double *lp = like_in + (size_t) s * K;
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) xl[k] = (k < nl) ? lp[k] : 0.0;
                        }
                        const double term = __dmul_rn(ain, tv);
                        // het-locus rule (HaploPair.cpp:53-60,77-84)
                        const u32 zero = (het && rev) ? hom : 0u;
                        if (het) hom = 0;
                        const u32 bpbase = (u32) s | (rev ? BP_REV : 0u);
                        double bl[KMAX];
                        u32 bb[KMAX];
                        int nl_cur = 0;
                        if (first) {
                            if (idx < CS) { co_ida[idx] = hpa; co_idb[idx] = hpb; co_alpha[idx] = term; sm.tp[idx] = tv; sm.ab[idx] = (unsigned char) ((rev ? y : x) | ((rev ? x : y) << 4)); }
                            else { o_ida[idx] = hpa; o_idb[idx] = hpb; o_alpha[idx] = term; o_tp[idx] = tv; o_ab[idx] = (unsigned char) ((rev ? y : x) | ((rev ? x : y) << 4)); }
                        }
                        else {
                            if (idx < CS) co_alpha[idx] = __dadd_rn(co_alpha[idx], term);
                            else o_alpha[idx] = __dadd_rn(o_alpha[idx], term);
                            // current list of the state
                            const u64 mc = idx < CL ? co_meta[idx] : meta_out[idx];
                            nl_cur = (int) (mc & 0xFF);
                            if (idx < CL) {
#pragma unroll
                                for (int k = 0; k < KMAX; ++k) { bl[k] = (k < nl_cur) ? co_like[idx * KMAX + k] : -1.0; bb[k] = sm.bp[idx * KMAX + k]; }
                            }
                            else {
#pragma unroll
                                for (int k = 0; k < KMAX; ++k) { bl[k] = (k < nl_cur) ? like_out[(size_t) idx * K + k] : -1.0; bb[k] = (k < nl_cur) ? o_bp[(size_t) idx * K + k] : 0u; }
                            }
                        }
                        if (first && zero == 0) {
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) {
                                bl[k] = (k < nl) ? __dmul_rn(xl[k], tv) : -1.0;
                                bb[k] = bpbase | ((u32) k << BP_PRED_BITS) | (((hom >> k) & 1u) ? BP_HOM : 0u);
                            }
                        }
                        else {
                            if (first) {
#pragma unroll
                                for (int k = 0; k < KMAX; ++k) { bl[k] = -1.0; bb[k] = 0; }
                            }
                            int k = 0;
                            while (k < nl) {
                                double lv = __dmul_rn(xl[0], tv);
                                if ((zero >> k) & 1u) lv = 0.0;
                                if (lv > bl[KMAX - 1]) topk_insert<KMAX>(bl, bb, lv, bpbase | ((u32) k << BP_PRED_BITS) | (((hom >> k) & 1u) ? BP_HOM : 0u));
                                else if (zero == 0) break;
#pragma unroll
                                for (int q = 0; q + 1 < KMAX; ++q) xl[q] = xl[q + 1];
                                ++k;
                            }
                        }
                        int nl_new = nl_cur + nl;
                        if (nl_new > K) nl_new = K;
                        u32 hom_out = 0;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < nl_new && (bb[k] & BP_HOM)) hom_out |= 1u << k;
                        const u64 mo = (u64) nl_new | ((u64) hom_out << 8);
                        if (idx < CL) {
                            co_meta[idx] = mo;
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) { co_like[idx * KMAX + k] = bl[k]; sm.bp[idx * KMAX + k] = bb[k]; }
                        }
                        else {
                            meta_out[idx] = mo;
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) if (k < nl_new) { like_out[(size_t) idx * K + k] = bl[k]; o_bp[(size_t) idx * K + k] = bb[k]; }
                        }
                    }
                    __syncwarp();
                }
                WMARK(4);
            }
            if (too_big) { status = ST_TOO_BIG; break; }
            if (S_out == 0) { S_in = 0; break; }       // every extension died: HaploBuilder.cpp:80-82
            c_arr += n_arr;

            const long long new_off = alloc(block_layout(S_out, K, p.estep).total);
            if (new_off < 0) { status = ST_NOMEM; break; }
            __syncwarp();

            // exponent-only rescale decision (mantissas stay those of the unscaled product)
            double factor = 1.0;
            int new_exp = cur_exp;
            if (p.scaled) {
                double amax = 0.0;
                for (int idx = lane; idx < S_out; idx += 32) amax = fmax(amax, idx < CS ? co_alpha[idx] : o_alpha[idx]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(FULL, amax, o));
                if (amax > 0.0 && amax < 0x1p-512) { factor = 0x1p+512; new_exp += 512; }
            }

            // ---- coalesced trellis write of the new index ----
            const BlockLayout blo = block_layout(S_out, K, p.estep);
            int *t_ida = (int *) (arena + new_off + blo.ida), *t_idb = (int *) (arena + new_off + blo.idb);
            double *t_alpha = (double *) (arena + new_off + blo.alpha);
            unsigned char *t_ab = (unsigned char *) (arena + new_off + blo.ab);
            u32 *t_bp = (u32 *) (arena + new_off + blo.bp);
            double *t_tp = (double *) (arena + new_off + blo.tp);
            for (int idx = lane; idx < S_out; idx += 32) {
                const bool c1 = idx < CS;
                double a = c1 ? co_alpha[idx] : o_alpha[idx];
                if (factor != 1.0) { a *= factor; if (c1) co_alpha[idx] = a; else o_alpha[idx] = a; }
                t_ida[idx] = c1 ? co_ida[idx] : o_ida[idx];
                t_idb[idx] = c1 ? co_idb[idx] : o_idb[idx];
                t_alpha[idx] = a;
                t_ab[idx] = c1 ? sm.ab[idx] : o_ab[idx];
                if (p.estep) t_tp[idx] = c1 ? sm.tp[idx] : o_tp[idx];
            }
            if (!p.estep || factor != 1.0) {
                const int tot = S_out * K;
                for (int q = lane; q < tot; q += 32) {
                    const int idx = q / K, k = q - idx * K;
                    const int nl = (int) ((idx < CL ? co_meta[idx] : meta_out[idx]) & 0xFF);
                    if (k < nl) {
                        if (!p.estep) t_bp[q] = idx < CL ? sm.bp[idx * KMAX + k] : o_bp[q];
                        if (factor != 1.0) { if (idx < CL) co_like[idx * KMAX + k] *= factor; else like_out[q] *= factor; }
                    }
                }
            }
            // spilled states become next step's input through the trellis block (ida/idb/alpha) and sc.like/meta
            if (lane == 0) {
                rec[i + 1].off = new_off; rec[i + 1].count = S_out;
                if (p.locus_exp) p.locus_exp[(size_t) w * (L + 1) + i + 1] = new_exp;
            }
            cur_exp = new_exp;
            ++c_steps; c_states += S_out; if (S_out > c_max) c_max = S_out;
            S_in = S_out; par ^= 1; cur_off = new_off;
            WMARK(5);
        }
        __syncwarp();
        if (status == ST_DONE && S_in > 0 && !p.estep) {
            // final link lists for the traceback kernel
            const long long fo = alloc((long long) S_in * K + S_in);
            if (fo < 0) status = ST_NOMEM;
            else {
                double *fl = (double *) (arena + fo);
                u64 *fm = arena + fo + (long long) S_in * K;
                for (int s = lane; s < S_in; s += 32) {
                    const u64 mt = s < CL ? sm.meta[par][s] : (par ? sc_meta1 : sc_meta0)[s];
                    fm[s] = mt;
                    const int nl = (int) (mt & 0xFF);
                    for (int k = 0; k < nl; ++k) fl[(size_t) s * K + k] = s < CL ? sm.like[par][s * KMAX + k] : (par ? sc_like1 : sc_like0)[(size_t) s * K + k];
                }
                if (lane == 0) { rec[0].off = fo; rec[0].count = S_in; }
            }
        }
        if (lane == 0) {
            p.status[w] = status;
            p.final_exp[w] = cur_exp;
            if (p.ind_cycles) p.ind_cycles[w] = clock64() - t_begin;
            if (status == ST_DONE) {
                atomicAdd((unsigned long long *) p.counters + 0, (unsigned long long) c_steps);
                atomicAdd((unsigned long long *) p.counters + 1, (unsigned long long) c_states);
                atomicMax((long long *) p.counters + 2, c_max);
                atomicAdd((unsigned long long *) p.counters + 3, (unsigned long long) c_arr);
                atomicAdd((unsigned long long *) p.counters + 5, (unsigned long long) c_spill);
#ifdef HMC_PROFILE_PASSES
                for (int q = 0; q < 8; ++q) atomicAdd((unsigned long long *) p.counters + 8 + q, (unsigned long long) wcyc[q]);
#endif
            }
        }
    }
}

}  // namespace hmc
