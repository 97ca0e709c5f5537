// K1 forward_resolve: the per-individual forward pass of HaploBuilder::resolve
// (HaploBuilder.cpp:35-86 with initHeadList :153-224, extendAll/extend/addHaploPair :226-261 and the
// HaploPair constructors / add, HaploPair.cpp:14-89).
//
// One CTA (NT threads, NT = 32 is one warp) owns one individual at a time and pulls individuals from
// a work queue (persistent grid).  Per locus step the lanes are spread over the candidate slots
// (predecessor state x ordered allele pair); the step is
//   A   one 128-bit gather per haplotype from the successor table {successor id, its transition
//       probability}, canonicalise (id_a <= id_b), insert into an open-addressing hash on (id_a,id_b);
//       the slot remembers the smallest candidate ordinal that hit it
//   B+C walk the candidates in ordinal order; the one whose ordinal equals the slot's minimum creates
//       the state -> new states get exactly the reference's creation order; the same block scan
//       yields the per-state arrival-segment offsets
//   D   scatter arrivals into the per-state segments
//   E   one lane per new state: arrivals in ordinal order -> alpha sum in the reference's order
//       (no FMA), K-best merge of the predecessors' sorted link lists, het-locus rule, trellis write.
// While a step fits (C <= CC candidate slots) the hash, candidate and segment arrays are in shared
// memory and addressed as such (the step body is instantiated twice: shared / global spill scratch);
// live states (ids, alpha) and, for small sets, their K-best lists are cached in shared memory too.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace hmc {

#ifdef HMC_PROFILE_PASSES
#define PASS_MARK(slot) do { const long long t_ = clock64(); if (tid == 0) { if (pass_fast) pass_cyc[slot] += t_ - t_mark; else pass_cyc[(slot) + 8] += t_ - t_mark; } t_mark = t_; } while (0)
#else
#define PASS_MARK(slot) do { } while (0)
#endif

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

template <int NT, int CC, int CL, int KMAX>
struct FwdSmem {
    static constexpr int HC = 2 * CC <= 256 ? 256 : 2 * CC <= 512 ? 512 : 2 * CC <= 1024 ? 1024 : 2 * CC <= 2048 ? 2048 : 4096;   // power of two >= 2 CC
    u64 keys[HC];
    double tpv[HC];
    double alpha[2][CC];
    double like[2][CL * KMAX];
    u64 meta[2][CL];
    long long ll[4];                // -, alloc result, chunk cursor, chunk end
    long long tmp[8];               // scan scratch
    u32 ord[HC], cnt[HC];
    alignas(16) int cand[CC];
    int arr[CC], aoff[CC + 1], slot_of[CC];
    int ida[2][CC], idb[2][CC];
    int gx[MAX_GROUPS], gy[MAX_GROUPS], gbase[MAX_GROUPS + 1];
    int ints[4];
};

template <int NT, int CC, int CL, int KMAX, int MINB>
__global__ void __launch_bounds__(NT, MINB) forward_kernel(FwdParams p)
{
    typedef FwdSmem<NT, CC, CL, KMAX> Smem;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);

    const int tid = threadIdx.x;
    const DevPanel &pn = p.panel;
    const DevModel &md = p.model;
    const int L = pn.L, A = md.Amax, K = p.K, head_len = md.head_len;
    const Scratch sc = p.scratch[blockIdx.x];
    double *const sc_like0 = sc.like[0], *const sc_like1 = sc.like[1];
    u64 *const sc_meta0 = sc.meta[0], *const sc_meta1 = sc.meta[1];
    u64 *const arena = p.arena.base;
    const int4 *const succ16 = reinterpret_cast<const int4 *>(md.succ16);

    if (tid == 0) { sm.ll[2] = 0; sm.ll[3] = 0; }

    // chunked bump allocation (thread 0 only); returns -1 when the arena is exhausted
    auto alloc = [&](long long units) -> long long {
        constexpr long long CHUNK = 1 << 15;
        units = (units + 1) & ~1LL;
        if (sm.ll[2] + units > sm.ll[3]) {
            const long long grab = units > CHUNK ? units : CHUNK;
            const unsigned long long base = atomicAdd(p.arena.cursor, (unsigned long long) grab);
            if ((long long) (base + grab) > p.arena.capacity) { atomicExch(p.arena.exhausted, 1); return -1; }
            sm.ll[2] = (long long) base; sm.ll[3] = (long long) base + grab;
        }
        const long long off = sm.ll[2];
        sm.ll[2] += units;
        return off;
    };

    for (;;) {
        block_sync<NT>();
        if (tid == 0) sm.ints[0] = atomicAdd(p.queue, 1);
        block_sync<NT>();
        const int w = sm.ints[0];
        if (w >= p.n_wave) break;
        if (*(volatile int *) p.arena.exhausted) { if (tid == 0) p.status[w] = ST_NOMEM; continue; }
        const long long t_begin = clock64();
        const int ind = p.wave[w];
        const u32 *gp = pn.packed + (size_t) ind * pn.words;
        LocusRec *rec = p.loci + (size_t) w * (L + 1);
        for (int l = tid; l <= L; l += NT) { rec[l].count = 0; rec[l].fwd_slots = 0; rec[l].off = 0; rec[l].fwd_off = 0; }
        const int gbits = pn.bits;
        const u32 gmask = (1u << gbits) - 1u;
        auto geno_word = [&](int l) -> u32 { const int bit = l * 2 * gbits; return gp[bit >> 5] >> (bit & 31); };

        // ---------------- head states (initHeadList), serial: at most A^head_len pairs ----------------
        int S_in = 0, par = 0, status = ST_DONE, cur_exp = 0;
        long long cur_off = 0;
        if (tid == 0) {
            int nh = 0, bad = 0;
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
                // odometer over the complementary allele strings, locus 0 most significant
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
                            if (nh < CC) { sm.ida[0][nh] = head; sm.idb[0][nh] = hp; }
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
            sm.ints[1] = nh; sm.ints[2] = st; sm.ll[1] = off;
        }
        block_sync<NT>();
        S_in = sm.ints[1]; status = sm.ints[2]; cur_off = sm.ll[1];
        if (status != ST_DONE) { if (tid == 0) p.status[w] = status; continue; }
        {
            const BlockLayout bl = block_layout(S_in, K, p.estep);
            int *t_ida = (int *) (arena + cur_off + bl.ida), *t_idb = (int *) (arena + cur_off + bl.idb);
            double *t_alpha = (double *) (arena + cur_off + bl.alpha);
            unsigned char *t_ab = (unsigned char *) (arena + cur_off + bl.ab);
            u32 *t_bp = (u32 *) (arena + cur_off + bl.bp);
            double *t_tp = (double *) (arena + cur_off + bl.tp);
            for (int s = tid; s < S_in; s += NT) {
                const int a = s < CC ? sm.ida[0][s] : sc.cand[2 * s], b = s < CC ? sm.idb[0][s] : sc.cand[2 * s + 1];
                const double f = __dmul_rn(md.freq[a], md.freq[b]);
                const double al = (a == b) ? f : __dmul_rn(f, 2.0);
                if (s < CC) sm.alpha[0][s] = al;
                t_ida[s] = a; t_idb[s] = b; t_alpha[s] = al;
                t_ab[s] = (unsigned char) (md.last[a] | (md.last[b] << 4));
                if (p.estep) t_tp[s] = f;
                else t_bp[(size_t) s * K] = BP_HEAD | (a == b ? BP_HOM : 0u);
                const u64 mt = 1ull | ((a == b ? 1ull : 0ull) << 8);
                if (S_in <= CL) { sm.like[0][s * KMAX] = f; sm.meta[0][s] = mt; }
                else { sc_like0[(size_t) s * K] = f; sc_meta0[s] = mt; }
            }
            if (tid == 0) { rec[head_len].off = cur_off; rec[head_len].count = S_in; }
        }
#ifdef HMC_PROFILE_PASSES
        long long pass_cyc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool pass_fast = true;
        long long t_mark = clock64();
#endif
        par = 0;
        long long c_steps = 0, c_states = S_in, c_max = S_in, c_arr = 0, c_spill = 0;
        if (p.locus_exp && tid == 0) for (int l = 0; l <= head_len; ++l) p.locus_exp[(size_t) w * (L + 1) + l] = 0;

        // genotype word and locus info of the next step are fetched one step ahead
        u32 nx_word = 0, nx_info = 0;
        if (head_len < L) { nx_word = geno_word(head_len); nx_info = pn.linfo[head_len]; }

        // ---------------- locus steps ----------------
        for (int i = head_len; i < L && S_in > 0; ++i) {
            PASS_MARK(7);
            const int g0 = (int) (nx_word & gmask) - 1, g1 = (int) ((nx_word >> gbits) & gmask) - 1;
            const int nA = (int) (nx_info & 0xFF);
            const u32 am = nx_info >> 8;
            if (i + 1 < L) { nx_word = geno_word(i + 1); nx_info = pn.linfo[i + 1]; }
            // allele-pair groups in the reference's extendAll order (HaploBuilder.cpp:48-79)
            int ng = 0, C = 0;
            auto add_group = [&](int x, int y) {
                if (tid == 0) { sm.gx[ng] = x; sm.gy[ng] = y; sm.gbase[ng] = C; }
                C += S_in * (x != y ? 2 : 1);
                ++ng;
            };
            if (g0 < 0 && g1 < 0) { for (int j = 0; j < nA; ++j) if ((am >> j) & 1u) for (int k = j; k < nA; ++k) if ((am >> k) & 1u) add_group(j, k); }
            else if (g0 < 0) { for (int j = 0; j < nA; ++j) if ((am >> j) & 1u) add_group(j, g1); }
            else if (g1 < 0) { for (int j = 0; j < nA; ++j) if ((am >> j) & 1u) add_group(j, g0); }
            else add_group(g0, g1);
            if (tid == 0) sm.gbase[ng] = C;
            int hsize = 64;
            while (hsize * 4 < 5 * C) hsize <<= 1;     // load factor <= 0.8 even if every candidate were a new state
            const bool fast = (C <= CC);
            if (!fast && (C > sc.ccap || hsize > sc.hcap)) { status = ST_TOO_BIG; break; }
            if (!fast) ++c_spill;
#ifdef HMC_PROFILE_PASSES
            pass_fast = fast;
            if (tid == 0) { if (fast) pass_cyc[6] += C; else pass_cyc[14] += C; }
#endif

            int S_out = 0, new_exp = cur_exp;
            long long new_off = 0;

            // The step body; FAST selects shared-memory work arrays (compile-time, so the accesses are LDS/STS/ATOMS).
            auto body = [&](auto fast_tag) -> int {
                constexpr bool FAST = decltype(fast_tag)::value;
                u64 *const keys = FAST ? sm.keys : sc.keys;
                double *const tpvs = FAST ? sm.tpv : sc.tpv;
                u32 *const ord = FAST ? sm.ord : sc.ord;
                u32 *const cnt = FAST ? sm.cnt : sc.cnt;
                int *const cand = FAST ? sm.cand : sc.cand;
                int *const arr = FAST ? sm.arr : sc.arr;
                int *const aoff = FAST ? sm.aoff : sc.aoff;
                int *const slot_of = FAST ? sm.slot_of : sc.slot_of;
                const u32 hmask = (u32) hsize - 1u;
                for (int h = tid; h < hsize; h += NT) { keys[h] = KEY_EMPTY; ord[h] = 0xFFFFFFFFu; cnt[h] = 0; }

                // input state views
                const BlockLayout bli = block_layout(S_in, K, p.estep);
                const int *gi_ida = (const int *) (arena + cur_off + bli.ida), *gi_idb = (const int *) (arena + cur_off + bli.idb);
                const double *gi_alpha = (const double *) (arena + cur_off + bli.alpha);
                const bool in_cache = FAST || S_in <= CC;       // FAST implies S_in <= C <= CC
                const int *ci_ida = sm.ida[par], *ci_idb = sm.idb[par];
                const double *ci_alpha = sm.alpha[par];
                const bool lk_in_cache = S_in <= CL;
                const double *like_in = par ? sc_like1 : sc_like0;
                const u64 *meta_in = par ? sc_meta1 : sc_meta0;
                double *like_out = par ? sc_like0 : sc_like1;
                u64 *meta_out = par ? sc_meta0 : sc_meta1;

                // forward-link block of index i (E-step mode)
                long long fwd_off = 0;
                if (p.estep && tid == 0) sm.ll[1] = alloc(((long long) C + 1) / 2);
                block_sync<NT>();       // hash cleared, groups visible, fwd allocation visible
                PASS_MARK(0);
                if (p.estep) {
                    fwd_off = sm.ll[1];
                    if (fwd_off < 0) { status = ST_NOMEM; return 2; }
                    if (tid == 0) { rec[i].fwd_off = fwd_off; rec[i].fwd_slots = C; }
                }
                int *fwd = (int *) (arena + fwd_off);

                auto decode = [&](int c, int &g, int &s, int &dir) {
                    g = 0;
                    while (g + 1 < ng && c >= sm.gbase[g + 1]) ++g;
                    const int r = c - sm.gbase[g];
                    if (sm.gx[g] != sm.gy[g]) { s = r >> 1; dir = r & 1; } else { s = r; dir = 0; }
                };

                // ---- pass A: gather successors (two candidates in flight per lane), canonicalise, hash-insert ----
                for (int c0 = tid; c0 < C; c0 += 2 * NT) {
                    const int cc[2] = {c0, c0 + NT};
                    bool live[2];
                    int4 ea[2], eb[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        live[q] = false;
                        if (cc[q] < C) {
                            int g, s, dir;
                            decode(cc[q], g, s, dir);
                            const int x = dir ? sm.gy[g] : sm.gx[g], y = dir ? sm.gx[g] : sm.gy[g];
                            const int pa = in_cache ? ci_ida[s] : gi_ida[s], pb = in_cache ? ci_idb[s] : gi_idb[s];
                            const double al = in_cache ? ci_alpha[s] : gi_alpha[s];
                            if (al > 0.0) {     // HaploBuilder.cpp:237
                                live[q] = true;
                                ea[q] = __ldcg(succ16 + (size_t) pa * A + x);     // L2 only: the 150 MB table would only evict the lists and hash from L1
                                eb[q] = __ldcg(succ16 + (size_t) pb * A + y);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (cc[q] >= C) continue;
                        int enc = -1;
                        if (live[q]) {
                            int hpa = ea[q].x, hpb = eb[q].x;
                            if (hpa >= 0 && hpb >= 0) {
                                const double tv = __dmul_rn(__hiloint2double(ea[q].w, ea[q].z), __hiloint2double(eb[q].w, eb[q].z));
                                u32 rev = 0;
                                if (hpa > hpb) { const int t = hpa; hpa = hpb; hpb = t; rev = 1; }
                                const u64 key = ((u64) (u32) hpa << 32) | (u32) hpb;
                                u32 h = mix64(key) & hmask;
                                for (;;) {
                                    const u64 prev = atomicCAS(keys + h, KEY_EMPTY, key);
                                    if (prev == KEY_EMPTY) { tpvs[h] = tv; break; }
                                    if (prev == key) break;
                                    h = (h + 1) & hmask;
                                }
                                atomicMin(ord + h, (u32) cc[q]);
                                atomicAdd(cnt + h, 1u);
                                enc = (int) (h | (rev << 30));
                            }
                        }
                        cand[cc[q]] = enc;
                        if (p.estep && enc < 0) fwd[cc[q]] = -1;
                    }
                }
                block_sync<NT>();
                PASS_MARK(1);

                // ---- pass B+C: creators in ordinal order -> new state indices and arrival-segment offsets ----
                // each lane owns UB consecutive ordinals, so the loads of a lane are independent and one block scan serves UB*NT slots
                int n_arr = 0;
                constexpr int UB = 4;
                for (int base = 0; base < C; base += NT * UB) {
                    const int c0 = base + tid * UB;
                    int enc[UB], nn[UB];
                    bool cr[UB];
                    u32 ov[UB];
                    if (c0 + UB <= C) { const int4 e4 = *reinterpret_cast<const int4 *>(cand + c0); enc[0] = e4.x; enc[1] = e4.y; enc[2] = e4.z; enc[3] = e4.w; }
                    else {
#pragma unroll
                        for (int u = 0; u < UB; ++u) enc[u] = (c0 + u < C) ? cand[c0 + u] : -1;
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) ov[u] = enc[u] >= 0 ? ord[enc[u] & 0x3FFFFFFF] : 0xFFFFFFFFu;
                    int ncr = 0, nar = 0;
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        cr[u] = enc[u] >= 0 && ov[u] == (u32) (c0 + u);
                        nn[u] = cr[u] ? (int) cnt[enc[u] & 0x3FFFFFFF] : 0;
                        ncr += cr[u] ? 1 : 0;
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) nar += nn[u];
                    // one scan for both: creators | arrivals
                    long long tot;
                    const long long ex = BlockScan<NT>::run(((long long) ncr << 40) | (long long) nar, tot, sm.tmp);
                    int idx = S_out + (int) ((ex >> 40) & 0xFFF), off = n_arr + (int) (ex & 0xFFFFFFFFFFll);
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        if (cr[u]) {
                            const int h = enc[u] & 0x3FFFFFFF;
                            if (idx < sc.scap) {
                                ord[h] = 0x80000000u | (u32) idx;     // never equals a later ordinal
                                slot_of[idx] = h;
                                aoff[idx] = off;
                                cnt[h] = 0;
                            }
                            ++idx; off += nn[u];
                        }
                    }
                    S_out += (int) ((tot >> 40) & 0xFFF);
                    n_arr += (int) (tot & 0xFFFFFFFFFFll);
                }
                if (S_out == 0) return 1;       // every extension died: m_haplopairs[i+1] empty, HaploBuilder.cpp:80-82
                if (S_out > sc.scap || S_out >= MAX_STATES_PER_LOCUS) { status = ST_TOO_BIG; return 2; }
                if (tid == 0) { sm.ll[1] = alloc(block_layout(S_out, K, p.estep).total); aoff[S_out] = n_arr; }
                c_arr += n_arr;
                block_sync<NT>();
                new_off = sm.ll[1];
                if (new_off < 0) { status = ST_NOMEM; return 2; }
                PASS_MARK(2);

                // ---- pass D: scatter arrivals (unordered inside a segment; sorted in E) ----
                for (int base = 0; base < C; base += NT * UB) {
                    const int c0 = base + tid * UB;
                    int enc[UB], idxv[UB], pos[UB];
                    if (c0 + UB <= C) { const int4 e4 = *reinterpret_cast<const int4 *>(cand + c0); enc[0] = e4.x; enc[1] = e4.y; enc[2] = e4.z; enc[3] = e4.w; }
                    else {
#pragma unroll
                        for (int u = 0; u < UB; ++u) enc[u] = (c0 + u < C) ? cand[c0 + u] : -1;
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        idxv[u] = 0; pos[u] = 0;
                        if (enc[u] >= 0) { const int h = enc[u] & 0x3FFFFFFF; idxv[u] = (int) (ord[h] & 0x7FFFFFFFu); pos[u] = (int) atomicAdd(cnt + h, 1u); }
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        if (enc[u] >= 0) {
                            const u32 rev = ((u32) enc[u] >> 30) & 1u;
                            arr[aoff[idxv[u]] + pos[u]] = (int) ((u32) (c0 + u) | (rev << 31));
                            if (p.estep) fwd[c0 + u] = (int) ((u32) idxv[u] | (rev << 31));
                        }
                    }
                }
                block_sync<NT>();
                PASS_MARK(3);

                // ---- pass E: one lane per new state ----
                const BlockLayout blo = block_layout(S_out, K, p.estep);
                int *t_ida = (int *) (arena + new_off + blo.ida), *t_idb = (int *) (arena + new_off + blo.idb);
                double *t_alpha = (double *) (arena + new_off + blo.alpha);
                unsigned char *t_ab = (unsigned char *) (arena + new_off + blo.ab);
                u32 *t_bp = (u32 *) (arena + new_off + blo.bp);
                double *t_tp = (double *) (arena + new_off + blo.tp);
                int *co_ida = sm.ida[par ^ 1], *co_idb = sm.idb[par ^ 1];
                double *co_alpha = sm.alpha[par ^ 1];
                const bool out_cache = S_out <= CC;
                const bool lk_out_cache = S_out <= CL;
                double amax = 0.0;
                for (int idx = tid; idx < S_out; idx += NT) {
                    const int h = slot_of[idx];
                    const u64 key = keys[h];
                    const int hpa = (int) (key >> 32), hpb = (int) (key & 0xFFFFFFFFu);
                    const double tpv = tpvs[h];
                    const int a0 = aoff[idx], n = aoff[idx + 1] - a0;
                    // ordinal order inside the segment: up to 8 arrivals are sorted in registers (one batch of loads, a
                    // 19-comparator network, one batch of stores); longer segments (< 1 % of arrivals) by insertion in place
                    if (n > 1 && n <= 8) {
                        int v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = (q < n) ? arr[a0 + q] : 0x7FFFFFFF;
                        auto cx = [&](int a, int b) { const int x = v[a], y = v[b]; const bool sw = (x & 0x7FFFFFFF) > (y & 0x7FFFFFFF); v[a] = sw ? y : x; v[b] = sw ? x : y; };
                        cx(0, 1); cx(2, 3); cx(4, 5); cx(6, 7);
                        cx(0, 2); cx(1, 3); cx(4, 6); cx(5, 7);
                        cx(1, 2); cx(5, 6); cx(0, 4); cx(3, 7);
                        cx(1, 5); cx(2, 6);
                        cx(1, 4); cx(3, 6);
                        cx(2, 4); cx(3, 5);
                        cx(3, 4);
#pragma unroll
                        for (int q = 0; q < 8; ++q) if (q < n) arr[a0 + q] = v[q];
                    }
                    else {
                        for (int j = 1; j < n; ++j) {
                            const int e = arr[a0 + j];
                            int q = j - 1;
                            while (q >= 0 && (arr[a0 + q] & 0x7FFFFFFF) > (e & 0x7FFFFFFF)) { arr[a0 + q + 1] = arr[a0 + q]; --q; }
                            arr[a0 + q + 1] = e;
                        }
                    }
                    double alpha = 0.0;
                    double bl[KMAX];
                    u32 bb[KMAX];
                    int nl_out = 0, xa = 0, ya = 0;
                    // software pipeline over the arrivals: the predecessor's alpha and list header of arrival j+1 are requested
                    // before arrival j is merged (the list itself is read in one batch of independent loads)
                    int nx_e = arr[a0], nx_g, nx_s, nx_dir;
                    decode(nx_e & 0x7FFFFFFF, nx_g, nx_s, nx_dir);
                    u64 nx_mt = lk_in_cache ? sm.meta[par][nx_s] : meta_in[nx_s];
                    double nx_ain = in_cache ? ci_alpha[nx_s] : gi_alpha[nx_s];
                    double nx_l0 = lk_in_cache ? sm.like[par][nx_s * KMAX] : like_in[(size_t) nx_s * K];    // every list has >= 1 link
                    for (int j = 0; j < n; ++j) {
                        const int e = nx_e, g = nx_g, s = nx_s, dir = nx_dir;
                        const u64 mt = nx_mt;
                        const double ain = nx_ain;
                        const double l0 = nx_l0;
                        const int nl = (int) (mt & 0xFF);
                        // the predecessor's list (sorted descending) is read straight from where it lives: the first arrival copies it in one
                        // batch of independent loads, later arrivals stream it with the next element always in flight
                        const double *lp = lk_in_cache ? &sm.like[par][s * KMAX] : like_in + (size_t) s * K;
                        if (j + 1 < n) {
                            nx_e = arr[a0 + j + 1];
                            decode(nx_e & 0x7FFFFFFF, nx_g, nx_s, nx_dir);
                            nx_mt = lk_in_cache ? sm.meta[par][nx_s] : meta_in[nx_s];
                            nx_ain = in_cache ? ci_alpha[nx_s] : gi_alpha[nx_s];
                            nx_l0 = lk_in_cache ? sm.like[par][nx_s * KMAX] : like_in[(size_t) nx_s * K];
                        }
                        const u32 rev = (u32) e >> 31;
                        const bool het = (sm.gx[g] != sm.gy[g]);
                        if (j == 0) { const int xx = dir ? sm.gy[g] : sm.gx[g], yy = dir ? sm.gx[g] : sm.gy[g]; xa = rev ? yy : xx; ya = rev ? xx : yy; }
                        u32 hom = (u32) (mt >> 8);
                        const double term = __dmul_rn(ain, tpv);
                        alpha = (j == 0) ? term : __dadd_rn(alpha, term);
                        nl_out += nl;
                        // het-locus rule (HaploPair.cpp:53-60,77-84): links still flagged homozygous lose the flag, and
                        // are zeroed when this arrival is reversed
                        const u32 zero = (het && rev) ? hom : 0u;
                        if (het) hom = 0;
                        const u32 bpbase = (u32) s | (rev ? BP_REV : 0u);
                        if (j == 0 && zero == 0) {
                            // scaling keeps the order: the first arrival's list is the initial top-K as it stands
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) {
                                bl[k] = (k < nl) ? __dmul_rn(lp[k], tpv) : -1.0;
                                bb[k] = bpbase | ((u32) k << BP_PRED_BITS) | (((hom >> k) & 1u) ? BP_HOM : 0u);
                            }
                        }
                        else {
                            if (j == 0) {
#pragma unroll
                                for (int k = 0; k < KMAX; ++k) { bl[k] = -1.0; bb[k] = 0; }
                            }
                            // sorted input: stop at the first element that does not make the list (unless some were zeroed)
                            int k = 0;
                            double cur = l0;
                            while (k < nl) {
                                const double nxt = (k + 1 < nl) ? lp[k + 1] : 0.0;
                                double lv = __dmul_rn(cur, tpv);
                                if ((zero >> k) & 1u) lv = 0.0;
                                if (lv > bl[KMAX - 1]) topk_insert<KMAX>(bl, bb, lv, bpbase | ((u32) k << BP_PRED_BITS) | (((hom >> k) & 1u) ? BP_HOM : 0u));
                                else if (zero == 0) break;
                                cur = nxt;
                                ++k;
                            }
                        }
                    }
                    if (nl_out > K) nl_out = K;
                    u32 hom_out = 0;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) {
                        if (k < nl_out) {
                            if (lk_out_cache) sm.like[par ^ 1][idx * KMAX + k] = bl[k];
                            else like_out[(size_t) idx * K + k] = bl[k];
                            if (!p.estep) t_bp[(size_t) idx * K + k] = bb[k];
                            if (bb[k] & BP_HOM) hom_out |= 1u << k;
                        }
                    }
                    const u64 mo = (u64) nl_out | ((u64) hom_out << 8);
                    if (lk_out_cache) sm.meta[par ^ 1][idx] = mo; else meta_out[idx] = mo;
                    t_ida[idx] = hpa; t_idb[idx] = hpb; t_alpha[idx] = alpha;
                    t_ab[idx] = (unsigned char) (xa | (ya << 4));
                    if (p.estep) t_tp[idx] = tpv;
                    if (out_cache) { co_ida[idx] = hpa; co_idb[idx] = hpb; co_alpha[idx] = alpha; }
                    amax = fmax(amax, alpha);
                }
                PASS_MARK(4);
                if (p.scaled) {     // exponent-only rescale: mantissas stay those of the unscaled product
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                    if (NT > 32) {
                        __syncthreads();
                        if ((tid & 31) == 0) ((double *) sm.tmp)[tid >> 5] = amax;
                        __syncthreads();
                        amax = 0.0;
                        for (int q = 0; q < NT / 32; ++q) amax = fmax(amax, ((double *) sm.tmp)[q]);
                    }
                    if (amax > 0.0 && amax < 0x1p-512) {
                        block_sync<NT>();
                        for (int idx = tid; idx < S_out; idx += NT) {
                            const double a2 = t_alpha[idx] * 0x1p+512;
                            t_alpha[idx] = a2;
                            if (out_cache) co_alpha[idx] = a2;
                            if (lk_out_cache) {
                                const int nl = (int) (sm.meta[par ^ 1][idx] & 0xFF);
                                for (int k = 0; k < nl; ++k) sm.like[par ^ 1][idx * KMAX + k] *= 0x1p+512;
                            }
                            else {
                                const int nl = (int) (meta_out[idx] & 0xFF);
                                for (int k = 0; k < nl; ++k) like_out[(size_t) idx * K + k] *= 0x1p+512;
                            }
                        }
                        new_exp += 512;
                    }
                }
                return 0;
            };
            const int rc = fast ? body(std::true_type()) : body(std::false_type());
            if (rc == 1) { S_in = 0; break; }
            if (rc == 2) break;
            if (tid == 0) {
                rec[i + 1].off = new_off; rec[i + 1].count = S_out;
                if (p.locus_exp) p.locus_exp[(size_t) w * (L + 1) + i + 1] = new_exp;
            }
            cur_exp = new_exp;
            ++c_steps; c_states += S_out; if (S_out > c_max) c_max = S_out;
            S_in = S_out; par ^= 1; cur_off = new_off;
            block_sync<NT>();
            PASS_MARK(5);
        }
        block_sync<NT>();
        if (status == ST_DONE && S_in > 0 && !p.estep) {
            // final link lists for the traceback kernel (selection over all final states' links)
            if (tid == 0) sm.ll[1] = alloc((long long) S_in * K + S_in);
            block_sync<NT>();
            const long long fo = sm.ll[1];
            if (fo < 0) status = ST_NOMEM;
            else {
                double *fl = (double *) (arena + fo);
                u64 *fm = arena + fo + (long long) S_in * K;
                const bool cached = S_in <= CL;
                for (int s = tid; s < S_in; s += NT) {
                    const u64 mt = cached ? sm.meta[par][s] : (par ? sc_meta1 : sc_meta0)[s];
                    fm[s] = mt;
                    const int nl = (int) (mt & 0xFF);
                    for (int k = 0; k < nl; ++k) fl[(size_t) s * K + k] = cached ? sm.like[par][s * KMAX + k] : (par ? sc_like1 : sc_like0)[(size_t) s * K + k];
                }
                if (tid == 0) { rec[0].off = fo; rec[0].count = S_in; }     // index 0 is unused by the trellis when head_len >= 1
            }
        }
        if (tid == 0) {
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
                for (int q = 0; q < 16; ++q) atomicAdd((unsigned long long *) p.counters + 8 + q, (unsigned long long) pass_cyc[q]);
#endif
            }
        }
    }
}

}  // namespace hmc
