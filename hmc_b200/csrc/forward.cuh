// K1 forward: the per-individual forward pass of HaploBuilder::resolve
// (HaploBuilder.cpp:35-86 with initHeadList :153-224, extendAll/extend/addHaploPair :226-261 and the
// HaploPair constructors / add, HaploPair.cpp:14-89).
//
// One CTA (NT threads) owns one individual at a time and pulls individuals from a work queue
// (persistent grid).  What the kernel materialises per trellis index is the state set in the
// reference's creation order with, per state, alpha (forward_likelihood), v1 = the likelihood of the
// best link (the head of m_best_links) and the list of arrivals (predecessor, reversed, heterozygous
// step) -- NOT the K-best link lists: those are a function of exactly this graph and are enumerated
// lazily from the end of the trellis by the K2 kernel (kbest.cuh), which touches ~K states per index
// instead of all of them.  alpha is accumulated over the arrivals in the reference's order with
// explicit __dmul_rn / __dadd_rn, so it is bit-identical; v1 = max over arrivals of v1(pred) * tp is
// order-free (rounding is monotone).
//
// Per locus step:
//   A   lanes over the predecessor states: one 128-bit gather per (haplotype, allele) from the successor
//       table {successor id, its transition probability}; per candidate (allele pair, orientation):
//       canonicalise (id_a <= id_b), insert (id_a,id_b) into an open-addressing hash, atomicMin of the
//       candidate ordinal, arrival count
//   C   candidates in ordinal order, warps take 128-candidate tiles round robin and chain their
//       (creators, arrivals) prefixes through a decoupled look-back ring in shared memory (no block
//       barrier): the candidate whose ordinal equals its slot's minimum creates the state -> new states
//       get exactly the reference's creation order; it writes ids / last alleles / arrival offset, and
//       when it is the state's only arrival (about two thirds of the states) alpha, v1 and the arrival
//       word as well, and frees its hash slot
//   D   arrivals of the remaining (multi-arrival) states are scattered into per-state segments
//   E   one lane per multi-arrival state: arrivals sorted by ordinal, alpha summed in that order,
//       v1 = max, arrival words written, hash slot freed.
// While a step fits (C <= CC candidate slots) the hash / candidate / segment arrays are in shared
// memory (the step body is instantiated twice: shared / global spill scratch); the live states'
// ids, alpha and v1 are cached in shared memory up to CC states.
#pragma once

#include <type_traits>

#ifndef FWD_SMEM_KEYS
#define FWD_SMEM_KEYS 2048      // hash slots whose keys (and, see FwdSmem, 32-bit fields) live in shared memory
#endif
#ifndef FWD_SMEM_META
#define FWD_SMEM_META 1
#endif
#ifndef FWD_SMEM_TPV
#define FWD_SMEM_TPV 0          // the slot's tp product too: 75 KB per CTA, the carve-out leaves 28 KB of L1 -- 111 vs 91 ms
#endif
// The successor rows a new state will gather in the next step are prefetched when pass C creates it: into L1 (and read through it) now that the
// hash of mid-size steps no longer competes for it -- 1.8 ms on C2 against the L2-only prefetch + ld.cg.
#ifndef FWD_PREFETCH_L1
#define FWD_PREFETCH_L1 1
#endif
// Register cap of the shipped 256-thread configuration.  Registers are allocated per SM sub-partition (16,384 each): the 6 forward warps a
// sub-partition holds at 3 CTAs per SM leave 1,024 registers at the 80 ptxas picks on its own (room for a 32-register warp: measured with a
// probe kernel) and 2,560 at 72, which is one streamed K-best warp at 80 registers beside them.  The cap costs the forward kernel 4 bytes
// more spill and 1.4 ms on C2; 64 / 64 (two K-best warps per sub-partition) costs it 7 ms.
#ifndef FWD_MAXNREG
#define FWD_MAXNREG 72
#endif

#include "common.cuh"

namespace hmc {

#ifdef HMC_PROFILE_PASSES
#define PASS_MARK(slot) do { const long long t_ = clock64(); if (tid == 0) { if (pass_fast) pass_cyc[slot] += t_ - t_mark; else pass_cyc[(slot) + 8] += t_ - t_mark; } t_mark = t_; } while (0)
#else
#define PASS_MARK(slot) do { } while (0)
#endif

// trellis outputs: written once, read much later (next step at the earliest, through L2) -> optionally streaming stores
template <typename T>
__device__ __forceinline__ void ST_OUT(T *p, T v)
{
#ifdef FWD_STREAM_OUT
    __stcs(p, v);
#else
    *p = v;
#endif
}

template <int NT>
__device__ __forceinline__ void block_sync()
{
    if (NT == 32) __syncwarp(); else __syncthreads();
}

// The per-step state hash: separate arrays per field, in shared memory (fast path) or in the CTA's global spill scratch.
// (One 32-byte record per slot was measured too: one sector per candidate instead of five, but the three or four atomics of a
// candidate then queue on the same L2 sector, and in shared memory a 32-byte stride puts every field on 4 bank groups -- slower.)
struct HashTab {
    u64 *keys;          // (id_a << 32 | id_b), KEY_EMPTY = free
    double *tpvs;       // transition-probability product of the state
    u32 *ord;           // smallest candidate ordinal that reached the slot; 0x80000000 | state index once created
    u32 *cnt;           // arrivals
    u32 *soffs;         // offset of the state's arrival segment
    __device__ __forceinline__ void release(u32 h) const { keys[h] = KEY_EMPTY; ord[h] = 0xFFFFFFFFu; cnt[h] = 0; }
    __device__ __forceinline__ void created(u32 h, u32 idx, u32 off) const { ord[h] = 0x80000000u | idx; cnt[h] = 0; soffs[h] = off; }
};

template <int NT, int CC>
struct FwdSmem {
    static constexpr int HC = ((CC + CC / 4 + CC / 16 + 63) >> 6) << 6;     // hash slots of a step with CC candidates (see hsize)
    // The KEYS of the hash stay in shared memory for steps of up to ~1,550 candidates (KB slots), far beyond the fast path: the insert's
    // compare-and-swap is the one access of the step a warp has to wait for slot by slot (probe chains), and on the L2-resident spill
    // hash it was a third of all warp time; the other slot fields of such a step live in the global scratch (reductions / later passes).
    static constexpr int KB = FWD_SMEM_KEYS;
    // ... and in the shipped 256-thread configuration (3 CTAs per SM, 65 KB each) so do the three 32-bit slot fields the later passes
    // update with atomics (smallest ordinal, arrival count, segment offset); the tp product of the slot is written once and read once and stays global
    static constexpr int MB = (FWD_SMEM_META && NT >= 256 && KB > HC) ? KB : HC;
    static constexpr int TB = FWD_SMEM_TPV ? MB : HC;
    u64 keys[KB > HC ? KB : HC];
    double tpv[TB];
    u32 ord[MB], cnt[MB], soff[MB];
    double alpha[2][CC], v1[2][CC];
    long long ll[4];                // -, alloc result, chunk cursor, chunk end
    unsigned long long look[2][8];  // look-back ring of pass C: tag << 48 | creators << 24 | arrivals (inclusive prefix of a tile); by step parity
    int cnts[2][4];                 // new states, arrivals, states with 2..8 / more arrivals of the step; by step parity (the other set is zeroed during a step)
    double red[NT / 32 + 1];
    alignas(16) int cand[CC];
    int sarr[CC];
    int2 mlist[CC / 2 + 1];
    int ida[2][CC], idb[2][CC];
    int gx[MAX_GROUPS], gy[MAX_GROUPS], gbase[MAX_GROUPS + 1];
    int ints[8];                    // 0 queue ticket, 1 head count, 2 status, 3 arena exhausted
};

template <int NT, int CC, int MINB, bool ESTEP>
__global__ void __launch_bounds__(NT, MINB) __maxnreg__(NT == 256 ? FWD_MAXNREG : (NT > 512 ? 64 : 255)) forward_kernel(FwdParams p)
{
    typedef FwdSmem<NT, CC> Smem;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    constexpr u32 FULL = 0xffffffffu;
    constexpr u32 SLOT_MASK = 0x3FFFFFFFu;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;
    const DevPanel &pn = p.panel;
    const DevModel &md = p.model;
    const int L = pn.L, A = md.Amax, head_len = md.head_len;
    constexpr bool estep = ESTEP;
    const Scratch sc = p.scratch[blockIdx.x];
    u64 *const arena = p.arena.base;
    const int4 *const succ16 = reinterpret_cast<const int4 *>(md.succ16);

    if (tid == 0) { sm.ll[2] = 0; sm.ll[3] = 0; }
    for (int h = tid; h < Smem::MB; h += NT) { sm.ord[h] = 0xFFFFFFFFu; sm.cnt[h] = 0; }
    for (int h = tid; h < (Smem::KB > Smem::HC ? Smem::KB : Smem::HC); h += NT) sm.keys[h] = KEY_EMPTY;

    // chunked bump allocation (thread 0 only); returns -1 when the arena is exhausted
    auto alloc = [&](long long units) -> long long {
        constexpr long long CHUNK = 1 << 15;
        units = (units + 1) & ~1LL;
        if (sm.ll[2] + units > sm.ll[3]) {
            const long long grab = units > CHUNK ? ((units + 15) & ~15LL) : CHUNK;       // whole 128-byte lines: see loci_stride
            const unsigned long long base = atomicAdd(p.arena.cursor, (unsigned long long) grab);
            if ((long long) (base + grab) > p.arena.capacity) {
                atomicAdd(p.arena.cursor, (unsigned long long) (-grab));       // give it back: the K-best kernel allocates behind the trellis
                atomicExch(p.arena.exhausted, 1);
                return -1;
            }
            sm.ll[2] = (long long) base; sm.ll[3] = (long long) base + grab;
        }
        const long long off = sm.ll[2];
        sm.ll[2] += units;
        return off;
    };

    for (;;) {
        block_sync<NT>();
        if (tid == 0) { sm.ints[0] = atomicAdd(p.queue, 1); sm.ints[3] = *(volatile int *) p.arena.exhausted; }     // one read per CTA: the decision is block-uniform
        block_sync<NT>();
        const int w = sm.ints[0];
        if (w >= p.n_wave) break;
        // hands wave position w to the K-best warps that run beside this kernel (thread 0, after the barrier that ends the individual's last
        // step: the fence makes the CTA's trellis writes visible before the list entry is)
        auto publish = [&](int w_) { if (p.done_list) { __threadfence(); atomicExch(p.done_list + atomicAdd(p.done_count, 1), w_); } };
        if (sm.ints[3]) { if (tid == 0) { p.status[w] = ST_NOMEM; publish(w); } continue; }
        if (tid == 0) sm.ll[2] = (sm.ll[2] + 15) & ~15LL;       // an individual's blocks start on a fresh 128-byte line (see loci_stride)
        const long long t_begin = clock64();
        long long t_begin_ns = 0;
        if (p.ind_cycles) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); t_begin_ns = (long long) t_; }
        const int ind = p.wave[w];
        const u32 *gp = pn.packed + (size_t) ind * pn.words;
        LocusRec *rec = p.loci + (size_t) w * loci_stride(L);
#pragma unroll 1
        for (int l = tid; l <= L; l += NT) { rec[l].count = 0; rec[l].fwd_slots = 0; rec[l].off = 0; rec[l].fwd_off = 0; }
        const int gbits = pn.bits;
        const u32 gmask = (1u << gbits) - 1u;
        auto geno_word = [&](int l) -> u32 { const int bit = l * 2 * gbits; return gp[bit >> 5] >> (bit & 31); };

        if (tid < 16) (&sm.look[0][0])[tid] = 0;
        if (tid < 8) (&sm.cnts[0][0])[tid] = 0;
        // ---------------- head states (initHeadList), serial: at most A^head_len pairs ----------------
        int S_in = 0, par = 0, status = ST_DONE, cur_exp = 0;
        long long cur_off = 0;
        if (tid == 0) {
            int nh = 0, bad = 0;
            int hg0[8], hg1[8];
#pragma unroll 1
            for (int j = 0; j < head_len; ++j) { const u32 wd = geno_word(j); hg0[j] = (int) (wd & gmask) - 1; hg1[j] = (int) ((wd >> gbits) & gmask) - 1; }
#pragma unroll 1
            for (int hi = 0; hi < md.n_heads && !bad; ++hi) {
                const int head = md.heads[hi];
                const signed char *ha = md.al + md.al_off[head];
                bool match = true;
#pragma unroll 1
                for (int j = 0; j < head_len; ++j) {
                    const int a = ha[j];
                    if (!(hg0[j] < 0 || hg1[j] < 0 || hg0[j] == a || hg1[j] == a)) { match = false; break; }
                }
                if (!match) continue;
                // odometer over the complementary allele strings, locus 0 most significant
                int lo[8], hi_[8], curv[8];
#pragma unroll 1
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
#pragma unroll 1
                    for (int j = 0; j < head_len; ++j) {
                        if (lo[j] != hi_[j] && !(((pn.linfo[j] >> 8) >> curv[j]) & 1u)) ok = false;
                        code += curv[j] * md.head_stride[j];
                    }
                    if (ok) {
                        const int hp = md.head_lut[code];
                        if (hp < 0) { bad = 1; break; }
                        if (hp >= head) {
                            if (nh < CC) { sm.ida[0][nh] = head; sm.idb[0][nh] = hp; }
                            if (2 * nh + 1 < sc.ccap) { sc.cand[2 * nh] = head; sc.cand[2 * nh + 1] = hp; }    // staging when nh >= CC
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
            else if (nh > 0) { off = alloc(block_layout(nh, 0, p.estep).total); if (off < 0) st = ST_NOMEM; }
            sm.ints[1] = nh; sm.ints[2] = st; sm.ll[1] = off;
        }
        block_sync<NT>();
        S_in = sm.ints[1]; status = sm.ints[2]; cur_off = sm.ll[1];
        if (status != ST_DONE) { if (tid == 0) { p.status[w] = status; publish(w); } continue; }
        {
            const BlockLayout bl = block_layout(S_in, 0, p.estep);
            int *t_ida = (int *) (arena + cur_off + bl.ida), *t_idb = (int *) (arena + cur_off + bl.idb);
            double *t_alpha = (double *) (arena + cur_off + bl.alpha);
            unsigned char *t_ab = (unsigned char *) (arena + cur_off + bl.ab);
            double *t_v1 = (double *) (arena + cur_off + bl.v1);
            int *t_aoff = (int *) (arena + cur_off + bl.aoff);
            double *t_tp = (double *) (arena + cur_off + bl.tp);
            for (int s = tid; s < S_in; s += NT) {
                const int a = s < CC ? sm.ida[0][s] : sc.cand[2 * s], b = s < CC ? sm.idb[0][s] : sc.cand[2 * s + 1];
                const double f = __dmul_rn(md.freq[a], md.freq[b]);
                const double al = (a == b) ? f : __dmul_rn(f, 2.0);
                if (s < CC) { sm.alpha[0][s] = al; sm.v1[0][s] = f; }
                t_ida[s] = a; t_idb[s] = b; t_alpha[s] = al;
                t_ab[s] = (unsigned char) (md.last[a] | (md.last[b] << 4));
                if (estep) t_tp[s] = f;
                else { t_v1[s] = f; t_aoff[s] = 0; if (s == 0) t_aoff[S_in] = 0; }
            }
            if (tid == 0) { rec[head_len].off = cur_off; rec[head_len].count = S_in; }
        }
#ifdef HMC_PROFILE_PASSES
        long long pass_cyc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool pass_fast = true;
        long long t_mark = clock64();
        long long hist_cyc[18], hist_n[18], t_step = 0;     // steps by candidate count: bucket b holds C in [2^(b-1), 2^b)
        for (int q = 0; q < 18; ++q) { hist_cyc[q] = 0; hist_n[q] = 0; }
        int hist_b = 0;
#endif
        par = 0;
        long long c_steps = 0, c_states = S_in, c_max = S_in, c_arr = 0, c_spill = 0;
        if (p.locus_exp && tid == 0) for (int l = 0; l <= head_len; ++l) p.locus_exp[(size_t) w * exp_stride(L) + l] = 0;

        // genotype word and locus info of the next step are fetched one step ahead
        u32 nx_word = 0, nx_info = 0;
        if (head_len < L) { nx_word = geno_word(head_len); nx_info = pn.linfo[head_len]; }

        // ---------------- locus steps ----------------
        // The step is written for a SMALL instruction footprint: one instantiation (shared-memory or spill arrays chosen at run time,
        // generic addressing), no unrolled per-lane batches -- the kernel is instruction-fetch bound otherwise (profiles/README.md:
        // the GPC-level instruction cache ran at 87 % of its request rate with the 130 KB body of the first round-2 version).
#pragma unroll 1
        for (int i = head_len; i < L && S_in > 0; ++i) {
            PASS_MARK(7);
            const int g0 = (int) (nx_word & gmask) - 1, g1 = (int) ((nx_word >> gbits) & gmask) - 1;
            const int nA = (int) (nx_info & 0xFF);
            const u32 am = nx_info >> 8;
            if (i + 1 < L) { nx_word = geno_word(i + 1); nx_info = pn.linfo[i + 1]; }
            // allele-pair groups in the reference's extendAll order (HaploBuilder.cpp:48-79): one group at a typed locus; at a locus with
            // missing alleles thread 0 lists them
            // (a typed locus -- 95 % of the steps -- keeps its one group in registers: no shared-memory traffic and no barrier for it)
            int ng = 1, C = S_in * (g0 != g1 ? 2 : 1);
            const bool typed = g0 >= 0 && g1 >= 0;
            if (!typed) {
                if (tid == 0) {
                    int n = 0, c = 0;
                    auto add_group = [&](int x, int y) { sm.gx[n] = x; sm.gy[n] = y; sm.gbase[n] = c; c += S_in * (x != y ? 2 : 1); ++n; };
                    if (g0 < 0 && g1 < 0) {
#pragma unroll 1
                        for (int j = 0; j < nA; ++j) {
                            if (!((am >> j) & 1u)) continue;
#pragma unroll 1
                            for (int k = j; k < nA; ++k) if ((am >> k) & 1u) add_group(j, k);
                        }
                    }
                    else {
                        const int known = g0 < 0 ? g1 : g0;
#pragma unroll 1
                        for (int j = 0; j < nA; ++j) if ((am >> j) & 1u) add_group(j, known);
                    }
                    sm.gbase[n] = c; sm.ints[1] = n;
                }
                block_sync<NT>();
                ng = sm.ints[1]; C = sm.gbase[ng];
            }
            // step counters and look-back ring of this step (zeroed during the previous one: no barrier needed before they are used)
            int *const cn = sm.cnts[par];
            unsigned long long *const lk = sm.look[par];
            // hash size: 1.3 C slots (load <= 0.77 even if every candidate were a new state, ~0.45 typically), NOT rounded to a power of
            // two: the slots a step touches lie 2-3x denser, and the spill hash of all resident CTAs has to live in L2 between passes
            const int hsize = ((C + C / 4 + C / 16 + 63) >> 6) << 6;
            const bool fast = (C <= CC);
            if (!fast && (C > sc.ccap || hsize > sc.hcap)) { status = ST_TOO_BIG; break; }
            if (!fast) ++c_spill;
#ifdef HMC_PROFILE_PASSES
            pass_fast = fast;
            if (tid == 0) { if (fast) pass_cyc[6] += C; else pass_cyc[14] += C; }
            hist_b = min(17, 32 - __clz(max(C, 1))); t_step = clock64();
#endif

            int S_out = 0, new_exp = cur_exp;
            long long new_off = 0;
            int rc = 0;         // 0 = ok, 1 = every extension died, 2 = status set (the hash arrays were reset)
            do {
                HashTab tab;
                tab.keys = (fast || hsize <= Smem::KB) ? sm.keys : sc.keys; tab.tpvs = (fast || hsize <= Smem::TB) ? sm.tpv : sc.tpv;
                const bool meta_sm = fast || hsize <= Smem::MB;
                tab.ord = meta_sm ? sm.ord : sc.ord; tab.cnt = meta_sm ? sm.cnt : sc.cnt; tab.soffs = meta_sm ? sm.soff : sc.soff;
                int *const cand = fast ? sm.cand : sc.cand;
                int *const sarr = fast ? sm.sarr : sc.sarr;
                int2 *const mlist = fast ? sm.mlist : sc.mlist;
                const int mcap = fast ? (CC / 2 + 1) : sc.scap;
                const u32 hsz = (u32) hsize;
                auto reset_table = [&]() {      // error exits: leave the hash arrays all-empty for the next individual
                    block_sync<NT>();
#pragma unroll 1
                    for (int h = tid; h < hsize; h += NT) tab.release((u32) h);
                };

                // input state views
                const BlockLayout bli = block_layout(S_in, 0, p.estep);
                const bool in_cache = S_in <= CC;
                const int *const in_ida = in_cache ? sm.ida[par] : (const int *) (arena + cur_off + bli.ida);
                const int *const in_idb = in_cache ? sm.idb[par] : (const int *) (arena + cur_off + bli.idb);
                const double *const in_alpha = in_cache ? sm.alpha[par] : (const double *) (arena + cur_off + bli.alpha);
                const double *const in_v1 = in_cache ? sm.v1[par] : (const double *) (arena + cur_off + bli.v1);

                // forward-link block of index i (E-step mode)
                long long fwd_off = 0;
                if (estep && tid == 0) sm.ll[1] = alloc(((long long) C + 1) / 2);
                if (estep) block_sync<NT>();       // fwd allocation visible
                PASS_MARK(0);
                if (estep) {
                    fwd_off = sm.ll[1];
                    if (fwd_off < 0) { status = ST_NOMEM; rc = 2; break; }       // nothing inserted yet: the hash arrays are clean
                    if (tid == 0) { rec[i].fwd_off = fwd_off; rec[i].fwd_slots = C; }
                }
                int *fwd = (int *) (arena + fwd_off);

                // (a typed locus is one group: the ordinal is the state, or state * 2 + orientation at a heterozygous one)
                const bool one_group = ng == 1;
                const int gx0 = typed ? g0 : sm.gx[0], gy0 = typed ? g1 : sm.gy[0];
                const bool het0 = gx0 != gy0;
                auto decode = [&](int c, int &g, int &s, int &dir) {
                    g = 0;
                    if (one_group) { s = het0 ? c >> 1 : c; dir = het0 ? (c & 1) : 0; return; }
#pragma unroll 1
                    while (g + 1 < ng && c >= sm.gbase[g + 1]) ++g;
                    const int r = c - sm.gbase[g];
                    if (sm.gx[g] != sm.gy[g]) { s = r >> 1; dir = r & 1; } else { s = r; dir = 0; }
                };

                // ---- pass A: lanes over predecessor states; gather successors, canonicalise, hash-insert ----
                // (two states per lane and round -- four compare-and-swaps in flight -- was measured: 18 % SLOWER; the inserts do not overlap)
                int my_new = 0, my_arr = 0;
                {
                    int nx_pa = 0, nx_pb = 0;
                    double nx_al = 0.0;
                    if (tid < S_in) { nx_pa = in_ida[tid]; nx_pb = in_idb[tid]; nx_al = in_alpha[tid]; }
#pragma unroll 1
                    for (int s = tid; s < S_in; s += NT) {
                        const int pa = nx_pa, pb = nx_pb;
                        const bool alive = nx_al > 0.0;       // HaploBuilder.cpp:237
                        if (s + NT < S_in) { nx_pa = in_ida[s + NT]; nx_pb = in_idb[s + NT]; nx_al = in_alpha[s + NT]; }      // next round's state, one round ahead
                        const int4 *ra = succ16 + (size_t) pa * A, *rb = succ16 + (size_t) pb * A;
#pragma unroll 1
                        for (int g = 0; g < ng; ++g) {
                            int x = gx0, y = gy0, gb = 0;
                            if (!one_group) { x = sm.gx[g]; y = sm.gy[g]; gb = sm.gbase[g]; }
                            const bool het = x != y;
                            const int c0 = gb + (het ? 2 * s : s);
                            const int nd = het ? 2 : 1;
                            int4 ea[2], eb[2];
                            ea[0] = ea[1] = eb[0] = eb[1] = make_int4(-1, 0, 0, 0);
                            if (alive) {        // (prefetched by the pass C that created the state)
#if FWD_PREFETCH_L1
                                ea[0] = ra[x]; eb[0] = rb[y];
                                if (het) { ea[1] = ra[y]; eb[1] = rb[x]; }
#else
                                ea[0] = __ldcg(ra + x); eb[0] = __ldcg(rb + y);
                                if (het) { ea[1] = __ldcg(ra + y); eb[1] = __ldcg(rb + x); }
#endif
                            }
                            // both orientations: keys first, then the two compare-and-swaps in flight together
                            u64 key[2];
                            u32 hh[2], rev[2];
                            double tv[2];
                            bool ok[2];
                            u64 prev[2];
#pragma unroll
                            for (int d = 0; d < 2; ++d) {
                                ok[d] = alive && d < nd && ea[d].x >= 0 && eb[d].x >= 0;
                                key[d] = 0; hh[d] = 0; rev[d] = 0; tv[d] = 0.0; prev[d] = 0;
                                if (ok[d]) {
                                    int hpa = ea[d].x, hpb = eb[d].x;
                                    tv[d] = __dmul_rn(__hiloint2double(ea[d].w, ea[d].z), __hiloint2double(eb[d].w, eb[d].z));
                                    if (hpa > hpb) { const int t = hpa; hpa = hpb; hpb = t; rev[d] = 1; }
                                    key[d] = ((u64) (u32) hpa << 32) | (u32) hpb;
                                    hh[d] = __umulhi(mix64(key[d]), hsz);
                                    prev[d] = atomicCAS(tab.keys + hh[d], KEY_EMPTY, key[d]);
                                }
                            }
#pragma unroll
                            for (int d = 0; d < 2; ++d) {
                                if (d >= nd) continue;
                                int enc = -1;
                                if (ok[d]) {
                                    u32 h = hh[d];
                                    u64 pv = prev[d];
#pragma unroll 1
                                    for (;;) {
                                        if (pv == KEY_EMPTY) { tab.tpvs[h] = tv[d]; ++my_new; break; }
                                        if (pv == key[d]) break;
                                        h = (h + 1 == hsz) ? 0u : h + 1;
                                        pv = atomicCAS(tab.keys + h, KEY_EMPTY, key[d]);
                                    }
                                    atomicMin(tab.ord + h, (u32) (c0 + d));
                                    atomicAdd(tab.cnt + h, 1u);
                                    ++my_arr;
                                    enc = (int) (h | (rev[d] << 30));
                                }
                                cand[c0 + d] = enc;
                                if (estep && enc < 0) fwd[c0 + d] = -1;
                            }
                        }
                    }
                }
                // new states / arrivals of the step
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { my_new += __shfl_xor_sync(FULL, my_new, o); my_arr += __shfl_xor_sync(FULL, my_arr, o); }
                if (lane == 0 && my_arr) { atomicAdd(&cn[0], my_new); atomicAdd(&cn[1], my_arr); }
                block_sync<NT>();
                PASS_MARK(1);
                S_out = cn[0];
                const int n_arr = cn[1];
                if (tid < 4) sm.cnts[par ^ 1][tid] = 0;         // the next step's set: its last users finished before the previous end-of-step barrier
                else if (tid < 12) sm.look[par ^ 1][tid - 4] = 0;
                if (S_out == 0) { rc = 1; break; }      // every extension died: m_haplopairs[i+1] empty, HaploBuilder.cpp:80-82 (nothing was inserted)
                if (S_out > sc.scap || S_out >= MAX_STATES_PER_LOCUS) { status = ST_TOO_BIG; reset_table(); rc = 2; break; }
                if (tid == 0) sm.ll[1] = alloc(block_layout(S_out, n_arr, p.estep).total);
                block_sync<NT>();
                new_off = sm.ll[1];
                if (new_off < 0) { status = ST_NOMEM; reset_table(); rc = 2; break; }
                c_arr += n_arr;

                const BlockLayout blo = block_layout(S_out, n_arr, p.estep);
                u64 *const oblk = arena + new_off;
                int *t_ida = (int *) (oblk + blo.ida), *t_idb = (int *) (oblk + blo.idb);
                double *t_alpha = (double *) (oblk + blo.alpha);
                double *t_v1 = (double *) (oblk + blo.v1);
                unsigned char *t_ab = (unsigned char *) (oblk + blo.ab);
                int *t_aoff = (int *) (oblk + blo.aoff);
                u32 *t_arr = (u32 *) (oblk + blo.arr);
                double *t_tp = (double *) (oblk + blo.tp);
                int *co_ida = sm.ida[par ^ 1], *co_idb = sm.idb[par ^ 1];
                double *co_alpha = sm.alpha[par ^ 1], *co_v1 = sm.v1[par ^ 1];
                const bool out_cache = S_out <= CC;
                double amax = 0.0;
                if (!estep && tid == 0) t_aoff[S_out] = n_arr;

                // ---- pass C: creators in ordinal order -> state indices and arrival offsets; single-arrival states are finished here ----
                // a warp takes 32 consecutive ordinals per tile; the (creators, arrivals) prefixes of the tiles are chained through the look-back ring
                {
                    const int ntiles = (C + 31) >> 5;
                    const u32 lt_mask = (1u << lane) - 1u;
                    int enc_nx = (wid < ntiles && (wid << 5) + lane < C) ? cand[(wid << 5) + lane] : -1;
#pragma unroll 1
                    for (int tile = wid; tile < ntiles; tile += NW) {
                        const int c = (tile << 5) + lane;
                        const int enc = enc_nx;
                        { const int c2 = c + (NW << 5); enc_nx = (c2 < C) ? cand[c2] : -1; }      // the next tile's slots, one trip ahead
                        const u32 h = (u32) enc & SLOT_MASK;
                        u32 so = 0xFFFFFFFFu, sn = 0;
                        u64 key = 0;
                        double tv = 0.0, pa = 0.0, pv = 0.0;
                        int g = 0, s = 0, dir = 0;
                        if (enc >= 0) {     // the slot's four fields and the predecessor's alpha / v1 (used if this turns out a single-arrival creator): one trip
                            so = tab.ord[h]; sn = tab.cnt[h]; key = tab.keys[h]; tv = tab.tpvs[h];
                            decode(c, g, s, dir);
                            pa = in_alpha[s];
                            if (!estep) pv = in_v1[s];
                        }
                        const bool cr = enc >= 0 && so == (u32) c;
                        const int nn = cr ? (int) sn : 0;
                        const u32 crm = __ballot_sync(FULL, cr);
                        int xa = nn;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const int ya = __shfl_up_sync(FULL, xa, o); if (lane >= o) xa += ya; }
                        const int tot_ar = __shfl_sync(FULL, xa, 31);
                        unsigned long long prevp = 0;
                        if (lane == 0) {
                            if (tile > 0) {
                                volatile unsigned long long *slot = &lk[(tile - 1) & 7];
                                const unsigned long long want = (unsigned long long) (tile & 0xFFFF);
                                do { prevp = *slot; } while ((prevp >> 48) != want);
                            }
                            const unsigned long long b_idx = (prevp >> 24) & 0xFFFFFFull, b_off = prevp & 0xFFFFFFull;
                            *(volatile unsigned long long *) &lk[tile & 7] =
                                ((unsigned long long) ((tile + 1) & 0xFFFF) << 48) | ((b_idx + (unsigned long long) __popc(crm)) << 24) | (b_off + (unsigned long long) tot_ar);
                        }
                        prevp = __shfl_sync(FULL, prevp, 0);
                        if (!cr) continue;
                        const int idx = (int) ((prevp >> 24) & 0xFFFFFFull) + __popc(crm & lt_mask);
                        const int off = (int) (prevp & 0xFFFFFFull) + xa - nn;
                        const u32 rev = ((u32) enc >> 30) & 1u;
                        const int gxv = one_group ? gx0 : sm.gx[g], gyv = one_group ? gy0 : sm.gy[g];
                        const int xx = dir ? gyv : gxv, yy = dir ? gxv : gyv;
                        const int hpa = (int) (key >> 32), hpb = (int) (key & 0xFFFFFFFFu);
                        ST_OUT(t_ida + idx, hpa); ST_OUT(t_idb + idx, hpb);
                        ST_OUT(t_ab + idx, (unsigned char) ((rev ? yy : xx) | ((rev ? xx : yy) << 4)));
                        if (estep) ST_OUT(t_tp + idx, tv); else ST_OUT(t_aoff + idx, off);
                        if (out_cache) { co_ida[idx] = hpa; co_idb[idx] = hpb; }
#ifndef FWD_NO_PREFETCH
                        // the next step gathers the successor rows of this state: start moving them to L2 now (a row of a biallelic model is one sector)
#if FWD_PREFETCH_L1
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(succ16 + (size_t) hpa * A));
                        asm volatile("prefetch.global.L1 [%0];" :: "l"(succ16 + (size_t) hpb * A));
#else
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(succ16 + (size_t) hpa * A));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(succ16 + (size_t) hpb * A));
#endif
#endif
                        if (nn == 1) {
                            const double alpha = __dmul_rn(pa, tv);
                            ST_OUT(t_alpha + idx, alpha);
                            if (out_cache) co_alpha[idx] = alpha;
                            amax = fmax(amax, alpha);
                            if (!estep) {
                                const double v = __dmul_rn(pv, tv);
                                ST_OUT(t_v1 + idx, v);
                                if (out_cache) co_v1[idx] = v;
                                ST_OUT(t_arr + off, (u32) s | (gxv != gyv ? ARR_HET : 0u) | (rev ? ARR_REV : 0u));
                            }
                            else fwd[c] = (int) ((u32) idx | (rev << 31));
                            tab.release(h);
                            cand[c] = -1;
                        }
                        else {
                            tab.created(h, (u32) idx, (u32) off);     // ord = state index (never equals an ordinal), cnt = 0 (the scatter counts again), soff
                            if (nn <= 8) mlist[atomicAdd(&cn[2], 1)] = make_int2(idx, (int) h);
                            else mlist[mcap - 1 - atomicAdd(&cn[3], 1)] = make_int2(idx, (int) h);
                        }
                    }
                }
                block_sync<NT>();
                PASS_MARK(2);
                const int n_light = cn[2], n_heavy = cn[3];

                // ---- pass D: scatter the arrivals of multi-arrival states (unordered inside a segment; ordered in E) ----
                if (n_light + n_heavy > 0) {
#pragma unroll 1
                    for (int c = tid; c < C; c += NT) {
                        const int enc = cand[c];
                        if (enc >= 0) {
                            const u32 h = (u32) enc & SLOT_MASK, rev = ((u32) enc >> 30) & 1u;
                            const u32 so = tab.soffs[h];
                            const u32 pos = atomicAdd(tab.cnt + h, 1u);
                            sarr[so + pos] = (int) ((u32) c | (rev << 31));
                            if (estep) fwd[c] = (int) ((tab.ord[h] & 0x7FFFFFFFu) | (rev << 31));
                        }
                    }
                    block_sync<NT>();
                }
                PASS_MARK(3);

                // ---- pass E (light): one lane per state with 2..8 arrivals: ordinals ordered in registers, then taken one by one ----
#pragma unroll 1
                for (int m = tid; m < n_light; m += NT) {
                    const int2 me = mlist[m];
                    const int idx = me.x;
                    const u32 h = (u32) me.y;
                    const double tv = tab.tpvs[h];
                    const int a0 = (int) tab.soffs[h], n = (int) tab.cnt[h];
                    tab.release(h);
                    int v[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = (q < n) ? sarr[a0 + q] : 0x7FFFFFFF;
                    auto cx = [&](int a, int b) { const int x = v[a], y = v[b]; const bool sw = (x & 0x7FFFFFFF) > (y & 0x7FFFFFFF); v[a] = sw ? y : x; v[b] = sw ? x : y; };
                    if (n > 2) {        // 19-comparator network for 8 (padding sorts to the end)
                        cx(0, 1); cx(2, 3); cx(4, 5); cx(6, 7);
                        cx(0, 2); cx(1, 3); cx(4, 6); cx(5, 7);
                        cx(1, 2); cx(5, 6); cx(0, 4); cx(3, 7);
                        cx(1, 5); cx(2, 6);
                        cx(1, 4); cx(3, 6);
                        cx(2, 4); cx(3, 5);
                        cx(3, 4);
                    }
                    else cx(0, 1);
                    // (fetching all predecessors' alpha / v1 in one trip before the ordered sum, or one arrival ahead, was measured: 1 ms SLOWER -- the
                    // kernel is bound by the instructions a warp issues per step more than by these trips)
                    double alpha = 0.0, vmax = 0.0;
#pragma unroll 1
                    for (int q = 0; q < n; ++q) {
                        const int e = v[0];
#pragma unroll
                        for (int r = 0; r < 7; ++r) v[r] = v[r + 1];       // (static register indexing: the list slides down)
                        int g, s, dir;
                        decode(e & 0x7FFFFFFF, g, s, dir);
                        const double term = __dmul_rn(in_alpha[s], tv);
                        alpha = (q == 0) ? term : __dadd_rn(alpha, term);
                        if (!estep) {
                            vmax = fmax(vmax, in_v1[s]);
                            ST_OUT(t_arr + a0 + q, (u32) s | ((one_group ? het0 : sm.gx[g] != sm.gy[g]) ? ARR_HET : 0u) | (((u32) e >> 31) ? ARR_REV : 0u));
                        }
                    }
                    ST_OUT(t_alpha + idx, alpha);
                    if (out_cache) co_alpha[idx] = alpha;
                    amax = fmax(amax, alpha);
                    if (!estep) {
                        const double vv = __dmul_rn(vmax, tv);       // = max over arrivals of v1(pred) * tp: rounding is monotone
                        ST_OUT(t_v1 + idx, vv);
                        if (out_cache) co_v1[idx] = vv;
                    }
                }
                // ---- pass E (heavy): one warp per state with more than 8 arrivals (about 1 % of the states); rank by counting, ordered copy in `cand` (free after D) ----
#pragma unroll 1
                for (int m = wid; m < n_heavy; m += NW) {
                    const int2 me = mlist[mcap - 1 - m];
                    const int idx = me.x;
                    const u32 h = (u32) me.y;
                    const double tv = tab.tpvs[h];
                    const int a0 = (int) tab.soffs[h], n = (int) tab.cnt[h];
                    __syncwarp();
                    if (lane == 0) tab.release(h);
#pragma unroll 1
                    for (int b0 = 0; b0 < n; b0 += 32) {
                        const int mine = (b0 + lane < n) ? sarr[a0 + b0 + lane] : 0x7FFFFFFF;
                        int rank = 0;
#pragma unroll 1
                        for (int q0 = 0; q0 < n; q0 += 32) {
                            const int other = (q0 + lane < n) ? sarr[a0 + q0 + lane] : 0x7FFFFFFF;
                            const int lim = min(32, n - q0);
#pragma unroll 1
                            for (int j = 0; j < lim; ++j) rank += ((__shfl_sync(FULL, other, j) & 0x7FFFFFFF) < (mine & 0x7FFFFFFF)) ? 1 : 0;
                        }
                        if (b0 + lane < n) cand[a0 + rank] = mine;
                    }
                    __syncwarp();
                    double alpha = 0.0, vmax = 0.0;
#pragma unroll 1
                    for (int b0 = 0; b0 < n; b0 += 32) {
                        double term = 0.0;
                        if (b0 + lane < n) {
                            const int e = cand[a0 + b0 + lane];
                            int g, s, dir;
                            decode(e & 0x7FFFFFFF, g, s, dir);
                            term = __dmul_rn(in_alpha[s], tv);
                            if (!estep) {
                                vmax = fmax(vmax, in_v1[s]);
                                ST_OUT(t_arr + a0 + b0 + lane, (u32) s | ((one_group ? het0 : sm.gx[g] != sm.gy[g]) ? ARR_HET : 0u) | (((u32) e >> 31) ? ARR_REV : 0u));
                            }
                        }
                        const int lim = min(32, n - b0);
#pragma unroll 1
                        for (int j = 0; j < lim; ++j) {
                            const double tj = __shfl_sync(FULL, term, j);
                            alpha = (b0 + j == 0) ? tj : __dadd_rn(alpha, tj);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_xor_sync(FULL, vmax, o));
                    if (lane == 0) {
                        ST_OUT(t_alpha + idx, alpha);
                        if (out_cache) co_alpha[idx] = alpha;
                        if (!estep) {
                            const double vv = __dmul_rn(vmax, tv);
                            ST_OUT(t_v1 + idx, vv);
                            if (out_cache) co_v1[idx] = vv;
                        }
                    }
                    amax = fmax(amax, alpha);
                }
                PASS_MARK(4);
                if (p.scaled) {     // exponent-only rescale: mantissas stay those of the unscaled product
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(FULL, amax, o));
                    if (NT > 32) {
                        __syncthreads();
                        if (lane == 0) sm.red[wid] = amax;
                        __syncthreads();
                        amax = 0.0;
#pragma unroll 1
                        for (int q = 0; q < NW; ++q) amax = fmax(amax, sm.red[q]);
                    }
                    if (amax > 0.0 && amax < 0x1p-512) {
                        block_sync<NT>();
#pragma unroll 1
                        for (int idx = tid; idx < S_out; idx += NT) {
                            const double a2 = t_alpha[idx] * 0x1p+512;
                            t_alpha[idx] = a2;
                            if (out_cache) co_alpha[idx] = a2;
                            if (!estep) {
                                const double v2 = t_v1[idx] * 0x1p+512;
                                t_v1[idx] = v2;
                                if (out_cache) co_v1[idx] = v2;
                            }
                        }
                        new_exp += 512;
                    }
                }
            } while (false);
            if (rc == 1) { S_in = 0; break; }
            if (rc == 2) break;
            if (tid == 0) {
                rec[i + 1].off = new_off; rec[i + 1].count = S_out;
                if (p.locus_exp) p.locus_exp[(size_t) w * exp_stride(L) + i + 1] = new_exp;
            }
            cur_exp = new_exp;
            ++c_steps; c_states += S_out; if (S_out > c_max) c_max = S_out;
            S_in = S_out; par ^= 1; cur_off = new_off;
            block_sync<NT>();
            PASS_MARK(5);
#ifdef HMC_PROFILE_PASSES
            hist_cyc[hist_b] += clock64() - t_step; ++hist_n[hist_b];
#endif
        }
        block_sync<NT>();
        if (tid == 0) {
            p.status[w] = status;
            p.final_exp[w] = cur_exp;
            if (p.ind_cycles) {     // diagnostics: cycles, states, max states, arrivals, spill steps, start / end time (ns), CTA
                long long *d = p.ind_cycles + (size_t) w * 8;
                d[0] = clock64() - t_begin; d[1] = c_states; d[2] = c_max; d[3] = c_arr; d[4] = c_spill; d[5] = t_begin_ns;
                unsigned long long t_end_ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end_ns)); d[6] = (long long) t_end_ns; d[7] = blockIdx.x;
            }
            if (status == ST_DONE) {
                if (p.work) p.work[ind] = clock64() - t_begin;      // (SM cycles: a better predictor of the next pass's time than the state count)
                atomicAdd((unsigned long long *) p.counters + 0, (unsigned long long) c_steps);
                atomicAdd((unsigned long long *) p.counters + 1, (unsigned long long) c_states);
                atomicMax((long long *) p.counters + 2, c_max);
                atomicAdd((unsigned long long *) p.counters + 3, (unsigned long long) c_arr);
                atomicAdd((unsigned long long *) p.counters + 5, (unsigned long long) c_spill);
#ifdef HMC_PROFILE_PASSES
                for (int q = 0; q < 16; ++q) atomicAdd((unsigned long long *) p.counters + 8 + q, (unsigned long long) pass_cyc[q]);
                for (int q = 0; q < 18; ++q) { atomicAdd((unsigned long long *) p.counters + 32 + q, (unsigned long long) hist_cyc[q]); atomicAdd((unsigned long long *) p.counters + 50 + q, (unsigned long long) hist_n[q]); }
#endif
            }
            publish(w);
        }
    }
}

}  // namespace hmc
