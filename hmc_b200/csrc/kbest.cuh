// K2 kbest_traceback: the K-best link lists of HaploPair (HaploPair.cpp:35-89), the final selection of
// HaploBuilder::resolve (HaploBuilder.cpp:87-116) and HaploPair::getGenotype (HaploPair.cpp:91-124),
// evaluated LAZILY from the end of the trellis that the forward kernel left (per state: ids, alpha, v1 =
// likelihood of its best link, list of arrivals).
//
// The reference keeps, for every state t, the K best links over its arrivals e:
//     list(t) = top-K of  U_e { g_e(list(pred_e)[j]) : j < K },     g_e(l) = l.likelihood * tp(t)
// (HaploPair.cpp:45-51,68-75,85-88; nth_element after every arrival = top-K of the union) and only the K
// best links over the final states are ever read.  Because g_e is monotone, list(t)[k] can be produced on
// demand by merging the arrival streams g_e(list(pred_e)[0..]) -- each stream is already sorted -- and a
// stream only needs list(pred_e)[j+1] after its j-th element was taken (Huang & Chiang's lazy k-best
// extraction on the trellis).  Every value is the same chain of products from the head state forward that
// the reference computes, so likelihoods are bit-identical; which of several equal-valued links is kept is
// left to libstdc++'s nth_element by the reference and to (stream order, index) here.
// Work: about K states per index instead of all of them times K.
//
// The one non-monotone rule -- at a heterozygous step links still flagged `homozygous` lose the flag and
// are zeroed when the arrival is reversed (HaploPair.cpp:53-60,77-84), and the final x2 for non-homozygous
// links (HaploBuilder.cpp:98) -- only concerns links whose whole path was homozygous, i.e. the indices up
// to and including the first typed heterozygous locus of the individual (every arrival of that step clears
// the flag).  For that short prefix the full lists of all states are built eagerly, exactly like the
// reference does; the lazy enumeration bottoms out on them.
//
// One warp per individual (persistent warps, work queue).  The control flow is warp-uniform: all lanes
// read the same node fields (broadcast loads), lane 0 stores; lanes split the work where there is some
// (arrival streams of a new node, arg-max over stream heads, eager lists, output rows).
#pragma once

#include "common.cuh"

namespace hmc {

struct KbParams {
    DevPanel panel;
    DevModel model;
    u64 *arena;
    long long capacity;             // units of 8 bytes
    unsigned long long *cursor;     // bump allocator: the forward kernel's own (serial mode: it has finished and the nodes go behind the trellis), or,
                                    // streamed, a second one over [region_begin, capacity) -- the two kernels give failed grabs back, which is only
                                    // sound among allocators with the same limit
    long long region_begin;
    const LocusRec *loci;           // [n_wave][L+1]
    const int *wave;
    int *status;
    const int *final_exp;
    const int *locus_exp;           // [n_wave][L+1] cumulative exponents (scaled mode) or null
    int n_wave, K;
    int first;                      // panel index of result row 0
    u64 *eoff;                      // [warps][L+2] eager list block of an index (indices <= base)
    u32 *chain;                     // [warps][L+2] head of the node chain of an index (indices >= base; L+1 = the root)
    u32 *stack;                     // [warps][stack_stride]: node stack of the enumeration (depth <= L + 2); result starts of the all-eager case (K)
    int stack_stride;
    int *queue;
    const int *done_list;           // streamed mode: wave positions in the order the forward kernel (running beside this one) finishes them, -1 = not yet; or null
    unsigned long long *work;       // [0] lazy-next calls, [1] nodes created, [2] eager list states
    // results, indexed by (panel individual - first)
    double *pg, *cov, *prior, *post;
    int *exp2, *nres;
    signed char *haps;              // [count][K][2][L] or null
};

constexpr int KB_WARPS = 4;
#ifndef KB1_MAXNREG
#define KB1_MAXNREG 80      // the streamed one-warp variant: see FWD_MAXNREG (forward.cuh)
#endif
constexpr u32 KB_NONE = 0xFFu;

// Node of the lazy enumeration: one visited (index, state).  Byte layout, K-dependent (see kb_node_bytes):
//   +0  int level, int state
//   +8  u32 next (chain of the level), u8 nD, u8 nS, u8 pending stream, u8 exhausted
//   +16 double tpv (transition probability of the state = factor of every stream)
//   +24 double fac (2^(exponent step) of the index in scaled mode, else 1)
//   +32 double head[K]     next candidate of every stream, -1 = exhausted
//   +.. double dv[K]       materialised links, descending
//   +.. u32 sword[K]       arrival word of the stream (predecessor state | reversed)
//   +.. u32 snode[K]       node of the predecessor (0 = not looked up yet)
//   +.. u32 dbp[K]         back-pointer word of the link: predecessor state | index << 24 | reversed (eager format at the base index)
//   +.. u32 dnode[K]       node of the link's predecessor (0 = unknown)
//   +.. u8  sj[K]          index of the stream's next candidate in the predecessor's list
__host__ __device__ inline int kb_node_bytes(int K) { return (32 + 33 * K + 63) & ~63; }

struct KbNode {
    unsigned char *b;
    int K;
    __device__ int &level() const { return *(int *) b; }
    __device__ int &state() const { return *(int *) (b + 4); }
    __device__ u32 &next() const { return *(u32 *) (b + 8); }
    __device__ unsigned char &nD() const { return b[12]; }
    __device__ unsigned char &nS() const { return b[13]; }
    __device__ unsigned char &pending() const { return b[14]; }
    __device__ unsigned char &exhausted() const { return b[15]; }
    __device__ double &tpv() const { return *(double *) (b + 16); }
    __device__ double &fac() const { return *(double *) (b + 24); }
    __device__ double *head() const { return (double *) (b + 32); }
    __device__ double *dv() const { return (double *) (b + 32 + 8 * K); }
    __device__ u32 *sword() const { return (u32 *) (b + 32 + 16 * K); }
    __device__ u32 *snode() const { return (u32 *) (b + 32 + 20 * K); }
    __device__ u32 *dbp() const { return (u32 *) (b + 32 + 24 * K); }
    __device__ u32 *dnode() const { return (u32 *) (b + 32 + 28 * K); }
    __device__ unsigned char *sj() const { return b + 32 + 32 * K; }
};

// Eager list block of one index (all states): len[S] bytes | val[S][K] | bp[S][K]
struct EagerLayout { long long len, val, bp, total; };
__host__ __device__ inline EagerLayout eager_layout(int S, int K)
{
    EagerLayout e;
    e.len = 0;
    e.val = (S + 7) / 8;
    e.bp = e.val + (long long) S * K;
    e.total = e.bp + ((long long) S * K + 1) / 2;
    return e;
}

// WARPS = 4: after the forward kernel, every warp slot of the GPU.  WARPS = 1 (streamed mode): resident BESIDE the forward kernel in the
// registers its three CTAs per SM leave over (4,096: two warps at 64 registers), fed through done_list as individuals finish.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) __maxnreg__(WARPS == 1 ? KB1_MAXNREG : 80) kbest_kernel(KbParams p)
{
    constexpr u32 FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const DevPanel &pn = p.panel;
    const DevModel &md = p.model;
    const int L = pn.L, K = p.K, head_len = md.head_len;
    u64 *const eoff = p.eoff + (size_t) gw * (L + 2);
    u32 *const chain = p.chain + (size_t) gw * (L + 2);
    u32 *const stack = p.stack + (size_t) gw * p.stack_stride;
    const int node_bytes = kb_node_bytes(K);
    unsigned long long n_next = 0, n_nodes = 0, n_eager = 0;

    // warp-level chunked bump allocation (64-byte aligned); returns -1 when the arena is exhausted
    long long ch_cur = 0, ch_end = 0;
    auto walloc = [&](long long units) -> long long {
        units = (units + 7) & ~7LL;
        if (ch_cur + units > ch_end) {
            constexpr long long CHUNK = 1 << 11;        // 16 KB
            const long long grab = (units > CHUNK ? ((units + 15) & ~15LL) : CHUNK) + 16;   // whole 128-byte lines (see loci_stride)
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(p.cursor, (unsigned long long) grab) + (unsigned long long) p.region_begin;
            base = __shfl_sync(FULL, base, 0);
            if ((long long) (base + grab) > p.capacity) {
                if (lane == 0) atomicAdd(p.cursor, (unsigned long long) (-grab));
                return -1;
            }
            ch_cur = ((long long) base + 7) & ~7LL; ch_end = (long long) base + grab;
        }
        const long long off = ch_cur;
        ch_cur += units;
        return off;
    };
    auto node_at = [&](u32 ref) -> KbNode { KbNode n; n.b = (unsigned char *) (p.arena + (size_t) ref * 8); n.K = K; return n; };

    // the `want` best of n candidates in (value descending, index ascending) order, warp-cooperatively; value_of(e) must not
    // be negative for a real candidate; calls emit(rank, e, value) (warp-uniform arguments)
    auto select_top = [&](int n, int want, auto value_of, auto emit) {
        double last_v = 0.0;
        int last_e = -1;
        for (int r = 0; r < want; ++r) {
            double bv = -1.0;
            int be = 0x7FFFFFFF;
            for (int e = lane; e < n; e += 32) {
                const double v = value_of(e);
                const bool after = (r == 0) || v < last_v || (v == last_v && e > last_e);
                if (after && v >= 0.0 && (v > bv || (v == bv && e < be))) { bv = v; be = e; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(FULL, bv, o);
                const int oe = __shfl_xor_sync(FULL, be, o);
                if (ov > bv || (ov == bv && oe < be)) { bv = ov; be = oe; }
            }
            if (be == 0x7FFFFFFF) break;
            emit(r, be, bv);
            last_v = bv; last_e = be;
        }
    };

    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(p.queue, 1);
        w = __shfl_sync(FULL, w, 0);
        if (w >= p.n_wave) break;
        if (p.done_list) {      // ticket w -> the w-th individual the forward kernel finishes
            if (lane == 0) {
                const volatile int *slot = p.done_list + w;
                while ((w = *slot) < 0) __nanosleep(1000);
            }
            w = __shfl_sync(FULL, w, 0);
            __threadfence();        // every lane reads the trellis: each one orders its loads behind the list entry
        }
        if (__ldcg(p.status + w) != ST_DONE) continue;
#ifdef HMC_PROFILE_PASSES
        const long long kt0 = clock64();
#endif
        const int row = p.wave[w] - p.first;
        const LocusRec *rec = p.loci + (size_t) w * loci_stride(L);
        const int *lexp = p.locus_exp ? p.locus_exp + (size_t) w * exp_stride(L) : nullptr;
        const int S_fin = rec[L].count;
        if (S_fin <= 0) {       // unresolvable: HaploBuilder.cpp:117-124
            if (lane == 0) { p.pg[row] = 0.0; p.cov[row] = 0.0; p.nres[row] = 0; p.exp2[row] = 0; }
            for (int r = lane; r < K; r += 32) { p.prior[(size_t) row * K + r] = 0.0; p.post[(size_t) row * K + r] = 0.0; }
            continue;
        }
        const u32 *gp = pn.packed + (size_t) p.wave[w] * pn.words;
        const int gbits = pn.bits;
        const u32 gmask = (1u << gbits) - 1u;

        // trellis views of an index
        auto blk_of = [&](int l) -> const u64 * { return p.arena + rec[l].off; };
        auto lay_of = [&](int l) -> BlockLayout { return block_layout(rec[l].count, 0, 0); };
        auto fac_of = [&](int l) -> double {        // exponent step applied at index l (forward kernel, scaled mode)
            if (!lexp || l <= head_len || l > L) return 1.0;
            return (lexp[l] != lexp[l - 1]) ? 0x1p+512 : 1.0;
        };

        // ---- P(G) = sum of alpha over the final states in creation order (serial, bit-identical) ----
        double total = 0.0;
        {
            const BlockLayout b = lay_of(L);
            const double *f_alpha = (const double *) (blk_of(L) + b.alpha);
            for (int s0 = 0; s0 < S_fin; s0 += 32) {
                const double v = (s0 + lane < S_fin) ? f_alpha[s0 + lane] : 0.0;
                const int m = min(32, S_fin - s0);
                for (int j = 0; j < m; ++j) total = __dadd_rn(total, __shfl_sync(FULL, v, j));
            }
        }

        // ---- base index: one past the first typed heterozygous locus (all homozygous flags are gone there) ----
        int base = L;
        for (int l0 = head_len; l0 < L; l0 += 32) {
            const int l = l0 + lane;
            bool het = false;
            if (l < L) {
                const int bit = l * 2 * gbits;
                const u32 wd = gp[bit >> 5] >> (bit & 31);
                const int g0 = (int) (wd & gmask) - 1, g1 = (int) ((wd >> gbits) & gmask) - 1;
                het = g0 >= 0 && g1 >= 0 && g0 != g1;
            }
            const u32 bal = __ballot_sync(FULL, het);
            if (bal) { base = l0 + __ffs(bal); break; }     // index of the locus + 1
        }
        if (base > L) base = L;
        bool nomem = false;

        // ---- eager lists for the indices head_len .. base ----
        for (int l = head_len; l <= base; ++l) {
            const int S = rec[l].count;
            const EagerLayout el = eager_layout(S, K);
            const long long eo = walloc(el.total);
            if (eo < 0) { nomem = true; break; }
            if (lane == 0) eoff[l] = (u64) eo;
            unsigned char *e_len = (unsigned char *) (p.arena + eo + el.len);
            double *e_val = (double *) (p.arena + eo + el.val);
            u32 *e_bp = (u32 *) (p.arena + eo + el.bp);
            const BlockLayout b = lay_of(l);
            const u64 *blk = blk_of(l);
            const int *ida = (const int *) (blk + b.ida), *idb = (const int *) (blk + b.idb);
            n_eager += (unsigned long long) S;
            if (l == head_len) {        // head link {null, 0, false, homozygous = (id_a == id_b), f_a * f_b}: HaploPair.cpp:26-32
                const double *v1 = (const double *) (blk + b.v1);
                for (int t = lane; t < S; t += 32) {
                    e_len[t] = 1; e_val[(size_t) t * K] = v1[t];
                    e_bp[(size_t) t * K] = BP_HEAD | (ida[t] == idb[t] ? BP_HOM : 0u);
                }
            }
            else {
                const int Sp = rec[l - 1].count;
                const EagerLayout pl = eager_layout(Sp, K);
                const long long po = (long long) eoff[l - 1];
                const unsigned char *p_len = (const unsigned char *) (p.arena + po + pl.len);
                const double *p_val = (const double *) (p.arena + po + pl.val);
                const u32 *p_bp = (const u32 *) (p.arena + po + pl.bp);
                const int *aoff = (const int *) (blk + b.aoff);
                const u32 *arr = (const u32 *) (blk + b.arr);
                const double fac = fac_of(l);
                for (int t = lane; t < S; t += 32) {
                    const double tpv = __dmul_rn(md.tp[ida[t]], md.tp[idb[t]]);
                    double *ov = e_val + (size_t) t * K;
                    u32 *ob = e_bp + (size_t) t * K;
                    int n = 0;
                    for (int a = aoff[t]; a < aoff[t + 1]; ++a) {
                        const u32 wd = arr[a];
                        const int s = (int) (wd & ARR_PRED_MASK);
                        const bool het = (wd & ARR_HET) != 0, rev = (wd & ARR_REV) != 0;
                        const int nl = p_len[s];
                        for (int j = 0; j < nl; ++j) {
                            double v = __dmul_rn(p_val[(size_t) s * K + j], tpv);
                            if (fac != 1.0) v *= fac;
                            bool hom = (p_bp[(size_t) s * K + j] & BP_HOM) != 0;
                            if (het) { if (hom && rev) v = 0.0; hom = false; }      // HaploPair.cpp:53-60,77-84
                            const u32 bw = (u32) s | ((u32) j << BP_PRED_BITS) | (rev ? BP_REV : 0u) | (hom ? BP_HOM : 0u);
                            // sorted insert, later equal values go behind earlier ones; the list is capped at K
                            int q = n;
                            if (n == K) { if (!(v > ov[K - 1])) continue; q = K - 1; }
                            while (q > 0 && v > ov[q - 1]) { ov[q] = ov[q - 1]; ob[q] = ob[q - 1]; --q; }
                            ov[q] = v; ob[q] = bw;
                            if (n < K) ++n;
                        }
                    }
                    e_len[t] = (unsigned char) n;
                }
            }
            __syncwarp();
        }
        if (nomem) { if (lane == 0) p.status[w] = ST_NOMEM_KBEST; continue; }

#ifdef HMC_PROFILE_PASSES
        const long long kt1 = clock64();
#endif
        // ---- the K best links over the final states ----
        double my_prior = 0.0;      // lane r: prior of result r
        int n_sel = 0;
        // start of path r (warp-uniform after the selection): lazy -> node + index, eager -> state + index
        u32 root = 0;

        // lazy machinery -----------------------------------------------------------------------------------
        auto new_node = [&](int level, int state) -> u32 {
            const long long off = walloc(node_bytes / 8);
            if (off < 0) { nomem = true; return 0; }
            const u32 ref = (u32) (off >> 3);
            KbNode nd = node_at(ref);
            if (lane == 0) {
                nd.level() = level; nd.state() = state; nd.next() = chain[level]; nd.nD() = 0; nd.nS() = 0; nd.pending() = KB_NONE; nd.exhausted() = 0;
                nd.tpv() = 1.0; nd.fac() = 1.0;
                chain[level] = ref;
            }
            ++n_nodes;
            return ref;
        };
        // streams of a new node: the (up to K) arrivals with the best heads v1(pred) * tpv (* fac)
        auto init_streams = [&](KbNode nd, int n, double tpv, double fac, const double *pv1, auto word_of) {
            int ns = 0;
            double h0 = -1.0;       // best head and its arrival word (stream 0), kept for the first link below
            u32 sw0 = 0;
            auto value_of = [&](int e) -> double {
                double v = __dmul_rn(pv1[word_of(e) & ARR_PRED_MASK], tpv);
                if (fac != 1.0) v *= fac;
                return v;
            };
            if (n == 1) {           // a single arrival (two thirds of the states): one stream, nothing to select
                const double v = value_of(0);
                if (v >= 0.0) {
                    sw0 = word_of(0); h0 = v; ns = 1;
                    if (lane == 0) { nd.head()[0] = v; nd.sword()[0] = sw0; nd.snode()[0] = 0; nd.sj()[0] = 0; }
                }
            }
            else
                select_top(n, K, value_of,
                           [&](int r, int e, double v) {
                               const u32 wd = word_of(e);
                               if (lane == 0) { nd.head()[r] = v; nd.sword()[r] = wd; nd.snode()[r] = 0; nd.sj()[r] = 0; }
                               if (r == 0) { h0 = v; sw0 = wd; }
                               ns = r + 1;
                           });
            if (lane == 0) {
                nd.tpv() = tpv; nd.fac() = fac; nd.nS() = (unsigned char) ns;
                if (ns > 0) {       // the first link is the best head: stream 0 (select_top emits best first)
                    const u32 sw = sw0;
                    nd.dv()[0] = h0;
                    nd.dbp()[0] = (sw & ARR_PRED_MASK) | ((sw & ARR_REV) ? BP_REV : 0u);
                    nd.dnode()[0] = 0;
                    nd.sj()[0] = 1;
                    nd.pending() = 0;
                    nd.nD() = 1;
                }
                else nd.exhausted() = 1;
            }
        };
        auto find_or_create = [&](int level, int state) -> u32 {
            u32 ref = chain[level];
            while (ref) { KbNode nd = node_at(ref); if (nd.state() == state) return ref; ref = nd.next(); }
            ref = new_node(level, state);
            if (!ref) return 0;
            KbNode nd = node_at(ref);
            if (level == base) {        // the eager list of the state, as it stands
                const EagerLayout el = eager_layout(rec[level].count, K);
                const long long eo = (long long) eoff[level];
                const int nl = ((const unsigned char *) (p.arena + eo + el.len))[state];
                const double *e_val = (const double *) (p.arena + eo + el.val) + (size_t) state * K;
                const u32 *e_bp = (const u32 *) (p.arena + eo + el.bp) + (size_t) state * K;
                if (lane < nl) { nd.dv()[lane] = e_val[lane]; nd.dbp()[lane] = e_bp[lane]; nd.dnode()[lane] = 0; }
                if (lane == 0) { nd.nD() = (unsigned char) nl; nd.exhausted() = 1; }
            }
            else {
                const BlockLayout b = lay_of(level);
                const u64 *blk = blk_of(level);
                const int ia = ((const int *) (blk + b.ida))[state], ib = ((const int *) (blk + b.idb))[state];
                const double tpv = __dmul_rn(md.tp[ia], md.tp[ib]);
                const int *aoff = (const int *) (blk + b.aoff);
                const int a0 = aoff[state], a1 = aoff[state + 1];
                const u32 *arr = (const u32 *) (blk + b.arr) + a0;
                const double *pv1 = (const double *) (blk_of(level - 1) + lay_of(level - 1).v1);
                init_streams(nd, a1 - a0, tpv, fac_of(level), pv1, [&](int e) -> u32 { return arr[e]; });
            }
            __syncwarp();
            return ref;
        };
        // appends one more link to the list of node `top` (or marks it exhausted); nodes below it are advanced as needed.
        // The chain of visits is a chain of dependent L2 round trips, so a visit is kept to two: (1) the 32-byte header and, with the lanes
        // spread over the streams (lane e holds head / arrival word / predecessor node / next index of stream e), the stream arrays -- issued
        // together, lanes past nS read inside the node and are masked afterwards; (2) the predecessor's header and the link the pending
        // stream wants.  A visit that comes back UP from a predecessor it had to advance takes that link from registers instead of (2).
        // Updates go back as single stores, nothing is re-read.
        auto lazy_next = [&](u32 top) {
            int sp = 0;
            u32 cur = top;
            bool up = false;            // the previous visit was of the pending stream's predecessor, which ...
            double up_v = -1.0;         // ... appended this link (< 0: it ran dry instead)
            // a visit that descends has already read its predecessor whole (header + streams, in the trip that fetched the wanted link): carried over
            bool have = false;
            uint4 c_hdr = make_uint4(0, 0, 0, 0);
            double2 c_tf = make_double2(0.0, 0.0);
            double c_head = -1.0;
            u32 c_sw = 0, c_sn = 0;
            int c_j = 0;
            for (;;) {
                __syncwarp();
                KbNode nd = node_at(cur);
                uint4 hdr;
                double2 tf;
                double my_head = -1.0;
                u32 my_sw = 0, my_sn = 0;
                int my_j = 0;
                if (have) { hdr = c_hdr; tf = c_tf; my_head = c_head; my_sw = c_sw; my_sn = c_sn; my_j = c_j; have = false; }
                else {
                    hdr = *reinterpret_cast<const uint4 *>(nd.b);           // level, state, next, {nD, nS, pending, exhausted}
                    tf = *reinterpret_cast<const double2 *>(nd.b + 16);     // tpv, fac
                    if (lane < K) { my_head = nd.head()[lane]; my_sw = nd.sword()[lane]; my_sn = nd.snode()[lane]; my_j = nd.sj()[lane]; }
                }
                const int ns = (int) ((hdr.w >> 8) & 0xFFu), pend = (int) ((hdr.w >> 16) & 0xFFu);
                const int k = (int) (hdr.w & 0xFFu);
                if (lane >= ns) { my_head = -1.0; my_sw = 0; my_sn = 0; my_j = 0; }
                if (pend != (int) KB_NONE) {
                    const int j = __shfl_sync(FULL, my_j, pend);
                    double hv = -1.0;
                    if (j < K) {
                        double pv = -1.0;
                        if (up) pv = up_v;      // the predecessor's list grew to index j a moment ago (it was at j: lists grow on demand only)
                        else {
                            u32 pref = __shfl_sync(FULL, my_sn, pend);
                            if (!pref) {
                                pref = find_or_create((int) hdr.x - 1, (int) (__shfl_sync(FULL, my_sw, pend) & ARR_PRED_MASK));
                                if (!pref) return;
                                if (lane == pend) { my_sn = pref; nd.snode()[pend] = pref; }
                                // the link this stream supplied (its first: index k - 1 here) did not know its predecessor's node yet: the traceback follows nodes
                                if (lane == 0 && k > 0) nd.dnode()[k - 1] = pref;
                            }
                            KbNode pd = node_at(pref);
                            c_hdr = *reinterpret_cast<const uint4 *>(pd.b);
                            c_tf = *reinterpret_cast<const double2 *>(pd.b + 16);
                            const double pj = pd.dv()[j];           // (read before it is known to be valid: inside the node)
                            c_head = -1.0; c_sw = 0; c_sn = 0; c_j = 0;
                            if (lane < K) { c_head = pd.head()[lane]; c_sw = pd.sword()[lane]; c_sn = pd.snode()[lane]; c_j = pd.sj()[lane]; }
                            const u32 pw = c_hdr.w;
                            if ((int) (pw & 0xFFu) > j) pv = pj;
                            else if (!(pw >> 24)) {     // the predecessor's list has to grow first
                                if (lane == 0) stack[sp] = cur;
                                ++sp;
                                cur = pref;
                                have = true;
                                continue;
                            }
                        }
                        if (pv >= 0.0) {
                            hv = __dmul_rn(pv, tf.x);
                            if (tf.y != 1.0) hv *= tf.y;
                        }
                    }
                    if (lane == pend) { my_head = hv; nd.head()[pend] = hv; }
                }
                // take the best stream head (ties: the earlier stream); a single stream needs no contest
                double bv = my_head;
                int be = lane;
                if (ns == 1) { bv = __shfl_sync(FULL, my_head, 0); be = 0; }
                else {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double ov = __shfl_xor_sync(FULL, bv, o);
                        const int oe = __shfl_xor_sync(FULL, be, o);
                        if (ov > bv || (ov == bv && oe < be)) { bv = ov; be = oe; }
                    }
                }
                if (bv < 0.0) { if (lane == 0) *reinterpret_cast<u32 *>(nd.b + 12) = (hdr.w & 0x0000FFFFu) | (KB_NONE << 16) | (1u << 24); }
                else {
                    if (lane == be) {
                        nd.dv()[k] = bv;
                        nd.dbp()[k] = (my_sw & ARR_PRED_MASK) | ((u32) my_j << BP_PRED_BITS) | ((my_sw & ARR_REV) ? BP_REV : 0u);
                        nd.dnode()[k] = my_sn;
                        nd.sj()[be] = (unsigned char) (my_j + 1);
                        *reinterpret_cast<u32 *>(nd.b + 12) = (u32) (k + 1) | ((u32) ns << 8) | ((u32) be << 16);       // nD, nS, pending, not exhausted
                    }
                }
                ++n_next;
                if (sp == 0) break;
                --sp;
                up = true; up_v = bv;       // (bv < 0: nothing was appended)
                __syncwarp();
                cur = stack[sp];
            }
            __syncwarp();
        };
        if (base == L) {
            // every index is eager (no typed heterozygous locus before the last index): selection over all final links, x2 unless homozygous
            const EagerLayout el = eager_layout(S_fin, K);
            const long long eo = (long long) eoff[L];
            const unsigned char *e_len = (const unsigned char *) (p.arena + eo + el.len);
            const double *e_val = (const double *) (p.arena + eo + el.val);
            const u32 *e_bp = (const u32 *) (p.arena + eo + el.bp);
            select_top(S_fin * K, K,
                       [&](int e) -> double {
                           const int t = e / K, j = e - t * K;
                           if (j >= e_len[t]) return -2.0;
                           const double v = e_val[e];
                           return (e_bp[e] & BP_HOM) ? v : __dmul_rn(v, 2.0);
                       },
                       [&](int r, int e, double v) {
                           if (lane == r) my_prior = v;
                           if (lane == 0) stack[r] = (u32) e;       // (state, index) of result r, read back below
                           n_sel = r + 1;
                       });
            __syncwarp();
        }
        else {
            for (int l = base + lane; l <= L + 1; l += 32) chain[l] = 0;
            __syncwarp();
            // the root merges the final states' lists (x2: no link is homozygous any more, HaploBuilder.cpp:98)
            root = new_node(L + 1, 0);
            if (root) {
                const double *fv1 = (const double *) (blk_of(L) + lay_of(L).v1);
                init_streams(node_at(root), S_fin, 2.0, 1.0, fv1, [&](int e) -> u32 { return (u32) e; });
                __syncwarp();
                for (int r = 0; r < K && !nomem; ++r) {
                    if (r > 0) lazy_next(root);         // (the first link was taken when the streams were built)
                    KbNode rn = node_at(root);
                    if ((int) rn.nD() <= r) break;
                    n_sel = r + 1;
                }
                __syncwarp();
                if (!nomem && lane < n_sel) my_prior = node_at(root).dv()[lane];
            }
        }
        if (nomem) { if (lane == 0) p.status[w] = ST_NOMEM_KBEST; continue; }

#ifdef HMC_PROFILE_PASSES
        const long long kt2 = clock64();
#endif
        // ---- priors, posteriors, coverage (HaploBuilder.cpp:105-116) ----
        double cov = 0.0;
        for (int r = 0; r < n_sel; ++r) {
            const double pr = __shfl_sync(FULL, my_prior, r);
            const double post = pr / total;
            cov = __dadd_rn(cov, post);
            if (lane == 0) { p.prior[(size_t) row * K + r] = pr; p.post[(size_t) row * K + r] = post; }
        }
        if (lane == 0) {
            p.pg[row] = total; p.cov[row] = cov; p.nres[row] = n_sel; p.exp2[row] = __ldcg(p.final_exp + w);
            for (int r = n_sel; r < K; ++r) { p.prior[(size_t) row * K + r] = 0.0; p.post[(size_t) row * K + r] = 0.0; }
        }
        if (!p.haps) continue;

        // ---- the paths (HaploPair::getGenotype): lane r walks path r.  Links with a node behind them are read from the node; a link
        // whose predecessor was never expanded is the predecessor's best link (index 0), found from the trellis directly with the
        // same (value, arrival order) rule the enumeration uses ----
        __syncwarp();
        if (lane < n_sel) {
            signed char *out = p.haps + ((size_t) row * K + lane) * 2 * L;
            int level = L, state = 0, k = 0;
            int a = 0;          // which output row currently plays pattern_a's role
            if (base < L) {
                KbNode rn = node_at(root);
                u32 bp = rn.dbp()[lane];
                u32 ref = rn.dnode()[lane];
                state = (int) (bp & ARR_PRED_MASK); k = (int) ((bp >> BP_PRED_BITS) & 31u);
                for (; level > base; --level) {
                    const BlockLayout b = lay_of(level);
                    const u64 *blk = blk_of(level);
                    const unsigned char ab = ((const unsigned char *) (blk + b.ab))[state];
                    out[(size_t) a * L + level - 1] = (signed char) (ab & 15);
                    out[(size_t) (a ^ 1) * L + level - 1] = (signed char) (ab >> 4);
                    if (ref) {
                        KbNode nd = node_at(ref);
                        bp = nd.dbp()[k];
                        ref = nd.dnode()[k];
                    }
                    else {      // best link of (level, state)
                        const int ia = ((const int *) (blk + b.ida))[state], ib = ((const int *) (blk + b.idb))[state];
                        const double tpv = __dmul_rn(md.tp[ia], md.tp[ib]);
                        const double fac = fac_of(level);
                        const int *aoff = (const int *) (blk + b.aoff);
                        const int a0 = aoff[state], a1 = aoff[state + 1];
                        const u32 *arr = (const u32 *) (blk + b.arr);
                        const double *pv1 = (const double *) (blk_of(level - 1) + lay_of(level - 1).v1);
                        double bv = -1.0;
                        u32 bw = 0;
                        for (int e = a0; e < a1; ++e) {
                            const u32 wd = arr[e];
                            double v = __dmul_rn(pv1[wd & ARR_PRED_MASK], tpv);
                            if (fac != 1.0) v *= fac;
                            if (v > bv) { bv = v; bw = wd; }
                        }
                        bp = (bw & ARR_PRED_MASK) | ((bw & ARR_REV) ? BP_REV : 0u);
                    }
                    if (bp & BP_REV) a ^= 1;
                    state = (int) (bp & ARR_PRED_MASK); k = (int) ((bp >> BP_PRED_BITS) & 31u);
                }
            }
            else {
                const int e = (int) stack[lane];
                state = e / K; k = e - state * K;
            }
            // eager indices: back-pointer words of the eager lists down to the head
            for (; level >= head_len; --level) {
                const EagerLayout el = eager_layout(rec[level].count, K);
                const u32 bp = ((const u32 *) (p.arena + (long long) eoff[level] + el.bp))[(size_t) state * K + k];
                const BlockLayout b = lay_of(level);
                const u64 *blk = blk_of(level);
                if (bp & BP_HEAD) {
                    const int ia = ((const int *) (blk + b.ida))[state], ib = ((const int *) (blk + b.idb))[state];
                    const signed char *pa = md.al + md.al_off[ia], *pb = md.al + md.al_off[ib];
                    for (int q = 0; q < level; ++q) { out[(size_t) a * L + q] = pa[q]; out[(size_t) (a ^ 1) * L + q] = pb[q]; }
                    break;
                }
                const unsigned char ab = ((const unsigned char *) (blk + b.ab))[state];
                out[(size_t) a * L + level - 1] = (signed char) (ab & 15);
                out[(size_t) (a ^ 1) * L + level - 1] = (signed char) (ab >> 4);
                if (bp & BP_REV) a ^= 1;
                state = (int) (bp & ARR_PRED_MASK); k = (int) ((bp >> BP_PRED_BITS) & 31u);
            }
        }
        __syncwarp();
#ifdef HMC_PROFILE_PASSES
        if (lane == 0) {    // cycles: P(G) + eager lists, enumeration, traceback; the longest individual
            const long long kt3 = clock64();
            atomicAdd(p.work + 3, (unsigned long long) (kt1 - kt0)); atomicAdd(p.work + 4, (unsigned long long) (kt2 - kt1)); atomicAdd(p.work + 5, (unsigned long long) (kt3 - kt2));
            atomicMax(p.work + 6, (unsigned long long) (kt3 - kt0));
        }
#endif
    }
    if (lane == 0) { atomicAdd(p.work + 0, n_next); atomicAdd(p.work + 1, n_nodes); atomicAdd(p.work + 2, n_eager); }
}

}  // namespace hmc
