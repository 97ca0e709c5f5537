// K3 backward and K4 estep_counts: HaploBuilder::calcBackwardLikelihood (HaploBuilder.cpp:263-272,
// HaploPair.cpp:126-136) and HaploBuilder::estimateFrequency (HaploBuilder.cpp:274-450) on the
// E-step trellis left by the forward kernel (ids, alpha, last alleles, transition probability and the
// forward-link block of every index).
#pragma once

#include "common.cuh"

namespace hmc {

// Flattened ForwardPatternTree (PatternTree.cpp:176-212): one trie per start locus, children by allele index.
struct DevTrie {
    int n_nodes, A, max_depth;
    const int *root;        // [L+1]
    const int *child;       // [n_nodes][A], -1 = none
    const int *data;        // [n_nodes] candidate pattern id or -1
};

struct BackParams {
    DevPanel panel;
    DevModel model;
    u64 *arena;
    const LocusRec *loci;
    const int *wave;
    const int *status;
    int n_wave, scaled;
    int *bexp;              // [n_wave][L+1] cumulative exponent of beta (scaled mode) or null
};

// Slot structure of the step leaving trellis index i (the forward kernels' candidate order):
// groups of allele pairs (x, y); per predecessor state, (x,y) then (y,x) when x != y.
struct StepShape {
    int mode, g0, g1, nA, S;
    u32 am;
    __device__ void init(const DevPanel &pn, const u32 *gp, int i, int S_in)
    {
        const int bit = i * 2 * pn.bits;
        const u32 wd = gp[bit >> 5] >> (bit & 31);
        const u32 m = (1u << pn.bits) - 1u;
        g0 = (int) (wd & m) - 1; g1 = (int) ((wd >> pn.bits) & m) - 1;
        const u32 info = pn.linfo[i];
        nA = (int) (info & 0xFF); am = info >> 8;
        mode = (g0 >= 0 && g1 >= 0) ? 0 : (g0 < 0 && g1 < 0) ? 3 : (g0 < 0 ? 1 : 2);
        S = S_in;
    }
    // calls f(slot, x, y) for every candidate slot of predecessor state s, in ordinal order
    template <typename F>
    __device__ void for_slots(int s, F f) const
    {
        if (mode == 0) {
            if (g0 != g1) { f(2 * s, g0, g1); f(2 * s + 1, g1, g0); }
            else f(s, g0, g0);
            return;
        }
        int base = 0;
        if (mode == 3) {
            for (int j = 0; j < nA; ++j) { if (!((am >> j) & 1u)) continue;
                for (int k = j; k < nA; ++k) { if (!((am >> k) & 1u)) continue;
                    if (j != k) { f(base + 2 * s, j, k); f(base + 2 * s + 1, k, j); base += 2 * S; }
                    else { f(base + s, j, j); base += S; }
                } }
        }
        else {
            const int known = (mode == 1) ? g1 : g0;
            for (int j = 0; j < nA; ++j) { if (!((am >> j) & 1u)) continue;
                if (j != known) { f(base + 2 * s, j, known); f(base + 2 * s + 1, known, j); base += 2 * S; }
                else { f(base + s, j, j); base += S; }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// K3: one warp per individual, indices L-1 .. head_len; lanes over the states of the index.
// beta(s) = sum over forward_links[0] then forward_links[1], each in push order, of beta(t) * tp(t).
__global__ void __launch_bounds__(128) backward_kernel(BackParams p)
{
    const int lane = threadIdx.x & 31;
    const int w = (int) ((blockIdx.x * (size_t) blockDim.x + threadIdx.x) >> 5);
    if (w >= p.n_wave || p.status[w] != ST_DONE) return;
    const DevPanel &pn = p.panel;
    const int L = pn.L, head_len = p.model.head_len;
    const LocusRec *rec = p.loci + (size_t) w * loci_stride(L);
    const u32 *gp = pn.packed + (size_t) p.wave[w] * pn.words;
    int cur_exp = 0;
    {   // index L: beta = 1 (HaploPair.cpp:17,38)
        const int S = rec[L].count;
        const BlockLayout b = block_layout(S, 1, 1);
        double *beta = (double *) (p.arena + rec[L].off + b.beta);
        for (int s = lane; s < S; s += 32) beta[s] = 1.0;
        if (p.bexp && lane == 0) p.bexp[(size_t) w * (L + 1) + L] = 0;
    }
    __syncwarp();
    for (int i = L - 1; i >= head_len; --i) {
        const int S = rec[i].count;
        if (S > 0) {
            const BlockLayout b = block_layout(S, 1, 1);
            double *beta = (double *) (p.arena + rec[i].off + b.beta);
            const int Sn = rec[i + 1].count;
            if (Sn <= 0 || rec[i].fwd_slots <= 0) {
                for (int s = lane; s < S; s += 32) beta[s] = 0.0;
            }
            else {
                const BlockLayout bn = block_layout(Sn, 1, 1);
                const double *beta_n = (const double *) (p.arena + rec[i + 1].off + bn.beta);
                const double *tp_n = (const double *) (p.arena + rec[i + 1].off + bn.tp);
                const int *fwd = (const int *) (p.arena + rec[i].fwd_off);
                StepShape sh;
                sh.init(pn, gp, i, S);
                double bmax = 0.0;
                for (int s = lane; s < S; s += 32) {
                    double acc = 0.0;
                    for (int r = 0; r < 2; ++r) {
                        sh.for_slots(s, [&](int c, int, int) {
                            const int f = fwd[c];
                            if (f != -1 && (int) ((u32) f >> 31) == r) {
                                const int t = f & 0x7FFFFFFF;
                                acc = __dadd_rn(acc, __dmul_rn(beta_n[t], tp_n[t]));
                            }
                        });
                    }
                    beta[s] = acc;
                    bmax = fmax(bmax, acc);
                }
                if (p.scaled) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) bmax = fmax(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
                    if (bmax > 0.0 && bmax < 0x1p-512) {
                        __syncwarp();
                        for (int s = lane; s < S; s += 32) beta[s] *= 0x1p+512;
                        cur_exp += 512;
                    }
                }
            }
        }
        if (p.bexp && lane == 0) p.bexp[(size_t) w * (L + 1) + i] = cur_exp;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
struct EstepScratch {       // per persistent warp
    double *acc;            // [3][smax] dense accumulators, zero between nodes
    int *ekey;              // [ecap] entry stack: state * 4 + class
    double *ew;             // [ecap] entry weights
    int smax, ecap;
};

struct EstepParams {
    DevPanel panel;
    DevModel model;
    DevTrie trie;
    const u64 *arena;
    const LocusRec *loci;
    const int *wave;
    const int *status;
    const int *aexp;        // [n_wave][L+1] forward cumulative exponents or null
    const int *bexp;        // [n_wave][L+1] backward cumulative exponents or null
    const int *final_exp;   // [n_wave]
    const double *pg;       // [count] genotype_probability (scaled by pg_exp) by (panel individual - first)
    const int *pg_exp;      // or null
    int first, n_wave;
    EstepScratch *scratch;  // [warps in the grid]
    double *counts;         // [2][Pc]: frequency sums then prefix sums
    int Pc;
    int *queue;
    int *error;             // set when an entry stack overflowed
    unsigned long long *work;   // parent entries read by the walk (HMC_CNT_ESTEP_ENTRIES)
};

constexpr int ESTEP_MAX_DEPTH = 160;
constexpr int ESTEP_WARPS = 4;
#ifndef ESTEP_MINB
#define ESTEP_MINB 8
#endif

// K4: one warp per (individual, start) task: depth-first walk of the candidate trie below root(start) with the three
// sparse weight vectors of HaploBuilder.cpp:334-450 ([0] both haplotypes match so far, [1] only the a-role one,
// [2] only the b-role one) kept as an entry stack; a node's entries are produced from its parent's by a scatter into
// dense accumulators plus a first-touch list.  Subtrees below a node with no entry are skipped: they only ever add 0.
__global__ void __launch_bounds__(ESTEP_WARPS * 32, ESTEP_MINB) estep_kernel(EstepParams p)
{
    __shared__ int s_node[ESTEP_WARPS][ESTEP_MAX_DEPTH], s_iter[ESTEP_WARPS][ESTEP_MAX_DEPTH];
    __shared__ int s_off[ESTEP_WARPS][ESTEP_MAX_DEPTH + 1];
    __shared__ double s_freq[ESTEP_WARPS][ESTEP_MAX_DEPTH];
    __shared__ int s_cnt[ESTEP_WARPS];
    constexpr u32 FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int gw = blockIdx.x * ESTEP_WARPS + wi;
    const EstepScratch sc = p.scratch[gw];
    const DevPanel &pn = p.panel;
    const DevTrie &tr = p.trie;
    const int L = pn.L, head_len = p.model.head_len, A = tr.A;
    int *node_st = s_node[wi], *iter_st = s_iter[wi], *off_st = s_off[wi];
    double *freq_st = s_freq[wi];
    const long long n_tasks = (long long) p.n_wave * L;
    unsigned long long work = 0;

    for (;;) {
        long long task = 0;
        if (lane == 0) task = atomicAdd((unsigned long long *) p.queue, 1ull);
        task = __shfl_sync(FULL, task, 0);
        if (task >= n_tasks) break;
        const int w = (int) (task / L), start = (int) (task % L);
        if (p.status[w] != ST_DONE) continue;
        const int root = tr.root[start];
        {   // any pattern starting here?
            bool any = false;
            for (int j = lane; j < A; j += 32) any |= tr.child[(size_t) root * A + j] >= 0;
            if (!__any_sync(FULL, any)) continue;
        }
        const LocusRec *rec = p.loci + (size_t) w * loci_stride(L);
        const int row = p.wave[w] - p.first;
        const double pg = p.pg[row];
        const int pge = p.pg_exp ? p.pg_exp[row] : 0;
        const u32 *gp = pn.packed + (size_t) p.wave[w] * pn.words;
        const int idx0 = start > head_len ? start : head_len;
        const int S0 = rec[idx0].count;
        const int ea0 = p.aexp ? p.aexp[(size_t) w * exp_stride(L) + idx0] : 0;
        // frame 0: class-0 weight alpha for every state at idx0 (HaploBuilder.cpp:300-305)
        if (S0 > sc.ecap) { if (lane == 0) atomicExch(p.error, 1); continue; }
        if (S0 > 0) {
            const BlockLayout b0 = block_layout(S0, 1, 1);
            const double *alpha0 = (const double *) (p.arena + rec[idx0].off + b0.alpha);
            for (int s = lane; s < S0; s += 32) { sc.ekey[s] = s * 4; sc.ew[s] = alpha0[s]; }
        }
        __syncwarp();
        int depth = 0;
        if (lane == 0) { node_st[0] = root; iter_st[0] = 0; off_st[0] = 0; off_st[1] = S0; freq_st[0] = 1.0; }
        __syncwarp();
        while (depth >= 0) {
            const int node = node_st[depth];
            int j = iter_st[depth];
            int child = -1;
            while (j < A) { child = tr.child[(size_t) node * A + j]; if (child >= 0) break; ++j; }
            if (j >= A) { --depth; continue; }
            __syncwarp();
            if (lane == 0) iter_st[depth] = j + 1;
            const int locus = start + depth;            // the locus this child's allele j sits at
            const int e0 = off_st[depth], e1 = off_st[depth + 1];
            const int out0 = e1;
            work += (unsigned long long) (e1 - e0);
            int n_out = 0;
            double fsum = 0.0;
            int corr = 0;
            if (e1 > e0) {
                if (locus < head_len) {
                    // head loci: test the head patterns' own alleles, states stay at index head_len (:345-372)
                    const int S = rec[head_len].count;
                    const BlockLayout b = block_layout(S, 1, 1);
                    const int *ida = (const int *) (p.arena + rec[head_len].off + b.ida), *idb = (const int *) (p.arena + rec[head_len].off + b.idb);
                    const double *beta = (const double *) (p.arena + rec[head_len].off + b.beta);
                    if (lane == 0) s_cnt[wi] = 0;
                    __syncwarp();
                    for (int e = e0 + lane; e < e1; e += 32) {
                        const int key = sc.ekey[e], s = key >> 2, cls = key & 3;
                        const double wgt = sc.ew[e];
                        const bool ma = p.model.al[p.model.al_off[ida[s]] + locus] == j, mb = p.model.al[p.model.al_off[idb[s]] + locus] == j;
                        int nc = -1;
                        double nw = wgt;
                        if (cls == 0) { if (ma && mb) nc = 0; else if (ma) { nc = 1; nw = wgt * 0.5; } else if (mb) { nc = 2; nw = wgt * 0.5; } }
                        else if (cls == 1) { if (ma) nc = 1; }
                        else { if (mb) nc = 2; }
                        if (nc >= 0) {
                            const int pos = atomicAdd(&s_cnt[wi], 1);
                            if (out0 + pos < sc.ecap) { sc.ekey[out0 + pos] = s * 4 + nc; sc.ew[out0 + pos] = nw; }
                            fsum += nw * beta[s];
                        }
                    }
                    __syncwarp();
                    n_out = s_cnt[wi];
                    corr = (p.bexp ? p.bexp[(size_t) w * (L + 1) + head_len] : 0) + ea0;
                }
                else {
                    // move along the forward links of index `locus` into index locus+1 (:374-425)
                    const int S = rec[locus].count, Sn = rec[locus + 1].count;
                    if (Sn > 0 && rec[locus].fwd_slots > 0) {
                        const BlockLayout bn = block_layout(Sn, 1, 1);
                        const u64 *blk = p.arena + rec[locus + 1].off;
                        const unsigned char *abn = (const unsigned char *) (blk + bn.ab);
                        const double *tpn = (const double *) (blk + bn.tp), *betan = (const double *) (blk + bn.beta);
                        const int *fwd = (const int *) (p.arena + rec[locus].fwd_off);
                        StepShape sh;
                        sh.init(pn, gp, locus, S);
                        if (lane == 0) s_cnt[wi] = 0;
                        __syncwarp();
                        const int smax = sc.smax;
                        for (int e = e0 + lane; e < e1; e += 32) {
                            const int key = sc.ekey[e], s = key >> 2, cls = key & 3;
                            const double wgt = sc.ew[e];
                            sh.for_slots(s, [&](int c, int, int) {
                                const int f = fwd[c];
                                if (f == -1) return;
                                const int t = f & 0x7FFFFFFF;
                                const bool rev = ((u32) f >> 31) != 0;
                                const int aa = abn[t] & 15, ab = abn[t] >> 4;
                                int nc = -1;
                                double nw = wgt * tpn[t];
                                if (cls == 0) { if (aa == j && ab == j) nc = 0; else if (aa == j) { nc = 1; nw *= 0.5; } else if (ab == j) { nc = 2; nw *= 0.5; } }
                                else if (cls == 1) { if (!rev) { if (aa == j) nc = 1; } else { if (ab == j) nc = 2; } }
                                else { if (!rev) { if (ab == j) nc = 2; } else { if (aa == j) nc = 1; } }
                                // weights are products of positive numbers: a zero one (underflow, tp = 0) changes no sum below it and is
                                // dropped, so "the accumulator was 0" identifies the first contribution to (state, class) exactly
                                if (nc >= 0 && nw > 0.0) {
                                    if (atomicAdd(sc.acc + (size_t) nc * smax + t, nw) == 0.0) {
                                        const int pos = atomicAdd(&s_cnt[wi], 1);
                                        if (out0 + pos < sc.ecap) sc.ekey[out0 + pos] = t * 4 + nc;
                                    }
                                }
                            });
                        }
                        __syncwarp();
                        n_out = s_cnt[wi];
                        if (out0 + n_out <= sc.ecap) {
                            for (int q = lane; q < n_out; q += 32) {
                                const int key = sc.ekey[out0 + q], t = key >> 2, nc = key & 3;
                                const double v = sc.acc[(size_t) nc * smax + t];
                                sc.ew[out0 + q] = v;
                                sc.acc[(size_t) nc * smax + t] = 0.0;
                                fsum += v * betan[t];
                            }
                        }
                        corr = (p.bexp ? p.bexp[(size_t) w * (L + 1) + locus + 1] : 0) + ea0;
                    }
                }
            }
            if (out0 + n_out > sc.ecap) { if (lane == 0) atomicExch(p.error, 1); break; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fsum += __shfl_xor_sync(FULL, fsum, o);
            // freq = sum(w * beta) / P(G)  (:427-433); exponent bookkeeping only in scaled mode
            double freq = fsum / pg;
            if (p.aexp) freq = ldexp(freq, pge - corr);
            const int pid = tr.data[child];
            if (lane == 0 && pid >= 0) {        // :435-439
                if (freq != 0.0) atomicAdd(p.counts + pid, freq);
                const double lf = freq_st[depth];
                if (lf != 0.0) atomicAdd(p.counts + p.Pc + pid, lf);
            }
            if (n_out > 0 && depth + 1 >= ESTEP_MAX_DEPTH && lane == 0) {
                bool more = false;
                for (int q = 0; q < A; ++q) more |= tr.child[(size_t) child * A + q] >= 0;
                if (more) atomicExch(p.error, 2);
            }
            if (n_out > 0 && depth + 1 < ESTEP_MAX_DEPTH) {
                if (lane == 0) { node_st[depth + 1] = child; iter_st[depth + 1] = 0; off_st[depth + 2] = out0 + n_out; freq_st[depth + 1] = freq; }
                ++depth;
            }
            __syncwarp();
        }
        __syncwarp();
    }
    if (lane == 0 && work) atomicAdd(p.work, work);
}

}  // namespace hmc
