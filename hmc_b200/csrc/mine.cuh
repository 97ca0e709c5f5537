// K5 mine_patterns: the pattern search of the M-step on the device.
// Reference: PatternManager::searchPattern / generateCandidates (PatternManager.cpp:100-144),
// checkFrequencyWithExtension (:195-265), getMatchingFrequency (:267-291).
//
// The reference walks one trie per start locus depth first; every candidate is extended by the alleles of its end locus
// and its frequency is a sum over the matching individuals (genotype branch: product of per-locus factors) or over the
// matching sampled haplotypes (sampling branch: their weights), in ascending index order.  Here the tries of ALL start
// loci grow level by level (pattern length d -> d + 1):
//   sweep<COUNT>  a warp per extended parent walks the parent's match list 32 entries at a time; lane j owns child j and
//                 adds that child's terms one by one in list order (fp64 sums round exactly like the reference's loop)
//   decide        release / extend rule per child, transition probability
//   (two prefix sums: list offsets of the children that will be extended, child slots of the next level)
//   sweep<FILL>   the same walk writes the match lists of those children
// and the reference's pattern ids (release order of its depth-first stack: starts descending, children in descending
// allele order, parent before children) are recovered afterwards from subtree sizes.
#pragma once

#include "common.cuh"
#include "../../include/hmc_b200.h"

namespace hmc {

struct MineNodes {                  // SoA over all trie nodes of all start loci, level after level
    int *parent, *start, *cnt, *cbase, *sub, *idb;
    unsigned char *len, *flags;     // flags: 1 released as a pattern, 2 extended
    signed char *al;
    double *freq;
    long long *loff;
};

struct MineParams {
    int n_items, L, Amax;           // n_items: individuals (genotype branch) or sampled haplotypes
    int s0, nroots;                 // the start loci [s0, s0 + nroots) of this batch: trie node r < nroots is the root of start s0 + r
    const signed char *a0, *a1;     // locus-major allele indices [L][n_items]; a1 only in the genotype branch
    const double *w;                // sample weights
    const double *af;               // [L][Amax]
    const u32 *linfo;               // per locus: n_alleles | mask(allele frequency > 0) << 8
    const int *min_len, *max_len;   // [L]
    double denom, min_freq;
    MineNodes nd;
    int *list_id[2];
    double *list_fr[2];
    const int *ext;                 // extended parents of the current level
};

constexpr int MINE_RELEASED = 1, MINE_EXTENDED = 2;

__device__ __forceinline__ void mine_rule(const MineParams &m, int start, int len, double freq, bool &released, bool &extended)
{
    const int mn = m.min_len[start], mx = m.max_len[start], e = start + len;
    extended = (freq >= m.min_freq || len < mn) && e < m.L && len < mx;           // PatternManager.cpp:112-124
    released = (freq >= m.min_freq || len <= mn) && len > 0 && len >= mn;         // :125-135
}

__global__ void mine_roots_kernel(MineParams m, long long *need, long long *packed)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m.nroots) return;
    const int st = m.s0 + s;
    bool rel, ext;
    mine_rule(m, st, 0, 1.0, rel, ext);
    m.nd.parent[s] = -1; m.nd.start[s] = st; m.nd.len[s] = 0; m.nd.al[s] = -1; m.nd.freq[s] = 1.0; m.nd.cnt[s] = 0;
    m.nd.flags[s] = ext ? MINE_EXTENDED : 0;
    need[s] = 0;
    const int nchild = ext ? __popc(m.linfo[st] >> 8) : 0;
    packed[s] = ((long long) (ext ? 1 : 0) << 32) | nchild;
}

template <bool SAMPLES, bool FILL>
__global__ void __launch_bounds__(256) mine_sweep_kernel(MineParams m, int n_ext, int par)
{
    __shared__ double s_val[SAMPLES && !FILL ? 8 : 1][32];     // sampling COUNT: the weights of the current 32 entries, per warp
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const u32 lt = (1u << lane) - 1u;
    double *const sv = s_val[SAMPLES && !FILL ? (threadIdx.x >> 5) : 0];
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_ext; k += warps) {
        const int p = m.ext[k];
        const int start = m.nd.start[p], plen = m.nd.len[p], e = start + plen;
        const u32 info = m.linfo[e];
        const int A = (int) (info & 0xFF);
        const u32 am = info >> 8;
        const bool full = plen == 0;                        // checkFrequency: scan everything (PatternManager.cpp:146-193)
        const int n_in = full ? m.n_items : m.nd.cnt[p];
        const long long poff = full ? 0 : m.nd.loff[p];
        const int *ids = m.list_id[par] + poff;
        const double *frs = m.list_fr[par] + poff;
        const signed char *a0 = m.a0 + (size_t) e * m.n_items;
        const signed char *a1 = SAMPLES ? nullptr : m.a1 + (size_t) e * m.n_items;
        const bool has_child = lane < A && ((am >> lane) & 1u);
        const int child = m.nd.cbase[p] + __popc(am & lt);
        const double afj = (!SAMPLES && lane < A) ? m.af[(size_t) e * m.Amax + lane] : 0.0;
        double total = 0.0;
        int cnt = 0, want = 0;
        long long ooff = 0;
        if (FILL && has_child) { want = (m.nd.flags[child] & MINE_EXTENDED) ? 1 : 0; ooff = m.nd.loff[child]; }
        if (FILL && !__any_sync(0xffffffffu, want)) continue;
        for (int base = 0; base < n_in; base += 32) {
            const int q = base + lane;
            const bool valid = q < n_in;
            int i = 0, x = 0, y = 0;
            double pf = 1.0, wv = 0.0;
            if (valid) {
                i = full ? q : ids[q];
                x = a0[i];
                if (SAMPLES) { if (!FILL) wv = m.w[i]; }
                else { y = a1[i]; if (!full) pf = frs[q]; }
            }
            if constexpr (SAMPLES && !FILL) {
                // a sample's weight does not depend on the child: park the 32 weights in shared memory, collect child j's members
                // as a ballot mask in lane j, and let every child lane add its own terms in list order -- the chains of the
                // children run side by side instead of one shuffle + predicated add per term for the whole warp
                sv[lane] = wv;
                u32 mine = 0u;
                for (int j = 0; j < A; ++j) {
                    if (!((am >> j) & 1u)) continue;
                    const u32 mask = __ballot_sync(0xffffffffu, valid && (x < 0 || x == j));
                    if (lane == j) { mine = mask; cnt += __popc(mask); }
                }
                __syncwarp();
                while (mine) { const int b = __ffs(mine) - 1; mine &= mine - 1; total += sv[b]; }
                __syncwarp();
            }
            else
            for (int j = 0; j < A; ++j) {
                if (!((am >> j) & 1u)) continue;
                bool emit;
                double v;
                if (SAMPLES) { emit = valid && (x < 0 || x == j); v = wv; }
                else {
                    emit = valid && (x < 0 || y < 0 || x == j || y == j);
                    const double a = __shfl_sync(0xffffffffu, afj, j);
                    double f = 0.0;                         // getMatchingFrequency, one locus (PatternManager.cpp:267-291)
                    if (x < 0) f += a; else if (x == j) f += 1.0;
                    if (y < 0) f += a; else if (y == j) f += 1.0;
                    const double tf = 0.5 * f;
                    v = full ? tf : pf * tf;
                }
                const u32 mask = __ballot_sync(0xffffffffu, emit);
                if (!FILL) {
                    u32 mm = mask;
                    while (mm) {                            // ordered sum: one term at a time, in list order
                        const int b = __ffs(mm) - 1;
                        mm &= mm - 1;
                        const double t = __shfl_sync(0xffffffffu, v, b);
                        if (lane == j) total += t;
                    }
                }
                else {
                    const int wj = __shfl_sync(0xffffffffu, want, j);
                    if (wj) {
                        const long long oj = __shfl_sync(0xffffffffu, ooff, j);
                        const int cj = __shfl_sync(0xffffffffu, cnt, j);
                        if (emit) {
                            const long long pos = oj + cj + __popc(mask & lt);
                            m.list_id[par ^ 1][pos] = i;
                            if (!SAMPLES) m.list_fr[par ^ 1][pos] = v;
                        }
                    }
                }
                if (lane == j) cnt += __popc(mask);
            }
        }
        if (!FILL && has_child) {
            m.nd.parent[child] = p; m.nd.start[child] = start; m.nd.len[child] = (unsigned char) (plen + 1); m.nd.al[child] = (signed char) lane;
            m.nd.freq[child] = total / m.denom;
            m.nd.cnt[child] = cnt;
        }
    }
}

__global__ void mine_decide_kernel(MineParams m, int c0, int n, long long *need, long long *packed)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = c0 + t;
    const int start = m.nd.start[c], len = m.nd.len[c];
    bool rel, ext;
    mine_rule(m, start, len, m.nd.freq[c], rel, ext);
    m.nd.flags[c] = (unsigned char) ((rel ? MINE_RELEASED : 0) | (ext ? MINE_EXTENDED : 0));
    need[t] = ext ? m.nd.cnt[c] : 0;
    const int nchild = ext ? __popc(m.linfo[start + len] >> 8) : 0;
    packed[t] = ((long long) (ext ? 1 : 0) << 32) | nchild;
}

struct MineTotals { long long list_total, n_ext, n_child; };

// after the two exclusive scans: list offsets, the next level's parent list and its child slots
__global__ void mine_link_kernel(MineParams m, int c0, int n, const long long *need, const long long *need_scan, const long long *packed,
                                 const long long *packed_scan, int *ext_next, int child_begin_next, MineTotals *tot)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = c0 + t;
    m.nd.loff[c] = need_scan[t];
    const long long ps = packed_scan[t];
    if (m.nd.flags[c] & MINE_EXTENDED) {
        ext_next[(int) (ps >> 32)] = c;
        m.nd.cbase[c] = child_begin_next + (int) (ps & 0xFFFFFFFFll);
    }
    else m.nd.cbase[c] = -1;
    if (t == n - 1) {
        const long long pt = ps + packed[t];
        tot->list_total = need_scan[t] + need[t];
        tot->n_ext = pt >> 32;
        tot->n_child = pt & 0xFFFFFFFFll;
    }
}

// released patterns in the subtree of every node, one level at a time from the deepest
__global__ void mine_sub_kernel(MineParams m, int c0, int n)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = c0 + t;
    const int fl = m.nd.flags[c];
    int sub = (fl & MINE_RELEASED) ? 1 : 0;
    if (fl & MINE_EXTENDED) {
        const int nchild = __popc(m.linfo[m.nd.start[c] + m.nd.len[c]] >> 8), cb = m.nd.cbase[c];
        for (int r = 0; r < nchild; ++r) sub += m.nd.sub[cb + r];
    }
    m.nd.sub[c] = sub;
}

// start loci are visited in descending order (the reference pops its candidate stack from the back)
__global__ void mine_root_ids_kernel(MineParams m, long long *P_out)
{
    if (blockIdx.x || threadIdx.x) return;
    long long run = 0;
    for (int s = m.nroots - 1; s >= 0; --s) { m.nd.idb[s] = (int) run; run += m.nd.sub[s]; }
    *P_out = run;
}

// a parent is released before its children; children are visited in descending allele order
__global__ void mine_ids_kernel(MineParams m, int c0, int n)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = c0 + t;
    const int fl = m.nd.flags[c];
    if (!(fl & MINE_EXTENDED)) return;
    const int nchild = __popc(m.linfo[m.nd.start[c] + m.nd.len[c]] >> 8), cb = m.nd.cbase[c];
    int run = m.nd.idb[c] + ((fl & MINE_RELEASED) ? 1 : 0);
    for (int r = nchild - 1; r >= 0; --r) { m.nd.idb[cb + r] = run; run += m.nd.sub[cb + r]; }
}

struct MineOut {
    int *start, *end;
    double *freq, *prefix, *tp;
    long long *len64, *al_off;      // [P], [P + 1]
    signed char *al;
    int *node_of;                   // [P] trie node of every pattern
};

__global__ void mine_emit_kernel(MineParams m, MineOut o, int c0, int n)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int c = c0 + t;
    if (!(m.nd.flags[c] & MINE_RELEASED)) return;
    const int id = m.nd.idb[c], len = m.nd.len[c], start = m.nd.start[c];
    const double f = m.nd.freq[c], pf = m.nd.freq[m.nd.parent[c]];
    double tp = pf > 0 ? f / pf : f;                         // PatternManager.cpp:128-131
    tp = tp < 1.0 ? tp : 1.0;                               // HaploPattern.h:47
    o.start[id] = start; o.end[id] = start + len; o.freq[id] = f; o.prefix[id] = pf; o.tp[id] = tp;
    o.len64[id] = len; o.node_of[id] = c;
}

__global__ void mine_alleles_kernel(MineParams m, MineOut o, long long P)
{
    const long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (id >= P) return;
    int c = o.node_of[id];
    signed char *dst = o.al + o.al_off[id];
    for (int k = (int) m.nd.len[c] - 1; k >= 0; --k) { dst[k] = m.nd.al[c]; c = m.nd.parent[c]; }
    if (id == P - 1) o.al_off[P] = o.al_off[id] + o.len64[id];
}

__global__ void mine_add_offset_kernel(long long *v, long long n, long long add)
{
    const long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (i < n) v[i] += add;
}

// the pattern list as the host's records (include/hmc_b200.h hmc_pattern_rec)
__global__ void mine_records_kernel(MineOut o, hmc_pattern_rec *rec, long long P)
{
    const long long id = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (id >= P) return;
    hmc_pattern_rec r;
    r.start = o.start[id]; r.end = o.end[id]; r.al = o.al_off[id]; r.frequency = o.freq[id]; r.prefix_freq = o.prefix[id]; r.transition_prob = o.tp[id];
    rec[id] = r;
}

// [n][L] -> [L][n] (genotype branch: the two haplotype rows of an individual go to a0 / a1)
__global__ void mine_transpose_kernel(const signed char *src, int n, int L, int rows_per_item, signed char *a0, signed char *a1)
{
    __shared__ signed char tile[2][32][33];
    const int i0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, l = l0 + threadIdx.x;
        if (i < n && l < L) {
            tile[0][r][threadIdx.x] = src[((size_t) i * rows_per_item) * L + l];
            if (rows_per_item == 2) tile[1][r][threadIdx.x] = src[((size_t) i * 2 + 1) * L + l];
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int l = l0 + r, i = i0 + threadIdx.x;
        if (i < n && l < L) {
            a0[(size_t) l * n + i] = tile[0][threadIdx.x][r];
            if (rows_per_item == 2) a1[(size_t) l * n + i] = tile[1][threadIdx.x][r];
        }
    }
}

// sampled haplotypes taken from the result rows of the last resolve pass: a0[l][s] = haps[rows[s]][l]
__global__ void mine_transpose_rows_kernel(const signed char *haps, const int *rows, int n, int L, signed char *a0)
{
    __shared__ signed char tile[32][33];
    const int i0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, l = l0 + threadIdx.x;
        if (i < n && l < L) tile[r][threadIdx.x] = haps[(size_t) rows[i] * L + l];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int l = l0 + r, i = i0 + threadIdx.x;
        if (i < n && l < L) a0[(size_t) l * n + i] = tile[threadIdx.x][r];
    }
}

}  // namespace hmc
