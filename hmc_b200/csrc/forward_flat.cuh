// K1 forward_resolve, level-synchronous ("flat") variant: one locus at a time for ALL individuals of the wave.
// Same algorithm and same results as forward.cuh (HaploBuilder::resolve, HaploBuilder.cpp:35-86,226-261 and
// HaploPair.cpp:35-89), but every pass is its own low-register, fully occupied kernel over flat arrays:
//   head_count/head_write  head states (initHeadList)
//   F0  per individual: candidate slots C = S x (ordered allele pairs), hash-region size; prefix sums -> bases
//   FA  per candidate slot: 128-bit successor gathers, canonicalise, insert into the individual's hash region
//   FB  per candidate slot: creator flag (slot ordinal == region minimum) and arrival count, tile-local prefix
//   FS  prefix over the tile sums (one CTA)
//   F2  per individual: new state count, arrival base, trellis block offsets (prefix sums: no allocator atomics)
//   FC  per creator: global state index (= the reference's creation order inside the individual), segment offset
//   FD  per candidate slot: scatter into the arrival segments
//   FE  per new state: ordinal order inside the segment, alpha in the reference's order, K-best merge, trellis write,
//       and the hash entries it used are reset (the table is clean again without a clearing pass)
//   F3  per individual: roll the per-individual state forward (and the exponent-only rescale decision)
// All sizes live in device memory; kernels are grid-stride, so the whole locus loop is enqueued without host syncs.
#pragma once

#include "common.cuh"
#include "forward.cuh"

namespace hmc {

constexpr int FLAT_TILE = 1024;         // candidates per scan tile (256 threads x 4)

struct FlatTotals {                     // device scalars
    long long C_tot, H_tot, S_in_tot, S_out_tot, A_tot;
    long long arena_cursor;
    int locus, abort_code;              // abort_code: 0 ok, ST_NOMEM, ST_TOO_BIG
    int any_rescale, pad;
};

struct FlatParams {
    DevPanel panel;
    DevModel model;
    u64 *arena;
    long long arena_cap;
    LocusRec *loci;
    const int *wave;
    int n_wave, K, scaled;
    int *status, *final_exp;
    long long *counters;
    FlatTotals *tot;
    // per individual
    int *S, *S_out, *C, *hsz, *expo, *rescale;
    long long *off, *new_off, *cbase, *hbase, *sbase_in, *sbase_out, *abase;
    u64 *amax;
    // per candidate / hash / state
    long long hcap, ccap, scap;
    u64 *keys;
    double *tpv;
    u32 *ord, *cnt;
    int *cand, *candw, *arr;
    u64 *ex_local, *tile_sum, *tile_pre;
    long long *aoff;
    int *slot_of, *st_w;
    double *like[2];
    u64 *meta[2];
    int par;                            // ping-pong parity of like/meta (host-side copy)
};

__device__ __forceinline__ void flat_geno(const DevPanel &pn, int ind, int l, int &g0, int &g1)
{
    const u32 *gp = pn.packed + (size_t) ind * pn.words;
    const int bit = l * 2 * pn.bits;
    const u32 wd = gp[bit >> 5] >> (bit & 31);
    const u32 m = (1u << pn.bits) - 1u;
    g0 = (int) (wd & m) - 1;
    g1 = (int) ((wd >> pn.bits) & m) - 1;
}

// candidate slots per predecessor state at a locus, in the reference's extendAll order (HaploBuilder.cpp:48-79)
struct FlatShape {
    int mode, g0, g1, nA, nlive;
    u32 am;
    __device__ void init(const DevPanel &pn, int ind, int l)
    {
        flat_geno(pn, ind, l, g0, g1);
        const u32 info = pn.linfo[l];
        nA = (int) (info & 0xFF); am = info >> 8;
        nlive = __popc(am & ((nA >= 32) ? 0xFFFFFFFFu : ((1u << nA) - 1u)));
        mode = (g0 >= 0 && g1 >= 0) ? 0 : (g0 < 0 && g1 < 0) ? 3 : (g0 < 0 ? 1 : 2);
    }
    __device__ int mult() const
    {
        if (mode == 0) return g0 != g1 ? 2 : 1;
        if (mode == 3) return nlive * nlive;
        const int known = (mode == 1) ? g1 : g0;
        return 2 * nlive - (((am >> known) & 1u) ? 1 : 0);
    }
    // slot -> (pred state, x, y); S = predecessor states of the individual
    __device__ void decode(int c, int S, int &s, int &x, int &y) const
    {
        if (mode == 0) {
            if (g0 != g1) { s = c >> 1; const int d = c & 1; x = d ? g1 : g0; y = d ? g0 : g1; }
            else { s = c; x = y = g0; }
            return;
        }
        int rem = c;
        if (mode == 3) {
            for (int j = 0; j < nA; ++j) { if (!((am >> j) & 1u)) continue;
                for (int k = j; k < nA; ++k) { if (!((am >> k) & 1u)) continue;
                    const int span = S * (j != k ? 2 : 1);
                    if (rem < span) { if (j != k) { s = rem >> 1; const int d = rem & 1; x = d ? k : j; y = d ? j : k; } else { s = rem; x = y = j; } return; }
                    rem -= span;
                } }
        }
        else {
            const int known = (mode == 1) ? g1 : g0;
            for (int j = 0; j < nA; ++j) { if (!((am >> j) & 1u)) continue;
                const int span = S * (j != known ? 2 : 1);
                if (rem < span) { if (j != known) { s = rem >> 1; const int d = rem & 1; x = d ? known : j; y = d ? j : known; } else { s = rem; x = y = j; } return; }
                rem -= span;
            }
        }
        s = 0; x = y = 0;
    }
};

// block-wide exclusive scan of one long long per thread (1024 threads max)
__device__ __forceinline__ long long flat_block_scan(long long v, long long &total, long long *s_tmp)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();
    if (lane == 31) s_tmp[wid] = x;
    __syncthreads();
    long long base = 0, tot = 0;
    for (int w = 0; w < nw; ++w) { const long long t = s_tmp[w]; if (w < wid) base += t; tot += t; }
    total = tot;
    return base + x - v;
}

// ---------------------------------------------------------------------------------------------
// head states: the serial enumeration of forward.cuh, run twice (count, then write)
template <bool WRITE>
__device__ int flat_heads(const FlatParams &p, int w, int *out_a, int *out_b)
{
    const DevPanel &pn = p.panel;
    const DevModel &md = p.model;
    const int head_len = md.head_len, ind = p.wave[w];
    int nh = 0;
    int hg0[8], hg1[8];
    for (int j = 0; j < head_len; ++j) flat_geno(pn, ind, j, hg0[j], hg1[j]);
    for (int hi = 0; hi < md.n_heads; ++hi) {
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
            else { const int v = hg0[j] < 0 ? 0 : hg0[j]; lo[j] = hi_[j] = v; }
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
                if (hp < 0) return -1;
                if (hp >= head) { if (WRITE) { out_a[nh] = head; out_b[nh] = hp; } ++nh; }
            }
            int j = head_len - 1;
            while (j >= 0) { if (++curv[j] <= hi_[j]) break; curv[j] = lo[j]; --j; }
            if (j < 0) break;
        }
    }
    return nh;
}

// one CTA: counts heads, lays out the head blocks, writes them
__global__ void __launch_bounds__(1024) flat_head_kernel(FlatParams p)
{
    __shared__ long long s_tmp[32];
    __shared__ long long s_run[2];
    const int L = p.panel.L, K = p.K, head_len = p.model.head_len;
    if (threadIdx.x == 0) { s_run[0] = 0; s_run[1] = p.tot->arena_cursor; }
    __syncthreads();
    for (int base = 0; base < p.n_wave; base += blockDim.x) {
        const int w = base + threadIdx.x;
        int nh = 0;
        if (w < p.n_wave) {
            LocusRec *rec = p.loci + (size_t) w * (L + 1);
            for (int l = 0; l <= L; ++l) { rec[l].count = 0; rec[l].fwd_slots = 0; rec[l].off = 0; rec[l].fwd_off = 0; }
            nh = flat_heads<false>(p, w, nullptr, nullptr);
            if (nh < 0) { p.status[w] = ST_NO_HEAD; nh = 0; }
            else p.status[w] = ST_DONE;
            p.expo[w] = 0; p.final_exp[w] = 0;
        }
        long long tot;
        const long long units = nh > 0 ? ((block_layout(nh, K, 0).total + 1) & ~1LL) : 0;
        const long long ex = flat_block_scan(((long long) nh << 40) | units, tot, s_tmp);
        if (w < p.n_wave) {
            const long long sb = s_run[0] + (ex >> 40), ob = s_run[1] + (ex & 0xFFFFFFFFFFll);
            p.S[w] = nh; p.sbase_in[w] = sb; p.off[w] = ob;
            if (nh > 0) {
                if (ob + units > p.arena_cap || sb + nh > p.scap) { atomicExch(&p.tot->abort_code, ob + units > p.arena_cap ? ST_NOMEM : ST_TOO_BIG); }
                else {
                    LocusRec *rec = p.loci + (size_t) w * (L + 1);
                    rec[head_len].off = ob; rec[head_len].count = nh;
                    const BlockLayout bl = block_layout(nh, K, 0);
                    int *t_ida = (int *) (p.arena + ob + bl.ida), *t_idb = (int *) (p.arena + ob + bl.idb);
                    double *t_alpha = (double *) (p.arena + ob + bl.alpha);
                    unsigned char *t_ab = (unsigned char *) (p.arena + ob + bl.ab);
                    u32 *t_bp = (u32 *) (p.arena + ob + bl.bp);
                    flat_heads<true>(p, w, t_ida, t_idb);
                    for (int s = 0; s < nh; ++s) {
                        const int a = t_ida[s], b = t_idb[s];
                        const double f = __dmul_rn(p.model.freq[a], p.model.freq[b]);
                        t_alpha[s] = (a == b) ? f : __dmul_rn(f, 2.0);
                        t_ab[s] = (unsigned char) (p.model.last[a] | (p.model.last[b] << 4));
                        t_bp[(size_t) s * K] = BP_HEAD | (a == b ? BP_HOM : 0u);
                        p.like[0][(size_t) (sb + s) * K] = f;
                        p.meta[0][sb + s] = 1ull | ((a == b ? 1ull : 0ull) << 8);
                        p.st_w[sb + s] = w;
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_run[0] += tot >> 40; s_run[1] += tot & 0xFFFFFFFFFFll; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.sbase_in[p.n_wave] = s_run[0];
        p.tot->S_in_tot = s_run[0];
        p.tot->arena_cursor = s_run[1];
        p.tot->locus = head_len;
        atomicAdd((unsigned long long *) p.counters + 1, (unsigned long long) s_run[0]);
    }
}

// F0: one CTA over the individuals: candidate counts and hash-region sizes, prefix sums
__global__ void __launch_bounds__(1024) flat_f0_kernel(FlatParams p)
{
    __shared__ long long s_tmp[32];
    __shared__ long long s_run[2];
    if (p.tot->abort_code) return;
    const int locus = p.tot->locus;
    if (threadIdx.x == 0) { s_run[0] = 0; s_run[1] = 0; }
    __syncthreads();
    for (int base = 0; base < p.n_wave; base += blockDim.x) {
        const int w = base + threadIdx.x;
        long long C = 0, H = 0;
        if (w < p.n_wave && p.status[w] == ST_DONE && p.S[w] > 0) {
            FlatShape sh;
            sh.init(p.panel, p.wave[w], locus);
            C = (long long) p.S[w] * sh.mult();
            H = 16;
            while (H * 4 < 5 * C) H <<= 1;
        }
        long long t1, t2;
        const long long e1 = flat_block_scan(C, t1, s_tmp);
        const long long e2 = flat_block_scan(H, t2, s_tmp);
        if (w < p.n_wave) { p.C[w] = (int) C; p.hsz[w] = (int) H; p.cbase[w] = s_run[0] + e1; p.hbase[w] = s_run[1] + e2; p.amax[w] = 0ull; }
        __syncthreads();
        if (threadIdx.x == 0) { s_run[0] += t1; s_run[1] += t2; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.cbase[p.n_wave] = s_run[0]; p.hbase[p.n_wave] = s_run[1];
        p.tot->C_tot = s_run[0]; p.tot->H_tot = s_run[1];
        if (s_run[0] > p.ccap || s_run[1] > p.hcap) atomicExch(&p.tot->abort_code, ST_TOO_BIG);
    }
}

__device__ __forceinline__ int flat_find(const long long *base, int n, long long c)
{
    int lo = 0, hi = n;     // largest w with base[w] <= c
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (base[mid] <= c) lo = mid; else hi = mid; }
    return lo;
}

// FA: per candidate slot
__global__ void __launch_bounds__(256) flat_fa_kernel(FlatParams p)
{
    if (p.tot->abort_code) return;
    const long long C_tot = p.tot->C_tot;
    const int locus = p.tot->locus, K = p.K, A = p.model.Amax;
    const int4 *succ16 = reinterpret_cast<const int4 *>(p.model.succ16);
    for (long long c = blockIdx.x * (long long) blockDim.x + threadIdx.x; c < C_tot; c += (long long) gridDim.x * blockDim.x) {
        const int w = flat_find(p.cbase, p.n_wave, c);
        const int lc = (int) (c - p.cbase[w]);
        const int S = p.S[w];
        FlatShape sh;
        sh.init(p.panel, p.wave[w], locus);
        int s, x, y;
        sh.decode(lc, S, s, x, y);
        const BlockLayout bl = block_layout(S, K, 0);
        const u64 *blk = p.arena + p.off[w];
        const int pa = ((const int *) (blk + bl.ida))[s], pb = ((const int *) (blk + bl.idb))[s];
        const double al = ((const double *) (blk + bl.alpha))[s];
        int enc = -1;
        if (al > 0.0) {     // HaploBuilder.cpp:237
            const int4 ea = __ldcg(succ16 + (size_t) pa * A + x), eb = __ldcg(succ16 + (size_t) pb * A + y);
            int hpa = ea.x, hpb = eb.x;
            if (hpa >= 0 && hpb >= 0) {
                const double tv = __dmul_rn(__hiloint2double(ea.w, ea.z), __hiloint2double(eb.w, eb.z));
                u32 rev = 0;
                if (hpa > hpb) { const int t = hpa; hpa = hpb; hpb = t; rev = 1; }
                const u64 key = ((u64) (u32) hpa << 32) | (u32) hpb;
                const u32 hmask = (u32) p.hsz[w] - 1u;
                const long long hb = p.hbase[w];
                u32 h = mix64(key) & hmask;
                for (;;) {
                    const u64 prev = atomicCAS(p.keys + hb + h, KEY_EMPTY, key);
                    if (prev == KEY_EMPTY) { p.tpv[hb + h] = tv; break; }
                    if (prev == key) break;
                    h = (h + 1) & hmask;
                }
                atomicMin(p.ord + hb + h, (u32) lc);
                atomicAdd(p.cnt + hb + h, 1u);
                enc = (int) (h | (rev << 30));
            }
        }
        p.cand[c] = enc;
        p.candw[c] = w;
    }
}

// FB: creator flags + arrival counts, tile-local exclusive prefix (creators << 40 | arrivals)
__global__ void __launch_bounds__(256) flat_fb_kernel(FlatParams p)
{
    __shared__ long long s_tmp[32];
    if (p.tot->abort_code) return;
    const long long C_tot = p.tot->C_tot;
    const long long n_tiles = (C_tot + FLAT_TILE - 1) / FLAT_TILE;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long c0 = tile * FLAT_TILE + threadIdx.x * 4;
        long long v[4];
        long long sum = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long c = c0 + u;
            v[u] = 0;
            if (c < C_tot) {
                const int enc = p.cand[c];
                if (enc >= 0) {
                    const int w = p.candw[c];
                    const long long hg = p.hbase[w] + (enc & 0x3FFFFFFF);
                    if (p.ord[hg] == (u32) (c - p.cbase[w])) v[u] = (1ll << 40) | (long long) p.cnt[hg];
                }
            }
            sum += v[u];
        }
        long long tot;
        long long ex = flat_block_scan(sum, tot, s_tmp);
#pragma unroll
        for (int u = 0; u < 4; ++u) { if (c0 + u < C_tot) p.ex_local[c0 + u] = (u64) ex; ex += v[u]; }
        if (threadIdx.x == 0) p.tile_sum[tile] = (u64) tot;
        __syncthreads();
    }
}

// FS: exclusive prefix over the tile sums (one CTA)
__global__ void __launch_bounds__(1024) flat_fs_kernel(FlatParams p)
{
    __shared__ long long s_tmp[32];
    __shared__ long long s_run;
    if (p.tot->abort_code) return;
    const long long n_tiles = (p.tot->C_tot + FLAT_TILE - 1) / FLAT_TILE;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (long long base = 0; base < n_tiles; base += blockDim.x) {
        const long long t = base + threadIdx.x;
        const long long v = t < n_tiles ? (long long) p.tile_sum[t] : 0;
        long long tot;
        const long long ex = flat_block_scan(v, tot, s_tmp);
        if (t < n_tiles) p.tile_pre[t] = (u64) (s_run + ex);
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.tot->S_out_tot = s_run >> 40;
        p.tot->A_tot = s_run & 0xFFFFFFFFFFll;
        if ((s_run >> 40) > p.scap) atomicExch(&p.tot->abort_code, ST_TOO_BIG);
    }
}

__device__ __forceinline__ long long flat_prefix(const FlatParams &p, long long c, long long C_tot)
{
    if (c >= C_tot) return (p.tot->S_out_tot << 40) | p.tot->A_tot;
    return (long long) p.tile_pre[c / FLAT_TILE] + (long long) p.ex_local[c];
}

// F2: one CTA over the individuals: new state counts, bases, trellis blocks of index locus+1
__global__ void __launch_bounds__(1024) flat_f2_kernel(FlatParams p)
{
    __shared__ long long s_tmp[32];
    __shared__ long long s_run;
    __shared__ long long s_cnt[4];
    if (p.tot->abort_code) return;
    const int L = p.panel.L, K = p.K, locus = p.tot->locus;
    const long long C_tot = p.tot->C_tot;
    if (threadIdx.x == 0) { s_run = p.tot->arena_cursor; s_cnt[0] = s_cnt[1] = s_cnt[2] = s_cnt[3] = 0; }
    __syncthreads();
    for (int base = 0; base < p.n_wave; base += blockDim.x) {
        const int w = base + threadIdx.x;
        int so = 0;
        long long units = 0;
        if (w < p.n_wave) {
            const long long p0 = flat_prefix(p, p.cbase[w], C_tot), p1 = flat_prefix(p, p.cbase[w + 1], C_tot);
            so = (int) ((p1 >> 40) - (p0 >> 40));
            p.S_out[w] = so;
            p.sbase_out[w] = p0 >> 40;
            p.abase[w] = p0 & 0xFFFFFFFFFFll;
            if (so > 0) units = (block_layout(so, K, 0).total + 1) & ~1LL;
            if (so >= MAX_STATES_PER_LOCUS) atomicExch(&p.tot->abort_code, ST_TOO_BIG);
        }
        long long tot;
        const long long ex = flat_block_scan(units, tot, s_tmp);
        if (w < p.n_wave) {
            const long long ob = s_run + ex;
            p.new_off[w] = ob;
            if (so > 0) {
                if (ob + units > p.arena_cap) atomicExch(&p.tot->abort_code, ST_NOMEM);
                else { LocusRec *rec = p.loci + (size_t) w * (L + 1); rec[locus + 1].off = ob; rec[locus + 1].count = so; }
            }
        }
        {   // wave statistics: warp reduction, one shared update per warp
            long long c0 = so > 0 ? 1 : 0, c1 = so, c2 = so, c3 = 0;
            if (w < p.n_wave && so > 0) c3 = (flat_prefix(p, p.cbase[w + 1], C_tot) & 0xFFFFFFFFFFll) - p.abase[w];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                c2 = max(c2, __shfl_xor_sync(0xffffffffu, c2, o)); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            }
            if ((threadIdx.x & 31) == 0 && c0 > 0) {
                atomicAdd((unsigned long long *) &s_cnt[0], (unsigned long long) c0);
                atomicAdd((unsigned long long *) &s_cnt[1], (unsigned long long) c1);
                atomicMax(&s_cnt[2], c2);
                atomicAdd((unsigned long long *) &s_cnt[3], (unsigned long long) c3);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.sbase_out[p.n_wave] = p.tot->S_out_tot;
        p.tot->arena_cursor = s_run;
        atomicAdd((unsigned long long *) p.counters + 0, (unsigned long long) s_cnt[0]);
        atomicAdd((unsigned long long *) p.counters + 1, (unsigned long long) s_cnt[1]);
        atomicMax((long long *) p.counters + 2, s_cnt[2]);
        atomicAdd((unsigned long long *) p.counters + 3, (unsigned long long) s_cnt[3]);
    }
}

// FC: creators publish their state
__global__ void __launch_bounds__(256) flat_fc_kernel(FlatParams p)
{
    if (p.tot->abort_code) return;
    const long long C_tot = p.tot->C_tot;
    for (long long c = blockIdx.x * (long long) blockDim.x + threadIdx.x; c < C_tot; c += (long long) gridDim.x * blockDim.x) {
        const int enc = p.cand[c];
        if (enc < 0) continue;
        const int w = p.candw[c];
        const long long hg = p.hbase[w] + (enc & 0x3FFFFFFF);
        if (p.ord[hg] != (u32) (c - p.cbase[w])) continue;
        const long long pre = (long long) p.tile_pre[c / FLAT_TILE] + (long long) p.ex_local[c];
        const long long gidx = pre >> 40;
        p.slot_of[gidx] = (int) hg;
        p.aoff[gidx] = pre & 0xFFFFFFFFFFll;
        p.st_w[gidx] = w;
        p.ord[hg] = 0x80000000u | (u32) (gidx - p.sbase_out[w]);
        p.cnt[hg] = 0;
    }
}

// FD: scatter arrivals
__global__ void __launch_bounds__(256) flat_fd_kernel(FlatParams p)
{
    if (p.tot->abort_code) return;
    const long long C_tot = p.tot->C_tot;
    for (long long c = blockIdx.x * (long long) blockDim.x + threadIdx.x; c < C_tot; c += (long long) gridDim.x * blockDim.x) {
        const int enc = p.cand[c];
        if (enc < 0) continue;
        const int w = p.candw[c];
        const long long hg = p.hbase[w] + (enc & 0x3FFFFFFF);
        const u32 rev = ((u32) enc >> 30) & 1u;
        const long long gidx = p.sbase_out[w] + (long long) (p.ord[hg] & 0x7FFFFFFFu);
        const u32 pos = atomicAdd(p.cnt + hg, 1u);
        p.arr[p.aoff[gidx] + pos] = (int) ((u32) (c - p.cbase[w]) | (rev << 31));
    }
}

// FE: per new state
template <int KMAX>
__global__ void __launch_bounds__(128) flat_fe_kernel(FlatParams p)
{
    if (p.tot->abort_code) return;
    const long long S_tot = p.tot->S_out_tot, A_tot = p.tot->A_tot;
    const int locus = p.tot->locus, K = p.K;
    const double *like_in = p.like[p.par];
    const u64 *meta_in = p.meta[p.par];
    double *like_out = p.like[p.par ^ 1];
    u64 *meta_out = p.meta[p.par ^ 1];
    for (long long gidx = blockIdx.x * (long long) blockDim.x + threadIdx.x; gidx < S_tot; gidx += (long long) gridDim.x * blockDim.x) {
        const int w = p.st_w[gidx];
        const int lidx = (int) (gidx - p.sbase_out[w]);
        const int hg = p.slot_of[gidx];
        const u64 key = p.keys[hg];
        const int hpa = (int) (key >> 32), hpb = (int) (key & 0xFFFFFFFFu);
        const double tpv = p.tpv[hg];
        const long long a0 = p.aoff[gidx];
        const int n = (int) ((gidx + 1 < S_tot ? p.aoff[gidx + 1] : A_tot) - a0);
        int *arr = p.arr + a0;
        if (n > 1 && n <= 8) {
            int v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = (q < n) ? arr[q] : 0x7FFFFFFF;
            auto cx = [&](int a, int b) { const int x = v[a], y = v[b]; const bool sw = (x & 0x7FFFFFFF) > (y & 0x7FFFFFFF); v[a] = sw ? y : x; v[b] = sw ? x : y; };
            cx(0, 1); cx(2, 3); cx(4, 5); cx(6, 7);
            cx(0, 2); cx(1, 3); cx(4, 6); cx(5, 7);
            cx(1, 2); cx(5, 6); cx(0, 4); cx(3, 7);
            cx(1, 5); cx(2, 6);
            cx(1, 4); cx(3, 6);
            cx(2, 4); cx(3, 5);
            cx(3, 4);
#pragma unroll
            for (int q = 0; q < 8; ++q) if (q < n) arr[q] = v[q];
        }
        else if (n > 8) {
            for (int j = 1; j < n; ++j) {
                const int e = arr[j];
                int q = j - 1;
                while (q >= 0 && (arr[q] & 0x7FFFFFFF) > (e & 0x7FFFFFFF)) { arr[q + 1] = arr[q]; --q; }
                arr[q + 1] = e;
            }
        }
        FlatShape sh;
        sh.init(p.panel, p.wave[w], locus);
        const int S_in = p.S[w];
        const long long sb_in = p.sbase_in[w];
        const BlockLayout bli = block_layout(S_in, K, 0);
        const double *gi_alpha = (const double *) (p.arena + p.off[w] + bli.alpha);
        double alpha = 0.0;
        double bl[KMAX];
        u32 bb[KMAX];
        int nl_out = 0, xa = 0, ya = 0;
        for (int j = 0; j < n; ++j) {
            const int e = arr[j];
            const u32 rev = (u32) e >> 31;
            int s, x, y;
            sh.decode(e & 0x7FFFFFFF, S_in, s, x, y);
            const bool het = (x != y);
            if (j == 0) { xa = rev ? y : x; ya = rev ? x : y; }
            const u64 mt = meta_in[sb_in + s];
            const int nl = (int) (mt & 0xFF);
            u32 hom = (u32) (mt >> 8);
            const double *lp = like_in + (size_t) (sb_in + s) * K;
            const double term = __dmul_rn(gi_alpha[s], tpv);
            alpha = (j == 0) ? term : __dadd_rn(alpha, term);
            nl_out += nl;
            const u32 zero = (het && rev) ? hom : 0u;       // het-locus rule, HaploPair.cpp:53-60,77-84
            if (het) hom = 0;
            const u32 bpbase = (u32) s | (rev ? BP_REV : 0u);
            if (j == 0 && zero == 0) {
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
                int k = 0;
                double cur = lp[0];
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
        const int S_out = p.S_out[w];
        const BlockLayout blo = block_layout(S_out, K, 0);
        u64 *blk = p.arena + p.new_off[w];
        u32 *t_bp = (u32 *) (blk + blo.bp);
        u32 hom_out = 0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < nl_out) {
                like_out[(size_t) gidx * K + k] = bl[k];
                t_bp[(size_t) lidx * K + k] = bb[k];
                if (bb[k] & BP_HOM) hom_out |= 1u << k;
            }
        }
        meta_out[gidx] = (u64) nl_out | ((u64) hom_out << 8);
        ((int *) (blk + blo.ida))[lidx] = hpa;
        ((int *) (blk + blo.idb))[lidx] = hpb;
        ((double *) (blk + blo.alpha))[lidx] = alpha;
        ((unsigned char *) (blk + blo.ab))[lidx] = (unsigned char) (xa | (ya << 4));
        if (p.scaled) atomicMax((unsigned long long *) &p.amax[w], (unsigned long long) __double_as_longlong(alpha));
        // leave the hash clean for the next locus
        p.keys[hg] = KEY_EMPTY; p.ord[hg] = 0xFFFFFFFFu; p.cnt[hg] = 0;
    }
}

// F3: one CTA over the individuals: roll forward
__global__ void __launch_bounds__(1024) flat_f3_kernel(FlatParams p)
{
    if (p.tot->abort_code) return;
    const int L = p.panel.L;
    int any = 0;
    for (int w = threadIdx.x; w < p.n_wave; w += blockDim.x) {
        const int so = p.S_out[w];
        int r = 0;
        if (p.scaled && so > 0) {
            const double am = __longlong_as_double((long long) p.amax[w]);
            if (am > 0.0 && am < 0x1p-512) { r = 1; p.expo[w] += 512; any = 1; }
        }
        p.rescale[w] = r;
        p.S[w] = so;
        p.off[w] = p.new_off[w];
        p.sbase_in[w] = p.sbase_out[w];
        if (p.tot->locus + 1 == L || so == 0) p.final_exp[w] = p.expo[w];
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) p.tot->any_rescale = 1;
    else if (threadIdx.x == 0) p.tot->any_rescale = 0;
    __syncthreads();
    if (threadIdx.x == 0) { p.sbase_in[p.n_wave] = p.tot->S_out_tot; p.tot->S_in_tot = p.tot->S_out_tot; p.tot->locus += 1; }
}

// exponent-only rescale of the states of flagged individuals (after F3: S/off/sbase_in already describe the new index);
// `par_new` = parity of the lists of the new index
__global__ void __launch_bounds__(256) flat_rescale_kernel(FlatParams p, int par_new)
{
    if (p.tot->abort_code || !p.tot->any_rescale) return;
    const long long S_tot = p.tot->S_in_tot;
    const int K = p.K;
    for (long long gidx = blockIdx.x * (long long) blockDim.x + threadIdx.x; gidx < S_tot; gidx += (long long) gridDim.x * blockDim.x) {
        const int w = p.st_w[gidx];
        if (!p.rescale[w]) continue;
        const int lidx = (int) (gidx - p.sbase_in[w]);
        const BlockLayout b = block_layout(p.S[w], K, 0);
        double *al = (double *) (p.arena + p.off[w] + b.alpha);
        al[lidx] *= 0x1p+512;
        const int nl = (int) (p.meta[par_new][gidx] & 0xFF);
        for (int k = 0; k < nl; ++k) p.like[par_new][(size_t) gidx * K + k] *= 0x1p+512;
    }
}

// final link lists for the traceback kernel: per individual a block [S x K likelihoods][S meta] recorded in rec[0]
__global__ void __launch_bounds__(1024) flat_final_alloc_kernel(FlatParams p)
{
    __shared__ long long s_tmp[32];
    __shared__ long long s_run;
    if (p.tot->abort_code) return;
    const int L = p.panel.L, K = p.K;
    if (threadIdx.x == 0) s_run = p.tot->arena_cursor;
    __syncthreads();
    for (int base = 0; base < p.n_wave; base += blockDim.x) {
        const int w = base + threadIdx.x;
        long long units = 0;
        int S = 0;
        if (w < p.n_wave && p.status[w] == ST_DONE) { S = p.loci[(size_t) w * (L + 1) + L].count; if (S > 0) units = ((long long) S * K + S + 1) & ~1LL; }
        long long tot;
        const long long ex = flat_block_scan(units, tot, s_tmp);
        if (w < p.n_wave && S > 0) {
            const long long ob = s_run + ex;
            if (ob + units > p.arena_cap) atomicExch(&p.tot->abort_code, ST_NOMEM);
            else { LocusRec *rec = p.loci + (size_t) w * (L + 1); rec[0].off = ob; rec[0].count = S; }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) p.tot->arena_cursor = s_run;
}

__global__ void __launch_bounds__(256) flat_final_copy_kernel(FlatParams p, int par_fin)
{
    if (p.tot->abort_code) return;
    const int L = p.panel.L, K = p.K;
    const long long S_tot = p.tot->S_in_tot;
    for (long long gidx = blockIdx.x * (long long) blockDim.x + threadIdx.x; gidx < S_tot; gidx += (long long) gridDim.x * blockDim.x) {
        const int w = p.st_w[gidx];
        const LocusRec *rec = p.loci + (size_t) w * (L + 1);
        if (rec[L].count <= 0) continue;        // only individuals alive at the last index own current states
        const int s = (int) (gidx - p.sbase_in[w]), S = rec[0].count;
        double *fl = (double *) (p.arena + rec[0].off);
        u64 *fm = p.arena + rec[0].off + (long long) S * K;
        const u64 mt = p.meta[par_fin][gidx];
        fm[s] = mt;
        const int nl = (int) (mt & 0xFF);
        for (int k = 0; k < nl; ++k) fl[(size_t) s * K + k] = p.like[par_fin][(size_t) gidx * K + k];
    }
}

}  // namespace hmc
