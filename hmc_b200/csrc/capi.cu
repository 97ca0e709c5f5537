// C-ABI of the hmc_b200 device path (include/hmc_b200.h): context, uploads (K0 pack / flatten),
// wave scheduling over the trellis arena, kernel launches, result download.
#include "../../include/hmc_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "forward.cuh"
#include "mine.cuh"
#include "automaton.cuh"
#include "compare.cuh"
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <climits>
#include <thread>
#include <atomic>
#include "kbest.cuh"
#include "estep.cuh"
#include "emit.cuh"

using namespace hmc;


#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return ctx->fail(HMC_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t want)
    {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        cudaError_t e = cudaMalloc((void **) &p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct ScratchPool {
    int slots = 0;
    int hcap = 0, ccap = 0, scap = 0;
    DevBuf<unsigned char> mem;
    DevBuf<Scratch> desc;
};

struct EstepTrie {      // host-built flattened ForwardPatternTree + its device copy
    DevBuf<int> d_root, d_child, d_data;
    DevTrie dev = {};
    void release() { d_root.release(); d_child.release(); d_data.release(); }
};

struct hmc_ctx {
    int device = 0;
    int sms = 0;
    int threads_opt = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaStream_t stream2 = nullptr;          // the K-best kernel of a resolve pass runs here, beside the forward kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;

    // panel
    int N = 0, L = 0, Amax = 0, bits = 0, words = 0;
    double missing_frac = 0;
    DevBuf<u32> d_packed, d_linfo;
    DevBuf<double> d_af;            // allele frequencies [L][Amax] (K5 mining)
    int panel_serial = 0;           // bumped by every hmc_set_panel
    struct MinePool *mine = nullptr;
    struct AutPool *aut = nullptr;
    DevBuf<signed char> d_cmp_real, d_cmp_infer;
    DevBuf<u32> d_best_packed;
    DevBuf<CompareOut> d_cmp_out;
    DevBuf<signed char> d_raw;      // staging of the unpacked genotypes (kept: cudaMalloc/cudaFree per call cost 100s of ms)
    DevBuf<int> d_order;
    std::vector<int> h_order, h_rank;
    DevBuf<long long> d_work;       // forward-kernel cycles per individual of the last pass that resolved it (0 = not yet): the next pass's work order
    std::vector<long long> h_work;
    long long seen_max_states = 0;  // most live states at one index over the passes on this panel: sizes the per-CTA spill scratch
    double est_units[2] = {0, 0};   // arena units (8 B) one individual's trellis took in the last waves, resolve / E-step mode: sizes the next wave
    std::vector<int> h_nall;

    // model
    bool have_model = false;
    DevModel dm = {};
    DevBuf<int> d_succ, d_start, d_heads, d_head_lut;
    DevBuf<int4> d_succ16;
    DevBuf<double> d_tp, d_freq;
    DevBuf<unsigned char> d_last;
    DevBuf<int> d_end_tmp, d_model_err, d_miss;
    DevBuf<double> d_counts_snap;
    // E-step trellis of the last single-wave hmc_estimate_counts call (alpha, beta, forward links stay in the arena): the E-passes of one
    // M-step differ only in their candidate list, so K1 + K3 run once and the later passes go straight to K4
    struct EstepCache { bool valid = false; int first = 0, count = 0, n_wave = 0; unsigned flags = 0; FwdParams fp; int64_t counters[HMC_CNT_N]; } ecache;
    DevBuf<long long> d_al_off;
    DevBuf<signed char> d_al;

    // arena + wave state
    size_t arena_req = 0;
    DevBuf<u64> d_arena;
    DevBuf<unsigned long long> d_cursor;     // [0] cursor
    DevBuf<int> d_flags;                     // [0] exhausted, [1] queue
    DevBuf<long long> d_counters;
    DevBuf<LocusRec> d_loci;
    DevBuf<int> d_wave, d_status, d_final_exp, d_locus_exp, d_done;
    DevBuf<long long> d_cycles;
    DevBuf<u64> d_kb_eoff;
    DevBuf<u32> d_kb_chain, d_kb_stack;
    int64_t kb_work[3] = {};
    std::vector<int> h_wave;                 // last wave (for hmc_get_trellis)
    int last_K = 0, last_estep = 0;
    ScratchPool pool, big_pool;

    // results of the last resolve
    int r_first = 0, r_count = 0, r_K = 0;
    DevBuf<double> d_pg, d_cov, d_prior, d_post;
    DevBuf<int> d_exp, d_nres;
    DevBuf<signed char> d_haps;
    DevBuf<double> d_lk;                     // K8: haplotype likelihoods, text emission
    DevBuf<int> d_emit_sym;
    DevBuf<unsigned char> d_emit_wid, d_emit_tmp;
    DevBuf<long long> d_emit_off;
    DevBuf<char> d_emit_text;

    // E-step
    DevBuf<double> d_counts, d_pgin;
    EstepTrie trie;
    DevBuf<int> d_bexp, d_eflags;
    DevBuf<unsigned char> d_escr;
    DevBuf<EstepScratch> d_edesc;
    DevBuf<int> d_pgexp;

    int64_t counters[HMC_CNT_N] = {};
    float kernel_ms[4] = {};

    int fail(int code, const char *fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

// ---------------------------------------------------------------------------------------------
// K0: genotype bit-packing, one thread per (individual, word)
__global__ void pack_kernel(const signed char *geno, int N, int L, int bits, int words, u32 *packed)
{
    const size_t t = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (t >= (size_t) N * words) return;
    const int ind = (int) (t / words), wd = (int) (t % words);
    const int per = 32 / (2 * bits);
    const signed char *g0 = geno + ((size_t) ind * 2) * L, *g1 = g0 + L;
    u32 v = 0;
    for (int q = 0; q < per; ++q) {
        const int l = wd * per + q;
        if (l < L) {
            const u32 c0 = (u32) (g0[l] + 1), c1 = (u32) (g1[l] + 1);
            v |= (c0 | (c1 << bits)) << (q * 2 * bits);
        }
    }
    packed[t] = v;
}

// the best resolution of every individual (row stride = K * 2 haplotype rows) in the packed panel layout
__global__ void pack_strided_kernel(const signed char *haps, long long row_stride, int N, int L, int bits, int words, u32 *packed)
{
    const size_t t = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (t >= (size_t) N * words) return;
    const int ind = (int) (t / words), wd = (int) (t % words);
    const int per = 32 / (2 * bits);
    const signed char *g0 = haps + (size_t) ind * row_stride * L, *g1 = g0 + L;
    u32 v = 0;
    for (int q = 0; q < per; ++q) {
        const int l = wd * per + q;
        if (l < L) v |= ((u32) (g0[l] + 1) | ((u32) (g1[l] + 1) << bits)) << (q * 2 * bits);
    }
    packed[t] = v;
}

// K0 for panels that arrive bit-packed (binary panel files): the inverse of pack_kernel into the unpacked staging copy that the
// miner and the comparison read, the missing-allele count per individual (work order) and a range check of the codes
__global__ void unpack_kernel(const u32 *packed, int N, int L, int bits, int words, int Amax, signed char *geno, int *miss, int *err)
{
    const size_t t = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (t >= (size_t) N * words) return;
    const int ind = (int) (t / words), wd = (int) (t % words);
    const int per = 32 / (2 * bits);
    const u32 m = (1u << bits) - 1u;
    signed char *g0 = geno + ((size_t) ind * 2) * L, *g1 = g0 + L;
    const u32 v = packed[t];
    int nm = 0;
    for (int q = 0; q < per; ++q) {
        const int l = wd * per + q;
        if (l >= L) break;
        const u32 c = v >> (q * 2 * bits);
        const int a = (int) (c & m) - 1, b = (int) ((c >> bits) & m) - 1;
        if (a >= Amax || b >= Amax) atomicMin(err, ind);
        g0[l] = (signed char) a; g1[l] = (signed char) b;
        nm += (a < 0) + (b < 0);
    }
    if (nm) atomicAdd(miss + ind, nm);
}

// K0: per-pattern consistency checks of an uploaded model and the last allele of every pattern, one thread per pattern;
// err[0] = smallest inconsistent pattern id (INT_MAX if none)
__global__ void model_check_kernel(int P, int L, const int *start, const int *end, const long long *al_off, long long total, const signed char *al,
                                   unsigned char *last, int *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const long long a0 = al_off[i], a1 = al_off[i + 1];
    const int s = start[i], e = end[i];
    if (a0 < 0 || a1 > total || a1 - a0 != (long long) e - s || e <= s || e > L || s < 0) { atomicMin(err, i); return; }
    last[i] = (unsigned char) al[a1 - 1];
}

// K0: successor table flattening, one thread per (pattern, allele)
__global__ void succ16_kernel(const int *succ, const double *tp, size_t n, int P, int4 *out, int *err)
{
    const size_t q = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (q >= n) return;
    int id = succ[q];
    if (id >= P || id < -1) { if (err) atomicMin(err + 1, (int) (q / (n / P))); id = -1; }
    const double t = id >= 0 ? tp[id] : 0.0;
    out[q] = make_int4(id, 0, __double2loint(t), __double2hiint(t));
}

// ---------------------------------------------------------------------------------------------
// Shared-memory carve-out (% of the SM's 228 KB) the forward kernel asks for: what MINB resident CTAs need.  The streamed K-best kernel asks
// for the same, since CTAs of two kernels only share an SM when their L1 / shared-memory split is the same.
static thread_local int t_fwd_carveout = 0;

template <int NT, int CC, int MINB, bool ESTEP>
static cudaError_t launch_forward_e(const FwdParams &p, int grid, cudaStream_t st, int *max_blocks_per_sm)
{
    auto kern = forward_kernel<NT, CC, MINB, ESTEP>;
    const size_t smem = sizeof(FwdSmem<NT, CC>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    t_fwd_carveout = (int) std::min<size_t>(100, (100 * MINB * (smem + 1024) + 228 * 1024 - 1) / (228 * 1024));
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, t_fwd_carveout);
    if (e != cudaSuccess) return e;
    if (max_blocks_per_sm) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, kern, NT, smem);
    kern<<<grid, NT, smem, st>>>(p);
    return cudaGetLastError();
}

template <int NT, int CC, int MINB>
static cudaError_t launch_forward_t(const FwdParams &p, int grid, cudaStream_t st, int *max_blocks_per_sm)
{
    return p.estep ? launch_forward_e<NT, CC, MINB, true>(p, grid, st, max_blocks_per_sm) : launch_forward_e<NT, CC, MINB, false>(p, grid, st, max_blocks_per_sm);
}

// threads per individual -> kernel configuration (candidate slots of the shared-memory fast path, resident CTAs per SM)
static cudaError_t launch_forward(int nt, const FwdParams &p, int grid, cudaStream_t st, int *occ)
{
#ifdef FWD_SWEEP
    if (const char *env = getenv("HMC_B200_FWD_CFG")) {     // kernel-configuration experiments
        switch (atoi(env)) {
            case 0: return launch_forward_t<256, 256, 3>(p, grid, st, occ);
            case 1: return launch_forward_t<256, 256, 4>(p, grid, st, occ);
            case 2: return launch_forward_t<256, 128, 4>(p, grid, st, occ);
            case 3: return launch_forward_t<256, 128, 3>(p, grid, st, occ);
            case 4: return launch_forward_t<128, 128, 5>(p, grid, st, occ);
            case 5: return launch_forward_t<128, 128, 4>(p, grid, st, occ);
            case 6: return launch_forward_t<384, 256, 2>(p, grid, st, occ);
            default: break;
        }
    }
#endif
    // measured on C2 (profiles/README.md, round 2): 256 threads per individual with 3-4 resident CTAs per SM is the best of the swept
    // configurations; throughput does not grow with more resident CTAs (the per-individual time grows in proportion)
    switch (nt) {
        case 32: return launch_forward_t<32, 128, 16>(p, grid, st, occ);
        case 64: return launch_forward_t<64, 128, 12>(p, grid, st, occ);
        case 128: return launch_forward_t<128, 256, 4>(p, grid, st, occ);
        case 1024: return launch_forward_t<1024, 256, 1>(p, grid, st, occ);     // the big-scratch wave: few CTAs, an SM each -- 32 warps instead of 8 on its giant steps
        default: return launch_forward_t<256, 256, 3>(p, grid, st, occ);
    }
}

// size / (re)build a scratch pool: one Scratch per persistent CTA.  The hash arrays start all-empty and every kernel leaves them so.
static int build_pool(hmc_ctx *ctx, ScratchPool &pl, int slots, int scap, int ccap, int hcap)
{
    if (pl.slots >= slots && pl.scap == scap && pl.ccap == ccap && pl.hcap == hcap) return HMC_OK;
    auto al = [](size_t x) { return (x + 255) & ~(size_t) 255; };
    const size_t b_keys = al((size_t) hcap * 8), b_u32 = al((size_t) hcap * 4), b_c = al((size_t) ccap * 4), b_m = al((size_t) scap * 8);
    const size_t per = 2 * b_keys + 3 * b_u32 + 2 * b_c + b_m;
    CK(pl.mem.ensure(per * slots));
    CK(pl.desc.ensure(slots));
    std::vector<Scratch> h(slots);
    for (int i = 0; i < slots; ++i) {
        unsigned char *b = pl.mem.p + per * i;
        Scratch &s = h[i];
        s.hcap = hcap; s.ccap = ccap; s.scap = scap;
        s.keys = (u64 *) b; b += b_keys;
        s.ord = (u32 *) b; b += b_u32;
        s.cnt = (u32 *) b; b += b_u32;
        s.tpv = (double *) b; b += b_keys;
        s.soff = (u32 *) b; b += b_u32;
        s.cand = (int *) b; b += b_c;
        s.sarr = (int *) b; b += b_c;
        s.mlist = (int2 *) b; b += b_m;
        CK(cudaMemsetAsync(s.keys, 0xFF, b_keys + b_u32, ctx->stream));      // keys = KEY_EMPTY, ord = none
        CK(cudaMemsetAsync(s.cnt, 0, b_u32, ctx->stream));
    }
    CK(cudaMemcpyAsync(pl.desc.p, h.data(), sizeof(Scratch) * slots, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    pl.slots = slots; pl.scap = scap; pl.ccap = ccap; pl.hcap = hcap;
    return HMC_OK;
}

static int ensure_arena(hmc_ctx *ctx)
{
    if (ctx->d_arena.p) return HMC_OK;
    size_t want = ctx->arena_req;
    if (want == 0) {
        size_t fr = 0, tot = 0;
        CK(cudaMemGetInfo(&fr, &tot));
        want = (size_t) (fr * 0.6);
        const char *env = getenv("HMC_B200_ARENA_MB");
        if (env) want = (size_t) atoll(env) << 20;
    }
    CK(ctx->d_arena.ensure(want / 8));
    return HMC_OK;
}

static int pick_threads(const hmc_ctx *ctx)
{
    const char *env = getenv("HMC_B200_THREADS");
    int nt = ctx->threads_opt > 0 ? ctx->threads_opt : (env ? atoi(env) : 0);
    if (nt == 0) nt = 256;
    if (nt != 32 && nt != 64 && nt != 128 && nt != 256) nt = 256;
    return nt;
}

// Runs the forward kernel (and, unless estep, the traceback) over individuals [first, first+count) in
// arena waves.  `after_wave`, when set, is called once per wave with the trellis still live (E-step).
template <typename AfterWave>
static int run_waves(hmc_ctx *ctx, int first, int count, int K, unsigned flags, int estep, AfterWave after_wave)
{
    if (ctx->N <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_set_panel has not been called");
    if (!ctx->have_model) return ctx->fail(HMC_ERR_INVALID, "hmc_set_model has not been called");
    if (first < 0 || count < 0 || first + count > ctx->N) return ctx->fail(HMC_ERR_INVALID, "individual range [%d,%d) outside the panel", first, first + count);
    if (K < 1 || K > HMC_MAX_K) return ctx->fail(HMC_ERR_LIMIT, "sample size %d outside [1,%d]", K, HMC_MAX_K);
    const int L = ctx->L;
    int rc = ensure_arena(ctx);
    if (rc) return rc;
    CK(ctx->d_cursor.ensure(2));         // [0] forward (and serial K-best), [1] streamed K-best
    CK(ctx->d_flags.ensure(4));          // [0] arena exhausted, [1] forward queue, [2] K-best queue, [3] done count
    CK(ctx->d_counters.ensure(72));
    ctx->ecache.valid = false;
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 72 * sizeof(long long), ctx->stream));
    memset(ctx->counters, 0, sizeof ctx->counters);
    memset(ctx->kernel_ms, 0, sizeof ctx->kernel_ms);

    const int nt = pick_threads(ctx);
    int occ = 1;
    FwdParams fp = {};
    fp.estep = estep;
    CK(launch_forward(nt, fp, 0, ctx->stream, &occ));
    if (occ < 1) occ = 1;
    const int max_slots = ctx->sms * occ;

    // wave capacity is bounded by the per-(individual, index) record table
    const size_t max_loci_bytes = (size_t) 6 << 30;
    int wave_cap = (int) std::min<size_t>((size_t) count, std::max<size_t>(1, max_loci_bytes / (loci_stride(L) * sizeof(LocusRec))));
    CK(ctx->d_loci.ensure((size_t) wave_cap * loci_stride(L)));
    CK(ctx->d_wave.ensure(wave_cap));
    CK(ctx->d_status.ensure(wave_cap));
    CK(ctx->d_final_exp.ensure(wave_cap));
    if (flags & HMC_FLAG_SCALED) CK(ctx->d_locus_exp.ensure((size_t) wave_cap * exp_stride(L)));

    // pending individuals, heaviest first: by the time the previous pass over the same panel spent on them (EM iterations and
    // the E-passes of an M-step resolve the same individuals again and again), else by missing genotypes
    std::vector<int> pending;
    pending.reserve(count);
    for (int q = 0; q < ctx->N; ++q) { const int ind = ctx->h_order[q]; if (ind >= first && ind < first + count) pending.push_back(ind); }
    CK(ctx->d_work.ensure(ctx->N));
    if ((int) ctx->h_work.size() != ctx->N) { ctx->h_work.assign(ctx->N, 0); CK(cudaMemsetAsync(ctx->d_work.p, 0, sizeof(long long) * ctx->N, ctx->stream)); }
    else std::stable_sort(pending.begin(), pending.end(), [&](int a, int b) { return ctx->h_work[a] > ctx->h_work[b]; });
    std::vector<int> big;
    std::vector<int> h_status;
    bool big_mode = false;
    int waves = 0, launches = 0;
    double fwd_share = 0.9;     // resolve mode: the part of the arena the forward kernel may fill; the rest is left to the K-best kernel's lists / nodes

    while (!pending.empty() || (!big_mode && !big.empty())) {
        if (pending.empty()) { big_mode = true; pending.swap(big); }
        // The arena is a bump allocator: everything a wave starts must fit, and what is still running when it fills up is lost work.  So a
        // wave takes only as many individuals as the footprint seen so far says will fit (the first wave of a panel starts blind).
        int n_wave = (int) std::min<size_t>(pending.size(), (size_t) wave_cap);
        {
            const double cap_units = (double) ctx->d_arena.n * (estep ? 1.0 : fwd_share);
            const double est = ctx->est_units[estep ? 1 : 0];
            if (est > 0 && !big_mode) n_wave = (int) std::max<double>(1.0, std::min<double>((double) n_wave, 0.97 * cap_units / est));
        }
        ScratchPool &pl = big_mode ? ctx->big_pool : ctx->pool;
        int slots;
        if (big_mode) { slots = std::min(n_wave, 16); rc = build_pool(ctx, pl, slots, 1 << 20, 1 << 22, 1 << 23); }
        // (131,072 states / 524,288 candidates per step and CTA, 35 MB of scratch each: an individual beyond that goes to the big-scratch wave, which
        // runs AFTER this one -- at 65,536 a single 71,425-state individual of an 8,000-individual cohort cost its rank a second wave of 39 ms)
        // Not more: an individual of several 100,000 states takes seconds on ONE 256-thread CTA that shares its SM, and the whole wave waits for it
        // (C4: 10,000 x 2,000 with individuals of up to 553,127 states ran 5.8 s with 262,144-state scratch, 3.1 s when they go to the big-scratch wave,
        // where 16 CTAs have an SM each).  HMC_B200_SCRATCH_SHIFT overrides (16..18).
        else {
            int sh = 17;
            if (const char *e = getenv("HMC_B200_SCRATCH_SHIFT")) sh = std::min(18, std::max(16, atoi(e)));
            slots = std::min(n_wave, max_slots); rc = build_pool(ctx, pl, slots, 1 << sh, 1 << (sh + 2), 1 << (sh + 3));
        }
        if (rc) return rc;
        ctx->h_wave.assign(pending.begin(), pending.begin() + n_wave);
        CK(cudaMemcpyAsync(ctx->d_wave.p, ctx->h_wave.data(), sizeof(int) * n_wave, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_cursor.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_flags.p, 0, 4 * sizeof(int), ctx->stream));
        CK(cudaMemsetAsync(ctx->d_status.p, 0, sizeof(int) * n_wave, ctx->stream));

        fp.panel.N = ctx->N; fp.panel.L = L; fp.panel.Amax = ctx->Amax; fp.panel.bits = ctx->bits; fp.panel.words = ctx->words;
        fp.panel.packed = ctx->d_packed.p; fp.panel.linfo = ctx->d_linfo.p; fp.panel.order = ctx->d_order.p;
        fp.model = ctx->dm;
        fp.arena.base = ctx->d_arena.p; fp.arena.capacity = estep ? (long long) ctx->d_arena.n : ((long long) ((double) ctx->d_arena.n * fwd_share) & ~15LL);
        fp.arena.cursor = ctx->d_cursor.p; fp.arena.exhausted = ctx->d_flags.p;
        fp.scratch = pl.desc.p;
        fp.loci = ctx->d_loci.p; fp.wave = ctx->d_wave.p; fp.n_wave = n_wave; fp.K = K; fp.estep = estep;
        fp.scaled = (flags & HMC_FLAG_SCALED) ? 1 : 0;
        fp.status = ctx->d_status.p; fp.final_exp = ctx->d_final_exp.p;
        fp.locus_exp = fp.scaled ? ctx->d_locus_exp.p : nullptr;
        fp.queue = ctx->d_flags.p + 1; fp.counters = ctx->d_counters.p;
        fp.work = ctx->d_work.p;
        fp.ind_cycles = nullptr;
        if (getenv("HMC_B200_DEBUG_CYCLES")) { CK(ctx->d_cycles.ensure((size_t) wave_cap * 8)); CK(cudaMemsetAsync(ctx->d_cycles.p, 0, sizeof(long long) * 8 * n_wave, ctx->stream)); fp.ind_cycles = ctx->d_cycles.p; }

        // K2 streamed: the K-best warps are resident beside the forward CTAs and take every individual as the forward kernel finishes it
        // (done list), so the enumeration hides behind the forward pass instead of following it
        const bool streamed = !estep && !getenv("HMC_B200_KBEST_SERIAL");
        fp.done_list = nullptr; fp.done_count = nullptr;
        if (streamed) {
            CK(ctx->d_done.ensure((size_t) wave_cap));
            CK(cudaMemsetAsync(ctx->d_done.p, 0xFF, sizeof(int) * n_wave, ctx->stream));
            fp.done_list = ctx->d_done.p; fp.done_count = ctx->d_flags.p + 3;
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        }
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        // (the wide configuration of the big-scratch wave is measured and parity-tested in resolve mode only: C4)
        CK(launch_forward(big_mode && !estep && !getenv("HMC_B200_BIG_NARROW") ? 1024 : nt, fp, slots, ctx->stream, nullptr));
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        ++launches;
        if (!estep) {
            // K2: lazy K-best enumeration + traceback, one warp per individual (persistent warps)
            // (streamed: at most 8 one-warp CTAs per SM, which even if they all became resident first would leave the forward kernel two CTAs per SM)
            int kb_warps = std::min(n_wave, ctx->sms * (streamed ? 8 : 16));
            if (const char *e = getenv("HMC_B200_KB_WARPS")) kb_warps = std::max(1, std::min(kb_warps, atoi(e)));
            const int kb_per = streamed ? 1 : KB_WARPS;
            const int kb_grid = (kb_warps + kb_per - 1) / kb_per;
            const size_t nw = (size_t) kb_grid * kb_per;
            CK(ctx->d_kb_eoff.ensure(nw * (L + 2))); CK(ctx->d_kb_chain.ensure(nw * (L + 2))); CK(ctx->d_kb_stack.ensure(nw * (L + 40)));
            if (!streamed) CK(cudaMemsetAsync(ctx->d_flags.p + 1, 0, sizeof(int), ctx->stream));        // the work queue is reused
            KbParams kp = {};
            kp.panel = fp.panel; kp.model = fp.model; kp.arena = ctx->d_arena.p; kp.capacity = (long long) ctx->d_arena.n; kp.cursor = ctx->d_cursor.p;
            kp.loci = ctx->d_loci.p; kp.wave = ctx->d_wave.p; kp.status = ctx->d_status.p; kp.final_exp = ctx->d_final_exp.p; kp.locus_exp = fp.locus_exp;
            kp.n_wave = n_wave; kp.K = K; kp.first = (flags & HMC_FLAG_PANEL_ROWS) ? 0 : first;
            kp.eoff = ctx->d_kb_eoff.p; kp.chain = ctx->d_kb_chain.p; kp.stack = ctx->d_kb_stack.p; kp.stack_stride = L + 40;
            kp.queue = ctx->d_flags.p + 1; kp.work = (unsigned long long *) (ctx->d_counters.p + 24);
            kp.pg = ctx->d_pg.p; kp.cov = ctx->d_cov.p; kp.prior = ctx->d_prior.p; kp.post = ctx->d_post.p; kp.exp2 = ctx->d_exp.p; kp.nres = ctx->d_nres.p;
            kp.haps = ctx->d_haps.p;
            if (streamed) {
                kp.queue = ctx->d_flags.p + 2; kp.done_list = ctx->d_done.p;
                kp.cursor = ctx->d_cursor.p + 1; kp.region_begin = fp.arena.capacity;      // its own part of the arena: the trellis is still growing
                CK(cudaFuncSetAttribute(kbest_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, t_fwd_carveout));
                CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
                kbest_kernel<1><<<kb_grid, 32, 0, ctx->stream2>>>(kp);
                CK(cudaGetLastError());
                CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
            }
            else kbest_kernel<KB_WARPS><<<kb_grid, KB_WARPS * 32, 0, ctx->stream>>>(kp);
            CK(cudaGetLastError());
            ++launches;
        }
        CK(cudaEventRecord(ctx->ev[2], ctx->stream));
        h_status.resize(n_wave);
        CK(cudaMemcpyAsync(h_status.data(), ctx->d_status.p, sizeof(int) * n_wave, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0, wave_fwd_ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->kernel_ms[0] += ms; wave_fwd_ms = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->kernel_ms[1] += ms;
        rc = after_wave(fp, n_wave, h_status, launches);
        if (rc) return rc;

        std::vector<int> again;
        int done = 0, kb_fail = 0;
        for (int q = 0; q < n_wave; ++q) {
            switch (h_status[q]) {
                case ST_DONE: ++done; break;
                case ST_NOMEM: case ST_PENDING: again.push_back(ctx->h_wave[q]); break;
                case ST_NOMEM_KBEST: again.push_back(ctx->h_wave[q]); ++kb_fail; break;
                case ST_TOO_BIG:
                    if (big_mode) return ctx->fail(HMC_ERR_LIMIT, "individual %d exceeds the per-locus state capacity", ctx->h_wave[q]);
                    big.push_back(ctx->h_wave[q]);
                    break;
                case ST_NO_HEAD: return ctx->fail(HMC_ERR_NO_HEAD, "individual %d: Can not find matching pattern!", ctx->h_wave[q]);
                default: return ctx->fail(HMC_ERR_CUDA, "individual %d: unexpected status %d", ctx->h_wave[q], h_status[q]);
            }
        }
        ++waves;
        {   // footprint estimate for the next wave: arena used / individuals that consumed it (the ones cut off were close to done: the
            // CTAs of a wave advance together), with a margin; heaviest individuals come first, so this leans high
            unsigned long long cur = 0;
            CK(cudaMemcpy(&cur, ctx->d_cursor.p, sizeof cur, cudaMemcpyDeviceToHost));
            const int consumers = done + kb_fail + std::min<int>((int) again.size() - kb_fail, slots);
            if (consumers > 0 && cur > 0 && !big_mode) {
                const double est = 1.06 * (double) cur / consumers;
                double &e = ctx->est_units[estep ? 1 : 0];
                e = (done + kb_fail == n_wave) ? std::max(est, 0.5 * e) : std::max(est, e);      // everything fitted: follow the estimate down slowly
            }
        }
        if (getenv("HMC_B200_PRINT_WAVES")) {
            int hist[8] = {0};
            for (int q = 0; q < n_wave; ++q) ++hist[h_status[q] & 7];
            unsigned long long cur = 0;
            cudaMemcpy(&cur, ctx->d_cursor.p, sizeof cur, cudaMemcpyDeviceToHost);
            fprintf(stderr, "wave %d (forward %.1f ms, %d CTAs%s): %d individuals -> pending %d done %d nomem %d too_big %d no_head %d nomem_kbest %d; arena cursor %.1f of %.1f GB (forward share %.2f)\n", waves,
                    wave_fwd_ms, slots, big_mode ? ", big scratch" : "", n_wave,
                    hist[ST_PENDING], hist[ST_DONE], hist[ST_NOMEM], hist[ST_TOO_BIG], hist[ST_NO_HEAD], hist[ST_NOMEM_KBEST], cur * 8 / 1e9, ctx->d_arena.n * 8 / 1e9, fwd_share);
        }
        // every individual of the wave ran out of arena: the first one had it to itself (the queue hands them out in order), so it can never fit
        if (kb_fail) fwd_share *= 0.5;      // the K-best kernel ran out of room behind the trellis: leave it more in the next wave
        if (done == 0 && !again.empty() && (int) again.size() == n_wave && n_wave == 1 && (kb_fail == 0 || fwd_share < 0.01))
            return ctx->fail(HMC_ERR_NOMEM, "trellis arena (%zu MB) cannot hold a single individual", ctx->d_arena.n * 8 >> 20);
        pending.erase(pending.begin(), pending.begin() + n_wave);
        pending.insert(pending.begin(), again.begin(), again.end());
    }
    long long hc[72];
    CK(cudaMemcpy(hc, ctx->d_counters.p, sizeof hc, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ctx->h_work.data(), ctx->d_work.p, sizeof(long long) * ctx->N, cudaMemcpyDeviceToHost));
    if (getenv("HMC_B200_PRINT_PASSES"))
        for (int b = 8; b < 24; b += 8)
            fprintf(stderr, "%s pass Mcycles: [0] %lld  [1] %lld  [2] %lld  [3] %lld  [4] %lld  [5] %lld  [7] %lld | [6] %lld\n", b == 8 ? "fast " : "spill",
                    hc[b] >> 20, hc[b + 1] >> 20, hc[b + 2] >> 20, hc[b + 3] >> 20, hc[b + 4] >> 20, hc[b + 5] >> 20, hc[b + 7] >> 20, hc[b + 6]);
#ifdef HMC_PROFILE_PASSES
    if (getenv("HMC_B200_PRINT_PASSES")) {
        long long tc = 0, tn = 0;
        for (int b = 0; b < 18; ++b) { tc += hc[32 + b]; tn += hc[50 + b]; }
        for (int b = 0; b < 18; ++b)
            if (hc[50 + b]) fprintf(stderr, "steps with < %6d candidates: %5.1f %% of the steps, %5.1f %% of the step cycles, %7.0f cycles per step\n", 1 << b, 100.0 * hc[50 + b] / tn,
                                    100.0 * hc[32 + b] / tc, (double) hc[32 + b] / hc[50 + b]);
    }
#endif
    ctx->seen_max_states = std::max<long long>(ctx->seen_max_states, hc[2]);
    ctx->counters[HMC_CNT_STEPS] = hc[0]; ctx->counters[HMC_CNT_STATES] = hc[1]; ctx->counters[HMC_CNT_MAX_STATES] = hc[2];
    ctx->counters[HMC_CNT_ARRIVALS] = hc[3]; ctx->counters[HMC_CNT_SPILL_STEPS] = hc[5]; ctx->counters[HMC_CNT_ESTEP_ENTRIES] = hc[7];
    ctx->counters[HMC_CNT_WAVES] = waves; ctx->counters[HMC_CNT_LAUNCHES] = launches;
    ctx->kb_work[0] = hc[24]; ctx->kb_work[1] = hc[25]; ctx->kb_work[2] = hc[26];
    if (getenv("HMC_B200_PRINT_KBEST")) fprintf(stderr, "kbest: %lld lazy-next visits, %lld nodes, %lld eager-list states\n", hc[24], hc[25], hc[26]);
#ifdef HMC_PROFILE_PASSES
    if (getenv("HMC_B200_PRINT_KBEST") && count > 0)
        fprintf(stderr, "kbest Mcycles per individual: P(G) + eager %.2f, enumeration %.2f, traceback %.2f; longest individual %.2f\n", hc[27] / 1e6 / count, hc[28] / 1e6 / count,
                hc[29] / 1e6 / count, hc[30] / 1e6);
#endif
    ctx->last_K = K; ctx->last_estep = estep;
    if (estep && waves == 1) {
        ctx->ecache.first = first; ctx->ecache.count = count; ctx->ecache.flags = flags; ctx->ecache.fp = fp; ctx->ecache.n_wave = fp.n_wave;
        memcpy(ctx->ecache.counters, ctx->counters, sizeof ctx->counters);
        ctx->ecache.valid = true;
    }
    return HMC_OK;
}

#include "mine_capi.inl"

// ---------------------------------------------------------------------------------------------
extern "C" {

const char *hmc_version(void) { return "hmc_b200 0.1 sm_100a"; }

int hmc_ctx_create(hmc_ctx **out, int device, size_t arena_bytes, int threads)
{
    if (!out) return HMC_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return HMC_ERR_NO_DEVICE;
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) return HMC_ERR_NO_DEVICE;
    if (pr.major != 10) return HMC_ERR_NO_DEVICE;      // the kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return HMC_ERR_NO_DEVICE;
    hmc_ctx *ctx = new hmc_ctx;
    ctx->device = device; ctx->sms = pr.multiProcessorCount; ctx->arena_req = arena_bytes; ctx->threads_opt = threads;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return HMC_ERR_CUDA; }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return HMC_ERR_CUDA; }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    *out = ctx;
    return HMC_OK;
}

void hmc_ctx_destroy(hmc_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->d_packed.release(); ctx->d_raw.release(); ctx->d_linfo.release(); ctx->d_order.release(); ctx->d_work.release(); ctx->d_succ16.release();
    ctx->d_succ.release(); ctx->d_start.release(); ctx->d_heads.release(); ctx->d_head_lut.release(); ctx->d_tp.release(); ctx->d_freq.release();
    ctx->d_last.release(); ctx->d_al_off.release(); ctx->d_al.release(); ctx->d_end_tmp.release(); ctx->d_model_err.release(); ctx->d_miss.release(); ctx->d_counts_snap.release();
    ctx->d_arena.release(); ctx->d_cursor.release(); ctx->d_flags.release(); ctx->d_counters.release(); ctx->d_loci.release();
    ctx->d_wave.release(); ctx->d_status.release(); ctx->d_final_exp.release(); ctx->d_locus_exp.release();
    ctx->pool.mem.release(); ctx->pool.desc.release(); ctx->big_pool.mem.release(); ctx->big_pool.desc.release();
    ctx->d_pg.release(); ctx->d_cov.release(); ctx->d_prior.release(); ctx->d_post.release(); ctx->d_exp.release(); ctx->d_nres.release(); ctx->d_haps.release(); ctx->d_lk.release(); ctx->d_emit_sym.release(); ctx->d_emit_wid.release(); ctx->d_emit_tmp.release(); ctx->d_emit_off.release(); ctx->d_emit_text.release();
    ctx->d_counts.release(); ctx->d_pgin.release(); ctx->d_pgexp.release();
    ctx->trie.release(); ctx->d_bexp.release(); ctx->d_eflags.release(); ctx->d_escr.release(); ctx->d_edesc.release(); ctx->d_cycles.release();
    ctx->d_kb_eoff.release(); ctx->d_kb_chain.release(); ctx->d_kb_stack.release();
    if (ctx->mine) { ctx->mine->release(); delete ctx->mine; }
    if (ctx->aut) { ctx->aut->release(); delete ctx->aut; }
    ctx->d_af.release(); ctx->d_best_packed.release(); ctx->d_cmp_real.release(); ctx->d_cmp_infer.release(); ctx->d_cmp_out.release();
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
}

const char *hmc_last_error(const hmc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int hmc_set_panel(hmc_ctx *ctx, int N, int L, int Amax, const int8_t *geno, const int32_t *n_alleles, const double *allele_freq)
{
    if (!ctx || N <= 0 || L <= 0 || !geno || !n_alleles || !allele_freq) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_set_panel: bad argument") : HMC_ERR_INVALID;
    if (Amax < 1 || Amax > HMC_MAX_ALLELES) return ctx->fail(HMC_ERR_LIMIT, "Amax %d outside [1,%d]", Amax, HMC_MAX_ALLELES);
    CK(cudaSetDevice(ctx->device));
    ctx->ecache.valid = false;
    ctx->have_model = false;        // the automaton tables (successor stride, head table) belong to the previous panel's allele tables
    ctx->r_count = 0;
    ctx->h_work.clear();
    ctx->est_units[0] = ctx->est_units[1] = 0; ctx->seen_max_states = 0;
    ctx->N = N; ctx->L = L; ctx->Amax = Amax;
    ctx->bits = Amax + 1 <= 4 ? 2 : (Amax + 1 <= 16 ? 4 : 8);
    const int per = 32 / (2 * ctx->bits);
    ctx->words = (L + per - 1) / per;
    // allele mask, work order (most missing genotypes first: they carry the largest state sets)
    std::vector<u32> amask(L, 0);
    ctx->h_nall.assign(n_alleles, n_alleles + L);
    for (int l = 0; l < L; ++l) {
        if (n_alleles[l] < 0 || n_alleles[l] > Amax) return ctx->fail(HMC_ERR_INVALID, "n_alleles[%d] = %d outside [0,%d]", l, n_alleles[l], Amax);
        for (int j = 0; j < n_alleles[l]; ++j) if (allele_freq[(size_t) l * Amax + j] > 0) amask[l] |= 1u << j;
        amask[l] = (u32) n_alleles[l] | (amask[l] << 8);
    }
    std::vector<int> miss(N, 0);
    size_t tot_miss = 0;
    for (int i = 0; i < N; ++i) {
        const int8_t *g0 = geno + ((size_t) i * 2) * L, *g1 = g0 + L;
        int m = 0;
        for (int l = 0; l < L; ++l) {
            if (g0[l] >= Amax || g1[l] >= Amax) return ctx->fail(HMC_ERR_INVALID, "allele index out of range at individual %d locus %d", i, l);
            m += (g0[l] < 0) + (g1[l] < 0);
        }
        miss[i] = m; tot_miss += m;
    }
    ctx->missing_frac = (double) tot_miss / ((double) N * L * 2);
    ctx->h_order.resize(N);
    for (int i = 0; i < N; ++i) ctx->h_order[i] = i;
    std::stable_sort(ctx->h_order.begin(), ctx->h_order.end(), [&](int a, int b) { return miss[a] > miss[b]; });
    CK(ctx->d_packed.ensure((size_t) N * ctx->words));
    CK(ctx->d_linfo.ensure(L));
    CK(ctx->d_order.ensure(N));
    DevBuf<signed char> &raw = ctx->d_raw;
    CK(raw.ensure((size_t) N * 2 * L));
    CK(cudaMemcpyAsync(raw.p, geno, (size_t) N * 2 * L, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_linfo.p, amask.data(), sizeof(u32) * L, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_af.ensure((size_t) L * Amax));
    CK(cudaMemcpyAsync(ctx->d_af.p, allele_freq, sizeof(double) * (size_t) L * Amax, cudaMemcpyHostToDevice, ctx->stream));
    ++ctx->panel_serial;
    CK(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.data(), sizeof(int) * N, cudaMemcpyHostToDevice, ctx->stream));
    const size_t nthr = (size_t) N * ctx->words;
    pack_kernel<<<(unsigned) ((nthr + 255) / 256), 256, 0, ctx->stream>>>(raw.p, N, L, ctx->bits, ctx->words, ctx->d_packed.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return HMC_OK;
}

int hmc_set_panel_packed(hmc_ctx *ctx, int N, int L, int Amax, int bits, const uint32_t *packed, const int32_t *n_alleles, const double *allele_freq)
{
    if (!ctx || N <= 0 || L <= 0 || !packed || !n_alleles || !allele_freq) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_set_panel_packed: bad argument") : HMC_ERR_INVALID;
    if (Amax < 1 || Amax > HMC_MAX_ALLELES) return ctx->fail(HMC_ERR_LIMIT, "Amax %d outside [1,%d]", Amax, HMC_MAX_ALLELES);
    if (bits != (Amax + 1 <= 4 ? 2 : (Amax + 1 <= 16 ? 4 : 8))) return ctx->fail(HMC_ERR_INVALID, "hmc_set_panel_packed: %d bits per allele do not match Amax %d", bits, Amax);
    CK(cudaSetDevice(ctx->device));
    ctx->ecache.valid = false;
    ctx->have_model = false;
    ctx->r_count = 0;
    ctx->h_work.clear();
    ctx->est_units[0] = ctx->est_units[1] = 0; ctx->seen_max_states = 0;
    ctx->N = N; ctx->L = L; ctx->Amax = Amax; ctx->bits = bits;
    const int per = 32 / (2 * bits);
    ctx->words = (L + per - 1) / per;
    std::vector<u32> amask(L, 0);
    ctx->h_nall.assign(n_alleles, n_alleles + L);
    for (int l = 0; l < L; ++l) {
        if (n_alleles[l] < 0 || n_alleles[l] > Amax) return ctx->fail(HMC_ERR_INVALID, "n_alleles[%d] = %d outside [0,%d]", l, n_alleles[l], Amax);
        for (int j = 0; j < n_alleles[l]; ++j) if (allele_freq[(size_t) l * Amax + j] > 0) amask[l] |= 1u << j;
        amask[l] = (u32) n_alleles[l] | (amask[l] << 8);
    }
    cudaStream_t st = ctx->stream;
    CK(ctx->d_packed.ensure((size_t) N * ctx->words)); CK(ctx->d_linfo.ensure(L)); CK(ctx->d_order.ensure(N)); CK(ctx->d_raw.ensure((size_t) N * 2 * L));
    CK(ctx->d_af.ensure((size_t) L * Amax)); CK(ctx->d_miss.ensure((size_t) N + 1));
    CK(cudaMemcpyAsync(ctx->d_packed.p, packed, sizeof(u32) * (size_t) N * ctx->words, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_linfo.p, amask.data(), sizeof(u32) * L, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_af.p, allele_freq, sizeof(double) * (size_t) L * Amax, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->d_miss.p, 0, sizeof(int) * N, st));
    const int big = INT_MAX;
    CK(cudaMemcpyAsync(ctx->d_miss.p + N, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    const size_t nthr = (size_t) N * ctx->words;
    unpack_kernel<<<(unsigned) ((nthr + 255) / 256), 256, 0, st>>>(ctx->d_packed.p, N, L, bits, ctx->words, Amax, ctx->d_raw.p, ctx->d_miss.p, ctx->d_miss.p + N);
    CK(cudaGetLastError());
    std::vector<int> miss((size_t) N + 1);
    CK(cudaMemcpyAsync(miss.data(), ctx->d_miss.p, sizeof(int) * ((size_t) N + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (miss[N] != INT_MAX) return ctx->fail(HMC_ERR_INVALID, "allele index out of range at individual %d", miss[N]);
    size_t tot_miss = 0;
    for (int i = 0; i < N; ++i) tot_miss += miss[i];
    ctx->missing_frac = (double) tot_miss / ((double) N * L * 2);
    ctx->h_order.resize(N);
    for (int i = 0; i < N; ++i) ctx->h_order[i] = i;
    std::stable_sort(ctx->h_order.begin(), ctx->h_order.end(), [&](int a, int b) { return miss[a] > miss[b]; });
    CK(cudaMemcpyAsync(ctx->d_order.p, ctx->h_order.data(), sizeof(int) * N, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    ++ctx->panel_serial;
    return HMC_OK;
}

int hmc_set_model(hmc_ctx *ctx, int P, int head_len, const int32_t *start, const int32_t *end, const double *freq, const double *tp,
                  const int32_t *succ, const int64_t *al_off, const int8_t *alleles, int n_heads, const int32_t *heads)
{
    if (!ctx || P <= 0 || !start || !end || !freq || !tp || !succ || !al_off || !alleles || !heads) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_set_model: bad argument") : HMC_ERR_INVALID;
    if (ctx->N <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_set_model before hmc_set_panel");
    if (head_len < 1 || head_len > 8 || head_len > ctx->L) return ctx->fail(HMC_ERR_LIMIT, "head length %d outside [1,8]", head_len);
    CK(cudaSetDevice(ctx->device));
    ctx->ecache.valid = false;
    const int A = ctx->Amax;
    // per-pattern checks, the last alleles and the successor range check run on the device after the upload (model_check_kernel,
    // succ16_kernel): host loops over millions of patterns would cost more than the copies
    if (al_off[0] != 0 || al_off[P] < P) return ctx->fail(HMC_ERR_INVALID, "pattern allele offsets are inconsistent");
    // head lookup: mixed-radix code over the first head_len loci, locus 0 most significant
    int stride[8] = {0};
    long long size = 1;
    for (int j = head_len - 1; j >= 0; --j) { stride[j] = (int) size; size *= std::max(1, ctx->h_nall[j]); if (size > (1 << 20)) return ctx->fail(HMC_ERR_LIMIT, "head table too large"); }
    std::vector<int> lut((size_t) size, -1);
    for (int h = 0; h < n_heads; ++h) {
        const int id = heads[h];
        if (id < 0 || id >= P || start[id] != 0 || end[id] != head_len || al_off[id] < 0 || al_off[id] + head_len > al_off[P])
            return ctx->fail(HMC_ERR_INVALID, "head %d is not a start-0 pattern of length head_len", id);
        long long code = 0;
        for (int j = 0; j < head_len; ++j) {
            const int a = alleles[al_off[id] + j];
            if (a < 0 || a >= std::max(1, ctx->h_nall[j])) return ctx->fail(HMC_ERR_INVALID, "head %d: allele %d at locus %d outside [0,%d)", id, a, j, ctx->h_nall[j]);
            code += (long long) a * stride[j];
        }
        lut[(size_t) code] = id;
    }
    CK(ctx->d_succ16.ensure((size_t) P * A));
    CK(ctx->d_succ.ensure((size_t) P * A)); CK(ctx->d_start.ensure(P)); CK(ctx->d_tp.ensure(P)); CK(ctx->d_freq.ensure(P)); CK(ctx->d_last.ensure(P));
    CK(ctx->d_al_off.ensure(P + 1)); CK(ctx->d_al.ensure((size_t) al_off[P] + 1)); CK(ctx->d_heads.ensure(std::max(1, n_heads))); CK(ctx->d_head_lut.ensure(lut.size()));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(ctx->d_succ.p, succ, sizeof(int) * (size_t) P * A, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_start.p, start, sizeof(int) * P, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_tp.p, tp, sizeof(double) * P, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_freq.p, freq, sizeof(double) * P, cudaMemcpyHostToDevice, st));
    CK(ctx->d_end_tmp.ensure(P)); CK(ctx->d_model_err.ensure(2));
    CK(cudaMemcpyAsync(ctx->d_end_tmp.p, end, sizeof(int) * P, cudaMemcpyHostToDevice, st));
    const int err_init[2] = {INT_MAX, INT_MAX};
    CK(cudaMemcpyAsync(ctx->d_model_err.p, err_init, sizeof err_init, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_al_off.p, al_off, sizeof(long long) * (P + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_al.p, alleles, (size_t) al_off[P], cudaMemcpyHostToDevice, st));
    if (n_heads > 0) CK(cudaMemcpyAsync(ctx->d_heads.p, heads, sizeof(int) * n_heads, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_head_lut.p, lut.data(), sizeof(int) * lut.size(), cudaMemcpyHostToDevice, st));
    // 16-byte successor entries {successor id, pad, transition_prob(successor)}: one 128-bit gather yields both (K0 flatten)
    {
        const size_t n = (size_t) P * A;
        model_check_kernel<<<(unsigned) ((P + 255) / 256), 256, 0, st>>>(P, ctx->L, ctx->d_start.p, ctx->d_end_tmp.p, ctx->d_al_off.p, (long long) al_off[P], ctx->d_al.p,
                                                                         ctx->d_last.p, ctx->d_model_err.p);
        succ16_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(ctx->d_succ.p, ctx->d_tp.p, n, P, ctx->d_succ16.p, ctx->d_model_err.p);
        CK(cudaGetLastError());
    }
    int err[2] = {INT_MAX, INT_MAX};
    CK(cudaMemcpyAsync(err, ctx->d_model_err.p, sizeof err, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (err[0] != INT_MAX || err[1] != INT_MAX) {
        ctx->have_model = false;
        if (err[0] != INT_MAX) return ctx->fail(HMC_ERR_INVALID, "pattern %d: inconsistent start/end/alleles", err[0]);
        return ctx->fail(HMC_ERR_INVALID, "successor id out of range (pattern %d)", err[1]);
    }
    DevModel &m = ctx->dm;
    m.P = P; m.head_len = head_len; m.n_heads = n_heads; m.Amax = A;
    m.succ = ctx->d_succ.p; m.succ16 = ctx->d_succ16.p; m.tp = ctx->d_tp.p; m.freq = ctx->d_freq.p; m.start = ctx->d_start.p; m.last = ctx->d_last.p;
    m.al_off = ctx->d_al_off.p; m.al = ctx->d_al.p; m.heads = ctx->d_heads.p; m.head_lut = ctx->d_head_lut.p;
    for (int j = 0; j < 8; ++j) m.head_stride[j] = stride[j];
    ctx->have_model = true;
    return HMC_OK;
}

int hmc_resolve_all_device(hmc_ctx *ctx, int first, int count, int K, unsigned flags)
{
    if (!ctx) return HMC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (count <= 0 || K < 1 || K > HMC_MAX_K) return ctx->fail(HMC_ERR_INVALID, "hmc_resolve_all: bad count/K");
    const int L = ctx->L;
    const bool panel_rows = (flags & HMC_FLAG_PANEL_ROWS) != 0;
    const size_t rows = panel_rows ? (size_t) ctx->N : (size_t) count;
    CK(ctx->d_pg.ensure(rows)); CK(ctx->d_cov.ensure(rows)); CK(ctx->d_exp.ensure(rows)); CK(ctx->d_nres.ensure(rows));
    CK(ctx->d_prior.ensure(rows * K)); CK(ctx->d_post.ensure(rows * K));
    CK(ctx->d_haps.ensure(rows * K * 2 * L));
    ctx->r_count = 0;           // fetch / compare / mine_resolved must not see partial results of a failed call
    ctx->r_first = panel_rows ? 0 : first; ctx->r_K = K;
    const int rc = run_waves(ctx, first, count, K, flags, 0, [](const FwdParams &, int, const std::vector<int> &, int &) { return 0; });
    if (rc == HMC_OK) ctx->r_count = (int) rows;
    return rc;
}

int hmc_fetch_results(hmc_ctx *ctx, double *out_pg, int32_t *out_exp, double *out_cov, int32_t *out_nres, double *out_prior, double *out_post,
                      int8_t *out_haps, double *out_ll)
{
    if (!ctx || ctx->r_count <= 0) return ctx ? ctx->fail(HMC_ERR_INVALID, "no results to fetch") : HMC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int n = ctx->r_count, K = ctx->r_K;
    cudaStream_t st = ctx->stream;
    std::vector<double> pg;
    std::vector<int> ex;
    if (out_ll && !out_pg) { pg.resize(n); out_pg = pg.data(); }
    if (out_ll && !out_exp) { ex.resize(n); out_exp = ex.data(); }
    if (out_pg) CK(cudaMemcpyAsync(out_pg, ctx->d_pg.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (out_exp) CK(cudaMemcpyAsync(out_exp, ctx->d_exp.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (out_cov) CK(cudaMemcpyAsync(out_cov, ctx->d_cov.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (out_nres) CK(cudaMemcpyAsync(out_nres, ctx->d_nres.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (out_prior) CK(cudaMemcpyAsync(out_prior, ctx->d_prior.p, sizeof(double) * (size_t) n * K, cudaMemcpyDeviceToHost, st));
    if (out_post) CK(cudaMemcpyAsync(out_post, ctx->d_post.p, sizeof(double) * (size_t) n * K, cudaMemcpyDeviceToHost, st));
    if (out_haps) CK(cudaMemcpyAsync(out_haps, ctx->d_haps.p, (size_t) n * K * 2 * ctx->L, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (out_ll) {   // HaploModel.cpp:110, in individual order
        double ll = 0;
        for (int i = 0; i < n; ++i) ll += out_exp[i] ? (std::log(out_pg[i]) - out_exp[i] * M_LN2) : std::log(out_pg[i]);
        *out_ll = ll;
    }
    return HMC_OK;
}

int hmc_result_device_ptrs(hmc_ctx *ctx, void *ptrs[7], int64_t row_bytes[7])
{
    if (!ctx || !ptrs || !row_bytes) return HMC_ERR_INVALID;
    if (ctx->r_count <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_result_device_ptrs: no resolve results on the device");
    const int64_t K = ctx->r_K, L = ctx->L;
    ptrs[0] = ctx->d_pg.p; row_bytes[0] = 8;
    ptrs[1] = ctx->d_exp.p; row_bytes[1] = 4;
    ptrs[2] = ctx->d_cov.p; row_bytes[2] = 8;
    ptrs[3] = ctx->d_nres.p; row_bytes[3] = 4;
    ptrs[4] = ctx->d_prior.p; row_bytes[4] = 8 * K;
    ptrs[5] = ctx->d_post.p; row_bytes[5] = 8 * K;
    ptrs[6] = ctx->d_haps.p; row_bytes[6] = K * 2 * L;
    return HMC_OK;
}

int hmc_fetch_counts(hmc_ctx *ctx, int Pc, double *out_freq, double *out_prefix)
{
    if (!ctx || Pc <= 0) return HMC_ERR_INVALID;
    if (ctx->d_counts.n < (size_t) 2 * Pc) return ctx->fail(HMC_ERR_INVALID, "hmc_fetch_counts: no count vector of %d candidates on the device", Pc);
    CK(cudaSetDevice(ctx->device));
    if (out_freq) CK(cudaMemcpyAsync(out_freq, ctx->d_counts.p, sizeof(double) * Pc, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_prefix) CK(cudaMemcpyAsync(out_prefix, ctx->d_counts.p + Pc, sizeof(double) * Pc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HMC_OK;
}

void *hmc_stream(hmc_ctx *ctx) { return ctx ? (void *) ctx->stream : nullptr; }

int hmc_resolve_all(hmc_ctx *ctx, int first, int count, int K, unsigned flags, double *out_pg, int32_t *out_exp, double *out_cov,
                    int32_t *out_nres, double *out_prior, double *out_post, int8_t *out_haps, double *out_ll)
{
    if (flags & HMC_FLAG_PANEL_ROWS) return ctx ? ctx->fail(HMC_ERR_INVALID, "HMC_FLAG_PANEL_ROWS belongs to hmc_resolve_all_device") : HMC_ERR_INVALID;
    int rc = hmc_resolve_all_device(ctx, first, count, K, flags);
    if (rc) return rc;
    return hmc_fetch_results(ctx, out_pg, out_exp, out_cov, out_nres, out_prior, out_post, out_haps, out_ll);
}

void hmc_normalise_counts(int Pc, int N, double *freq, double *prefix, double *tp)
{
    for (int i = 0; i < Pc; ++i) {      // HaploBuilder.cpp:317-331
        double f = std::min(freq[i], (double) N);
        const double pf = std::min(prefix[i], (double) N);
        f = std::min(f, pf);
        freq[i] = f / N;
        prefix[i] = pf / N;
        const double t = pf > 0 ? f / pf : f / N;
        if (tp) tp[i] = t < 1.0 ? t : 1.0;      // HaploPattern::setTransitionProb clamp, HaploPattern.h:47
    }
}

int hmc_mine_patterns(hmc_ctx *ctx, int n_samples, const int8_t *sample_haps, const double *sample_weights, double total_weight,
                      double min_freq, const int32_t *min_len, const int32_t *max_len, int64_t *out_P, int64_t *out_n_alleles)
{
    if (!ctx || !min_len || !max_len || !out_P || !out_n_alleles || n_samples < 0 || (n_samples > 0 && (!sample_haps || !sample_weights)))
        return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_mine_patterns: bad argument") : HMC_ERR_INVALID;
    if (ctx->N <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_set_panel has not been called");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->mine) ctx->mine = new MinePool;
    const int rc = run_mine(ctx, *ctx->mine, n_samples, sample_haps, nullptr, sample_weights, total_weight, min_freq, min_len, max_len);
    if (rc) return rc;
    *out_P = ctx->mine->P; *out_n_alleles = ctx->mine->n_alleles;
    return HMC_OK;
}

int hmc_mine_patterns_resolved(hmc_ctx *ctx, int n_samples, const int32_t *rows, const double *sample_weights, double total_weight,
                               double min_freq, const int32_t *min_len, const int32_t *max_len, int64_t *out_P, int64_t *out_n_alleles)
{
    if (!ctx || !min_len || !max_len || !out_P || !out_n_alleles || n_samples <= 0 || !rows || !sample_weights)
        return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_mine_patterns_resolved: bad argument") : HMC_ERR_INVALID;
    if (ctx->N <= 0) return ctx->fail(HMC_ERR_INVALID, "hmc_set_panel has not been called");
    if (ctx->r_count <= 0 || !ctx->d_haps.p) return ctx->fail(HMC_ERR_INVALID, "hmc_mine_patterns_resolved: no resolve results on the device");
    const long long n_rows = (long long) ctx->r_count * ctx->r_K * 2;
    for (int i = 0; i < n_samples; ++i) if (rows[i] < 0 || rows[i] >= n_rows) return ctx->fail(HMC_ERR_INVALID, "sample row %d outside the %lld result rows", rows[i], n_rows);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->mine) ctx->mine = new MinePool;
    const int rc = run_mine(ctx, *ctx->mine, n_samples, nullptr, rows, sample_weights, total_weight, min_freq, min_len, max_len);
    if (rc) return rc;
    *out_P = ctx->mine->P; *out_n_alleles = ctx->mine->n_alleles;
    return HMC_OK;
}

int hmc_result_dims(hmc_ctx *ctx, int32_t *count, int32_t *K)
{
    if (!ctx) return HMC_ERR_INVALID;
    if (count) *count = ctx->r_count;
    if (K) *K = ctx->r_K;
    return HMC_OK;
}

int hmc_fetch_best_haps(hmc_ctx *ctx, int8_t *out_best)
{
    if (!ctx || !out_best || ctx->r_count <= 0) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_fetch_best_haps: no results to fetch") : HMC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const size_t row = (size_t) 2 * ctx->L;
    CK(cudaMemcpy2DAsync(out_best, row, ctx->d_haps.p, row * ctx->r_K, row, ctx->r_count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HMC_OK;
}

int hmc_fetch_best_packed(hmc_ctx *ctx, uint32_t *out_packed)
{
    if (!ctx || !out_packed || ctx->r_count <= 0) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_fetch_best_packed: no results to fetch") : HMC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t) ctx->r_count * ctx->words;
    CK(ctx->d_best_packed.ensure(n));
    pack_strided_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_haps.p, (long long) ctx->r_K * 2, ctx->r_count, ctx->L, ctx->bits, ctx->words,
                                                                              ctx->d_best_packed.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_packed, ctx->d_best_packed.p, sizeof(u32) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HMC_OK;
}

int hmc_fetch_patterns(hmc_ctx *ctx, int32_t *start, int32_t *end, double *freq, double *prefix_freq, double *tp, int64_t *al_off, int8_t *alleles)
{
    if (!ctx || !start || !end || !freq || !prefix_freq || !tp || !al_off || !alleles) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_fetch_patterns: bad argument") : HMC_ERR_INVALID;
    if (!ctx->mine) return ctx->fail(HMC_ERR_INVALID, "hmc_fetch_patterns before hmc_mine_patterns");
    MinePool &mp = *ctx->mine;
    if (mp.handed_over) return ctx->fail(HMC_ERR_INVALID, "hmc_fetch_patterns: the list was handed to hmc_build_automaton; fetch before building");
    CK(cudaSetDevice(ctx->device));
    const size_t P = (size_t) mp.P;
    if (P == 0) { al_off[0] = 0; return HMC_OK; }
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(start, mp.o_start.p, 4 * P, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(end, mp.o_end.p, 4 * P, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(freq, mp.o_freq.p, 8 * P, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(prefix_freq, mp.o_prefix.p, 8 * P, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(tp, mp.o_tp.p, 8 * P, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(al_off, mp.o_aloff.p, 8 * (P + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(alleles, mp.o_al.p, (size_t) mp.n_alleles, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return HMC_OK;
}

int hmc_map_patterns(hmc_ctx *ctx, const hmc_pattern_rec **records, const int8_t **alleles)
{
    if (!ctx || !records || !alleles) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_map_patterns: bad argument") : HMC_ERR_INVALID;
    if (!ctx->mine) return ctx->fail(HMC_ERR_INVALID, "hmc_map_patterns before hmc_mine_patterns");
    MinePool &mp = *ctx->mine;
    CK(cudaSetDevice(ctx->device));
    const size_t rb = ((size_t) mp.P * sizeof(hmc_pattern_rec) + 255) & ~(size_t) 255, total = rb + (size_t) mp.n_alleles + 256;
    if (total > mp.pin_bytes) {
        if (mp.pin) cudaFreeHost(mp.pin);
        mp.pin = nullptr; mp.pin_bytes = 0;
        const size_t cap = total + total / 2;
        CK(cudaHostAlloc(&mp.pin, cap, cudaHostAllocDefault));
        mp.pin_bytes = cap;
    }
    if (!mp.pin_valid && mp.P > 0) {
        if (mp.handed_over) return ctx->fail(HMC_ERR_INVALID, "hmc_map_patterns: the list was handed to hmc_build_automaton; map before building");
        CK(cudaMemcpyAsync(mp.pin, mp.o_rec.p, (size_t) mp.P * sizeof(hmc_pattern_rec), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync((char *) mp.pin + rb, mp.o_al.p, (size_t) mp.n_alleles, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        mp.pin_valid = true;
    }
    *records = (const hmc_pattern_rec *) mp.pin;
    *alleles = (const int8_t *) ((char *) mp.pin + rb);
    return HMC_OK;
}

int hmc_build_automaton(hmc_ctx *ctx, int head_len, int32_t *out_n_heads)
{
    if (!ctx) return HMC_ERR_INVALID;
    if (!ctx->mine) return ctx->fail(HMC_ERR_INVALID, "hmc_build_automaton before hmc_mine_patterns");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->aut) ctx->aut = new AutPool;
    return run_build_automaton(ctx, *ctx->mine, *ctx->aut, head_len, out_n_heads);
}

int hmc_fetch_automaton(hmc_ctx *ctx, int32_t *succ, int32_t *heads)
{
    if (!ctx || !ctx->have_model) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_fetch_automaton: no model on the device") : HMC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (succ) CK(cudaMemcpyAsync(succ, ctx->dm.succ, sizeof(int) * (size_t) ctx->dm.P * ctx->dm.Amax, cudaMemcpyDeviceToHost, ctx->stream));
    if (heads && ctx->dm.n_heads > 0) CK(cudaMemcpyAsync(heads, ctx->dm.heads, sizeof(int) * ctx->dm.n_heads, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return HMC_OK;
}

int hmc_get_mine_stats(hmc_ctx *ctx, double *ms, int64_t *nodes, int64_t *levels, int64_t *max_list, int64_t *swept, int64_t *written)
{
    if (!ctx || !ctx->mine) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_get_mine_stats before hmc_mine_patterns") : HMC_ERR_INVALID;
    if (ms) *ms = ctx->mine->ms + (ctx->aut ? ctx->aut->ms : 0.0f);
    if (nodes) *nodes = ctx->mine->nodes;
    if (levels) *levels = ctx->mine->levels;
    if (max_list) *max_list = ctx->mine->max_list;
    if (swept) *swept = ctx->mine->swept;
    if (written) *written = ctx->mine->written;
    return HMC_OK;
}

int hmc_compare(hmc_ctx *ctx, int N, int L, const int8_t *real, const int8_t *infer, int64_t out[6], int32_t *bad_individual, int32_t *bad_locus)
{
    if (!ctx || !real || !out || N <= 0 || L <= 0) return ctx ? ctx->fail(HMC_ERR_INVALID, "hmc_compare: bad argument") : HMC_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t bytes = (size_t) N * 2 * L;
    CK(ctx->d_cmp_real.ensure(bytes)); CK(ctx->d_cmp_out.ensure(1));
    CK(cudaMemcpyAsync(ctx->d_cmp_real.p, real, bytes, cudaMemcpyHostToDevice, st));
    const signed char *inf = nullptr;
    long long inf_stride = 2;
    if (infer) {
        CK(ctx->d_cmp_infer.ensure(bytes));
        CK(cudaMemcpyAsync(ctx->d_cmp_infer.p, infer, bytes, cudaMemcpyHostToDevice, st));
        inf = ctx->d_cmp_infer.p;
    }
    else {      // the best resolutions of the last resolve call, where they lie: row (i * K + 0) * 2 + h
        if (ctx->r_count != N || ctx->L != L || !ctx->d_haps.p) return ctx->fail(HMC_ERR_INVALID, "hmc_compare: no resolve results for %d x %d on the device", N, L);
        inf = ctx->d_haps.p; inf_stride = 2ll * ctx->r_K;
    }
    CompareOut h = {};
    h.bad = ~0ull;
    CK(cudaMemcpyAsync(ctx->d_cmp_out.p, &h, sizeof h, cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned) std::min<long long>(((long long) N * 32 + 255) / 256, (long long) ctx->sms * 16);
    compare_kernel<<<grid, 256, 0, st>>>(ctx->d_cmp_real.p, 2, inf, inf_stride, N, L, ctx->d_cmp_out.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h, ctx->d_cmp_out.p, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out[0] = h.se_num; out[1] = h.se_den; out[2] = h.ig_num; out[3] = h.ig_den; out[4] = h.ih_num; out[5] = h.ih_den;
    const int bi = h.bad == ~0ull ? -1 : (int) (h.bad >> 32), bl = h.bad == ~0ull ? -1 : (int) (h.bad & 0xFFFFFFFFu);
    if (bad_individual) *bad_individual = bi;
    if (bad_locus) *bad_locus = bl;
    if (bi >= 0) return ctx->fail(HMC_ERR_INVALID, "Inconsistent genotypes at locus %d! (individual %d)", bl, bi);
    return HMC_OK;
}

int hmc_get_counters(hmc_ctx *ctx, int64_t out[HMC_CNT_N])
{
    if (!ctx || !out) return HMC_ERR_INVALID;
    memcpy(out, ctx->counters, sizeof ctx->counters);
    return HMC_OK;
}

// diagnostics (HMC_B200_DEBUG_CYCLES=1): per individual of the last wave, in wave order, 8 values {SM cycles, states, max states,
// arrivals, spill steps, start ns, end ns, CTA}; `out` holds 8 * cap values; returns the count
int hmc_debug_cycles(hmc_ctx *ctx, long long *out, int *inds, int cap)
{
    if (!ctx || !ctx->d_cycles.p) return 0;
    const int n = std::min<int>(cap, (int) ctx->h_wave.size());
    cudaMemcpy(out, ctx->d_cycles.p, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) inds[i] = ctx->h_wave[i];
    return n;
}

int hmc_get_kernel_ms(hmc_ctx *ctx, float out[4])
{
    if (!ctx || !out) return HMC_ERR_INVALID;
    memcpy(out, ctx->kernel_ms, sizeof ctx->kernel_ms);
    return HMC_OK;
}

int64_t hmc_get_trellis(hmc_ctx *ctx, int individual, int32_t *counts, int64_t capacity, int32_t *ida, int32_t *idb, double *alpha, double *beta)
{
    if (!ctx || !counts) return HMC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    int w = -1;
    for (size_t q = 0; q < ctx->h_wave.size(); ++q) if (ctx->h_wave[q] == individual) { w = (int) q; break; }
    if (w < 0) return ctx->fail(HMC_ERR_INVALID, "individual %d is not in the last wave", individual);
    const int L = ctx->L;
    std::vector<LocusRec> rec(L + 1);
    if (cudaMemcpy(rec.data(), ctx->d_loci.p + (size_t) w * loci_stride(L), sizeof(LocusRec) * (L + 1), cudaMemcpyDeviceToHost) != cudaSuccess)
        return ctx->fail(HMC_ERR_CUDA, "trellis record copy failed");
    int64_t tot = 0;
    for (int l = 0; l <= L; ++l) {
        const int S = (l < ctx->dm.head_len) ? 0 : rec[l].count;
        counts[l] = S;
        if (S <= 0) continue;
        if (tot + S > capacity) return ctx->fail(HMC_ERR_INVALID, "trellis capacity too small");
        const BlockLayout b = block_layout(S, 0, ctx->last_estep);
        const u64 *blk = ctx->d_arena.p + rec[l].off;
        cudaError_t e = cudaSuccess;
        if (ida) e = cudaMemcpy(ida + tot, blk + b.ida, sizeof(int) * S, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && idb) e = cudaMemcpy(idb + tot, blk + b.idb, sizeof(int) * S, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && alpha) e = cudaMemcpy(alpha + tot, blk + b.alpha, sizeof(double) * S, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && beta && ctx->last_estep) e = cudaMemcpy(beta + tot, blk + b.beta, sizeof(double) * S, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return ctx->fail(HMC_ERR_CUDA, "trellis copy failed: %s", cudaGetErrorString(e));
        tot += S;
    }
    return tot;
}

}  // extern "C"

#include "estep_capi.inl"
#include "emit_capi.inl"
