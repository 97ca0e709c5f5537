// K2 traceback_topk: final selection of HaploBuilder::resolve (HaploBuilder.cpp:87-116) and
// HaploPair::getGenotype (HaploPair.cpp:91-124).
//
// One warp per individual.  total = sum of alpha over the final states in creation order (serial, so
// P(G) is bit-identical to the reference); the K best links over all final states (likelihood x2 when
// not homozygous) are selected with a per-lane register top-K followed by K rounds of warp arg-max
// (ties: smaller (state, index) first -- the reference leaves ties to libstdc++'s nth_element);
// then lane k walks path k back through the back-pointer words of the trellis.
#pragma once

#include "common.cuh"

namespace hmc {

struct TraceParams {
    DevPanel panel;
    DevModel model;
    const u64 *arena;
    const LocusRec *loci;       // [n_wave][L+1]
    const int *wave;
    const int *status;
    const int *final_exp;
    int n_wave, K;
    int first;                  // panel index of result row 0
    // results, indexed by (panel individual - first)
    double *pg, *cov, *prior, *post;
    int *exp2, *nres;
    signed char *haps;          // [count][K][2][L] or null
};

template <int KMAX>
__global__ void __launch_bounds__(128) traceback_kernel(TraceParams p)
{
    const int lane = threadIdx.x & 31;
    const int w = (int) ((blockIdx.x * (size_t) blockDim.x + threadIdx.x) >> 5);
    if (w >= p.n_wave) return;
    if (p.status[w] != ST_DONE) return;
    const int L = p.panel.L, K = p.K, head_len = p.model.head_len;
    const int row = p.wave[w] - p.first;
    const LocusRec *rec = p.loci + (size_t) w * (L + 1);
    const int S = rec[L].count;
    if (S <= 0) {       // unresolvable: HaploBuilder.cpp:117-124
        if (lane == 0) { p.pg[row] = 0.0; p.cov[row] = 0.0; p.nres[row] = 0; p.exp2[row] = 0; }
        return;
    }
    const BlockLayout blL = block_layout(S, K, 0);
    const double *f_alpha = (const double *) (p.arena + rec[L].off + blL.alpha);
    const double *fl = (const double *) (p.arena + rec[0].off);
    const u64 *fm = p.arena + rec[0].off + (long long) S * K;

    double total = 0.0;
    if (lane == 0) for (int s = 0; s < S; ++s) total = __dadd_rn(total, f_alpha[s]);
    total = __shfl_sync(0xffffffffu, total, 0);

    // per-lane top-K over its strided share of (state, k) candidates, in increasing ordinal
    double bl[KMAX];
    u32 bo[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { bl[k] = -1.0; bo[k] = 0xFFFFFFFFu; }
    for (int s = lane; s < S; s += 32) {
        const u64 mt = fm[s];
        const int nl = (int) (mt & 0xFF);
        const u32 hom = (u32) (mt >> 8);
        for (int k = 0; k < nl; ++k) {
            double v = fl[(size_t) s * K + k];
            if (!((hom >> k) & 1u)) v = __dmul_rn(v, 2.0);
            if (v > bl[KMAX - 1]) {
                const u32 o = (u32) s * 32u + (u32) k;
#pragma unroll
                for (int j = KMAX - 1; j >= 1; --j) {
                    const bool shift = v > bl[j - 1];
                    const bool place = !shift && (v > bl[j]);
                    if (shift) { bl[j] = bl[j - 1]; bo[j] = bo[j - 1]; }
                    else if (place) { bl[j] = v; bo[j] = o; }
                }
                if (v > bl[0]) { bl[0] = v; bo[0] = o; }
            }
        }
    }
    // K rounds of warp arg-max over the lane heads
    double my_v = -1.0;
    u32 my_o = 0xFFFFFFFFu;
    int n_sel = 0;
    double cov = 0.0;
    for (int r = 0; r < K; ++r) {
        double hv = bl[0];
        u32 ho = bo[0];
        double bv = hv;
        u32 bo_ = ho;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const u32 oo = __shfl_xor_sync(0xffffffffu, bo_, o);
            if (ov > bv || (ov == bv && oo < bo_)) { bv = ov; bo_ = oo; }
        }
        if (bv < 0.0) break;
        if (hv == bv && ho == bo_) {    // winner pops its head
#pragma unroll
            for (int j = 0; j < KMAX - 1; ++j) { bl[j] = bl[j + 1]; bo[j] = bo[j + 1]; }
            bl[KMAX - 1] = -1.0; bo[KMAX - 1] = 0xFFFFFFFFu;
        }
        if (lane == r) { my_v = bv; my_o = bo_; }
        const double post = bv / total;
        cov = __dadd_rn(cov, post);
        if (lane == 0) { p.prior[(size_t) row * K + r] = bv; p.post[(size_t) row * K + r] = post; }
        ++n_sel;
    }
    if (lane == 0) {
        p.pg[row] = total; p.cov[row] = cov; p.nres[row] = n_sel; p.exp2[row] = p.final_exp[w];
        for (int r = n_sel; r < K; ++r) { p.prior[(size_t) row * K + r] = 0.0; p.post[(size_t) row * K + r] = 0.0; }
    }
    if (!p.haps) return;

    // ---- lane r walks path r (HaploPair::getGenotype) ----
    if (lane < n_sel) {
        signed char *out = p.haps + ((size_t) row * K + lane) * 2 * L;
        int s = (int) (my_o >> 5), kk = (int) (my_o & 31u);
        int a = 0;      // which output haplotype currently plays pattern_a's role
        for (int idx = L; idx >= head_len; --idx) {
            const int Sx = rec[idx].count;
            const BlockLayout b = block_layout(Sx, K, 0);
            const u64 *blk = p.arena + rec[idx].off;
            const u32 bp = ((const u32 *) (blk + b.bp))[(size_t) s * K + kk];
            if (bp & BP_HEAD) {
                const int ia = ((const int *) (blk + b.ida))[s], ib = ((const int *) (blk + b.idb))[s];
                const signed char *pa = p.model.al + p.model.al_off[ia], *pb = p.model.al + p.model.al_off[ib];
                for (int q = 0; q < idx; ++q) { out[(size_t) a * L + q] = pa[q]; out[(size_t) (a ^ 1) * L + q] = pb[q]; }
                break;
            }
            const unsigned char ab = ((const unsigned char *) (blk + b.ab))[s];
            out[(size_t) a * L + idx - 1] = (signed char) (ab & 15);
            out[(size_t) (a ^ 1) * L + idx - 1] = (signed char) (ab >> 4);
            if (bp & BP_REV) a ^= 1;
            s = (int) (bp & ((1u << BP_PRED_BITS) - 1u));
            kk = (int) ((bp >> BP_PRED_BITS) & 31u);
        }
    }
}

}  // namespace hmc
