// K7 compare: HaploComp (HaploComp.cpp:29-76) -- switch error, incorrect-genotype and incorrect-haplotype percentages of
// inferred against real (target) phased genotypes; Genotype::getSwitchDistanceIgnoreMissing (Genotype.cpp:235-279),
// getDiffNumIgnoreMissing (:160-175), checkGenotype (:44-56), Allele::isMatch (Allele.h:20: missing matches anything).
// One warp per individual: 32 loci per step as three ballots (skip / matches straight / matches reversed); the orientation
// state machine of the switch distance then runs over the bits of the two match masks.
#pragma once

#include "common.cuh"

namespace hmc {

struct CompareOut {     // device counters
    long long se_num, se_den, ig_num, ig_den, ih_num, ih_den;
    unsigned long long bad;     // (individual << 32 | locus) of the first inconsistent individual, ~0 if none
};

// real / infer: allele codes (negative = missing); row(i, h) of individual i = base + (i * stride_ind + h) * L
__global__ void __launch_bounds__(256) compare_kernel(const signed char *real, long long real_stride, const signed char *infer, long long infer_stride,
                                                      int N, int L, CompareOut *out)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < N; i += warps) {
        const signed char *r0 = real + (size_t) i * real_stride * L, *r1 = r0 + L;
        const signed char *f0 = infer + (size_t) i * infer_stride * L, *f1 = f0 + L;
        int het = 0, miss = 0, d_rev = 0, d_fwd = 0, sd = 0;
        int state = 0;          // 0 = no informative locus yet, 1 = straight, 2 = reversed
        int bad = -1;
        for (int base = 0; base < L && bad < 0; base += 32) {
            const int l = base + lane;
            const bool in = l < L;
            const int a0 = in ? r0[l] : -1, a1 = in ? r1[l] : -1, b0 = in ? f0[l] : -1, b1 = in ? f1[l] : -1;
            auto am = [](int x, int y) { return x < 0 || y < 0 || x == y; };
            const bool hm = a0 < 0 || a1 < 0;
            const bool mf = am(a0, b0) && am(a1, b1), mr = am(a0, b1) && am(a1, b0);
            const u32 valid = __ballot_sync(0xffffffffu, in);
            const u32 skip = __ballot_sync(0xffffffffu, hm) | ~valid;
            const u32 F = __ballot_sync(0xffffffffu, mf), R = __ballot_sync(0xffffffffu, mr);
            het += __popc(__ballot_sync(0xffffffffu, in && !am(a0, a1)));
            miss += __popc(__ballot_sync(0xffffffffu, in && hm));
            d_rev += __popc(~skip & ~R);
            d_fwd += __popc(~skip & ~F);
            // loci that can move the state machine: not skipped and (before the start) not matching both ways
            u32 todo = ~skip;
            while (todo) {
                const int b = __ffs(todo) - 1;
                todo &= todo - 1;
                const bool f = (F >> b) & 1u, r = (R >> b) & 1u;
                if (state == 0) {
                    if (f && r) continue;
                    if (r) state = 2; else if (f) state = 1; else { bad = base + b; break; }      // reversed is tried first (Genotype.cpp:252-257)
                }
                else {
                    const bool cur = state == 2 ? r : f, oth = state == 2 ? f : r;
                    if (cur) continue;
                    if (oth) { state = 3 - state; ++sd; }
                    else { bad = base + b; break; }
                }
            }
        }
        if (lane == 0) {
            if (bad >= 0) atomicMin(&out->bad, ((unsigned long long) i << 32) | (unsigned) bad);
            else {
                atomicAdd((unsigned long long *) &out->se_num, (unsigned long long) sd);
                atomicAdd((unsigned long long *) &out->se_den, (unsigned long long) (long long) (het - 1));
                atomicAdd((unsigned long long *) &out->ig_num, (unsigned long long) min(d_rev, d_fwd));
                atomicAdd((unsigned long long *) &out->ig_den, (unsigned long long) (L - miss));
                if (sd > 0) atomicAdd((unsigned long long *) &out->ih_num, 1ull);
                if (het > 1) atomicAdd((unsigned long long *) &out->ih_den, 1ull);
            }
        }
    }
}

}  // namespace hmc
