// K8 emit: what follows the resolve pass on the way out of the device.
//   likelihood_kernel  HaploBuilder::getLikelihood(const Haplotype &) (HaploBuilder.cpp:128-146): the product of the transition
//                      probabilities of the longest matching pattern ending at every index head_len .. L.  The reference descends the
//                      backward trie once per index; here one thread per haplotype walks the automaton (the successor of the longest
//                      match at i under allele a[i] IS the longest match at i + 1, the property the forward pass rests on too).
//   emit_len / emit_text  the allele lines of `.reconstructed` (HaploFile::writeGenoData, HaploFile.cpp:146-148, through
//                      AlleleSequence::writeAllele, Allele.cpp:82-101): the symbol and a blank at an 'S' locus ('?' when missing), "%d "
//                      elsewhere.  One warp per haplotype line, 32 loci per round, positions by a warp scan of the text widths.
#pragma once

#include "common.cuh"

namespace hmc {

// haplotype q lies at haps + ((q >> 1) * pair_stride + (q & 1)) * L; flag[0] is set when an allele is missing or out of range
__global__ void __launch_bounds__(128) likelihood_kernel(DevModel md, int L, const u32 *linfo, const signed char *haps, long long pair_stride, int n,
                                                         double *out, int *flag)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const signed char *h = haps + ((size_t) (q >> 1) * pair_stride + (q & 1)) * L;
    const int A = md.Amax;
    long long code = 0;
    for (int j = 0; j < md.head_len; ++j) {
        const int a = h[j];
        if (a < 0 || a >= (int) (linfo[j] & 0xFF)) { atomicExch(flag, 1); out[q] = 0; return; }
        code += (long long) a * md.head_stride[j];
    }
    int p = md.head_lut[code];
    double lk = p >= 0 ? md.tp[p] : 0.0;
    for (int i = md.head_len; i < L && p >= 0; ++i) {
        const int a = h[i];
        if (a < 0 || a >= (int) (linfo[i] & 0xFF)) { atomicExch(flag, 1); out[q] = 0; return; }
        p = md.succ[(size_t) p * A + a];
        lk = p >= 0 ? __dmul_rn(lk, md.tp[p]) : 0.0;
    }
    out[q] = lk;
}

struct EmitParams {
    int n, L, Amax, K;              // n individuals (result rows), alleles per locus bound, resolutions per individual
    const signed char *haps;        // [n][K][2][L] resolutions, best first
    const signed char *raw;         // [n][2][L] the input genotypes of the same individuals (written for an unresolved individual, HaploBuilder.cpp:117-124)
    const int *nres;                // [n]
    const int *sym;                 // [L][Amax] allele value: the character at an 'S' locus, the integer elsewhere
    const unsigned char *wid;       // [L][Amax + 1] bytes of the allele's text with its blank; entry Amax = missing
    const char *type;               // [L]
    long long *off;                 // [2n + 1] line offsets (emit_len writes the lengths, a scan turns them into offsets)
    char *text;
};

__device__ __forceinline__ const signed char *emit_row(const EmitParams &e, int line)
{
    const int i = line >> 1, h = line & 1;
    return e.nres[i] > 0 ? e.haps + (((size_t) i * e.K) * 2 + h) * e.L : e.raw + ((size_t) i * 2 + h) * e.L;
}

__global__ void __launch_bounds__(256) emit_len_kernel(EmitParams e)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; line < 2 * e.n; line += warps) {
        const signed char *row = emit_row(e, line);
        int len = 0;
        for (int l = lane; l < e.L; l += 32) { const int k = row[l]; len += e.wid[(size_t) l * (e.Amax + 1) + (k < 0 ? e.Amax : k)]; }
        for (int d = 16; d; d >>= 1) len += __shfl_xor_sync(0xffffffffu, len, d);
        if (lane == 0) e.off[line] = len + 1;       // + the newline
    }
}

__global__ void __launch_bounds__(256) emit_text_kernel(EmitParams e)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int line = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; line < 2 * e.n; line += warps) {
        const signed char *row = emit_row(e, line);
        char *dst = e.text + e.off[line];
        for (int base = 0; base < e.L; base += 32) {
            const int l = base + lane;
            int k = -1, w = 0;
            if (l < e.L) { k = row[l]; w = e.wid[(size_t) l * (e.Amax + 1) + (k < 0 ? e.Amax : k)]; }
            int incl = w;
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
            if (l < e.L) {
                char *o = dst + incl - w;
                const int v = k >= 0 ? e.sym[(size_t) l * e.Amax + k] : -1;
                if (e.type[l] == 'S') { o[0] = v < 0 ? '?' : (char) v; o[1] = ' '; }
                else {      // "%d "
                    unsigned u = v < 0 ? 0u - (unsigned) v : (unsigned) v;
                    int at = w - 1;
                    o[at--] = ' ';
                    do { o[at--] = (char) ('0' + u % 10u); u /= 10u; } while (u);
                    if (v < 0) o[at] = '-';
                }
            }
            dst += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) *dst = '\n';
    }
}

}  // namespace hmc
