// K6 build_automaton: PatternManager::initialize on the device (PatternManager.cpp:293-318).
// For every pattern p = (start, end, a[0..len)) and every allele j of locus `end`, the successor is the LONGEST pattern
// that ends at end + 1, whose string is a trailing part of a ++ j and whose start is >= p.start
// (BackwardPatternTree::findLongestMatchPattern, PatternTree.cpp:98-134: the walk goes back from j through a and keeps the
// deepest node that carries a pattern).  The reference descends one backward trie per end locus; here all patterns go
// into one open-addressing hash table keyed by (start, end, polynomial hash of the string), and a thread per pattern
// tries the candidates start' = start, start + 1, ... with rolling suffix hashes; every hit is verified against the
// stored allele string, so hash collisions cannot change the result.
#pragma once

#include "common.cuh"

namespace hmc {

constexpr u64 AUT_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr u64 AUT_BASE = 0x9E3779B97F4A7C15ull;      // odd multiplier of the string hash (mod 2^64)
constexpr int AUT_MAX_LEN = 256;

struct AutParams {
    long long P;
    int L, Amax;
    const int *start, *end;
    const long long *al_off;
    const signed char *al;
    const u32 *linfo;
    u64 *keys;
    int *vals;
    u64 hmask;
    int *succ;              // [P][Amax]
    unsigned char *last;
    const u64 *bpow;        // AUT_BASE^t, t = 0..AUT_MAX_LEN
};

__device__ __forceinline__ u64 aut_key(u64 hstr, int start, int end)
{
    u64 k = hstr ^ ((u64) (u32) start << 32) ^ (u64) (u32) end * 0xD6E8FEB86659FD93ull;
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;     // full 64-bit finaliser
    return k == AUT_EMPTY ? 0 : k;
}

__global__ void aut_insert_kernel(AutParams a)
{
    const long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (p >= a.P) return;
    const signed char *s = a.al + a.al_off[p];
    const int len = a.end[p] - a.start[p];
    u64 h = 0;
    for (int t = 0; t < len; ++t) h = h * AUT_BASE + (u64) (s[t] + 1);
    a.last[p] = (unsigned char) s[len - 1];
    const u64 key = aut_key(h, a.start[p], a.end[p]);
    u64 slot = key & a.hmask;
    for (;;) {
        const u64 old = atomicCAS((unsigned long long *) &a.keys[slot], (unsigned long long) AUT_EMPTY, (unsigned long long) key);
        if (old == AUT_EMPTY) { a.vals[slot] = (int) p; return; }
        slot = (slot + 1) & a.hmask;        // occupied (an equal key is a different pattern: keep both)
    }
}

__global__ void aut_succ_kernel(AutParams a)
{
    const long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (p >= a.P) return;
    const int start = a.start[p], end = a.end[p], len = end - start, A = a.Amax;
    int *out = a.succ + (size_t) p * A;
    for (int j = 0; j < A; ++j) out[j] = -1;
    if (end >= a.L) return;
    const signed char *s = a.al + a.al_off[p];
    const int nA = (int) (a.linfo[end] & 0xFF);
    u64 h0 = 0;
    for (int t = 0; t < len; ++t) h0 = h0 * AUT_BASE + (u64) (s[t] + 1);
    for (int j = 0; j < nA; ++j) {
        u64 h = h0;                                     // hash of s[k..len)
        int result = -1;
        for (int k = 0; k <= len && result < 0; ++k) {
            const int clen = len - k + 1;               // candidate: start + k, end + 1, s[k..len) ++ j
            const u64 key = aut_key(h * AUT_BASE + (u64) (j + 1), start + k, end + 1);
            u64 slot = key & a.hmask;
            for (;;) {
                const u64 kk = a.keys[slot];
                if (kk == AUT_EMPTY) break;
                if (kk == key) {
                    const int c = a.vals[slot];
                    if (a.start[c] == start + k && a.end[c] == end + 1) {
                        const signed char *cs = a.al + a.al_off[c];
                        bool same = cs[clen - 1] == j;
                        for (int t = 0; t < clen - 1 && same; ++t) same = cs[t] == s[k + t];
                        if (same) { result = c; break; }
                    }
                }
                slot = (slot + 1) & a.hmask;
            }
            if (k < len) h -= (u64) (s[k] + 1) * a.bpow[len - 1 - k];
        }
        out[j] = result;
    }
}

// head patterns: start 0, length head_len, in id order (PatternManager.cpp:303-306); flags for the compaction
__global__ void aut_head_flags_kernel(AutParams a, int head_len, unsigned char *flags)
{
    const long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x;
    if (p >= a.P) return;
    flags[p] = (a.start[p] == 0 && a.end[p] == head_len) ? 1 : 0;
}

struct HeadStride { int s[8]; };

__global__ void aut_head_lut_kernel(AutParams a, const int *heads, int n_heads, int head_len, HeadStride st, int *lut)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_heads) return;
    const int id = heads[h];
    const signed char *s = a.al + a.al_off[id];
    long long code = 0;
    for (int j = 0; j < head_len; ++j) code += (long long) s[j] * st.s[j];
    lut[code] = id;
}

}  // namespace hmc
