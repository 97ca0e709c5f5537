// Shared device-side definitions of the hmc_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hmc {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int MAX_ALLELES = 16;
constexpr int MAX_GROUPS = MAX_ALLELES * (MAX_ALLELES + 1) / 2;   // allele pairs tried at a fully missing locus
constexpr u64 KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;

// back-pointer word of one K-best link (HaploPairLink, HaploPair.h:14-20):
//   bits 0..23 predecessor state index at the previous trellis index, 24..28 index in its list,
//   29 reversed, 30 homozygous, 31 head (no predecessor)
constexpr u32 BP_HEAD = 0x80000000u;
constexpr u32 BP_HOM = 0x40000000u;
constexpr u32 BP_REV = 0x20000000u;
constexpr int BP_PRED_BITS = 24;
constexpr int MAX_STATES_PER_LOCUS = 1 << BP_PRED_BITS;

// individual status written by the forward kernel
enum : int { ST_PENDING = 0, ST_DONE = 1, ST_NOMEM = 2, ST_TOO_BIG = 3, ST_NO_HEAD = 4 };

// The panel on the device: genotypes bit-packed per individual x locus (K0).
struct DevPanel {
    int N, L, Amax;
    int bits;                   // bits per allele code (code = index + 1, 0 = missing): 2, 4 or 8
    int words;                  // 32-bit words per individual
    const u32 *packed;          // [N][words]
    const u32 *linfo;           // [L] allele_num(l) | mask << 8, mask bit j set <=> allele_frequency(l, j) > 0
    const int *order;           // [N] work order (heaviest individuals first)
};

// The automaton in SoA form.
struct DevModel {
    int P, head_len, n_heads, Amax;
    const int *succ;            // [P][Amax]
    const void *succ16;         // [P][Amax] 16-byte entries {int successor id; int pad; double transition_prob of the successor}
    const double *tp;           // [P] transition_prob
    const double *freq;         // [P] frequency
    const int *start;           // [P]
    const unsigned char *last;  // [P] last allele index
    const long long *al_off;    // [P+1]
    const signed char *al;      // concatenated alleles
    const int *heads;           // [n_heads]
    const int *head_lut;        // mixed-radix allele code of a head-length string -> head pattern id or -1
    int head_stride[8];
};

// Where the trellis of one (individual, index) lives in the arena (8-byte units).
struct LocusRec {
    long long off;              // state block
    int count;                  // live states
    int fwd_slots;              // candidate slots of the step leaving this index (E-step mode)
    long long fwd_off;          // forward-link block (E-step mode)
};

// Arena with chunked bump allocation: a CTA takes CHUNK units at a time and sub-allocates.
struct Arena {
    u64 *base;
    long long capacity;         // units of 8 bytes
    unsigned long long *cursor; // next free unit
    int *exhausted;             // set once an allocation failed; later individuals stop early
};

// Per-CTA (persistent slot) spill scratch in global memory, used when a step does not fit the
// shared-memory fast path, plus the K-best likelihood ping-pong.
struct Scratch {
    int hcap, ccap, scap;       // hash slots (power of two), candidate slots, states
    u64 *keys;                  // [hcap]
    double *tpv;                // [hcap] transition-probability product of the slot's state
    double *otp;                // [scap] warp kernel: transition probability of spilled output states
    u32 *obp;                   // [scap*K] warp kernel: back-pointer words of spilled output states
    u32 *ord, *cnt, *nidx;      // [hcap]
    int *cand, *arr;            // [ccap]
    int *aoff, *slot_of;        // [scap+1], [scap]
    int *order;                 // [scap] pass-E visiting order (light states first)
    double *like[2];            // [scap*K]
    u64 *meta[2];               // [scap]  (links | hom mask << 8)
};

struct FwdParams {
    DevPanel panel;
    DevModel model;
    Arena arena;
    Scratch *scratch;           // [gridDim.x]
    LocusRec *loci;             // [n_ind][L+1] for the individuals of this wave (indexed by wave position)
    const int *wave;            // [n_wave] panel individual of each wave position
    int n_wave;
    int K;
    int estep;                  // keep tp/beta/forward links instead of back-pointers
    int scaled;
    int *status;                // [n_wave]
    int *final_exp;             // [n_wave]
    int *locus_exp;             // [n_wave][L+1] cumulative exponent (scaled mode, E-step) or null
    int *queue;                 // work-queue head
    long long *counters;        // steps, states, max_states, arrivals, -, spill steps
    long long *ind_cycles;      // [n_wave] SM clock cycles spent on each individual (diagnostics) or null
};

// trellis state-block layout (units of 8 bytes) for S states
struct BlockLayout {
    long long ida, idb, alpha, ab, bp, tp, beta, total;
};
__host__ __device__ inline BlockLayout block_layout(int S, int K, int estep)
{
    BlockLayout b;
    long long s2 = (S + 1) / 2;     // S ints in 8-byte units
    b.ida = 0;
    b.idb = b.ida + s2;
    b.alpha = b.idb + s2;
    b.ab = b.alpha + S;
    long long o = b.ab + (S + 7) / 8;
    b.bp = b.tp = b.beta = o;
    if (estep) { b.tp = o; b.beta = o + S; o += 2LL * S; }
    else { b.bp = o; o += ((long long) S * K + 1) / 2; }
    b.total = o;
    return b;
}

__device__ __forceinline__ u32 mix64(u64 k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (u32) k;
}

}  // namespace hmc
