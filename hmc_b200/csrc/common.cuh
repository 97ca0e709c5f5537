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

// arrival word of the per-state arrival lists (resolve mode): bits 0..23 predecessor state index,
// 30 the step's allele pair was heterozygous (HaploPair.cpp:53,77), 31 reversed
constexpr u32 ARR_HET = 0x40000000u;
constexpr u32 ARR_REV = 0x80000000u;
constexpr u32 ARR_PRED_MASK = (1u << BP_PRED_BITS) - 1u;

// individual status written by the forward kernel
enum : int { ST_PENDING = 0, ST_DONE = 1, ST_NOMEM = 2, ST_TOO_BIG = 3, ST_NO_HEAD = 4, ST_NOMEM_KBEST = 5 };

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

// Row strides of the per-individual tables [n_wave][L + 1], whole 128-byte lines: the K-best kernel reads the rows of an individual
// while the forward kernel still writes its neighbours' (streamed mode), and no cached sector may hold data of two individuals.
__host__ __device__ __forceinline__ size_t loci_stride(int L) { return (size_t) ((L + 1 + 15) & ~15); }      // LocusRec (24 B) x 16 = 3 lines
__host__ __device__ __forceinline__ size_t exp_stride(int L) { return (size_t) ((L + 1 + 31) & ~31); }       // int x 32 = 1 line

// Arena with chunked bump allocation: a CTA takes CHUNK units at a time and sub-allocates.
struct Arena {
    u64 *base;
    long long capacity;         // units of 8 bytes
    unsigned long long *cursor; // next free unit
    int *exhausted;             // set once an allocation failed; later individuals stop early
};

// Per-CTA (persistent slot) spill scratch in global memory, used when a step does not fit the
// shared-memory fast path.  The hash arrays are all-empty between steps (slots are released by their owners).
struct Scratch {
    int hcap, ccap, scap;       // hash slots (power of two), candidate slots, states
    u64 *keys;                  // [hcap]
    double *tpv;                // [hcap] transition-probability product of the slot's state
    u32 *ord, *cnt, *soff;      // [hcap] smallest candidate ordinal -> state index; arrivals; arrival-segment offset
    int *cand, *sarr;           // [ccap] slot | rev << 30 of every candidate; arrival ordinals grouped by state
    int2 *mlist;                // [scap] {state index, slot} of the states with more than one arrival (2..8 from the front, more from the back)
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
    int *locus_exp;             // [n_wave][exp_stride(L)] cumulative exponent (scaled mode) or null
    int *queue;                 // work-queue head
    long long *counters;        // steps, states, max_states, arrivals, -, spill steps
    int *done_list, *done_count; // streamed K-best: every individual the kernel is through with (any status) is appended, in completion order; or null
    long long *work;            // [N] SM cycles the forward kernel spent on every individual it resolved (orders the next pass: longest first)
    long long *ind_cycles;      // [n_wave][8] per-individual diagnostics (cycles, states, ...) or null
};

// trellis state-block layout (units of 8 bytes) for S states with X arrivals.
//   resolve mode: ida | idb | alpha | v1 (likelihood of the best path into the state) | last alleles | aoff[S+1] | arrival words[X]
//   E-step mode:  ida | idb | alpha | last alleles | tp | beta        (X unused; forward links live in their own block)
struct BlockLayout {
    long long ida, idb, alpha, v1, ab, aoff, arr, tp, beta, total;
};
__host__ __device__ inline BlockLayout block_layout(int S, long long X, int estep)
{
    BlockLayout b;
    long long s2 = (S + 1) / 2;     // S ints in 8-byte units
    b.ida = 0;
    b.idb = b.ida + s2;
    b.alpha = b.idb + s2;
    long long o = b.alpha + S;
    b.v1 = b.tp = b.beta = b.aoff = b.arr = o;
    if (estep) {
        b.ab = o; o += (S + 7) / 8;
        b.tp = o; b.beta = o + S; o += 2LL * S;
    }
    else {
        b.v1 = o; o += S;
        b.ab = o; o += (S + 7) / 8;
        b.aoff = o; o += (S + 2) / 2;
        b.arr = o; o += (X + 1) / 2;
    }
    b.total = o;
    return b;
}

__device__ __forceinline__ u32 mix64(u64 k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (u32) k;
}

}  // namespace hmc
