/*
 * hmc_b200 -- C-ABI of the B200 (sm_100a) HaploBuilder / EM E-step path of HMC.
 *
 * This is the drop-in boundary for the reference's HaploBuilder (HaploBuilder.h:33-54) as driven by
 * HaploModel::resolveAll (HaploModel.cpp:79-115) and PatternManager::estimateFrequency /
 * estimatePatterns (PatternManager.cpp:347-410).  The reference has no FFI of its own (it is one
 * single-threaded C++ program); the calls are batched over individuals because a per-individual
 * call (the reference's `resolve`) would serialise the GPU.  INTEGRATION.md shows the reference-side
 * glue a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every buffer is caller-owned HOST memory unless the name says
 *     otherwise; the context owns all device memory;
 *   - every function returns HMC_OK (0) or a negative hmc_status; nothing calls exit() (the
 *     reference does: Logger::error + exit(1), e.g. HaploBuilder.cpp:215-218);
 *   - alleles are per-locus allele INDICES in GenoData::getAlleleIndex order (GenoData.cpp:27-44,
 *     symbols sorted ascending, GenoData.cpp:96-98); -1 is the missing allele (Allele.h:13);
 *   - pattern ids are positions in the arrays handed to hmc_set_model, exactly the ids
 *     PatternManager::initialize assigns (PatternManager.cpp:300-302);
 *   - all probabilities are IEEE fp64; products and sums are evaluated in the reference's order
 *     without FMA contraction, so alpha, P(G) and priors are bit-identical to the reference;
 *   - one host thread per context (the reference is non-reentrant too).
 *
 * There is no CPU fallback: without a CUDA device hmc_ctx_create fails with HMC_ERR_NO_DEVICE.
 */
#ifndef HMC_B200_H
#define HMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hmc_ctx hmc_ctx;

typedef enum hmc_status {
    HMC_OK = 0,
    HMC_ERR_NO_DEVICE = -1,     /* no CUDA device / wrong architecture */
    HMC_ERR_CUDA = -2,          /* a CUDA call failed; see hmc_last_error */
    HMC_ERR_INVALID = -3,       /* bad argument or call order */
    HMC_ERR_LIMIT = -4,         /* a documented capacity was exceeded (alleles per locus, K, states per locus) */
    HMC_ERR_NOMEM = -5,         /* the trellis arena cannot hold even one individual */
    HMC_ERR_NO_HEAD = -6        /* "Can not find matching pattern!" (HaploBuilder.cpp:215-218) */
} hmc_status;

/* flags for hmc_resolve_all / hmc_estimate_counts */
#define HMC_FLAG_SCALED        1u   /* exponent-only rescaling by 2^512 when max alpha < 2^-512; off = the
                                       reference's raw fp64 products, bit for bit */
#define HMC_FLAG_KEEP_TRELLIS  2u   /* keep the last wave's trellis on the device for hmc_get_trellis (tests) */
#define HMC_FLAG_PANEL_ROWS    4u   /* hmc_resolve_all_device only: the device result block has one row per PANEL individual (row =
                                       individual, not individual - first); the call fills rows [first, first+count) and the rows of the
                                       other shards are filled by an all-gather on the pointers of hmc_result_device_ptrs (multi-GPU) */

#define HMC_MAX_ALLELES 16          /* alleles per locus (Amax) supported by the packed genotype / pair encoding */
#define HMC_MAX_K       32          /* sample_size */

/* hmc_get_counters slots */
#define HMC_CNT_STEPS       0       /* (individual, locus) extensions that produced >= 1 state */
#define HMC_CNT_STATES      1       /* sum of live states over all trellis indices (head included) */
#define HMC_CNT_MAX_STATES  2       /* largest live-state set */
#define HMC_CNT_ARRIVALS    3       /* addHaploPair calls (new state + merges), HaploBuilder.cpp:246 */
#define HMC_CNT_WAVES       4       /* arena waves used by the last call */
#define HMC_CNT_SPILL_STEPS 5       /* steps that left the shared-memory fast path */
#define HMC_CNT_LAUNCHES    6       /* kernels launched by the last call */
#define HMC_CNT_ESTEP_ENTRIES 7     /* E-step: (node, matched state, class) entries read by the trie walk (K4 traffic model) */
#define HMC_CNT_N           8

/* Create a context on CUDA device `device`.  arena_bytes = 0 picks ~60% of free device memory for
 * the trellis arena.  threads = 0 picks the threads-per-individual group size from the panel. */
int hmc_ctx_create(hmc_ctx **out, int device, size_t arena_bytes, int threads);
void hmc_ctx_destroy(hmc_ctx *ctx);
const char *hmc_last_error(const hmc_ctx *ctx);
/* version string of the loaded library, e.g. "hmc_b200 0.1 sm_100a" */
const char *hmc_version(void);

/* The panel: replaces HaploBuilder::setGenoData (HaploBuilder.h:47) + the GenoData allele tables
 * (GenoData.h:43-49).  geno: [N][2][L] allele indices; n_alleles: [L]; allele_freq: [L][Amax].
 * Genotypes are bit-packed on the device (K0). */
int hmc_set_panel(hmc_ctx *ctx, int N, int L, int Amax, const int8_t *geno, const int32_t *n_alleles, const double *allele_freq);

/* The same for a panel that is already bit-packed in the device layout (binary panel files): per individual
 * `words` = ceil(L / (32 / (2*bits))) 32-bit words, locus l at bit (l % per) * 2*bits of word l / per, low field = first
 * allele, each field = allele index + 1 (0 = missing); bits = 2 for Amax <= 3, 4 for Amax <= 15, else 8.  One H2D copy of a
 * quarter (biallelic) of the unpacked size and no host pass over the genotypes. */
int hmc_set_panel_packed(hmc_ctx *ctx, int N, int L, int Amax, int bits, const uint32_t *packed, const int32_t *n_alleles,
                         const double *allele_freq);
/* The best resolution of every individual of the last resolve call in the same packed layout: out_packed [count][words]. */
int hmc_fetch_best_packed(hmc_ctx *ctx, uint32_t *out_packed);

/* The automaton: the HaploPattern fields HaploBuilder reads (HaploPattern.h:19-24,36-37) after
 * PatternManager::initialize (PatternManager.cpp:293-318).  succ: [P][Amax], -1 = NULL successor;
 * alleles: concatenated allele indices, al_off: [P+1]; heads: head_list ids in order
 * (PatternManager.cpp:304-306). */
int hmc_set_model(hmc_ctx *ctx, int P, int head_len, const int32_t *start, const int32_t *end, const double *freq,
                  const double *tp, const int32_t *succ, const int64_t *al_off, const int8_t *alleles, int n_heads, const int32_t *heads);

/* HaploModel::resolveAll (HaploModel.cpp:79-115) over individuals [first, first+count) of the panel:
 * one HaploBuilder::resolve (HaploBuilder.cpp:35-126) each, K = sample_size.
 *   out_pg      [count]        genotype_probability (scaled by 2^out_exp when HMC_FLAG_SCALED)
 *   out_exp     [count]        power-of-two exponent carried (0 without HMC_FLAG_SCALED)
 *   out_cov     [count]        coverage returned by resolve()
 *   out_nres    [count]        res_list.size()
 *   out_prior   [count][K]     prior_probability of each entry, descending
 *   out_post    [count][K]     posterior_probability
 *   out_haps    [count][K][2][L] phased allele indices of each entry (may be NULL to skip the copy)
 *   out_ll      [1]            sum of log P(G) over the range, in individual order
 * Any out_* may be NULL.  Unresolvable individuals (no final state) get pg = 0, nres = 0. */
int hmc_resolve_all(hmc_ctx *ctx, int first, int count, int K, unsigned flags, double *out_pg, int32_t *out_exp, double *out_cov,
                    int32_t *out_nres, double *out_prior, double *out_post, int8_t *out_haps, double *out_ll);

/* Device-resident variant used for kernel-only timing: runs the same kernels, leaves the results
 * in device buffers owned by the context; hmc_fetch_results copies them out afterwards. */
int hmc_resolve_all_device(hmc_ctx *ctx, int first, int count, int K, unsigned flags);
int hmc_fetch_results(hmc_ctx *ctx, double *out_pg, int32_t *out_exp, double *out_cov, int32_t *out_nres, double *out_prior,
                      double *out_post, int8_t *out_haps, double *out_ll);

/* Multi-GPU exchange without host staging (one process per GPU): device pointers of the result block of the last
 * hmc_resolve_all_device call -- ptrs[7] = {pg, exp, coverage, nres, prior, posterior, haps}, row_bytes[7] = bytes per individual --
 * for an in-place all-gather over the individuals (ncclAllGather / torch.distributed on the same device).  With HMC_FLAG_PANEL_ROWS the
 * block has N rows; after the exchange hmc_fetch_results / hmc_fetch_best_haps / hmc_mine_patterns_resolved / hmc_compare see all of
 * them (this replaces the gather that HaploModel::resolveAll does by being one loop, HaploModel.cpp:86-109). */
int hmc_result_device_ptrs(hmc_ctx *ctx, void *ptrs[7], int64_t row_bytes[7]);
/* Raw count vector of the last hmc_estimate_counts_device call (after the caller's in-place all-reduce): out_freq / out_prefix [Pc]. */
int hmc_fetch_counts(hmc_ctx *ctx, int Pc, double *out_freq, double *out_prefix);
/* The CUDA stream the context's kernels run on (cudaStream_t), so that a collective can be ordered after them. */
void *hmc_stream(hmc_ctx *ctx);

/* HaploBuilder::estimateFrequency(vector<HaploPattern*>&) (HaploBuilder.cpp:274-332): expected
 * pattern counts over individuals [first, first+count) for an explicit candidate list (which may
 * hold patterns absent from the automaton, PatternManager.cpp:377-388).
 *   start [Pc], al_off [Pc+1], alleles: the candidates
 *   pg [count], pg_exp [count] (may be NULL = 0): genotype_probability from the preceding resolve_all
 *   out_freq / out_prefix [Pc]: RAW sums over the range (before the normalisation of
 *   HaploBuilder.cpp:317-331), so shards can be added (ncclAllReduce) before hmc_normalise_counts. */
int hmc_estimate_counts(hmc_ctx *ctx, int first, int count, unsigned flags, int Pc, const int32_t *start, const int64_t *al_off,
                        const int8_t *alleles, const double *pg, const int32_t *pg_exp, double *out_freq, double *out_prefix);
/* Same, but the raw sums stay on the device: *dev_counts receives a device pointer to 2*Pc doubles
 * (freq then prefix) owned by the context, for an in-place NCCL all-reduce. */
int hmc_estimate_counts_device(hmc_ctx *ctx, int first, int count, unsigned flags, int Pc, const int32_t *start, const int64_t *al_off,
                               const int8_t *alleles, const double *pg, const int32_t *pg_exp, double **dev_counts);
/* HaploBuilder.cpp:317-331 on host arrays: clamp to N, freq <= prefix, divide by N, transition probability. */
void hmc_normalise_counts(int Pc, int N, double *freq, double *prefix, double *tp);

/* K5: the pattern search of the M-step, PatternManager::searchPattern with checkFrequencyWithExtension /
 * getMatchingFrequency (PatternManager.cpp:100-291), for the frequency-threshold models (findPatternByFreq,
 * findPatternBlock).  n_samples == 0: the genotype branch over the panel of hmc_set_panel (:156-179,205-251);
 * n_samples > 0: the sampling branch over weighted sampled haplotypes (:180-191,252-263),
 *   sample_haps [n_samples][L] allele indices (-1 missing), sample_weights [n_samples], total_weight = HaploData's.
 * min_len / max_len [L]: per start locus (PatternManager.cpp:33-34).  Patterns come back in the reference's id order
 * (release order of its depth-first stack) with bit-identical frequency, prefix_freq and transition_prob.
 * Two calls: hmc_mine_patterns returns the sizes, hmc_fetch_patterns copies into caller buffers
 *   start/end/freq/prefix_freq/tp [P], al_off [P+1], alleles [n_alleles]. */
int hmc_mine_patterns(hmc_ctx *ctx, int n_samples, const int8_t *sample_haps, const double *sample_weights, double total_weight,
                      double min_freq, const int32_t *min_len, const int32_t *max_len, int64_t *out_P, int64_t *out_n_alleles);
int hmc_fetch_patterns(hmc_ctx *ctx, int32_t *start, int32_t *end, double *freq, double *prefix_freq, double *tp, int64_t *al_off,
                       int8_t *alleles);
/* Sampling branch without moving the sampled haplotypes: rows [n_samples] index the haplotype rows
 * ((individual - first) * K + k) * 2 + h that the last hmc_resolve_all / hmc_resolve_all_device call left on the device
 * (HaploModel::resolveAll fills HaploData from exactly these rows, HaploModel.cpp:94-108). */
int hmc_mine_patterns_resolved(hmc_ctx *ctx, int n_samples, const int32_t *rows, const double *sample_weights, double total_weight,
                               double min_freq, const int32_t *min_len, const int32_t *max_len, int64_t *out_P, int64_t *out_n_alleles);
/* Individuals and K of the result block the last resolve call left on the device (either pointer may be NULL). */
int hmc_result_dims(hmc_ctx *ctx, int32_t *count, int32_t *K);
/* The best resolution (k = 0) of every individual of the last resolve call: out_best [count][2][L]. */
int hmc_fetch_best_haps(hmc_ctx *ctx, int8_t *out_best);
/* The same list as records in page-locked host memory owned by the context (one D2H copy at link speed, no caller
 * buffers): valid until the next hmc_mine_patterns / hmc_ctx_destroy.  `al` is the offset of the allele string. */
typedef struct hmc_pattern_rec {
    int32_t start, end;
    int64_t al;
    double frequency, prefix_freq, transition_prob;
} hmc_pattern_rec;
int hmc_map_patterns(hmc_ctx *ctx, const hmc_pattern_rec **records, const int8_t **alleles);
/* K6: PatternManager::initialize (PatternManager.cpp:293-318) over the list of the last hmc_mine_patterns call, on the
 * device: successor table by longest trailing match with start >= the pattern's start (PatternTree.cpp:98-134), head
 * list (:303-306); the result is installed as the context's model exactly as hmc_set_model would (no host round trip).
 * Large lists (> 2 GB) are not copied: their device arrays become the model's, so hmc_fetch_patterns / hmc_map_patterns must
 * come BEFORE this call (a view already mapped stays valid).
 * hmc_fetch_automaton copies succ [P][Amax] and heads [n_heads] out for parity checks / export; either may be NULL. */
int hmc_build_automaton(hmc_ctx *ctx, int head_len, int32_t *out_n_heads);
int hmc_fetch_automaton(hmc_ctx *ctx, int32_t *succ, int32_t *heads);
/* Device time (ms, mining + automaton) and trie statistics of the last hmc_mine_patterns call; any pointer may be NULL.
 * swept / written: match-list entries read by the sweeps / written into child lists (the roofline's traffic model). */
int hmc_get_mine_stats(hmc_ctx *ctx, double *ms, int64_t *nodes, int64_t *levels, int64_t *max_list, int64_t *swept, int64_t *written);

/* K7: HaploComp (HaploComp.cpp:29-76, as used by HaploModel::run :134 and HMC::compare, HMC.cpp:246-257) of inferred
 * against real (target) phased genotypes, both [N][2][L] allele codes in any per-locus coding shared by the two sides
 * (negative = missing).  infer == NULL compares the best resolutions of the last resolve call where they lie on the device.
 * out: switch-error numerator / denominator, incorrect-genotype num / den, incorrect-haplotype num / den; the three ratios
 * are num / den in double (HaploComp.cpp:145-147).  Data the reference would exit on ("Inconsistent genotypes at locus
 * %d!", Genotype.cpp:258-261,272-275) returns HMC_ERR_INVALID with the first offending individual / locus. */
int hmc_compare(hmc_ctx *ctx, int N, int L, const int8_t *real, const int8_t *infer, int64_t out[6], int32_t *bad_individual, int32_t *bad_locus);

/* K8a: HaploBuilder::getLikelihood(const Haplotype &) (HaploBuilder.cpp:128-146; the Genotype overload :148-151 is the product of the
 * two haplotypes' values): likelihood of n fully typed haplotypes [n][L] (allele indices) under the model on the device -- the product
 * of the transition probabilities of the longest matching pattern ending at every index head_len .. L, 0 when no pattern matches.
 * haps == NULL takes both haplotypes of the best resolution of every individual of the last resolve call where they lie on the
 * device (n = 2 * result rows).  A missing or out-of-range allele returns HMC_ERR_INVALID. */
int hmc_get_likelihood(hmc_ctx *ctx, int n, const int8_t *haps, double *out);

/* K8b: the allele lines of `.reconstructed` (HaploFile::writeGenoData, HaploFile.cpp:146-148 through AlleleSequence::write /
 * writeAllele, Allele.cpp:82-101) formatted on the device from the best resolution of every individual of the last resolve call
 * (the input genotype for an individual that could not be resolved, HaploBuilder.cpp:117-124).  symbols [L][Amax]: the value of
 * every allele index -- the character at an 'S' locus, the integer at any other; allele_type [L].  Two lines per result row, each
 * ending in '\n'; line_off [2 * rows + 1] (may be NULL) receives the byte offset of every line, *total the size of the text.
 * out == NULL only sizes; a buffer smaller than the text returns HMC_ERR_LIMIT with *total set. */
int hmc_format_best(hmc_ctx *ctx, const int32_t *symbols, const char *allele_type, char *out, int64_t capacity, int64_t *line_off,
                    int64_t *total);

/* Structural counters of the last resolve/estimate call (HMC_CNT_*), for parity and the roofline. */
int hmc_get_counters(hmc_ctx *ctx, int64_t out[HMC_CNT_N]);
/* Device time (ms, CUDA events) of the kernels of the last call: [0] forward, [1] traceback, [2] backward, [3] E-step. */
int hmc_get_kernel_ms(hmc_ctx *ctx, float out[4]);

/* Parity hook: per-index live-state sets of one individual from the trellis kept by
 * HMC_FLAG_KEEP_TRELLIS.  counts: [L+1]; the other arrays are concatenated over indices and must
 * hold `capacity` entries; returns the total number of states or a negative status. */
int64_t hmc_get_trellis(hmc_ctx *ctx, int individual, int32_t *counts, int64_t capacity, int32_t *ida, int32_t *idb, double *alpha,
                        double *beta);

#ifdef __cplusplus
}
#endif
#endif /* HMC_B200_H */
