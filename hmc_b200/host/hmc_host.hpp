// Host side of hmc_b200: the reference's HMC classes (GenoData, HaploFile, PatternManager,
// HaploBuilder, HaploModel) re-written Boost-free around the C-ABI device path (include/hmc_b200.h).
// Same names, option fields, defaults and error behaviour as the reference; data is held in flat
// index arrays (allele indices per locus) instead of Allele/Haplotype/Genotype objects.
#pragma once

#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hmc_b200.h"

namespace hmc_host {

struct HmcError : std::runtime_error {      // the reference calls Logger::error + exit(1); we throw, main() reports and exits 1
    explicit HmcError(const std::string &m) : std::runtime_error(m) {}
};

// GenoData (GenoData.h:17-66): the panel.  Raw allele values (-1 missing) plus, after checkAlleleSymbol(),
// the per-locus symbol table (sorted ascending, GenoData.cpp:96-98), its frequencies and the index form.
class GenoData {
public:
    int N = 0, L = 0;
    std::string allele_type;                    // one char per locus: 'S' or 'M'
    std::vector<int> position;                  // allele_postition
    std::vector<std::string> ids;
    std::vector<int32_t> raw;                   // [N][2][L] allele values as read
    // derived
    int Amax = 0;
    std::vector<int32_t> n_alleles;             // [L]
    std::vector<std::vector<int32_t> > symbols; // [L][n_alleles]
    std::vector<double> allele_freq;            // [L][Amax]
    std::vector<int8_t> idx;                    // [N][2][L] allele indices, -1 missing
    std::vector<double> genotype_probability;   // [N], written by resolveAll (HaploModel.cpp:109)
    // bit-packed indices in the device layout (filled by the binary reader; hmc_set_panel_packed takes them as they are)
    int packed_bits = 0, packed_words = 0;
    std::vector<uint32_t> packed;               // [N][packed_words]

    void resize(int n, int l);
    int &allele(int i, int h, int l) { return raw[((size_t) i * 2 + h) * L + l]; }
    int allele(int i, int h, int l) const { return raw[((size_t) i * 2 + h) * L + l]; }
    int getAlleleIndex(int locus, int value) const;     // GenoData.cpp:27-44
    void checkAlleleSymbol();                           // GenoData.cpp:78-118
};

// HaploFile (PHASE format only; HaploFile.cpp:54-153)
class HaploFile {
    std::string m_filename;
public:
    explicit HaploFile(const std::string &filename) : m_filename(filename) {}
    const std::string &filename() const { return m_filename; }
    void readGenoData(GenoData &genos) const;   // PHASE text, or the binary panel format below (recognised by its magic)
    // Binary panel ("HMCB"): header, per-locus symbol tables and frequencies, ids, then the genotypes bit-packed exactly as the
    // device holds them (2 x bits per locus, code 0 = missing), so ingest is one read + one H2D copy and no text parsing
    // (PHASE text caps L at ~200 k characters per line, HaploFile.cpp:10, and is tens of GB at C5 scale).
    static bool isBinary(const std::string &path);
    void readBinary(GenoData &genos) const;
    static void writeBinary(const std::string &path, const GenoData &genos, const std::vector<int8_t> &idx);
    // writes allele INDICES `res` ([N][2][L], mapped back through the symbol table) to filename + suffix
    void writeGenoData(const GenoData &genos, const std::vector<int8_t> &res, const char *suffix = "") const;
    // the same file from allele lines that arrive formatted (hmc_format_best): line 2 * i + h of `text` is haplotype h of individual i
    void writeGenoText(const GenoData &genos, const char *text, const int64_t *line_off, const char *suffix = "") const;
};

// HaploPattern (HaploPattern.h:19-24) as a plain record; alleles live in the manager's pool.
struct Pattern {
    int start = 0, end = 0;
    int64_t al = 0;                 // offset into PatternSet::alleles
    double frequency = 1.0, prefix_freq = 1.0, transition_prob = 1.0;
    int length() const { return end - start; }
    void setTransitionProb(double p) { transition_prob = p < 1.0 ? p : 1.0; }   // HaploPattern.h:47
};

struct PatternSet {
    std::vector<Pattern> pats;
    std::vector<int8_t> alleles;    // concatenated allele indices
    const int8_t *al(const Pattern &p) const { return alleles.data() + p.al; }
    void clear() { pats.clear(); alleles.clear(); }
    int add(int start, const int8_t *a, int len, double f, double pf, double tp);
    size_t size() const { return pats.size(); }
};

// Weighted sampled haplotypes (HaploData as used by the sampling M-step, HaploModel.cpp:94-108)
struct HaploData {
    int L = 0;
    std::vector<int8_t> h;      // [n][L]; empty while the haplotypes are only on the device (`rows` names them)
    std::vector<double> w;
    std::vector<int32_t> rows;  // device-resident samples: row of each haplotype in the result block of the last resolve pass
    double total_weight = 0;
    bool empty() const { return w.empty(); }
    bool on_device() const { return !rows.empty() && h.empty(); }
    void clear() { h.clear(); w.clear(); rows.clear(); total_weight = 0; }
};

class HaploBuilder;

// PatternManager (PatternManager.h:18-79): pattern mining, automaton build, M-step glue.
class PatternManager {
    HaploBuilder &m_builder;
    PatternSet m_patterns;
    double m_min_freq = 0;
    std::vector<int> m_min_len, m_max_len;
    mutable std::vector<int32_t> m_succ;        // [P][Amax]
    mutable std::vector<int32_t> m_head_list;
    mutable bool m_host_automaton_stale = false;    // the automaton was built on the device (K6); host copies are fetched on demand
    int m_dev_heads = 0;
    bool m_dev_list_fresh = false;      // m_patterns is exactly the list of the last hmc_mine_patterns call
    void syncAutomaton() const;
    int m_threads = 0;
    int m_device_mining = -1;           // 1 device (K5), 0 host threads, -1 device when a B200 context can be created

public:
    explicit PatternManager(HaploBuilder &hb) : m_builder(hb) {}
    const PatternSet &patterns() const { return m_patterns; }
    const std::vector<int32_t> &successors() const { syncAutomaton(); return m_succ; }
    const std::vector<int32_t> &head_list() const { syncAutomaton(); return m_head_list; }
    bool last_automaton_on_device = false;
    int head_len() const { return m_min_len[0]; }
    int size() const { return (int) m_patterns.size(); }
    double min_freq() const { return m_min_freq; }
    void setThreads(int t) { m_threads = t; }
    void setDeviceMining(int mode) { m_device_mining = mode; }
    bool deviceMiningWanted() const;    // device mining enabled and not overridden by HMC_B200_HOST_MINING
    double last_mine_seconds = 0;       // wall time of the last searchPattern
    bool last_mine_on_device = false;

    void findPatternByFreq(double min_freq, int min_len = 2, int max_len = 0);
    void findPatternByNum(int max_num, int min_len = 2, int max_len = 0);
    void findPatternBlock(int len);
    void estimateFrequency();       // PatternManager.cpp:347-362
    void estimatePatterns();        // PatternManager.cpp:364-410
    // replace the pattern list (tests / import) and rebuild the automaton
    void setPatterns(const PatternSet &ps, int min_len, int max_len, double min_freq);

protected:
    void setLens(int min_len, int max_len);
    void searchPattern(bool by_num, int max_num);
    bool searchPatternDevice();     // hmc_mine_patterns; false when no device context is available in auto mode
    void initialize();              // PatternManager.cpp:293-318
    void extendPatterns(PatternSet &patterns, std::vector<int> &seeds, const PatternSet &last, const std::vector<int> &last_seeds);
};

// HaploBuilder (HaploBuilder.h:20-68) backed by the device context.
class HaploBuilder {
protected:
    GenoData *m_genos = nullptr;
    PatternManager m_patterns;
    HaploData m_samples;
    hmc_ctx *m_ctx = nullptr;
    int m_device = 0, m_dev_threads = 0;
    size_t m_arena_bytes = 0;
    bool m_model_dirty = true, m_panel_dirty = true;
    unsigned m_flags = 0;
    int m_shard_first = 0, m_shard_count = -1;     // this rank's block of individuals (-1 = all)

public:
    explicit HaploBuilder(int device = 0, size_t arena_bytes = 0, int threads = 0);
    virtual ~HaploBuilder();
    HaploBuilder(const HaploBuilder &) = delete;

    GenoData *genos() { return m_genos; }
    const GenoData *genos() const { return m_genos; }
    HaploData *samples() { return &m_samples; }
    const HaploData *samples() const { return &m_samples; }
    PatternManager &patternManager() { return m_patterns; }
    int pattern_num() const { return m_patterns.size(); }
    int genotype_num() const { return m_genos->N; }
    int genotype_len() const { return m_genos->L; }
    hmc_ctx *ctx();                             // created on first use; throws without a B200 (there is no CPU path)
    void setScaled(bool on) { m_flags = on ? (m_flags | HMC_FLAG_SCALED) : (m_flags & ~HMC_FLAG_SCALED); }
    unsigned flags() const { return m_flags; }
    void markModelDirty() { m_model_dirty = true; }
    void markModelClean() { m_model_dirty = false; }     // the device built and installed the automaton itself (hmc_build_automaton)

    // Multi-GPU (one process per GPU): individuals [first, first+count) are this rank's; the collectives are supplied by the
    // launcher (torch.distributed / NCCL / MPI).  reduce_counts: in-place sum over ranks of n doubles.  gather: every rank
    // contributes its individuals' records (elem_bytes each, contiguous) and receives all N records in individual order.
    void setShard(int first, int count) { m_shard_first = first; m_shard_count = count; }
    std::function<void(double *, size_t)> reduce_counts;
    std::function<void(const void *, void *, size_t)> gather;
    // The same two exchanges on DEVICE memory (preferred when set: nothing is staged through the host).  reduce_counts_device: in-place
    // sum over ranks of n doubles at a device pointer (ncclAllReduce).  gather_device: in-place all-gather over the individuals of a
    // device array with one row of row_bytes per panel individual -- this rank's rows [first, first+count) are valid on entry, all N on
    // return (ncclAllGather / grouped broadcasts).  Both run after the context's stream has been synchronised.
    std::function<void(double *, size_t)> reduce_counts_device;
    std::function<void(void *, size_t)> gather_device;
    bool sharded() const { return m_shard_count >= 0 && m_shard_count < (m_genos ? m_genos->N : 0); }

    void setGenoData(GenoData &genos);          // HaploBuilder.cpp:19-23
    void uploadPanel();                         // hmc_set_panel when the panel changed
    void uploadModel();                         // hmc_set_model when the automaton changed

    struct ResolveResult {
        int K = 0;
        std::vector<double> pg, coverage, prior, posterior;
        std::vector<int32_t> exp2, nres;
        std::vector<int8_t> haps;               // [N][K][2][L]
        double ll = 0;
    };
    // resolve() of every individual in one call (HaploModel::resolveAll drives this)
    void resolveAll(int sample_size, ResolveResult &out, bool want_haps = true);
    // device-resident samples -> host copies (only the host miner needs them)
    void materializeSamples();
    // HaploBuilder::estimateFrequency(vector<HaploPattern*>&): fills frequency / prefix_freq / transition_prob
    void estimateFrequency(PatternSet &patterns);
    std::vector<int32_t> pg_exp;                // exponents carried next to genos()->genotype_probability (scaled mode)
};

// HaploComp (HaploComp.h:10-45, HaploComp.cpp:29-76) on the device (K7): switch error, incorrect-haplotype and
// incorrect-genotype percentages of `infer` against `real`.
struct HaploComp {
    int64_t counts[6] = {0, 0, 0, 0, 0, 0};     // numerators / denominators in the order SE, IGP, IHP
    double switch_error() const { return (double) counts[0] / counts[1]; }
    double incorrect_genotype_percentage() const { return (double) counts[2] / counts[3]; }
    double incorrect_haplotype_percentage() const { return (double) counts[4] / counts[5]; }
    // both sides as allele INDICES of the same panel ([N][2][L], -1 missing); infer == nullptr: the best resolutions of the
    // last resolve pass, compared where they lie in device memory
    HaploComp(HaploBuilder &hb, int N, int L, const int8_t *real, const int8_t *infer);
    // two panels with their own symbol tables (HMC::compare, HMC.cpp:246-257): a joint per-locus coding is built first
    HaploComp(HaploBuilder &hb, const GenoData &real, const GenoData &infer);
};

// HaploModel (HaploModel.h:10-40): option fields + EM driver.
class HaploModel : public HaploBuilder {
protected:
    std::string m_model = "MV";

public:
    // HaploModel.cpp:11-25 defaults; HMC.cpp:35-47 are the CLI defaults
    double min_freq = -1;
    double min_freq_abs = -1;
    int num_patterns = -1;
    int min_pattern_len = 1;
    int max_pattern_len = -1;
    int mc_order = 1;
    int max_iteration = -1;
    int min_iteration = -1;
    int sample_size = 1;
    int max_sample_size = 1;
    int final_sample_size = 1;
    bool exact_estimate = false;

    std::vector<double> ll_trace;
    std::vector<int> npat_trace;
    std::vector<double> seconds;        // [0] mining+automaton, [1] device resolve/E-step, [2] host M-step glue
    bool verbose = false;

    explicit HaploModel(int device = 0, size_t arena_bytes = 0, int threads = 0) : HaploBuilder(device, arena_bytes, threads), seconds(3, 0.0) {}
    void setModel(const std::string &model);
    void setCliDefaults();      // HMC.cpp:35-47
    // HaploModel::run (HaploModel.cpp:117-155); resolutions: [N][2][L] allele indices
    void run(const GenoData &genos, std::vector<int8_t> &resolutions);
    // `.reconstructed` (HMC.cpp:228): the allele lines are formatted on the device when it holds the best resolution of every
    // individual (K8b, hmc_format_best); `res` (what run() returned) is written through the host formatter otherwise.  Returns
    // true when the device formatted the text.
    bool m_device_rows_current = false;     // the best rows of the last device resolve are what run() returned
    bool writeReconstructed(const HaploFile &file, const GenoData &genos, const std::vector<int8_t> &res, const char *suffix = ".reconstructed");
    GenoData &working() { return m_unphased; }
    void build(GenoData &genos);
    void findPatterns();
    double resolveAll(std::vector<int8_t> &resolved);

protected:
    GenoData m_unphased;        // kept alive for the builder (the reference hands out a pointer to a local)
};

}  // namespace hmc_host
