// Flat C API over the host classes for the Python layer (hmc_b200/host.py) -- tests and bench.py drive
// HaploModel through this, exactly as the CLI (hmc_cli.cpp) does.
#include <cstring>
#include <string>

#include "hmc_host.hpp"

using namespace hmc_host;

struct Handle {
    HaploModel *hm = nullptr;
    GenoData input;
    std::vector<int8_t> resolutions;
    std::string err, path;
};

#define GUARD(body)                                                        \
    try { body; return 0; }                                                \
    catch (const std::exception &e) { h->err = e.what(); return -1; }

extern "C" {

void *hmch_create(int device, size_t arena_bytes, int threads, char *errbuf, int errlen)
{
    Handle *h = new Handle;
    try { h->hm = new HaploModel(device, arena_bytes, threads); h->hm->setCliDefaults(); }
    catch (const std::exception &e) { if (errbuf && errlen > 0) { strncpy(errbuf, e.what(), errlen - 1); errbuf[errlen - 1] = 0; } delete h; return nullptr; }
    return h;
}
void hmch_destroy(void *p) { Handle *h = (Handle *) p; if (h) { delete h->hm; delete h; } }
const char *hmch_last_error(void *p) { return ((Handle *) p)->err.c_str(); }
void *hmch_ctx(void *p)
{
    Handle *h = (Handle *) p;
    try { return h->hm->ctx(); } catch (const std::exception &e) { h->err = e.what(); return nullptr; }
}

int hmch_load_phase(void *p, const char *path)
{
    Handle *h = (Handle *) p;
    GUARD(h->path = path; HaploFile(path).readGenoData(h->input));
}
// raw allele values [N][2][L] (-1 missing), types: L characters ('S'/'M')
int hmch_set_panel_raw(void *p, int N, int L, const char *types, const int32_t *raw)
{
    Handle *h = (Handle *) p;
    GUARD(h->input.resize(N, L); h->input.allele_type.assign(types, types + L); memcpy(h->input.raw.data(), raw, sizeof(int32_t) * (size_t) N * 2 * L);
          for (int i = 0; i < N; ++i) h->input.ids[i] = "ind" + std::to_string(i);
          h->input.checkAlleleSymbol());
}
int hmch_set_options(void *p, const char *model, double min_freq, double min_freq_abs, int num_patterns, int min_len, int max_len, int mc_order,
                     int max_iter, int sample_size, int exact, int scaled, int host_threads)
{
    Handle *h = (Handle *) p;
    GUARD(HaploModel &m = *h->hm; m.setModel(model); m.min_freq = min_freq; m.min_freq_abs = min_freq_abs; m.num_patterns = num_patterns;
          m.min_pattern_len = min_len; m.max_pattern_len = max_len; m.mc_order = mc_order; m.max_iteration = max_iter; m.sample_size = sample_size;
          m.exact_estimate = exact != 0; m.setScaled(scaled != 0); m.patternManager().setThreads(host_threads));
}
// pattern search placement: 1 device (K5), 0 host threads, -1 device when a context can be created
int hmch_set_device_mining(void *p, int mode) { Handle *h = (Handle *) p; GUARD(h->hm->patternManager().setDeviceMining(mode)); }
int hmch_mined_on_device(void *p) { Handle *h = (Handle *) p; return h->hm->patternManager().last_mine_on_device ? 1 : 0; }
void hmch_dims(void *p, int *N, int *L, int *Amax) { Handle *h = (Handle *) p; *N = h->input.N; *L = h->input.L; *Amax = h->input.Amax; }
void hmch_get_panel(void *p, int8_t *idx, int32_t *n_alleles, double *allele_freq)
{
    Handle *h = (Handle *) p;
    memcpy(idx, h->input.idx.data(), h->input.idx.size());
    memcpy(n_alleles, h->input.n_alleles.data(), sizeof(int32_t) * h->input.L);
    memcpy(allele_freq, h->input.allele_freq.data(), sizeof(double) * h->input.allele_freq.size());
}
// HaploModel::build on a working copy of the panel (mining + automaton + uploads)
int hmch_build(void *p)
{
    Handle *h = (Handle *) p;
    GUARD(h->hm->working() = h->input; h->hm->samples()->clear(); h->hm->build(h->hm->working()));
}
int hmch_find_patterns(void *p) { Handle *h = (Handle *) p; GUARD(h->hm->findPatterns()); }
int hmch_upload(void *p) { Handle *h = (Handle *) p; GUARD(h->hm->uploadModel()); }
int hmch_run(void *p) { Handle *h = (Handle *) p; GUARD(h->hm->run(h->input, h->resolutions)); }
int hmch_resolve_all(void *p, double *ll)
{
    Handle *h = (Handle *) p;
    GUARD(*ll = h->hm->resolveAll(h->resolutions));
}
int hmch_estimate_patterns(void *p) { Handle *h = (Handle *) p; GUARD(h->hm->patternManager().estimatePatterns()); }
int hmch_num_patterns(void *p) { return ((Handle *) p)->hm->pattern_num(); }
int hmch_head_len(void *p) { return ((Handle *) p)->hm->patternManager().head_len(); }
int hmch_num_heads(void *p) { return (int) ((Handle *) p)->hm->patternManager().head_list().size(); }
long long hmch_total_pattern_len(void *p)
{
    long long t = 0;
    for (const Pattern &q : ((Handle *) p)->hm->patternManager().patterns().pats) t += q.length();
    return t;
}
void hmch_export_model(void *p, int32_t *start, int32_t *end, double *freq, double *prefix, double *tp, int32_t *succ, int64_t *al_off, int8_t *al, int32_t *heads)
{
    Handle *h = (Handle *) p;
    PatternManager &pm = h->hm->patternManager();
    const PatternSet &ps = pm.patterns();
    int64_t off = 0;
    for (size_t i = 0; i < ps.size(); ++i) {
        const Pattern &q = ps.pats[i];
        start[i] = q.start; end[i] = q.end; freq[i] = q.frequency; prefix[i] = q.prefix_freq; tp[i] = q.transition_prob;
        al_off[i] = off;
        memcpy(al + off, ps.al(q), q.length());
        off += q.length();
    }
    al_off[ps.size()] = off;
    memcpy(succ, pm.successors().data(), sizeof(int32_t) * pm.successors().size());
    memcpy(heads, pm.head_list().data(), sizeof(int32_t) * pm.head_list().size());
}
void hmch_get_resolutions(void *p, int8_t *out) { Handle *h = (Handle *) p; memcpy(out, h->resolutions.data(), h->resolutions.size()); }
int hmch_trace_len(void *p) { return (int) ((Handle *) p)->hm->ll_trace.size(); }
void hmch_get_trace(void *p, double *ll, int *npat)
{
    Handle *h = (Handle *) p;
    for (size_t i = 0; i < h->hm->ll_trace.size(); ++i) { ll[i] = h->hm->ll_trace[i]; npat[i] = h->hm->npat_trace[i]; }
}
void hmch_get_seconds(void *p, double *out3) { Handle *h = (Handle *) p; for (int i = 0; i < 3; ++i) out3[i] = h->hm->seconds[i]; }
int hmch_set_shard(void *p, int first, int count) { Handle *h = (Handle *) p; GUARD(h->hm->setShard(first, count)); }
typedef void (*hmch_reduce_fn)(double *buf, size_t n);
typedef void (*hmch_gather_fn)(const void *send, void *recv_all, size_t elem_bytes);
int hmch_set_collectives(void *p, hmch_reduce_fn reduce, hmch_gather_fn gather)
{
    Handle *h = (Handle *) p;
    GUARD(if (reduce) h->hm->reduce_counts = [reduce](double *b, size_t n) { reduce(b, n); }; else h->hm->reduce_counts = nullptr;
          if (gather) h->hm->gather = [gather](const void *s, void *r, size_t e) { gather(s, r, e); }; else h->hm->gather = nullptr);
}
typedef void (*hmch_reduce_dev_fn)(void *dev_buf, size_t n);
typedef void (*hmch_gather_dev_fn)(void *dev_base, size_t row_bytes);
// device-side exchanges (see HaploBuilder::reduce_counts_device / gather_device); NULL removes one
int hmch_set_device_collectives(void *p, hmch_reduce_dev_fn reduce, hmch_gather_dev_fn gather)
{
    Handle *h = (Handle *) p;
    GUARD(if (reduce) h->hm->reduce_counts_device = [reduce](double *b, size_t n) { reduce((void *) b, n); }; else h->hm->reduce_counts_device = nullptr;
          if (gather) h->hm->gather_device = [gather](void *b, size_t e) { gather(b, e); }; else h->hm->gather_device = nullptr);
}
int hmch_write_reconstructed(void *p, const char *path, const char *suffix)
{
    Handle *h = (Handle *) p;
    GUARD(h->hm->writeReconstructed(HaploFile(path), h->input, h->resolutions, suffix));
}

}  // extern "C"
