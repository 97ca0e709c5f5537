// TEST INFRASTRUCTURE ONLY -- flat C API over oracle/hmc_oracle.hpp for ctypes (oracle/pyoracle.py).
#include "hmc_oracle.hpp"
#include <chrono>
#include <cstring>

using namespace orc;

struct Handle {
    Panel input;
    HaploModel hm;
    std::vector<Resolution> resolutions;
    std::vector<Pat> estep_pats;
};

extern "C" {

void *orc_create(int N, int L, const int8_t *g)
{
    Handle *h = new Handle;
    h->input.N = N; h->input.L = L;
    h->input.g.assign(g, g + (size_t) N * 2 * L);
    h->input.computeAlleleTable();
    h->hm.setPanel(h->input);
    return h;
}
void orc_destroy(void *p) { delete (Handle *) p; }

void orc_set_options(void *p, const char *model, double min_freq, double min_freq_abs, int num_patterns, int min_len, int max_len,
                     int mc_order, int max_iter, int sample_size, int exact, int scaled)
{
    Options &o = ((Handle *) p)->hm.opt;
    o.model = model; o.min_freq = min_freq; o.min_freq_abs = min_freq_abs; o.num_patterns = num_patterns;
    o.min_pattern_len = min_len; o.max_pattern_len = max_len; o.mc_order = mc_order; o.max_iteration = max_iter;
    o.sample_size = sample_size; o.exact_estimate = exact != 0; o.scaled = scaled != 0;
}
int orc_amax(void *p) { return ((Handle *) p)->hm.panel.Amax; }
void orc_get_allele_table(void *p, int *nall, double *afreq)
{
    Panel &pn = ((Handle *) p)->hm.panel;
    memcpy(nall, pn.nall.data(), sizeof(int) * pn.L);
    memcpy(afreq, pn.afreq.data(), sizeof(double) * pn.afreq.size());
}
// override the allele table (a sub-sample must keep the full panel's alleles and frequencies)
void orc_set_allele_table(void *p, int Amax, const int *nall, const double *afreq)
{
    Handle *h = (Handle *) p;
    for (Panel *pn : {&h->input, &h->hm.panel}) {
        pn->Amax = Amax;
        pn->nall.assign(nall, nall + pn->L);
        pn->afreq.assign(afreq, afreq + (size_t) pn->L * Amax);
    }
}
void orc_find_patterns(void *p) { Handle *h = (Handle *) p; h->hm.samples.clear(); h->hm.findPatterns(); }
int orc_num_patterns(void *p) { return ((Handle *) p)->hm.model.P(); }
int orc_head_len(void *p) { return ((Handle *) p)->hm.model.head_len(); }
int orc_num_heads(void *p) { return (int) ((Handle *) p)->hm.model.heads.size(); }
long long orc_total_pattern_len(void *p)
{
    long long t = 0;
    for (const Pat &q : ((Handle *) p)->hm.model.pats) t += q.len();
    return t;
}
void orc_export_model(void *p, int *start, int *end, double *freq, double *prefix, double *tp, int *succ, long long *al_off, int8_t *al, int *heads)
{
    Model &m = ((Handle *) p)->hm.model;
    const int A = m.panel->Amax;
    long long off = 0;
    for (int i = 0; i < m.P(); ++i) {
        const Pat &q = m.pats[i];
        start[i] = q.start; end[i] = q.end; freq[i] = q.freq; prefix[i] = q.prefix; tp[i] = q.tp;
        for (int j = 0; j < A; ++j) succ[(size_t) i * A + j] = q.successor(j);
        al_off[i] = off;
        memcpy(al + off, q.al.data(), q.al.size());
        off += q.len();
    }
    al_off[m.P()] = off;
    for (size_t i = 0; i < m.heads.size(); ++i) heads[i] = m.heads[i];
}
// replace the model by an explicit pattern list (ids = positions), then rebuild the automaton
void orc_import_model(void *p, int P, const int *start, const int *end, const double *freq, const double *prefix, const double *tp,
                      const long long *al_off, const int8_t *al, int min_len, int max_len, double min_freq)
{
    Handle *h = (Handle *) p;
    Model &m = h->hm.model;
    m.min_len.assign(m.panel->L, min_len);
    m.max_len.assign(m.panel->L, max_len);
    m.min_freq = min_freq;
    m.pats.assign(P, Pat());
    for (int i = 0; i < P; ++i) {
        Pat &q = m.pats[i];
        q.start = start[i]; q.end = end[i]; q.freq = freq[i]; q.prefix = prefix[i]; q.tp = tp[i];
        q.al.assign(al + al_off[i], al + al_off[i + 1]);
    }
    m.initialize();
}
double orc_resolve_all(void *p)
{
    Handle *h = (Handle *) p;
    h->resolutions.assign(h->hm.panel.N, Resolution());
    h->hm.builder.cnt = Counters();
    return h->hm.resolveAll(h->resolutions);
}
void orc_counters(void *p, long long *out4)
{
    Counters &c = ((Handle *) p)->hm.builder.cnt;
    out4[0] = c.steps; out4[1] = c.states; out4[2] = c.max_states; out4[3] = c.arrivals;
}
// HaploBuilder::getLikelihood(const Haplotype &) of n haplotypes [n][L]
void orc_get_likelihood(void *p, int n, const int8_t *haps, double *out)
{
    Handle *h = (Handle *) p;
    for (int q = 0; q < n; ++q) out[q] = h->hm.builder.getLikelihood(haps + (size_t) q * h->hm.panel.L);
}
int orc_num_res(void *p, int i) { return (int) ((Handle *) p)->hm.klists[i].size(); }
// K-list of individual i after orc_resolve_all: priors/posteriors [n], haps [n][2][L]; returns P(G) (scaled) and exponent
void orc_get_individual(void *p, int i, double *pg, int *exp2, double *prior, double *post, int8_t *haps)
{
    Handle *h = (Handle *) p;
    const int L = h->hm.panel.L;
    const std::vector<Resolution> &kl = h->hm.klists[i];
    *pg = h->hm.panel.pg[i];
    *exp2 = h->hm.exps[i];
    for (size_t k = 0; k < kl.size(); ++k) {
        prior[k] = kl[k].prior; post[k] = kl[k].posterior;
        memcpy(haps + (k * 2) * L, kl[k].h[0].data(), L);
        memcpy(haps + (k * 2 + 1) * L, kl[k].h[1].data(), L);
    }
}
void orc_get_resolution(void *p, int i, int8_t *haps)
{
    Handle *h = (Handle *) p;
    const int L = h->hm.panel.L;
    memcpy(haps, h->resolutions[i].h[0].data(), L);
    memcpy(haps + L, h->resolutions[i].h[1].data(), L);
}
// trellis of individual i (fresh resolve with K = sample_size): per-index state counts [L+1]; returns total states
long long orc_trellis_run(void *p, int i, int K, int with_beta, int *counts)
{
    Handle *h = (Handle *) p;
    std::vector<Resolution> rl;
    h->hm.builder.scaled = h->hm.opt.scaled;
    h->hm.builder.resolve(i, rl, K);
    if (with_beta) h->hm.builder.calcBackward();
    long long t = 0;
    for (int l = 0; l <= h->hm.panel.L; ++l) { counts[l] = (int) h->hm.builder.at[l].size(); t += counts[l]; }
    return t;
}
void orc_trellis_get(void *p, int *ida, int *idb, double *alpha, double *tp, double *beta, int *narr)
{
    Handle *h = (Handle *) p;
    size_t k = 0;
    for (int l = 0; l <= h->hm.panel.L; ++l)
        for (const State &s : h->hm.builder.at[l]) {
            ida[k] = s.ida; idb[k] = s.idb; alpha[k] = s.alpha; tp[k] = s.tp; beta[k] = s.beta;
            narr[k] = (int) (s.fwd[0].size() + s.fwd[1].size()) | (s.nin << 8);
            ++k;
        }
}
// E-pass over an explicit candidate list (HaploBuilder::estimateFrequency(vector<HaploPattern*>&)); uses the pg[] of the last resolve_all
void orc_estep(void *p, int Pc, const int *start, const long long *al_off, const int8_t *al, double *freq, double *prefix, double *tp)
{
    Handle *h = (Handle *) p;
    std::vector<Pat> pats(Pc);
    for (int i = 0; i < Pc; ++i) {
        pats[i].start = start[i];
        pats[i].al.assign(al + al_off[i], al + al_off[i + 1]);
        pats[i].end = start[i] + pats[i].len();
    }
    EStep es(h->hm.builder, pats);
    es.run();
    for (int i = 0; i < Pc; ++i) { freq[i] = pats[i].freq; prefix[i] = pats[i].prefix; tp[i] = pats[i].tp; }
}
void orc_estimate_patterns(void *p) { ((Handle *) p)->hm.estimatePatterns(); }
// full HaploModel::run; returns number of iterations executed
int orc_run(void *p)
{
    Handle *h = (Handle *) p;
    h->hm.run(h->input, h->resolutions);
    return (int) h->hm.ll_trace.size();
}
void orc_get_trace(void *p, double *ll, int *npat)
{
    Handle *h = (Handle *) p;
    for (size_t i = 0; i < h->hm.ll_trace.size(); ++i) { ll[i] = h->hm.ll_trace[i]; npat[i] = h->hm.npat_trace[i]; }
}
// HaploComp over allele codes [N][2][L]; out6 = se_num, se_den, ig_num, ig_den, ih_num, ih_den; returns the first inconsistent individual or -1
int orc_compare(const int8_t *real, const int8_t *infer, int N, int L, long long *out6, int *bad_locus)
{
    const orc::CompareResult r = orc::compare_genotypes(real, infer, N, L);
    out6[0] = r.se_num; out6[1] = r.se_den; out6[2] = r.ig_num; out6[3] = r.ig_den; out6[4] = r.ih_num; out6[5] = r.ih_den;
    if (bad_locus) *bad_locus = r.bad_locus;
    return r.bad_individual;
}

double orc_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // extern "C"
