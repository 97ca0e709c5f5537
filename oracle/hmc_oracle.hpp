// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle port") of the reference HMC hot path.
// Nothing under hmc_b200/ (the product) may include, link or execute this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
//
// Parity status: PINNED.  Every function below is checked bit-for-bit (structure, ids, state
// order, alpha, P(G), K-lists including libstdc++ nth_element/sort tie behaviour) against the
// unmodified reference built by oracle/build_ref.sh (tests/test_oracle_vs_ref.py, run in the build
// container where /root/reference exists) and against the committed dumps of that binary under
// tests/golden/ (tests/test_oracle_golden.py, runs anywhere).  The reference itself ships no golden
// vectors, tests or fixtures, so the reference binary's own output is the anchor.
//
// Layout: flat index-based arrays instead of the reference's heap objects; every routine cites the
// reference lines it restates.  Alleles are per-locus allele INDICES (GenoData.cpp:27-44 order, i.e.
// symbols sorted ascending, GenoData.cpp:96-98); -1 is the missing allele (Allele.h:13,17).
//
// One deliberate extension: `scaled` mode rescales alpha and the path likelihoods by 2^512
// whenever the per-locus maximum alpha drops below 2^-512 (exponent-only, mantissas untouched), and
// carries the exponent per individual.  The reference has no rescaling and underflows to P(G)=0
// for L >~ 2000.  With scaling off the port is the reference, bit for bit.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Panel: genotypes as allele indices + the per-locus allele table (GenoData.h:43-49).
struct Panel {
    int N = 0, L = 0, Amax = 0;
    std::vector<int8_t> g;          // [N][2][L]
    std::vector<int> nall;          // [L]      GenoData::allele_num
    std::vector<double> afreq;      // [L][Amax] GenoData::allele_frequency
    std::vector<double> pg;         // [N] genotype_probability written by resolveAll (HaploModel.cpp:109)

    int al(int i, int h, int l) const { return g[((size_t) i * 2 + h) * L + l]; }
    double af(int l, int j) const { return afreq[(size_t) l * Amax + j]; }
    bool isMissing(int i, int l) const { return al(i, 0, l) < 0 && al(i, 1, l) < 0; }      // Genotype.h:115
    bool hasMissing(int i, int l) const { return al(i, 0, l) < 0 || al(i, 1, l) < 0; }     // Genotype.h:120
    bool hasAllele(int i, int l, int a) const { return al(i, 0, l) == a || al(i, 1, l) == a; } // Genotype.h:125 (a is never missing here)
    static bool amatch(int x, int y) { return x < 0 || y < 0 || x == y; }                   // Allele.h:17
    bool matchAllele(int i, int l, int a) const { return amatch(al(i, 0, l), a) || amatch(al(i, 1, l), a); } // Genotype.h:130
    bool isHet(int i, int l) const { return !amatch(al(i, 0, l), al(i, 1, l)); }            // Genotype.h:64

    // GenoData::checkAlleleSymbol (GenoData.cpp:78-118) on values already mapped to indices:
    // frequency = count / (non-missing alleles at the locus), haplotype weight 1.
    void computeAlleleTable()
    {
        Amax = 0;
        nall.assign(L, 0);
        for (size_t k = 0; k < g.size(); ++k) if (g[k] >= 0) { int l = (int) (k % L); nall[l] = std::max(nall[l], g[k] + 1); }
        for (int l = 0; l < L; ++l) Amax = std::max(Amax, nall[l]);
        afreq.assign((size_t) L * Amax, 0.0);
        std::vector<double> tot(L, 0.0);
        for (int i = 0; i < N; ++i) for (int h = 0; h < 2; ++h) for (int l = 0; l < L; ++l) {
            int a = al(i, h, l);
            if (a >= 0) { afreq[(size_t) l * Amax + a] += 1.0; tot[l] += 1.0; }
        }
        for (int l = 0; l < L; ++l) for (int j = 0; j < nall[l]; ++j) afreq[(size_t) l * Amax + j] /= tot[l];
        pg.assign(N, 0.0);
    }
};

// HaploData as used by the sampling M-step (HaploModel.cpp:94-108): weighted sampled haplotypes.
struct Samples {
    int L = 0;
    std::vector<int8_t> h;      // [n][L]
    std::vector<double> w;      // Haplotype::weight
    double total = 0;           // HaploData::checkTotalWeight (HaploData.cpp:120-126)
    bool empty() const { return w.empty(); }
    void clear() { h.clear(); w.clear(); total = 0; }
    int size() const { return (int) w.size(); }
    const int8_t *hap(int i) const { return &h[(size_t) i * L]; }
};

// HaploPattern (HaploPattern.h:19-24): allele string on [start,end), frequencies, successor table.
struct Pat {
    int start = 0, end = 0;
    std::vector<int8_t> al;
    double freq = 1.0, prefix = 1.0, tp = 1.0;      // HaploPattern.h:105-107 defaults
    std::vector<int> succ;                          // by allele index at locus `end`; -1 = NULL
    int len() const { return (int) al.size(); }
    void setTp(double p) { tp = p < 1.0 ? p : 1.0; }    // HaploPattern.h:47
    int successor(int j) const { return (j >= 0 && j < (int) succ.size()) ? succ[j] : -1; }  // HaploPattern.h:36
};

// BackwardPatternTree (PatternTree.cpp:14-134): one trie per end locus, keyed by alleles walking backwards.
struct BackTrie {
    int A = 0;
    std::vector<int> child;     // [node][A]
    std::vector<int> data;      // pattern id or -1
    std::vector<int> root;      // [L+1]
    int newNode() { data.push_back(-1); child.insert(child.end(), A, -1); return (int) data.size() - 1; }
    void init(int L, int Amax) { A = Amax; child.clear(); data.clear(); root.resize(L + 1); for (int i = 0; i <= L; ++i) root[i] = newNode(); }
    int getChild(int n, int i) const { return child[(size_t) n * A + i]; }
    int addChild(int n, int i) { int c = child[(size_t) n * A + i]; if (c < 0) { c = newNode(); child[(size_t) n * A + i] = c; } return c; }
    // PatternTree.cpp:24-48 (patterns never hold a missing allele, so only the else-branch is live)
    void addPattern(const Pat &p, int id)
    {
        int node = root[p.end];
        for (int k = p.len() - 1; k >= 0; --k) {
            node = addChild(node, p.al[k]);
            if (k == 0) data[node] = id;    // setChild(i, hp), Tree.h:87-98
        }
    }
    // PatternTree.cpp:98-134. `as` is indexed by ll (local), lg is the global locus (only used for the
    // wildcard fan-out), len bounds the depth.
    int findLongest(int node, const int8_t *as, const std::vector<Pat> &pats, const std::vector<int> &nall, int lg, int ll, int len) const
    {
        int result = data[node];
        auto consider = [&](int i) {
            int c = getChild(node, i);
            if (c < 0) return;
            int temp = len > 1 ? findLongest(c, as, pats, nall, lg - 1, ll - 1, len - 1) : data[c];
            if (result < 0 || (temp >= 0 && pats[temp].len() > pats[result].len())) result = temp;
        };
        if (as[ll] < 0) { for (int i = 0; i < nall[lg]; ++i) consider(i); }
        else consider(as[ll]);
        return result;
    }
    // PatternTree.cpp:62-72
    int findLongestMatch(int end, const int8_t *as, int as_len, int as_start, const std::vector<Pat> &pats, const std::vector<int> &nall, int len = 0) const
    {
        if (end <= as_start || end > as_start + as_len) return -1;
        int max_len = end - as_start;
        if (len <= 0 || len > max_len) len = max_len;
        return findLongest(root[end], as, pats, nall, end - 1, max_len - 1, len);
    }
};

struct Options {            // HaploModel.cpp:11-25 fields with the CLI defaults of HMC.cpp:35-47
    std::string model = "MV";
    double min_freq = -1, min_freq_abs = 1.5;
    int num_patterns = -1, min_pattern_len = 1, max_pattern_len = 30, mc_order = 1;
    int max_iteration = 1, sample_size = 10;
    bool exact_estimate = false;
    bool scaled = false;    // extension, see header
};

typedef std::vector<std::pair<int, double> > MatchingState;     // PatternManager.h:19

// ---------------------------------------------------------------------------------------------
// PatternManager (mining + automaton) -- PatternManager.cpp
struct Model {
    const Panel *panel = nullptr;
    const Samples *samples = nullptr;
    std::vector<Pat> pats;
    double min_freq = 0;
    std::vector<int> min_len, max_len;
    std::vector<int> heads;
    BackTrie trie;
    struct Cand { Pat p; MatchingState ms; };
    std::vector<Cand> cands;

    int head_len() const { return min_len[0]; }
    int P() const { return (int) pats.size(); }

    // PatternManager.cpp:267-291
    double getMatchingFrequency(int gi, const int8_t *pa, int start, int len) const
    {
        double total_freq = 1.0;
        for (int i = 0; i < len; ++i) {
            if (pa[i] >= 0) {
                double freq = 0;
                for (int j = 0; j < 2; ++j) {
                    int b = panel->al(gi, j, start + i);
                    if (b < 0) freq += panel->af(start + i, pa[i]);
                    else if (b == pa[i]) freq += 1.0;
                }
                total_freq *= 0.5 * freq;
            }
        }
        return total_freq;
    }
    bool genoMatch(int gi, const Pat &p, int start, int len) const     // Genotype::isMatch(as,...) Genotype.cpp:87-103
    {
        if (start + len > panel->L || start - p.start + len > p.len()) return false;
        for (int i = 0; i < len; ++i) if (!panel->matchAllele(gi, start + i, p.al[start - p.start + i])) return false;
        return true;
    }
    bool hapMatch(const int8_t *h, const Pat &p, int start, int len) const  // AlleleSequence::isMatch Allele.cpp:10-27
    {
        if (start + len > samples->L || start - p.start + len > p.len()) return false;
        for (int i = 0; i < len; ++i) if (!Panel::amatch(h[start + i], p.al[start - p.start + i])) return false;
        return true;
    }
    // PatternManager.cpp:146-193 (isPhased() is never true)
    void checkFrequency(Pat &hp, MatchingState &ms) const
    {
        ms.clear();
        if (hp.len() == 0) { hp.freq = 1.0; return; }
        double total_freq = 0;
        if (samples->empty()) {
            for (int i = 0; i < panel->N; ++i) {
                if (genoMatch(i, hp, hp.start, hp.len())) {
                    double freq = getMatchingFrequency(i, hp.al.data(), hp.start, hp.len());
                    total_freq += freq;
                    ms.push_back(std::make_pair(i, freq));
                }
            }
            hp.freq = total_freq / panel->N;
        }
        else {
            for (int i = 0; i < samples->size(); ++i) {
                if (hapMatch(samples->hap(i), hp, hp.start, hp.len())) { total_freq += samples->w[i]; ms.push_back(std::make_pair(i, 0.0)); }
            }
            hp.freq = total_freq / samples->total;
        }
    }
    // PatternManager.cpp:195-265
    void checkFrequencyWithExtension(Pat &hp, MatchingState &ms, const MatchingState &oms, int start, int len = 1) const
    {
        if (hp.len() == 0 || oms.empty()) { checkFrequency(hp, ms); return; }
        ms.clear();
        double total_freq = 0;
        if (samples->empty()) {
            for (size_t k = 0; k < oms.size(); ++k) {
                int gi = oms[k].first;
                double freq = oms[k].second;
                if (genoMatch(gi, hp, start, len)) {
                    freq *= getMatchingFrequency(gi, &hp.al[start - hp.start], start, len);
                    total_freq += freq;
                    ms.push_back(std::make_pair(gi, freq));
                }
            }
            hp.freq = total_freq / panel->N;
        }
        else {
            for (size_t k = 0; k < oms.size(); ++k) {
                if (hapMatch(samples->hap(oms[k].first), hp, start, len)) { total_freq += samples->w[oms[k].first]; ms.push_back(std::make_pair(oms[k].first, 0.0)); }
            }
            hp.freq = total_freq / samples->total;
        }
    }
    void generateCandidates()   // PatternManager.cpp:90-98
    {
        cands.clear();
        for (int i = 0; i < panel->L; ++i) { Cand c; c.p.start = c.p.end = i; cands.push_back(c); }
    }
    // PatternManager.cpp:100-144
    void searchPattern(bool reserve_candidates = false)
    {
        const int L = panel->L;
        std::vector<Cand> new_candidates;
        while (!cands.empty()) {
            Cand pc;
            std::swap(pc, cands.back());
            cands.pop_back();
            const Pat &hp = pc.p;
            if (hp.freq >= min_freq || hp.len() < min_len[hp.start]) {
                if (hp.end < L && hp.len() < max_len[hp.start]) {
                    for (int i = 0; i < panel->nall[hp.end]; ++i) {
                        if (panel->af(hp.end, i) > 0) {
                            Cand nc;
                            nc.p.start = hp.start; nc.p.end = hp.end + 1; nc.p.al = hp.al; nc.p.al.push_back((int8_t) i);   // HaploPattern::assign
                            nc.p.freq = hp.freq;
                            checkFrequencyWithExtension(nc.p, nc.ms, pc.ms, hp.end);
                            nc.p.prefix = hp.freq;
                            if (nc.p.prefix > 0) nc.p.setTp(nc.p.freq / nc.p.prefix);
                            else nc.p.setTp(nc.p.freq);
                            cands.push_back(std::move(nc));
                        }
                    }
                }
            }
            if (hp.freq >= min_freq || hp.len() <= min_len[hp.start]) {
                if (hp.len() > 0 && hp.len() >= min_len[hp.start]) pats.push_back(pc.p);
            }
            else if (reserve_candidates) new_candidates.push_back(std::move(pc));
        }
        cands.swap(new_candidates);
    }
    void setLens(int mn, int mx)
    {
        // vector::resize(n, v) only fills NEW elements (PatternManager.cpp:33-34): lengths stick after the first call
        min_len.resize(panel->L, mn);
        max_len.resize(panel->L, mx);
    }
    void findPatternByFreq(double mf, int mn = 2, int mx = 0)   // PatternManager.cpp:27-42
    {
        mx = mx <= 0 ? panel->L : mx;
        mn = std::max(mn, 1);
        mx = std::max(mx, mn);
        setLens(mn, mx);
        pats.clear();
        generateCandidates();
        min_freq = mf;
        searchPattern();
        cands.clear();
        initialize();
    }
    void findPatternByNum(int max_num, int mn = 2, int mx = 0)  // PatternManager.cpp:44-73
    {
        int last_size = 0;
        mx = mx <= 0 ? panel->L : mx;
        mn = std::max(mn, 1);
        mx = std::max(mx, mn);
        setLens(mn, mx);
        pats.clear();
        generateCandidates();
        min_freq = 1.0;
        searchPattern(true);
        max_num = std::max(max_num, (int) pats.size());
        while ((int) pats.size() < max_num && min_freq > 1e-38) {
            last_size = (int) pats.size();
            min_freq *= 0.9;
            searchPattern(true);
        }
        if ((int) pats.size() > max_num) {
            std::sort(pats.begin() + last_size, pats.end(), [](const Pat &a, const Pat &b) { return a.freq > b.freq; });
            pats.resize(max_num);
        }
        cands.clear();
        initialize();
    }
    void findPatternBlock(int len)  // PatternManager.cpp:75-88
    {
        len = std::max(1, len);
        setLens(len, len);
        pats.clear();
        generateCandidates();
        min_freq = -1.0;
        searchPattern();
        cands.clear();
        initialize();
    }
    // PatternManager.cpp:293-318
    void initialize()
    {
        const int L = panel->L;
        trie.init(L, panel->Amax);
        heads.clear();
        for (int i = 0; i < P(); ++i) {
            trie.addPattern(pats[i], i);
            if (pats[i].start == 0 && pats[i].len() == head_len()) heads.push_back(i);
        }
        std::vector<int8_t> temp;
        for (int i = 0; i < P(); ++i) {
            Pat &hp = pats[i];
            hp.succ.clear();
            if (hp.end < L) {
                temp.assign(hp.al.begin(), hp.al.end());
                temp.push_back(-1);
                hp.succ.assign(panel->nall[hp.end], -1);
                for (int j = 0; j < panel->nall[hp.end]; ++j) {
                    temp.back() = (int8_t) j;
                    hp.succ[j] = trie.findLongestMatch(hp.end + 1, temp.data(), (int) temp.size(), hp.start, pats, panel->nall);
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// HaploPair trellis (HaploPair.h:14-52,55-66)
struct Link {
    int link;           // predecessor state index at the previous locus, -1 at the head (HaploPairLink::link)
    int index;
    bool reversed, homozygous;
    double likelihood;
};
struct LinkGreater { bool operator()(const Link &a, const Link &b) const { return a.likelihood > b.likelihood; } };   // HaploPair.h:47-50

struct State {
    int ida, idb;
    double tp, alpha, beta;
    int nin = 0;                    // arrivals (addHaploPair calls) into this state -- statistics only
    std::vector<Link> links;        // m_best_links
    std::vector<int> fwd[2];        // m_forward_links[reversed]
};

struct Resolution {
    std::vector<int8_t> h[2];
    double prior = 0, posterior = 0, pg = 0;
};

struct Counters { long long steps = 0, states = 0, max_states = 0, arrivals = 0; };

// HaploBuilder (HaploBuilder.cpp) over a Model
struct Builder {
    const Panel *panel;
    Model *model;
    bool scaled = false;
    int sample_size = 1;
    std::vector<std::vector<State> > at;    // m_haplopairs, index 0..L
    std::vector<int> scale_exp;             // cumulative power-of-two exponent applied up to each index (scaled mode)
    int final_exp = 0;
    Counters cnt;

    Builder(const Panel *p, Model *m) : panel(p), model(m) {}

    int lastA(int id) const { return model->pats[id].al.back(); }

    // HaploBuilder.cpp:128-146 (dead in the reference: its only callers are commented out, HMC.cpp:222-223): one trie descent per index
    double getLikelihood(const int8_t *hap) const
    {
        double likelihood = 1.0;
        for (int i = model->head_len(); i <= panel->L; ++i) {
            const int hp = model->trie.findLongestMatch(i, hap, panel->L, 0, model->pats, panel->nall);
            if (hp >= 0) likelihood *= model->pats[hp].tp;
            else { likelihood = 0; break; }
        }
        return likelihood;
    }

    // HaploBuilder.cpp:153-224 + HaploPair head ctor HaploPair.cpp:14-33
    void initHeadList(int gi)
    {
        const int head_len = model->head_len();
        std::vector<std::vector<int8_t> > last_list, new_list;
        for (size_t hi = 0; hi < model->heads.size(); ++hi) {
            const int head = model->heads[hi];
            const Pat &hd = model->pats[head];
            if (!model->genoMatch(gi, hd, hd.start, hd.len())) continue;
            last_list.clear();
            last_list.push_back(std::vector<int8_t>());
            for (int j = 0; j < head_len; ++j) {
                new_list.clear();
                if (panel->isMissing(gi, j) || (panel->hasMissing(gi, j) && panel->hasAllele(gi, j, hd.al[j]))) {
                    for (size_t s = 0; s < last_list.size(); ++s)
                        for (int k = 0; k < panel->nall[j]; ++k)
                            if (panel->af(j, k) > 0) { new_list.push_back(last_list[s]); new_list.back().push_back((int8_t) k); }
                }
                else if (panel->isHet(gi, j)) {
                    for (size_t s = 0; s < last_list.size(); ++s) {
                        new_list.push_back(last_list[s]);
                        new_list.back().push_back((int8_t) (hd.al[j] == panel->al(gi, 0, j) ? panel->al(gi, 1, j) : panel->al(gi, 0, j)));
                    }
                }
                else {
                    for (size_t s = 0; s < last_list.size(); ++s) { new_list.push_back(last_list[s]); new_list.back().push_back((int8_t) panel->al(gi, 0, j)); }
                }
                last_list.swap(new_list);
            }
            for (size_t s = 0; s < last_list.size(); ++s) {
                int hp = model->trie.findLongestMatch(head_len, last_list[s].data(), (int) last_list[s].size(), 0, model->pats, panel->nall);
                if (hp >= 0 && model->pats[hp].start == 0) {
                    if (hp >= head) {
                        State st;
                        st.ida = head; st.idb = hp;
                        st.beta = 1.0;
                        st.alpha = st.tp = model->pats[head].freq * model->pats[hp].freq;
                        bool homo = (head == hp);
                        if (!homo) st.alpha *= 2.0;
                        Link lk = { -1, 0, false, homo, st.tp };
                        st.links.push_back(lk);
                        at[head_len].push_back(st);
                    }
                }
                else { fprintf(stderr, "oracle: Can not find matching pattern!\n"); exit(1); }
            }
        }
    }

    // HaploBuilder.cpp:226-261 + HaploPair.cpp:35-89
    void extendAll(int i, int a1, int a2, std::unordered_map<uint64_t, int> &best_pair)
    {
        std::vector<State> &cur = at[i];
        std::vector<State> &nxt = at[i + 1];
        for (size_t s = 0; s < cur.size(); ++s) {
            for (int dir = 0; dir < (a1 != a2 ? 2 : 1); ++dir) {
                int x = dir ? a2 : a1, y = dir ? a1 : a2;
                if (cur[s].alpha <= 0) continue;
                int hpa = model->pats[cur[s].ida].successor(x);
                int hpb = model->pats[cur[s].idb].successor(y);
                if (hpa < 0 || hpb < 0) continue;
                bool reversed = false;
                if (hpa > hpb) { reversed = true; std::swap(hpa, hpb); }
                cnt.arrivals++;
                uint64_t key = ((uint64_t) (uint32_t) hpa << 32) | (uint32_t) hpb;
                std::unordered_map<uint64_t, int>::iterator it = best_pair.find(key);
                State *t;
                size_t k0;
                const std::vector<Link> &pl = cur[s].links;
                if (it == best_pair.end()) {
                    nxt.push_back(State());
                    best_pair[key] = (int) nxt.size() - 1;
                    t = &nxt.back();
                    t->ida = hpa; t->idb = hpb; t->beta = 1.0;
                    t->tp = model->pats[hpa].tp * model->pats[hpb].tp;
                    t->alpha = cur[s].alpha * t->tp;
                    k0 = 0;
                    cur[s].fwd[reversed ? 1 : 0].push_back((int) nxt.size() - 1);
                }
                else {
                    t = &nxt[it->second];
                    t->alpha += cur[s].alpha * t->tp;
                    k0 = t->links.size();
                    cur[s].fwd[reversed ? 1 : 0].push_back(it->second);
                }
                t->nin++;
                t->links.insert(t->links.end(), pl.begin(), pl.end());
                for (size_t k = k0; k < t->links.size(); ++k) {
                    t->links[k].link = (int) s;
                    t->links[k].index = (int) (k - k0);
                    t->links[k].reversed = reversed;
                    t->links[k].likelihood *= t->tp;
                }
                if (lastA(t->ida) != lastA(t->idb)) {
                    for (size_t k = k0; k < t->links.size(); ++k) {
                        if (t->links[k].homozygous) {
                            if (reversed) t->links[k].likelihood = 0;
                            t->links[k].homozygous = false;
                        }
                    }
                }
                if (k0 > 0 && (int) t->links.size() > sample_size) {
                    std::nth_element(t->links.begin(), t->links.begin() + sample_size - 1, t->links.end(), LinkGreater());
                    t->links.resize(sample_size);
                }
            }
        }
    }

    // HaploPair::getGenotype HaploPair.cpp:91-124
    Resolution getGenotype(int si, int index) const
    {
        const int L = panel->L;
        Resolution r;
        r.h[0].assign(L, -1); r.h[1].assign(L, -1);
        int a = 0, b = 1, i = L - 1, loc = L;
        const State *hp = &at[loc][si];
        r.prior = hp->links[index].homozygous ? hp->links[index].likelihood : hp->links[index].likelihood * 2.0;
        while (true) {
            const Pat &pa = model->pats[hp->ida], &pb = model->pats[hp->idb];
            if (index < (int) hp->links.size() && hp->links[index].link >= 0) {
                r.h[a][i] = pa.al[i - pa.start];
                r.h[b][i] = pb.al[i - pb.start];
                if (hp->links[index].reversed) std::swap(a, b);
                int nsi = hp->links[index].link;
                index = hp->links[index].index;
                --loc; --i;
                hp = &at[loc][nsi];
            }
            else {
                for (int k = 0; k < i + 1 - pa.start; ++k) r.h[a][pa.start + k] = pa.al[k];
                for (int k = 0; k < i + 1 - pb.start; ++k) r.h[b][pb.start + k] = pb.al[k];
                break;
            }
        }
        return r;
    }

    // HaploBuilder::resolve HaploBuilder.cpp:35-126.  Returns coverage.  In scaled mode priors/PG are
    // the scaled values; the true value is x * 2^-final_exp.
    double resolve(int gi, std::vector<Resolution> &res_list, int ssize = 1)
    {
        const int L = panel->L, head_len = model->head_len();
        sample_size = ssize > 1 ? ssize : 1;
        at.assign(L + 1, std::vector<State>());
        scale_exp.assign(L + 1, 0);
        final_exp = 0;
        initHeadList(gi);
        std::unordered_map<uint64_t, int> best_pair;
        int i;
        for (i = head_len; i < L; ++i) {
            best_pair.clear();
            const int g0 = panel->al(gi, 0, i), g1 = panel->al(gi, 1, i);
            if (g0 < 0 && g1 < 0) {
                for (int j = 0; j < panel->nall[i]; ++j) if (panel->af(i, j) > 0)
                    for (int k = j; k < panel->nall[i]; ++k) if (panel->af(i, k) > 0) extendAll(i, j, k, best_pair);
            }
            else if (g0 < 0) { for (int j = 0; j < panel->nall[i]; ++j) if (panel->af(i, j) > 0) extendAll(i, j, g1, best_pair); }
            else if (g1 < 0) { for (int j = 0; j < panel->nall[i]; ++j) if (panel->af(i, j) > 0) extendAll(i, j, g0, best_pair); }
            else extendAll(i, g0, g1, best_pair);
            scale_exp[i + 1] = scale_exp[i];
            if (at[i + 1].empty()) break;
            cnt.steps++;
            if (scaled) {
                double mx = 0;
                for (size_t s = 0; s < at[i + 1].size(); ++s) mx = std::max(mx, at[i + 1][s].alpha);
                if (mx > 0 && mx < std::ldexp(1.0, -512)) {
                    for (size_t s = 0; s < at[i + 1].size(); ++s) {
                        at[i + 1][s].alpha *= std::ldexp(1.0, 512);
                        for (size_t k = 0; k < at[i + 1][s].links.size(); ++k) at[i + 1][s].links[k].likelihood *= std::ldexp(1.0, 512);
                    }
                    scale_exp[i + 1] += 512;
                }
            }
        }
        for (int l = head_len; l <= L; ++l) { cnt.states += at[l].size(); cnt.max_states = std::max(cnt.max_states, (long long) at[l].size()); }
        final_exp = scale_exp[L];
        double coverage = 0;
        res_list.clear();
        if (!at[L].empty()) {
            double total = 0;
            std::vector<Link> res_link;
            for (size_t s = 0; s < at[L].size(); ++s) {
                total += at[L][s].alpha;
                size_t k = res_link.size(), n = at[L][s].links.size();
                res_link.insert(res_link.end(), at[L][s].links.begin(), at[L][s].links.end());
                for (size_t q = k; q < k + n; ++q) {
                    res_link[q].link = (int) s;
                    res_link[q].index = (int) (q - k);
                    if (!res_link[q].homozygous) res_link[q].likelihood *= 2.0;
                }
                if ((int) res_link.size() > sample_size) {
                    std::nth_element(res_link.begin(), res_link.begin() + sample_size - 1, res_link.end(), LinkGreater());
                    res_link.resize(sample_size);
                }
            }
            std::sort(res_link.begin(), res_link.end(), LinkGreater());
            for (size_t q = 0; q < res_link.size(); ++q) {
                Resolution r = getGenotype(res_link[q].link, res_link[q].index);
                r.posterior = r.prior / total;
                r.pg = total;
                coverage += r.posterior;
                res_list.push_back(r);
            }
        }
        return coverage;
    }

    // HaploBuilder.cpp:263-272 + HaploPair.cpp:126-136 (unscaled; the E-step is pinned on L <= ~1000)
    void calcBackward()
    {
        const int L = panel->L, head_len = model->head_len();
        for (int i = L - 1; i >= head_len; --i) {
            for (size_t s = 0; s < at[i].size(); ++s) {
                State &st = at[i][s];
                st.beta = 0;
                for (int r = 0; r < 2; ++r)
                    for (size_t k = 0; k < st.fwd[r].size(); ++k) { const State &nx = at[i + 1][st.fwd[r][k]]; st.beta += nx.beta * nx.tp; }
            }
        }
    }
};

// ForwardPatternTree (PatternTree.cpp:176-212) for the E-step
struct FwdTrie {
    int A = 0;
    std::vector<int> child, data, root;
    int newNode() { data.push_back(-1); child.insert(child.end(), A, -1); return (int) data.size() - 1; }
    void init(int L, int Amax) { A = Amax; child.clear(); data.clear(); root.resize(L + 1); for (int i = 0; i <= L; ++i) root[i] = newNode(); }
    void addPattern(const Pat &p, int id)
    {
        int node = root[p.start];
        for (int k = 0; k < p.len(); ++k) {
            int c = child[(size_t) node * A + p.al[k]];
            if (c < 0) { c = newNode(); child[(size_t) node * A + p.al[k]] = c; }
            node = c;
            if (k == p.len() - 1) data[node] = id;
        }
    }
};

// HaploBuilder::estimateFrequency (HaploBuilder.cpp:274-450).  Maps are keyed by state index within
// a locus instead of HaploPair* (heap-address order is not reproducible); sums therefore
// agree with the reference to rounding, not to the bit.
struct EStep {
    Builder &b;
    std::vector<Pat> &pats;
    FwdTrie tree;
    double cur_pg = 0;
    typedef std::map<int, double> WMap;

    EStep(Builder &bb, std::vector<Pat> &pp) : b(bb), pats(pp) {}

    double rec(int node, int locus, int a, double last_freq, const WMap last_match[3])
    {
        WMap match_list[3];
        const int head_len = b.model->head_len();
        const std::vector<Pat> &mp = b.model->pats;
        if (locus < head_len) {
            const std::vector<State> &st = b.at[head_len];
            for (WMap::const_iterator it = last_match[0].begin(); it != last_match[0].end(); ++it) {
                const State &hp = st[it->first];
                double weight = it->second;
                if (mp[hp.ida].al[locus] == a) {
                    if (mp[hp.idb].al[locus] == a) match_list[0][it->first] += weight;
                    else match_list[1][it->first] += weight * 0.5;
                }
                else if (mp[hp.idb].al[locus] == a) match_list[2][it->first] += weight * 0.5;
            }
            for (WMap::const_iterator it = last_match[1].begin(); it != last_match[1].end(); ++it)
                if (mp[st[it->first].ida].al[locus] == a) match_list[1][it->first] += it->second;
            for (WMap::const_iterator it = last_match[2].begin(); it != last_match[2].end(); ++it)
                if (mp[st[it->first].idb].al[locus] == a) match_list[2][it->first] += it->second;
        }
        else {
            const std::vector<State> &cur = b.at[locus];
            const std::vector<State> &nxt = b.at[locus + 1];
            for (WMap::const_iterator it = last_match[0].begin(); it != last_match[0].end(); ++it) {
                const State &last_hp = cur[it->first];
                double weight = it->second;
                for (int i = 0; i < 2; ++i) {
                    for (size_t k = 0; k < last_hp.fwd[i].size(); ++k) {
                        int t = last_hp.fwd[i][k];
                        const State &hp = nxt[t];
                        int aa = b.lastA(hp.ida), ab = b.lastA(hp.idb);
                        if (aa == a) {
                            if (ab == a) match_list[0][t] += weight * hp.tp;
                            else match_list[1][t] += weight * hp.tp * 0.5;
                        }
                        else if (ab == a) match_list[2][t] += weight * hp.tp * 0.5;
                    }
                }
            }
            for (WMap::const_iterator it = last_match[1].begin(); it != last_match[1].end(); ++it) {
                const State &last_hp = cur[it->first];
                double weight = it->second;
                for (size_t k = 0; k < last_hp.fwd[0].size(); ++k) { int t = last_hp.fwd[0][k]; if (b.lastA(nxt[t].ida) == a) match_list[1][t] += weight * nxt[t].tp; }
                for (size_t k = 0; k < last_hp.fwd[1].size(); ++k) { int t = last_hp.fwd[1][k]; if (b.lastA(nxt[t].idb) == a) match_list[2][t] += weight * nxt[t].tp; }
            }
            for (WMap::const_iterator it = last_match[2].begin(); it != last_match[2].end(); ++it) {
                const State &last_hp = cur[it->first];
                double weight = it->second;
                for (size_t k = 0; k < last_hp.fwd[0].size(); ++k) { int t = last_hp.fwd[0][k]; if (b.lastA(nxt[t].idb) == a) match_list[2][t] += weight * nxt[t].tp; }
                for (size_t k = 0; k < last_hp.fwd[1].size(); ++k) { int t = last_hp.fwd[1][k]; if (b.lastA(nxt[t].ida) == a) match_list[1][t] += weight * nxt[t].tp; }
            }
        }
        const int idx = locus < head_len ? head_len : locus + 1;
        double freq = 0;
        for (int i = 0; i < 3; ++i)
            for (WMap::const_iterator it = match_list[i].begin(); it != match_list[i].end(); ++it) freq += it->second * b.at[idx][it->first].beta;
        freq /= cur_pg;
        if (tree.data[node] >= 0) { Pat &hp = pats[tree.data[node]]; hp.freq += freq; hp.prefix += last_freq; }
        for (int i = 0; i < tree.A; ++i) {
            int c = tree.child[(size_t) node * tree.A + i];
            if (c >= 0) rec(c, locus + 1, i, freq, match_list);
        }
        return freq;
    }

    void run()
    {
        const Panel &pn = *b.panel;
        const int L = pn.L, N = pn.N, head_len = b.model->head_len();
        tree.init(L, pn.Amax);
        for (size_t i = 0; i < pats.size(); ++i) { tree.addPattern(pats[i], (int) i); pats[i].freq = 0; pats[i].prefix = 0; }
        std::vector<Resolution> res_list;
        const bool was_scaled = b.scaled;
        b.scaled = false;       // the E-step restatement is unscaled like the reference (pinned on L <= ~1000)
        for (int geno = 0; geno < N; ++geno) {
            b.resolve(geno, res_list, 1);
            b.calcBackward();
            cur_pg = pn.pg[geno];
            for (int start = 0; start < L; ++start) {
                WMap match_list[3];
                int end = std::max(start, head_len);
                for (size_t i = 0; i < b.at[end].size(); ++i) match_list[0][(int) i] = b.at[end][i].alpha;
                int node = tree.root[start];
                for (int i = 0; i < tree.A; ++i) {
                    int c = tree.child[(size_t) node * tree.A + i];
                    if (c >= 0) rec(c, start, i, 1.0, match_list);
                }
            }
        }
        b.scaled = was_scaled;
        for (size_t i = 0; i < pats.size(); ++i) {     // HaploBuilder.cpp:317-331
            Pat &hp = pats[i];
            double freq = std::min(hp.freq, (double) N);
            double prefix_freq = std::min(hp.prefix, (double) N);
            freq = std::min(freq, prefix_freq);
            hp.freq = freq / N;
            hp.prefix = prefix_freq / N;
            if (prefix_freq > 0) hp.setTp(freq / prefix_freq);
            else hp.setTp(freq / N);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// HaploModel (HaploModel.cpp) -- EM driver
struct HaploModel {
    Panel panel;            // owned working copy ("unphased")
    Samples samples;
    Model model;
    Builder builder;
    Options opt;
    std::vector<std::vector<Resolution> > klists;   // last resolveAll's K-lists (for tests)
    std::vector<double> ll_trace;
    std::vector<int> npat_trace;
    std::vector<int> exps;                          // per-individual final exponent (scaled mode)

    HaploModel() : builder(&panel, &model) {}

    void setPanel(const Panel &p) { panel = p; model.panel = &panel; model.samples = &samples; samples.L = panel.L; builder.panel = &panel; builder.model = &model; }

    void findPatterns()     // HaploModel.cpp:52-77
    {
        if (opt.min_freq_abs > 0) opt.min_freq = opt.min_freq_abs / (2.0 * panel.N);
        if (opt.model == "MV" || opt.model == "MA") {
            if (opt.num_patterns > 0) model.findPatternByNum(opt.num_patterns, opt.min_pattern_len, opt.max_pattern_len);
            else model.findPatternByFreq(opt.min_freq, opt.min_pattern_len, opt.max_pattern_len);
        }
        else if (opt.model == "MC") model.findPatternBlock(opt.mc_order + 1);
    }

    // HaploModel.cpp:79-115.  log P(G) is taken on the scaled total minus the carried exponent.
    double resolveAll(std::vector<Resolution> &resolutions)
    {
        double log_likelihood = 0;
        samples.clear();
        builder.scaled = opt.scaled;
        klists.assign(panel.N, std::vector<Resolution>());
        exps.assign(panel.N, 0);
        for (int i = 0; i < panel.N; ++i) {
            std::vector<Resolution> &res_list = klists[i];
            double coverage = builder.resolve(i, res_list, opt.sample_size);
            exps[i] = builder.final_exp;
            if (res_list.empty()) {
                Resolution r;      // resolution = input genotype with zero probabilities (HaploBuilder.cpp:117-124)
                r.h[0].assign(&panel.g[((size_t) i * 2) * panel.L], &panel.g[((size_t) i * 2 + 1) * panel.L]);
                r.h[1].assign(&panel.g[((size_t) i * 2 + 1) * panel.L], &panel.g[((size_t) i * 2 + 2) * panel.L]);
                resolutions[i] = r;
            }
            else {
                resolutions[i] = res_list.front();
                for (size_t j = 0; j < res_list.size(); ++j) {
                    double w = res_list[j].posterior / coverage;
                    for (int h = 0; h < 2; ++h) { samples.h.insert(samples.h.end(), res_list[j].h[h].begin(), res_list[j].h[h].end()); samples.w.push_back(w); }
                }
            }
            panel.pg[i] = resolutions[i].pg;
            log_likelihood += opt.scaled ? (std::log(resolutions[i].pg) - builder.final_exp * M_LN2) : std::log(resolutions[i].pg);
        }
        samples.total = 0;
        for (size_t i = 0; i < samples.w.size(); ++i) samples.total += samples.w[i];
        return log_likelihood;
    }

    // PatternManager::estimateFrequency PatternManager.cpp:347-362
    void estimateFrequencyFixed()
    {
        std::vector<Pat> patterns = model.pats;
        EStep es(builder, patterns);
        es.run();
        for (int i = 0; i < model.P(); ++i) { model.pats[i].freq = patterns[i].freq; model.pats[i].prefix = patterns[i].prefix; model.pats[i].setTp(patterns[i].tp); }
    }
    // PatternManager.cpp:412-438
    void extendPatterns(std::vector<Pat> &patterns, std::vector<Pat> &seeds)
    {
        std::vector<Pat> new_seeds;
        int level = 4;
        // the reference holds pointers, so a seed's frequency is whatever the last E-pass wrote into it;
        // seeds produced inside this loop have not been through an E-pass and carry the parent's frequency
        while (level-- > 0) {
            for (size_t i = 0; i < seeds.size(); ++i) {
                const Pat &hp = seeds[i];
                if (hp.end < panel.L && hp.len() < model.max_len[hp.start]) {
                    if (hp.freq >= model.min_freq) {
                        for (int j = 0; j < panel.nall[hp.end]; ++j) {
                            Pat n;      // fresh HaploPattern: prefix = tp = 1.0 (HaploPattern.h:105-107)
                            n.start = hp.start; n.end = hp.end + 1; n.al = hp.al; n.al.push_back((int8_t) j);
                            n.freq = hp.freq;
                            patterns.push_back(n);
                            new_seeds.push_back(n);
                        }
                    }
                }
            }
            seeds.clear();
            seeds.swap(new_seeds);
        }
    }
    // PatternManager.cpp:364-410
    void estimatePatterns()
    {
        if (model.min_freq < 0) { estimateFrequencyFixed(); return; }
        std::vector<Pat> patterns, candidates;
        std::vector<int> seed_idx;     // indices into `patterns` (the reference aliases the same objects)
        for (int i = 0; i < model.P(); ++i) {
            const Pat &hp = model.pats[i];
            patterns.push_back(hp);
            if (hp.end < panel.L && hp.len() < model.max_len[hp.start]) {
                for (int j = 0; j < panel.nall[hp.end]; ++j) {
                    int s = hp.successor(j);
                    if (s < 0 || model.pats[s].start != hp.start) {
                        Pat n;
                        n.start = hp.start; n.end = hp.end + 1; n.al = hp.al; n.al.push_back((int8_t) j);
                        n.freq = hp.freq;
                        patterns.push_back(n);
                        seed_idx.push_back((int) patterns.size() - 1);
                    }
                }
            }
        }
        while (!patterns.empty()) {
            EStep es(builder, patterns);
            es.run();
            std::vector<Pat> seeds;
            for (size_t k = 0; k < seed_idx.size(); ++k) seeds.push_back(patterns[seed_idx[k]]);
            candidates.insert(candidates.end(), patterns.begin(), patterns.end());
            patterns.clear();
            extendPatterns(patterns, seeds);
            // after extendPatterns the surviving seeds are the last level's new patterns, which sit at
            // the tail of `patterns` in the same order
            seed_idx.clear();
            for (size_t k = 0; k < seeds.size(); ++k) seed_idx.push_back((int) (patterns.size() - seeds.size() + k));
        }
        model.pats.clear();
        for (size_t i = 0; i < candidates.size(); ++i) {
            const Pat &hp = candidates[i];
            if (hp.freq >= model.min_freq || hp.len() <= model.min_len[hp.start]) { model.pats.push_back(hp); model.pats.back().succ.clear(); }
        }
        model.initialize();
    }

    // HaploModel::run HaploModel.cpp:117-155
    void run(const Panel &genos, std::vector<Resolution> &resolutions)
    {
        setPanel(genos);
        samples.clear();
        findPatterns();     // build()
        std::vector<Resolution> resolved(panel.N);
        resolutions.assign(panel.N, Resolution());
        ll_trace.clear(); npat_trace.clear();
        double old_ll = -DBL_MAX;
        for (int iter = 1; iter <= opt.max_iteration; ++iter) {
            npat_trace.push_back(model.P());
            double ll = resolveAll(resolved);
            ll_trace.push_back(ll);
            if (ll >= old_ll) resolutions = resolved;
            if (iter < opt.max_iteration && ll >= old_ll && (old_ll - ll) / old_ll > 0.0001) {
                if (opt.exact_estimate) estimatePatterns();
                else findPatterns();
                old_ll = ll;
            }
            else break;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// HaploComp (HaploComp.cpp:29-76) over allele codes (any per-locus coding shared by both sides, negative = missing):
// real = the target / true data, infer = the resolutions.  Alleles "match" when either is missing (Allele.h:20).
struct CompareResult {
    long long se_num = 0, se_den = 0, ig_num = 0, ig_den = 0, ih_num = 0, ih_den = 0;
    int bad_individual = -1, bad_locus = -1;        // "Inconsistent genotypes at locus %d!" (Genotype.cpp:258-261,272-275): the reference exits
    double switch_error() const { return (double) se_num / se_den; }                       // HaploComp.cpp:145-147 (0/0 is NaN there too)
    double incorrect_genotype_percentage() const { return (double) ig_num / ig_den; }
    double incorrect_haplotype_percentage() const { return (double) ih_num / ih_den; }
};

inline CompareResult compare_genotypes(const int8_t *real, const int8_t *infer, int N, int L)
{
    CompareResult r;
    auto am = [](int a, int b) { return a < 0 || b < 0 || a == b; };       // Allele::isMatch
    for (int i = 0; i < N; ++i) {
        const int8_t *r0 = real + ((size_t) i * 2) * L, *r1 = r0 + L, *f0 = infer + ((size_t) i * 2) * L, *f1 = f0 + L;
        auto match = [&](int l, bool rev) { return rev ? (am(r0[l], f1[l]) && am(r1[l], f0[l])) : (am(r0[l], f0[l]) && am(r1[l], f1[l])); };   // Genotype.cpp:98-118
        auto has_missing = [&](int l) { return r0[l] < 0 || r1[l] < 0; };  // Genotype.h:120-123
        int het = 0, miss = 0;                                              // Genotype::checkGenotype, Genotype.cpp:44-56
        for (int l = 0; l < L; ++l) { if (!am(r0[l], r1[l])) ++het; if (has_missing(l)) ++miss; }
        // getSwitchDistanceIgnoreMissing, Genotype.cpp:235-279
        int sd = 0, start = L;
        for (int l = 0; l < L; ++l) { if (has_missing(l) || (match(l, true) && match(l, false))) continue; start = l; break; }
        if (start < L) {
            bool reversed;
            if (match(start, true)) reversed = true;
            else if (match(start, false)) reversed = false;
            else { if (r.bad_individual < 0) { r.bad_individual = i; r.bad_locus = start; } continue; }
            bool bad = false;
            for (int l = start + 1; l < L; ++l) {
                if (has_missing(l) || match(l, reversed)) continue;
                if (match(l, !reversed)) { reversed = !reversed; ++sd; }
                else { if (r.bad_individual < 0) { r.bad_individual = i; r.bad_locus = l; } bad = true; break; }
            }
            if (bad) continue;
        }
        // getDiffNumIgnoreMissing, Genotype.cpp:160-175
        int d1 = 0, d2 = 0;
        for (int l = 0; l < L; ++l) if (!has_missing(l)) { if (!match(l, true)) ++d1; if (!match(l, false)) ++d2; }
        r.se_num += sd; r.se_den += het - 1;                                // HaploComp.cpp:58-61
        r.ig_num += d1 < d2 ? d1 : d2; r.ig_den += L - miss;                // :62-65
        if (sd > 0) ++r.ih_num;                                             // :66-68
        if (het > 1) ++r.ih_den;
    }
    return r;
}

}  // namespace orc
