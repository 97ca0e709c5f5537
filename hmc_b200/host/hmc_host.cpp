// Host side of hmc_b200 (see hmc_host.hpp).  Mining and automaton construction follow the reference's
// arithmetic and id order exactly (PatternManager.cpp) but run multi-threaded over start / end loci on
// flat arrays; everything per-individual goes through the C-ABI to the GPU.
#include "hmc_host.hpp"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>

namespace hmc_host {

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int hw_threads(int want)
{
    if (want > 0) return want;
    const char *env = getenv("HMC_B200_HOST_THREADS");
    if (env && atoi(env) > 0) return atoi(env);
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int) std::min(n, 64u) : 4;
}

template <typename F>
static void parallel_for(int n, int threads, F f)
{
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) { for (int i = 0; i < n; ++i) f(i, 0); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() { for (;;) { int i = next.fetch_add(1); if (i >= n) break; f(i, t); } });
    for (auto &th : pool) th.join();
}

// ---------------------------------------------------------------------------------------------
// GenoData
void GenoData::resize(int n, int l)
{
    N = n; L = l;
    raw.assign((size_t) N * 2 * L, -1);
    ids.assign(N, std::string());
    allele_type.assign(L, 'M');                     // GenoData.cpp:71
    position.resize(L);
    for (int i = 0; i < L; ++i) position[i] = i * 1000;     // Constant::average_marker_distance, Constant.cpp:5
    genotype_probability.assign(N, 0.0);
    packed.clear(); packed_bits = packed_words = 0;
}

int GenoData::getAlleleIndex(int locus, int value) const
{
    const std::vector<int32_t> &s = symbols[locus];
    for (size_t i = 0; i < s.size(); ++i) if (s[i] == value) return (int) i;
    return -1;
}

void GenoData::checkAlleleSymbol()
{
    symbols.assign(L, std::vector<int32_t>());
    n_alleles.assign(L, 0);
    for (int l = 0; l < L; ++l) {
        std::vector<int32_t> &s = symbols[l];
        for (int i = 0; i < N; ++i) for (int h = 0; h < 2; ++h) {
            const int v = allele(i, h, l);
            if (v >= 0 && std::find(s.begin(), s.end(), v) == s.end()) s.push_back(v);
        }
        std::sort(s.begin(), s.end());
        n_alleles[l] = (int) s.size();
    }
    Amax = 0;
    for (int l = 0; l < L; ++l) Amax = std::max(Amax, n_alleles[l]);
    if (Amax > HMC_MAX_ALLELES) throw HmcError("more than " + std::to_string(HMC_MAX_ALLELES) + " alleles at one locus");
    allele_freq.assign((size_t) L * std::max(Amax, 1), 0.0);
    idx.assign((size_t) N * 2 * L, -1);
    std::vector<double> total(L, 0.0);
    for (int i = 0; i < N; ++i) for (int h = 0; h < 2; ++h) for (int l = 0; l < L; ++l) {
        const int v = allele(i, h, l);
        if (v >= 0) {
            const int k = getAlleleIndex(l, v);
            idx[((size_t) i * 2 + h) * L + l] = (int8_t) k;
            allele_freq[(size_t) l * Amax + k] += 1.0;      // haplotype weight 1 (Haplotype.h:33)
            total[l] += 1.0;
        }
    }
    for (int l = 0; l < L; ++l) for (int j = 0; j < n_alleles[l]; ++j) allele_freq[(size_t) l * Amax + j] /= total[l];
    genotype_probability.assign(N, 0.0);
}

// ---------------------------------------------------------------------------------------------
// HaploFile: PHASE text
static const char *DELIM = " \t\r\n";

static bool read_line(FILE *fp, std::string &out)
{
    out.clear();
    char buf[65536];
    bool got = false;
    while (fgets(buf, sizeof buf, fp)) {
        got = true;
        out += buf;
        if (!out.empty() && out.back() == '\n') break;
    }
    return got;
}

// ---------------------------------------------------------------------------------------------
// HaploFile: binary panel
namespace {
const char HMCB_MAGIC[8] = {'H', 'M', 'C', 'B', '0', '0', '1', '\n'};
struct HmcbHeader { char magic[8]; int32_t N, L, Amax, bits, words, reserved[3]; };

int bits_for(int Amax) { return Amax + 1 <= 4 ? 2 : (Amax + 1 <= 16 ? 4 : 8); }

void pack_indices(const std::vector<int8_t> &idx, int N, int L, int bits, int words, std::vector<uint32_t> &out)
{
    const int per = 32 / (2 * bits);
    out.assign((size_t) N * words, 0u);
    for (int i = 0; i < N; ++i) {
        const int8_t *g0 = &idx[((size_t) i * 2) * L], *g1 = g0 + L;
        uint32_t *w = &out[(size_t) i * words];
        for (int l = 0; l < L; ++l) {
            const uint32_t c = (uint32_t) (g0[l] + 1) | ((uint32_t) (g1[l] + 1) << bits);
            w[l / per] |= c << ((l % per) * 2 * bits);
        }
    }
}

void unpack_indices(const std::vector<uint32_t> &packed, int N, int L, int bits, int words, std::vector<int8_t> &idx)
{
    const int per = 32 / (2 * bits);
    const uint32_t m = (1u << bits) - 1u;
    idx.assign((size_t) N * 2 * L, -1);
    for (int i = 0; i < N; ++i) {
        const uint32_t *w = &packed[(size_t) i * words];
        int8_t *g0 = &idx[((size_t) i * 2) * L], *g1 = g0 + L;
        for (int l = 0; l < L; ++l) {
            const uint32_t c = w[l / per] >> ((l % per) * 2 * bits);
            g0[l] = (int8_t) ((int) (c & m) - 1);
            g1[l] = (int8_t) ((int) ((c >> bits) & m) - 1);
        }
    }
}
}  // namespace

bool HaploFile::isBinary(const std::string &path)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    char m[8] = {0};
    const size_t n = fread(m, 1, 8, fp);
    fclose(fp);
    return n == 8 && memcmp(m, HMCB_MAGIC, 8) == 0;
}

void HaploFile::writeBinary(const std::string &path, const GenoData &g, const std::vector<int8_t> &idx)
{
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) throw HmcError("Can not open file " + path + "!");
    HmcbHeader h = {};
    memcpy(h.magic, HMCB_MAGIC, 8);
    h.N = g.N; h.L = g.L; h.Amax = std::max(g.Amax, 1); h.bits = bits_for(h.Amax);
    const int per = 32 / (2 * h.bits);
    h.words = (g.L + per - 1) / per;
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1;
    ok = ok && fwrite(g.allele_type.data(), 1, g.L, fp) == (size_t) g.L;
    ok = ok && fwrite(g.position.data(), sizeof(int), g.L, fp) == (size_t) g.L;
    ok = ok && fwrite(g.n_alleles.data(), sizeof(int32_t), g.L, fp) == (size_t) g.L;
    std::vector<int32_t> sym((size_t) g.L * h.Amax, -1);
    for (int l = 0; l < g.L; ++l) for (int j = 0; j < g.n_alleles[l]; ++j) sym[(size_t) l * h.Amax + j] = g.symbols[l][j];
    ok = ok && fwrite(sym.data(), sizeof(int32_t), sym.size(), fp) == sym.size();
    ok = ok && fwrite(g.allele_freq.data(), sizeof(double), (size_t) g.L * h.Amax, fp) == (size_t) g.L * h.Amax;
    for (int i = 0; i < g.N && ok; ++i) {
        const uint16_t n = (uint16_t) std::min<size_t>(g.ids[i].size(), 65535);
        ok = fwrite(&n, 2, 1, fp) == 1 && (n == 0 || fwrite(g.ids[i].data(), 1, n, fp) == n);
    }
    std::vector<uint32_t> packed;
    pack_indices(idx, g.N, g.L, h.bits, h.words, packed);
    ok = ok && fwrite(packed.data(), sizeof(uint32_t), packed.size(), fp) == packed.size();
    fclose(fp);
    if (!ok) throw HmcError("Can not write file " + path + "!");
}

void HaploFile::readBinary(GenoData &g) const
{
    FILE *fp = fopen(m_filename.c_str(), "rb");
    if (!fp) throw HmcError("Can not open file " + m_filename + "!");
    HmcbHeader h;
    auto bad = [&](const char *why) { fclose(fp); throw HmcError(std::string("Invalid binary panel (") + why + "): " + m_filename); };
    if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, HMCB_MAGIC, 8) != 0) bad("header");
    if (h.N <= 0 || h.L <= 0 || h.Amax < 1 || h.Amax > HMC_MAX_ALLELES || h.bits != bits_for(h.Amax)) bad("dimensions");
    const int per = 32 / (2 * h.bits);
    if (h.words != (h.L + per - 1) / per) bad("row length");
    g.resize(h.N, h.L);
    g.Amax = h.Amax;
    if (fread(&g.allele_type[0], 1, h.L, fp) != (size_t) h.L) bad("types");
    if (fread(g.position.data(), sizeof(int), h.L, fp) != (size_t) h.L) bad("positions");
    g.n_alleles.assign(h.L, 0);
    if (fread(g.n_alleles.data(), sizeof(int32_t), h.L, fp) != (size_t) h.L) bad("allele counts");
    std::vector<int32_t> sym((size_t) h.L * h.Amax);
    if (fread(sym.data(), sizeof(int32_t), sym.size(), fp) != sym.size()) bad("symbols");
    g.symbols.assign(h.L, std::vector<int32_t>());
    for (int l = 0; l < h.L; ++l) {
        if (g.n_alleles[l] < 0 || g.n_alleles[l] > h.Amax) bad("allele counts");
        g.symbols[l].assign(sym.begin() + (size_t) l * h.Amax, sym.begin() + (size_t) l * h.Amax + g.n_alleles[l]);
    }
    g.allele_freq.assign((size_t) h.L * h.Amax, 0.0);
    if (fread(g.allele_freq.data(), sizeof(double), g.allele_freq.size(), fp) != g.allele_freq.size()) bad("frequencies");
    for (int i = 0; i < h.N; ++i) {
        uint16_t n = 0;
        if (fread(&n, 2, 1, fp) != 1) bad("ids");
        g.ids[i].resize(n);
        if (n && fread(&g.ids[i][0], 1, n, fp) != n) bad("ids");
    }
    g.packed_bits = h.bits; g.packed_words = h.words;
    g.packed.resize((size_t) h.N * h.words);
    if (fread(g.packed.data(), sizeof(uint32_t), g.packed.size(), fp) != g.packed.size()) bad("genotypes");
    fclose(fp);
    unpack_indices(g.packed, h.N, h.L, h.bits, h.words, g.idx);
    for (int i = 0; i < h.N; ++i) for (int hp = 0; hp < 2; ++hp) for (int l = 0; l < h.L; ++l) {     // raw values as a text reader would have seen them
        const int k = g.idx[((size_t) i * 2 + hp) * h.L + l];
        if (k >= g.n_alleles[l]) bad("allele index");
        g.allele(i, hp, l) = k >= 0 ? g.symbols[l][k] : -1;
    }
    g.genotype_probability.assign(h.N, 0.0);
}

void HaploFile::readGenoData(GenoData &genos) const
{
    if (isBinary(m_filename)) { readBinary(genos); return; }
    FILE *fp = fopen(m_filename.c_str(), "r");
    if (!fp) throw HmcError("Can not open file " + m_filename + "!");
    int n = 0, l = 0;
    if (fscanf(fp, "%d\n", &n) != 1 || fscanf(fp, "%d\n", &l) != 1 || n <= 0 || l <= 0) { fclose(fp); throw HmcError("Invalid file type!"); }
    genos.resize(n, l);
    std::string line;
    read_line(fp, line);
    const char *s = line.c_str() + strspn(line.c_str(), DELIM);
    if (s[0] == 'P') {                                  // optional marker positions
        s += strcspn(s, DELIM); s += strspn(s, DELIM);
        for (int i = 0; i < l; ++i) { genos.position[i] = atoi(s); s += strcspn(s, DELIM); s += strspn(s, DELIM); }
        read_line(fp, line);
        s = line.c_str() + strspn(line.c_str(), DELIM);
    }
    for (int i = 0; i < l; ++i) {                       // one type character per locus, blanks allowed between them
        if (!*s) { fclose(fp); throw HmcError("Invalid locus type line!"); }
        genos.allele_type[i] = s[0];
        ++s; s += strspn(s, DELIM);
    }
    for (int i = 0; i < n; ++i) {
        if (!read_line(fp, line)) { fclose(fp); throw HmcError("Incorrect haplotype data for individual " + std::to_string(i) + "!"); }
        char idbuf[4096];
        if (sscanf(line.c_str(), "%4095s", idbuf) > 0) { genos.ids[i] = idbuf; std::replace(genos.ids[i].begin(), genos.ids[i].end(), ' ', '_'); }
        for (int h = 0; h < 2; ++h) {
            if (!read_line(fp, line)) { fclose(fp); throw HmcError("Incorrect haplotype data for individual " + std::to_string(i) + "!"); }
            const char *b = line.c_str();
            for (int k = 0; k < l; ++k) {               // AlleleSequence::readAllele, Allele.cpp:55-80
                b += strspn(b, DELIM);
                if (!*b) { fclose(fp); throw HmcError("Incorrect haplotype data for individual " + std::to_string(i) + "!"); }
                int v;
                if (genos.allele_type[k] == 'S') { v = (b[0] == '-' || b[0] == '?') ? -1 : (int) (unsigned char) b[0]; ++b; }
                else {
                    if (b[0] == '-' || b[0] == '?') v = -1;
                    else { const int a = atoi(b); v = a > 0 ? a : -1; }
                    b += strcspn(b, DELIM);
                }
                genos.allele(i, h, k) = v;
            }
        }
    }
    fclose(fp);
    genos.checkAlleleSymbol();
}

void HaploFile::writeGenoData(const GenoData &genos, const std::vector<int8_t> &res, const char *suffix) const
{
    const std::string out = m_filename + suffix;
    FILE *fp = fopen(out.c_str(), "w");
    if (!fp) throw HmcError("Can not open file " + out + "!");
    fprintf(fp, "%d\n%d\nP", genos.N, genos.L);
    for (int i = 0; i < genos.L; ++i) fprintf(fp, " %d", genos.position[i]);
    fprintf(fp, "\n%s\n", genos.allele_type.c_str());
    std::string buf;
    for (int i = 0; i < genos.N; ++i) {
        const std::string &id = genos.ids[i];
        if (!id.empty() && id[0] >= '0' && id[0] <= '9') fprintf(fp, "#%s\n", id.c_str());
        else fprintf(fp, "%s\n", id.c_str());
        for (int h = 0; h < 2; ++h) {
            buf.clear();
            for (int l = 0; l < genos.L; ++l) {         // AlleleSequence::writeAllele, Allele.cpp:82-101
                const int k = res[((size_t) i * 2 + h) * genos.L + l];
                const int v = k >= 0 ? genos.symbols[l][k] : -1;
                if (genos.allele_type[l] == 'S') { buf += v < 0 ? '?' : (char) v; buf += ' '; }
                else { buf += std::to_string(v); buf += ' '; }
            }
            fprintf(fp, "%s\n", buf.c_str());
        }
    }
    fclose(fp);
}

void HaploFile::writeGenoText(const GenoData &genos, const char *text, const int64_t *line_off, const char *suffix) const
{
    const std::string out = m_filename + suffix;
    FILE *fp = fopen(out.c_str(), "w");
    if (!fp) throw HmcError("Can not open file " + out + "!");
    fprintf(fp, "%d\n%d\nP", genos.N, genos.L);
    for (int i = 0; i < genos.L; ++i) fprintf(fp, " %d", genos.position[i]);
    fprintf(fp, "\n%s\n", genos.allele_type.c_str());
    for (int i = 0; i < genos.N; ++i) {
        const std::string &id = genos.ids[i];
        if (!id.empty() && id[0] >= '0' && id[0] <= '9') fprintf(fp, "#%s\n", id.c_str());
        else fprintf(fp, "%s\n", id.c_str());
        fwrite(text + line_off[2 * i], 1, (size_t) (line_off[2 * i + 2] - line_off[2 * i]), fp);      // both lines, newlines included
    }
    fclose(fp);
}

// ---------------------------------------------------------------------------------------------
int PatternSet::add(int start, const int8_t *a, int len, double f, double pf, double tp)
{
    Pattern p;
    p.start = start; p.end = start + len; p.al = (int64_t) alleles.size();
    p.frequency = f; p.prefix_freq = pf; p.transition_prob = tp;
    alleles.insert(alleles.end(), a, a + len);
    pats.push_back(p);
    return (int) pats.size() - 1;
}

// ---------------------------------------------------------------------------------------------
// Mining.  One depth-first walk per start locus (PatternManager::searchPattern, PatternManager.cpp:100-144):
// a candidate is extended by every allele of its end locus, released as a pattern BEFORE its children
// are visited, and the children are visited in descending allele order (the reference pops a stack);
// starts are visited in descending order.  Matching states (individual, factor) live on a per-thread
// arena stack; the factor arithmetic is checkFrequencyWithExtension / getMatchingFrequency
// (PatternManager.cpp:195-291) operation for operation.
namespace {

struct Miner {
    const GenoData &g;
    const HaploData &samples;
    double min_freq;
    const std::vector<int> &min_len, &max_len;
    bool use_samples;
    // locus-major copies for the sweeps
    std::vector<int8_t> g0, g1;     // [L][N]
    std::vector<int8_t> sh;         // [L][n_samples]

    Miner(const GenoData &gd, const HaploData &hd, double mf, const std::vector<int> &mn, const std::vector<int> &mx)
        : g(gd), samples(hd), min_freq(mf), min_len(mn), max_len(mx), use_samples(!hd.empty())
    {
        const int N = g.N, L = g.L;
        if (!use_samples) {
            g0.resize((size_t) L * N); g1.resize((size_t) L * N);
            for (int i = 0; i < N; ++i) for (int l = 0; l < L; ++l) {
                g0[(size_t) l * N + i] = g.idx[((size_t) i * 2) * L + l];
                g1[(size_t) l * N + i] = g.idx[((size_t) i * 2 + 1) * L + l];
            }
        }
        else {
            const size_t n = samples.w.size();
            sh.resize((size_t) L * n);
            for (size_t i = 0; i < n; ++i) for (int l = 0; l < L; ++l) sh[(size_t) l * n + i] = samples.h[i * L + l];
        }
    }

    // stack of matching states; grows without initialising (std::vector::resize would zero-fill every stripe)
    template <typename T>
    struct RawVec {
        T *p = nullptr;
        size_t n = 0, cap = 0;
        RawVec() {}
        RawVec(const RawVec &) = delete;
        RawVec &operator=(const RawVec &) = delete;
        ~RawVec() { free(p); }
        size_t size() const { return n; }
        void resize(size_t m)
        {
            if (m > cap) { cap = std::max(m, cap * 2 + 1024); p = (T *) realloc(p, cap * sizeof(T)); if (!p) throw HmcError("out of host memory while mining"); }
            n = m;
        }
        T &operator[](size_t i) { return p[i]; }
        const T &operator[](size_t i) const { return p[i]; }
        T *begin() { return p; }
        void assign(const std::vector<T> &v) { resize(v.size()); if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T)); }
    };
    struct Arena { RawVec<int32_t> id; RawVec<double> fr; };

    struct Frame { double freq; size_t off, len; bool scanned; };

    // children of candidate (start, path[0..len), freq, ms) at locus e = start + len; returns per-allele frames
    void extend(Arena &ar, int e, const Frame &parent, int plen, Frame *out) const
    {
        const int A = g.n_alleles[e];
        for (int j = 0; j < A; ++j) out[j] = Frame{0.0, 0, 0, false};
        const bool full = (plen == 0) || parent.len == 0;       // checkFrequency: rescan everything (PatternManager.cpp:197-199)
        double total[HMC_MAX_ALLELES];
        for (int j = 0; j < A; ++j) total[j] = 0.0;
        if (!use_samples) {
            const int N = g.N;
            const int8_t *a0 = &g0[(size_t) e * N], *a1 = &g1[(size_t) e * N];
            const double *af = &g.allele_freq[(size_t) e * g.Amax];
            if (full && plen > 0) {
                // no individual matched the parent, so none matches a child: frequency 0, empty state
                for (int j = 0; j < A; ++j) out[j] = Frame{0.0 / N, ar.id.size(), 0, true};
                return;
            }
            const size_t n_in = full ? (size_t) N : parent.len;
            // one sweep: child j's list gets its own stripe of n_in entries on the arena stack (an individual matches a child at most once)
            const size_t base = ar.id.size();
            for (int j = 0; j < A; ++j) { out[j].off = base + (size_t) j * n_in; out[j].len = 0; }
            ar.id.resize(base + (size_t) A * n_in); ar.fr.resize(base + (size_t) A * n_in);
            for (size_t q = 0; q < n_in; ++q) {
                const int i = full ? (int) q : ar.id[parent.off + q];
                const double pf = full ? 1.0 : ar.fr[parent.off + q];
                const int x = a0[i], y = a1[i];
                auto emit = [&](int j) {
                    double f = 0;                       // getMatchingFrequency, one locus
                    if (x < 0) f += af[j]; else if (x == j) f += 1.0;
                    if (y < 0) f += af[j]; else if (y == j) f += 1.0;
                    double tf = 1.0;
                    tf *= 0.5 * f;
                    const double nf = full ? tf : pf * tf;
                    total[j] += nf;
                    const size_t pos = out[j].off + out[j].len++;
                    ar.id[pos] = i; ar.fr[pos] = nf;
                };
                if (x < 0 || y < 0) { for (int j = 0; j < A; ++j) emit(j); }
                else { emit(x); if (y != x) emit(y); }
            }
            for (int j = 0; j < A; ++j) { out[j].freq = total[j] / N; out[j].scanned = true; }
        }
        else {
            const size_t n = samples.w.size();
            const int8_t *hh = &sh[(size_t) e * n];
            if (full && plen > 0) { for (int j = 0; j < A; ++j) out[j] = Frame{0.0 / samples.total_weight, ar.id.size(), 0, true}; return; }
            const size_t n_in = full ? n : parent.len;
            const size_t base = ar.id.size();
            for (int j = 0; j < A; ++j) { out[j].off = base + (size_t) j * n_in; out[j].len = 0; }
            ar.id.resize(base + (size_t) A * n_in); ar.fr.resize(base + (size_t) A * n_in);
            for (size_t q = 0; q < n_in; ++q) {
                const int i = full ? (int) q : ar.id[parent.off + q];
                const int x = hh[i];
                auto emit = [&](int j) { total[j] += samples.w[i]; const size_t pos = out[j].off + out[j].len++; ar.id[pos] = i; ar.fr[pos] = 0.0; };
                if (x < 0) { for (int j = 0; j < A; ++j) emit(j); } else emit(x);
            }
            for (int j = 0; j < A; ++j) { out[j].freq = total[j] / samples.total_weight; out[j].scanned = true; }
        }
    }

    // visit = "pop this candidate" in the reference
    void visit(Arena &ar, PatternSet &out, int start, std::vector<int8_t> &path, const Frame &me, double prefix, double tp) const
    {
        const int len = (int) path.size(), e = start + len, L = g.L;
        Frame kids[HMC_MAX_ALLELES];
        const size_t mark = ar.id.size();
        bool extended = false;
        if (me.freq >= min_freq || len < min_len[start]) {
            if (e < L && len < max_len[start]) { extend(ar, e, me, len, kids); extended = true; }
        }
        if (me.freq >= min_freq || len <= min_len[start]) {
            if (len > 0 && len >= min_len[start]) out.add(start, path.data(), len, me.freq, prefix, tp);
        }
        if (extended) {
            const double *af = &g.allele_freq[(size_t) e * g.Amax];
            for (int j = g.n_alleles[e] - 1; j >= 0; --j) {
                if (!(af[j] > 0)) continue;
                Pattern t;
                const double pf = me.freq;
                if (pf > 0) t.setTransitionProb(kids[j].freq / pf); else t.setTransitionProb(kids[j].freq);
                path.push_back((int8_t) j);
                visit(ar, out, start, path, kids[j], pf, t.transition_prob);
                path.pop_back();
            }
        }
        ar.id.resize(mark); ar.fr.resize(mark);
    }
};

}  // namespace

void PatternManager::setLens(int mn, int mx)
{
    // vector::resize(n, v) fills only NEW elements (PatternManager.cpp:33-34): the lengths of the first call stick
    const int L = m_builder.genotype_len();
    m_min_len.resize(L, mn);
    m_max_len.resize(L, mx);
}

void PatternManager::searchPattern(bool by_num, int max_num)
{
    const GenoData &g = *m_builder.genos();
    const int L = g.L;
    last_mine_on_device = false;
    m_dev_list_fresh = false;
    if (!by_num && m_device_mining != 0 && searchPatternDevice()) { last_mine_on_device = true; m_dev_list_fresh = true; return; }
    m_patterns.clear();
    m_builder.materializeSamples();
    if (!by_num) {
        Miner miner(g, *m_builder.samples(), m_min_freq, m_min_len, m_max_len);
        std::vector<PatternSet> per_start(L);
        const int threads = hw_threads(m_threads);
        std::vector<Miner::Arena> arenas_store(threads);
        Miner::Arena *arenas = arenas_store.data();
        parallel_for(L, threads, [&](int k, int t) {
            const int start = L - 1 - k;                // heaviest (longest) walks are not special; any order works
            std::vector<int8_t> path;
            Miner::Frame root{1.0, 0, 0, false};        // empty pattern: frequency 1 (HaploPattern.h:105)
            miner.visit(arenas[t], per_start[start], start, path, root, 1.0, 1.0);
        });
        size_t np = 0, na = 0;
        for (int s = 0; s < L; ++s) { np += per_start[s].pats.size(); na += per_start[s].alleles.size(); }
        m_patterns.pats.reserve(np); m_patterns.alleles.reserve(na);
        for (int s = L - 1; s >= 0; --s) {              // the reference pops start L-1 first
            const int64_t shift = (int64_t) m_patterns.alleles.size();
            for (Pattern p : per_start[s].pats) { p.al += shift; m_patterns.pats.push_back(p); }
            m_patterns.alleles.insert(m_patterns.alleles.end(), per_start[s].alleles.begin(), per_start[s].alleles.end());
            PatternSet().pats.swap(per_start[s].pats);
        }
        return;
    }
    // findPatternByNum (PatternManager.cpp:44-73): repeated searches with a shrinking threshold over the candidates
    // that were neither released nor extended.  Sequential: the candidate stack is shared between rounds.
    struct Cand { int start; std::vector<int8_t> path; Miner::Frame fr; double prefix, tp; std::vector<int32_t> id; std::vector<double> f; };
    std::vector<Cand> cands, keep;
    for (int i = 0; i < L; ++i) cands.push_back(Cand{i, {}, Miner::Frame{1.0, 0, 0, false}, 1.0, 1.0, {}, {}});
    m_min_freq = 1.0;
    int last_size = 0;
    auto search = [&]() {
        Miner miner(g, *m_builder.samples(), m_min_freq, m_min_len, m_max_len);
        Miner::Arena ar;
        keep.clear();
        while (!cands.empty()) {
            Cand pc = std::move(cands.back());
            cands.pop_back();
            const int len = (int) pc.path.size(), e = pc.start + len;
            if (pc.fr.freq >= m_min_freq || len < m_min_len[pc.start]) {
                if (e < L && len < m_max_len[pc.start]) {
                    ar.id.assign(pc.id); ar.fr.assign(pc.f);
                    Miner::Frame me{pc.fr.freq, 0, pc.id.size(), false};
                    Miner::Frame kids[HMC_MAX_ALLELES];
                    miner.extend(ar, e, me, len, kids);
                    for (int j = 0; j < g.n_alleles[e]; ++j) {
                        if (!(g.allele_freq[(size_t) e * g.Amax + j] > 0)) continue;
                        Cand c;
                        c.start = pc.start; c.path = pc.path; c.path.push_back((int8_t) j);
                        c.fr = Miner::Frame{kids[j].freq, 0, 0, false};
                        c.prefix = pc.fr.freq;
                        Pattern t;
                        if (c.prefix > 0) t.setTransitionProb(c.fr.freq / c.prefix); else t.setTransitionProb(c.fr.freq);
                        c.tp = t.transition_prob;
                        c.id.assign(ar.id.begin() + kids[j].off, ar.id.begin() + kids[j].off + kids[j].len);
                        c.f.assign(ar.fr.begin() + kids[j].off, ar.fr.begin() + kids[j].off + kids[j].len);
                        cands.push_back(std::move(c));
                    }
                }
            }
            if (pc.fr.freq >= m_min_freq || len <= m_min_len[pc.start]) {
                if (len > 0 && len >= m_min_len[pc.start]) m_patterns.add(pc.start, pc.path.data(), len, pc.fr.freq, pc.prefix, pc.tp);
            }
            else keep.push_back(std::move(pc));
        }
        cands.swap(keep);
    };
    search();
    max_num = std::max(max_num, (int) m_patterns.size());
    while ((int) m_patterns.size() < max_num && m_min_freq > 1e-38) {
        last_size = (int) m_patterns.size();
        m_min_freq *= 0.9;
        search();
    }
    if ((int) m_patterns.size() > max_num) {
        std::sort(m_patterns.pats.begin() + last_size, m_patterns.pats.end(), [](const Pattern &a, const Pattern &b) { return a.frequency > b.frequency; });
        m_patterns.pats.resize(max_num);
    }
}

// The same search on the device (K5, csrc/mine.cuh): identical ids and bit-identical frequencies.
bool PatternManager::deviceMiningWanted() const
{
    const char *env = getenv("HMC_B200_HOST_MINING");
    return m_device_mining != 0 && !(env && atoi(env) != 0);
}

bool PatternManager::searchPatternDevice()
{
    if (!deviceMiningWanted()) return false;
    hmc_ctx *c = nullptr;
    try { c = m_builder.ctx(); }
    catch (const HmcError &) { if (m_device_mining > 0) throw; return false; }
    m_builder.uploadPanel();
    const HaploData &sm = *m_builder.samples();
    const int n = (int) sm.w.size();
    int64_t P = 0, na = 0;
    const double t0 = now_s();
    int rc = sm.on_device()
        ? hmc_mine_patterns_resolved(c, n, sm.rows.data(), sm.w.data(), sm.total_weight, m_min_freq, m_min_len.data(), m_max_len.data(), &P, &na)
        : hmc_mine_patterns(c, n, n ? sm.h.data() : nullptr, n ? sm.w.data() : nullptr, sm.total_weight, m_min_freq, m_min_len.data(), m_max_len.data(), &P, &na);
    if (rc < 0) throw HmcError(std::string("hmc_mine_patterns: ") + hmc_last_error(c));
    const double t1 = now_s();
    const hmc_pattern_rec *recs = nullptr;
    const int8_t *als = nullptr;
    rc = hmc_map_patterns(c, &recs, &als);
    if (rc < 0) throw HmcError(std::string("hmc_map_patterns: ") + hmc_last_error(c));
    const double t2 = now_s();
    static_assert(sizeof(Pattern) == sizeof(hmc_pattern_rec) && offsetof(Pattern, al) == offsetof(hmc_pattern_rec, al) &&
                  offsetof(Pattern, transition_prob) == offsetof(hmc_pattern_rec, transition_prob), "Pattern must mirror hmc_pattern_rec");
    m_patterns.pats.resize((size_t) P);
    m_patterns.alleles.resize((size_t) na);
    const int threads = hw_threads(m_threads), chunks = std::max(1, std::min(threads * 4, (int) (P / 65536) + 1));
    parallel_for(chunks, threads, [&](int k, int) {
        const size_t p0 = (size_t) P * k / chunks, p1 = (size_t) P * (k + 1) / chunks, a0 = (size_t) na * k / chunks, a1 = (size_t) na * (k + 1) / chunks;
        if (p1 > p0) memcpy((void *) (m_patterns.pats.data() + p0), recs + p0, (p1 - p0) * sizeof(Pattern));
        if (a1 > a0) memcpy(m_patterns.alleles.data() + a0, als + a0, a1 - a0);
    });
    if (getenv("HMC_B200_HOST_TIMING")) {
        double ms = 0; int64_t nodes = 0, levels = 0, mx = 0;
        hmc_get_mine_stats(c, &ms, &nodes, &levels, &mx, nullptr, nullptr);
        fprintf(stderr, "[host] device mining: call %.3f s (kernels %.1f ms, %lld trie nodes, %lld levels, largest level list %lld), fetch %.3f s, unpack %.3f s\n",
                t1 - t0, ms, (long long) nodes, (long long) levels, (long long) mx, t2 - t1, now_s() - t2);
    }
    return true;
}

void PatternManager::findPatternByFreq(double min_freq, int min_len, int max_len)
{
    const int L = m_builder.genotype_len();
    max_len = max_len <= 0 ? L : max_len;
    min_len = std::max(min_len, 1);
    max_len = std::max(max_len, min_len);
    setLens(min_len, max_len);
    m_min_freq = min_freq;
    const double t0 = now_s();
    searchPattern(false, 0);
    const double t1 = now_s();
    initialize();
    if (getenv("HMC_B200_HOST_TIMING")) fprintf(stderr, "[host] searchPattern %.3f s (%s), initialize %.3f s (%s), patterns %d\n", t1 - t0, last_mine_on_device ? "device" : "host threads", now_s() - t1,
                                                last_automaton_on_device ? "device" : "host threads", size());
}

void PatternManager::findPatternByNum(int max_num, int min_len, int max_len)
{
    const int L = m_builder.genotype_len();
    max_len = max_len <= 0 ? L : max_len;
    min_len = std::max(min_len, 1);
    max_len = std::max(max_len, min_len);
    setLens(min_len, max_len);
    searchPattern(true, max_num);
    initialize();
}

void PatternManager::findPatternBlock(int len)
{
    len = std::max(1, len);
    setLens(len, len);
    m_min_freq = -1.0;
    searchPattern(false, 0);
    initialize();
}

void PatternManager::setPatterns(const PatternSet &ps, int min_len, int max_len, double min_freq)
{
    m_patterns = ps;
    m_min_len.assign(m_builder.genotype_len(), min_len);
    m_max_len.assign(m_builder.genotype_len(), max_len);
    m_min_freq = min_freq;
    initialize();
}

// PatternManager::initialize (PatternManager.cpp:293-318): ids are positions; one backward trie per end locus
// (BackwardPatternTree, PatternTree.cpp:14-48); successor j of a pattern = the longest pattern ending one locus
// further that is a suffix of pattern+allele_j and starts no earlier (findLongestMatchPattern, :62-72,98-134).
void PatternManager::initialize()
{
    const GenoData &g = *m_builder.genos();
    const int L = g.L, A = std::max(g.Amax, 1), P = size();
    const int threads = hw_threads(m_threads);
    const bool fresh = m_dev_list_fresh;
    m_dev_list_fresh = false;
    m_host_automaton_stale = false;
    last_automaton_on_device = false;
    if (fresh && P > 0 && !getenv("HMC_B200_HOST_AUTOMATON")) {
        // the list came straight from the device miner and is still there: build and install the automaton in place (K6)
        hmc_ctx *c = m_builder.ctx();
        int32_t nh = 0;
        if (hmc_build_automaton(c, head_len(), &nh) < 0) throw HmcError(std::string("hmc_build_automaton: ") + hmc_last_error(c));
        m_dev_heads = nh;
        m_succ.clear(); m_head_list.clear();
        m_host_automaton_stale = true;
        last_automaton_on_device = true;
        m_builder.markModelClean();
        return;
    }
    m_head_list.clear();
    std::vector<std::vector<int32_t> > by_end(L + 1);
    for (int i = 0; i < P; ++i) {
        const Pattern &p = m_patterns.pats[i];
        by_end[p.end].push_back(i);
        if (p.start == 0 && p.length() == head_len()) m_head_list.push_back(i);
    }
    struct Trie { std::vector<int32_t> child, data; };
    std::vector<Trie> tries(L + 1);
    parallel_for(L + 1, threads, [&](int e, int) {
        Trie &t = tries[e];
        t.child.assign(A, -1); t.data.assign(1, -1);
        for (int id : by_end[e]) {
            const Pattern &p = m_patterns.pats[id];
            const int8_t *al = m_patterns.al(p);
            int node = 0;
            for (int k = p.length() - 1; k >= 0; --k) {
                int c = t.child[(size_t) node * A + al[k]];
                if (c < 0) { c = (int) t.data.size(); t.data.push_back(-1); t.child.insert(t.child.end(), A, -1); t.child[(size_t) node * A + al[k]] = c; }
                node = c;
            }
            t.data[node] = id;
        }
    });
    m_succ.assign((size_t) P * A, -1);
    const int chunk = 4096;
    parallel_for((P + chunk - 1) / chunk, threads, [&](int c, int) {
        for (int i = c * chunk; i < std::min(P, (c + 1) * chunk); ++i) {
            const Pattern &p = m_patterns.pats[i];
            if (p.end >= L) continue;
            const Trie &t = tries[p.end + 1];
            const int8_t *al = m_patterns.al(p);
            for (int j = 0; j < g.n_alleles[p.end]; ++j) {
                int node = t.child[j], result = -1;         // root has no data
                int k = p.length() - 1;
                while (node >= 0) {
                    if (t.data[node] >= 0) result = t.data[node];
                    if (k < 0) break;
                    node = t.child[(size_t) node * A + al[k]];
                    --k;
                }
                m_succ[(size_t) i * A + j] = result;
            }
        }
    });
    m_builder.markModelDirty();
}

void PatternManager::syncAutomaton() const
{
    if (!m_host_automaton_stale) return;
    hmc_ctx *c = m_builder.ctx();
    const int A = std::max(m_builder.genos()->Amax, 1);
    m_succ.assign((size_t) size() * A, -1);
    m_head_list.assign((size_t) m_dev_heads, -1);
    if (hmc_fetch_automaton(c, m_succ.data(), m_head_list.data()) < 0) throw HmcError(std::string("hmc_fetch_automaton: ") + hmc_last_error(c));
    m_host_automaton_stale = false;
}

void PatternManager::estimateFrequency()
{
    PatternSet copy = m_patterns;
    m_builder.estimateFrequency(copy);
    for (int i = 0; i < size(); ++i) {
        m_patterns.pats[i].frequency = copy.pats[i].frequency;
        m_patterns.pats[i].prefix_freq = copy.pats[i].prefix_freq;
        m_patterns.pats[i].setTransitionProb(copy.pats[i].transition_prob);
    }
    m_builder.markModelDirty();
}

// PatternManager::extendPatterns (PatternManager.cpp:412-438): up to four levels of one-allele extensions below every
// seed whose frequency reached min_freq; inner levels inherit the parent's frequency without an E-pass.
void PatternManager::extendPatterns(PatternSet &patterns, std::vector<int> &seeds, const PatternSet &last, const std::vector<int> &last_seeds)
{
    const GenoData &g = *m_builder.genos();
    const int L = g.L;
    // level 0 walks the seeds of the last round (in `last`), the later levels the patterns this call has just added (in
    // `patterns`); a candidate is (set, index): its string is read where it lies and copied once into the child
    std::vector<int> cur(last_seeds), nxt;
    const PatternSet *from = &last;
    std::vector<int8_t> buf;
    size_t room = 0;
    for (int sidx : last_seeds) room += (size_t) last.pats[sidx].length() + 1;
    patterns.alleles.reserve(patterns.alleles.size() + room * 4);
    int level = 4;
    while (level-- > 0) {
        nxt.clear();
        for (int idx : cur) {
            const Pattern hp = from->pats[idx];         // by value: `patterns` grows below
            const int len = hp.length(), end = hp.end;
            if (end < L && len < m_max_len[hp.start] && hp.frequency >= m_min_freq) {
                buf.assign(from->al(hp), from->al(hp) + len);
                buf.push_back(0);
                for (int j = 0; j < g.n_alleles[end]; ++j) {
                    buf[len] = (int8_t) j;
                    nxt.push_back(patterns.add(hp.start, buf.data(), len + 1, hp.frequency, 1.0, 1.0));
                }
            }
        }
        cur.swap(nxt);
        from = &patterns;
    }
    seeds = cur;            // the last level's new patterns are the next round's seeds
}

void PatternManager::estimatePatterns()
{
    if (m_min_freq < 0) { estimateFrequency(); return; }
    const GenoData &g = *m_builder.genos();
    const int L = g.L, A = std::max(g.Amax, 1);
    PatternSet patterns, candidates;
    std::vector<int> seeds;
    const std::vector<int32_t> &succ = successors();
    const double t_begin = now_s();
    std::vector<int8_t> al;
    patterns.pats.reserve((size_t) size() * 2); patterns.alleles.reserve(m_patterns.alleles.size() * 2 + (size_t) size());
    for (int i = 0; i < size(); ++i) {
        const Pattern &hp = m_patterns.pats[i];
        patterns.add(hp.start, m_patterns.al(hp), hp.length(), hp.frequency, hp.prefix_freq, hp.transition_prob);
        if (hp.end < L && hp.length() < m_max_len[hp.start]) {
            al.assign(m_patterns.al(hp), m_patterns.al(hp) + hp.length());
            al.push_back(0);
            for (int j = 0; j < g.n_alleles[hp.end]; ++j) {
                const int s = succ[(size_t) i * A + j];
                if (s < 0 || m_patterns.pats[s].start != hp.start) {    // unseen one-allele extension
                    al.back() = (int8_t) j;
                    seeds.push_back(patterns.add(hp.start, al.data(), (int) al.size(), hp.frequency, 1.0, 1.0));
                }
            }
        }
    }
    const bool timing = getenv("HMC_B200_HOST_TIMING") != nullptr;
    double t_mark = now_s();
    if (timing) fprintf(stderr, "[host] estimatePatterns: %d patterns + %zu unseen extensions prepared in %.3f s\n", size(), seeds.size(), t_mark - t_begin);
    while (!patterns.pats.empty()) {
        m_builder.estimateFrequency(patterns);
        const double t_e = now_s();
        {   // candidates += patterns (same order), strings appended in one piece
            const int64_t shift = (int64_t) candidates.alleles.size();
            candidates.alleles.insert(candidates.alleles.end(), patterns.alleles.begin(), patterns.alleles.end());
            const size_t n0 = candidates.pats.size();
            candidates.pats.insert(candidates.pats.end(), patterns.pats.begin(), patterns.pats.end());
            for (size_t q = n0; q < candidates.pats.size(); ++q) candidates.pats[q].al += shift;
        }
        PatternSet next;
        std::vector<int> next_seeds;
        extendPatterns(next, next_seeds, patterns, seeds);
        if (timing) fprintf(stderr, "[host]   E-pass over %zu candidates %.3f s, collect + extend %.3f s -> %zu new candidates\n", patterns.size(), t_e - t_mark,
                            now_s() - t_e, next.size());
        t_mark = now_s();
        patterns = std::move(next);
        seeds = std::move(next_seeds);
    }
    m_patterns.clear();
    for (const Pattern &hp : candidates.pats)
        if (hp.frequency >= m_min_freq || hp.length() <= m_min_len[hp.start])
            m_patterns.add(hp.start, candidates.al(hp), hp.length(), hp.frequency, hp.prefix_freq, hp.transition_prob);
    const double t_i = now_s();
    initialize();
    if (timing) fprintf(stderr, "[host]   selection %.3f s, initialize %.3f s, %d patterns kept\n", t_i - t_mark, now_s() - t_i, size());
}

// ---------------------------------------------------------------------------------------------
// HaploBuilder
HaploBuilder::HaploBuilder(int device, size_t arena_bytes, int threads)
    : m_patterns(*this), m_device(device), m_dev_threads(threads), m_arena_bytes(arena_bytes)
{
}

HaploBuilder::~HaploBuilder() { if (m_ctx) hmc_ctx_destroy(m_ctx); }

hmc_ctx *HaploBuilder::ctx()
{
    if (!m_ctx) {
        const int rc = hmc_ctx_create(&m_ctx, m_device, m_arena_bytes, m_dev_threads);
        if (rc != HMC_OK) { m_ctx = nullptr; throw HmcError("hmc_ctx_create failed (status " + std::to_string(rc) + "): a B200 (sm_100a) device is required, there is no CPU path"); }
    }
    return m_ctx;
}

static void ck(hmc_ctx *ctx, int rc, const char *what)
{
    if (rc < 0) throw HmcError(std::string(what) + ": " + hmc_last_error(ctx));
}

void HaploBuilder::setGenoData(GenoData &genos)
{
    m_genos = &genos;
    m_samples.clear();
    m_samples.L = genos.L;
    pg_exp.assign(genos.N, 0);
    m_panel_dirty = true;
    m_model_dirty = true;
}

void HaploBuilder::uploadPanel()
{
    if (!m_panel_dirty) return;
    const GenoData &genos = *m_genos;
    if (!genos.packed.empty())      // read from a binary panel: the rows are already in the device layout
        ck(ctx(), hmc_set_panel_packed(ctx(), genos.N, genos.L, genos.Amax, genos.packed_bits, genos.packed.data(), genos.n_alleles.data(), genos.allele_freq.data()),
           "hmc_set_panel_packed");
    else
        ck(ctx(), hmc_set_panel(ctx(), genos.N, genos.L, genos.Amax, genos.idx.data(), genos.n_alleles.data(), genos.allele_freq.data()), "hmc_set_panel");
    m_panel_dirty = false;
}

void HaploBuilder::uploadModel()
{
    uploadPanel();
    if (!m_model_dirty) return;
    const PatternSet &ps = m_patterns.patterns();
    const int P = (int) ps.size();
    std::vector<int32_t> start(P), end(P);
    std::vector<double> freq(P), tp(P);
    std::vector<int64_t> off(P + 1);
    std::vector<int8_t> al;
    al.reserve(ps.alleles.size());
    for (int i = 0; i < P; ++i) {
        const Pattern &p = ps.pats[i];
        start[i] = p.start; end[i] = p.end; freq[i] = p.frequency; tp[i] = p.transition_prob;
        off[i] = (int64_t) al.size();
        al.insert(al.end(), ps.al(p), ps.al(p) + p.length());
    }
    off[P] = (int64_t) al.size();
    const std::vector<int32_t> &heads = m_patterns.head_list();
    ck(m_ctx, hmc_set_model(m_ctx, P, m_patterns.head_len(), start.data(), end.data(), freq.data(), tp.data(), m_patterns.successors().data(),
                            off.data(), al.data(), (int) heads.size(), heads.data()), "hmc_set_model");
    m_model_dirty = false;
}

void HaploBuilder::resolveAll(int sample_size, ResolveResult &r, bool want_haps)
{
    uploadModel();
    const int N = m_genos->N, L = m_genos->L, K = sample_size > 1 ? sample_size : 1;
    const int first = m_shard_count >= 0 ? m_shard_first : 0, count = m_shard_count >= 0 ? m_shard_count : N;
    r.K = K;
    r.pg.assign(N, 0); r.coverage.assign(N, 0); r.exp2.assign(N, 0); r.nres.assign(N, 0);
    r.prior.assign((size_t) N * K, 0); r.posterior.assign((size_t) N * K, 0);
    if (want_haps) r.haps.assign((size_t) N * K * 2 * L, -1); else r.haps.clear();
    double ll_shard = 0;
    if (count < N && gather_device) {
        // sharded, exchange on the device: the result block has one row per panel individual, this rank fills its own rows, the
        // all-gather fills the others in place (7 arrays), and only then the host copies are made -- every rank ends up with every
        // individual's result and with all K sampled pairs still on its device for the next pattern search
        if (count > 0) ck(m_ctx, hmc_resolve_all_device(m_ctx, first, count, K, m_flags | HMC_FLAG_PANEL_ROWS), "hmc_resolve_all_device");
        else ck(m_ctx, hmc_resolve_all_device(m_ctx, 0, 1, K, m_flags | HMC_FLAG_PANEL_ROWS), "hmc_resolve_all_device");      // (a rank without individuals still owns a result block)
        void *ptrs[7];
        int64_t row_bytes[7];
        ck(m_ctx, hmc_result_device_ptrs(m_ctx, ptrs, row_bytes), "hmc_result_device_ptrs");
        for (int a = 0; a < 7; ++a) gather_device(ptrs[a], (size_t) row_bytes[a]);
        ck(m_ctx, hmc_fetch_results(m_ctx, r.pg.data(), r.exp2.data(), r.coverage.data(), r.nres.data(), r.prior.data(), r.posterior.data(),
                                    want_haps ? r.haps.data() : nullptr, nullptr), "hmc_fetch_results");
        r.ll = 0;
        for (int i = 0; i < N; ++i) r.ll += r.exp2[i] ? (std::log(r.pg[i]) - r.exp2[i] * M_LN2) : std::log(r.pg[i]);
        return;
    }
    if (count > 0)
        ck(m_ctx, hmc_resolve_all(m_ctx, first, count, K, m_flags, r.pg.data() + first, r.exp2.data() + first, r.coverage.data() + first, r.nres.data() + first,
                                  r.prior.data() + (size_t) first * K, r.posterior.data() + (size_t) first * K,
                                  want_haps ? r.haps.data() + (size_t) first * K * 2 * L : nullptr, &ll_shard), "hmc_resolve_all");
    if (count < N && gather) {      // every rank ends up with every individual's result, in individual order
        auto all = [&](auto &v, size_t per) { std::vector<typename std::remove_reference<decltype(v)>::type::value_type> full(v.size());
                                              gather(v.data() + (size_t) first * per, full.data(), per * sizeof(v[0])); v.swap(full); };
        all(r.pg, 1); all(r.exp2, 1); all(r.coverage, 1); all(r.nres, 1); all(r.prior, K); all(r.posterior, K);
        if (want_haps) all(r.haps, (size_t) K * 2 * L);
    }
    // HaploModel.cpp:110: sum of log P(G) in individual order (identical on every rank and for every shard count)
    r.ll = 0;
    for (int i = 0; i < N; ++i) r.ll += r.exp2[i] ? (std::log(r.pg[i]) - r.exp2[i] * M_LN2) : std::log(r.pg[i]);
    if (count == N) r.ll = ll_shard;
}

void HaploBuilder::materializeSamples()
{
    if (!m_samples.on_device()) return;
    const int N = m_genos->N, L = m_genos->L;
    std::vector<int8_t> all;
    int rk = 0;
    ck(m_ctx, hmc_result_dims(m_ctx, nullptr, &rk), "hmc_result_dims");
    all.resize((size_t) N * rk * 2 * L);
    ck(m_ctx, hmc_fetch_results(m_ctx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, all.data(), nullptr), "hmc_fetch_results");
    m_samples.h.resize(m_samples.rows.size() * (size_t) L);
    for (size_t s = 0; s < m_samples.rows.size(); ++s) memcpy(&m_samples.h[s * L], &all[(size_t) m_samples.rows[s] * L], (size_t) L);
    m_samples.rows.clear();
}

void HaploBuilder::estimateFrequency(PatternSet &patterns)
{
    materializeSamples();       // the E-pass reuses the device result block the sample rows point into
    uploadModel();
    const int Pc = (int) patterns.size(), N = m_genos->N;
    std::vector<int32_t> start(Pc);
    std::vector<int64_t> off(Pc + 1);
    std::vector<int8_t> al;
    al.reserve(patterns.alleles.size());
    for (int i = 0; i < Pc; ++i) {
        const Pattern &p = patterns.pats[i];
        start[i] = p.start; off[i] = (int64_t) al.size();
        al.insert(al.end(), patterns.al(p), patterns.al(p) + p.length());
    }
    off[Pc] = (int64_t) al.size();
    const int first = m_shard_count >= 0 ? m_shard_first : 0, count = m_shard_count >= 0 ? m_shard_count : N;
    std::vector<double> raw((size_t) 2 * Pc, 0.0), tp(Pc);
    double *f = raw.data(), *pf = raw.data() + Pc;
    if (count < N && reduce_counts_device && count > 0) {
        // the one exchange of an exact E-pass, on the device: the raw sums stay in HBM, are all-reduced in place and come back once
        double *dev = nullptr;
        ck(m_ctx, hmc_estimate_counts_device(m_ctx, first, count, m_flags, Pc, start.data(), off.data(), al.data(), m_genos->genotype_probability.data() + first,
                                             (m_flags & HMC_FLAG_SCALED) ? pg_exp.data() + first : nullptr, &dev), "hmc_estimate_counts_device");
        reduce_counts_device(dev, (size_t) 2 * Pc);
        ck(m_ctx, hmc_fetch_counts(m_ctx, Pc, f, pf), "hmc_fetch_counts");
    }
    else {
    if (count > 0)
        ck(m_ctx, hmc_estimate_counts(m_ctx, first, count, m_flags, Pc, start.data(), off.data(), al.data(), m_genos->genotype_probability.data() + first,
                                      (m_flags & HMC_FLAG_SCALED) ? pg_exp.data() + first : nullptr, f, pf), "hmc_estimate_counts");
    if (count < N && reduce_counts) reduce_counts(raw.data(), raw.size());     // the one exchange of an exact E-pass
    }
    hmc_normalise_counts(Pc, N, f, pf, tp.data());
    for (int i = 0; i < Pc; ++i) { patterns.pats[i].frequency = f[i]; patterns.pats[i].prefix_freq = pf[i]; patterns.pats[i].setTransitionProb(tp[i]); }
}

// ---------------------------------------------------------------------------------------------
// HaploModel
void HaploModel::setModel(const std::string &model)
{
    if (model == "MV" || model == "MC" || model == "MA") m_model = model;
    else throw HmcError("Unknown model " + model + "!");
}

void HaploModel::setCliDefaults()
{
    min_freq_abs = 1.5; min_pattern_len = 1; max_pattern_len = 30; mc_order = 1;
    sample_size = 10; final_sample_size = 1; max_iteration = 1; min_iteration = -1;
}

void HaploModel::build(GenoData &genos)
{
    setGenoData(genos);
    findPatterns();
}

void HaploModel::findPatterns()
{
    const double t0 = now_s();
    if (min_freq_abs > 0) min_freq = min_freq_abs / (2.0 * genos()->N);
    if (m_model == "MV" || m_model == "MA") {       // MA only adds adjustFrequency, which just logs (PatternManager.cpp:320-345)
        if (num_patterns > 0) m_patterns.findPatternByNum(num_patterns, min_pattern_len, max_pattern_len);
        else m_patterns.findPatternByFreq(min_freq, min_pattern_len, max_pattern_len);
    }
    else if (m_model == "MC") m_patterns.findPatternBlock(mc_order + 1);
    seconds[0] += now_s() - t0;
    if (verbose) fprintf(stderr, "Found haplotype patterns: %d\n", m_patterns.size());
}

double HaploModel::resolveAll(std::vector<int8_t> &resolved)
{
    GenoData &g = *m_genos;
    const int N = g.N, L = g.L;
    ResolveResult r;
    double t0 = now_s();
    m_device_rows_current = false;
    // One GPU holds every individual and the next search runs on it: the K sampled pairs stay on the device and are named by
    // their result rows (hmc_mine_patterns_resolved); only the best pair of every individual comes back.
    // (sharded runs with a device-side all-gather hold every individual's rows on every rank too)
    const bool dev_samples = m_patterns.deviceMiningWanted() && num_patterns <= 0 && !exact_estimate &&
                             (m_shard_count < 0 || m_shard_count == N || (bool) gather_device) && !getenv("HMC_B200_HOST_SAMPLES");
    HaploBuilder::resolveAll(sample_size, r, !dev_samples);
    std::vector<int8_t> best;
    if (dev_samples) {
        best.resize((size_t) N * 2 * L);
        ck(m_ctx, hmc_fetch_best_haps(m_ctx, best.data()), "hmc_fetch_best_haps");
    }
    seconds[1] += now_s() - t0;
    t0 = now_s();
    const int K = r.K;
    m_samples.clear();
    m_samples.L = L;
    resolved.assign((size_t) N * 2 * L, -1);
    for (int i = 0; i < N; ++i) {
        const int n = r.nres[i];
        if (n == 0) {       // unable to resolve: the resolution is the input genotype (HaploBuilder.cpp:117-124)
            if (verbose) fprintf(stderr, "Unable to resolve Genotype[%d]: %s!\n", i, g.ids[i].c_str());
            memcpy(&resolved[(size_t) i * 2 * L], &g.idx[(size_t) i * 2 * L], (size_t) 2 * L);
        }
        else {
            memcpy(&resolved[(size_t) i * 2 * L], dev_samples ? &best[(size_t) i * 2 * L] : &r.haps[((size_t) i * K) * 2 * L], (size_t) 2 * L);
            for (int j = 0; j < n; ++j) {
                const double wgt = r.posterior[(size_t) i * K + j] / r.coverage[i];
                if (dev_samples) { m_samples.rows.push_back((i * K + j) * 2); m_samples.rows.push_back((i * K + j) * 2 + 1); }
                else {
                    const int8_t *h = &r.haps[((size_t) i * K + j) * 2 * L];
                    m_samples.h.insert(m_samples.h.end(), h, h + 2 * L);
                }
                m_samples.w.push_back(wgt);
                m_samples.w.push_back(wgt);
            }
        }
        g.genotype_probability[i] = r.pg[i];
        pg_exp[i] = r.exp2[i];
    }
    m_samples.total_weight = 0;
    for (double w : m_samples.w) m_samples.total_weight += w;      // HaploData::checkTotalWeight
    seconds[2] += now_s() - t0;
    return r.ll;
}

bool HaploModel::writeReconstructed(const HaploFile &file, const GenoData &genos, const std::vector<int8_t> &res, const char *suffix)
{
    int32_t rows = 0, K = 0;
    if (getenv("HMC_B200_HOST_FORMAT") || !m_device_rows_current || hmc_result_dims(m_ctx, &rows, &K) != HMC_OK || rows != genos.N) {
        file.writeGenoData(genos, res, suffix);
        return false;
    }
    const int L = genos.L, A = genos.Amax;
    std::vector<int32_t> sym((size_t) L * A, 0);
    for (int l = 0; l < L; ++l) for (size_t k = 0; k < genos.symbols[l].size() && (int) k < A; ++k) sym[(size_t) l * A + k] = genos.symbols[l][k];
    std::vector<int64_t> off((size_t) 2 * rows + 1);
    int64_t total = 0;
    ck(m_ctx, hmc_format_best(m_ctx, sym.data(), genos.allele_type.c_str(), nullptr, 0, nullptr, &total), "hmc_format_best");
    std::vector<char> text((size_t) total);
    ck(m_ctx, hmc_format_best(m_ctx, sym.data(), genos.allele_type.c_str(), text.data(), total, off.data(), &total), "hmc_format_best");
    file.writeGenoText(genos, text.data(), off.data(), suffix);
    return true;
}

// ---------------------------------------------------------------------------------------------
// HaploComp
HaploComp::HaploComp(HaploBuilder &hb, int N, int L, const int8_t *real, const int8_t *infer)
{
    hmc_ctx *c = hb.ctx();
    int32_t bi = -1, bl = -1;
    if (hmc_compare(c, N, L, real, infer, counts, &bi, &bl) < 0) throw HmcError(hmc_last_error(c));
}

HaploComp::HaploComp(HaploBuilder &hb, const GenoData &real, const GenoData &infer)
{
    if (real.N != infer.N || real.L != infer.L) throw HmcError("Attempt to compare inconsistent haplotype data!");      // HaploComp.cpp:36-40
    const int N = real.N, L = real.L;
    std::vector<int8_t> a((size_t) N * 2 * L), b((size_t) N * 2 * L);
    for (int l = 0; l < L; ++l) {
        std::vector<int32_t> u(real.symbols[l]);
        u.insert(u.end(), infer.symbols[l].begin(), infer.symbols[l].end());
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        if (u.size() > 127) throw HmcError("too many distinct alleles at locus " + std::to_string(l));
        auto code = [&](int v) { return v < 0 ? -1 : (int) (std::lower_bound(u.begin(), u.end(), v) - u.begin()); };
        for (int i = 0; i < N; ++i) for (int h = 0; h < 2; ++h) {
            a[((size_t) i * 2 + h) * L + l] = (int8_t) code(real.allele(i, h, l));
            b[((size_t) i * 2 + h) * L + l] = (int8_t) code(infer.allele(i, h, l));
        }
    }
    hmc_ctx *c = hb.ctx();
    int32_t bi = -1, bl = -1;
    if (hmc_compare(c, N, L, a.data(), b.data(), counts, &bi, &bl) < 0) throw HmcError(hmc_last_error(c));
}

void HaploModel::run(const GenoData &genos, std::vector<int8_t> &resolutions)
{
    m_unphased = genos;
    build(m_unphased);
    resolutions = m_unphased.idx;
    std::vector<int8_t> resolved;
    ll_trace.clear(); npat_trace.clear();
    double old_ll = -DBL_MAX;
    for (int iter = 1; iter <= max_iteration; ++iter) {
        npat_trace.push_back(m_patterns.size());
        const double ll = resolveAll(resolved);
        ll_trace.push_back(ll);
        if (ll >= old_ll) resolutions = resolved;
        m_device_rows_current = ll >= old_ll;          // the rows on the device are the resolutions just adopted (or a worse pass that was not)
        if (verbose) {      // HaploModel.cpp:134-137
            HaploComp compare(*this, m_unphased.N, m_unphased.L, m_unphased.idx.data(), resolutions.data());
            fprintf(stderr, "  iteration %d: patterns %d, Switch Error = %f, IHP = %f, IGP = %f, LL = %f\n", iter, m_patterns.size(), compare.switch_error(),
                    compare.incorrect_haplotype_percentage(), compare.incorrect_genotype_percentage(), ll);
        }
        if (iter < max_iteration && ll >= old_ll && (old_ll - ll) / old_ll > 0.0001) {
            if (exact_estimate) { const double t0 = now_s(); m_patterns.estimatePatterns(); seconds[2] += now_s() - t0; }
            else findPatterns();
            old_ll = ll;
        }
        else break;
    }
}

}  // namespace hmc_host
