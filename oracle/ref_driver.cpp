// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the product path.
//
// Driver for the UNMODIFIED reference HMC core (built by oracle/build_ref.sh into oracle/_ref/).
// It replaces HMC.cpp/Main.cpp/Options.cpp (which need boost::program_options) with a hand-rolled
// option loop using the same option names and defaults (HMC.cpp:19-56), mirrors HaploModel::run
// (HaploModel.cpp:117-155) statement by statement -- but keeps the working GenoData alive (the
// reference hands the builder a pointer to a local) -- and dumps what the parity tests
// need: allele tables, the pattern automaton, per-individual P(G)/K-list/trellis, E-step counts.
// Every number printed is produced by the reference's own functions; the driver does no arithmetic
// of its own except wall-clock timing (std::chrono) and integer counters read from the trellis.
//
// Dump format (text, doubles as C99 hex floats so they round-trip exactly): see tests/refdump.py.

#include "HaploModel.h"
#include "HaploFile.h"
#include "HaploComp.h"
#include "GenoData.h"

#include <chrono>
#include <cfloat>
#include <string>
#include <vector>

using namespace std;

static double now_s()
{
	return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

struct DriverOpts {
	string input, dump, out_suffix;
	bool dump_states = false, dump_beta = false, dump_estep = false, dump_patterns = true, dump_klist = true, dump_likelihood = false;
	int time_resolve = 0, time_estep = 0, time_offset = 0, time_steps = 1, time_warmup = 0;
	DriverOpts() : out_suffix(".reconstructed") {}
};

class Probe : public HaploModel {
public:
	GenoData unphased, resolved;
	FILE *fp;
	DriverOpts opt;

	Probe() : fp(0) {}

	int alleleIndex(int locus, const Allele &a) const { return a.isMissing() ? -1 : m_genos->getAlleleIndex(locus, a); }

	void dumpPanel()
	{
		if (!fp) return;
		const GenoData &g = *m_genos;
		fprintf(fp, "panel %d %d %d\n", g.genotype_num(), g.genotype_len(), g.max_allele_num());
		for (int l = 0; l < g.genotype_len(); ++l) {
			fprintf(fp, "locus %d %c %d", l, g.allele_type(l), g.allele_num(l));
			for (int j = 0; j < g.allele_num(l); ++j) fprintf(fp, " %d %a", g.allele_symbol(l, j).asInt(), g.allele_frequency(l, j));
			fprintf(fp, "\n");
		}
		for (int i = 0; i < g.genotype_num(); ++i) {
			fprintf(fp, "geno %d %s", i, g[i].id().c_str());
			for (int h = 0; h < 2; ++h) {
				fprintf(fp, " ");
				for (int l = 0; l < g.genotype_len(); ++l) fprintf(fp, "%d%s", alleleIndex(l, g[i](h)[l]), l + 1 < g.genotype_len() ? "," : "");
			}
			fprintf(fp, "\n");
		}
	}

	void dumpPatterns(const char *tag)
	{
		if (!fp) return;
		int pn = pattern_num();
		fprintf(fp, "patterns %s %d %d %d\n", tag, pn, m_patterns.head_len(), (int) m_patterns.head_list().size());
		if (!opt.dump_patterns) return;
		for (int i = 0; i < pn; ++i) {
			const HaploPattern *hp = m_patterns[i];
			fprintf(fp, "p %u %d %d %a %a %a", hp->id(), hp->start(), hp->end(), hp->frequency(), hp->prefix_freq(), hp->transition_prob());
			int an = hp->end() < genotype_len() ? m_genos->allele_num(hp->end()) : 0;
			fprintf(fp, " %d", an);
			for (int j = 0; j < an; ++j) {
				const HaploPattern *s = hp->successors(j);
				fprintf(fp, " %d", s ? (int) s->id() : -1);
			}
			fprintf(fp, " ");
			for (int k = 0; k < hp->length(); ++k) fprintf(fp, "%d%s", hp->getAlleleIndex(k), k + 1 < hp->length() ? "," : "");
			fprintf(fp, "\n");
		}
		fprintf(fp, "heads");
		for (size_t i = 0; i < m_patterns.head_list().size(); ++i) fprintf(fp, " %u", m_patterns.head_list()[i]->id());
		fprintf(fp, "\n");
	}

	// one more resolve() per individual on the current model, to expose res_list and the trellis
	void dumpIndividuals(int sample)
	{
		if (!fp) return;
		Genotype res;
		vector<Genotype> res_list;
		long long tot_states = 0, tot_arrivals = 0, max_states = 0, steps = 0;
		for (int i = 0; i < unphased.genotype_num(); ++i) {
			double cov = resolve(unphased[i], res, res_list, sample);
			fprintf(fp, "ind %d %a %a %d\n", i, res.genotype_probability(), cov, (int) res_list.size());
			if (opt.dump_klist) {
				for (size_t k = 0; k < res_list.size(); ++k) {
					fprintf(fp, "res %d %a %a", (int) k, res_list[k].prior_probability(), res_list[k].posterior_probability());
					for (int h = 0; h < 2; ++h) {
						fprintf(fp, " ");
						for (int l = 0; l < genotype_len(); ++l) fprintf(fp, "%d%s", alleleIndex(l, res_list[k](h)[l]), l + 1 < genotype_len() ? "," : "");
					}
					fprintf(fp, "\n");
				}
			}
			if (opt.dump_beta) calcBackwardLikelihood();
			long long st = 0, arr = 0, mx = 0;
			for (int l = m_patterns.head_len(); l <= genotype_len(); ++l) {
				const vector<HaploPair*> &v = m_haplopairs[l];
				if (l > m_patterns.head_len() && !v.empty()) steps++;
				st += v.size();
				if ((long long) v.size() > mx) mx = v.size();
				for (size_t s = 0; s < v.size(); ++s) arr += v[s]->forward_links(0).size() + v[s]->forward_links(1).size();
				if (opt.dump_states) {
					fprintf(fp, "states %d %d %d\n", i, l, (int) v.size());
					for (size_t s = 0; s < v.size(); ++s) {
						fprintf(fp, "s %d %d %a %a", v[s]->id_a(), v[s]->id_b(), v[s]->forward_likelihood(), v[s]->transition_prob());
						if (opt.dump_beta) fprintf(fp, " %a", v[s]->backward_likelihood());
						fprintf(fp, "\n");
					}
				}
			}
			fprintf(fp, "cnt %d %lld %lld %lld\n", i, st, mx, arr);
			tot_states += st; tot_arrivals += arr; if (mx > max_states) max_states = mx;
		}
		fprintf(fp, "counters %lld %lld %lld %lld\n", steps, tot_states, max_states, tot_arrivals);
	}

	// PatternManager::estimateFrequency (PatternManager.cpp:347-362) on clones, dumped instead of stored
	void dumpEstep()
	{
		if (!fp) return;
		vector<HaploPattern*> patterns;
		int n = pattern_num();
		for (int i = 0; i < n; ++i) patterns.push_back(new HaploPattern(*m_patterns[i]));
		double t0 = now_s();
		estimateFrequency(patterns);
		double t1 = now_s();
		fprintf(fp, "estep %d %a\n", n, t1 - t0);
		for (int i = 0; i < n; ++i) fprintf(fp, "e %d %a %a %a\n", i, patterns[i]->frequency(), patterns[i]->prefix_freq(), patterns[i]->transition_prob());
		DeleteAll_Clear()(patterns);
	}

	void run2(const GenoData &genos, GenoData &resolutions)
	{
		int iter;
		double ll, old_ll;
		unphased = resolved = genos;
		double t0 = now_s();
		build(unphased);
		double t_build = now_s() - t0;
		resolutions = unphased;
		dumpPanel();
		if (fp) fprintf(fp, "time build %a\n", t_build);
		fprintf(stderr, "[hmc_ref] build %.3f s, patterns %d\n", t_build, pattern_num());

		old_ll = -DBL_MAX;
		for (iter = 1; iter <= max_iteration; ++iter) {
			if (fp) fprintf(fp, "iter %d\n", iter);
			dumpPatterns("model");
			t0 = now_s();
			ll = resolveAll(unphased, resolved);
			double t_res = now_s() - t0;
			if (ll >= old_ll) resolutions = resolved;
			HaploComp compare(&genos, &resolutions);
			if (fp) {
				fprintf(fp, "ll %d %a\n", iter, ll);
				fprintf(fp, "time resolveAll %a\n", t_res);
				fprintf(fp, "accuracy %a %a %a\n", compare.switch_error(), compare.incorrect_haplotype_percentage(), compare.incorrect_genotype_percentage());
				fprintf(fp, "pg");
				for (int i = 0; i < unphased.genotype_num(); ++i) fprintf(fp, " %a", unphased[i].genotype_probability());
				fprintf(fp, "\n");
			}
			fprintf(stderr, "[hmc_ref] iter %d  resolveAll %.3f s  LL = %f  switch = %f\n", iter, t_res, ll, compare.switch_error());
			if (opt.dump_klist || opt.dump_states) dumpIndividuals(sample_size);
			if (opt.dump_estep) dumpEstep();
			if (opt.dump_likelihood && fp) {	// HaploBuilder::getLikelihood (HaploBuilder.cpp:128-151) of both haplotypes of every resolution, under this iteration's model
				for (int i = 0; i < resolved.genotype_num(); ++i)
					fprintf(fp, "lk %d %a %a %a\n", i, getLikelihood(resolved[i](0)), getLikelihood(resolved[i](1)), getLikelihood(resolved[i]));
			}

			if (iter < max_iteration && ll >= old_ll && (old_ll - ll) / old_ll > 0.0001) {
				t0 = now_s();
				if (exact_estimate) {
					m_patterns.estimatePatterns();
				}
				else {
					findPatterns();
				}
				if (fp) fprintf(fp, "time mstep %a\n", now_s() - t0);
				old_ll = ll;
				if (m_model == "MA") {
					m_patterns.adjustFrequency();
				}
			}
			else {
				break;
			}
		}
		if (fp) fprintf(fp, "end %d\n", iter);
	}

	// CPU-baseline legs: the model is mined on the full panel; the timed calls see the first n individuals
	void timeLegs(const GenoData &genos)
	{
		unphased = genos;
		double t0 = now_s();
		build(unphased);
		double t_build = now_s() - t0;
		int head_len = m_patterns.head_len();
		printf("{\"build_s\": %.6f, \"patterns\": %d, \"N\": %d, \"L\": %d", t_build, pattern_num(), genos.genotype_num(), genos.genotype_len());
		if (opt.time_resolve > 0) {
			// sub-panel = individuals [offset, offset+n) of the full panel, allele tables and model of the full panel
			int off = min(opt.time_offset, genos.genotype_num() - 1);
			int n = min(opt.time_resolve, genos.genotype_num() - off);
			GenoData sub = unphased;
			for (int i = 0; i < n; ++i) sub[i] = unphased[off + i];
			sub.setGenotypeNum(n);
			GenoData res = sub;
			setGenoData(sub);
			printf(", \"resolve_n\": %d, \"resolve_offset\": %d, \"resolve_steps\": %lld, \"resolve_s\": [", n, off, (long long) n * (genos.genotype_len() - head_len));
			double ll = 0;
			for (int it = 0; it < opt.time_warmup + opt.time_steps; ++it) {
				t0 = now_s();
				ll = resolveAll(sub, res);
				double t = now_s() - t0;
				if (it >= opt.time_warmup) printf("%s%.6f", it > opt.time_warmup ? ", " : "", t);
			}
			printf("], \"resolve_ll\": %.17g", ll);
			if (opt.time_estep > 0) {
				int m = min(opt.time_estep, n);
				sub.setGenotypeNum(m);
				vector<HaploPattern*> patterns;
				for (int i = 0; i < pattern_num(); ++i) patterns.push_back(new HaploPattern(*m_patterns[i]));
				t0 = now_s();
				estimateFrequency(patterns);
				double t = now_s() - t0;
				printf(", \"estep_n\": %d, \"estep_s\": %.6f, \"estep_steps\": %lld", m, t, (long long) m * (genos.genotype_len() - head_len));
				DeleteAll_Clear()(patterns);
			}
			setGenoData(unphased);
		}
		printf("}\n");
	}
};

static void usage()
{
	fprintf(stderr,
		"usage: hmc_ref [options] phasefile\n"
		"  model options (names and defaults of HMC.cpp:19-56): -m/--model MV|MC|MA, -r/--min-freq-rel x,\n"
		"    -a/--min-freq-abs 1.5, -n/--num-patterns n, --min-pattern-len 1, --max-pattern-len 30, -o/--mc-order 1,\n"
		"    --exact-estimate, --sample-size 10, -i/--max-iteration 1, -d/--debug 2\n"
		"  driver options: --dump FILE [--dump-states] [--dump-beta] [--dump-estep] [--dump-likelihood] [--no-patterns] [--no-klist]\n"
		"    --out-suffix S (default .reconstructed; 'none' to skip), --time-resolve n [--time-estep m]\n"
		"    --compare REAL.phase: HaploComp of the positional (inferred) file against REAL, printed as hex floats\n");
}

int main(int argc, char **argv)
{
	Probe m;
	// HMC.cpp:35-47 defaults
	string model = "MV";
	m.min_freq_abs = 1.5;
	m.min_pattern_len = 1;
	m.max_pattern_len = 30;
	m.mc_order = 1;
	m.sample_size = 10;
	m.final_sample_size = 1;
	m.max_iteration = 1;
	m.min_iteration = -1;
	int debug = 2;
	string compare_with;
	for (int i = 1; i < argc; ++i) {
		string a = argv[i];
		auto next = [&]() -> const char * { if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
		if (a == "-m" || a == "--model") model = next();
		else if (a == "-r" || a == "--min-freq-rel") m.min_freq = atof(next());
		else if (a == "-a" || a == "--min-freq-abs") m.min_freq_abs = atof(next());
		else if (a == "-n" || a == "--num-patterns") { m.num_patterns = atoi(next()); }
		else if (a == "--min-pattern-len") m.min_pattern_len = atoi(next());
		else if (a == "--max-pattern-len") m.max_pattern_len = atoi(next());
		else if (a == "-o" || a == "--mc-order") m.mc_order = atoi(next());
		else if (a == "--exact-estimate") m.exact_estimate = true;
		else if (a == "--sample-size") m.sample_size = atoi(next());
		else if (a == "-i" || a == "--max-iteration") m.max_iteration = atoi(next());
		else if (a == "-d" || a == "--debug") debug = atoi(next());
		else if (a == "--dump") m.opt.dump = next();
		else if (a == "--dump-states") m.opt.dump_states = true;
		else if (a == "--dump-beta") m.opt.dump_beta = true;
		else if (a == "--dump-estep") m.opt.dump_estep = true;
		else if (a == "--dump-likelihood") m.opt.dump_likelihood = true;
		else if (a == "--no-patterns") m.opt.dump_patterns = false;
		else if (a == "--no-klist") m.opt.dump_klist = false;
		else if (a == "--out-suffix") m.opt.out_suffix = next();
		else if (a == "--time-resolve") m.opt.time_resolve = atoi(next());
		else if (a == "--time-estep") m.opt.time_estep = atoi(next());
		else if (a == "--time-offset") m.opt.time_offset = atoi(next());
		else if (a == "--time-steps") m.opt.time_steps = atoi(next());
		else if (a == "--time-warmup") m.opt.time_warmup = atoi(next());
		else if (a == "--compare") compare_with = next();
		else if (a[0] == '-') { usage(); return 2; }
		else m.opt.input = a;
	}
	if (m.opt.input.empty()) { usage(); return 2; }
	Logger::setLogLevel(debug);
	m.setModel(model);

	HaploFile file(m.opt.input);
	GenoData genos, resolutions;
	file.readGenoData(genos);

	if (!compare_with.empty()) {
		// HMC::compare (HMC.cpp:246-257): the positional file is the inferred data, --compare names the real (target) data
		GenoData target;
		HaploFile(compare_with).readGenoData(target);
		HaploComp hc(&target, &genos);
		printf("{\"switch_error\": \"%a\", \"ihp\": \"%a\", \"igp\": \"%a\"}\n", hc.switch_error(), hc.incorrect_haplotype_percentage(),
		       hc.incorrect_genotype_percentage());
		return 0;
	}
	if (m.opt.time_resolve > 0) {
		m.timeLegs(genos);
		return 0;
	}
	if (!m.opt.dump.empty()) {
		m.fp = fopen(m.opt.dump.c_str(), "w");
		if (!m.fp) { fprintf(stderr, "cannot open %s\n", m.opt.dump.c_str()); return 1; }
		fprintf(m.fp, "HMCDUMP 1\n");
	}
	m.run2(genos, resolutions);
	if (m.fp) fclose(m.fp);
	if (m.opt.out_suffix != "none") file.writeGenoData(resolutions, m.opt.out_suffix.c_str());
	return 0;
}
