"""Golden cases: panel recipe, reference CLI options (oracle/ref_driver.cpp) and the same options
as keyword arguments for the oracle port / the product host layer."""

# name: (panel kwargs, PHASE kind, half-missing fraction, reference driver options, model options)
CASES = {
    "g1_clean":   (dict(n_ind=40, n_loci=30, founders=6, seed=11), "S", 0.0,
                   ["--dump-states", "--dump-beta", "--dump-estep"], dict()),
    "g2_missing": (dict(n_ind=36, n_loci=32, founders=8, missing=0.06, seed=12), "S", 0.02,
                   ["--dump-states", "--dump-beta", "--dump-estep"], dict()),
    "g3_multi":   (dict(n_ind=50, n_loci=30, alleles=4, founders=8, missing=0.03, seed=13), "M", 0.0,
                   ["--dump-states"], dict()),
    "g4_em":      (dict(n_ind=60, n_loci=40, founders=6, missing=0.03, seed=14), "S", 0.0,
                   ["-i", "6", "--no-klist"], dict(max_iteration=6)),
    "g5_exact":   (dict(n_ind=30, n_loci=20, founders=4, missing=0.03, seed=15), "S", 0.0,
                   ["-i", "4", "--exact-estimate", "--no-klist"], dict(max_iteration=4, exact_estimate=True)),
    "g6_mc":      (dict(n_ind=40, n_loci=30, founders=6, seed=16), "S", 0.0,
                   ["-m", "MC", "-o", "2", "-i", "3", "--exact-estimate", "--dump-estep"],
                   dict(model="MC", mc_order=2, max_iteration=3, exact_estimate=True)),
    "g7_bynum":   (dict(n_ind=40, n_loci=30, founders=6, seed=17), "S", 0.0,
                   ["-n", "400", "--no-klist"], dict(num_patterns=400)),
    "g8_head2":   (dict(n_ind=40, n_loci=24, alleles=3, founders=6, missing=0.10, seed=18), "M", 0.05,
                   ["--min-pattern-len", "2", "--dump-states", "--dump-beta", "--dump-estep"], dict(min_pattern_len=2)),
    "g9_k1":      (dict(n_ind=40, n_loci=30, founders=6, missing=0.04, seed=19), "S", 0.0,
                   ["--sample-size", "1", "--max-pattern-len", "8"], dict(sample_size=1, max_pattern_len=8)),
}
