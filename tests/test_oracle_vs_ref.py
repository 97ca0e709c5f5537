"""The oracle port against the UNMODIFIED reference binary run live on fresh random panels (beyond the committed dumps).
Needs oracle/_ref/hmc_ref (built by oracle/build_ref.sh where /root/reference exists); skipped otherwise.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

import refdump
from hmc_b200.synth import make_panel, write_phase
from oracle.pyoracle import Oracle

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "hmc_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/hmc_ref not built")

CASES = [
    (dict(n_ind=30, n_loci=25, founders=5, missing=0.05, seed=101), "S", ["--sample-size", "4"], dict(sample_size=4)),
    (dict(n_ind=24, n_loci=20, alleles=5, founders=6, missing=0.08, seed=102), "M", ["--max-pattern-len", "6"], dict(max_pattern_len=6)),
    (dict(n_ind=36, n_loci=28, founders=4, seed=103), "S", ["-i", "4"], dict(max_iteration=4)),
    (dict(n_ind=20, n_loci=16, founders=3, missing=0.04, seed=104), "S", ["-i", "3", "--exact-estimate"], dict(max_iteration=3, exact_estimate=True)),
]


@pytest.mark.parametrize("kw,kind,ref_opts,opts", CASES)
def test_live_reference_run(tmp_path, kw, kind, ref_opts, opts):
    g = make_panel(**kw)
    phase = str(tmp_path / "p.phase")
    dump = str(tmp_path / "p.dump")
    write_phase(phase, g, kind=kind)
    subprocess.check_call([REF, phase, "--dump", dump, "--out-suffix", "none", "--dump-states", "-d", "1"] + ref_opts, stderr=subprocess.DEVNULL)
    d = refdump.load(dump)
    o = Oracle(d["geno"], **opts)
    if opts.get("max_iteration", 1) > 1:
        ll, npat = o.run()
        ref_ll = np.array([it["ll"] for it in d["iters"]])
        assert len(ll) == len(ref_ll) and np.array_equal(npat, [it["P"] for it in d["iters"]])
        if opts.get("exact_estimate"):
            assert np.allclose(ll, ref_ll, rtol=1e-9, atol=0)
        else:
            assert np.array_equal(ll, ref_ll)
        return
    o.find_patterns()
    m, r = o.export_model(), d["iters"][0]["patterns"]
    for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads"):
        assert np.array_equal(m[k], r[k]), k
    assert o.resolve_all() == d["iters"][0]["ll"]
    assert o.counters() == d["iters"][0]["counters"]
    for i in range(d["N"]):
        a, b = o.individual(i), d["iters"][0]["inds"][i]
        assert a["pg"] == b["pg"] and np.array_equal(a["prior"], np.array(b["prior"]))
        for k in range(a["n"]):
            assert np.array_equal(a["haps"][k], b["haps"][k])


def test_config0_c1_200x100_defaults(tmp_path):
    """BASELINE.json configs[0]: 200 individuals x 100 SNPs, default options -- oracle port == reference binary, bit for bit."""
    g = make_panel(200, 100, founders=8, seed=1)
    phase = str(tmp_path / "c1.phase")
    dump = str(tmp_path / "c1.dump")
    write_phase(phase, g, kind="S")
    subprocess.check_call([REF, phase, "--dump", dump, "--out-suffix", "none", "-d", "1"], stderr=subprocess.DEVNULL)
    d = refdump.load(dump)
    o = Oracle(d["geno"])
    o.find_patterns()
    m, r = o.export_model(), d["iters"][0]["patterns"]
    assert m["P"] == r["P"] == 49878
    for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads"):
        assert np.array_equal(m[k], r[k]), k
    assert o.resolve_all() == d["iters"][0]["ll"]
    assert o.counters() == d["iters"][0]["counters"]
    for i in range(200):
        a, b = o.individual(i), d["iters"][0]["inds"][i]
        assert a["pg"] == b["pg"] and np.array_equal(a["prior"], np.array(b["prior"]))
        for k in range(a["n"]):
            assert np.array_equal(a["haps"][k], b["haps"][k])


@pytest.mark.parametrize("kw,kind", [(dict(n_ind=40, n_loci=30, founders=5, missing=0.05, seed=111), "S"),
                                     (dict(n_ind=30, n_loci=24, alleles=5, founders=6, missing=0.04, seed=112), "M")])
def test_get_likelihood_of_the_reference(tmp_path, kw, kind):
    """HaploBuilder::getLikelihood (HaploBuilder.cpp:128-151) is dead code in the reference's CLI; the driver calls it on both haplotypes of
    every resolution and on the genotype.  The oracle's restatement (one trie descent per index) equals it bit for bit -- and the device's
    automaton walk equals the oracle's (tests/test_gpu_emit.py), which pins hmc_get_likelihood to the reference."""
    g = make_panel(**kw)
    phase = str(tmp_path / "p.phase")
    dump = str(tmp_path / "p.dump")
    write_phase(phase, g, kind=kind)
    subprocess.check_call([REF, phase, "--dump", dump, "--out-suffix", "none", "--dump-likelihood", "-d", "1"], stderr=subprocess.DEVNULL)
    d = refdump.load(dump)
    lk = np.array(d["iters"][0]["lk"])                       # [N][3]: haplotype 0, haplotype 1, genotype
    assert lk.shape == (d["N"], 3) and (lk[:, :2] > 0).all()
    o = Oracle(d["geno"])
    o.find_patterns()
    o.resolve_all()
    haps = np.stack([o.resolution(i) for i in range(d["N"])])      # the reference's resolutions (bit-identical, checked above)
    mine = o.likelihood(haps.reshape(-1, d["L"])).reshape(-1, 2)
    assert np.array_equal(mine, lk[:, :2])
    assert np.array_equal(mine[:, 0] * mine[:, 1], lk[:, 2])        # the Genotype overload is the product of the two
