"""The oracle port (oracle/hmc_oracle.hpp) against dumps of the UNMODIFIED reference binary
(tests/golden/*.dump.gz, written by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from cases import CASES
from oracle.pyoracle import Oracle
from util import GOLDEN, golden, model_opts, read_phase_idx, rel_close, same_unordered_pair

SINGLE = ["g1_clean", "g2_missing", "g3_multi", "g7_bynum", "g8_head2", "g9_k1"]


@pytest.mark.parametrize("name", list(CASES))
def test_allele_table_and_first_model(name):
    d = golden(name)
    o = Oracle(d["geno"], **model_opts(name))
    nall, afreq = o.allele_table()
    assert np.array_equal(nall, d["nall"]) and np.array_equal(afreq, d["afreq"])
    o.find_patterns()
    m, r = o.export_model(), d["iters"][0]["patterns"]
    assert m["P"] == r["P"] and m["head_len"] == r["head_len"]
    for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads"):
        assert np.array_equal(m[k], r[k]), k
    assert np.array_equal(np.concatenate(r["al"]), m["al"])


@pytest.mark.parametrize("name", SINGLE)
def test_resolve_all_bit_exact(name):
    d = golden(name)
    it = d["iters"][0]
    o = Oracle(d["geno"], **model_opts(name))
    o.find_patterns()
    ll = o.resolve_all()
    assert ll == it["ll"] or (np.isinf(ll) and np.isinf(it["ll"]))
    if "counters" in it:
        assert o.counters() == it["counters"]
    for i in range(d["N"]):
        a = o.individual(i)
        assert a["pg"] == it["pg"][i]
        if it["inds"]:
            b = it["inds"][i]
            assert a["n"] == b["n"]
            assert np.array_equal(a["prior"], np.array(b["prior"])) and np.array_equal(a["posterior"], np.array(b["posterior"]))
            for k in range(a["n"]):
                assert np.array_equal(a["haps"][k], b["haps"][k]), (i, k)


@pytest.mark.parametrize("name", ["g1_clean", "g2_missing", "g3_multi", "g8_head2"])
def test_trellis_bit_exact(name):
    d = golden(name)
    it = d["iters"][0]
    o = Oracle(d["geno"], **model_opts(name))
    o.find_patterns()
    o.resolve_all()
    for i in range(d["N"]):
        b = it["inds"][i]
        with_beta = len(next(iter(b["states"].values()))["beta"]) > 0 if b["states"] else False
        tr = o.trellis(i, K=o.opts["sample_size"], with_beta=with_beta)
        off = 0
        for l in range(d["L"] + 1):
            c = int(tr["counts"][l])
            s = b["states"].get(l)
            if s is None:
                assert c == 0 or l < o.opts["min_pattern_len"]
            else:
                assert c == len(s["ida"]), (i, l)
                sl = slice(off, off + c)
                assert np.array_equal(tr["ida"][sl], s["ida"]) and np.array_equal(tr["idb"][sl], s["idb"])
                assert np.array_equal(tr["alpha"][sl], s["alpha"]) and np.array_equal(tr["tp"][sl], s["tp"])
                if with_beta:
                    assert np.array_equal(tr["beta"][sl], s["beta"]), (i, l)   # sums run in link order, like the reference
            off += c


@pytest.mark.parametrize("name", ["g1_clean", "g2_missing", "g6_mc", "g8_head2"])
def test_estep_counts(name):
    d = golden(name)
    it = d["iters"][0]
    o = Oracle(d["geno"], **model_opts(name))
    o.find_patterns()
    o.resolve_all()
    m = o.export_model()
    f, pf, tp = o.estep(m["start"], m["al_off"], m["al"])
    e = it["estep"]
    # summation order over std::map<HaploPair*,...> follows heap addresses in the reference
    assert rel_close(f, e["freq"]) and rel_close(pf, e["prefix"]) and rel_close(tp, e["tp"])


@pytest.mark.parametrize("name", ["g4_em", "g5_exact", "g6_mc"])
def test_em_trace(name):
    d = golden(name)
    o = Oracle(d["geno"], **model_opts(name))
    ll, npat = o.run()
    ref_ll = np.array([it["ll"] for it in d["iters"]])
    ref_np = np.array([it["P"] for it in d["iters"]])
    assert len(ll) == len(ref_ll)
    assert np.array_equal(npat, ref_np)
    if model_opts(name).get("exact_estimate"):
        assert rel_close(ll, ref_ll)
    else:
        assert np.array_equal(ll, ref_ll)       # sampling EM is deterministic to the bit
    rec = read_phase_idx(os.path.join(GOLDEN, name + ".phase.reconstructed"), d["sym"])
    for i in range(d["N"]):
        assert same_unordered_pair(o.resolution(i), rec[i]), i


@pytest.mark.parametrize("name", ["g1_clean", "g2_missing", "g3_multi", "g8_head2", "g9_k1"])
def test_get_likelihood_golden(name):
    """getLikelihood(Haplotype) / (Genotype) of the reference on its own reconstructions (tests/golden/likelihood.json, written by
    tests/golden/make_likelihood.py from the unmodified reference) == the oracle's restatement, bit for bit."""
    import json
    d = golden(name)
    want = np.array([[float.fromhex(x) for x in row] for row in json.load(open(os.path.join(GOLDEN, "likelihood.json")))[name]])
    rec = read_phase_idx(os.path.join(GOLDEN, name + ".phase.reconstructed"), d["sym"])        # what the reference resolved
    o = Oracle(d["geno"], **model_opts(name))
    o.find_patterns()
    got = o.likelihood(rec.reshape(-1, d["L"])).reshape(-1, 2)
    assert np.array_equal(got, want[:, :2])
    assert np.array_equal(got[:, 0] * got[:, 1], want[:, 2])
