"""-m gpu: K3 backward + K4 E-step counts through the C-ABI against the reference dumps and the oracle port."""
import numpy as np
import pytest

from hmc_b200 import DeviceBuilder, FLAG_SCALED
from hmc_b200.synth import make_panel
from oracle.pyoracle import Oracle
from util import golden, model_opts, rel_close

pytestmark = pytest.mark.gpu


def setup(d_geno, o, threads=0):
    dev = DeviceBuilder(threads=threads)
    nall, afreq = o.allele_table()
    dev.set_panel(d_geno, nall, afreq)
    m = o.export_model()
    dev.set_model(m)
    return dev, m


@pytest.mark.parametrize("threads", [32, 128])
@pytest.mark.parametrize("name", ["g1_clean", "g2_missing", "g6_mc", "g8_head2"])
def test_estep_vs_reference_dump(name, threads):
    d = golden(name)
    it = d["iters"][0]
    o = Oracle(d["geno"], **model_opts(name))
    o.find_patterns()
    dev, m = setup(d["geno"], o, threads)
    r = dev.resolve_all(K=o.opts["sample_size"], haps=False)
    assert np.array_equal(r["pg"], it["pg"])
    f, pf = dev.estimate_counts(m["start"], m["al_off"], m["al"], r["pg"])
    f, pf, tp = dev.normalise_counts(f, pf, d["N"])
    e = it["estep"]
    # the reference sums over std::map<HaploPair*,...> in heap-address order: 1e-9 relative is the bar
    assert rel_close(f, e["freq"]) and rel_close(pf, e["prefix"]) and rel_close(tp, e["tp"])
    # backward likelihoods are summed in forward-link order like the reference: bit-exact
    for i in range(d["N"]):
        b = it["inds"][i]
        if not b["states"] or not next(iter(b["states"].values()))["beta"]:
            continue
        tr = dev.trellis(i, beta=True)
        off = 0
        for l in range(d["L"] + 1):
            c = int(tr["counts"][l])
            s = b["states"].get(l)
            if s is not None:
                assert c == len(s["ida"])
                assert np.array_equal(tr["alpha"][off:off + c], s["alpha"])
                assert np.array_equal(tr["beta"][off:off + c], s["beta"]), (i, l)
            off += c
    dev.close()


def test_estep_candidates_outside_the_automaton():
    """estimatePatterns' first E-pass: current patterns + unseen one-allele extensions (PatternManager.cpp:373-389)."""
    g = make_panel(48, 40, founders=6, missing=0.03, seed=21)
    o = Oracle(g)
    o.find_patterns()
    o.resolve_all()
    dev, m = setup(g, o)
    r = dev.resolve_all(K=10, haps=False)
    # candidates: every pattern plus its extensions by each allele when the successor does not keep the start
    starts, als = [], []
    nall, _ = o.allele_table()
    for p in range(m["P"]):
        a = m["al"][m["al_off"][p]:m["al_off"][p + 1]]
        starts.append(m["start"][p]); als.append(a)
        if m["end"][p] < g.shape[2] and len(a) < 30:
            for j in range(nall[m["end"][p]]):
                s = m["succ"][p, j]
                if s < 0 or m["start"][s] != m["start"][p]:
                    starts.append(m["start"][p]); als.append(np.append(a, j).astype(np.int8))
    off = np.zeros(len(als) + 1, np.int64)
    np.cumsum([len(a) for a in als], out=off[1:])
    al = np.concatenate(als).astype(np.int8)
    starts = np.array(starts, np.int32)
    ef, epf, etp = o.estep(starts, off, al)
    f, pf = dev.estimate_counts(starts, off, al, r["pg"])
    f, pf, tp = dev.normalise_counts(f, pf, 48)
    assert rel_close(f, ef) and rel_close(pf, epf) and rel_close(tp, etp)
    dev.close()


def test_estep_sharded_counts_add_up():
    """Individuals shard: raw counts of [0,n) == counts of [0,k) + counts of [k,n) (the all-reduce identity)."""
    g = make_panel(60, 50, founders=8, missing=0.04, seed=22)
    o = Oracle(g)
    o.find_patterns()
    dev, m = setup(g, o)
    r = dev.resolve_all(K=1, haps=False)
    f, pf = dev.estimate_counts(m["start"], m["al_off"], m["al"], r["pg"])
    fa, pfa = dev.estimate_counts(m["start"], m["al_off"], m["al"], r["pg"][:23], first=0, count=23)
    fb, pfb = dev.estimate_counts(m["start"], m["al_off"], m["al"], r["pg"][23:], first=23, count=37)
    assert rel_close(f, fa + fb, 1e-12) and rel_close(pf, pfa + pfb, 1e-12)
    dev.close()


def test_estep_scaled_matches_unscaled():
    """L = 900: alpha dips below 2^-512 for most individuals; the exponent-carrying E-step must equal the raw fp64 one (still finite here)."""
    g = make_panel(16, 900, founders=6, seed=23)
    o = Oracle(g, max_pattern_len=10)
    o.find_patterns()
    ll = o.resolve_all()
    assert np.isfinite(ll)
    dev, m = setup(g, o)
    r0 = dev.resolve_all(K=1, haps=False)
    f0, pf0 = dev.estimate_counts(m["start"], m["al_off"], m["al"], r0["pg"])
    r1 = dev.resolve_all(K=1, haps=False, flags=FLAG_SCALED)
    assert r1["exp"].max() > 0
    f1, pf1 = dev.estimate_counts(m["start"], m["al_off"], m["al"], r1["pg"], r1["exp"], flags=FLAG_SCALED)
    assert rel_close(f0, f1) and rel_close(pf0, pf1)
    # a second scaled pass reuses the trellis together with its forward / backward exponent tables
    f2, pf2 = dev.estimate_counts(m["start"], m["al_off"], m["al"], r1["pg"], r1["exp"], flags=FLAG_SCALED)
    assert dev.kernel_ms()["forward"] == 0 and rel_close(f1, f2) and rel_close(pf1, pf2)
    # ... but not across a change of the scaling flag
    f3, pf3 = dev.estimate_counts(m["start"], m["al_off"], m["al"], r0["pg"])
    assert dev.kernel_ms()["forward"] > 0 and rel_close(f0, f3)
    dev.close()


def test_later_epasses_reuse_the_trellis(monkeypatch):
    """The E-passes of one M-step share model and individuals: after the first pass the E-step trellis (alpha, beta, links) stays
    in the arena and the next hmc_estimate_counts calls go straight to K4.  Same counts as a from-scratch pass; the reuse ends
    with any call that changes the model, the panel or the arena."""
    g = make_panel(60, 50, founders=6, missing=0.04, seed=61)
    o = Oracle(g)
    o.find_patterns()
    o.resolve_all()
    dev, m = setup(g, o)
    r = dev.resolve_all(K=1, haps=False)
    P = m["P"]
    # pass 1: all patterns; pass 2: a different candidate list (every other pattern)
    f1, pf1 = dev.estimate_counts(m["start"], m["al_off"], m["al"], r["pg"])
    k1, c1 = dev.kernel_ms(), dev.counters()
    assert k1["forward"] > 0 and k1["backward"] > 0 and c1["estep_entries"] > 0
    keep = np.arange(0, P, 2)
    lens = (m["al_off"][1:] - m["al_off"][:-1])[keep]
    off2 = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    al2 = np.concatenate([m["al"][m["al_off"][i]:m["al_off"][i + 1]] for i in keep]).astype(np.int8)
    st2 = np.ascontiguousarray(m["start"][keep])
    f2, pf2 = dev.estimate_counts(st2, off2, al2, r["pg"])
    k2, c2 = dev.kernel_ms(), dev.counters()
    assert k2["forward"] == 0 and k2["backward"] == 0 and k2["estep"] > 0          # reused
    assert c2["steps"] == c1["steps"] and c2["states"] == c1["states"] and 0 < c2["estep_entries"] <= c1["estep_entries"]
    monkeypatch.setenv("HMC_B200_NO_ESTEP_CACHE", "1")
    f3, pf3 = dev.estimate_counts(st2, off2, al2, r["pg"])
    assert dev.kernel_ms()["forward"] > 0
    assert rel_close(f2, f3) and rel_close(pf2, pf3)
    monkeypatch.delenv("HMC_B200_NO_ESTEP_CACHE")
    # a pattern's own count does not depend on which other candidates are in the list
    assert rel_close(f2, f1[keep])
    # the oracle agrees
    ef, epf, _ = o.estep(st2, off2, al2)
    fn, pfn, _ = dev.normalise_counts(f2, pf2, 60)
    assert rel_close(fn, ef) and rel_close(pfn, epf)
    # model upload ends the reuse
    dev.set_model(m)
    dev.estimate_counts(st2, off2, al2, r["pg"])
    assert dev.kernel_ms()["forward"] > 0
    # and so does a resolve pass
    dev.estimate_counts(st2, off2, al2, r["pg"])
    assert dev.kernel_ms()["forward"] == 0
    dev.resolve_all(K=1, haps=False)
    dev.estimate_counts(st2, off2, al2, r["pg"])
    assert dev.kernel_ms()["forward"] > 0
    dev.close()
