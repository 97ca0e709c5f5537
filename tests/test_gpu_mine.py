"""K5 (device pattern mining, csrc/mine.cuh) against the reference's dumps and the host miner: pattern ids, allele
strings, frequency / prefix_freq / transition_prob must be bit-identical (PatternManager.cpp:100-291)."""
import os

import numpy as np
import pytest

from cases import CASES
from hmc_b200 import DeviceBuilder
from hmc_b200.host import HaploModel
from hmc_b200.synth import make_panel
from util import GOLDEN, golden, model_opts

pytestmark = pytest.mark.gpu

KEYS = ("start", "end", "freq", "prefix", "tp", "succ", "heads", "al")


@pytest.fixture(params=[None, "96"], ids=["one-batch", "start-batches"])
def mine_batches(request, monkeypatch):
    """The start loci are mined in batches of bounded (item, start) pairs; "96" forces batches of 1-3 start loci on the small panels."""
    if request.param:
        monkeypatch.setenv("HMC_B200_MINE_PAIRS", request.param)
    return request.param


@pytest.mark.parametrize("name", [n for n in CASES if n != "g7_bynum"])
def test_first_model_matches_reference_dump(name, mine_batches):
    """Genotype branch (checkFrequencyWithExtension :205-251 / getMatchingFrequency) on every golden panel."""
    d = golden(name)
    hm = HaploModel(**{**model_opts(name), "device_mining": 1})
    hm.load_phase(os.path.join(GOLDEN, name + ".phase"))
    hm.build()
    assert hm.mined_on_device()
    m, r = hm.export_model(), d["iters"][0]["patterns"]
    assert m["P"] == r["P"] and m["head_len"] == r["head_len"]
    for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads"):
        assert np.array_equal(m[k], r[k]), k
    assert np.array_equal(np.concatenate(r["al"]), m["al"])
    hm.close()


def test_by_num_model_stays_on_the_host():
    hm = HaploModel(**{**model_opts("g7_bynum"), "device_mining": 1})
    hm.load_phase(os.path.join(GOLDEN, "g7_bynum.phase"))
    hm.build()
    assert not hm.mined_on_device()
    hm.close()


def _models(g, kind="S", **opts):
    out = []
    for dm in (0, 1):
        hm = HaploModel(**{**opts, "device_mining": dm})
        hm.set_panel(g, kind)
        hm.build()
        assert hm.mined_on_device() == bool(dm)
        out.append(hm.export_model())
        hm.close()
    return out


@pytest.mark.parametrize("cfg", [
    dict(n_ind=300, n_loci=200, founders=16, missing=0.05, seed=2),
    dict(n_ind=64, n_loci=90, founders=4, seed=3),
    dict(n_ind=120, n_loci=60, alleles=6, founders=8, missing=0.02, seed=4),
    dict(n_ind=1, n_loci=5, founders=2, seed=5),
    dict(n_ind=33, n_loci=1, founders=2, seed=6),
])
@pytest.mark.parametrize("opts", [dict(), dict(max_pattern_len=4), dict(min_pattern_len=2, max_pattern_len=12), dict(min_freq=0.05, min_freq_abs=-1.0),
                                  dict(model="MC", mc_order=2)])
def test_device_equals_host_miner(cfg, opts, mine_batches):
    c = dict(cfg)
    kind = "M" if c.get("alleles", 2) > 2 else "S"
    if c["n_loci"] < opts.get("min_pattern_len", 1):
        pytest.skip("panel shorter than the head")
    g = make_panel(c.pop("n_ind"), c.pop("n_loci"), **c)
    host, dev = _models(g, kind, **opts)
    assert host["P"] == dev["P"]
    for k in KEYS:
        assert np.array_equal(host[k], dev[k]), k


@pytest.mark.parametrize("name", ["g4_em", "g2_missing"])
def test_sampling_em_identical_with_device_mining(name, mine_batches):
    """Sampling branch (:180-191,252-263): every EM iteration re-mines on the K sampled haplotypes; the whole trace,
    the pattern counts and the final phasing must not depend on where the search ran."""
    out = []
    for dm in (0, 1):
        hm = HaploModel(**{**model_opts(name), "max_iteration": 5, "device_mining": dm})
        hm.load_phase(os.path.join(GOLDEN, name + ".phase"))
        ll, npat = hm.run()
        out.append((ll, npat, hm.resolutions(), hm.export_model()))
        hm.close()
    (l0, n0, r0, m0), (l1, n1, r1, m1) = out
    assert np.array_equal(l0, l1) and np.array_equal(n0, n1) and np.array_equal(r0, r1)
    for k in KEYS:
        assert np.array_equal(m0[k], m1[k]), k


def test_direct_abi_call_with_samples():
    """hmc_mine_patterns through the C-ABI with an explicit weighted sample, checked against a brute-force count."""
    rng = np.random.default_rng(7)
    n, L = 500, 12
    founders = rng.integers(0, 2, (5, L), dtype=np.int8)
    h = founders[rng.integers(0, 5, n)].copy()
    h[rng.random((n, L)) < 0.01] ^= 1
    h[rng.random((n, L)) < 0.01] = -1                       # unresolved alleles match every pattern
    w = rng.random(n)
    tw = float(np.sum(w))
    g = np.stack([h[:50], h[50:100]], axis=1)               # any panel with the same loci: only the allele tables are used
    nall = np.full(L, 2, np.int32)
    af = np.full((L, 2), 0.5)
    dev = DeviceBuilder()
    dev.set_panel(np.ascontiguousarray(g), nall, af)
    r = dev.mine_patterns(0.02, 1, 6, samples=h, weights=w, total_weight=tw)
    assert r["P"] > 0
    for i in range(r["P"]):
        s, e = r["start"][i], r["end"][i]
        al = r["alleles"][r["al_off"][i]:r["al_off"][i + 1]]
        match = np.all((h[:, s:e] == al) | (h[:, s:e] < 0), axis=1)
        tot = 0.0
        for x in w[match]:
            tot += x                                         # the reference's sequential sum
        assert r["freq"][i] == tot / tw
        assert e - s == 1 or r["freq"][i] >= 0.02
    st = dev.mine_stats()
    assert st["nodes"] >= r["P"] and st["levels"] >= 2
    dev.close()


def test_device_resident_samples_equal_host_samples(monkeypatch):
    """The sampled haplotypes normally stay on the device between the resolve pass and the next search
    (hmc_mine_patterns_resolved); HMC_B200_HOST_SAMPLES=1 moves them through HaploData on the host as the reference does."""
    out = []
    for host_samples in (False, True):
        if host_samples:
            monkeypatch.setenv("HMC_B200_HOST_SAMPLES", "1")
        hm = HaploModel(**{**model_opts("g4_em"), "max_iteration": 4, "device_mining": 1})
        hm.load_phase(os.path.join(GOLDEN, "g4_em.phase"))
        ll, npat = hm.run()
        out.append((ll, npat, hm.resolutions(), hm.export_model()))
        hm.close()
    (l0, n0, r0, m0), (l1, n1, r1, m1) = out
    assert len(l0) > 1 and np.array_equal(l0, l1) and np.array_equal(n0, n1) and np.array_equal(r0, r1)
    for k in KEYS:
        assert np.array_equal(m0[k], m1[k]), k
