"""-m gpu: size-independent properties at BASELINE config 1's FULL size (C2: 2,000 x 1,000, 5 % missing, default options), where the
oracle would take hours: consistency with the genotypes, ordering / normalisation of the K-lists, K-prefix stability,
sum_s alpha(s) beta(s) = P(G) at every trellis index, shard invariance, and the E-step mass identities."""
import numpy as np
import pytest

from hmc_b200 import HaploModel
from hmc_b200.synth import make_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    g, kind = make_config("C2")
    hm = HaploModel()
    hm.set_panel(g, kind)
    hm.build()
    hm.upload()
    m = hm.export_model()
    dev = hm.device()
    r10 = dev.resolve_all(K=10)
    yield dict(g=g, hm=hm, m=m, dev=dev, r10=r10)
    hm.close()


def test_resolutions_are_consistent_with_the_genotypes(c2):
    g, r = c2["g"], c2["r10"]
    assert (r["nres"] > 0).all() and np.isfinite(r["ll"])
    N, _, L = g.shape
    typed = (g[:, 0] >= 0) & (g[:, 1] >= 0)
    for k in range(3):                                  # the three best phasings of every individual
        h = r["haps"][:, k]
        valid = (r["nres"] > k)[:, None] & typed
        lo, hi = np.minimum(h[:, 0], h[:, 1]), np.maximum(h[:, 0], h[:, 1])
        glo, ghi = np.minimum(g[:, 0], g[:, 1]), np.maximum(g[:, 0], g[:, 1])
        assert ((lo == glo) & (hi == ghi))[valid].all()  # unordered allele pair of every typed locus is preserved
        assert (h[r["nres"] > k] >= 0).all()            # missing genotypes are imputed


def test_klists_are_sorted_normalised_and_prefix_stable(c2):
    r, dev = c2["r10"], c2["dev"]
    for i in range(0, 2000, 7):
        n = r["nres"][i]
        p = r["prior"][i, :n]
        assert (np.diff(p) <= 0).all() and (p >= 0).all()
        assert np.array_equal(r["posterior"][i, :n], p / r["pg"][i])
        assert 0 < r["coverage"][i] <= 1.0 + 1e-12
    r1 = dev.resolve_all(K=1)
    r4 = dev.resolve_all(K=4)
    assert np.array_equal(r1["pg"], r["pg"]) and r1["ll"] == r["ll"]             # P(G) does not depend on K
    assert np.array_equal(r1["prior"][:, 0], r["prior"][:, 0])                   # best path likelihood is K-independent, bit for bit
    assert np.array_equal(r4["prior"], r["prior"][:, :4])


def test_streamed_kbest_equals_the_serial_order(c2, monkeypatch):
    """K2 runs beside the forward kernel and takes individuals as they finish (done list, own arena region); running it after the
    forward kernel instead (HMC_B200_KBEST_SERIAL=1) must give the same bits -- full C2, every array."""
    r, dev = c2["r10"], c2["dev"]
    monkeypatch.setenv("HMC_B200_KBEST_SERIAL", "1")
    s = dev.resolve_all(K=10)
    monkeypatch.delenv("HMC_B200_KBEST_SERIAL")
    for k in ("pg", "exp", "coverage", "nres", "prior", "posterior", "haps"):
        assert np.array_equal(s[k], r[k]), k
    again = dev.resolve_all(K=10)                       # and the streamed pass is reproducible run to run
    for k in ("pg", "prior", "haps"):
        assert np.array_equal(again[k], r[k]), k


def test_shard_invariance(c2):
    r, dev = c2["r10"], c2["dev"]
    a = dev.resolve_all(K=10, first=0, count=700)
    b = dev.resolve_all(K=10, first=700, count=1300)
    assert np.array_equal(np.concatenate([a["pg"], b["pg"]]), r["pg"])
    assert np.array_equal(np.concatenate([a["prior"], b["prior"]]), r["prior"])
    assert np.array_equal(np.concatenate([a["haps"], b["haps"]]), r["haps"])


def test_forward_backward_identity_and_estep_mass(c2):
    m, dev, r, g = c2["m"], c2["dev"], c2["r10"], c2["g"]
    n = 48
    f, pf = dev.estimate_counts(m["start"], m["al_off"], m["al"], r["pg"][:n], first=0, count=n, flags=2)
    # sum_s alpha(s) * beta(s) = P(G) at every index with live states (any individual of the kept wave)
    for i in (0, 17, 47):
        tr = dev.trellis(i, capacity=1 << 24, beta=True)
        off = 0
        for l, c in enumerate(tr["counts"]):
            if c:
                s = float(np.dot(tr["alpha"][off:off + c], tr["beta"][off:off + c]))
                assert abs(s - r["pg"][i]) <= 1e-9 * r["pg"][i], (i, l)
            off += c
    L = g.shape[2]
    single = (m["end"] - m["start"]) == 1
    # length-1 patterns: prefix sum = number of individuals (last_freq = 1 at depth 1), and the alleles of a locus share the mass 1 per individual
    assert np.allclose(pf[single], n, rtol=1e-12)
    mass = np.zeros(L)
    np.add.at(mass, m["start"][single], f[single])
    assert np.allclose(mass, n, rtol=1e-9)
    # a pattern can never be more frequent than its prefix
    assert (f <= pf * (1 + 1e-9) + 1e-12).all()


def test_device_mined_model_equals_host_mined_model_at_full_size(c2):
    """K5 + K6 at C2's full size (4.7 M patterns, 9.4 M trie nodes): the model the fixture built on the device must be the one
    the threaded host miner + host automaton build produce, bit for bit (ids, strings, frequencies, successors, heads)."""
    assert c2["hm"].mined_on_device()
    hm0 = HaploModel(device_mining=0)
    hm0.set_panel(c2["g"], "S")
    hm0.build()
    assert not hm0.mined_on_device()
    a, b = c2["m"], hm0.export_model()
    hm0.close()
    assert a["P"] == b["P"] and a["head_len"] == b["head_len"]
    for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads", "al", "al_off"):
        assert np.array_equal(a[k], b[k]), k


def test_switch_error_of_the_c2_run_is_small_and_symmetric(c2):
    """K7 at full size: the panel holds the true phase, so HaploComp of the best resolutions against it is the phasing error of
    the run; comparing where the resolutions lie on the device equals comparing a host copy, and swapping the two haplotype
    rows of every inferred individual changes nothing (switch distance and genotype difference are orientation-free)."""
    g, dev, r = c2["g"], c2["dev"], c2["r10"]
    best = np.ascontiguousarray(r["haps"][:, 0])
    a = dev.compare(g, best)
    b = dev.compare(g, np.ascontiguousarray(best[:, ::-1]))
    assert np.array_equal(a["counts"], b["counts"])
    assert 0.0 <= a["switch_error"] < 0.2 and a["counts"][1] > 0
