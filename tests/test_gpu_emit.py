"""-m gpu: K8 -- hmc_get_likelihood (HaploBuilder::getLikelihood, HaploBuilder.cpp:128-151) and hmc_format_best (the allele lines of
`.reconstructed`, HaploFile.cpp:146-148 / Allele.cpp:82-101) against the oracle and the host formatter."""
import os

import numpy as np
import pytest

from hmc_b200 import DeviceBuilder, HaploModel, HmcError
from hmc_b200.synth import make_panel, write_phase
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


def setup(N, L, **kw):
    g = make_panel(N, L, **kw)
    o = Oracle(g, sample_size=5)
    o.find_patterns()
    nall, afreq = o.allele_table()
    dev = DeviceBuilder()
    dev.set_panel(g, nall, afreq)
    dev.set_model(o.export_model())
    return g, o, dev


@pytest.mark.parametrize("kw", [dict(founders=6, seed=11), dict(alleles=5, founders=5, missing=0.05, seed=12), dict(founders=8, missing=0.1, seed=13)])
def test_likelihood_equals_the_trie_descent(kw):
    """One automaton walk per haplotype on the device = one backward-trie descent per index in the reference, bit for bit: the best
    resolutions (on the device), and random haplotypes."""
    g, o, dev = setup(60, 80, **kw)
    r = dev.resolve_all(K=5)
    o.resolve_all()
    best = r["haps"][:, 0]                                          # [N][2][L]
    ok = r["nres"] > 0
    want = o.likelihood(best[ok].reshape(-1, 80)).reshape(-1, 2)
    got_dev = dev.likelihood()                                      # where the results lie
    got = dev.likelihood(best[ok])
    assert np.array_equal(got, want) and np.array_equal(got_dev[ok], want)
    # getLikelihood(Genotype) = product of the two; the top-1 prior is that product (x 2 for a heterozygous pair) up to the order of the multiplications
    het = (best[ok, 0] != best[ok, 1]).any(axis=1)
    prod = want[:, 0] * want[:, 1] * np.where(het, 2.0, 1.0)
    assert np.allclose(prod, r["prior"][ok, 0], rtol=1e-12, atol=0)
    rng = np.random.default_rng(5)
    nall = o.allele_table()[0]
    rnd = np.stack([rng.integers(0, nall[l], 300) for l in range(80)], axis=1).astype(np.int8)
    a, b = dev.likelihood(rnd), o.likelihood(rnd)
    assert np.array_equal(a, b)


def test_likelihood_rejects_missing_alleles():
    g, o, dev = setup(30, 40, founders=4, seed=3)
    h = np.zeros((2, 40), np.int8)
    h[1, 7] = -1
    with pytest.raises(HmcError, match="missing or out-of-range"):
        dev.likelihood(h)
    h[1, 7] = 9
    with pytest.raises(HmcError, match="missing or out-of-range"):
        dev.likelihood(h)


@pytest.mark.parametrize("kind,kw", [("S", dict(founders=6, missing=0.05, seed=21)), ("M", dict(alleles=6, founders=7, missing=0.03, seed=22))])
def test_reconstructed_text_from_the_device_equals_the_host_formatter(kind, kw, tmp_path, monkeypatch):
    g = make_panel(57, 133, **kw)
    phase = str(tmp_path / "p.phase")
    write_phase(phase, g, kind=kind)
    hm = HaploModel(max_iteration=2)
    hm.load_phase(phase)
    hm.run()
    hm.write_reconstructed(phase, ".dev")
    monkeypatch.setenv("HMC_B200_HOST_FORMAT", "1")
    hm.write_reconstructed(phase, ".host")
    monkeypatch.delenv("HMC_B200_HOST_FORMAT")
    a, b = open(phase + ".dev", "rb").read(), open(phase + ".host", "rb").read()
    assert a == b and len(a) > 57 * 2 * 133
    hm.close()


def test_format_best_writes_the_input_genotype_of_an_unresolved_individual():
    """HaploBuilder.cpp:117-124: an individual no head pattern matches keeps its input genotype, missing alleles included ('?' / -1)."""
    g = make_panel(40, 30, founders=4, seed=31)
    o = Oracle(g, sample_size=3)
    o.find_patterns()
    nall, afreq = o.allele_table()
    g2 = g.copy()
    dev = DeviceBuilder()
    m = o.export_model()
    # a model that lost every pattern ending in allele 0 of locus 14: the states of an individual carrying it there die out (no successor)
    keep = ~((m["end"] == 15) & (m["al"][m["al_off"][1:] - 1] == 0))
    al = np.concatenate([m["al"][m["al_off"][i]:m["al_off"][i + 1]] for i in np.nonzero(keep)[0]])
    ln = (m["al_off"][1:] - m["al_off"][:-1])[keep]
    m2 = dict(P=int(keep.sum()), head_len=m["head_len"], start=m["start"][keep].copy(), end=m["end"][keep].copy(), freq=m["freq"][keep].copy(),
              prefix=m["prefix"][keep].copy(), tp=m["tp"][keep].copy(), al_off=np.concatenate([[0], np.cumsum(ln)]).astype(np.int64), al=al.astype(np.int8))
    g2[7, 1, 3] = -1
    o2 = Oracle(g2, sample_size=3)
    o2.set_allele_table(nall, afreq)
    o2.import_model(m2)                                            # successors and heads of the cut-down model
    dev.set_panel(g2, nall, afreq)
    dev.set_model(o2.export_model())
    r = dev.resolve_all(K=3)
    assert (r["nres"] == 0).any() and (r["nres"] > 0).any()
    # no pattern ends at index 15 with allele 0 -> likelihood 0 (HaploBuilder.cpp:141-144)
    hz = np.zeros((1, 30), np.int8)
    assert dev.likelihood(hz)[0] == 0.0
    sym = np.zeros((30, int(nall.max())), np.int32)
    sym[:, 0] = ord("A"); sym[:, 1] = ord("G")
    text, off = dev.format_best(sym, "S" * 30)
    lines = text.decode().split("\n")[:-1]
    assert len(lines) == 80 and off[-1] == len(text)
    for i in range(40):
        src = r["haps"][i, 0] if r["nres"][i] > 0 else g2[i]
        for h in range(2):
            want = "".join(("?" if k < 0 else "AG"[k]) + " " for k in src[h])
            assert lines[2 * i + h] == want, (i, h)
    with pytest.raises(HmcError, match="the buffer holds"):
        import ctypes as C
        from hmc_b200.device import lib
        buf = np.zeros(10, np.uint8); tot = C.c_int64()
        dev._ck(lib().hmc_format_best(dev._h, sym.ctypes.data_as(C.c_void_p), b"S" * 30, buf.ctypes.data_as(C.c_void_p), 10, None, C.byref(tot)))
    dev.close()
