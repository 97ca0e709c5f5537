"""-m gpu: edge cases of the C-ABI path (tiny / degenerate panels, limits, misuse), each against the oracle port."""
import numpy as np
import pytest

from hmc_b200 import DeviceBuilder, HmcError, FLAG_SCALED
from hmc_b200.synth import make_panel
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


def both(g, K=10, threads=0, flags=0, **opts):
    o = Oracle(g, sample_size=K, scaled=bool(flags & FLAG_SCALED), **opts)
    o.find_patterns()
    ll = o.resolve_all()
    dev = DeviceBuilder(threads=threads)
    nall, afreq = o.allele_table()
    dev.set_panel(g, nall, afreq)
    dev.set_model(o.export_model())
    r = dev.resolve_all(K=K, flags=flags)
    return o, ll, dev, r


def check(o, ll, r, n):
    assert (np.isneginf(ll) and np.isneginf(r["ll"])) or ll == r["ll"]
    for i in range(n):
        a = o.individual(i)
        assert a["pg"] == r["pg"][i] and a["n"] == r["nres"][i]
        assert np.array_equal(a["prior"], r["prior"][i, :a["n"]])


@pytest.mark.parametrize("threads", [32, 64, 128])
def test_single_individual_two_loci(threads):
    g = np.array([[[0, 1], [1, 0]]], np.int8)
    o, ll, dev, r = both(g, threads=threads)
    check(o, ll, r, 1)
    dev.close()


@pytest.mark.parametrize("threads", [32, 64, 128])
def test_one_locus_panel_has_no_steps(threads):
    g = make_panel(12, 1, founders=2, seed=31)
    g[0, 0, 0], g[0, 1, 0] = 0, 1       # make sure both alleles occur
    o, ll, dev, r = both(g, threads=threads)
    check(o, ll, r, 12)
    assert dev.counters()["steps"] == 0
    dev.close()


@pytest.mark.parametrize("threads", [32, 64, 128])
def test_fully_missing_individual(threads):
    g = make_panel(20, 16, founders=4, seed=32)
    g[3] = -1                            # every genotype of individual 3 is missing: all allele pairs at every locus
    g[7, :, :8] = -1
    o, ll, dev, r = both(g, threads=threads, max_pattern_len=6)
    check(o, ll, r, 20)
    assert (r["haps"][3, 0] >= 0).all()
    dev.close()


def test_k_extremes():
    g = make_panel(16, 30, founders=4, missing=0.05, seed=33)
    for K in (1, 2, 16, 32):
        o, ll, dev, r = both(g, K=K)
        check(o, ll, r, 16)
        dev.close()


def test_sixteen_alleles_and_the_limit():
    g = make_panel(40, 10, alleles=16, founders=24, seed=34)
    o, ll, dev, r = both(g, max_pattern_len=4)
    check(o, ll, r, 40)
    dev.close()
    dev = DeviceBuilder()
    with pytest.raises(HmcError):
        dev.set_panel(np.zeros((2, 2, 3), np.int8), np.full(3, 17, np.int32), np.zeros((3, 17)))
    dev.close()


def test_underflow_without_scaling_is_the_references_zero():
    """L = 3000 unscaled: alpha underflows, every extension dies (HaploBuilder.cpp:237), P(G) = 0, no resolution -- like the reference."""
    g = make_panel(6, 3000, founders=6, seed=35)
    o, ll, dev, r = both(g, max_pattern_len=8)
    assert np.isneginf(ll) and np.isneginf(r["ll"])
    assert (r["nres"] == 0).all() and (r["pg"] == 0).all()
    dev.close()
    o, ll, dev, r = both(g, flags=FLAG_SCALED, max_pattern_len=8)
    assert np.isfinite(ll) and ll == r["ll"] and (r["nres"] > 0).all()
    dev.close()


def test_misuse_is_an_error_code_not_a_crash():
    dev = DeviceBuilder()
    with pytest.raises(HmcError):
        dev.resolve_all(K=1, count=1)               # no panel
    g = make_panel(8, 10, founders=2, seed=36)
    o = Oracle(g)
    o.find_patterns()
    nall, afreq = o.allele_table()
    dev.set_panel(g, nall, afreq)
    with pytest.raises(HmcError):
        dev.resolve_all(K=1)                        # no model
    m = o.export_model()
    dev.set_model(m)
    with pytest.raises(HmcError):
        dev.resolve_all(K=33)                       # K limit
    with pytest.raises(HmcError):
        dev.resolve_all(K=1, first=4, count=8)      # range outside the panel
    bad = dict(m)
    bad["succ"] = m["succ"].copy()
    bad["succ"][0, 0] = m["P"] + 5
    with pytest.raises(HmcError):
        dev.set_model(bad)
    with pytest.raises(HmcError):
        dev.resolve_all(K=1)                        # a rejected model is not left installed
    bad = dict(m)
    bad["end"] = m["end"].copy()
    bad["end"][m["P"] // 2] += 1                    # start/end no longer agree with the allele string
    with pytest.raises(HmcError, match="inconsistent start/end/alleles"):
        dev.set_model(bad)
    dev.set_model(m)
    r = dev.resolve_all(K=3, first=2, count=4)      # a sub-range works and is indexed from `first`
    o.set_options(sample_size=3)
    o.resolve_all()
    for i in range(4):
        assert o.individual(2 + i)["pg"] == r["pg"][i]
    dev.close()


def test_packed_panel_ingest_and_packed_emission(tmp_path):
    """A binary (bit-packed) panel goes to the device as it is (hmc_set_panel_packed) and the whole run -- device mining,
    automaton, resolve -- equals the run from the text panel; `hmc` writes the resolutions of a binary input in binary, and
    hmc_fetch_best_packed packs the best resolutions on the device exactly like the host packs them."""
    import ctypes as C
    import os
    import subprocess
    from hmc_b200.build import CLI
    from hmc_b200.device import lib
    from hmc_b200.host import HaploModel
    from hmc_b200.synth import write_phase
    g = make_panel(70, 90, alleles=5, founders=6, missing=0.04, seed=41)
    phase, binp = str(tmp_path / "p.phase"), str(tmp_path / "p.hmcb")
    write_phase(phase, g, kind="M")
    subprocess.run([CLI, "--to-binary", binp, phase, "-d", "0"], check=True)
    out = []
    for path in (phase, binp):
        hm = HaploModel(max_iteration=3, device_mining=1)
        hm.load_phase(path)
        ll, npat = hm.run()
        out.append((ll, npat, hm.resolutions()))
        hm.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    # CLI on the binary input: binary output that reads back to the same resolutions as the text run's text output
    subprocess.run([CLI, phase, "-d", "0", "-i", "3"], check=True)
    subprocess.run([CLI, binp, "-d", "0", "-i", "3"], check=True)
    a, b = HaploModel(), HaploModel()
    a.load_phase(phase + ".reconstructed")
    b.load_phase(binp + ".reconstructed")
    with open(binp + ".reconstructed", "rb") as f:
        assert f.read(4) == b"HMCB"
    assert np.array_equal(a.panel()[0], b.panel()[0])
    a.close(); b.close()
    # device-side packing of the best resolutions
    o = Oracle(g)
    o.find_patterns()
    nall, afreq = o.allele_table()
    dev = DeviceBuilder()
    dev.set_panel(g, nall, afreq)
    dev.set_model(o.export_model())
    r = dev.resolve_all(K=4)
    bits = 4
    per = 32 // (2 * bits)
    words = (90 + per - 1) // per
    packed = np.zeros((70, words), np.uint32)
    rc = lib().hmc_fetch_best_packed(dev._h, packed.ctypes.data_as(C.c_void_p))
    assert rc == 0
    best = r["haps"][:, 0].astype(np.int64) + 1                    # [N][2][L] codes
    want = np.zeros((70, words), np.uint64)
    for l in range(90):
        want[:, l // per] |= ((best[:, 0, l] | (best[:, 1, l] << bits)).astype(np.uint64)) << np.uint64((l % per) * 2 * bits)
    assert np.array_equal(packed, want.astype(np.uint32))
    dev.close()


@pytest.mark.parametrize("threads", [32, 64, 128, 256])
@pytest.mark.parametrize("K", [12, 16, 25, 32])
def test_sample_sizes_above_ten(K, threads):
    """--sample-size up to the documented limit of 32 runs the KMAX = 16 / 32 instantiations of every forward variant."""
    from test_gpu_forward import klist_equiv
    g = make_panel(24, 40, founders=6, missing=0.05, seed=50 + K)
    o, ll, dev, r = both(g, K=K, threads=threads)
    assert ll == r["ll"]
    most = 0
    for i in range(24):
        a = o.individual(i)
        assert a["pg"] == r["pg"][i] and a["n"] == r["nres"][i]
        klist_equiv(r["prior"][i, :a["n"]], r["haps"][i, :a["n"]], a["prior"], a["haps"])
        assert np.all(np.diff(r["prior"][i, :a["n"]]) <= 0)
        most = max(most, a["n"])
    assert most > 10                                        # the long lists are really exercised
    dev.close()
