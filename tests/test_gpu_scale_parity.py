"""-m gpu: oracle parity AT the BASELINE shapes.  The model is the full-size one (mined on the device from the whole panel); the
oracle port (oracle/hmc_oracle.hpp) resolves a subsample of individuals against exactly that model -- the individuals with the most
missing genotypes, the ones with the largest live state sets and the ones the kernel spent the most states on, plus a stride over
the rest -- and P(G), priors, posteriors and counters must be bit-identical, the K-lists tie-aware identical.

  C2  2,000 x 1,000, 5 % missing, defaults            (configs[1], the bench workload; state sets up to 54,355, spill steps)
  C4  10,000 x 2,000, 6 alleles ('M' loci), 1 % missing   (configs[3]; scaled: fp64 underflows at L = 2,000)
  C3  deep + scaled: 5,000 loci, max-pattern-len 100, 32 founders  (configs[2] at a reduced cohort: the oracle subsample fixes the cost)
"""
import ctypes as C
import os

import numpy as np
import pytest

from hmc_b200 import FLAG_SCALED, HaploModel
from hmc_b200.device import lib as dev_lib
from hmc_b200.synth import make_config, make_panel
from oracle.pyoracle import Oracle
from test_gpu_forward import klist_equiv

pytestmark = pytest.mark.gpu


def per_individual_work(dev, N):
    """{individual: (states, max states, spill steps)} of the last resolve pass (HMC_B200_DEBUG_CYCLES diagnostics)."""
    d = np.zeros((N, 8), np.int64)
    inds = np.zeros(N, np.int32)
    n = dev_lib().hmc_debug_cycles(dev._h, d.ctypes.data_as(C.c_void_p), inds.ctypes.data_as(C.c_void_p), N)
    return {int(inds[q]): (int(d[q, 1]), int(d[q, 2]), int(d[q, 4])) for q in range(n)}


def pick(g, work, n_each, stride):
    miss = (g[:, 0] < 0).sum(axis=1)
    by_missing = list(np.argsort(-miss, kind="stable")[:n_each])
    by_max = sorted(work, key=lambda i: -work[i][1])[:n_each]
    by_states = sorted(work, key=lambda i: -work[i][0])[:n_each]
    rest = list(range(0, g.shape[0], stride))
    sel = []
    for i in by_missing + by_max + by_states + rest:
        if int(i) not in sel:
            sel.append(int(i))
    return sel


def check_against_oracle(g, kind, sel, r, m, na, af, K, scaled, max_len):
    idx_sel = np.ascontiguousarray(g[sel])
    o = Oracle(idx_sel, sample_size=K, scaled=scaled, max_pattern_len=max_len)
    o.set_allele_table(na, af)                       # the full panel's alleles and frequencies, not the subsample's own
    o.import_model(m, min_len=1, max_len=max_len, min_freq=0.0)
    o.resolve_all()
    for q, i in enumerate(sel):
        a = o.individual(q)
        assert a["pg"] == r["pg"][i], (i, a["pg"], r["pg"][i])
        assert a["exp"] == r["exp"][i] and a["n"] == r["nres"][i]
        n = a["n"]
        assert np.array_equal(a["posterior"], r["posterior"][i, :n])
        klist_equiv(r["prior"][i, :n], r["haps"][i, :n], a["prior"], a["haps"])
    return o


def run_config(g, kind, K, scaled, max_len, n_each, stride):
    os.environ["HMC_B200_DEBUG_CYCLES"] = "1"
    try:
        hm = HaploModel(scaled=scaled, max_pattern_len=max_len, sample_size=K)
        hm.set_panel(g, kind)
        hm.build()
        hm.upload()
        assert hm.mined_on_device()
        m = hm.export_model()
        idx, na, af = hm.panel()
        dev = hm.device()
        r = dev.resolve_all(K=K, flags=FLAG_SCALED if scaled else 0)
        c = dev.counters()
        # (several arena waves: the diagnostics only cover the last one -- the selection then goes by missing genotypes and stride)
        work = per_individual_work(dev, g.shape[0]) if c["waves"] == 1 else {}
    finally:
        del os.environ["HMC_B200_DEBUG_CYCLES"]
    sel = pick(idx, work, n_each, stride)
    o = check_against_oracle(idx, kind, sel, r, m, na, af, K, scaled, max_len)
    # structural counters of the subsample: every selected individual alone on the device, summed, against the oracle's
    states = arrivals = 0
    for i in sel:
        rs = dev.resolve_all(K=K, flags=FLAG_SCALED if scaled else 0, first=i, count=1)
        assert rs["pg"][0] == r["pg"][i]                # (and a pass over one individual gives what the full pass gave)
        ci = dev.counters()
        states += ci["states"]; arrivals += ci["arrivals"]
    oc = o.counters()
    assert oc["states"] == states and oc["arrivals"] == arrivals
    hm.close()
    return c, work, sel, o


def test_c2_full_model_subsample_vs_oracle():
    g, kind = make_config("C2")
    c, work, sel, o = run_config(g, kind, 10, False, 30, n_each=6, stride=125)
    assert c["steps"] == 2000 * 999 and c["spill_steps"] > 0 and c["max_states"] > 30000
    assert max(work[i][1] for i in sel) == c["max_states"]          # the individual with the largest state set is in the subsample
    assert sum(work[i][2] for i in sel) > 0                         # ... and steps that left the shared-memory fast path
    oc = o.counters()
    assert oc["states"] == sum(work[i][0] for i in sel) and oc["max_states"] == c["max_states"]


def test_c4_shape_multiallelic_scaled_vs_oracle():
    g, kind = make_config("C4")
    assert g.shape == (10000, 2, 2000) and int(g.max()) == 5
    c, work, sel, o = run_config(g, kind, 10, True, 30, n_each=3, stride=1250)
    assert c["steps"] == 10000 * 1999


def test_c3_shape_deep_scaled_vs_oracle():
    """configs[2]'s shape per individual (5,000 SNPs, 32 founders, --max-pattern-len 100, exponent rescaling) on a 3,000-individual cohort."""
    g = make_panel(3000, 5000, founders=32, seed=3)
    c, work, sel, o = run_config(g, "S", 10, True, 100, n_each=2, stride=750)
    assert c["steps"] == 3000 * 4999
