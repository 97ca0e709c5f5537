"""K7 (HaploComp on the device, csrc/compare.cuh) through the C-ABI against the reference's numbers and the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle
from cases import CASES
from hmc_b200 import DeviceBuilder
from hmc_b200.build import CLI
from hmc_b200.synth import make_panel, write_phase
from test_compare_cpu import FIX, check, golden_pair
from util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = DeviceBuilder()
    yield d
    d.close()


@pytest.mark.parametrize("name", list(CASES))
def test_golden_reconstructions(dev, name):
    real, infer = golden_pair(name)
    r = dev.compare(real, infer)
    assert r["bad_individual"] < 0
    check(r, FIX["golden"][name])
    assert np.array_equal(r["counts"], pyoracle.compare(real, infer)["counts"])


@pytest.mark.parametrize("k", range(len(FIX["planted"])))
def test_planted_errors(dev, k):
    p = FIX["planted"][k]
    real, infer = np.array(p["real"], np.int8), np.array(p["infer"], np.int8)
    r = dev.compare(real, infer)
    o = pyoracle.compare(real, infer)
    if "error" in p["ref"]:
        assert r["bad_individual"] == o["bad_individual"] and r["bad_locus"] == o["bad_locus"] and r["bad_individual"] >= 0
    else:
        check(r, p["ref"])
        assert np.array_equal(r["counts"], o["counts"])


@pytest.mark.parametrize("shape", [(300, 517, 2), (64, 2000, 5), (1, 33, 2), (7, 31, 3)])
def test_random_switches_and_missing_vs_oracle(dev, shape):
    n, L, alleles = shape
    rng = np.random.default_rng(n * 1000 + L)
    real = make_panel(n, L, alleles=alleles, founders=6, seed=L)
    infer = real.copy()
    flip = np.cumsum(rng.random((n, L)) < 0.03, axis=1) % 2 == 1
    infer[:, 0, :] = np.where(flip, real[:, 1, :], real[:, 0, :])
    infer[:, 1, :] = np.where(flip, real[:, 0, :], real[:, 1, :])
    real[rng.random(real.shape) < 0.04] = -1
    infer[rng.random(infer.shape) < 0.02] = -1
    r, o = dev.compare(real, infer), pyoracle.compare(real, infer)
    assert o["bad_individual"] < 0 and np.array_equal(r["counts"], o["counts"])
    assert r["counts"][0] > 0 or n < 10


def test_best_resolutions_compared_on_the_device():
    """infer = None: the resolutions of the last resolve pass are compared where they lie (no D2H of the haplotypes)."""
    from oracle.pyoracle import Oracle
    g = make_panel(80, 60, founders=6, missing=0.03, seed=21)
    o = Oracle(g)
    o.find_patterns()
    nall, afreq = o.allele_table()
    d = DeviceBuilder()
    d.set_panel(g, nall, afreq)
    d.set_model(o.export_model())
    res = d.resolve_all(K=10)
    a = d.compare(g)                                    # real = the (phased) input panel, as in HaploModel::run
    b = d.compare(g, np.ascontiguousarray(res["haps"][:, 0]))
    assert np.array_equal(a["counts"], b["counts"]) and np.array_equal(a["counts"], pyoracle.compare(g, res["haps"][:, 0])["counts"])
    d.close()


def test_cli_compare_prints_the_reference_line(tmp_path):
    name = "g2_missing"
    out = subprocess.run([CLI, "-e", os.path.join(GOLDEN, name + ".phase.reconstructed"), os.path.join(GOLDEN, name + ".phase"), "-d", "0"],
                         capture_output=True, text=True, check=True).stdout
    ref = FIX["golden"][name]
    want = "Switch Error = %f, IHP = %f, IGP = %f" % (float.fromhex(ref["switch_error"]), float.fromhex(ref["ihp"]), float.fromhex(ref["igp"]))
    assert want in out, out
