"""-m gpu: the host layer end to end -- HaploModel::run (sampling and exact EM) and the hmc CLI -- against the reference dumps."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from hmc_b200 import HaploModel
from hmc_b200.build import CLI
from oracle.pyoracle import Oracle
from util import GOLDEN, golden, model_opts, path_prior, read_phase_idx, rel_close, same_unordered_pair

pytestmark = pytest.mark.gpu


def tie_ok(name, i, ours, theirs, d):
    """Top-1 differs: acceptable only if the reference's own K-list holds our pair with a prior bitwise equal to its top-1."""
    o = Oracle(d["geno"], **{**model_opts(name), "max_iteration": 1})
    o.find_patterns()
    o.resolve_all()
    a = o.individual(i)
    for k in range(a["n"]):
        if a["prior"][k] == a["prior"][0] and same_unordered_pair(a["haps"][k], ours):
            return True
    # more tied paths than the K-list holds: our pair must carry exactly the reference's top-1 prior under the same automaton
    return path_prior(o.export_model(), ours[0], ours[1]) == a["prior"][0]


@pytest.mark.parametrize("name", ["g1_clean", "g2_missing", "g3_multi", "g8_head2", "g9_k1", "g7_bynum"])
def test_single_pass_run_matches_reconstructed(name):
    d = golden(name)
    hm = HaploModel(**model_opts(name))
    hm.load_phase(os.path.join(GOLDEN, name + ".phase"))
    ll, npat = hm.run()
    assert len(ll) == 1 and ll[0] == d["iters"][0]["ll"] and npat[0] == d["iters"][0]["P"]
    rec = read_phase_idx(os.path.join(GOLDEN, name + ".phase.reconstructed"), d["sym"])
    res = hm.resolutions()
    for i in range(d["N"]):
        if not same_unordered_pair(res[i], rec[i]):
            assert tie_ok(name, i, res[i], rec[i], d), i
    hm.close()


@pytest.mark.parametrize("name", ["g4_em", "g5_exact", "g6_mc"])
def test_em_trace(name):
    d = golden(name)
    hm = HaploModel(**model_opts(name))
    hm.load_phase(os.path.join(GOLDEN, name + ".phase"))
    ll, npat = hm.run()
    ref_ll = np.array([it["ll"] for it in d["iters"]])
    ref_np = np.array([it["P"] for it in d["iters"]])
    assert len(ll) == len(ref_ll), (ll, ref_ll)
    assert np.array_equal(npat, ref_np)
    assert ll[0] == ref_ll[0]
    # later iterations: K-lists are equal up to which of several exactly tied paths sits at the K-th place, and
    # the exact E-step sums in a different order than the reference's pointer-keyed maps (F7)
    assert rel_close(ll, ref_ll, 1e-9)
    rec = read_phase_idx(os.path.join(GOLDEN, name + ".phase.reconstructed"), d["sym"])
    res = hm.resolutions()
    # differing top-1 pairs must be ties: both pairs carry the same prior under the final automaton (to 1e-9: the exact M-step's
    # frequencies agree with the reference's to rounding only)
    m = hm.export_model()
    bad = [i for i in range(d["N"]) if not same_unordered_pair(res[i], rec[i])]
    not_tied = [i for i in bad if not rel_close(path_prior(m, res[i][0], res[i][1]), path_prior(m, rec[i][0], rec[i][1]), 1e-9)]
    assert len(ll) < model_opts(name)["max_iteration"] or not not_tied or len(not_tied) <= 1, (bad, not_tied)
    if len(ll) < model_opts(name)["max_iteration"] and ll[-1] < ll[-2]:
        pass        # the kept resolutions belong to the previous iteration's automaton; the trace above is the check
    else:
        assert not not_tied, (bad, not_tied)
    hm.close()


@pytest.mark.parametrize("name", ["g1_clean", "g3_multi", "g4_em"])
def test_cli_writes_the_reference_output(name, tmp_path):
    d = golden(name)
    phase = str(tmp_path / (name + ".phase"))
    shutil.copy(os.path.join(GOLDEN, name + ".phase"), phase)
    opts = []
    mo = model_opts(name)
    if "max_iteration" in mo:
        opts += ["-i", str(mo["max_iteration"])]
    subprocess.check_call([CLI, "--nologo", "-d", "1"] + opts + [phase])
    ours = open(phase + ".reconstructed").read().split("\n")
    ref = open(os.path.join(GOLDEN, name + ".phase.reconstructed")).read().split("\n")
    assert ours[:4] == ref[:4]                      # N, L, positions, types
    assert ours[4::3] == ref[4::3]                  # id lines
    a = read_phase_idx(phase + ".reconstructed", d["sym"])
    b = read_phase_idx(os.path.join(GOLDEN, name + ".phase.reconstructed"), d["sym"])
    diff = [i for i in range(d["N"]) if not same_unordered_pair(a[i], b[i])]
    assert len(diff) <= max(1, d["N"] // 20), diff  # exact-tie individuals only


def test_no_missing_allele_in_output():
    d = golden("g2_missing")
    hm = HaploModel()
    hm.load_phase(os.path.join(GOLDEN, "g2_missing.phase"))
    hm.run()
    assert (hm.resolutions() >= 0).all()
    hm.close()
