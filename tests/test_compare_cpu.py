"""The oracle's HaploComp restatement (oracle/hmc_oracle.hpp compare_genotypes) against the UNMODIFIED reference's
HaploComp (tests/golden/compare.json, written by tests/golden/make_compare.py through oracle/_ref/hmc_ref --compare)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import pyoracle
from cases import CASES
from util import GOLDEN, golden, read_phase_idx

FIX = json.load(open(os.path.join(GOLDEN, "compare.json")))


def hexf(s: str) -> float:
    return float("nan") if "nan" in s else float.fromhex(s)


def same(x: float, y: float) -> bool:
    return (math.isnan(x) and math.isnan(y)) or x == y


def check(r: dict, ref: dict):
    assert same(r["switch_error"], hexf(ref["switch_error"])), (r, ref)
    assert same(r["ihp"], hexf(ref["ihp"])), (r, ref)
    assert same(r["igp"], hexf(ref["igp"])), (r, ref)


def golden_pair(name):
    d = golden(name)
    real = read_phase_idx(os.path.join(GOLDEN, name + ".phase"), d["sym"])
    infer = read_phase_idx(os.path.join(GOLDEN, name + ".phase.reconstructed"), d["sym"])
    return real, infer


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_compare_on_golden_reconstructions(name):
    real, infer = golden_pair(name)
    r = pyoracle.compare(real, infer)
    assert r["bad_individual"] < 0
    check(r, FIX["golden"][name])


@pytest.mark.parametrize("k", range(len(FIX["planted"])))
def test_oracle_compare_on_planted_errors(k):
    p = FIX["planted"][k]
    r = pyoracle.compare(np.array(p["real"], np.int8), np.array(p["infer"], np.int8))
    if "error" in p["ref"]:
        assert r["bad_individual"] >= 0         # the reference exits with "Inconsistent genotypes at locus %d!"
    else:
        assert r["bad_individual"] < 0
        check(r, p["ref"])
