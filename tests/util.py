"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np

import refdump
from cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def golden(name: str) -> dict:
    if name not in _cache:
        _cache[name] = refdump.load(os.path.join(GOLDEN, name + ".dump.gz"))
    return _cache[name]


def model_opts(name: str) -> dict:
    return dict(CASES[name][4])


def read_phase_idx(path: str, sym: list) -> np.ndarray:
    """Read a PHASE file (HaploFile.cpp:54-118 layout) into allele indices using the dump's symbol tables."""
    with open(path) as f:
        n = int(f.readline()); L = int(f.readline())
        line = f.readline()
        if line.lstrip().startswith("P"):
            line = f.readline()
        types = line.strip()
        g = np.full((n, 2, L), -1, np.int8)
        for i in range(n):
            f.readline()
            for h in range(2):
                tok = f.readline().split()
                for l in range(L):
                    if types[l] == "S":
                        v = -1 if tok[l] in "?-" else ord(tok[l])
                    else:
                        v = int(tok[l]) if tok[l] not in "?-" else -1
                        v = v if v > 0 else -1
                    g[i, h, l] = sym[l].index(v) if v >= 0 else -1
    return g


def same_unordered_pair(a: np.ndarray, b: np.ndarray) -> bool:
    return (np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])) or (np.array_equal(a[0], b[1]) and np.array_equal(a[1], b[0]))


def rel_close(x, y, tol=1e-9):
    x = np.asarray(x, float); y = np.asarray(y, float)
    den = np.maximum(np.abs(x), np.abs(y))
    ok = (np.abs(x - y) <= tol * den) | (x == y)
    return bool(np.all(ok))


def path_prior(m: dict, h0: np.ndarray, h1: np.ndarray) -> float:
    """prior_probability of one phased pair under the automaton `m` (export_model layout): the product the reference builds along
    the path -- head f_a*f_b (HaploPair.cpp:27), then *= tp_a*tp_b per locus (HaploPair.cpp:42,51) -- doubled unless the pair is
    homozygous throughout (HaploBuilder.cpp:98).  Returns 0.0 when the path leaves the automaton."""
    hl = int(m["head_len"])
    heads = {}
    for h in m["heads"]:
        heads[bytes(m["al"][m["al_off"][h]:m["al_off"][h + 1]])] = int(h)
    a = heads.get(bytes(h0[:hl].astype(np.int8)))
    b = heads.get(bytes(h1[:hl].astype(np.int8)))
    if a is None or b is None:
        return 0.0
    like = float(m["freq"][a]) * float(m["freq"][b])
    for l in range(hl, len(h0)):
        a = int(m["succ"][a, h0[l]]); b = int(m["succ"][b, h1[l]])
        if a < 0 or b < 0:
            return 0.0
        like *= float(m["tp"][a]) * float(m["tp"][b])
    return like if np.array_equal(h0, h1) else like * 2.0
