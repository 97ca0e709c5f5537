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
