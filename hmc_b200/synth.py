"""Synthetic genotype panels: a founder-mosaic population written in the PHASE
text format that the reference's ``HaploFile::readGenoData`` (HaploFile.cpp:54-118) parses.

Each haplotype copies one of ``founders`` random founder haplotypes and switches founder with
probability ``switch`` per locus; each allele then mutates with probability ``mutation``; with
probability ``missing`` a whole genotype (both alleles) is blanked.  The two truth haplotypes are
written in order, so switch-error against the input is meaningful.
"""
from __future__ import annotations

import numpy as np


def make_panel(n_ind: int, n_loci: int, *, alleles: int = 2, founders: int = 8, missing: float = 0.0,
               switch: float = 0.02, mutation: float = 0.005, seed: int = 1, shard: int = 0) -> np.ndarray:
    """Return int8 array [n_ind, 2, n_loci] of allele codes 0..alleles-1, -1 = missing.
    shard > 0 draws a fresh set of individuals from the SAME founder population (multi-GPU weak-scaling shards)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    fnd = rng.integers(0, alleles, size=(founders, n_loci), dtype=np.int8)
    if shard > 0:
        rng = np.random.Generator(np.random.PCG64(seed * 1000003 + shard))
    h = n_ind * 2
    cur = rng.integers(0, founders, size=h)
    sw = rng.random((h, n_loci)) < switch
    newf = rng.integers(0, founders, size=(h, n_loci))
    src = np.empty((h, n_loci), dtype=np.int64)
    for l in range(n_loci):
        cur = np.where(sw[:, l], newf[:, l], cur)
        src[:, l] = cur
    hap = fnd[src, np.arange(n_loci)[None, :]]
    mut = rng.random((h, n_loci)) < mutation
    alt = rng.integers(0, alleles, size=(h, n_loci), dtype=np.int8)
    hap = np.where(mut, alt, hap).astype(np.int8)
    g = hap.reshape(n_ind, 2, n_loci).copy()
    if missing > 0:
        miss = rng.random((n_ind, n_loci)) < missing
        g[:, 0, :][miss] = -1
        g[:, 1, :][miss] = -1
    return g


def write_phase(path: str, g: np.ndarray, *, kind: str = "S", ids=None) -> None:
    """Write allele codes as PHASE text. kind 'S': one character per allele ('0'..'9', '?' missing);
    kind 'M': integers code+1, -1 for missing."""
    n, _, L = g.shape
    with open(path, "w") as f:
        f.write(f"{n}\n{L}\n")
        f.write("P " + " ".join(str(1000 * i) for i in range(L)) + "\n")
        f.write(kind * L + "\n")
        if kind == "S":
            lut = np.array([ord("?")] + [ord("0") + i for i in range(10)], dtype=np.uint8)
            for i in range(n):
                f.write((ids[i] if ids else f"ind{i}") + "\n")
                for hp in range(2):
                    row = lut[g[i, hp].astype(np.int64) + 1]
                    f.write(" ".join(row.tobytes().decode()) + "\n")
        else:
            for i in range(n):
                f.write((ids[i] if ids else f"ind{i}") + "\n")
                for hp in range(2):
                    f.write(" ".join(str(int(a) + 1) if a >= 0 else "-1" for a in g[i, hp]) + "\n")


CONFIGS = {
    # name: (N, L, alleles, founders, missing, seed, kind)   -- SURVEY.md §8d
    "C1": (200, 100, 2, 8, 0.0, 1, "S"),
    "C2": (2000, 1000, 2, 16, 0.05, 2, "S"),
    "C3": (50000, 5000, 2, 32, 0.0, 3, "S"),
    "C4": (10000, 2000, 6, 16, 0.01, 4, "M"),
}


def make_config(name: str, n_ind: int | None = None, n_loci: int | None = None):
    N, L, A, F, m, seed, kind = CONFIGS[name]
    g = make_panel(n_ind or N, n_loci or L, alleles=A, founders=F, missing=m, seed=seed)
    return g, kind
