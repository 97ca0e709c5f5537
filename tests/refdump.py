"""Parser for the text dumps written by oracle/ref_driver.cpp (the unmodified reference's output)."""
from __future__ import annotations

import gzip

import numpy as np

fh = float.fromhex


def _ints(s):
    return np.array([int(x) for x in s.split(",")], dtype=np.int8) if s else np.zeros(0, np.int8)


def load(path: str) -> dict:
    op = gzip.open if path.endswith(".gz") else open
    d = dict(iters=[])
    it = None
    ind = None
    with op(path, "rt") as f:
        assert f.readline().startswith("HMCDUMP")
        for line in f:
            t = line.split()
            k = t[0]
            if k == "panel":
                d["N"], d["L"], d["Amax"] = int(t[1]), int(t[2]), int(t[3])
                d["nall"] = np.zeros(d["L"], np.int32)
                d["afreq"] = np.zeros((d["L"], d["Amax"]))
                d["sym"] = [None] * d["L"]
                d["geno"] = np.zeros((d["N"], 2, d["L"]), np.int8)
                d["ids"] = [None] * d["N"]
            elif k == "locus":
                l, n = int(t[1]), int(t[3])
                d["nall"][l] = n
                d["sym"][l] = [int(t[4 + 2 * j]) for j in range(n)]
                for j in range(n):
                    d["afreq"][l, j] = fh(t[5 + 2 * j])
            elif k == "geno":
                i = int(t[1])
                d["ids"][i] = t[2]
                d["geno"][i, 0] = _ints(t[3])
                d["geno"][i, 1] = _ints(t[4])
            elif k == "iter":
                it = dict(patterns=None, inds=[], estep=None)
                d["iters"].append(it)
            elif k == "patterns":
                P = int(t[2])
                it["P"], it["head_len"] = P, int(t[3])
                it["patterns"] = dict(P=P, head_len=int(t[3]), start=np.zeros(P, np.int32), end=np.zeros(P, np.int32), freq=np.zeros(P),
                                      prefix=np.zeros(P), tp=np.zeros(P), succ=np.full((P, d["Amax"]), -1, np.int32), al=[None] * P, full=False)
            elif k == "p":
                m = it["patterns"]
                m["full"] = True
                i = int(t[1])
                m["start"][i], m["end"][i] = int(t[2]), int(t[3])
                m["freq"][i], m["prefix"][i], m["tp"][i] = fh(t[4]), fh(t[5]), fh(t[6])
                an = int(t[7])
                for j in range(an):
                    m["succ"][i, j] = int(t[8 + j])
                m["al"][i] = _ints(t[8 + an])
            elif k == "heads":
                it["patterns"]["heads"] = np.array([int(x) for x in t[1:]], np.int32)
            elif k == "ll":
                it["ll"] = fh(t[2])
            elif k == "pg":
                it["pg"] = np.array([fh(x) for x in t[1:]])
            elif k == "accuracy":
                it["accuracy"] = tuple(fh(x) for x in t[1:4])
            elif k == "ind":
                ind = dict(pg=fh(t[2]), coverage=fh(t[3]), n=int(t[4]), prior=[], posterior=[], haps=[], states={})
                it["inds"].append(ind)
            elif k == "res":
                ind["prior"].append(fh(t[2]))
                ind["posterior"].append(fh(t[3]))
                ind["haps"].append(np.stack([_ints(t[4]), _ints(t[5])]))
            elif k == "states":
                cur = dict(ida=[], idb=[], alpha=[], tp=[], beta=[])
                ind["states"][int(t[2])] = cur
            elif k == "s":
                cur["ida"].append(int(t[1])); cur["idb"].append(int(t[2])); cur["alpha"].append(fh(t[3])); cur["tp"].append(fh(t[4]))
                if len(t) > 5:
                    cur["beta"].append(fh(t[5]))
            elif k == "cnt":
                ind["cnt"] = tuple(int(x) for x in t[2:5])
            elif k == "counters":
                it["counters"] = dict(steps=int(t[1]), states=int(t[2]), max_states=int(t[3]), arrivals=int(t[4]))
            elif k == "lk":
                it.setdefault("lk", []).append((fh(t[2]), fh(t[3]), fh(t[4])))
            elif k == "estep":
                P = int(t[1])
                it["estep"] = dict(freq=np.zeros(P), prefix=np.zeros(P), tp=np.zeros(P))
            elif k == "e":
                i = int(t[1])
                it["estep"]["freq"][i], it["estep"]["prefix"][i], it["estep"]["tp"][i] = fh(t[2]), fh(t[3]), fh(t[4])
            elif k == "end":
                d["n_iter"] = int(t[1])
    return d


def model_arrays(m: dict) -> dict:
    """pattern dict from load() -> CSR arrays as used by Oracle.import_model / hmc_set_model."""
    lens = np.array([len(a) for a in m["al"]], np.int64)
    off = np.zeros(m["P"] + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    out = dict(m)
    out["al_off"] = off
    out["al"] = np.concatenate(m["al"]).astype(np.int8) if m["P"] else np.zeros(0, np.int8)
    return out
