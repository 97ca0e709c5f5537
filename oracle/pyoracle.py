"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the oracle port (oracle/libhmc_oracle.so).

Imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by hmc_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libhmc_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "hmc_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libhmc_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_options.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double] + [C.c_int] * 8
        L.orc_total_pattern_len.restype = C.c_longlong
        L.orc_resolve_all.restype = C.c_double
        L.orc_trellis_run.restype = C.c_longlong
        L.orc_now.restype = C.c_double
        for name in ("orc_amax", "orc_get_allele_table", "orc_find_patterns", "orc_num_patterns", "orc_head_len", "orc_num_heads",
                     "orc_total_pattern_len", "orc_export_model", "orc_resolve_all", "orc_counters", "orc_num_res",
                     "orc_get_individual", "orc_get_resolution", "orc_trellis_get", "orc_estimate_patterns", "orc_run", "orc_get_trace"):
            fn = getattr(L, name)
            if fn.argtypes is None:
                fn.argtypes = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One panel + HaploModel options; mirrors the reference's HaploModel surface."""

    def __init__(self, geno_idx: np.ndarray, **opts):
        g = np.ascontiguousarray(geno_idx, dtype=np.int8)
        assert g.ndim == 3 and g.shape[1] == 2
        self.N, _, self.L = g.shape
        self._g = g
        self._h = C.c_void_p(lib().orc_create(self.N, self.L, _p(g)))
        self.opts = dict(model="MV", min_freq=-1.0, min_freq_abs=1.5, num_patterns=-1, min_pattern_len=1, max_pattern_len=30,
                         mc_order=1, max_iteration=1, sample_size=10, exact_estimate=False, scaled=False)
        self.set_options(**opts)
        self.Amax = lib().orc_amax(self._h)

    def __del__(self):
        try:
            lib().orc_destroy(self._h)
        except Exception:
            pass

    def set_options(self, **opts):
        self.opts.update(opts)
        o = self.opts
        lib().orc_set_options(self._h, o["model"].encode(), o["min_freq"], o["min_freq_abs"], o["num_patterns"], o["min_pattern_len"],
                              o["max_pattern_len"], o["mc_order"], o["max_iteration"], o["sample_size"], int(o["exact_estimate"]), int(o["scaled"]))

    def allele_table(self):
        nall = np.zeros(self.L, np.int32)
        afreq = np.zeros((self.L, self.Amax), np.float64)
        lib().orc_get_allele_table(self._h, _p(nall), _p(afreq))
        return nall, afreq

    def set_allele_table(self, nall, afreq):
        nall = np.ascontiguousarray(nall, np.int32)
        afreq = np.ascontiguousarray(afreq, np.float64)
        self.Amax = afreq.shape[1]
        lib().orc_set_allele_table(self._h, C.c_int(self.Amax), _p(nall), _p(afreq))

    def find_patterns(self):
        lib().orc_find_patterns(self._h)

    def export_model(self) -> dict:
        L_ = lib()
        P = L_.orc_num_patterns(self._h)
        tl = L_.orc_total_pattern_len(self._h)
        H = L_.orc_num_heads(self._h)
        m = dict(P=P, head_len=L_.orc_head_len(self._h), Amax=self.Amax,
                 start=np.zeros(P, np.int32), end=np.zeros(P, np.int32), freq=np.zeros(P), prefix=np.zeros(P), tp=np.zeros(P),
                 succ=np.zeros((P, self.Amax), np.int32), al_off=np.zeros(P + 1, np.int64), al=np.zeros(tl, np.int8), heads=np.zeros(H, np.int32))
        L_.orc_export_model(self._h, _p(m["start"]), _p(m["end"]), _p(m["freq"]), _p(m["prefix"]), _p(m["tp"]), _p(m["succ"]),
                            _p(m["al_off"]), _p(m["al"]), _p(m["heads"]))
        return m

    def import_model(self, m: dict, min_len=1, max_len=30, min_freq=0.0):
        lib().orc_import_model(self._h, C.c_int(m["P"]), _p(m["start"]), _p(m["end"]), _p(m["freq"]), _p(m["prefix"]), _p(m["tp"]),
                               _p(m["al_off"]), _p(m["al"]), C.c_int(min_len), C.c_int(max_len), C.c_double(min_freq))

    def resolve_all(self) -> float:
        return lib().orc_resolve_all(self._h)

    def counters(self) -> dict:
        out = np.zeros(4, np.int64)
        lib().orc_counters(self._h, _p(out))
        return dict(steps=int(out[0]), states=int(out[1]), max_states=int(out[2]), arrivals=int(out[3]))

    def individual(self, i: int) -> dict:
        n = lib().orc_num_res(self._h, i)
        pg = C.c_double()
        ex = C.c_int()
        prior = np.zeros(max(n, 1))
        post = np.zeros(max(n, 1))
        haps = np.zeros((max(n, 1), 2, self.L), np.int8)
        lib().orc_get_individual(self._h, i, C.byref(pg), C.byref(ex), _p(prior), _p(post), _p(haps))
        return dict(pg=pg.value, exp=ex.value, n=n, prior=prior[:n], posterior=post[:n], haps=haps[:n])

    def likelihood(self, haps: np.ndarray) -> np.ndarray:
        """HaploBuilder::getLikelihood(const Haplotype &) (HaploBuilder.cpp:128-146) of every row of haps [n][L]."""
        haps = np.ascontiguousarray(haps, np.int8).reshape(-1, self.L)
        out = np.zeros(len(haps))
        lib().orc_get_likelihood(self._h, C.c_int(len(haps)), _p(haps), _p(out))
        return out

    def resolution(self, i: int) -> np.ndarray:
        h = np.zeros((2, self.L), np.int8)
        lib().orc_get_resolution(self._h, i, _p(h))
        return h

    def trellis(self, i: int, K: int = 1, with_beta: bool = False) -> dict:
        counts = np.zeros(self.L + 1, np.int32)
        tot = lib().orc_trellis_run(self._h, i, K, int(with_beta), _p(counts))
        t = dict(counts=counts, ida=np.zeros(tot, np.int32), idb=np.zeros(tot, np.int32), alpha=np.zeros(tot), tp=np.zeros(tot),
                 beta=np.zeros(tot), narr=np.zeros(tot, np.int32))
        lib().orc_trellis_get(self._h, _p(t["ida"]), _p(t["idb"]), _p(t["alpha"]), _p(t["tp"]), _p(t["beta"]), _p(t["narr"]))
        return t

    def estep(self, start: np.ndarray, al_off: np.ndarray, al: np.ndarray):
        Pc = len(start)
        start = np.ascontiguousarray(start, np.int32)
        al_off = np.ascontiguousarray(al_off, np.int64)
        al = np.ascontiguousarray(al, np.int8)
        f, pf, tp = np.zeros(Pc), np.zeros(Pc), np.zeros(Pc)
        lib().orc_estep(self._h, C.c_int(Pc), _p(start), _p(al_off), _p(al), _p(f), _p(pf), _p(tp))
        return f, pf, tp

    def estimate_patterns(self):
        lib().orc_estimate_patterns(self._h)

    def run(self):
        n = lib().orc_run(self._h)
        ll = np.zeros(n)
        npat = np.zeros(n, np.int32)
        lib().orc_get_trace(self._h, _p(ll), _p(npat))
        return ll, npat


def compare(real: np.ndarray, infer: np.ndarray) -> dict:
    """HaploComp (HaploComp.cpp:29-76) over allele codes [N][2][L] (negative = missing): counters and the three ratios."""
    real = np.ascontiguousarray(real, np.int8)
    infer = np.ascontiguousarray(infer, np.int8)
    assert real.shape == infer.shape and real.ndim == 3 and real.shape[1] == 2
    out = np.zeros(6, np.int64)
    bad_locus = C.c_int(-1)
    L_ = lib()
    L_.orc_compare.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    bad = L_.orc_compare(_p(real), _p(infer), real.shape[0], real.shape[2], _p(out), C.byref(bad_locus))
    with np.errstate(divide="ignore", invalid="ignore"):
        f = out.astype(np.float64)
        return dict(counts=out, switch_error=float(f[0] / f[1]), igp=float(f[2] / f[3]), ihp=float(f[4] / f[5]), bad_individual=bad,
                    bad_locus=bad_locus.value)
