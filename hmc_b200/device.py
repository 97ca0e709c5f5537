"""ctypes binding of the C-ABI (include/hmc_b200.h) -- the device path of HaploBuilder.

There is no CPU fallback: if libhmc_b200.so is missing or no B200 is present, constructing a
``DeviceBuilder`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("HMC_B200_LIB") or os.path.join(_HERE, "libhmc_b200.so")     # (the override is for kernel experiments)
_LIB = None

FLAG_SCALED = 1
FLAG_KEEP_TRELLIS = 2
FLAG_PANEL_ROWS = 4
CNT_NAMES = ("steps", "states", "max_states", "arrivals", "waves", "spill_steps", "launches", "estep_entries")

EXPORTS = [
    "hmc_version", "hmc_ctx_create", "hmc_ctx_destroy", "hmc_last_error", "hmc_set_panel", "hmc_set_model", "hmc_resolve_all",
    "hmc_resolve_all_device", "hmc_fetch_results", "hmc_estimate_counts", "hmc_estimate_counts_device", "hmc_normalise_counts",
    "hmc_get_counters", "hmc_get_kernel_ms", "hmc_get_trellis", "hmc_mine_patterns", "hmc_fetch_patterns", "hmc_map_patterns", "hmc_mine_patterns_resolved", "hmc_fetch_best_haps", "hmc_result_dims", "hmc_compare", "hmc_set_panel_packed", "hmc_fetch_best_packed",
    "hmc_build_automaton", "hmc_fetch_automaton", "hmc_get_mine_stats", "hmc_result_device_ptrs", "hmc_fetch_counts", "hmc_stream",
    "hmc_get_likelihood", "hmc_format_best",
]


class HmcError(RuntimeError):
    pass


def library_path() -> str:
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise HmcError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        L = C.CDLL(_SO)
        L.hmc_version.restype = C.c_char_p
        L.hmc_last_error.restype = C.c_char_p
        L.hmc_last_error.argtypes = [C.c_void_p]
        L.hmc_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_int]
        L.hmc_ctx_destroy.argtypes = [C.c_void_p]
        L.hmc_set_panel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hmc_set_model.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
        L.hmc_resolve_all.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint] + [C.c_void_p] * 8
        L.hmc_resolve_all_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint]
        L.hmc_fetch_results.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.hmc_estimate_counts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_int] + [C.c_void_p] * 7
        L.hmc_estimate_counts_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_void_p)]
        L.hmc_normalise_counts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hmc_normalise_counts.restype = None
        L.hmc_mine_patterns.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.hmc_fetch_patterns.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.hmc_build_automaton.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
        L.hmc_fetch_automaton.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hmc_get_mine_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)] + [C.POINTER(C.c_int64)] * 5
        L.hmc_compare.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.hmc_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.hmc_get_kernel_ms.argtypes = [C.c_void_p, C.c_void_p]
        L.hmc_get_trellis.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong] + [C.c_void_p] * 4
        L.hmc_get_trellis.restype = C.c_longlong
        L.hmc_result_device_ptrs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hmc_fetch_counts.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hmc_stream.restype = C.c_void_p
        L.hmc_stream.argtypes = [C.c_void_p]
        if hasattr(L, "hmc_get_likelihood"):       # (absent from older builds loaded through HMC_B200_LIB for A/B timing)
            L.hmc_get_likelihood.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.hmc_format_best.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceBuilder:
    """The HaploBuilder surface (HaploBuilder.h:33-54) batched over individuals, on one B200."""

    def __init__(self, device: int = 0, arena_bytes: int = 0, threads: int = 0):
        self._h = C.c_void_p()
        rc = lib().hmc_ctx_create(C.byref(self._h), device, arena_bytes, threads)
        if rc != 0:
            self._h = None
            raise HmcError(f"hmc_ctx_create failed with status {rc} (no B200 / sm_100a device?) -- there is no CPU fallback")
        self.N = self.L = self.Amax = 0

    def close(self):
        if getattr(self, "_h", None):
            if not getattr(self, "_borrowed", False):
                lib().hmc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise HmcError(f"status {rc}: {lib().hmc_last_error(self._h).decode()}")
        return rc

    def set_panel(self, geno: np.ndarray, n_alleles: np.ndarray, allele_freq: np.ndarray):
        g = np.ascontiguousarray(geno, np.int8)
        self.N, _, self.L = g.shape
        na = np.ascontiguousarray(n_alleles, np.int32)
        af = np.ascontiguousarray(allele_freq, np.float64)
        self.Amax = af.shape[1]
        self._ck(lib().hmc_set_panel(self._h, self.N, self.L, self.Amax, _p(g), _p(na), _p(af)))

    def set_model(self, m: dict):
        """m: start,end,freq,tp,succ[P][Amax],al_off,al,heads,head_len (the layout of Oracle.export_model / HaploModel.export_model)."""
        a = {k: np.ascontiguousarray(m[k], dt) for k, dt in (("start", np.int32), ("end", np.int32), ("freq", np.float64), ("tp", np.float64),
                                                               ("succ", np.int32), ("al_off", np.int64), ("al", np.int8), ("heads", np.int32))}
        self._ck(lib().hmc_set_model(self._h, int(m["P"]), int(m["head_len"]), _p(a["start"]), _p(a["end"]), _p(a["freq"]), _p(a["tp"]),
                                     _p(a["succ"]), _p(a["al_off"]), _p(a["al"]), len(a["heads"]), _p(a["heads"])))

    def _alloc(self, count, K, haps):
        r = dict(pg=np.zeros(count), exp=np.zeros(count, np.int32), coverage=np.zeros(count), nres=np.zeros(count, np.int32),
                 prior=np.zeros((count, K)), posterior=np.zeros((count, K)), ll=np.zeros(1))
        r["haps"] = np.zeros((count, K, 2, self.L), np.int8) if haps else None
        return r

    def resolve_all(self, K: int = 10, first: int = 0, count: int | None = None, flags: int = 0, haps: bool = True, out: dict | None = None) -> dict:
        """`out`: caller-owned result arrays as returned by `alloc_results` (e.g. page-locked), filled in place -- the C-ABI writes
        into caller buffers; without it fresh arrays are allocated per call."""
        count = self.N - first if count is None else count
        r = dict(out) if out is not None else self._alloc(count, K, haps)
        if out is not None:
            assert r["prior"].shape == (count, K) and r["pg"].shape == (count,) and (r["haps"] is None or r["haps"].shape == (count, K, 2, self.L))
            r["ll"] = np.zeros(1)
        self._ck(lib().hmc_resolve_all(self._h, first, count, K, flags, _p(r["pg"]), _p(r["exp"]), _p(r["coverage"]), _p(r["nres"]),
                                       _p(r["prior"]), _p(r["posterior"]), _p(r["haps"]), _p(r["ll"])))
        r["ll"] = float(r["ll"][0])
        return r

    def alloc_results(self, count: int, K: int, haps: bool = True) -> dict:
        return self._alloc(count, K, haps)

    def resolve_all_device(self, K: int = 10, first: int = 0, count: int | None = None, flags: int = 0):
        count = self.N - first if count is None else count
        self._ck(lib().hmc_resolve_all_device(self._h, first, count, K, flags))
        self._last = (self.N if flags & FLAG_PANEL_ROWS else count, K)

    def result_device_ptrs(self):
        """[(device pointer, bytes per individual)] of {pg, exp, coverage, nres, prior, posterior, haps} of the last resolve_all_device call."""
        ptrs = (C.c_void_p * 7)()
        rb = (C.c_int64 * 7)()
        self._ck(lib().hmc_result_device_ptrs(self._h, ptrs, rb))
        return [(int(ptrs[i] or 0), int(rb[i])) for i in range(7)]

    def fetch_results(self, haps: bool = True, out: dict | None = None) -> dict:
        count, K = self._last
        r = dict(out) if out is not None else self._alloc(count, K, haps)
        if out is not None:
            assert r["prior"].shape == (count, K) and r["pg"].shape == (count,)
            r["ll"] = np.zeros(1)
        self._ck(lib().hmc_fetch_results(self._h, _p(r["pg"]), _p(r["exp"]), _p(r["coverage"]), _p(r["nres"]), _p(r["prior"]),
                                         _p(r["posterior"]), _p(r["haps"]), _p(r["ll"])))
        r["ll"] = float(r["ll"][0])
        return r

    def estimate_counts(self, start, al_off, al, pg, pg_exp=None, first: int = 0, count: int | None = None, flags: int = 0):
        count = self.N - first if count is None else count
        start = np.ascontiguousarray(start, np.int32); al_off = np.ascontiguousarray(al_off, np.int64); al = np.ascontiguousarray(al, np.int8)
        pg = np.ascontiguousarray(pg, np.float64)
        pe = None if pg_exp is None else np.ascontiguousarray(pg_exp, np.int32)
        Pc = len(start)
        f, pf = np.zeros(Pc), np.zeros(Pc)
        self._ck(lib().hmc_estimate_counts(self._h, first, count, flags, Pc, _p(start), _p(al_off), _p(al), _p(pg), _p(pe), _p(f), _p(pf)))
        return f, pf

    def estimate_counts_device(self, start, al_off, al, pg, pg_exp=None, first: int = 0, count: int | None = None, flags: int = 0):
        """E-pass whose raw sums stay in HBM: returns a zero-copy torch.float64 CUDA tensor of 2*Pc entries (frequency sums, then prefix
        sums) owned by the context -- all-reduce it in place (NCCL) before fetching / normalising."""
        import torch
        count = self.N - first if count is None else count
        start = np.ascontiguousarray(start, np.int32); al_off = np.ascontiguousarray(al_off, np.int64); al = np.ascontiguousarray(al, np.int8)
        pg = np.ascontiguousarray(pg, np.float64)
        pe = None if pg_exp is None else np.ascontiguousarray(pg_exp, np.int32)
        Pc = len(start)
        ptr = C.c_void_p()
        self._ck(lib().hmc_estimate_counts_device(self._h, first, count, flags, Pc, _p(start), _p(al_off), _p(al), _p(pg), _p(pe), C.byref(ptr)))

        class _Buf:
            __cuda_array_interface__ = {"shape": (2 * Pc,), "typestr": "<f8", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(_Buf(), device="cuda")

    @staticmethod
    def normalise_counts(freq, prefix, N):
        f = np.array(freq, np.float64); pf = np.array(prefix, np.float64); tp = np.zeros_like(f)
        lib().hmc_normalise_counts(len(f), N, _p(f), _p(pf), _p(tp))
        return f, pf, tp

    def mine_patterns(self, min_freq: float, min_len, max_len, samples=None, weights=None, total_weight: float = 0.0) -> dict:
        """PatternManager::searchPattern on the device (K5).  samples [n][L] int8 + weights [n] select the sampling branch;
        min_len / max_len: scalars or per-start arrays [L].  Returns the pattern list in the reference's id order."""
        L = self.L
        mn = np.ascontiguousarray(np.broadcast_to(np.asarray(min_len, np.int32), (L,)))
        mx = np.ascontiguousarray(np.broadcast_to(np.asarray(max_len, np.int32), (L,)))
        n = 0
        if samples is not None:
            samples = np.ascontiguousarray(samples, np.int8)
            weights = np.ascontiguousarray(weights, np.float64)
            n = samples.shape[0]
            assert samples.shape == (n, L) and weights.shape == (n,)
        P, na = C.c_int64(), C.c_int64()
        self._ck(lib().hmc_mine_patterns(self._h, n, _p(samples), _p(weights), float(total_weight), float(min_freq), _p(mn), _p(mx), C.byref(P), C.byref(na)))
        P, na = P.value, na.value
        r = dict(P=P, start=np.zeros(P, np.int32), end=np.zeros(P, np.int32), freq=np.zeros(P), prefix=np.zeros(P), tp=np.zeros(P),
                 al_off=np.zeros(P + 1, np.int64), alleles=np.zeros(max(na, 1), np.int8))
        self._ck(lib().hmc_fetch_patterns(self._h, _p(r["start"]), _p(r["end"]), _p(r["freq"]), _p(r["prefix"]), _p(r["tp"]), _p(r["al_off"]), _p(r["alleles"])))
        r["alleles"] = r["alleles"][:na]
        return r

    def build_automaton(self, head_len: int = 1) -> dict:
        """PatternManager::initialize on the device over the last mined list; the model is installed in the context."""
        nh = C.c_int32()
        self._ck(lib().hmc_build_automaton(self._h, head_len, C.byref(nh)))
        return dict(n_heads=nh.value)

    def fetch_automaton(self, P: int, n_heads: int) -> dict:
        succ = np.zeros((P, self.Amax), np.int32)
        heads = np.zeros(max(n_heads, 1), np.int32)
        self._ck(lib().hmc_fetch_automaton(self._h, _p(succ), _p(heads)))
        return dict(succ=succ, heads=heads[:n_heads])

    def compare(self, real: np.ndarray, infer: np.ndarray | None = None) -> dict:
        """HaploComp of `infer` (None = the best resolutions of the last resolve call, on the device) against `real`."""
        real = np.ascontiguousarray(real, np.int8)
        n, _, l = real.shape
        if infer is not None:
            infer = np.ascontiguousarray(infer, np.int8)
            assert infer.shape == real.shape
        out = np.zeros(6, np.int64)
        bi, bl = C.c_int32(-1), C.c_int32(-1)
        rc = lib().hmc_compare(self._h, n, l, _p(real), _p(infer), _p(out), C.byref(bi), C.byref(bl))
        if rc < 0 and bi.value < 0:
            self._ck(rc)
        with np.errstate(divide="ignore", invalid="ignore"):
            f = out.astype(np.float64)
            return dict(counts=out, switch_error=float(f[0] / f[1]), igp=float(f[2] / f[3]), ihp=float(f[4] / f[5]), bad_individual=bi.value,
                        bad_locus=bl.value)

    def likelihood(self, haps: np.ndarray | None = None) -> np.ndarray:
        """HaploBuilder::getLikelihood(const Haplotype &) (HaploBuilder.cpp:128-146) of every row of haps [n][L]; None = both haplotypes of
        the best resolution of every individual of the last resolve call, on the device ([rows][2])."""
        if haps is None:
            cnt, k = C.c_int32(), C.c_int32()
            self._ck(lib().hmc_result_dims(self._h, C.byref(cnt), C.byref(k)))
            out = np.zeros((cnt.value, 2))
            self._ck(lib().hmc_get_likelihood(self._h, 2 * cnt.value, None, _p(out)))
            return out
        haps = np.ascontiguousarray(haps, np.int8)
        out = np.zeros(haps.shape[:-1])
        self._ck(lib().hmc_get_likelihood(self._h, int(out.size), _p(haps), _p(out)))
        return out

    def format_best(self, symbols: np.ndarray, allele_type: str):
        """The allele lines of `.reconstructed` for the last resolve call, formatted on the device (HaploFile.cpp:146-148): returns
        (text bytes, line offsets [2 * rows + 1])."""
        symbols = np.ascontiguousarray(symbols, np.int32)
        cnt, k = C.c_int32(), C.c_int32()
        self._ck(lib().hmc_result_dims(self._h, C.byref(cnt), C.byref(k)))
        off = np.zeros(2 * cnt.value + 1, np.int64)
        total = C.c_int64()
        self._ck(lib().hmc_format_best(self._h, _p(symbols), allele_type.encode(), None, 0, _p(off), C.byref(total)))
        buf = np.zeros(total.value, np.uint8)
        self._ck(lib().hmc_format_best(self._h, _p(symbols), allele_type.encode(), _p(buf), total.value, _p(off), C.byref(total)))
        return buf.tobytes(), off

    def mine_stats(self) -> dict:
        ms, nodes, levels, mx, sw, wr = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(lib().hmc_get_mine_stats(self._h, C.byref(ms), C.byref(nodes), C.byref(levels), C.byref(mx), C.byref(sw), C.byref(wr)))
        return dict(ms=ms.value, nodes=nodes.value, levels=levels.value, max_list=mx.value, swept=sw.value, written=wr.value)

    def counters(self) -> dict:
        out = np.zeros(8, np.int64)
        self._ck(lib().hmc_get_counters(self._h, _p(out)))
        return {k: int(v) for k, v in zip(CNT_NAMES, out) if k != "_"}

    def kernel_ms(self) -> dict:
        out = np.zeros(4, np.float32)
        self._ck(lib().hmc_get_kernel_ms(self._h, _p(out)))
        return dict(forward=float(out[0]), traceback=float(out[1]), backward=float(out[2]), estep=float(out[3]))

    def trellis(self, individual: int, capacity: int = 1 << 22, beta: bool = False) -> dict:
        counts = np.zeros(self.L + 1, np.int32)
        ida = np.zeros(capacity, np.int32); idb = np.zeros(capacity, np.int32); alpha = np.zeros(capacity)
        bt = np.zeros(capacity) if beta else None
        tot = self._ck(lib().hmc_get_trellis(self._h, individual, _p(counts), capacity, _p(ida), _p(idb), _p(alpha), _p(bt)))
        return dict(counts=counts, ida=ida[:tot], idb=idb[:tot], alpha=alpha[:tot], beta=None if bt is None else bt[:tot])
