"""ctypes binding of the host library (hmc_b200/host/*): the reference's HaploModel surface
(HaploModel.h:10-40) -- option fields, run(), resolveAll(), findPatterns() -- driving the device path.
Mining and automaton construction run on host threads; everything per-individual runs on the GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .device import DeviceBuilder, HmcError, lib as dev_lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhmc_host.so")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            raise HmcError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        dev_lib()       # libhmc_b200.so first
        L = C.CDLL(_SO)
        L.hmch_create.restype = C.c_void_p
        L.hmch_create.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_char_p, C.c_int]
        L.hmch_last_error.restype = C.c_char_p
        L.hmch_ctx.restype = C.c_void_p
        L.hmch_total_pattern_len.restype = C.c_longlong
        for name in ("hmch_destroy", "hmch_last_error", "hmch_ctx", "hmch_load_phase", "hmch_build", "hmch_find_patterns", "hmch_upload", "hmch_run",
                     "hmch_estimate_patterns", "hmch_num_patterns", "hmch_head_len", "hmch_num_heads", "hmch_total_pattern_len", "hmch_trace_len"):
            getattr(L, name).argtypes = [C.c_void_p] + ([C.c_char_p] if name == "hmch_load_phase" else [])
        L.hmch_set_panel_raw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p]
        L.hmch_set_options.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double] + [C.c_int] * 9
        L.hmch_set_device_mining.argtypes = [C.c_void_p, C.c_int]
        L.hmch_mined_on_device.argtypes = [C.c_void_p]
        L.hmch_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.hmch_get_panel.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.hmch_resolve_all.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.hmch_export_model.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        L.hmch_get_resolutions.argtypes = [C.c_void_p, C.c_void_p]
        L.hmch_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hmch_get_seconds.argtypes = [C.c_void_p, C.c_void_p]
        L.hmch_write_reconstructed.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class HaploModel:
    """HaploModel with the CLI defaults of HMC.cpp:35-47 (MV, min-freq-abs 1.5, max-pattern-len 30, sample-size 10, 1 iteration)."""

    DEFAULTS = dict(model="MV", min_freq=-1.0, min_freq_abs=1.5, num_patterns=-1, min_pattern_len=1, max_pattern_len=30, mc_order=1,
                    max_iteration=1, sample_size=10, exact_estimate=False, scaled=False, host_threads=0,
                    device_mining=-1)      # pattern search: 1 device kernels, 0 host threads, -1 device when a B200 context exists

    def __init__(self, device: int = 0, arena_bytes: int = 0, threads: int = 0, **opts):
        err = C.create_string_buffer(512)
        self._h = lib().hmch_create(device, arena_bytes, threads, err, 512)
        if not self._h:
            raise HmcError(err.value.decode())
        self._h = C.c_void_p(self._h)
        self.device_index = device
        self.opts = dict(self.DEFAULTS)
        self.set_options(**opts)
        self.N = self.L = self.Amax = 0

    def close(self):
        if getattr(self, "_h", None):
            lib().hmch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HmcError(lib().hmch_last_error(self._h).decode())

    def set_options(self, **opts):
        self.opts.update(opts)
        o = self.opts
        self._ck(lib().hmch_set_options(self._h, o["model"].encode(), o["min_freq"], o["min_freq_abs"], o["num_patterns"], o["min_pattern_len"],
                                        o["max_pattern_len"], o["mc_order"], o["max_iteration"], o["sample_size"], int(o["exact_estimate"]),
                                        int(o["scaled"]), o["host_threads"]))
        self._ck(lib().hmch_set_device_mining(self._h, int(o["device_mining"])))

    def mined_on_device(self) -> bool:
        return bool(lib().hmch_mined_on_device(self._h))

    def _dims(self):
        n, l, a = C.c_int(), C.c_int(), C.c_int()
        lib().hmch_dims(self._h, C.byref(n), C.byref(l), C.byref(a))
        self.N, self.L, self.Amax = n.value, l.value, a.value

    def load_phase(self, path: str):
        self._ck(lib().hmch_load_phase(self._h, path.encode()))
        self._dims()

    def set_panel(self, codes: np.ndarray, kind: str = "S"):
        """codes: [N][2][L] allele codes 0..A-1 (-1 missing), stored as the PHASE values ('0'+code for S loci, code+1 for M loci)."""
        g = np.asarray(codes)
        n, _, l = g.shape
        raw = np.where(g < 0, -1, g.astype(np.int32) + (ord("0") if kind == "S" else 1)).astype(np.int32)
        raw = np.ascontiguousarray(raw)
        self._ck(lib().hmch_set_panel_raw(self._h, n, l, (kind * l).encode(), _p(raw)))
        self._dims()

    def panel(self):
        idx = np.zeros((self.N, 2, self.L), np.int8)
        na = np.zeros(self.L, np.int32)
        af = np.zeros((self.L, max(self.Amax, 1)), np.float64)
        lib().hmch_get_panel(self._h, _p(idx), _p(na), _p(af))
        return idx, na, af

    def build(self):
        """HaploModel::build: mine the patterns and build the automaton (host threads)."""
        self._ck(lib().hmch_build(self._h))

    def upload(self):
        self._ck(lib().hmch_upload(self._h))

    def find_patterns(self):
        self._ck(lib().hmch_find_patterns(self._h))

    def estimate_patterns(self):
        self._ck(lib().hmch_estimate_patterns(self._h))

    def export_model(self) -> dict:
        L_ = lib()
        P = L_.hmch_num_patterns(self._h)
        tl = L_.hmch_total_pattern_len(self._h)
        H = L_.hmch_num_heads(self._h)
        A = max(self.Amax, 1)
        m = dict(P=P, head_len=L_.hmch_head_len(self._h), Amax=A, start=np.zeros(P, np.int32), end=np.zeros(P, np.int32), freq=np.zeros(P),
                 prefix=np.zeros(P), tp=np.zeros(P), succ=np.zeros((P, A), np.int32), al_off=np.zeros(P + 1, np.int64), al=np.zeros(tl, np.int8),
                 heads=np.zeros(H, np.int32))
        L_.hmch_export_model(self._h, _p(m["start"]), _p(m["end"]), _p(m["freq"]), _p(m["prefix"]), _p(m["tp"]), _p(m["succ"]), _p(m["al_off"]),
                             _p(m["al"]), _p(m["heads"]))
        return m

    def resolve_all(self) -> float:
        ll = C.c_double()
        self._ck(lib().hmch_resolve_all(self._h, C.byref(ll)))
        return ll.value

    def run(self):
        self._ck(lib().hmch_run(self._h))
        n = lib().hmch_trace_len(self._h)
        ll = np.zeros(n)
        npat = np.zeros(n, np.int32)
        lib().hmch_get_trace(self._h, _p(ll), _p(npat))
        return ll, npat

    def resolutions(self) -> np.ndarray:
        r = np.zeros((self.N, 2, self.L), np.int8)
        lib().hmch_get_resolutions(self._h, _p(r))
        return r

    def seconds(self) -> dict:
        s = np.zeros(3)
        lib().hmch_get_seconds(self._h, _p(s))
        return dict(mining=float(s[0]), device=float(s[1]), host_mstep=float(s[2]))

    def write_reconstructed(self, path: str, suffix: str = ".reconstructed"):
        self._ck(lib().hmch_write_reconstructed(self._h, path.encode(), suffix.encode()))

    def device(self) -> DeviceBuilder:
        """The underlying device context (owned by this model) as a DeviceBuilder view, for direct C-ABI calls."""
        h = lib().hmch_ctx(self._h)
        if not h:
            raise HmcError(lib().hmch_last_error(self._h).decode())
        d = DeviceBuilder.__new__(DeviceBuilder)
        d._h = C.c_void_p(h)
        d._borrowed = True
        d.N, d.L, d.Amax = self.N, self.L, max(self.Amax, 1)
        return d
