import sys, os, time, ctypes as C
os.environ["HMC_B200_DEBUG_CYCLES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmc_b200 import DeviceBuilder
from hmc_b200.device import lib
from hmc_b200.synth import make_panel
from oracle.pyoracle import Oracle
N, L, miss, K = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
g = make_panel(N, L, founders=16, missing=miss, seed=2)
o = Oracle(g); o.find_patterns(); m = o.export_model(); nall, afreq = o.allele_table()
for nt in (32, 128):
  for n_use in (N, N // 4):
    dev = DeviceBuilder(threads=nt); dev.set_panel(g, nall, afreq); dev.set_model(m)
    dev.resolve_all_device(K=K, count=n_use); dev.resolve_all_device(K=K, count=n_use)
    cyc = np.zeros(N, np.int64); inds = np.zeros(N, np.int32)
    n = lib().hmc_debug_cycles(dev._h, cyc.ctypes.data_as(C.c_void_p), inds.ctypes.data_as(C.c_void_p), N)
    cyc = cyc[:n] / 1.9e3  # us at ~1.9 GHz
    print(f"nt={nt} n={n_use} fwd {dev.kernel_ms()['forward']:.2f} ms; per-individual us: mean {cyc.mean():.0f} p50 {np.percentile(cyc,50):.0f} p90 {np.percentile(cyc,90):.0f} p99 {np.percentile(cyc,99):.0f} max {cyc.max():.0f}  sum/slots {cyc.sum()/1e3/min(n_use, 148*(16 if nt==32 else 4)):.2f} ms")
    dev.close()
