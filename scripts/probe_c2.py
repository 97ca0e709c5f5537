"""Perf probe on a C2-like panel: python scripts/probe_c2.py N L missing K threads[,threads...] [cycles]"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
N, L, miss, K = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
nts = [int(x) for x in sys.argv[5].split(",")]
cycles = len(sys.argv) > 6
if cycles:
    os.environ["HMC_B200_DEBUG_CYCLES"] = "1"
from hmc_b200 import HaploModel, DeviceBuilder
from hmc_b200.device import lib
from hmc_b200.synth import make_panel
g = make_panel(N, L, founders=16, missing=miss, seed=2)
hm = HaploModel(); hm.set_panel(g, "S"); t = time.time(); hm.build(); m = hm.export_model(); print("build %.1fs P=%d" % (time.time() - t, m["P"]), flush=True)
idx, na, af = hm.panel()
for nt in nts:
    dev = DeviceBuilder(threads=nt); dev.set_panel(idx, na, af); dev.set_model(m)
    dev.resolve_all_device(K=K)
    t = time.time(); dev.resolve_all_device(K=K); wall = time.time() - t
    ms = dev.kernel_ms(); c = dev.counters(); steps = c["steps"]
    print(f"K={K} nt={nt}: fwd {ms['forward']:.2f} ms tb {ms['traceback']:.2f} ms wall {wall*1e3:.1f} ms steps/s {steps/wall:.3e} mean S {c['states']/steps:.1f} X {c['arrivals']/steps:.1f} max {c['max_states']} spill {c['spill_steps']} waves {c['waves']}", flush=True)
    if cycles:
        cyc = np.zeros(N, np.int64); inds = np.zeros(N, np.int32)
        n = lib().hmc_debug_cycles(dev._h, cyc.ctypes.data_as(C.c_void_p), inds.ctypes.data_as(C.c_void_p), N)
        cyc = cyc[:n] / 1.9e6
        print(f"   last wave n={n}: per-individual ms: mean {cyc.mean():.2f} p50 {np.percentile(cyc,50):.2f} p90 {np.percentile(cyc,90):.2f} p99 {np.percentile(cyc,99):.2f} max {cyc.max():.2f} sum {cyc.sum():.0f}", flush=True)
    dev.close()
