"""Scratch perf probe (not the bench): model mined by the oracle port, kernels timed by CUDA events."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmc_b200 import DeviceBuilder
from hmc_b200.synth import make_panel
from oracle.pyoracle import Oracle

cases = [("clean 2000x300", dict(n_ind=2000, n_loci=300, founders=16, seed=2)),
         ("miss5 2000x300", dict(n_ind=2000, n_loci=300, founders=16, missing=0.05, seed=2))]
for name, kw in cases:
    g = make_panel(**kw)
    o = Oracle(g)
    t = time.time(); o.find_patterns(); m = o.export_model(); print(name, "mine %.1fs P=%d" % (time.time() - t, m["P"]), flush=True)
    nall, afreq = o.allele_table()
    for K in (10, 1):
        for nt in (32, 64, 128, 256):
            dev = DeviceBuilder(threads=nt)
            dev.set_panel(g, nall, afreq); dev.set_model(m)
            dev.resolve_all_device(K=K)
            t = time.time(); dev.resolve_all_device(K=K); wall = time.time() - t
            ms = dev.kernel_ms(); c = dev.counters()
            steps = c["steps"]
            print(f"  K={K} nt={nt}: fwd {ms['forward']:.2f} ms  tb {ms['traceback']:.2f} ms  wall {wall*1e3:.1f} ms  steps/s(fwd) {steps/ms['forward']*1e3:.3e}  "
                  f"mean S {c['states']/steps:.1f} X/step {c['arrivals']/steps:.1f} max {c['max_states']} spill {c['spill_steps']} waves {c['waves']}", flush=True)
            dev.close()
