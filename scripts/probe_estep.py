import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmc_b200 import HaploModel
from hmc_b200.synth import make_panel
N, L, miss, n_sub = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
g = make_panel(N, L, founders=16, missing=miss, seed=2)
hm = HaploModel(); hm.set_panel(g, "S"); hm.build(); hm.upload(); m = hm.export_model(); dev = hm.device()
r = dev.resolve_all(K=1, haps=False)
for n in (n_sub,):
    t = time.time(); f, pf = dev.estimate_counts(m["start"], m["al_off"], m["al"], r["pg"][:n], first=0, count=n); wall = time.time() - t
    ms = dev.kernel_ms(); c = dev.counters()
    print(f"E-pass n={n} P={m['P']}: wall {wall:.3f}s kernels fwd {ms['forward']:.1f} back {ms['backward']:.1f} estep {ms['estep']:.1f} ms  steps/s {c['steps']/wall:.3e}  waves {c['waves']}", flush=True)
