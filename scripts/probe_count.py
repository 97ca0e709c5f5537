"""One resolve pass over the first COUNT individuals of a synthetic panel: python scripts/probe_count.py N L missing threads K count"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N, L, miss, nt, K, cnt = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
from hmc_b200 import HaploModel, DeviceBuilder
from hmc_b200.synth import make_panel
g = make_panel(N, L, founders=16, missing=miss, seed=2)
hm = HaploModel(); hm.set_panel(g, "S"); hm.build(); m = hm.export_model()
idx, na, af = hm.panel()
dev = DeviceBuilder(threads=nt); dev.set_panel(idx, na, af); dev.set_model(m)
import os
reps = int(os.environ.get("REPS", "1"))
f, k = [], []
for i in range(1 + reps):
    dev.resolve_all_device(K=K, count=cnt)
    ms = dev.kernel_ms()
    if i: f.append(ms["forward"]); k.append(ms["traceback"])
c = dev.counters()
import statistics
print(f"K={K} nt={nt} count={cnt}: fwd {statistics.median(f):.2f} ms kbest {statistics.median(k):.2f} ms (median of {reps}; min {min(f):.2f} + {min(k):.2f}; pass median {statistics.median([a + b for a, b in zip(f, k)]):.2f}) {c}", flush=True)
