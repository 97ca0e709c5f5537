"""One resolve pass on a synthetic panel (for ncu captures): python scripts/probe_one.py N L missing threads K [passes]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N, L, miss, nt, K = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
passes = int(sys.argv[6]) if len(sys.argv) > 6 else 2
from hmc_b200 import HaploModel, DeviceBuilder
from hmc_b200.synth import make_panel
g = make_panel(N, L, founders=16, missing=miss, seed=2)
hm = HaploModel(); hm.set_panel(g, "S"); hm.build(); m = hm.export_model()
idx, na, af = hm.panel()
dev = DeviceBuilder(threads=nt); dev.set_panel(idx, na, af); dev.set_model(m)
for _ in range(passes):
    dev.resolve_all_device(K=K)
ms = dev.kernel_ms(); c = dev.counters()
print(f"K={K} nt={nt}: fwd {ms['forward']:.2f} ms kbest {ms['traceback']:.2f} ms {c}", flush=True)
