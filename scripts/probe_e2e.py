import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hmc_b200 import HaploModel, DeviceBuilder
from hmc_b200.synth import make_config
g, kind = make_config("C2")
hm = HaploModel(); hm.set_panel(g, kind); hm.build(); m = hm.export_model(); idx, na, af = hm.panel()
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep = []
hidx = pin(idx); keep.append(hidx[0]); idxp = hidx[1]
mp = dict(m)
for k in ("start", "end", "freq", "tp", "succ", "al_off", "al", "heads"):
    t, mp[k] = pin(m[k]); keep.append(t)
dev = DeviceBuilder()
for it in range(3):
    t0 = time.perf_counter(); dev.set_panel(idxp, na, af); t1 = time.perf_counter(); dev.set_model(mp); t2 = time.perf_counter()
    dev.resolve_all_device(K=10); t3 = time.perf_counter(); r = dev.fetch_results(); t4 = time.perf_counter()
    print(f"set_panel {1e3*(t1-t0):.1f} ms  set_model {1e3*(t2-t1):.1f} ms  resolve_device {1e3*(t3-t2):.1f} ms  fetch {1e3*(t4-t3):.1f} ms", flush=True)
