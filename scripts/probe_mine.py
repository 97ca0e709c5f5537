"""Device (K5) vs host-thread pattern mining on a BASELINE config: python scripts/probe_mine.py [C2] [iterations]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HMC_B200_HOST_TIMING"] = "1"
import numpy as np
from hmc_b200 import HaploModel
from hmc_b200.synth import make_config
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g, kind = make_config(cfg)
res = []
for dm in (1, 0):
    hm = HaploModel(max_iteration=iters, device_mining=dm)
    hm.set_panel(g, kind)
    t = time.time(); ll, npat = hm.run(); w = time.time() - t
    print(f"device_mining={dm}: {len(ll)} iterations in {w:.2f}s  LL {ll.tolist()}  patterns {npat.tolist()}  seconds {hm.seconds()}", flush=True)
    res.append((ll, npat, hm.resolutions()))
    hm.close()
print("identical:", np.array_equal(res[0][0], res[1][0]), np.array_equal(res[0][1], res[1][1]), np.array_equal(res[0][2], res[1][2]))
