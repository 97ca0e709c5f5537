"""Where the exact M-step spends its time on C2 (HMC_B200_HOST_TIMING breakdown): python scripts/probe_exact.py [iterations]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HMC_B200_HOST_TIMING"] = "1"
from hmc_b200 import HaploModel
from hmc_b200.synth import make_config
g, kind = make_config("C2")
hm = HaploModel(max_iteration=int(sys.argv[1]) if len(sys.argv) > 1 else 2, exact_estimate=True)
hm.set_panel(g, kind)
t = time.time(); ll, npat = hm.run(); w = time.time() - t
print(f"exact EM: {len(ll)} iterations in {w:.2f}s  LL {ll.tolist()}  patterns {npat.tolist()}  seconds {hm.seconds()}", flush=True)
