import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmc_b200 import HaploModel
from hmc_b200.synth import make_config
g, kind = make_config("C2")
for exact in (False, True):
    hm = HaploModel(max_iteration=3, exact_estimate=exact)
    hm.set_panel(g, kind)
    t = time.time(); ll, npat = hm.run(); w = time.time() - t
    print(f"exact={exact}: {len(ll)} iterations in {w:.2f}s  LL {ll.tolist()}  patterns {npat.tolist()}  seconds {hm.seconds()}", flush=True)
    hm.close()
