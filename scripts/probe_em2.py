import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmc_b200 import HaploModel
from hmc_b200.synth import make_config
g, kind = make_config("C2")
hm = HaploModel()
hm.set_panel(g, kind)
t = time.time(); hm.build(); print("build", time.time() - t)
for i in range(3):
    t = time.time(); ll = hm.resolve_all(); print("host resolve_all", time.time() - t, ll, hm.seconds(), flush=True)
t = time.time(); hm.find_patterns(); print("find_patterns (samples branch)", time.time() - t, flush=True)
t = time.time(); ll = hm.resolve_all(); print("host resolve_all after M-step", time.time() - t, ll, flush=True)
