"""One resolve_all on a synthetic panel, for ncu: python scripts/probe_one.py N L missing threads K [founders]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmc_b200 import DeviceBuilder
from hmc_b200.synth import make_panel
from oracle.pyoracle import Oracle
N, L, miss, nt, K = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
F = int(sys.argv[6]) if len(sys.argv) > 6 else 16
g = make_panel(N, L, founders=F, missing=miss, seed=2)
o = Oracle(g); o.find_patterns(); m = o.export_model(); nall, afreq = o.allele_table()
dev = DeviceBuilder(threads=nt); dev.set_panel(g, nall, afreq); dev.set_model(m)
for _ in range(2):
    dev.resolve_all_device(K=K)
print(dev.kernel_ms(), dev.counters())
