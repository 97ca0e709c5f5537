"""The weak-scaling panel of N GPUs on ONE GPU, block by block: python scripts/probe_shards.py N
Per block of 2,000 individuals: forward / K-best ms and the slowest individuals of the forward kernel (HMC_B200_DEBUG_CYCLES)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["HMC_B200_DEBUG_CYCLES"] = "1"
n_gpus = int(sys.argv[1])
import bench
from hmc_b200 import HaploModel
from hmc_b200.device import lib
g, kind, n_ind, n_loci = bench.make_workload(n_gpus)
hm = HaploModel(min_freq_abs=1.5 * n_gpus)
hm.set_panel(g, kind); hm.build(); hm.upload()
dev = hm.device()
for b in range(n_gpus):
    for _ in range(2):
        dev.resolve_all_device(K=10, first=b * 2000, count=2000)
    ms = dev.kernel_ms(); c = dev.counters()
    d = np.zeros((2000, 8), np.int64); inds = np.zeros(2000, np.int32)
    n = lib().hmc_debug_cycles(dev._h, d.ctypes.data_as(C.c_void_p), inds.ctypes.data_as(C.c_void_p), 2000)
    dur = (d[:n, 6] - d[:n, 5]) / 1e6
    o = np.argsort(-dur)[:4]
    print(f"block {b}: fwd {ms['forward']:.1f} ms kbest {ms['traceback']:.1f} ms; states {c['states']/c['steps']:.0f}/step max {c['max_states']}; sum of individual ms / 444 = {dur.sum()/444:.1f}; "
          f"slowest individuals (ms, max states): " + ", ".join(f"{dur[q]:.1f}/{d[q,2]}" for q in o), flush=True)
