"""Per-individual tail of the forward kernel on a C2-like panel: python scripts/probe_tail.py N L missing K [threads]
Writes gpurun_out/tail_<N>x<L>.npy = [n][9] {individual, cycles, states, max states, arrivals, spill steps, start ns, end ns, CTA}."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
N, L, miss, K = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
nt = int(sys.argv[5]) if len(sys.argv) > 5 else 0
os.environ["HMC_B200_DEBUG_CYCLES"] = "1"
from hmc_b200 import HaploModel, DeviceBuilder
from hmc_b200.device import lib
from hmc_b200.synth import make_panel
g = make_panel(N, L, founders=16, missing=miss, seed=2)
hm = HaploModel(); hm.set_panel(g, "S"); t = time.time(); hm.build(); m = hm.export_model(); print("build %.1fs P=%d" % (time.time() - t, m["P"]), flush=True)
idx, na, af = hm.panel()
dev = DeviceBuilder(threads=nt); dev.set_panel(idx, na, af); dev.set_model(m)
cnt = int(os.environ.get('COUNT', N))
dev.resolve_all_device(K=K, count=cnt)
dev.resolve_all_device(K=K, count=cnt)
ms = dev.kernel_ms(); c = dev.counters()
print(f"K={K} nt={nt}: fwd {ms['forward']:.2f} ms tb {ms['traceback']:.2f} ms {c}", flush=True)
d = np.zeros((N, 8), np.int64); inds = np.zeros(N, np.int32)
n = lib().hmc_debug_cycles(dev._h, d.ctypes.data_as(C.c_void_p), inds.ctypes.data_as(C.c_void_p), N)
d = d[:n]; inds = inds[:n]
t0 = d[:, 5].min(); span = (d[:, 6].max() - t0) / 1e6
dur = (d[:, 6] - d[:, 5]) / 1e6
ctas = int(d[:, 7].max()) + 1
print(f"n={n} CTAs={ctas} span {span:.2f} ms  sum of individual ms {dur.sum():.0f}  sum/CTAs {dur.sum()/ctas:.2f} ms")
print(f"per-individual ms: mean {dur.mean():.2f} p50 {np.percentile(dur,50):.2f} p90 {np.percentile(dur,90):.2f} p99 {np.percentile(dur,99):.2f} max {dur.max():.2f}")
o = np.argsort(-dur)[:15]
for q in o:
    print(f"  ind {inds[q]:5d} wave pos {q:5d}: {dur[q]:7.2f} ms start {(d[q,5]-t0)/1e6:7.2f} end {(d[q,6]-t0)/1e6:7.2f} states {d[q,1]:9d} max {d[q,2]:6d} arrivals {d[q,3]:9d} spill steps {d[q,4]:4d}  ns/state {dur[q]*1e6/d[q,1]:.1f}")
# when does each CTA finish?
fin = np.zeros(ctas)
for q in range(n): fin[d[q, 7]] = max(fin[d[q, 7]], (d[q, 6] - t0) / 1e6)
print("CTA finish ms: min %.2f p10 %.2f p50 %.2f p90 %.2f max %.2f" % (fin.min(), np.percentile(fin, 10), np.percentile(fin, 50), np.percentile(fin, 90), fin.max()))
# time per state as a function of the individual's size
st = d[:, 1].astype(float)
for lo, hi in ((0, 5e4), (5e4, 1e5), (1e5, 2e5), (2e5, 5e5), (5e5, 1e6), (1e6, 1e9)):
    sel = (st >= lo) & (st < hi)
    if sel.any(): print(f"  individuals with {lo:.0e}..{hi:.0e} states: {sel.sum():5d}  share of states {st[sel].sum()/st.sum():.3f} share of time {dur[sel].sum()/dur.sum():.3f} ns/state {dur[sel].sum()*1e6/st[sel].sum():.1f}")
os.makedirs("gpurun_out", exist_ok=True)
np.save(f"gpurun_out/tail_{N}x{L}.npy", np.concatenate([inds[:, None].astype(np.int64), d], axis=1))
