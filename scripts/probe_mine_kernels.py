"""Steady-state device time of K5 (mining) + K6 (automaton) on a BASELINE config, genotype and sampling branch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmc_b200 import HaploModel, DeviceBuilder
from hmc_b200.synth import make_config
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
if cfg == "custom":     # custom N L alleles founders missing
    from hmc_b200.synth import make_panel
    n_, l_, a_, f_, m_ = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
    g, kind = make_panel(n_, l_, alleles=a_, founders=f_, missing=m_, seed=4), ("M" if a_ > 2 else "S")
else:
    g, kind = make_config(cfg)
hm = HaploModel(device_mining=0); hm.set_panel(g, kind)
t = time.time(); hm.build(); print(f"host threads: mining + automaton {time.time() - t:.3f} s", flush=True)
idx, na, af = hm.panel()
N, _, L = idx.shape
t = time.time(); dev = DeviceBuilder(); print(f"ctx create {time.time() - t:.3f} s", flush=True)
t = time.time(); dev.set_panel(idx, na, af); print(f"set_panel {time.time() - t:.3f} s", flush=True)
prev_aut = 0.0
min_freq = 1.5 / (2.0 * N)      # HaploModel.cpp:57-64 with the CLI default --min-freq-abs 1.5
for rep in range(3):
    t = time.time(); r = dev.mine_patterns(min_freq, 1, 30); w = time.time() - t
    st = dev.mine_stats(); st['ms'] -= prev_aut
    t = time.time(); a = dev.build_automaton(1); w2 = time.time() - t
    prev_aut = dev.mine_stats()['ms'] - st['ms']
    print(f"genotype branch rep {rep}: P={r['P']} mine+fetch wall {w:.3f} s, kernels {st['ms']:.1f} ms, nodes {st['nodes']}, max list {st['max_list']}; automaton wall {w2*1e3:.1f} ms (automaton kernels {prev_aut:.1f} ms)", flush=True)
m = hm.export_model()
assert r["P"] == m["P"] and np.array_equal(r["freq"], m["freq"]) and np.array_equal(r["start"], m["start"])
fa = dev.fetch_automaton(r["P"], a["n_heads"])
assert np.array_equal(fa["succ"], m["succ"]) and np.array_equal(fa["heads"], m["heads"])
res = dev.resolve_all(K=10)
hs = res["haps"].reshape(N * 10 * 2, L)
w = np.repeat(res["posterior"] / res["coverage"][:, None], 2, axis=1).reshape(-1)
keep = np.repeat(np.arange(10)[None, :] < res["nres"][:, None], 2, axis=1).reshape(-1)
hs, w = np.ascontiguousarray(hs[keep]), np.ascontiguousarray(w[keep])
tw = float(np.sum(w))
for rep in range(3):
    t = time.time(); r = dev.mine_patterns(min_freq, 1, 30, samples=hs, weights=w, total_weight=tw); wl = time.time() - t
    st = dev.mine_stats(); st['ms'] -= prev_aut
    t = time.time(); a = dev.build_automaton(1); w2 = time.time() - t
    prev_aut = dev.mine_stats()['ms'] - st['ms']
    print(f"sampling branch rep {rep}: n={len(w)} P={r['P']} mine+fetch wall {wl:.3f} s, kernels {st['ms']:.1f} ms, nodes {st['nodes']}, max list {st['max_list']}; automaton wall {w2*1e3:.1f} ms (automaton kernels {prev_aut:.1f} ms)", flush=True)
