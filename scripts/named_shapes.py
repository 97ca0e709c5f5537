"""BASELINE.json configs[2..4] at their NAMED shapes on one GPU (or the largest stated fraction that fits): model mined on the device from
the whole panel, one resolveAll pass timed, the oracle port timed on a few individuals against the same model (and compared bit for
bit) as the reference-subsample baseline.

  python scripts/named_shapes.py C4            10,000 x 2,000, 6 alleles, 1 % missing, defaults (scaled: fp64 underflows at L = 2,000)
  python scripts/named_shapes.py C3 [N]        50,000 x 5,000, 32 founders, --max-pattern-len 100, scaled  (N individuals of it, default all)
  python scripts/named_shapes.py C5 N          N x 10,000, 32 founders, defaults, scaled; through the binary panel ingest (HMCB rows)
Prints one JSON line per config.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmc_b200 import FLAG_SCALED, HaploModel
from hmc_b200.synth import make_panel

name = sys.argv[1]
if name == "C4":
    N, L, kw, kind, opts = 10000, 2000, dict(alleles=6, founders=16, missing=0.01, seed=4), "M", dict(scaled=True)
elif name == "C3":
    N, L, kw, kind, opts = 50000, 5000, dict(founders=32, seed=3), "S", dict(scaled=True, max_pattern_len=100)
else:
    N, L, kw, kind, opts = 1000, 10000, dict(founders=32, seed=5), "S", dict(scaled=True)
if len(sys.argv) > 2:
    N = int(sys.argv[2])
K = 10
t = time.time()
g = make_panel(N, L, **kw)
t_gen = time.time() - t
hm = HaploModel(**opts)
hm.set_panel(g, kind)
t = time.time(); hm.build(); t_build = time.time() - t
hm.upload()
dev = hm.device()
P = int(__import__("hmc_b200.host", fromlist=["lib"]).lib().hmch_num_patterns(hm._h))
print(f"{name}: {N} x {L} generated in {t_gen:.1f} s, model {P} patterns in {t_build:.1f} s (device mining: {hm.mined_on_device()})", file=sys.stderr, flush=True)
dev.resolve_all_device(K=K, flags=FLAG_SCALED)
t = time.time(); dev.resolve_all_device(K=K, flags=FLAG_SCALED); wall = time.time() - t
c = dev.counters(); ms = dev.kernel_ms()
r = dev.fetch_results(haps=False)
out = dict(config=name, individuals=N, loci=L, options=opts, patterns=P, model_build_s=t_build, pass_s=wall, steps=c["steps"], steps_per_s=c["steps"] / wall,
           mean_states=c["states"] / max(c["steps"], 1), max_states=c["max_states"], waves=c["waves"], kernel_ms=ms, ll=r["ll"], unresolved=int((r["nres"] == 0).sum()))
# reference-subsample baseline: the oracle port on a few individuals with the full model (needs the model on the host: skipped when it is huge)
if P <= 60_000_000 and not os.environ.get("HMC_SKIP_ORACLE"):
    from oracle.pyoracle import Oracle
    m = hm.export_model(); idx, na, af = hm.panel()
    sel = list(range(0, N, max(1, N // 6)))[:6]
    o = Oracle(np.ascontiguousarray(idx[sel]), sample_size=K, scaled=True, max_pattern_len=opts.get("max_pattern_len", 30))
    o.set_allele_table(na, af); o.import_model(m, min_len=1, max_len=opts.get("max_pattern_len", 30), min_freq=0.0)
    t = time.time(); o.resolve_all(); secs = time.time() - t
    rr = dev.fetch_results(haps=False)
    ok = all(o.individual(q)["pg"] == rr["pg"][i] and np.array_equal(o.individual(q)["prior"], rr["prior"][i, :o.individual(q)["n"]]) for q, i in enumerate(sel))
    out["cpu_baseline"] = dict(kind="port", cores=1, sample=f"{len(sel)} individuals of the panel, full model", steps_per_s=o.counters()["steps"] / secs, seconds=secs, parity_checked=bool(ok))
    assert ok
print(json.dumps(out), flush=True)
hm.close()
