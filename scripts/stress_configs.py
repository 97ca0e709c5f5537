"""Other BASELINE configs at reduced size through the whole pipeline (host mining -> device): crash / capacity check + throughput."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hmc_b200 import HaploModel, FLAG_SCALED
from hmc_b200.synth import make_panel
cases = [
    ("C1 200x100 biallelic", dict(n_ind=200, n_loci=100, founders=8, seed=1), "S", dict(), 0),
    ("C4-like 3000x600, 6 alleles, 1% missing", dict(n_ind=3000, n_loci=600, alleles=6, founders=16, missing=0.01, seed=4), "M", dict(), 0),
    ("C3-like 8000x2500, 32 founders, max-pattern-len 100, scaled", dict(n_ind=8000, n_loci=2500, founders=32, seed=3), "S", dict(max_pattern_len=100, scaled=True), FLAG_SCALED),
]
for name, kw, kind, opts, flags in (cases if "--em-only" not in sys.argv else []):
    g = make_panel(**kw)
    hm = HaploModel(**opts); hm.set_panel(g, kind)
    t = time.time(); hm.build(); tb = time.time() - t
    hm.upload(); m = hm.export_model(); dev = hm.device()
    dev.resolve_all_device(K=10, flags=flags)
    t = time.time(); dev.resolve_all_device(K=10, flags=flags); w = time.time() - t
    c = dev.counters(); r = dev.fetch_results(haps=False)
    print(f"{name}: P={m['P']} build {tb:.1f}s | resolve {w*1e3:.1f} ms = {c['steps']/w:.3e} steps/s, mean S {c['states']/max(c['steps'],1):.1f} max {c['max_states']} waves {c['waves']} "
          f"LL {r['ll']:.3f} unresolved {(r['nres']==0).sum()}", flush=True)
    hm.close()

# the same panels through a 2-iteration sampling EM (device mining of the sampled haplotypes, device automaton, device-resident samples)
os.environ["HMC_B200_HOST_TIMING"] = "1"
for name, kw, kind, opts, flags in cases[1:]:
    g = make_panel(**kw)
    hm = HaploModel(max_iteration=2, **opts); hm.set_panel(g, kind)
    t = time.time(); ll, npat = hm.run(); w = time.time() - t
    print(f"{name}: EM {len(ll)} iterations in {w:.1f} s, LL {ll.tolist()}, patterns {npat.tolist()}, seconds {hm.seconds()}", flush=True)
    hm.close()
