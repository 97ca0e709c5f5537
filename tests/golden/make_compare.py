"""Regenerates tests/golden/compare.json: HaploComp results (switch error, IHP, IGP as hex floats) of the UNMODIFIED
reference (oracle/_ref/hmc_ref --compare, i.e. HMC::compare, HMC.cpp:246-257) for
  * every golden case: its .phase (the true phase, with its missing data) against its .phase.reconstructed;
  * synthetic pairs with planted switch / genotype errors and missing alleles (arrays stored in the fixture).
Run in the build container only.

    python tests/golden/make_compare.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from hmc_b200.synth import make_panel, write_phase  # noqa: E402
from cases import CASES  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "hmc_ref")


def ref_compare(infer_path, real_path):
    out = subprocess.check_output([REF, infer_path, "--compare", real_path, "-d", "0"]).decode()
    return json.loads(out.strip().splitlines()[-1])


def planted(seed, n, L, alleles, kind, p_switch, p_err, p_miss_real, p_miss_infer):
    rng = np.random.Generator(np.random.PCG64(seed))
    real = make_panel(n, L, alleles=alleles, founders=4, seed=seed)
    infer = real.copy()
    for i in range(n):
        rev = rng.random() < 0.5
        for l in range(L):
            if rng.random() < p_switch:
                rev = not rev
            if rev:
                infer[i, 0, l], infer[i, 1, l] = real[i, 1, l], real[i, 0, l]
    # genotype errors that stay consistent with HaploComp's contract: turn a HOMOZYGOUS real site into another homozygous call
    for i in range(n):
        for l in range(L):
            if real[i, 0, l] == real[i, 1, l] and rng.random() < p_err:
                infer[i, :, l] = (real[i, 0, l] + 1) % alleles
    real[rng.random(real.shape) < p_miss_real] = -1
    infer[rng.random(infer.shape) < p_miss_infer] = -1
    return real, infer


def main():
    assert os.path.exists(REF), "build the reference first: bash oracle/build_ref.sh"
    out = {"golden": {}, "planted": []}
    for name in CASES:
        real = os.path.join(HERE, name + ".phase")
        out["golden"][name] = ref_compare(real + ".reconstructed", real)
    # a genotype error makes the reference exit ("Inconsistent genotypes", Genotype.cpp:258-261): only the last pair plants them
    specs = [(1, 12, 40, 2, "S", 0.05, 0.0, 0.0, 0.0), (2, 16, 60, 2, "S", 0.10, 0.0, 0.05, 0.0), (3, 10, 50, 4, "M", 0.08, 0.0, 0.03, 0.02),
             (4, 8, 30, 2, "S", 0.0, 0.0, 0.0, 0.0), (5, 6, 25, 3, "M", 0.2, 0.0, 0.3, 0.1), (7, 9, 33, 2, "S", 0.5, 0.0, 0.0, 0.3),
             (8, 3, 1, 2, "S", 0.0, 0.0, 0.0, 0.0), (6, 5, 8, 2, "S", 0.0, 0.5, 0.0, 0.0)]
    with tempfile.TemporaryDirectory() as td:
        for spec in specs:
            seed, n, L, alleles, kind = spec[:5]
            real, infer = planted(*spec)
            pr, pi = os.path.join(td, "real.phase"), os.path.join(td, "infer.phase")
            write_phase(pr, real, kind=kind)
            write_phase(pi, infer, kind=kind)
            try:
                res = ref_compare(pi, pr)
            except subprocess.CalledProcessError:
                res = {"error": "inconsistent"}
            out["planted"].append({"spec": list(spec), "real": real.tolist(), "infer": infer.tolist(), "ref": res})
    with open(os.path.join(HERE, "compare.json"), "w") as f:
        json.dump(out, f)
    print({k: v for k, v in out["golden"].items()})
    print([p["ref"] for p in out["planted"]])


if __name__ == "__main__":
    main()
