"""Regenerates tests/golden/likelihood.json: HaploBuilder::getLikelihood (HaploBuilder.cpp:128-151) of both haplotypes of every
resolution and of the genotype, computed by the UNMODIFIED reference (oracle/_ref/hmc_ref --dump-likelihood) on the committed golden
panels, first iteration.  Hex floats.  Run in the build container only.

    python tests/golden/make_likelihood.py
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refdump  # noqa: E402
from cases import CASES  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "hmc_ref")
NAMES = ["g1_clean", "g2_missing", "g3_multi", "g8_head2", "g9_k1"]     # single-pass models (the haplotypes are the committed reconstructions)


def main():
    assert os.path.exists(REF), "build the reference first: bash oracle/build_ref.sh"
    out = {}
    for name in NAMES:
        opts = [o for o in CASES[name][3] if not o.startswith("--dump")]
        with tempfile.TemporaryDirectory() as tmp:
            dump = os.path.join(tmp, "d.dump")
            subprocess.check_call([REF, os.path.join(HERE, name + ".phase"), "--dump", dump, "--out-suffix", "none", "--no-patterns", "--dump-likelihood", "-d", "1"] + opts,
                                  stderr=subprocess.DEVNULL)
            d = refdump.load(dump)
        out[name] = [[float(x).hex() for x in row] for row in d["iters"][0]["lk"]]
    with open(os.path.join(HERE, "likelihood.json"), "w") as f:
        json.dump(out, f, indent=0)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
