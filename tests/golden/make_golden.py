"""Regenerates tests/golden/*.phase and *.dump.gz by running the UNMODIFIED reference
(oracle/_ref/hmc_ref, built by oracle/build_ref.sh from /root/reference).  Run in the build
container only; the committed outputs are what the tests read (the GPU box has no reference).

    python tests/golden/make_golden.py
"""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from hmc_b200.synth import make_panel, write_phase  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "hmc_ref")

from cases import CASES  # noqa: E402


def main():
    assert os.path.exists(REF), "build the reference first: bash oracle/build_ref.sh"
    for name, (kw, kind, half, opts, _) in CASES.items():
        g = make_panel(**kw)
        if half > 0:    # blank single alleles too (exercises the half-missing branches, HaploBuilder.cpp:62-77,159)
            rng = np.random.Generator(np.random.PCG64(kw["seed"] + 1000))
            m = rng.random(g.shape) < half
            g[m] = -1
        phase = os.path.join(HERE, name + ".phase")
        write_phase(phase, g, kind=kind)
        dump = os.path.join(HERE, name + ".dump")
        subprocess.check_call([REF, phase, "--dump", dump, "--out-suffix", ".reconstructed"] + opts)
        with open(dump, "rb") as fi, gzip.GzipFile(dump + ".gz", "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
        os.remove(dump)
        print(name, os.path.getsize(dump + ".gz"), "bytes")


if __name__ == "__main__":
    main()
