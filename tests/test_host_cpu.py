"""CPU: the host layer (PHASE I/O, mining, automaton) against the reference dumps; the C-ABI library's exports."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cases import CASES
from hmc_b200.device import EXPORTS, library_path
from hmc_b200.host import HaploModel
from util import GOLDEN, golden, model_opts


@pytest.mark.parametrize("name", list(CASES))
def test_phase_reader_and_first_model(name):
    d = golden(name)
    hm = HaploModel(**model_opts(name))
    hm.load_phase(os.path.join(GOLDEN, name + ".phase"))
    idx, na, af = hm.panel()
    assert np.array_equal(idx, d["geno"]) and np.array_equal(na, d["nall"]) and np.array_equal(af, d["afreq"])
    hm.build()          # mining + automaton only: no device needed until the first resolve
    m, r = hm.export_model(), d["iters"][0]["patterns"]
    assert m["P"] == r["P"] and m["head_len"] == r["head_len"]
    for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads"):
        assert np.array_equal(m[k], r[k]), k
    assert np.array_equal(np.concatenate(r["al"]), m["al"])
    hm.close()


def test_mining_is_thread_count_independent():
    d = golden("g2_missing")
    ms = []
    for t in (1, 3, 8):
        hm = HaploModel(host_threads=t, device_mining=0)
        hm.load_phase(os.path.join(GOLDEN, "g2_missing.phase"))
        hm.build()
        ms.append(hm.export_model())
        hm.close()
    for m in ms[1:]:
        for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads", "al"):
            assert np.array_equal(m[k], ms[0][k]), k


def test_set_panel_from_codes_equals_phase_file():
    d = golden("g3_multi")
    hm = HaploModel()
    hm.set_panel(d["geno"], kind="M")       # codes == indices here only if every allele value occurs; compare through the tables
    idx, na, af = hm.panel()
    assert np.array_equal(na, d["nall"]) and np.array_equal(af, d["afreq"]) and np.array_equal(idx, d["geno"])
    hm.close()


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "hmc_b200.h")).read()
    declared = set(re.findall(r"\b(hmc_[a-z_]+)\s*\(", hdr))
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    L = C.CDLL(library_path())
    for name in declared:
        assert hasattr(L, name), name
    L.hmc_version.restype = C.c_char_p
    assert b"sm_100a" in L.hmc_version()


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hmc_b200 import DeviceBuilder, HmcError
    with pytest.raises(HmcError):
        DeviceBuilder()
    hm = HaploModel()
    hm.load_phase(os.path.join(GOLDEN, "g1_clean.phase"))
    hm.build()
    with pytest.raises(HmcError):
        hm.resolve_all()
    hm.close()


def test_normalise_counts_matches_reference_rule():
    from hmc_b200 import DeviceBuilder
    f = np.array([0.0, 3.0, 50.0, 2.0, 1.0])
    pf = np.array([0.0, 4.0, 45.0, 2.0, 0.5])
    nf, npf, tp = DeviceBuilder.normalise_counts(f, pf, 40)
    # HaploBuilder.cpp:317-331: clamp to N, freq <= prefix, /N, tp = freq/prefix (or freq/N), tp <= 1
    assert np.allclose(nf, [0, 3 / 40, 1.0, 2 / 40, 0.5 / 40]) and np.allclose(npf, [0, 0.1, 1.0, 0.05, 0.0125])
    assert np.allclose(tp, [0, 0.75, 1.0, 1.0, 1.0])


def test_phase_reader_variants(tmp_path):
    """HaploFile.cpp:54-153: optional 'P' line, blanks between the type characters, '-' / '?' as missing, 'M' loci with <= 0 as missing,
    ids starting with '#' or a digit."""
    p1 = tmp_path / "a.phase"
    p1.write_text("3\n4\nS S M M\n1abc\n0 1 12 7\n1 1 12 -1\nx\n? 0 7 7\n- 1 0 12\ny\n0 0 12 12\n1 0 7 7\n")
    hm = HaploModel()
    hm.load_phase(str(p1))
    idx, na, af = hm.panel()
    assert idx.shape == (3, 2, 4) and list(na) == [2, 2, 2, 2]
    assert idx[0].tolist() == [[0, 1, 1, 0], [1, 1, 1, -1]]         # symbols sorted: '0'<'1', 7<12
    assert idx[1].tolist() == [[-1, 0, 0, 0], [-1, 1, -1, 1]]       # '?', '-', 0 and -1 are missing
    assert abs(af[2, 0] - 2 / 5) < 1e-15 and abs(af[2, 1] - 3 / 5) < 1e-15
    hm.close()
    p2 = tmp_path / "b.phase"
    p2.write_text("2\n3\nP 10 20 30\nSSS\n#7\n0 1 0\n1 1 0\nq\n1 0 0\n1 0 1\n")
    hm = HaploModel()
    hm.load_phase(str(p2))
    idx2, na2, _ = hm.panel()
    assert idx2.tolist() == [[[0, 1, 0], [1, 1, 0]], [[1, 0, 0], [1, 0, 1]]] and list(na2) == [2, 2, 2]
    hm.close()


def test_binary_panel_round_trip(tmp_path):
    """PHASE text -> HMCB binary -> the same panel: allele tables, frequencies (bit-identical), indices, ids; and a model mined
    from the binary panel equals the one mined from the text panel."""
    import subprocess
    from hmc_b200.build import CLI
    for name in ("g2_missing", "g3_multi", "g8_head2"):
        phase = os.path.join(GOLDEN, name + ".phase")
        binp = str(tmp_path / (name + ".hmcb"))
        subprocess.run([CLI, "--to-binary", binp, phase, "-d", "0"], check=True)
        with open(binp, "rb") as f:
            assert f.read(8) == b"HMCB001\n"
        a, b = HaploModel(**{**model_opts(name), "device_mining": 0}), HaploModel(**{**model_opts(name), "device_mining": 0})
        a.load_phase(phase)
        b.load_phase(binp)
        for x, y in zip(a.panel(), b.panel()):
            assert np.array_equal(x, y)
        a.build(); b.build()
        ma, mb = a.export_model(), b.export_model()
        for k in ("start", "end", "freq", "prefix", "tp", "succ", "heads", "al"):
            assert np.array_equal(ma[k], mb[k]), k
        a.close(); b.close()
        assert os.path.getsize(binp) < os.path.getsize(phase)


def test_binary_panel_rejects_garbage(tmp_path):
    p = tmp_path / "bad.hmcb"
    p.write_bytes(b"HMCB001\n" + b"\x00" * 16)
    hm = HaploModel()
    with pytest.raises(Exception):
        hm.load_phase(str(p))
    hm.close()
