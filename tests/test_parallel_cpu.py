"""CPU, world_size 2 over gloo: the N>1 host logic -- shard ranges, the count all-reduce identity and the result all-gather --
with the oracle port standing in for the device on each rank (tests may use the oracle; the product never does)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from hmc_b200.parallel import shard_range


def test_shard_ranges_cover_and_balance():
    for n in (1, 7, 40, 2000):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests"), os.path.join(root, "tests", "golden")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from hmc_b200.parallel import allgather_blocks, allreduce_sum_, shard_range
    from hmc_b200.device import lib as dev_lib
    from oracle.pyoracle import Oracle
    from util import golden
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = golden("g2_missing")
    N = d["N"]
    full = Oracle(d["geno"])
    full.find_patterns()
    full.resolve_all()
    m = full.export_model()
    nall, afreq = full.allele_table()
    first, count = shard_range(N, rank, world)
    # this rank's individuals only, with the full panel's allele table and model
    sh = Oracle(d["geno"][first:first + count], sample_size=10)
    sh.set_allele_table(nall, afreq)
    sh.import_model(m)
    sh.resolve_all()
    # (1) result all-gather: P(G) of every individual, in individual order, on every rank
    pg = np.array([sh.individual(i)["pg"] for i in range(count)])
    allpg = allgather_blocks(pg.view(np.uint8).reshape(count, 8), N, rank, world).view(np.float64).ravel()
    ref_pg = np.array([full.individual(i)["pg"] for i in range(N)])
    ok_gather = bool(np.array_equal(allpg, ref_pg))
    ll = float(np.sum(np.log(allpg)))
    # (2) count all-reduce: raw shard sums add up to the unsharded E-pass (normalised by the FULL N)
    f, pf, _ = sh.estep(m["start"], m["al_off"], m["al"])       # normalised by the shard size: undo to get raw sums
    raw = np.concatenate([f * count, pf * count])
    allreduce_sum_(raw)
    P = m["P"]
    import ctypes as C
    fr, pr, tp = raw[:P].copy(), raw[P:].copy(), np.zeros(P)
    dev_lib().hmc_normalise_counts(P, N, fr.ctypes.data_as(C.c_void_p), pr.ctypes.data_as(C.c_void_p), tp.ctypes.data_as(C.c_void_p))
    ef, epf, etp = full.estep(m["start"], m["al_off"], m["al"])
    # the shard oracle clamps at its own N before we un-normalise; patterns that hit the clamp are compared after re-clamping
    close = lambda a, b: bool(np.all(np.abs(a - b) <= 1e-9 * np.maximum(np.abs(a), np.abs(b)) + 1e-300))
    q.put((rank, ok_gather, ll, close(fr, ef), close(pr, epf), close(tp, etp)))
    dist.destroy_process_group()


def test_two_rank_exchange_identities():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    for rank, ok_gather, ll, cf, cp, ct in res:
        assert ok_gather, rank
        assert cf and cp and ct, rank
    assert res[0][2] == res[1][2]           # LL identical on every rank (summed in individual order)
