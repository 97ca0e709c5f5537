"""-m gpu: the sharded EM must reproduce the single-process run: same LL trace, same pattern counts, same resolutions.
  gloo, two ranks sharing cuda:0 (independent kernels, no inter-kernel waiting): host-staged exchanges and device-side exchanges
  nccl, two ranks on two GPUs (skipped on a one-GPU box): device-side exchanges -- in-place all-gather of the result rows,
        in-place all-reduce of the raw count vector"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, q, backend="gloo", device_collectives=False):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests"), os.path.join(root, "tests", "golden")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from hmc_b200 import HaploModel
    from hmc_b200.parallel import attach
    from util import GOLDEN, model_opts
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import traceback
    try:
        _work(rank, world, port, name, q, backend, device_collectives)
    except Exception:
        q.put((rank, "error", traceback.format_exc(), b""))


def _work(rank, world, port, name, q, backend, device_collectives):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist
    from hmc_b200 import HaploModel
    from hmc_b200.parallel import attach
    from util import GOLDEN, model_opts
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    hm = HaploModel(device=dev, **model_opts(name))
    hm.load_phase(os.path.join(GOLDEN, name + ".phase"))
    attach(hm, device_collectives=device_collectives)
    ll, npat = hm.run()
    q.put((rank, ll.tolist(), npat.tolist(), hm.resolutions().tobytes()))
    hm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["gloo-host", "gloo-device", "nccl-device"])
@pytest.mark.parametrize("name", ["g4_em", "g5_exact"])
def test_two_ranks_equal_one(name, mode):
    import torch
    backend, device_collectives = mode.split("-")[0], mode.endswith("device")
    if backend == "nccl" and torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from hmc_b200 import HaploModel
    from util import GOLDEN, model_opts, rel_close
    hm = HaploModel(device=0, **model_opts(name))
    hm.load_phase(os.path.join(GOLDEN, name + ".phase"))
    ll1, np1 = hm.run()
    res1 = hm.resolutions().tobytes()
    hm.close()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q, backend, device_collectives)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=420) for _ in range(world))
    assert all(o[1] != "error" for o in out), [o[2] for o in out if o[1] == "error"]
    for p in procs:
        p.join(timeout=60)
    assert out[0][1:] == out[1][1:]                        # every rank holds the same trace and the same resolutions
    assert out[0][2] == np1.tolist()
    if model_opts(name).get("exact_estimate"):
        assert rel_close(out[0][1], ll1, 1e-12)             # fp64 atomics: counts agree to rounding, not to the bit
    else:
        assert out[0][1] == ll1.tolist() and out[0][3] == res1
