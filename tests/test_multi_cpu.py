"""world_size=2 test of the multi-GPU plumbing on CPU (gloo): shard -> 'fit' -> one all-gather.
The per-fit work is replaced by a deterministic stub because the CUDA path needs a GPU; what is
covered is exactly what differs between N=1 and N>1: the partition and the record exchange."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from gpedm_b200 import dist as gd


def _stub_result(i):
    d = 1 + i % 4
    return dict(status=0, iter=10 + i, nevals=12 + i, nllpost=-float(i), lp=0.5 * i, SS=1.0 + i, logdet=-2.0 * i,
                pars=np.arange(d + 3) + i, parst=np.arange(d + 3) - i, grad=np.ones(d + 3) * i)


def _worker(rank, world, port, nfits, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = gd.shard_indices(nfits, world, rank)
    local = [(i, _stub_result(i)) for i in mine]
    allres = gd.gather_records(local, nfits)
    ok = all(r is not None and r["iter"] == 10 + i and np.array_equal(r["pars"], _stub_result(i)["pars"])
             for i, r in enumerate(allres))
    q.put((rank, len(mine), ok))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    nfits, world, port = 11, 2, 29571
    procs = [ctx.Process(target=_worker, args=(r, world, port, nfits, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(n for _, n, _ in out) == [5, 6]
    assert all(ok for _, _, ok in out)


def test_gather_without_process_group_is_identity():
    res = gd.gather_records([(1, _stub_result(1)), (0, _stub_result(0))], 2)
    assert res[0]["iter"] == 10 and res[1]["iter"] == 11
