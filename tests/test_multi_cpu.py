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


def _worker_skewed(rank, world, port, nfits, q):
    """rank 0 holds all but one record (what a dynamic queue produces when the other rank draws one
    long fit): the slab must be sized from the data, not from nfits/world (round-1 4-GPU crash)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(range(nfits - 1)) if rank == 0 else [nfits - 1]
    allres = gd.gather_records([(i, _stub_result(i)) for i in mine], nfits)
    ok = all(r is not None and r["iter"] == 10 + i and np.array_equal(r["grad"], _stub_result(i)["grad"])
             for i, r in enumerate(allres))
    # a hole (nobody reports the last fit) must raise on every rank, not return None entries
    try:
        gd.gather_records([(i, _stub_result(i)) for i in mine if i != nfits - 1], nfits)
        raised = False
    except RuntimeError:
        raised = True
    # rank 1 holding nothing at all is fine too
    allres2 = gd.gather_records([(i, _stub_result(i)) for i in range(nfits)] if rank == 0 else [], nfits)
    ok = ok and all(r["nevals"] == 12 + i for i, r in enumerate(allres2))
    q.put((rank, len(mine), ok and raised))
    dist.barrier()
    dist.destroy_process_group()


def _worker_queue(rank, world, port, nfits, q):
    """Two successive dynamic queues in ONE process group: every index handed out exactly once per
    queue (the second queue must not see the first one's counter)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    store = dist.distributed_c10d._get_default_store()
    ok = True
    for rep in range(2):
        wq = gd.WorkQueue(nfits, store, start=world)
        mine = [rank]  # static first pick
        while True:
            i = wq.next()
            if i is None:
                break
            mine.append(i)
        allres = gd.gather_records([(i, _stub_result(i)) for i in mine], nfits)
        ok = ok and all(r["iter"] == 10 + i for i, r in enumerate(allres))
        dist.barrier()
    q.put((rank, len(mine), ok))
    dist.barrier()
    dist.destroy_process_group()


def _run(target, nfits, world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port, nfits, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_gather_skewed_split_world2_gloo():
    out = _run(_worker_skewed, 40, 2, 29573)
    assert sorted(n for _, n, _ in out) == [1, 39]
    assert all(ok for _, _, ok in out)


def test_dynamic_queue_twice_in_one_group_world2_gloo():
    out = _run(_worker_queue, 23, 2, 29575)
    assert all(ok for _, _, ok in out)


def test_gather_rejects_duplicates_and_holes_without_group():
    import pytest
    with pytest.raises(ValueError):
        gd.gather_records([(0, _stub_result(0)), (0, _stub_result(0))], 2)
    with pytest.raises(RuntimeError):
        gd.gather_records([(0, _stub_result(0))], 2)
    q = gd.WorkQueue(3)
    assert [q.next(), q.next(), q.next(), q.next()] == [0, 1, 2, None]


def test_gather_without_process_group_is_identity():
    res = gd.gather_records([(1, _stub_result(1)), (0, _stub_result(0))], 2)
    assert res[0]["iter"] == 10 and res[1]["iter"] == 11
