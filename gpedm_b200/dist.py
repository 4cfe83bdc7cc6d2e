"""One-process-per-GPU plumbing for sharded batches of independent fits (torch.distributed:
NCCL on GPUs, gloo in CPU tests).  The data path has no collective: each rank fits its own
share of the grid; the only exchange is one all-gather of the fixed-width result records."""
import numpy as np

RECORD_WIDTH = 8 + 3 * 35  # status, iter, nevals, p, nll, lp, SS, logdet + pars/parst/grad (padded)


def shard_indices(nfits, world_size, rank, costs=None):
    """Static longest-processing-time assignment of fit indices to ranks (all ranks compute the same
    answer without communicating).  costs: per-fit estimated cost; default equal."""
    costs = np.ones(nfits) if costs is None else np.asarray(costs, dtype=float)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world_size)
    owner = np.empty(nfits, dtype=int)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return [int(i) for i in range(nfits) if owner[i] == rank]


def pack_record(index, res):
    rec = np.full(RECORD_WIDTH + 1, np.nan)
    rec[0] = index
    rec[1:9] = [res["status"], res["iter"], res["nevals"], len(res["pars"]), res["nllpost"], res["lp"], res["SS"],
                res["logdet"]]
    p = len(res["pars"])
    rec[9:9 + p] = res["pars"]
    rec[9 + 35:9 + 35 + p] = res["parst"]
    rec[9 + 70:9 + 70 + p] = res["grad"]
    return rec


def unpack_record(rec):
    p = int(rec[4])
    return int(rec[0]), dict(status=int(rec[1]), iter=int(rec[2]), nevals=int(rec[3]), nllpost=float(rec[5]),
                             lp=float(rec[6]), SS=float(rec[7]), logdet=float(rec[8]), pars=rec[9:9 + p].copy(),
                             parst=rec[9 + 35:9 + 35 + p].copy(), grad=rec[9 + 70:9 + 70 + p].copy())


def gather_records(local, nfits, device=None):
    """All-gather (one collective) of each rank's result records.  local: list of (index, result dict).
    Returns the list of nfits result dicts on every rank.  Works with any initialised
    torch.distributed backend; without an initialised process group it is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [None] * nfits
        for i, r in local:
            out[i] = r
        return out
    world = dist.get_world_size()
    slab = (nfits + world - 1) // world + 1
    buf = np.full((slab, RECORD_WIDTH + 1), np.nan)
    buf[:, 0] = -1
    for q, (i, r) in enumerate(local):
        buf[q] = pack_record(i, r)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    out = [None] * nfits
    for part in allt:
        for rec in part.cpu().numpy():
            if rec[0] >= 0:
                i, r = unpack_record(rec)
                out[i] = r
    return out
