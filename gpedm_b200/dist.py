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
    """All-gather of each rank's result records.  local: list of (index, result dict).
    Returns the list of nfits result dicts on every rank.  Works with any initialised
    torch.distributed backend; without an initialised process group it is the identity.

    The ranks may hold ANY split of the fits (a dynamic work queue hands short fits to whichever rank
    is free, so one rank can finish far more than nfits/world of them): the per-rank record counts are
    exchanged first and the record slab is sized to the largest of them.  Every fit must be reported by
    exactly one rank; a missing or duplicated index raises instead of returning a hole."""
    import torch
    import torch.distributed as dist
    for i, _ in local:
        if not 0 <= i < nfits:
            raise ValueError("gather_records: fit index %d outside 0..%d" % (i, nfits - 1))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [None] * nfits
        for i, r in local:
            if out[i] is not None:
                raise ValueError("gather_records: fit %d reported twice" % i)
            out[i] = r
        _check_complete(out)
        return out
    world = dist.get_world_size()
    cnt = torch.tensor([len(local)], dtype=torch.int64)
    if device is not None:
        cnt = cnt.to(device)
    counts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    if sum(counts) != nfits:
        raise RuntimeError("gather_records: ranks hold %s records, expected %d in total" % (counts, nfits))
    slab = max(1, max(counts))
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
    for part, n in zip(allt, counts):
        for rec in part.cpu().numpy()[:n]:
            i, r = unpack_record(rec)
            if out[i] is not None:
                raise RuntimeError("gather_records: fit %d reported by two ranks" % i)
            out[i] = r
    _check_complete(out)
    return out


def _check_complete(out):
    missing = [i for i, r in enumerate(out) if r is None]
    if missing:
        raise RuntimeError("gather_records: no rank reported fits %s" % missing)


class WorkQueue:
    """Dynamic hand-out of fit indices 0..n-1 to the ranks of a process group through an atomic counter
    in the group's store (control plane only; no data-path collective).  Each instance uses its own
    store key (a per-process-group sequence number, identical on every rank because every rank
    constructs its queues in the same order), so a second grid in the same job starts from zero."""
    _seq = 0

    def __init__(self, n, store=None, start=0):
        import itertools
        import threading
        WorkQueue._seq += 1
        self.n, self.store, self.start = n, store, start
        self.key = "gpedm_queue_%d" % WorkQueue._seq
        self._local, self._lock = itertools.count(), threading.Lock()

    def next(self):
        """Next unclaimed index, or None when the queue is exhausted."""
        if self.store is not None:
            q = self.start + self.store.add(self.key, 1) - 1
        else:
            with self._lock:
                q = self.start + next(self._local)
        return q if q < self.n else None
