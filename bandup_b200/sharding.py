"""SC-k-point sharding across GPUs (one process per GPU, torch.distributed for the plumbing).

The unit of work of the unfolding path is one SC k-point (loop `do i_SCKPT`, main_BandUP.f90:133):
its coefficient block is private and every pc-kpt folds onto exactly one SC-kpt
(band_unfolding_routines_mod.f90:317-324, 396-403), so ranks never exchange coefficients and the
delta_N rows they produce are disjoint.  The only collective of the path is the gather of those
rows (a few hundred rows x nener doubles, <= ~2 MB): NCCL all_gather over NVLink on GPUs, gloo in
the CPU tests.  Symmetry averaging over needed_dirs (band_unfolding_routines_mod.f90:840-853)
happens after the gather, on the host, unchanged.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def partition_kpts(costs: Sequence[float], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous blocks [start, end) of the SC-kpt list, one per rank, balanced by cost
    (n_bands * n_pw(K) * n_spinor).  Contiguity keeps the reader sequential in the wavefunction
    file (read_wavecar_mod.f90 skips k-points it does not need).  Every rank gets a block (possibly
    empty when there are fewer k-points than ranks); blocks tile 0..len(costs) exactly."""
    n = len(costs)
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    c = np.asarray(costs, dtype=np.float64)
    if n and np.any(c < 0):
        raise ValueError("costs must be non-negative")
    total = float(c.sum())
    if total <= 0.0:
        c = np.ones(n)
        total = float(n)
    bounds = [0]
    prefix = np.concatenate([[0.0], np.cumsum(c)])
    for r in range(1, world_size):
        target = total * r / world_size
        lo = bounds[-1]
        j = lo + int(np.argmin(np.abs(prefix[lo:] - target)))   # closest cut at or after the previous one
        bounds.append(j)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class KptTickets:
    """Hands out SC-kpt indices 0..n_kpts-1 to whoever asks first: a counter in the
    torch.distributed key-value store (the rendezvous store every rank already shares), so that a
    rank whose host->device link is slower -- on the 8-GPU boxes measured here four GPUs get 23 GB/s
    and four 35 GB/s when all copy at once -- simply ends up with fewer k-points instead of holding
    the job back.  The output does not depend on who computed what: rows are gathered by row id.
    Without a store (single process) it is a local counter."""

    def __init__(self, n_kpts: int, store=None, job: str = "0", offset: int = 0, stride: int = 1):
        self.n_kpts = int(n_kpts)
        self.store = store
        # without a store: offset, offset+stride, ... (round-robin over ranks when there are several)
        self._local, self._stride = int(offset), max(1, int(stride))
        self.key = f"bandup_b200/next_kpt/{job}"
        if store is not None:
            # A counter is never reset, so every use of a job name gets a counter of its own: each
            # rank counts how many times IT has opened `job` (a key only it increments) and all
            # ranks open their jobs in the same order, so they agree on the epoch without talking.
            epoch = int(store.add(f"bandup_b200/epoch/{job}/{int(offset)}", 1))
            self.key = f"bandup_b200/next_kpt/{job}/{epoch}"
            store.add(self.key, 0)        # creates the counter; every rank may do it

    @classmethod
    def from_default_group(cls, n_kpts: int, job: str = "0") -> "KptTickets":
        """Counter in the store of the default process group.  One process: a local counter.  Several
        processes whose store cannot be reached: static round-robin (rank, rank + world, ...)."""
        store, rank, world = None, 0, 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                rank, world = dist.get_rank(), dist.get_world_size()
                from torch.distributed import distributed_c10d as c10d
                store = c10d._get_default_store()
        except Exception:
            store = None
        return cls(n_kpts, store, job, offset=rank, stride=world)

    def take(self) -> Optional[int]:
        if self.store is None:
            i = self._local
            self._local += self._stride
        else:
            i = int(self.store.add(self.key, 1)) - 1
        return i if i < self.n_kpts else None


@dataclass
class GatheredDeltaN:
    dN: np.ndarray            # (n_rows_total, nener) in global row order
    row_ids: np.ndarray       # (n_rows_total,) the global pc-kpt row id of every row
    rows_per_rank: List[int]


def gather_delta_N(local_dN: np.ndarray, local_row_ids: np.ndarray, nener: int, device=None,
                   group=None) -> GatheredDeltaN:
    """All ranks contribute their (n_local, nener) delta_N rows + global row ids; every rank gets
    the full table sorted by row id.  Rows are disjoint by construction; a duplicate id is an
    error (it would mean a pc-kpt was unfolded twice).  Pads to the largest shard so that one
    all_gather moves everything."""
    import torch
    import torch.distributed as dist
    local_dN = np.ascontiguousarray(local_dN, dtype=np.float64).reshape(-1, nener)
    ids = np.ascontiguousarray(local_row_ids, dtype=np.int64).reshape(-1)
    if ids.shape[0] != local_dN.shape[0]:
        raise ValueError("one row id per delta_N row")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        order = np.argsort(ids, kind="stable")
        _check_disjoint(ids[order])
        return GatheredDeltaN(local_dN[order], ids[order], [int(ids.shape[0])])
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device("cpu")
    n_local = torch.tensor([ids.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(max(counts), 1)
    # one payload: [row id (as f64, exact below 2^53), nener values] per row, padded
    payload = torch.zeros((n_max, nener + 1), dtype=torch.float64, device=dev)
    if ids.shape[0]:
        payload[:ids.shape[0], 0] = torch.from_numpy(ids.astype(np.float64)).to(dev)
        payload[:ids.shape[0], 1:] = torch.from_numpy(local_dN).to(dev)
    out = torch.empty((world, n_max, nener + 1), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out.view(world * n_max, nener + 1), payload, group=group)
    out = out.cpu().numpy()
    rows, rid = [], []
    for r in range(world):
        rows.append(out[r, :counts[r], 1:])
        rid.append(out[r, :counts[r], 0].astype(np.int64))
    dN = np.concatenate(rows, axis=0) if rows else np.zeros((0, nener))
    rid = np.concatenate(rid) if rid else np.zeros(0, dtype=np.int64)
    order = np.argsort(rid, kind="stable")
    _check_disjoint(rid[order])
    return GatheredDeltaN(dN[order], rid[order], counts)


def _check_disjoint(sorted_ids: np.ndarray) -> None:
    if sorted_ids.size > 1 and np.any(np.diff(sorted_ids) == 0):
        dup = sorted_ids[1:][np.diff(sorted_ids) == 0][0]
        raise ValueError(f"pc-kpt row {int(dup)} was produced by more than one SC-kpt / rank")


class ShardedUnfoldJob:
    """Runs this rank's block of SC k-points through an unfolder with double buffering and gathers
    the delta_N rows.  `unfolder` is anything with submit(ikpt, wf, g0=/folding_G=) / fetch(ikpt)
    (bandup_b200.unfold.GpuUnfolder on a GPU box; a stand-in in the CPU tests).

    load_kpt(iK) -> (wf, folding_G, row_ids): the wavefunction of SC-kpt iK (or a callable
    (unfolder, iK, folding_G) that submits it), the Cartesian folding vectors of the pc-kpts that
    fold onto it and their global row ids in the output table."""

    def __init__(self, unfolder, n_kpts: int, costs: Sequence[float], load_kpt: Callable,
                 rank: int = 0, world_size: int = 1, depth: int = 2,
                 after_fetch: Optional[Callable] = None, prefetch: bool = False,
                 tickets: Optional[KptTickets] = None):
        self.unfolder = unfolder
        # tickets: dynamic assignment (KptTickets shared by all ranks) instead of the static
        # contiguous blocks; the k-points this rank ended up with are in self.done_kpts
        self.tickets = tickets
        self.done_kpts: List[int] = []
        self.load_kpt = load_kpt
        # after_fetch(iK, result, row_ids): called right after a k-point is fetched, while it is
        # still resident on the device (rho, Pauli vectors, spectral function can be asked for)
        self.after_fetch = after_fetch
        # prefetch: load_kpt(iK+1) runs on a helper thread while iK is submitted / older k-points are
        # fetched (load_kpt must then not touch the unfolder)
        self.prefetch = prefetch
        self.rank, self.world_size = rank, world_size
        self.blocks = partition_kpts(costs, world_size)
        self.my_block = self.blocks[rank]
        self.depth = max(1, depth)
        if len(costs) != n_kpts:
            raise ValueError("one cost per SC-kpt")

    def run_local(self):
        """(dN rows, row ids, weights per row) of this rank's SC-kpts; submit runs `depth` ahead of
        fetch so the upload of k-point j+1 overlaps the kernels of j."""
        if self.tickets is not None:
            source = iter(self.tickets.take, None)
        else:
            source = iter(range(*self.my_block))
        inflight: List[Tuple[int, np.ndarray]] = []
        rows, ids, weights = [], [], []
        self.done_kpts = []

        def drain_one():
            iK, rid = inflight.pop(0)
            res = self.unfolder.fetch(iK)
            if np.any(res.n_selected == 0):
                bad = rid[np.nonzero(res.n_selected == 0)[0][0]]
                # main_BandUP.f90:167-171 (stop_if_pckpt_cannot_be_parsed = .TRUE.)
                raise RuntimeError(f"pc-kpt row {int(bad)} cannot be parsed (no plane wave selected)")
            rows.append(res.dN)
            weights.append(res.spectral_weight)
            ids.append(rid)
            if self.after_fetch is not None:
                self.after_fetch(iK, res, rid)

        pool = ThreadPoolExecutor(1) if self.prefetch else None
        iK = next(source, None)
        ahead = pool.submit(self.load_kpt, iK) if pool and iK is not None else None
        while iK is not None:
            nxt = next(source, None)      # with tickets: this rank holds one k-point ahead
            if pool:
                wf, folding_G, rid = ahead.result()
                # the next one goes into the buffer of iK-2, fetched two iterations ago (depth 2)
                ahead = pool.submit(self.load_kpt, nxt) if nxt is not None else None
            else:
                wf, folding_G, rid = self.load_kpt(iK)
            if len(inflight) >= self.depth:
                drain_one()
            if callable(wf):      # the loader submits by itself (e.g. raw file records, device-made G list)
                wf(self.unfolder, iK, folding_G)
            else:
                self.unfolder.submit(iK, wf, folding_G=folding_G)
            inflight.append((iK, np.asarray(rid, dtype=np.int64)))
            self.done_kpts.append(iK)
            iK = nxt
        while inflight:
            drain_one()
        if pool:
            pool.shutdown()
        nener = self.unfolder.nener
        dN = np.concatenate(rows, axis=0) if rows else np.zeros((0, nener))
        rid = np.concatenate(ids) if ids else np.zeros(0, dtype=np.int64)
        return dN, rid, weights

    def run(self, device=None, group=None) -> GatheredDeltaN:
        dN, rid, _ = self.run_local()
        return gather_delta_N(dN, rid, self.unfolder.nener, device=device, group=group)
