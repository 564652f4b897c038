"""Host-side multi-GPU logic on CPU: partitioning of the SC-kpt list and the gather of the disjoint
delta_N rows over a world_size-2 gloo group.  The per-k-point numbers come from a stand-in
unfolder backed by the oracle (the real GpuUnfolder needs a B200; the collective and the
bookkeeping are what is under test here)."""
import os
import socket

import numpy as np
import pytest

from bandup_b200 import synth
from bandup_b200.sharding import GatheredDeltaN, ShardedUnfoldJob, gather_delta_N, partition_kpts


def test_partition_tiles_and_balances():
    rng = np.random.default_rng(0)
    for n, w in [(138, 8), (30, 4), (5, 8), (1, 2), (0, 3), (100, 1)]:
        costs = rng.uniform(0.9, 1.1, n)
        blocks = partition_kpts(costs, w)
        assert len(blocks) == w and blocks[0][0] == 0 and blocks[-1][1] == n
        for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
            assert b == c and a <= b
        if n >= 4 * w:
            loads = [costs[a:b].sum() for a, b in blocks]
            assert max(loads) <= 1.0 / w * costs.sum() + 1.2 * costs.max()
    # heavily skewed costs still give every k-point to exactly one rank
    blocks = partition_kpts([100, 1, 1, 1, 1, 1], 3)
    assert sorted(i for a, b in blocks for i in range(a, b)) == list(range(6))
    with pytest.raises(ValueError):
        partition_kpts([1.0, -1.0], 2)


def test_gather_single_process_sorts_and_rejects_duplicates():
    dN = np.arange(12, dtype=float).reshape(4, 3)
    g = gather_delta_N(dN, np.array([7, 2, 9, 4]), 3)
    assert isinstance(g, GatheredDeltaN) and g.row_ids.tolist() == [2, 4, 7, 9]
    assert np.array_equal(g.dN[0], dN[1])
    with pytest.raises(ValueError):
        gather_delta_N(dN, np.array([1, 2, 2, 3]), 3)


class OracleUnfolder:
    """Stand-in with GpuUnfolder's submit/fetch surface; numbers from the CPU oracle."""

    def __init__(self, sc):
        from oracle import oracle as O
        self.O, self.sc = O, sc
        self.grid = O.real_seq(*sc.energy_window())
        self.nener = len(self.grid)
        self.pending = {}
        self.max_inflight = 0

    def submit(self, ikpt, wf, folding_G=None, g0=None):
        self.pending[ikpt] = (wf, np.atleast_2d(folding_G))
        self.max_inflight = max(self.max_inflight, len(self.pending))
        assert len(self.pending) <= 2, "more k-points in flight than the pipeline depth"

    def fetch(self, ikpt):
        from bandup_b200.unfold import UnfoldedKpt
        wf, fG = self.pending.pop(ikpt)
        w, dN, ns, _ = self.O.unfold_kpt(wf.pw_coeffs, wf.G_frac @ self.sc.B_sc, wf.band_energies,
                                         self.sc.b_pc, fG, self.grid, norm_mode=1)
        return UnfoldedKpt(w, dN, ns, np.zeros(wf.n_bands))


def _job(sc, rank, world):
    from bandup_b200.unfold import PwWavefunction

    def load(iK):
        coeffs, gf, ener = synth.make_wavefunction(sc, iK)
        return PwWavefunction(coeffs, gf, ener), sc.folding_G(iK), np.asarray(sc.folds[iK])

    costs = [sc.n_bands * sc.gvecs(iK).shape[0] for iK in range(sc.n_sc_kpts)]
    return ShardedUnfoldJob(OracleUnfolder(sc), sc.n_sc_kpts, costs, load, rank=rank, world_size=world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synth.make_scenario("graphene4x4", scale=0.06, n_kpts=12)
        job = _job(sc, rank, world)
        g = job.run()
        ret[rank] = (g.dN, g.row_ids, g.rows_per_rank, job.my_block, job.unfolder.max_inflight)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    sc = synth.make_scenario("graphene4x4", scale=0.06, n_kpts=12)
    single = _job(sc, 0, 1).run()
    n_rows = sum(len(f) for f in sc.folds)
    assert single.row_ids.tolist() == list(range(n_rows))          # every pc-kpt exactly once
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    for dN, ids, per_rank, block, inflight in (r0, r1):
        assert np.array_equal(ids, single.row_ids)
        assert np.array_equal(dN, single.dN)                        # same oracle, bit-identical rows
        assert sum(per_rank) == n_rows and min(per_rank) > 0
        assert inflight == 2                                        # double buffering was exercised
    assert r0[3][1] == r1[3][0] and r0[3][0] == 0 and r1[3][1] == sc.n_sc_kpts


# ---- dynamic assignment: ranks pull SC-kpts from a shared counter ---------------------------
def test_tickets_local_counter():
    from bandup_b200.sharding import KptTickets
    t = KptTickets(3)
    assert [t.take(), t.take(), t.take(), t.take(), t.take()] == [0, 1, 2, None, None]
    t = KptTickets(7, offset=1, stride=3)          # no shared store: round-robin share of rank 1 of 3
    assert [t.take(), t.take(), t.take(), t.take()] == [1, 4, None, None]
    sc = synth.make_scenario("graphene4x4", scale=0.06, n_kpts=5)
    job = _job(sc, 0, 1)
    static = job.run()
    job2 = _job(sc, 0, 1)
    job2.tickets = KptTickets(sc.n_sc_kpts)
    dyn = job2.run()
    assert job2.done_kpts == list(range(sc.n_sc_kpts))
    assert np.array_equal(dyn.dN, static.dN) and np.array_equal(dyn.row_ids, static.row_ids)


def _worker_dynamic(rank, world, port, ret):
    import time
    import torch.distributed as dist
    from bandup_b200.sharding import KptTickets
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synth.make_scenario("graphene4x4", scale=0.06, n_kpts=24)
        job = _job(sc, rank, world)
        job.tickets = KptTickets.from_default_group(sc.n_sc_kpts, job="t1")
        assert job.tickets.store is not None
        if rank == 1:                      # a rank with a slow link: every fetch takes much longer
            fetch = job.unfolder.fetch
            job.unfolder.fetch = lambda ikpt: (time.sleep(0.4), fetch(ikpt))[1]
        dist.barrier()
        g = job.run()
        # a second job under the SAME name starts from a fresh counter (a counter in the store is
        # never reset: every use of a name gets its own epoch)
        again = KptTickets.from_default_group(5, job="t1")
        dist.barrier()
        mine = list(iter(again.take, None))
        dist.barrier()
        ret[rank] = (g.dN, g.row_ids, list(job.done_kpts), mine)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_dynamic_tickets_balance_a_slow_rank():
    import torch.multiprocessing as mp
    sc = synth.make_scenario("graphene4x4", scale=0.06, n_kpts=24)
    single = _job(sc, 0, 1).run()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_dynamic, args=(2, _free_port(), ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    assert sorted(r0[3] + r1[3]) == list(range(5))                  # the re-used job name handed out 0..4 again
    for dN, ids, _, _ in (r0, r1):
        assert np.array_equal(ids, single.row_ids)
        assert np.array_equal(dN, single.dN)                        # who computed what does not matter
    assert sorted(r0[2] + r1[2]) == list(range(sc.n_sc_kpts))       # every SC-kpt exactly once
    assert len(r0[2]) > len(r1[2]) >= 1                             # the slow rank ended up with fewer
