"""Two processes, two GPUs: SC-kpt shards unfolded independently, delta_N rows gathered with the
library's own ncclAllGather (bandup_gpu_gather_dN).  Skipped on boxes with fewer than 2 GPUs."""
import os
import tempfile
import time
from pathlib import Path

import numpy as np
import pytest

from bandup_b200 import synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, tmp, ret):
    from bandup_b200.sharding import ShardedUnfoldJob
    from bandup_b200.unfold import GpuUnfolder, PwWavefunction
    sc = synth.make_scenario("graphene4x4", scale=0.08, n_kpts=12)
    idf = Path(tmp) / "nccl_id"
    if rank == 0:
        uid = GpuUnfolder.comm_unique_id()
        (Path(tmp) / "nccl_id.tmp").write_bytes(uid)
        os.replace(Path(tmp) / "nccl_id.tmp", idf)          # the host's own way of sharing 128 bytes
    else:
        t0 = time.time()
        while not idf.exists():
            if time.time() - t0 > 60:
                raise TimeoutError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = idf.read_bytes()
    with GpuUnfolder(rank) as u:
        u.verify_commens(sc.b_pc, sc.B_sc)
        u.set_energy_grid(*sc.energy_window())
        u.comm_init(world, rank, uid)

        def load(iK):
            coeffs, gf, ener = synth.make_wavefunction(sc, iK)
            return PwWavefunction(coeffs, gf, ener), sc.folding_G(iK), np.asarray(sc.folds[iK])

        costs = [sc.n_bands * sc.gvecs(iK).shape[0] for iK in range(sc.n_sc_kpts)]
        job = ShardedUnfoldJob(u, sc.n_sc_kpts, costs, load, rank=rank, world_size=world)
        dN, rid, _ = job.run_local()
        allrows, allids = u.gather_dN(dN, rid)
        # rows of another shape through the same collective (what the spin columns use)
        extra = np.stack([dN * (rank + 1), -dN, dN + 1.0], axis=2)          # (n, nener, 3)
        xrows, xids = u.gather_rows(extra, rid)
        assert np.array_equal(xids, allids) and xrows.shape == (len(allids), dN.shape[1], 3)
        assert np.array_equal(xrows[:, :, 1], -allrows) and np.array_equal(xrows[:, :, 2], allrows + 1.0)
        # the device-resident form of the collective (what bench.py --gpus N uses once per job)
        import torch
        dev = torch.device("cuda", rank)
        torch.cuda.set_device(dev)
        send = torch.arange(7, dtype=torch.float64, device=dev) + 100.0 * (rank + 1)
        recv = torch.zeros((world, 7), dtype=torch.float64, device=dev)
        u.allgather_device(torch.cuda.current_stream(dev).cuda_stream, send.data_ptr(), recv.data_ptr(), 7)
        torch.cuda.synchronize()
        want = np.arange(7)[None, :] + 100.0 * (np.arange(world)[:, None] + 1)
        assert np.array_equal(recv.cpu().numpy(), want)
        ret[rank] = (allrows, allids, job.my_block, len(rid))
        u.comm_finalize()


def test_two_gpus_shard_and_nccl_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import oracle as O
    sc = synth.make_scenario("graphene4x4", scale=0.08, n_kpts=12)
    with tempfile.TemporaryDirectory() as tmp, mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, tmp, ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    n_rows = sum(len(f) for f in sc.folds)
    assert r0[1].tolist() == list(range(n_rows)) and np.array_equal(r0[1], r1[1])
    assert np.array_equal(r0[0], r1[0])                       # every rank holds the same table
    assert r0[3] > 0 and r1[3] > 0 and r0[3] + r1[3] == n_rows
    assert r0[2][1] == r1[2][0]
    grid = O.real_seq(*sc.energy_window())
    for iK in range(sc.n_sc_kpts):
        coeffs, gf, ener = synth.make_wavefunction(sc, iK)
        _, dN, _, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid, norm_mode=1)
        got = r0[0][np.asarray(sc.folds[iK])]
        assert np.max(np.abs(got - dN) / np.maximum(np.abs(dN), 1e-12)) <= 1e-6


@pytest.mark.parametrize("spinor", [False, True])
def test_host_exe_two_ranks(tmp_path, gpu_lib, spinor):
    """BandUP_gpu.x as two processes on two GPUs (-n_ranks 2): the gathered output file (with the
    five spin columns for a spinor WAVECAR) equals the one-rank output."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from bandup_b200 import bandup_files as BF
    from tests.test_host_exe import EXE, _job
    _job(tmp_path, spinor)
    common = [str(EXE), "-no_symm_sckpts", "-no_symm_avg"]
    r = subprocess.run(common + ["-out_file_nosymm", "one.dat"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    procs = [subprocess.Popen(common + ["-out_file_nosymm", "two.dat", "-n_ranks", "2", "-rank", str(k),
                                        "-nccl_id_file", "ncclid"], cwd=tmp_path, stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for k in (0, 1)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert np.array_equal(BF.read_band_struc(tmp_path / "one.dat"), BF.read_band_struc(tmp_path / "two.dat"))
    # each rank took a share of the SC-kpts
    n0 = outs[0][0].count("SC-kpt #")
    n1 = outs[1][0].count("SC-kpt #")
    assert n0 > 0 and n1 > 0
    # dynamic assignment: both ranks pull SC-kpts from a counter file; same output, every k-point once
    procs = [subprocess.Popen(common + ["-out_file_nosymm", "dyn.dat", "-n_ranks", "2", "-rank", str(k),
                                        "-nccl_id_file", "ncclid2", "-ticket_file", "tickets"], cwd=tmp_path,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for k in (0, 1)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert np.array_equal(BF.read_band_struc(tmp_path / "one.dat"), BF.read_band_struc(tmp_path / "dyn.dat"))
    assert outs[0][0].count("SC-kpt #") + outs[1][0].count("SC-kpt #") == n0 + n1
    # the unfolding-density operator file with two ranks (part files merged by rank 0)
    udo = ["-write_unf_dens_op", "-output_file_only_user_selec_direcs_unf_dens_op"]
    r = subprocess.run(common + ["-out_file_nosymm", "one_b.dat"] + udo + ["one_udo.dat"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    procs = [subprocess.Popen(common + ["-out_file_nosymm", "two_b.dat", "-n_ranks", "2", "-rank", str(k),
                                        "-nccl_id_file", "ncclid3"] + udo + ["two_udo.dat"], cwd=tmp_path,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for k in (0, 1)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    one, two = (tmp_path / "one_udo.dat").read_text(), (tmp_path / "two_udo.dat").read_text()
    assert one == two and "UnfDensOp[m1,m2]" in one
    assert not list(tmp_path.glob("two_udo.dat.part*"))


def test_bench_two_ranks_prints_one_json_line():
    """The driver's N>1 launch line: torchrun, one rank per GPU, rank 0 prints exactly one JSON line."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(root / "bench.py"), "--gpus", "2",
                        "--steps", "3", "--warmup", "3", "--workload", "si333_vacancy", "--kpts-per-step", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["cpu_baseline"] is None                     # rank 0 at N = 1 only
    assert "bandup_gpu_allgather_device" in d["config"]["parallelism"]
    e = d["e2e"]
    assert e["e2e_checked"] is True and len(e["h2d_GBps_per_rank"]) == 2 and sorted(e["gpu_of_rank"]) == [0, 1]
    assert sum(e["kpts_per_rank"]) == 2 * 2 * e["steps"]


def test_host_exe_rank_without_kpoints(tmp_path, gpu_lib):
    """Two ranks, a job with a single SC k-point: one rank has nothing to unfold and still takes part
    in the collectives; the output equals the one-rank output (static blocks and shared counter)."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from bandup_b200 import bandup_files as BF
    from tests.test_host_exe import EXE, _job
    _job(tmp_path, True)
    BF.write_pckpts(tmp_path / "KPOINTS_prim_cell.in", [([0.0, 0.0, 0.0], [0.0, 0.0, 0.0])], [1])   # Gamma only
    common = [str(EXE), "-no_symm_sckpts", "-no_symm_avg", "-write_unf_dens_op"]
    r = subprocess.run(common + ["-out_file_nosymm", "one.dat", "-output_file_only_user_selec_direcs_unf_dens_op", "one_udo.dat"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("SC-kpt #") == 1
    for tag, extra in (("st", []), ("dy", ["-ticket_file", "tk"])):
        procs = [subprocess.Popen(common + ["-out_file_nosymm", f"{tag}.dat", "-output_file_only_user_selec_direcs_unf_dens_op",
                                            f"{tag}_udo.dat", "-n_ranks", "2", "-rank", str(k), "-nccl_id_file", f"id_{tag}"] + extra,
                                  cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for k in (0, 1)]
        outs = [p.communicate(timeout=300) for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        assert sorted(o[0].count("SC-kpt #") for o in outs) == [0, 1]
        assert np.array_equal(BF.read_band_struc(tmp_path / "one.dat"), BF.read_band_struc(tmp_path / f"{tag}.dat"))
        assert (tmp_path / "one_udo.dat").read_text() == (tmp_path / f"{tag}_udo.dat").read_text()
        assert not (tmp_path / f"id_{tag}").exists() and not (tmp_path / "tk").exists()     # rank 0 cleaned up
