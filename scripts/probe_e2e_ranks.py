#!/usr/bin/env python
"""Under torchrun: every rank copies page-locked host blocks to its GPU at the same time, first
with plain cudaMemcpyAsync (torch), then through bandup_gpu_submit_kpt/fetch_kpt; per-rank GB/s.
Separates 'the box cannot feed N GPUs' from 'the library's host path does not scale'."""
import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import synth
from bandup_b200.unfold import GpuUnfolder, PwWavefunction, pinned_empty

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device(f"cuda:{lr}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def gather(x):
    if world == 1:
        return [x]
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


sc = synth.make_scenario("sige555")
u = GpuUnfolder(lr)
u.verify_commens(sc.b_pc, sc.B_sc)
grid = u.set_energy_grid(*sc.energy_window())
nb, nsp = sc.n_bands, sc.n_spinor
gf = sc.gvecs(rank)
n_pw = gf.shape[0]
nbytes = nb * n_pw * nsp * 8
hosts = [pinned_empty(nbytes) for _ in range(2)]
for h in hosts:
    h[:] = 1
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
# 1. plain copies
th = [torch.from_numpy(h) for h in hosts]
d.copy_(th[0], non_blocking=True)
sync()
t0 = time.perf_counter()
for i in range(8):
    d.copy_(th[i & 1], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
r = gather(8 * nbytes / dt / 1e9)
if rank == 0:
    print("plain cudaMemcpyAsync per rank GB/s:", [round(v, 1) for v in r], "total", round(sum(r), 1), flush=True)
# 2. through the library
rng = np.random.default_rng(1)
ener = np.sort(rng.uniform(-20, 5, nb))
for h in hosts:
    h.view(np.float32)[:] = 0.01
g0, _ = u.folding_g0(sc.folding_G(rank))
wfs = [PwWavefunction(h.view(np.complex64).reshape(nb, n_pw, nsp), gf, ener) for h in hosts]
u.submit(0, wfs[0], g0=g0); u.fetch(0)
sync()
t0 = time.perf_counter()
n = 8
u.submit(1, wfs[0], g0=g0)
for j in range(n):
    if j + 1 < n:
        u.submit(2 + j, wfs[(j + 1) & 1], g0=g0)
    u.fetch(1 + j)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
r = gather(n * nbytes / dt / 1e9)
if rank == 0:
    print("submit/fetch per rank GB/s:", [round(v, 1) for v in r], "total", round(sum(r), 1), flush=True)
u.close()
if world > 1:
    dist.destroy_process_group()
