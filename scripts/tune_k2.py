#!/usr/bin/env python
"""Sweeps the K2 tuning knobs (tile size, CTAs per SM, stage-ring depth) on one full-size
SC k-point held in HBM and prints the CUDA-event time of k2_weights_reduce per setting."""
import itertools
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import synth  # noqa: E402
from bandup_b200.unfold import GpuUnfolder, K_WEIGHTS, LAYOUT_FORTRAN_INMEM  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "sige555"
sc = synth.make_scenario(workload)
dev = torch.device("cuda:0")
u = GpuUnfolder(0)
u.verify_commens(sc.b_pc, sc.B_sc)
grid = u.set_energy_grid(*sc.energy_window())
stream = torch.cuda.current_stream(dev).cuda_stream
nb, nsp = sc.n_bands, sc.n_spinor
items = []
for iK in (0, 1):   # two blocks, alternated, each far larger than L2
    gf = sc.gvecs(iK)
    n_pw = gf.shape[0]
    g0, _ = u.folding_g0(sc.folding_G(iK))
    d_c = torch.empty((nb, n_pw * nsp, 2), dtype=torch.float32, device=dev)
    u.synth_fill_device(stream, d_c.data_ptr(), n_pw * nsp, nb, n_pw * nsp * 8, 7, iK)
    items.append((n_pw, g0, d_c, torch.from_numpy(gf).to(dev),
                  torch.from_numpy(np.sort(np.random.default_rng(iK).uniform(-20, 5, nb))).to(dev)))
nener = len(grid)
out = []
settings = [(0, 0, 0)] + list(itertools.product([2048, 4096, 8192, 16384], [1, 2, 3], [3, 4, 6, 8, 12]))
for cv, ctas, st in settings:
    u.set_tuning(cv, ctas, st)
    bufs = []
    for n_pw, g0, d_c, d_g, d_e in items:
        nf = g0.shape[0]
        ws = u.workspace_bytes(nsp, n_pw, nb, nf)
        bufs.append((torch.empty((nf, nb), dtype=torch.float64, device=dev),
                     torch.empty((nf, nener), dtype=torch.float64, device=dev),
                     torch.empty(nf, dtype=torch.int32, device=dev),
                     torch.empty(ws, dtype=torch.uint8, device=dev), ws))

    def run(n):
        for r in range(n):
            for (n_pw, g0, d_c, d_g, d_e), (d_w, d_dn, d_ns, d_ws, ws) in zip(items, bufs):
                u.unfold_device(stream, nsp, n_pw, nb, LAYOUT_FORTRAN_INMEM, n_pw * nsp * 8, d_c.data_ptr(),
                                d_g.data_ptr(), d_e.data_ptr(), g0, d_w.data_ptr(), d_dn.data_ptr(),
                                d_ns.data_ptr(), d_ws.data_ptr(), ws)
    run(2)
    torch.cuda.synchronize()
    u.set_profiling(True)
    u.reset_kernel_times()
    run(5)
    ms, n = u.kernel_time(K_WEIGHTS)
    u.set_profiling(False)
    byts = sum(nb * it[0] * nsp * 8 for it in items) / len(items)
    gbs = byts / (ms / n * 1e-3) / 1e9
    out.append(dict(chunk_vecs=cv, ctas_per_sm=ctas, n_stages=st, ms=ms / n, gbs=gbs))
    print(f"chunk_vecs={cv:6d} ctas={ctas} stages={st:2d}  {ms / n:.4f} ms  {gbs:7.1f} GB/s", flush=True)
best = max(out, key=lambda r: r["gbs"])
print("BEST", json.dumps(best))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/tune_k2_%s.json" % workload).write_text(json.dumps(out, indent=1))
