#!/usr/bin/env python
"""Host-call cost vs device time of one batched step, per workload (no profiling events)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import synth
from bandup_b200.unfold import GpuUnfolder, LAYOUT_FORTRAN_INMEM

for wl, kps in (("graphene4x4", 8), ("si333_vacancy", 8), ("sige555", 4)):
    sc = synth.make_scenario(wl)
    dev = torch.device("cuda:0")
    u = GpuUnfolder(0)
    u.verify_commens(sc.b_pc, sc.B_sc)
    grid = u.set_energy_grid(*sc.energy_window())
    stream = torch.cuda.current_stream(dev).cuda_stream
    nb, nsp = sc.n_bands, sc.n_spinor
    ks, keep = [], []
    for iK in range(kps):
        gf = sc.gvecs(iK); n_pw = gf.shape[0]
        g0, _ = u.folding_g0(sc.folding_G(iK)); nf = g0.shape[0]
        d_c = torch.empty((nb, n_pw * nsp, 2), dtype=torch.float32, device=dev)
        u.synth_fill_device(stream, d_c.data_ptr(), n_pw * nsp, nb, n_pw * nsp * 8, 3, iK)
        t = dict(c=d_c, g=torch.from_numpy(gf).to(dev), e=torch.from_numpy(np.sort(np.random.default_rng(iK).uniform(-20, 5, nb))).to(dev),
                 w=torch.empty((nf, nb), dtype=torch.float64, device=dev), dn=torch.empty((nf, len(grid)), dtype=torch.float64, device=dev),
                 ns=torch.empty(nf, dtype=torch.int32, device=dev))
        keep.append(t)
        ks.append(dict(n_spinor=nsp, n_pw=n_pw, n_bands=nb, layout=LAYOUT_FORTRAN_INMEM, band_stride_bytes=n_pw * nsp * 8,
                       d_coeffs=t["c"].data_ptr(), d_g_frac=t["g"].data_ptr(), d_band_energies=t["e"].data_ptr(), g0=g0,
                       d_weights=t["w"].data_ptr(), d_dN=t["dn"].data_ptr(), d_n_selected=t["ns"].data_ptr()))
    descs, k2 = u.make_batch(ks)
    ws = u.batch_workspace_bytes(descs)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    for _ in range(5):
        u.unfold_device_batch(stream, descs, d_ws.data_ptr(), ws)
    torch.cuda.synchronize()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        u.unfold_device_batch(stream, descs, d_ws.data_ptr(), ws)
    t_host = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    print(f"{wl}: host call {t_host / n * 1e6:.1f} us/step, device {e0.elapsed_time(e1) / n * 1e3:.1f} us/step, "
          f"bytes/step {sum(nb * k['n_pw'] * nsp * 8 for k in ks) / 1e6:.1f} MB", flush=True)
    u.close()
