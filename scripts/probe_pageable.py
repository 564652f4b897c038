#!/usr/bin/env python
"""Upload rate of bandup_gpu_submit_kpt from pageable, registered and library-pinned memory."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import synth
from bandup_b200.unfold import GpuUnfolder, PwWavefunction, host_register, host_unregister, pinned_empty

sc = synth.make_scenario("sige555")
u = GpuUnfolder(0)
u.verify_commens(sc.b_pc, sc.B_sc)
u.set_energy_grid(*sc.energy_window())
gf = sc.gvecs(0)
nb, n_pw = sc.n_bands // 2, gf.shape[0]
rng = np.random.default_rng(0)
ener = np.sort(rng.uniform(-20, 5, nb))
g0, _ = u.folding_g0(sc.folding_G(0))
a = np.empty((nb, n_pw, 1), dtype=np.complex64)
a.view(np.float32)[:] = 0.01
pin = pinned_empty(a.nbytes)
pin.view(np.float32)[:] = 0.01


def rate(arr, label, n=4):
    wf = PwWavefunction(arr, gf, ener)
    u.perform_unfolding(wf, g0=g0, want_selected=False)
    t0 = time.perf_counter()
    for k in range(n):
        u.perform_unfolding(wf, g0=g0, want_selected=False, ikpt=100 + k)
    dt = (time.perf_counter() - t0) / n
    print(f"{label}: {arr.nbytes / dt / 1e9:.1f} GB/s ({arr.nbytes / 1e9:.2f} GB in {dt * 1e3:.1f} ms)", flush=True)


rate(a, "pageable (staged through the library's pinned ring)")
host_register(a)
rate(a, "registered in place")
host_unregister(a)
rate(pin.view(np.complex64).reshape(nb, n_pw, 1), "bandup_gpu_pinned_alloc")
u.close()
