#!/usr/bin/env python
"""File-driven throughput: a synthetic WAVECAR on local disk -> unfold_wavecar (parallel preadv into
page-locked buffers on a helper thread, raw records uploaded as they are, plane-wave lists made on
the device) -> delta_N.  Reports SC-kpts/s and GB/s of coefficient data, page cache warm and cold."""
import os, sys, tempfile, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import synth
from bandup_b200.unfold import GpuUnfolder
from bandup_b200.wavecar import Wavecar, unfold_wavecar, write_wavecar

name = sys.argv[1] if len(sys.argv) > 1 else "si333_vacancy"
n_k = int(sys.argv[2]) if len(sys.argv) > 2 else 24
sc = synth.make_scenario(name, n_kpts=2 * n_k)
ks = list(range(min(n_k, sc.n_sc_kpts)))
rng = np.random.default_rng(1)
tmp = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))
path = Path(tmp) / "WAVECAR"
t0 = time.perf_counter()
coeffs, eners = [], []
for iK in ks:
    gf = sc.gvecs(iK)
    c = (rng.standard_normal((sc.n_bands, gf.shape[0], sc.n_spinor, 2)).astype(np.float32))
    coeffs.append(c.view(np.complex64).reshape(sc.n_bands, gf.shape[0], sc.n_spinor))
    eners.append(np.sort(rng.uniform(-20, 5, sc.n_bands)))
write_wavecar(path, sc.A_sc, sc.ecut, [sc.sc_kpts_frac[i] for i in ks], coeffs, eners)
size = path.stat().st_size
print(f"wrote {size / 1e9:.2f} GB in {time.perf_counter() - t0:.1f} s", flush=True)
del coeffs
pc = np.concatenate([sc.pc_kpts_cart[sc.folds[i]] for i in ks])
with GpuUnfolder(0) as u:
    wc = Wavecar(path)
    for label in ("page cache warm", "page cache warm (2nd)"):
        t0 = time.perf_counter()
        got, grid = unfold_wavecar(u, wc, sc.b_pc, pc, *sc.energy_window())
        dt = time.perf_counter() - t0
        print(f"{label}: {len(ks) / dt:.1f} SC-kpts/s, {size / dt / 1e9:.2f} GB/s of file, rows {got.dN.shape}", flush=True)
    try:
        os.system("sync; echo 3 > /proc/sys/vm/drop_caches 2>/dev/null")
        t0 = time.perf_counter()
        got, grid = unfold_wavecar(u, wc, sc.b_pc, pc, *sc.energy_window())
        dt = time.perf_counter() - t0
        print(f"after drop_caches: {len(ks) / dt:.1f} SC-kpts/s, {size / dt / 1e9:.2f} GB/s of file", flush=True)
    except Exception as e:
        print("cold run failed:", e)
os.remove(path)
