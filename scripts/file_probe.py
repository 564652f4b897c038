import os, sys, tempfile, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import synth
from bandup_b200.unfold import GpuUnfolder, pinned_empty, pinned_free
from bandup_b200.wavecar import Wavecar, write_wavecar
import bandup_b200.wavecar as W
sc = synth.make_scenario("sige555", n_kpts=8)
ks = [0, 1, 2]
rng = np.random.default_rng(1)
tmp = tempfile.mkdtemp()
path = Path(tmp) / "WAVECAR"
coeffs, eners = [], []
for iK in ks:
    gf = sc.gvecs(iK)
    c = np.ones((sc.n_bands, gf.shape[0], 1), dtype=np.complex64)
    coeffs.append(c); eners.append(np.sort(rng.uniform(-20, 5, sc.n_bands)))
write_wavecar(path, sc.A_sc, sc.ecut, [sc.sc_kpts_frac[i] for i in ks], coeffs, eners)
del coeffs
wc = Wavecar(path)
n = wc.band_block_bytes()
t0 = time.perf_counter(); buf = pinned_empty(n); t1 = time.perf_counter()
print(f"pinned alloc {n/1e9:.2f} GB: {t1-t0:.2f} s")
for thr in (1, 4, 8, 16):
    W.READ_THREADS = thr
    t0 = time.perf_counter(); wc.read_band_records(2, out=buf); dt = time.perf_counter() - t0
    print(f"read {thr} threads: {n/dt/1e9:.2f} GB/s")
pag = np.empty(n, dtype=np.uint8)
t0 = time.perf_counter(); wc.read_band_records(2, out=pag); dt = time.perf_counter() - t0
print(f"read into pageable (first touch): {n/dt/1e9:.2f} GB/s")
t0 = time.perf_counter(); wc.read_band_records(2, out=pag); dt = time.perf_counter() - t0
print(f"read into pageable (touched): {n/dt/1e9:.2f} GB/s")
with GpuUnfolder(0) as u:
    u.verify_commens(sc.b_pc, sc.B_sc); u.set_energy_grid(*sc.energy_window())
    for rep in range(3):
        t0 = time.perf_counter()
        wc.submit_to(u, 1, 2, sc.folding_G(1), raw=buf[:n]); t1 = time.perf_counter()
        u.fetch(1); t2 = time.perf_counter()
        print(f"submit {t1-t0:.3f} s fetch {t2-t1:.3f} s  -> {n/(t2-t0)/1e9:.1f} GB/s")
t0 = time.perf_counter(); pinned_free(buf); print(f"pinned free {time.perf_counter()-t0:.2f} s")
os.remove(path)
