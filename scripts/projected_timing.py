#!/usr/bin/env python
"""BASELINE config 4 (projected-unfold on the Si 3x3x3 vacancy cell, 215 atoms x 9 orbitals, 500
bands): time per SC k-point of the device path (rho, orbital-contribution matrix, N Re Tr(rho O))
next to the NumPy restatement of the reference's Python loops on the host (the oracle; the checker,
timed here only as the CPU figure)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import synth
from bandup_b200.unfold import GpuUnfolder, PwWavefunction

sc = synth.make_scenario("si333_vacancy")
u = GpuUnfolder(0)
u.verify_commens(sc.b_pc, sc.B_sc)
grid = u.set_energy_grid(*sc.energy_window())
iK = 11
coeffs, gf, ener = synth.make_wavefunction(sc, iK)
nb = coeffs.shape[0]
fG = sc.folding_G(iK)
rng = np.random.default_rng(12)
n_at, n_orb = 215, 9
proj = (rng.standard_normal((n_at * n_orb, nb)) + 1j * rng.standard_normal((n_at * n_orb, nb))) * \
    np.repeat(rng.uniform(0.02, 0.2, n_at), n_orb)[:, None]
dual = np.linalg.pinv(proj)
mask = np.zeros(n_at * n_orb, np.uint8)
mask.reshape(n_at, n_orb)[:108, 1:4] = 1
wf = PwWavefunction(coeffs, gf, ener)


def once():
    t = [time.perf_counter()]
    out = u.perform_unfolding(wf, folding_G=fG, ikpt=3, want_selected=False); t.append(time.perf_counter())
    res = u.calc_rho(3); t.append(time.perf_counter())
    O = u.get_orbital_contribution_matrix(dual, proj, mask, ener); t.append(time.perf_counter())
    vals = u.unfold_operator(3, fG.shape[0], min_abs_rho=1e-4, clip=True); t.append(time.perf_counter())
    return out, res, O, vals, np.diff(t)


once()
ts = np.array([once()[4] for _ in range(5)])
out, res, O, vals, _ = once()
med = np.median(ts, axis=0) * 1e3
print(f"device path per SC-kpt (ms): unfold {med[0]:.2f}  rho {med[1]:.2f}  orbital matrix {med[2]:.2f}  "
      f"N Re Tr(rho O) {med[3]:.2f}  total {med.sum():.2f}   [{res.n_energies} operators, {len(res.rho)} stored elements]")
from oracle import oracle_np as ON
t0 = time.perf_counter()
O_ref = ON.orb_contrib_matrix(dual, proj, mask, ener)
t1 = time.perf_counter()
n = 0
for f in range(fG.shape[0]):
    for ie in res.energies_of(f):
        s = (res.fold == f) & (res.iener == ie) & (res.m1 <= res.m2)
        ON.unfold_rho_operator(list(res.m1[s] - 1), list(res.m2[s] - 1), list(res.rho[s]), O_ref, out.dN[f, ie - 1],
                               min_abs_rho=1e-4, clip=True)
        n += 1
t2 = time.perf_counter()
print(f"NumPy restatement of the reference's Python on the host: orbital matrix {t1 - t0:.2f} s, "
      f"N Re Tr(rho O) for {n} operators {t2 - t1:.2f} s")
u.close()
