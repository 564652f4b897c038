#!/usr/bin/env python
"""Time-bounded randomised differential test of the CUDA path against the CPU oracle, wider than
tests/test_gpu_parity.py::test_random_cells_and_shapes: random triclinic cells and integer SC
matrices (both hands), scalar / spinor, both coefficient layouts with padded strides, complex64 /
complex128, box / Gaussian delta_N, 1..90 pc-kpts per SC-kpt with repeated cosets (several passes
when more than 64 are distinct), the three K2 accumulator choices, dependent launch on/off, raw
records with the plane-wave list made on the device, rho and Pauli vectors on some cases.

    gpurun -- 'python scripts/fuzz_gpu.py --seconds 150 --seed 1 > gpurun_out/fuzz.log'

Prints one line per failure (with the seed of the case, re-runnable with --only) and a summary;
exit code 1 when anything failed."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bandup_b200 import synth  # noqa: E402
from bandup_b200._capi import CONST  # noqa: E402
from bandup_b200.unfold import (GpuUnfolder, LAYOUT_VASP_RECORD, PwWavefunction)  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import oracle_np as ON  # noqa: E402


def rel(a, b, floor=1e-12):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if np.size(a) else 0.0


def one_case(u, seed):
    rng = np.random.default_rng(seed)
    while True:
        a_pc = np.eye(3) * rng.uniform(2.5, 4.5, 3) + rng.uniform(-0.7, 0.7, (3, 3))
        M = rng.integers(-4, 5, (3, 3))
        det = int(round(np.linalg.det(M)))
        if 2 <= abs(det) <= 130 and abs(np.linalg.det(a_pc)) > 5.0:
            break
    n_spinor = int(rng.integers(1, 3))
    dp = rng.random() < 0.3
    smear = None if rng.random() < 0.6 else float(rng.uniform(0.01, 0.3))
    kf = rng.uniform(-0.5, 0.5, (3, 3))
    sc = synth.Scenario(f"fuzz{seed}", a_pc, M, float(rng.uniform(15.0, 45.0)), int(rng.integers(1, 40)), n_spinor,
                        kf @ synth.rec_latt(a_pc), e_fermi=float(rng.uniform(-2, 2)), emin=-6.0, emax=4.0,
                        dE=float(rng.choice([0.05, 0.1, 0.37])))
    what = dict(seed=seed, det=det, n_spinor=n_spinor, dp=dp, smear=smear, n_bands=sc.n_bands)
    u.set_complex128(dp)
    u.set_option(CONST["BANDUP_GPU_OPT_K2_VARIANT"], int(rng.integers(0, 3)))
    u.set_option(CONST["BANDUP_GPU_OPT_DEPENDENT_LAUNCH"], int(rng.integers(0, 4)))
    Mgot = u.verify_commens(sc.b_pc, sc.B_sc)
    assert np.array_equal(Mgot, M if det > 0 else -M), "M"
    grid = u.set_energy_grid(*sc.energy_window(), smearing_for_delta_function=smear)
    iK = int(rng.integers(0, sc.n_sc_kpts))
    coeffs, gf, ener = synth.make_wavefunction(sc, iK, seed=int(rng.integers(1, 10**6)))
    n_pw = gf.shape[0]
    if n_pw < 4:
        return None
    if rng.random() < 0.3:                       # an odd / short plane-wave count, rows off 16-byte alignment
        cut = int(rng.integers(1, min(4, n_pw - 2)))
        coeffs, gf = np.ascontiguousarray(coeffs[:, :-cut]), gf[:-cut]
        n_pw -= cut
    nb = coeffs.shape[0]
    ener = rng.uniform(sc.e_fermi - 7, sc.e_fermi + 5, nb)
    if dp:
        coeffs = coeffs.astype(np.complex128) * (1.0 + 1e-7 * rng.standard_normal(coeffs.shape))
    # pc-kpts: random SC reciprocal lattice shifts of K, some repeated cosets
    n_fold = int(rng.choice([1, 2, 3, 5, 9, 33, 70, 200]))
    g0 = rng.integers(-5, 6, (n_fold, 3)).astype(np.int32)
    if n_fold > 2:
        g0[-1] = g0[0] + M[:, 0] - 2 * M[:, 2]                   # b_i = sum_j M_ji B_j: the same coset as the first
    fG = g0.astype(np.float64) @ sc.B_sc
    what.update(n_pw=n_pw, n_fold=n_fold)
    gc = gf @ sc.B_sc
    if dp:
        w, dN, ns, norms = O.unfold_kpt_dp(coeffs, gc, ener, sc.b_pc, fG, grid, smearing=smear or 0.0)
        tol = 1e-9
    else:
        w, dN, ns, _ = O.unfold_kpt(coeffs, gc, ener, sc.b_pc, fG, grid, smearing=smear or 0.0, norm_mode=1)
        tol = 1e-6
    esz = 16 if dp else 8
    mode = rng.random()
    if mode < 0.35:                              # raw records, plane-wave list on the device
        nrecl = n_pw * n_spinor * esz + esz * int(rng.integers(0, 4))
        rec = np.zeros((nb, nrecl), dtype=np.uint8)
        rows = np.ascontiguousarray(coeffs.transpose(0, 2, 1)).reshape(nb, n_spinor * n_pw)
        rec[:, :n_pw * n_spinor * esz] = rows.view(np.uint8).reshape(nb, -1)
        full = synth.gen_gvecs(sc.sc_kpts_frac[iK], sc.ecut, sc.B_sc)
        if full.shape[0] == n_pw:                # the device list must be THE list of this k-point
            nsp, npw = u.submit_vasp(1, rec, nrecl, n_pw * n_spinor, nb, sc.sc_kpts_frac[iK], sc.ecut, ener, g0=g0)
            assert (nsp, npw) == (n_spinor, n_pw), "device plane-wave count"
            assert np.array_equal(u.fetch_gvecs(1, n_pw), gf), "device plane-wave list"
            res = u.fetch(1, want_selected=True)
            what["path"] = "vasp records + device G"
        else:
            wf = PwWavefunction(rec, gf, ener, n_pw=n_pw, n_bands=nb, n_spinor=n_spinor, layout=LAYOUT_VASP_RECORD,
                                band_stride_bytes=nrecl)
            res = u.perform_unfolding(wf, g0=g0, ikpt=1)
            what["path"] = "vasp records"
    else:
        res = u.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=g0, ikpt=1)
        what["path"] = "in-memory"
    assert np.array_equal(res.n_selected, ns), "n_selected"
    if n_fold <= 9:
        for f in range(n_fold):
            sel, _ = O.select_coeffs(gc, fG[f], sc.b_pc)
            assert np.array_equal(res.selected_coeff_indices[f], sel), "selected indices"
    assert rel(res.spectral_weight, w) <= tol, f"weights {rel(res.spectral_weight, w):.2e}"
    # delta_N: the relative tolerance, plus the absolute rounding of lambda itself -- with a Gaussian
    # smearing a far tail is the difference of two erf values next to +-1 (each good to an ulp of 1),
    # so a 1e-11 tail is only defined to ~1e-16 absolute in the reference's own arithmetic
    slack = 4 * np.finfo(np.float64).eps * np.sum(np.abs(w), axis=1, keepdims=True)
    bad = np.abs(res.dN - dN) > tol * np.abs(dN) + slack
    assert not bad.any(), f"delta_N {rel(res.dN, dN):.2e} (abs {np.max(np.abs(res.dN - dN)):.2e})"
    if rng.random() < 0.3 and nb <= 24 and n_fold <= 5:     # rho (+ Pauli vectors) on the resident k-point
        ops = u.calc_rho(1)
        cren = O.renormalize(coeffs.astype(np.complex64), 1)[0] if not dp else None
        norms_ = np.sqrt(np.sum(np.abs(coeffs) ** 2, axis=(1, 2)))
        d = grid[1] - grid[0]
        n = 0
        for f in range(n_fold):
            sel, _ = O.select_coeffs(gc, fG[f], sc.b_pc)
            for ie in np.nonzero(res.dN[f] >= 1e-3)[0]:
                if dp:
                    rho = ON.calc_rho(coeffs / norms_[:, None, None], sel, ener, res.dN[f, ie], grid[ie], d)
                else:
                    rho = ON.calc_rho(cren, sel, ener, res.dN[f, ie], grid[ie], d)
                got = ops.dense(f, int(ie) + 1, nb)
                keep = np.abs(rho) >= 1.2e-5
                assert not keep.any() or np.max(np.abs(got[keep] - rho[keep])) <= 3e-6, "rho"
                assert not np.any(np.abs(got[np.abs(rho) < 0.8e-5])), "rho below the threshold stored"
                n += 1
        what["rho_ops"] = n
        if n_spinor == 2:
            sig = u.pauli_vectors(1, n_fold)
            assert np.all(np.isfinite(sig)), "Pauli vectors not finite"
            cr = cren if not dp else coeffs / norms_[:, None, None]
            for f in range(n_fold):
                rhos = {int(ie) + 1: ops.dense(f, int(ie) + 1, nb) for ie in np.nonzero(res.dN[f] >= 1e-3)[0]}
                for ie, v in ON.unfolded_pauli_vectors(cr, ener, grid, rhos, d).items():
                    err = np.max(np.abs(v - sig[f, ie - 1]))
                    assert err <= 1e-5 * max(1.0, np.max(np.abs(v))), f"Pauli vectors {err:.2e} of {np.max(np.abs(v)):.2e}"
                rest = np.setdiff1d(np.arange(len(grid)), np.array(sorted(rhos), dtype=int) - 1)
                assert not np.any(sig[f, rest]), "Pauli vector where there is no operator"
    return what


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--only", type=int, default=None, help="re-run one case seed")
    args = ap.parse_args()
    t0 = time.time()
    n_ok = n_fail = n_skip = 0
    seeds = [args.only] if args.only is not None else None
    i = 0
    u = GpuUnfolder(0)
    try:
        while (seeds is None and time.time() - t0 < args.seconds) or (seeds is not None and i < len(seeds)):
            seed = seeds[i] if seeds is not None else args.seed * 1_000_003 + i
            i += 1
            try:
                r = one_case(u, seed)
                if r is None:
                    n_skip += 1
                else:
                    n_ok += 1
                    if args.only is not None:
                        print("ok", r)
            except Exception as exc:                      # noqa: BLE001 -- every failure is reported
                n_fail += 1
                print(f"FAIL seed {seed}: {type(exc).__name__}: {exc}", flush=True)
                u.close()                                # a fresh context: nothing of the failed case is left in flight
                u = GpuUnfolder(0)
    finally:
        u.close()
    print(f"fuzz: {n_ok} ok, {n_fail} failed, {n_skip} skipped in {time.time() - t0:.0f} s (seed {args.seed})")
    sys.exit(1 if n_fail else 0)


if __name__ == "__main__":
    main()
