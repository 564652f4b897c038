"""The CUDA path, through the C ABI, against golden vectors produced by EXECUTING the reference's
own Fortran sources (tests/golden/make_fortran_golden.py; interpreter: oracle/f90interp.py).
No oracle in between: what is compared is the reference's output and the library's.
Tolerances (BASELINE.json north_star): selected indices and the plane-wave list bit-exact; weights,
delta_N, rho, Pauli vectors within 1e-6 relative (abs floor 1e-12 / 1e-9 for single-precision
sums) -- the golden run renormalises with the reference's single-precision dot_product, the
library with an fp64 norm, which is where the last 1e-7 comes from."""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden"
CASES = json.loads((GOLD / "fortran_golden.json").read_text())["cases"]
REL = 1e-6


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "fortran_golden.npz")


def case(gold, name):
    pre = name + "/"
    return {k[len(pre):]: gold[k] for k in gold.files if k.startswith(pre)}


def rel(a, b, floor=1e-12):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if np.size(a) else 0.0


def configure(u, g, smearing=None):
    u.set_complex128(bool(g["dp_build"]))                              # the reference's kind_cplx_coeffs = dp build
    M = u.verify_commens(g["b_pc"], g["B_sc"])
    assert np.array_equal(M, np.rint(g["M"]).astype(np.int32))       # nint(M) of the reference's M
    grid = g["energy_grid"]
    got = u.set_energy_grid(grid[0], grid[-1], grid[1] - grid[0], smearing_for_delta_function=smearing)
    assert np.array_equal(got, grid)                                  # real_seq, bit for bit
    return grid


def wavefunction(g):
    from bandup_b200.unfold import PwWavefunction
    c = np.ascontiguousarray(np.transpose(g["coeffs_in"], (2, 1, 0)))  # (n_bands, n_pw, n_spinor)
    return PwWavefunction(c, g["g_frac"], g["band_energies"])


def selected_lists(g):
    off = np.concatenate([[0], np.cumsum(g["n_selected"])])
    return [g["selected"][off[f]:off[f + 1]] for f in range(len(g["n_selected"]))]


@pytest.mark.parametrize("name", CASES)
def test_unfold_against_the_reference_run(unfolder, gold, name):
    g = case(gold, name)
    configure(unfolder, g)
    unfolder.set_perform_unfold(bool(g["perform_unfold"]))
    res = unfolder.perform_unfolding(wavefunction(g), folding_G=g["folding_G"], ikpt=1, want_selected=True)
    assert np.array_equal(res.n_selected, g["n_selected"])
    for a, b in zip(res.selected_coeff_indices, selected_lists(g)):
        assert np.array_equal(a, b)                                    # coset indices, bit-exact
    assert rel(res.spectral_weight, g["weights"]) <= REL
    assert rel(res.dN, g["delta_N"]) <= REL
    if bool(g["dp_build"]):                                            # a true fp64 path agrees far better
        assert rel(res.spectral_weight, g["weights"]) <= 1e-10 and rel(res.dN, g["delta_N"]) <= 1e-10
    if name == "fcc_negative_det":                                     # the all-zero band stays unscaled
        assert res.band_norms[2] == 0.0 and not np.any(res.spectral_weight[:, 2])
    if not bool(g["perform_unfold"]):
        assert np.array_equal(res.spectral_weight, np.ones_like(res.spectral_weight))


@pytest.mark.parametrize("name", CASES)
def test_device_plane_wave_list_against_the_reference_reader(unfolder, gold, name):
    """bandup_gpu_submit_kpt_vasp: raw band records in, plane-wave list regenerated on the device."""
    g = case(gold, name)
    configure(unfolder, g)
    unfolder.set_perform_unfold(bool(g["perform_unfold"]))
    cin = g["coeffs_in"]                                               # (n_spinor, n_pw, n_bands)
    nsp, npw, nb = cin.shape
    nplane = nsp * npw
    esz = cin.dtype.itemsize                                           # 8, or 16 in a dp build
    nrecl = esz * nplane + 48                                          # padded records, as in a real file
    rec = np.zeros((nb, nrecl), dtype=np.uint8)
    for b in range(nb):
        row = np.ascontiguousarray(np.concatenate([cin[s, :, b] for s in range(nsp)]))
        rec[b, :esz * nplane] = row.view(np.uint8)
    got_nsp, got_npw = unfolder.submit_vasp(3, rec, nrecl, nplane, nb, g["kpt_frac"], float(g["ecut"]),
                                            g["band_energies"], folding_G=g["folding_G"])
    assert (got_nsp, got_npw) == (nsp, npw)
    assert np.array_equal(unfolder.fetch_gvecs(3, npw), g["g_frac"])   # count, order, indices
    res = unfolder.fetch(3, want_selected=True)
    assert np.array_equal(res.n_selected, g["n_selected"])
    for a, b in zip(res.selected_coeff_indices, selected_lists(g)):
        assert np.array_equal(a, b)
    assert rel(res.spectral_weight, g["weights"]) <= REL
    assert rel(res.dN, g["delta_N"]) <= REL


def test_gaussian_smearing_and_spectral_function_against_the_reference_run(unfolder, gold):
    g = case(gold, "triclinic_scalar")
    configure(unfolder, g, smearing=float(g["smearing"]))
    res = unfolder.perform_unfolding(wavefunction(g), folding_G=g["folding_G"], ikpt=1)
    assert rel(res.dN, g["delta_N_smeared"]) <= REL
    sf = unfolder.calc_spectral_function(1, len(g["n_selected"]), float(g["sf_std"]))
    assert rel(sf[0], g["spectral_function"], 1e-12) <= REL


@pytest.mark.parametrize("name", CASES)
def test_unfolding_density_operator_against_the_reference_run(unfolder, gold, name):
    g = case(gold, name)
    configure(unfolder, g)
    unfolder.set_perform_unfold(bool(g["perform_unfold"]))             # False: calc_rho's other branch
    unfolder.perform_unfolding(wavefunction(g), folding_G=g["folding_G"], ikpt=1)
    ops = unfolder.calc_rho(1)
    want = g["rho"]
    # which elements are stored, and in which order (fold, energy, m1, m2): identical
    assert np.array_equal(ops.fold, want[:, 0].astype(np.int32))
    assert np.array_equal(ops.iener, want[:, 1].astype(np.int32))
    assert np.array_equal(ops.m1, want[:, 2].astype(np.int32))
    assert np.array_equal(ops.m2, want[:, 3].astype(np.int32))
    ref = (want[:, 4] + 1j * want[:, 5])
    # rho is a complex(sp) quantity in the reference: 1e-6 of its scale (|rho| <= 1), elements near the
    # 1e-5 storage threshold included
    assert np.max(np.abs(ops.rho - ref)) <= 2e-6


def test_pauli_vectors_against_the_reference_run(unfolder, gold):
    from bandup_b200.unfold import calc_spin_projections
    g = case(gold, "hex_spinor")
    configure(unfolder, g)
    unfolder.perform_unfolding(wavefunction(g), folding_G=g["folding_G"], ikpt=1)
    unfolder.calc_rho(1)
    n_fold = len(g["n_selected"])
    sig = unfolder.pauli_vectors(1, n_fold)
    assert sig.shape == g["sigma"].shape
    assert np.max(np.abs(sig - g["sigma"])) <= 5e-6                    # sp inner products over all plane waves
    assert np.any(np.abs(g["sigma"]) > 1e-3)
    for f in range(n_fold):
        for ie in np.nonzero(np.any(g["sigma"][f] != 0.0, axis=1))[0]:
            perp, para = calc_spin_projections(sig[f, ie], g["pckpts_cart"][f], np.array([0.0, 0.0, 1.0]),
                                               origin=np.zeros(3))
            assert abs(perp - g["spin_proj"][f, ie, 0]) <= 1e-5
            assert abs(para - g["spin_proj"][f, ie, 1]) <= 1e-5


@pytest.mark.parametrize("name", [c for c in CASES if c.startswith("symm_avg")])
def test_symmetry_average_against_the_reference_run(unfolder, gold, name):
    """K8/K9 (bandup_gpu_symm_average, bandup_gpu_symm_average_rho) fed with the reference's own
    per-direction delta_N / rho / spin columns, against what get_delta_Ns_for_output made of them."""
    from bandup_b200.unfold import UnfoldDensityOps
    g = case(gold, name)
    n_dirs = int(g["n_dirs"])
    n_k = len(g["n_selected"]) // n_dirs
    w = g["dir_weights"]
    avg = unfolder.symm_average(w, g["delta_N"].reshape(n_dirs, n_k, -1))
    assert np.array_equal(avg, g["avg_delta_N"])                       # bit for bit
    ops = []
    for d in range(n_dirs):
        r = g["rho"][(g["rho"][:, 0] >= d * n_k) & (g["rho"][:, 0] < (d + 1) * n_k)]
        ops.append(UnfoldDensityOps(0, (r[:, 0] - d * n_k).astype(np.int32), r[:, 1].astype(np.int32),
                                    r[:, 2].astype(np.int32), r[:, 3].astype(np.int32),
                                    (r[:, 4] + 1j * r[:, 5]).astype(np.complex64)))
    res = unfolder.symm_average_rho(w, ops, avg, g["weights"].shape[1])
    want = g["avg_rho"]
    for got, col in zip((res.fold, res.iener, res.m1, res.m2), range(4)):
        assert np.array_equal(got, want[:, col].astype(np.int32))
    ref = (want[:, 4] + 1j * want[:, 5]).astype(np.complex64)
    assert np.array_equal(res.rho.view(np.uint32), ref.view(np.uint32))  # complex(sp), bit for bit
    if "avg_sigma" in g:
        sig = g["sigma"].reshape((n_dirs, n_k) + g["sigma"].shape[1:])
        assert np.array_equal(unfolder.symm_average(w, sig), g["avg_sigma"])
        pr = g["spin_proj"].reshape((n_dirs, n_k) + g["spin_proj"].shape[1:])
        assert np.array_equal(unfolder.symm_average(w, pr), g["avg_spin_proj"])


@pytest.mark.parametrize("name", [c for c in CASES if c.startswith("symm_avg")])
def test_unfold_then_average_end_to_end(unfolder, gold, name):
    """The whole post-read flow on the device for one SC-kpt: unfold every (direction, pc-kpt), rho,
    then the averages -- against the reference's averaged output."""
    from bandup_b200.unfold import UnfoldDensityOps
    g = case(gold, name)
    configure(unfolder, g)
    n_dirs = int(g["n_dirs"])
    n_fold = len(g["n_selected"])
    n_k = n_fold // n_dirs
    res = unfolder.perform_unfolding(wavefunction(g), folding_G=g["folding_G"], ikpt=1)
    rho = unfolder.calc_rho(1)
    w = g["dir_weights"]
    avg = unfolder.symm_average(w, res.dN.reshape(n_dirs, n_k, -1))
    assert rel(avg, g["avg_delta_N"]) <= REL
    ops = []
    for d in range(n_dirs):
        s = (rho.fold >= d * n_k) & (rho.fold < (d + 1) * n_k)
        ops.append(UnfoldDensityOps(0, rho.fold[s] - d * n_k, rho.iener[s], rho.m1[s], rho.m2[s], rho.rho[s]))
    out = unfolder.symm_average_rho(w, ops, avg, g["weights"].shape[1])
    want = g["avg_rho"]
    assert np.array_equal(np.stack([out.fold, out.iener, out.m1, out.m2], axis=1), want[:, :4].astype(np.int32))
    assert np.max(np.abs(out.rho - (want[:, 4] + 1j * want[:, 5]))) <= 2e-6


@pytest.mark.parametrize("name", [c for c in CASES if "wavecar" in c])
def test_wavecar_file_to_delta_N_against_the_reference_run(unfolder, gold, name, tmp_path):
    """The reference's whole loop body on an identical synthetic WAVECAR: its own read_wavecar (records,
    spinor decision, 2m/hbar^2 adjustment), renormalisation, selection, perform_unfolding -- against
    the file handed to the library as raw band records with the plane-wave list made on the device."""
    from bandup_b200.wavecar import Wavecar
    g = case(gold, name)
    path = tmp_path / "WAVECAR"
    path.write_bytes(g["wavecar_bytes"].tobytes())
    wc = Wavecar(path)
    unfolder.set_complex128(wc.double)
    assert wc.double == bool(g["dp_build"])
    M = unfolder.verify_commens(g["b_pc"], wc.B_matrix)               # the SC lattice as the file gives it
    assert np.array_equal(M, np.rint(g["M"]).astype(np.int32))
    eg = g["energy_grid"]
    unfolder.set_energy_grid(eg[0], eg[-1], eg[1] - eg[0])
    ik = int(g["wavecar_ikpt"])
    nsp, npw = wc.submit_to(unfolder, 5, ik, g["folding_G"])
    assert (nsp, npw) == (int(g["n_spinor"]), len(g["g_frac"]))
    assert np.array_equal(unfolder.fetch_gvecs(5, npw), g["g_frac"])
    res = unfolder.fetch(5, want_selected=True)
    assert np.array_equal(res.n_selected, g["n_selected"])
    assert np.array_equal(np.concatenate(res.selected_coeff_indices), g["selected"])
    assert rel(res.spectral_weight, g["weights"]) <= REL
    assert rel(res.dN, g["delta_N"]) <= REL
    ops = unfolder.calc_rho(5)
    want = g["rho"]
    assert np.array_equal(np.stack([ops.fold, ops.iener, ops.m1, ops.m2], axis=1), want[:, :4].astype(np.int32))
    assert np.max(np.abs(ops.rho - (want[:, 4] + 1j * want[:, 5]))) <= 2e-6
    if nsp == 2:
        assert np.max(np.abs(unfolder.pauli_vectors(5, len(g["n_selected"])) - g["sigma"])) <= 5e-6


@pytest.mark.parametrize("name", json.loads((GOLD / "fortran_golden.json").read_text())["gur_cases"])
def test_fold_map_against_the_reference_run(unfolder, gold, name):
    """K7 (bandup_gpu_fold_map) against the reference's get_geom_unfolding_relations run: the SC-kpt
    every pc-kpt folds onto, skipped SC-kpts, and the tolerance the escalation ended on."""
    g = case(gold, name)
    unfolder.verify_commens(g["b_pc"], g["B_sc"])
    fs, fe, tol = unfolder.get_geom_unfolding_relations(g["pckpts_cart"], g["sckpts_cart"],
                                                        n_sckpts_to_skip=int(g["n_sckpts_to_skip"]))
    assert np.array_equal(fs, g["fold_sckpt"]) and np.array_equal(fe, g["fold_sckpt"])
    assert abs(tol - (1e-5 + (int(g["n_attempts"]) - 1) * 0.05e-5)) < 1e-15
