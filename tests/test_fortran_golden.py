"""The oracle (C and NumPy restatements) against golden vectors produced by EXECUTING the
reference's own Fortran sources (tests/golden/make_fortran_golden.py, through the interpreter
oracle/f90interp.py -- the build image has no Fortran compiler).  This is what pins the
Fortran-derived half of the oracle; tests/test_gpu_fortran_golden.py holds the CUDA path against
the same vectors."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from oracle import oracle_np as ON

GOLD = Path(__file__).resolve().parent / "golden"
CASES = json.loads((GOLD / "fortran_golden.json").read_text())["cases"]
GUR_CASES = json.loads((GOLD / "fortran_golden.json").read_text())["gur_cases"]
DP_CASES = [c for c in CASES if c.startswith("dp_build")]        # the reference with kind_cplx_coeffs = dp
SP_CASES = [c for c in CASES if c not in DP_CASES]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD / "fortran_golden.npz")


def case(gold, name):
    pre = name + "/"
    return {k[len(pre):]: gold[k] for k in gold.files if k.startswith(pre)}


def bands_first(c):
    """(n_spinor, n_pw, n_bands) as the Fortran array is indexed -> the oracle's (n_bands, n_pw, n_spinor)."""
    return np.ascontiguousarray(np.transpose(c, (2, 1, 0)))


def rel(a, b, floor=1e-300):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if np.size(a) else 0.0


def folds(g):
    off = np.concatenate([[0], np.cumsum(g["n_selected"])])
    return [(f, g["selected"][off[f]:off[f + 1]]) for f in range(len(g["n_selected"]))]


def test_manifest_names_the_reference_files():
    m = json.loads((GOLD / "fortran_golden.json").read_text())
    assert m["fortran_statements_executed"] > 500000
    assert set(m["reference_files_sha256"]) >= {"band_unfolding_routines_mod.f90", "io_routines_mod.f90",
                                               "crystals_mod.f90", "math_mod.f90", "lists_and_seqs_mod.f90"}
    ref = Path("/root/reference/src")
    if ref.exists():      # build container only: the fixtures belong to THIS reference tree
        import hashlib
        for f, h in m["reference_files_sha256"].items():
            assert hashlib.sha256((ref / f).read_bytes()).hexdigest() == h, f


@pytest.mark.parametrize("name", CASES)
def test_cells_and_commensurability(gold, name):
    g = case(gold, name)
    assert np.array_equal(O.rec_latt(g["a_pc"]), g["b_pc"])          # get_rec_latt, bit for bit
    assert np.array_equal(O.rec_latt(g["A_sc"]), g["B_sc"])
    assert np.allclose(ON.rec_latt(g["a_pc"]), g["b_pc"], rtol=1e-15, atol=1e-18)
    ok, M, iM = O.commensurate(g["b_pc"], g["B_sc"])
    assert ok and bool(g["commensurate"])
    assert np.array_equal(M, g["M"])                                  # transpose(matmul(b, inv(B)))
    ok2, M2 = ON.commensurate(g["b_pc"], g["B_sc"])[:2]
    assert ok2 and np.allclose(M2, g["M"], rtol=0, atol=1e-12)
    # the supercell really is M a -- up to the sign get_rec_latt drops for a left-handed cell (it divides by
    # |a1.(a2 x a3)|, crystals_mod.f90:194-197, so such a cell gets -B: the same lattice, and M changes sign)
    hand = np.sign(np.linalg.det(g["A_sc"]) * np.linalg.det(g["a_pc"]))
    assert np.allclose(hand * (iM @ g["a_pc"]), g["A_sc"], atol=1e-9)
    if name == "fcc_negative_det":
        assert hand == -1.0


def test_non_commensurate_cells_are_reported():
    ok, M, _ = O.commensurate(np.eye(3) * 1.0, np.eye(3) * 0.37)
    assert not ok


@pytest.mark.parametrize("name", CASES)
def test_plane_wave_list_is_the_readers(gold, name):
    g = case(gold, name)
    # the file of this case holds one plane wave more than the stock 2m/hbar^2 yields; read_wavecar's
    # adjustment loop (read_wavecar_mod.f90:459-472) ended on the constant + 1e-10
    c = float(np.float32(0.262465831)) + 1e-10 if name == "wavecar_spinor_adjusted_cutoff" else 0.0
    gf, gc = O.gen_gvecs(g["kpt_frac"], float(g["ecut"]), g["B_sc"], c)
    assert np.array_equal(gf, g["g_frac"])                           # count, order and indices
    assert np.array_equal(gc, g["g_cart"])                           # ig1p*b1 + ig2p*b2 + ig3p*b3, bit for bit


@pytest.mark.parametrize("name", SP_CASES)
def test_renormalisation(gold, name):
    g = case(gold, name)
    cin, want = bands_first(g["coeffs_in"]), bands_first(g["coeffs_renormalised"])
    faithful, norms_sp = O.renormalize(cin, norm_mode=0)
    # the reference's own arithmetic (sp dot_product, dp scale, rounded to sp): bit for bit
    assert np.array_equal(faithful.view(np.float32), want.view(np.float32))
    exact, norms = O.renormalize(cin, norm_mode=1)
    assert rel(norms_sp, norms, 1e-30) < 5e-6                         # the sp sum is 1e-6..5e-6 off
    assert np.max(np.abs(exact - want)) < 5e-6 * np.max(np.abs(want))
    np_faithful = ON.renormalize(cin, norm_mode=0)[0]
    assert np.array_equal(np_faithful.view(np.float32), want.view(np.float32))
    if name == "fcc_negative_det":                                    # its band 3 is all zeros: left alone (:1325)
        assert not np.any(want[2]) and norms[2] == 0.0


@pytest.mark.parametrize("name", CASES)
def test_selection(gold, name):
    g = case(gold, name)
    Binv = np.linalg.inv(g["B_sc"])
    iM = np.rint(g["M"]).astype(np.int32)
    for f, sel in folds(g):
        got, tol = O.select_coeffs(g["g_cart"], g["folding_G"][f], g["b_pc"])
        assert np.array_equal(got, sel)                               # float vec_in_latt recipe
        assert np.array_equal(ON.select_coeffs(g["g_cart"], g["folding_G"][f], g["b_pc"])[0], sel)
        g0 = np.rint(g["folding_G"][f] @ Binv).astype(np.int32)
        assert np.array_equal(O.select_coeffs_int(g["g_frac"], g0, iM), sel)   # integer cosets (what K1 does)
        if name == "tolerance_ladder":
            assert tol > 1.05e-5                                      # the escalation loop really ran
        else:
            assert abs(tol - 1e-5) < 1e-12
    assert np.all(g["n_selected"] > 0)


@pytest.mark.parametrize("name", SP_CASES)
def test_spectral_weights_and_delta_N(gold, name):
    g = case(gold, name)
    cren = bands_first(g["coeffs_renormalised"])
    for f, sel in folds(g):
        if bool(g["perform_unfold"]):
            w = O.spectral_weights(cren, sel)
            assert rel(w, g["weights"][f]) < 1e-15                    # same terms, same order, in dp
            # the NumPy twin takes abs() with NumPy's vectorised fp32 hypot: one ulp of fp32 per term at most
            assert rel(ON.spectral_weights(cren, sel), g["weights"][f]) < 2e-7
        else:
            assert np.array_equal(g["weights"][f], np.ones_like(g["weights"][f]))   # -dont_unfold (:1346-1351)
        dn = O.delta_N(g["energy_grid"], g["band_energies"], g["weights"][f])
        assert np.max(np.abs(dn - g["delta_N"][f])) <= 1e-15 * max(1.0, np.max(g["delta_N"][f]))
        dn_np = ON.delta_N(g["energy_grid"], g["band_energies"], g["weights"][f])
        assert np.max(np.abs(dn_np - g["delta_N"][f])) <= 1e-13 * max(1.0, np.max(g["delta_N"][f]))
    # a band exactly on a bin edge counts half in both bins (lambda = 1/2): the cases plant one
    e, grid = g["band_energies"], g["energy_grid"]
    d = grid[1] - grid[0]
    on_edge = [m for m in range(len(e)) if np.any(e[m] == grid + 0.5 * d)]
    if name != "tolerance_ladder":
        assert on_edge
    for m in on_edge:
        i = int(np.nonzero(e[m] == grid + 0.5 * d)[0][0])
        eps = float(np.finfo(np.float64).eps)          # sigma = epsilon(1.0_dp): the box
        assert O.lam(grid[i], e[m], d, eps) == 0.5 and O.lam(grid[i + 1], e[m], d, eps) == 0.5


def test_whole_kpt_entry_point_of_the_oracle(gold):
    """bo_unfold_kpt (what bench.py's CPU arm times) from the raw inputs, against the golden run."""
    for name in SP_CASES:
        g = case(gold, name)
        if not bool(g["perform_unfold"]):
            continue
        w, dn, ns, _ = O.unfold_kpt(bands_first(g["coeffs_in"]), g["g_cart"], g["band_energies"], g["b_pc"],
                                    g["folding_G"], g["energy_grid"], norm_mode=0)
        assert np.array_equal(ns, g["n_selected"])
        assert rel(w, g["weights"]) < 1e-15
        assert np.max(np.abs(dn - g["delta_N"])) <= 1e-15
        w1, dn1, _, _ = O.unfold_kpt(bands_first(g["coeffs_in"]), g["g_cart"], g["band_energies"], g["b_pc"],
                                     g["folding_G"], g["energy_grid"], norm_mode=1)
        assert rel(w1, g["weights"], 1e-12) < 5e-6 and np.max(np.abs(dn1 - g["delta_N"])) < 5e-6


def test_gaussian_smearing_and_spectral_function(gold):
    g = case(gold, "triclinic_scalar")
    s = float(g["smearing"])
    for f in range(len(g["n_selected"])):
        dn = O.delta_N(g["energy_grid"], g["band_energies"], g["weights"][f], smearing=s)
        assert np.max(np.abs(dn - g["delta_N_smeared"][f])) <= 1e-15
        assert np.max(np.abs(ON.delta_N(g["energy_grid"], g["band_energies"], g["weights"][f], smearing=s)
                             - g["delta_N_smeared"][f])) <= 1e-13
    sf = O.spectral_function(g["energy_grid"], g["band_energies"], g["weights"][0], float(g["sf_std"]))
    assert rel(sf, g["spectral_function"], 1e-200) < 1e-14
    sf_np = ON.spectral_function(g["energy_grid"], g["band_energies"], g["weights"][0], float(g["sf_std"]))
    assert rel(sf_np, g["spectral_function"], 1e-200) < 1e-12


def rho_lists(g, f, sel, oracle_rho):
    """The (iener, m1, m2, value) entries perform_unfolding appends for fold f (:1372-1428)."""
    grid, dn = g["energy_grid"], g["delta_N"][f]
    d = grid[1] - grid[0]
    out = []
    for ie in range(len(grid)):
        if not dn[ie] >= 1e-3:
            continue
        rho = oracle_rho(sel, dn[ie], grid[ie], d)
        for m1 in range(rho.shape[0]):
            for m2 in range(rho.shape[1]):
                if np.abs(np.complex64(rho[m1, m2])) < np.float32(1e-5):
                    continue
                out.append((ie + 1, m1 + 1, m2 + 1, rho[m1, m2]))
    return out


@pytest.mark.parametrize("name", SP_CASES)
def test_unfolding_density_operator(gold, name):
    g = case(gold, name)
    cren = bands_first(g["coeffs_renormalised"])
    n = 0
    unfold = bool(g["perform_unfold"])
    for f, sel in folds(g):
        want = g["rho"][g["rho"][:, 0] == f]
        if unfold:
            c_rho = lambda s, dn, E, d: O.calc_rho(cren, s, g["band_energies"], dn, E, d)
        else:                                                          # calc_rho's other branch (:1136-1154)
            c_rho = lambda s, dn, E, d: O.calc_rho_dont_unfold(g["band_energies"], dn, E, d)
        got = rho_lists(g, f, sel, c_rho)
        assert [(int(r[1]), int(r[2]), int(r[3])) for r in want] == [t[:3] for t in got]    # which, and in which order
        for r, t in zip(want, got):
            v = np.complex64(complex(r[4], r[5]))
            assert t[3] == v                                           # complex(sp) arithmetic, bit for bit
        got_np = rho_lists(g, f, sel, lambda s, dn, E, d: ON.calc_rho(cren, s, g["band_energies"], dn, E, d,
                                                                      perform_unfold=unfold))
        assert [t[:3] for t in got_np] == [t[:3] for t in got]
        for a, b in zip(got_np, got):
            # the NumPy twin forms the inner products in fp64: off by the reference's own sp rounding,
            # which scales with the terms of the sum (|rho| <= 1), not with a small result
            assert abs(complex(a[3]) - complex(b[3])) < 1e-6
        n += len(got)
    assert n == len(g["rho"]) > 0


def test_pauli_vectors_and_spin_projections(gold):
    g = case(gold, "hex_spinor")
    from bandup_b200.unfold import calc_spin_projections      # 3-vector algebra the host does (:1266-1286)
    cren = bands_first(g["coeffs_renormalised"])
    grid = g["energy_grid"]
    d = grid[1] - grid[0]
    nb = cren.shape[0]
    checked = 0
    for f, sel in folds(g):
        rows = g["rho"][g["rho"][:, 0] == f]
        rhos = {}
        for r in rows:
            rhos.setdefault(int(r[1]), np.zeros((nb, nb), dtype=np.complex128))[int(r[2]) - 1, int(r[3]) - 1] = \
                complex(r[4], r[5])
        # energies with delta_N >= 1e-3 but no stored element still get a (zero) Pauli vector
        for ie in np.nonzero(g["delta_N"][f] >= 1e-3)[0]:
            rhos.setdefault(int(ie) + 1, np.zeros((nb, nb), dtype=np.complex128))
        vecs = ON.unfolded_pauli_vectors(cren, g["band_energies"], grid, rhos, d)
        for ie, v in vecs.items():
            want = g["sigma"][f, ie - 1]
            assert np.max(np.abs(v - want)) < 5e-6 * max(1.0, np.max(np.abs(want)))   # sp inner products
            perp, para = calc_spin_projections(want, g["pckpts_cart"][f], np.array([0.0, 0.0, 1.0]),
                                               origin=np.zeros(3))
            assert abs(perp - g["spin_proj"][f, ie - 1, 0]) < 1e-12
            assert abs(para - g["spin_proj"][f, ie - 1, 1]) < 1e-12
            checked += 1
        untouched = np.setdiff1d(np.arange(len(grid)), np.array(sorted(vecs)) - 1)
        assert not np.any(g["sigma"][f, untouched])
    assert checked >= 3


@pytest.mark.parametrize("name", DP_CASES)
def test_double_precision_build(gold, name):
    """`kind_cplx_coeffs = dp` (constants_and_types_mod.f90:60): complex128 coefficients, every
    dot_product, abs and rescale in double precision -- the dp entry points of both oracles against
    the reference run in that configuration."""
    g = case(gold, name)
    assert g["coeffs_in"].dtype == np.complex128 and bool(g["dp_build"])
    cin = bands_first(g["coeffs_in"])
    w, dn, ns, norms = O.unfold_kpt_dp(cin, g["g_cart"], g["band_energies"], g["b_pc"], g["folding_G"],
                                       g["energy_grid"])
    assert np.array_equal(ns, g["n_selected"])
    assert rel(w, g["weights"]) < 1e-14 and np.max(np.abs(dn - g["delta_N"])) < 1e-14
    w2, dn2, ns2, norms2 = ON.unfold_kpt_dp(cin, g["g_cart"], g["band_energies"], g["b_pc"], g["folding_G"],
                                            g["energy_grid"])
    assert np.array_equal(ns2, ns) and rel(w2, g["weights"]) < 1e-13 and np.max(np.abs(dn2 - g["delta_N"])) < 1e-13
    assert rel(norms, norms2, 1e-30) < 1e-14
    # renormalised coefficients: (1/sqrt(s)) * C in double precision
    want = bands_first(g["coeffs_renormalised"])
    got = cin / np.sqrt(norms)[:, None, None]
    assert np.max(np.abs(got - want)) <= 4e-16 * np.max(np.abs(want))
    # a single-precision treatment of the same numbers is 1e-8 or more away: this really is the dp path
    w_sp = O.unfold_kpt(cin.astype(np.complex64), g["g_cart"], g["band_energies"], g["b_pc"], g["folding_G"],
                        g["energy_grid"], norm_mode=1)[0]
    assert rel(w_sp, g["weights"]) > 1e-9
    # rho in complex(dp): the NumPy twin (fp64 inner products) to 1e-13
    off = np.concatenate([[0], np.cumsum(g["n_selected"])])
    n = 0
    for f in range(len(ns)):
        sel = g["selected"][off[f]:off[f + 1]]
        rows = g["rho"][g["rho"][:, 0] == f]
        got_np = rho_lists(g, f, sel, lambda s_, dn_, E, d: ON.calc_rho(want, s_, g["band_energies"], dn_, E, d))
        assert [t[:3] for t in got_np] == [(int(r[1]), int(r[2]), int(r[3])) for r in rows]
        for t, r in zip(got_np, rows):
            assert abs(complex(t[3]) - complex(r[4], r[5])) < 1e-13
        n += len(rows)
    assert n == len(g["rho"]) > 0


WAVECAR_CASES = [c for c in CASES if "wavecar" in c]


@pytest.mark.parametrize("name", WAVECAR_CASES)
def test_wavecar_parser_against_the_reference_reader(gold, name, tmp_path):
    """The same synthetic WAVECAR through bandup_b200.wavecar (the host-side parser of the raw-record
    path) and through the reference's read_wavecar (interpreted): header fields, scalar/spinor
    decision, plane-wave list -- with the 2m/hbar^2 adjustment in one case -- energies, coefficients."""
    from bandup_b200.wavecar import Wavecar
    g = case(gold, name)
    path = tmp_path / "WAVECAR"
    path.write_bytes(g["wavecar_bytes"].tobytes())
    wc = Wavecar(path)
    ik = int(g["wavecar_ikpt"])
    assert wc.nrecl == int(g["wavecar_nrecl"]) and wc.nspin == int(g["reader_n_spin"]) == 1
    assert wc.encut == float(g["reader_encut"]) == float(g["ecut"])
    assert np.array_equal(wc.A_matrix, g["reader_A_matrix"]) and np.array_equal(wc.A_matrix, g["A_sc"])
    assert np.allclose(wc.B_matrix, g["reader_B_matrix"], rtol=1e-15, atol=0)
    # the reader rebuilds B from the file's A (twopi*cross/Vcell): the same matrix get_rec_latt gave
    assert np.allclose(g["reader_B_matrix"], g["B_sc"], rtol=1e-15, atol=1e-18)
    h = wc.kpt_header(ik)
    assert np.array_equal(h.kpt_frac_coords, g["reader_kpt_frac"])
    assert np.array_equal(h.band_energies, g["band_energies"])
    assert np.array_equal(h.band_occupations, g["reader_occupations"])
    gf, nsp = wc.gvecs(ik)
    assert nsp == int(g["n_spinor"]) == (2 if bool(g["reader_is_spinor"]) else 1)
    assert np.array_equal(gf, g["g_frac"])
    assert h.nplane == nsp * len(gf)
    assert np.array_equal(wc.coefficients(ik), bands_first(g["coeffs_in"]))      # (n_bands, n_pw, n_spinor)
    assert wc.double == bool(g["dp_build"]) == (g["coeffs_in"].dtype == np.complex128)
    if name == "wavecar_spinor_adjusted_cutoff":
        stock = O.gen_gvecs(g["kpt_frac"], float(g["ecut"]), g["B_sc"])[0]
        assert len(stock) == len(gf) - 1                                   # the loop had work to do


AVG_CASES = [c for c in CASES if c.startswith("symm_avg")]


def per_direction(g):
    """The unfolding-density operators of each needed direction, rows renumbered to the pc-kpt."""
    n_dirs = int(g["n_dirs"])
    n_k = len(g["n_selected"]) // n_dirs
    out = []
    for d in range(n_dirs):
        rows = g["rho"][(g["rho"][:, 0] >= d * n_k) & (g["rho"][:, 0] < (d + 1) * n_k)]
        out.append([(int(r[0]) - d * n_k, int(r[1]), int(r[2]), int(r[3]), np.complex64(complex(r[4], r[5])))
                    for r in rows])
    return n_dirs, n_k, out


@pytest.mark.parametrize("name", AVG_CASES)
def test_symmetry_average_over_the_needed_directions(gold, name):
    """get_delta_Ns_for_output (band_unfolding_routines_mod.f90:789-1043) as the reference ran it."""
    g = case(gold, name)
    n_dirs, n_k, ents = per_direction(g)
    w = g["dir_weights"]
    dn = g["delta_N"].reshape(n_dirs, n_k, -1)
    avg = O.symm_average(w, dn)
    assert np.array_equal(avg, g["avg_delta_N"])                       # dp, direction order: bit for bit
    assert np.array_equal(ON.symm_average(w, dn), g["avg_delta_N"])
    want = [(int(r[0]), int(r[1]), int(r[2]), int(r[3]), np.complex64(complex(r[4], r[5]))) for r in g["avg_rho"]]
    got_np = ON.symm_average_rho(w, ents, avg)
    assert [t[:4] for t in got_np] == [t[:4] for t in want]            # which elements, in which order
    assert all(a[4] == b[4] for a, b in zip(got_np, want))             # complex(sp) values, bit for bit
    cols = [np.array([e[i] for d in ents for e in d], dtype=np.int32) for i in range(4)]
    rho = np.array([e[4] for d in ents for e in d], dtype=np.complex64)
    off = np.concatenate([[0], np.cumsum([len(d) for d in ents])])
    nb = g["weights"].shape[1]
    got_c = O.symm_average_rho(w, off, *cols, rho, avg, nb)
    assert [tuple(int(c[i]) for c in got_c[:4]) for i in range(len(got_c[0]))] == [t[:4] for t in want]
    assert np.array_equal(got_c[4].view(np.uint32), np.array([t[4] for t in want], dtype=np.complex64).view(np.uint32))
    assert len(want) > 0 and len(want) < len(g["rho"])                  # the directions share elements
    if "avg_sigma" in g:
        sig = g["sigma"].reshape((n_dirs, n_k) + g["sigma"].shape[1:])
        assert np.array_equal(O.symm_average(w, sig), g["avg_sigma"])
        pr = g["spin_proj"].reshape((n_dirs, n_k) + g["spin_proj"].shape[1:])
        assert np.array_equal(O.symm_average(w, pr), g["avg_spin_proj"])


@pytest.mark.parametrize("name", AVG_CASES)
def test_unfolded_ebs_file_is_what_the_reference_writes(gold, name, tmp_path):
    """write_band_struc (io_routines_mod.f90:889-977), executed from the reference on the averaged
    quantities of the case, against bandup_files.write_band_struc (what BandUP_gpu.x and unfold_task
    write): every byte after the first header line (which carries version and time stamp)."""
    from bandup_b200 import bandup_files as BF
    g = case(gold, name)
    want = g["ebs_file"].tobytes().decode()
    n_dirs = int(g["n_dirs"])
    n_k = len(g["n_selected"]) // n_dirs
    kpath = BF.PcKptsPath([g["pckpts_cart"][:n_k]], float(g["ebs_zero_of_kpts_scale"]))   # the first needed direction
    out = tmp_path / "ebs.dat"
    spinor = "avg_sigma" in g
    BF.write_band_struc(out, kpath, g["energy_grid"], g["avg_delta_N"], e_fermi=float(g["ebs_e_fermi"]),
                        sigma=g["avg_sigma"] if spinor else None, spin_proj=g["avg_spin_proj"] if spinor else None)
    got = out.read_text()
    assert got.splitlines()[0].startswith("# File created by") and want.splitlines()[0].startswith("# File created by")
    assert got.split("\n", 1)[1] == want.split("\n", 1)[1]
    data = BF.read_band_struc(out)                                   # and it loads the way `bandup plot` loads it
    assert data.shape == (n_k * len(g["energy_grid"]), 8 if spinor else 3)


@pytest.mark.parametrize("name", GUR_CASES)
def test_fold_search_against_the_reference_run(gold, name):
    """get_geom_unfolding_relations as the reference ran it (-no_symm_sckpts): the first SC-kpt of the
    list -- the skipped ones excepted -- whose difference to the pc-kpt is an SC reciprocal lattice
    vector, with the wrapper's tolerance steps of 0.05e-5.  Brute force with the oracle's vec_in_latt."""
    g = case(gold, name)
    skip, attempts = int(g["n_sckpts_to_skip"]), int(g["n_attempts"])
    tol = 1e-5 + (attempts - 1) * 0.05e-5

    def brute(t):
        out = np.full(len(g["pckpts_cart"]), -1)
        for i, k in enumerate(g["pckpts_cart"]):
            for s_ in range(skip, len(g["sckpts_cart"])):
                if O.vec_in_latt(k - g["sckpts_cart"][s_], g["B_sc"], t):
                    out[i] = s_
                    break
        return out
    assert np.array_equal(brute(tol), g["fold_sckpt"])
    assert int(g["n_folding_pckpts"]) == int(g["n_pckpts"]) == len(g["fold_sckpt"])
    assert np.array_equal(g["folds"].sum(axis=0), np.ones(len(g["fold_sckpt"])))       # one SC-kpt each
    assert np.array_equal(g["sckpt_used"], np.isin(np.arange(len(g["sckpts_cart"])), g["fold_sckpt"]))
    assert np.array_equal(g["scoords"], g["pckpts_cart"])                              # identity operation only
    if attempts > 1:
        assert np.any(brute(tol - 0.05e-5) < 0)                                        # the attempt before failed
        assert np.all(g["fold_sckpt"] >= skip) and skip > 0


@pytest.mark.skipif(not Path("/root/reference/src").exists(),
                    reason="the reference sources exist in the build container only")
@pytest.mark.parametrize("name", ["hex_spinor", "dont_unfold"])
def test_goldens_regenerate_from_the_reference_sources(gold, name):
    """Re-executes the reference's Fortran for two of the cases and compares with the committed
    fixture, array by array, bit by bit (the generator is deterministic)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_fortran_golden", GOLD / "make_fortran_golden.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    s = [x for x in gen.SPECS if x[0] == name][0]
    out, n_statements = gen.run_case(*s[:10], **s[10])
    assert n_statements > 50000
    g = case(gold, name)
    assert set(out) == set(g)
    for k, v in out.items():
        assert np.array_equal(np.asarray(v), g[k], equal_nan=True), k


@pytest.mark.skipif(not Path("/root/reference/src").exists(),
                    reason="the reference sources exist in the build container only")
def test_fold_search_golden_regenerates(gold):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_fortran_golden", GOLD / "make_fortran_golden.py")
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out, _ = gen.run_gur_case(*gen.GUR_SPECS[0])
    g = case(gold, gen.GUR_SPECS[0][0])
    assert set(out) == set(g)
    for k, v in out.items():
        assert np.array_equal(np.asarray(v), g[k]), k


def _sample_plot_check_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_sample_plot_writer_check", GOLD / "make_sample_plot_writer_check.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not Path("/root/reference/utils/sample_input_files_task_plot").exists(),
                    reason="the reference tree exists in the build container only")
def test_interpreted_writer_reproduces_a_file_written_by_a_compiled_bandup():
    """The one output of a real (compiled) BandUP run in the reference tree,
    utils/sample_input_files_task_plot/unfolded_EBS_symmetry-averaged.dat (63 pc-kpts x 581 energies):
    real_seq + write_band_struc EXECUTED from the reference sources on its path, window and delta_N
    column print every one of its 36 603 data lines, text for text -- the interpreter's arithmetic and
    edit descriptors against a compiled build, on the routine that also wrote the golden EBS files."""
    mod = _sample_plot_check_module()
    mine, ref, n_statements = mod.run()
    assert len(ref) == 63 * 581 and n_statements > 100000
    assert mine == ref
    rec = json.loads((GOLD / "sample_plot_writer_check.json").read_text())
    assert rec["sha256_reference_file_lines"] == mod.digest(ref) == rec["sha256_interpreted_writer_lines"]


def test_product_writer_reproduces_a_file_written_by_a_compiled_bandup():
    """The same file from the product's writer (bandup_files.write_band_struc: what unfold_task and
    BandUP_gpu.x print) on the committed inputs and the file's delta_N column (sample_plot_dN.npz):
    the SHA-256 of its 36 603 data lines is the one of the compiled program's file
    (sample_plot_writer_check.json, made where the reference tree exists).  Runs anywhere."""
    mod = _sample_plot_check_module()
    z = np.load(GOLD / "sample_plot_dN.npz")
    nener, n_k = (int(x) for x in z["shape"])
    d = np.zeros(nener * n_k)
    d[z["idx"]] = z["val"].astype(np.float64)
    lines = mod.product_writer_lines(d.reshape(nener, n_k))
    rec = json.loads((GOLD / "sample_plot_writer_check.json").read_text())
    assert rec["differing_lines"] == 0 and len(lines) == rec["data_lines"] == 36603
    assert mod.digest(lines) == rec["sha256_reference_file_lines"]
