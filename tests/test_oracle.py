"""CPU tests of the oracle itself: C restatement vs the independent NumPy restatement,
plus the analytic invariants that stand in for the reference's missing test-suite.
(The reference ships no golden vectors for the Fortran path; tests/test_fortran_golden.py holds
the oracle against vectors made by executing the reference's sources with oracle/f90interp.py.)"""
import math

import numpy as np
import pytest

from bandup_b200 import synth
from oracle import oracle as O
from oracle import oracle_np as ON

SCEN = [("graphene4x4", 0.2), ("si333_vacancy", 0.25), ("mos2_8x8_spinor", 0.08), ("sige555", 0.05)]


@pytest.mark.parametrize("name,scale", SCEN)
def test_commensurability_matrix(name, scale):
    sc = synth.make_scenario(name, scale=scale)
    ok, M, iM = O.commensurate(sc.b_pc, sc.B_sc)
    ok2, M2, iM2 = ON.commensurate(sc.b_pc, sc.B_sc)
    assert ok and ok2
    assert np.array_equal(iM, iM2) and np.array_equal(iM, sc.M)
    assert np.max(np.abs(M - iM)) < 1e-9
    # A_i = sum_j M_ij a_j
    assert np.allclose(iM @ sc.a_pc, sc.A_sc, atol=1e-9)


def test_tutorial_graphene_rectangular_cell():
    """the tutorial's rectangular graphene SC gives M=[[2,2,0],[-3,3,0],[0,0,1]]."""
    a = 2.467 * np.array([[0.866025403, -0.5, 0.0], [0.866025403, 0.5, 0.0], [0.0, 0.0, 6.080259424]])
    A = 2.467 * np.array([[3.464101614, 0.0, 0.0], [0.0, 3.0, 0.0], [0.0, 0.0, 6.080259424]])
    ok, M, iM = O.commensurate(O.rec_latt(a), O.rec_latt(A))
    assert ok
    assert np.array_equal(iM, [[2, 2, 0], [-3, 3, 0], [0, 0, 1]])
    assert np.allclose(O.rec_latt(a), ON.rec_latt(a), rtol=1e-14)


def test_not_commensurate():
    a = synth.hex_cell(2.467, 15.0)
    A = np.diag([4.0, 4.1, 1.0]) @ a
    ok, _, _ = O.commensurate(O.rec_latt(a), O.rec_latt(A))
    assert not ok


@pytest.mark.parametrize("name,scale", SCEN)
def test_gvec_enumeration_matches_generator(name, scale):
    sc = synth.make_scenario(name, scale=scale)
    for iK in (0, sc.n_sc_kpts // 2, sc.n_sc_kpts - 1):
        gf, gc = O.gen_gvecs(sc.sc_kpts_frac[iK], sc.ecut, sc.B_sc)
        assert np.array_equal(gf, sc.gvecs(iK))
        assert np.allclose(gc, gf @ sc.B_sc, rtol=1e-13, atol=1e-13)
        # reader order: first wave is G=0, ig1 runs 0..n then -n..-1
        assert tuple(gf[0]) == (0, 0, 0)


@pytest.mark.parametrize("name,scale", SCEN)
def test_float_selection_equals_integer_cosets(name, scale):
    """The reference's tolerance test (vec_in_latt) and the integer coset test pick the same
    plane waves; C and NumPy restatements agree index for index."""
    sc = synth.make_scenario(name, scale=scale)
    Binv = np.linalg.inv(sc.B_sc)
    _, _, iM = O.commensurate(sc.b_pc, sc.B_sc)
    for iK in sorted(set([0, sc.n_sc_kpts // 3, sc.n_sc_kpts - 1])):
        gf, gc = O.gen_gvecs(sc.sc_kpts_frac[iK], sc.ecut, sc.B_sc)
        for fG in sc.folding_G(iK):
            sel, tol = O.select_coeffs(gc, fG, sc.b_pc)
            sel_np, tol_np = ON.select_coeffs(gc, fG, sc.b_pc)
            g0 = np.rint(fG @ Binv).astype(np.int32)
            assert np.max(np.abs(g0 @ sc.B_sc - fG)) < 1e-9
            sel_int = O.select_coeffs_int(gf, g0, iM)
            assert abs(tol - 1e-5) < 1e-15 and abs(tol_np - 1e-5) < 1e-15
            assert np.array_equal(sel, sel_np)
            assert np.array_equal(sel, sel_int)
            assert len(sel) > 0 and np.all(np.diff(sel) > 0)


def test_selection_tolerance_escalation():
    """A folding vector 4.55e-5 off the SC lattice is only accepted after the tolerance grows
    (band_unfolding_routines_mod.f90:582-601)."""
    sc = synth.make_scenario("graphene4x4", scale=0.1)
    gf, gc = O.gen_gvecs(sc.sc_kpts_frac[1], sc.ecut, sc.B_sc)
    fG = sc.folding_G(1)[0]
    sel0, _ = O.select_coeffs(gc, fG, sc.b_pc)
    sel, tol = O.select_coeffs(gc, fG + np.array([4.55e-5, 0.0, 0.0]), sc.b_pc)
    sel_np, tol_np = ON.select_coeffs(gc, fG + np.array([4.55e-5, 0.0, 0.0]), sc.b_pc)
    assert np.array_equal(sel, sel0) and np.array_equal(sel, sel_np)
    assert 4.55e-5 <= tol <= 4.7e-5 and abs(tol - tol_np) < 1e-12
    none, tol_max = O.select_coeffs(gc, fG + np.array([3e-3, 0.0, 0.0]), sc.b_pc)
    assert len(none) == 0 and tol_max > 1e-3


def test_vec_in_latt_cases():
    b = synth.rec_latt(synth.fcc_cell(5.43))
    for v, want in [(2 * b[0] - 3 * b[2], True), (0.5 * b[0], False), (b[1] + 2e-6, True),
                    (b[1] + np.array([2e-5, 0, 0]), False), (np.zeros(3), True), (-b[0] - b[1] - b[2], True)]:
        assert O.vec_in_latt(v, b) == want
        assert ON.vec_in_latt(v, b) == want


def test_real_seq_and_grid():
    g = O.real_seq(-20.0 + 1.2345, 5.0 + 1.2345, 0.05)
    assert len(g) == 501 and np.array_equal(g, ON.real_seq(-20.0 + 1.2345, 5.0 + 1.2345, 0.05))
    assert g[0] == -20.0 + 1.2345 and g[7] == (-20.0 + 1.2345) + 7.0 * 0.05


def test_lambda_box_edges():
    """sigma = epsilon turns lambda into a box: 1 inside, exactly 1/2 on an edge, 0 outside."""
    eps = np.finfo(float).eps
    E, de = 0.35, 0.05
    assert O.lam(E, E, de, eps) == 1.0
    assert O.lam(E, E + 0.5 * de, de, eps) == 0.5
    assert O.lam(E, E - 0.5 * de, de, eps) == 0.5
    assert O.lam(E, E + 0.5 * de + 1e-9, de, eps) == 0.0
    assert O.lam(E, E + 0.024, de, eps) == 1.0
    for args in [(E, E + 0.01, de, eps), (E, E + 0.03, de, 0.02), (E, E - 0.06, de, 0.05)]:
        assert abs(O.lam(*args) - ON.lam(*args)) < 1e-15
    # Gaussian smearing integrates to ~1 over all bins
    grid = O.real_seq(-1.0, 1.0, 0.05)
    tot = sum(O.lam(e, 0.013, 0.05, 0.03) for e in grid)
    assert abs(tot - 1.0) < 1e-12


@pytest.mark.parametrize("name,scale", [("graphene4x4", 0.12), ("mos2_8x8_spinor", 0.05)])
def test_renorm_weights_deltaN_c_vs_numpy(name, scale):
    sc = synth.make_scenario(name, scale=scale)
    iK = 0
    coeffs, gf, ener = synth.make_wavefunction(sc, iK)
    gc = gf @ sc.B_sc
    for mode in (0, 1):
        c_c, n_c = O.renormalize(coeffs, mode)
        c_n, n_n = ON.renormalize(coeffs, mode)
        assert np.allclose(n_c, n_n, rtol=1e-12 if mode else 0, atol=0) if mode else np.array_equal(n_c, n_n)
        assert np.array_equal(c_c, c_n)
    c_ren, norms = O.renormalize(coeffs, 1)
    grid = O.real_seq(*sc.energy_window())
    total = np.zeros(sc.n_bands)
    for fG in sc.folding_G(iK):
        sel, _ = O.select_coeffs(gc, fG, sc.b_pc)
        w_c = O.spectral_weights(c_ren, sel)
        w_n = ON.spectral_weights(c_ren, sel)
        # abs() of a complex64 is rounded to fp32: hypotf (C) and np.abs may differ by 1 ulp per term
        assert np.allclose(w_c, w_n, rtol=2e-7, atol=0)
        dn_c = O.delta_N(grid, ener, w_c)
        dn_n = ON.delta_N(grid, ener, w_n)
        assert np.allclose(dn_c, dn_n, rtol=2e-7, atol=1e-300)
        # sum_E delta_N == sum of weights of bands inside the window (box lambda)
        inside = (ener > grid[0] - 0.025) & (ener < grid[-1] + 0.025)
        assert abs(dn_c.sum() - w_c[inside].sum()) < 1e-9
        dn_g = O.delta_N(grid, ener, w_c, smearing=0.04)
        dn_gn = ON.delta_N(grid, ener, w_n, smearing=0.04)
        assert np.allclose(dn_g, dn_gn, rtol=2e-7, atol=1e-14)
        total += w_c
    assert np.all(total <= 1.0 + 1e-6)


def test_sum_over_all_cosets_is_one():
    """After renormalisation the weights of the det(M) cosets add up to 1 for every band."""
    sc = synth.make_scenario("graphene4x4", scale=0.1)
    coeffs, gf, ener = synth.make_wavefunction(sc, 2)
    gc = gf @ sc.B_sc
    c_ren, _ = O.renormalize(coeffs, 1)
    # one representative folding vector per coset: n1*B1 + n2*B2, n in 0..3
    total = np.zeros(sc.n_bands)
    nsel = 0
    for n1 in range(4):
        for n2 in range(4):
            fG = n1 * sc.B_sc[0] + n2 * sc.B_sc[1]
            sel, _ = O.select_coeffs(gc, fG, sc.b_pc)
            nsel += len(sel)
            total += O.spectral_weights(c_ren, sel)
    assert nsel == gf.shape[0]
    assert np.allclose(total, 1.0, rtol=0, atol=2e-6)


def test_perfect_supercell_weights_are_zero_or_one():
    """Bands built from a single coset unfold to P in {0,1} and delta_N counts degeneracies."""
    sc = synth.make_scenario("graphene4x4", scale=0.1)
    iK = 1
    _, gf, _ = synth.make_wavefunction(sc, iK)
    gc = gf @ sc.B_sc
    lab = synth.coset_labels(gf, sc.M)
    rng = np.random.default_rng(5)
    nb = 6
    fG = sc.folding_G(iK)[0]
    sel, _ = O.select_coeffs(gc, fG, sc.b_pc)
    target = lab[sel[0] - 1]
    coeffs = np.zeros((nb, gf.shape[0], 1), dtype=np.complex64)
    for m in range(nb):
        use = target if m % 2 == 0 else (target + 1) % (lab.max() + 1)
        idx = np.nonzero(lab == use)[0]
        coeffs[m, idx, 0] = (rng.standard_normal(len(idx)) + 1j * rng.standard_normal(len(idx)))
    ener = np.array([-3.0, -3.0, -1.0, -1.0, 0.5, 2.0])
    w, dN, ns, _ = O.unfold_kpt(coeffs, gc, ener, sc.b_pc, fG[None, :], O.real_seq(-5.0, 5.0, 0.5))
    assert np.allclose(w[0], [1, 0, 1, 0, 1, 0], atol=1e-6)
    assert abs(dN[0].sum() - 3.0) < 1e-5
    assert ns[0] == len(sel)


def test_faithful_fp32_norm_differs_from_exact():
    """the reference's single-precision norm is itself ~1e-6 off; both modes exist."""
    sc = synth.make_scenario("si333_vacancy", scale=0.3)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    _, n0 = O.renormalize(coeffs[:16], 0)
    _, n1 = O.renormalize(coeffs[:16], 1)
    rel = np.abs(n0 - n1) / n1
    assert rel.max() < 1e-4 and rel.max() > 0.0


def test_calc_rho_c_vs_numpy_and_trace():
    sc = synth.make_scenario("graphene4x4", scale=0.2)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    gc = gf @ sc.B_sc
    c_ren, _ = O.renormalize(coeffs, 1)
    grid = O.real_seq(*sc.energy_window())
    fG = sc.folding_G(0)[0]
    sel, _ = O.select_coeffs(gc, fG, sc.b_pc)
    w = O.spectral_weights(c_ren, sel)
    dN = O.delta_N(grid, ener, w)
    de = grid[1] - grid[0]
    nb = sc.n_bands
    top_bin = int(np.argmin(np.abs(grid - ener[-1])))
    cand = [i for i in np.argsort(-dN) if dN[i] >= 1e-3 and i != top_bin]
    ie = int(cand[0])
    rho_c = O.calc_rho(c_ren, sel, ener, dN[ie], grid[ie], de)
    rho_n = ON.calc_rho(c_ren, sel, ener, dN[ie], grid[ie], de)
    assert np.allclose(rho_c, rho_n, rtol=2e-5, atol=2e-6)
    # Tr rho ~ 1 (io_routines_mod.f90:1832-1839 prints it)
    assert abs(np.trace(rho_c).real - 1.0) < 1e-4
    assert np.allclose(rho_c, rho_c.conj().T)
    # Reference quirk (band_unfolding_routines_mod.f90:1096): m1 stops at nb-1, so a bin that
    # holds only the top band yields an all-zero rho.
    if dN[top_bin] >= 1e-3 and np.sum(np.abs(ener - grid[top_bin]) <= 0.5 * de) == 1:
        assert not np.any(O.calc_rho(c_ren, sel, ener, dN[top_bin], grid[top_bin], de))
    with pytest.raises(ZeroDivisionError):
        O.calc_rho(c_ren, sel, ener, 0.0, grid[ie], de)


def test_synth_fill_is_deterministic():
    a = O.synth_fill(1001, 3, 7, 2)
    b = O.synth_fill(1001, 3, 7, 2)
    c = O.synth_fill(1001, 3, 7, 3)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.all(np.abs(a.real) <= 1.0) and a.std() > 0.05


def test_spectral_function_c_vs_numpy():
    rng = np.random.default_rng(2)
    grid = O.real_seq(-3.0, 2.0, 0.05)
    ener = np.sort(rng.uniform(-4.0, 3.0, 40))
    w = rng.uniform(0.0, 1.0, 40)
    for sd in (0.025, 0.1):
        a = O.spectral_function(grid, ener, w, sd)
        b = ON.spectral_function(grid, ener, w, sd)
        assert np.allclose(a, b, rtol=1e-12, atol=1e-300)
    assert np.array_equal(O.spectral_function(grid, ener, w), O.spectral_function(grid, ener, w, 0.025))
    # integrates to the total weight of the bands well inside the window
    inside = (ener > -2.5) & (ener < 1.5)
    fine = O.real_seq(-4.5, 3.5, 0.002)
    assert abs(np.sum(O.spectral_function(fine, ener[inside], w[inside], 0.025)) * 0.002 - w[inside].sum()) < 1e-6


# ---- symmetry averaging over the needed directions (band_unfolding_routines_mod.f90:789-1043) ----
def test_symm_average_c_equals_numpy_twin():
    from tests._symm_cases import make_case
    case = make_case()
    a = O.symm_average(case["weights"], case["dN"])
    b = ON.symm_average(case["weights"], case["dN"])
    assert np.array_equal(a, b)
    one = O.symm_average([1.0], case["dN"][:1])              # -no_symm_avg: one direction of weight 1
    assert np.array_equal(one, case["dN"][0])


@pytest.mark.parametrize("seed,n_dirs,equal", [(7, 4, False), (8, 1, False), (9, 6, True)])
def test_symm_average_rho_c_equals_numpy_twin(seed, n_dirs, equal):
    from tests._symm_cases import make_case
    case = make_case(seed=seed, n_dirs=n_dirs, equal_weights=equal)
    avg = O.symm_average(case["weights"], case["dN"])
    ref = ON.symm_average_rho(case["weights"], case["per_dir"], avg)
    got = O.symm_average_rho(case["weights"], case["off"], *case["cols"], avg, case["nb"])
    assert len(ref) == len(got[0]) > 0
    for i, r in enumerate(ref):
        assert r[:4] == tuple(int(got[c][i]) for c in range(4))
        assert r[4] == got[4][i]                             # bit-equal complex64
    # energies below the averaged-delta_N cut carry no operator
    kept = {(int(k), int(e)) for k, e in zip(got[0], got[1])}
    assert all(avg[k, e - 1] >= ON.MIN_DN_SYMM_AVG for k, e in kept)
    if n_dirs == 1:                                          # weight 1: the surviving operators unchanged
        k, e, m1, m2, rho = case["cols"]
        keep = avg[k, e - 1] >= ON.MIN_DN_SYMM_AVG
        assert np.array_equal(got[4], rho[keep]) and np.array_equal(got[2], m1[keep])


def test_dp_build_oracle_c_equals_numpy_and_tracks_the_sp_build():
    """The `kind_cplx_coeffs = dp` restatement (complex128 coefficients, everything in double):
    C == NumPy twin to rounding; fed with float-representable coefficients it agrees with the
    default single-precision build to the latter's own rounding (~1e-6)."""
    from bandup_b200 import synth
    from oracle import oracle as O
    from oracle import oracle_np as ON
    for name, scale in (("graphene4x4", 0.15), ("mos2_8x8_spinor", 0.05)):
        sc = synth.make_scenario(name, scale=scale)
        coeffs, gf, ener = synth.make_wavefunction(sc, 1)
        gc = gf @ sc.B_sc
        fG = sc.folding_G(1)
        grid = O.real_seq(*sc.energy_window())
        rng = np.random.default_rng(5)
        c128 = coeffs.astype(np.complex128) * (1.0 + 1e-9 * rng.standard_normal(coeffs.shape))   # not float-representable
        w, dN, ns, norms = O.unfold_kpt_dp(c128, gc, ener, sc.b_pc, fG, grid)
        w2, dN2, ns2, norms2 = ON.unfold_kpt_dp(c128, gc, ener, sc.b_pc, fG, grid)
        assert np.array_equal(ns, ns2) and ns.min() > 0
        assert np.allclose(w, w2, rtol=1e-12, atol=1e-300) and np.allclose(norms, norms2, rtol=1e-12)
        assert np.allclose(dN, dN2, rtol=1e-12, atol=1e-300)
        w_sp, dN_sp, ns_sp, _ = O.unfold_kpt(coeffs, gc, ener, sc.b_pc, fG, grid, norm_mode=1)
        w_dp, _, ns_dp, _ = O.unfold_kpt_dp(coeffs.astype(np.complex128), gc, ener, sc.b_pc, fG, grid)
        assert np.array_equal(ns_sp, ns_dp)
        assert np.max(np.abs(w_sp - w_dp) / np.maximum(np.abs(w_dp), 1e-12)) < 5e-6
