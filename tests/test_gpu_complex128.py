"""GPU parity of the double-precision coefficient path -- the reference's `kind_cplx_coeffs = dp`
build (constants_and_types_mod.f90:60), the one a double-precision WAVECAR (nprec = 45210) needs
(read_wavecar_mod.f90:363-377): complex128 blocks through the C ABI against the dp oracle.  The
coefficients are NOT float-representable, and weights / norms / delta_N must agree to 1e-10 --
a level only a true fp64 path reaches (the complex64 path's own rounding is ~1e-7)."""
import numpy as np
import pytest

from bandup_b200 import synth

pytestmark = pytest.mark.gpu

TOL_DP = 1e-10


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-14))) if a.size else 0.0


def _c128(coeffs, seed):
    rng = np.random.default_rng(seed)
    return coeffs.astype(np.complex128) * (1.0 + 1e-7 * rng.standard_normal(coeffs.shape)) + \
        1e-9j * rng.standard_normal(coeffs.shape)


@pytest.mark.parametrize("name,scale,iK,trim", [("graphene4x4", 0.2, 1, 0), ("graphene4x4", 0.2, 0, 1),
                                                ("si333_vacancy", 0.25, 0, 0), ("mos2_8x8_spinor", 0.08, 1, 0),
                                                ("mos2_8x8_spinor", 0.08, 0, 1), ("sige555", 0.05, 2, 0)])
@pytest.mark.parametrize("variant", [1, 2])
def test_complex128_unfold_matches_dp_oracle(unfolder, name, scale, iK, trim, variant):
    from bandup_b200._capi import CONST
    from bandup_b200.unfold import LAYOUT_VASP_RECORD, PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario(name, scale=scale)
    unfolder.verify_commens(sc.b_pc, sc.B_sc)
    grid = unfolder.set_energy_grid(*sc.energy_window())
    coeffs, gf, ener = synth.make_wavefunction(sc, iK)
    if trim:                                            # an odd / different plane-wave count
        coeffs, gf = np.ascontiguousarray(coeffs[:, :-trim]), gf[:-trim]
    c = _c128(coeffs, 7 + iK)
    nb, n_pw, nsp = c.shape
    fG = sc.folding_G(iK)
    unfolder.set_complex128(True)
    unfolder.set_option(CONST["BANDUP_GPU_OPT_K2_VARIANT"], variant)
    res = unfolder.perform_unfolding(PwWavefunction(c, gf, ener), folding_G=fG, ikpt=1)
    w, dN, ns, norms = O.unfold_kpt_dp(c, gf @ sc.B_sc, ener, sc.b_pc, fG, grid)
    assert res.spectral_weight.shape == w.shape and np.array_equal(res.n_selected, ns)
    for f in range(fG.shape[0]):
        sel, _ = O.select_coeffs(gf @ sc.B_sc, fG[f], sc.b_pc)
        assert np.array_equal(res.selected_coeff_indices[f], sel)
    assert rel_err(res.band_norms, norms) <= TOL_DP
    assert rel_err(res.spectral_weight, w) <= TOL_DP
    assert rel_err(res.dN, dN) <= TOL_DP
    # the complex64 restatement of the same numbers is 1e-8..1e-6 away: the tolerance above is a real fp64 claim
    w_sp, _, _, _ = O.unfold_kpt(c.astype(np.complex64), gf @ sc.B_sc, ener, sc.b_pc, fG, grid, norm_mode=1)
    assert rel_err(w_sp, w) > 10 * TOL_DP
    if nsp == 2:                                        # raw-record layout: [band][spinor][pw], padded stride
        nrecl = 2 * n_pw * 16 + 48
        rec = np.zeros((nb, nrecl), dtype=np.uint8)
        rec[:, :2 * n_pw * 16] = np.ascontiguousarray(c.transpose(0, 2, 1)).reshape(nb, 2 * n_pw).view(np.uint8)
        r2 = unfolder.perform_unfolding(PwWavefunction(rec, gf, ener, n_pw=n_pw, n_bands=nb, n_spinor=2,
                                                       layout=LAYOUT_VASP_RECORD, band_stride_bytes=nrecl),
                                        folding_G=fG, ikpt=2, want_selected=False)
        assert rel_err(r2.spectral_weight, w) <= TOL_DP and rel_err(r2.dN, dN) <= TOL_DP
    unfolder.set_option(CONST["BANDUP_GPU_OPT_K2_VARIANT"], 0)
    unfolder.set_complex128(False)


def test_complex128_many_cosets_device_entry_and_errors(unfolder):
    """40 pc-kpts (shared-memory bins) through the device-resident entry point; misaligned blocks and
    strides that are not multiples of 16 bytes are refused; switching back restores complex64."""
    import torch
    from bandup_b200.unfold import BandupGpuError, LAYOUT_FORTRAN_INMEM, PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario("si333_vacancy", scale=0.2)
    unfolder.verify_commens(sc.b_pc, sc.B_sc)
    grid = unfolder.set_energy_grid(*sc.energy_window())
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    c = _c128(coeffs, 3)
    nb, n_pw, _ = c.shape
    rng = np.random.default_rng(3)
    g0 = rng.integers(-4, 5, size=(40, 3)).astype(np.int32)
    fG = g0.astype(np.float64) @ sc.B_sc
    w, dN, ns, norms = O.unfold_kpt_dp(c, gf @ sc.B_sc, ener, sc.b_pc, fG, grid)
    unfolder.set_complex128(True)
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    d_c = torch.from_numpy(c.view(np.float64).reshape(nb, n_pw, 2)).to(dev)
    d_g, d_e = torch.from_numpy(gf).to(dev), torch.from_numpy(ener).to(dev)
    nf = g0.shape[0]
    d_w = torch.zeros((nf, nb), dtype=torch.float64, device=dev)
    d_dn = torch.zeros((nf, len(grid)), dtype=torch.float64, device=dev)
    d_ns = torch.zeros(nf, dtype=torch.int32, device=dev)
    ws = unfolder.workspace_bytes(1, n_pw, nb, nf)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    unfolder.unfold_device(stream, 1, n_pw, nb, LAYOUT_FORTRAN_INMEM, n_pw * 16, d_c.data_ptr(), d_g.data_ptr(),
                           d_e.data_ptr(), g0, d_w.data_ptr(), d_dn.data_ptr(), d_ns.data_ptr(), d_ws.data_ptr(), ws)
    torch.cuda.synchronize()
    assert np.array_equal(d_ns.cpu().numpy(), ns)
    assert rel_err(d_w.cpu().numpy(), w) <= TOL_DP and rel_err(d_dn.cpu().numpy(), dN) <= TOL_DP
    with pytest.raises(BandupGpuError) as ei:           # stride not a multiple of 16
        unfolder.unfold_device(stream, 1, n_pw, nb, LAYOUT_FORTRAN_INMEM, n_pw * 16 + 8, d_c.data_ptr(), d_g.data_ptr(),
                               d_e.data_ptr(), g0, d_w.data_ptr(), d_dn.data_ptr(), d_ns.data_ptr(), d_ws.data_ptr(), ws)
    assert ei.value.code == 1
    with pytest.raises(BandupGpuError) as ei:           # block not 16-byte aligned
        unfolder.unfold_device(stream, 1, n_pw - 1, nb, LAYOUT_FORTRAN_INMEM, n_pw * 16, d_c.data_ptr() + 8,
                               d_g.data_ptr(), d_e.data_ptr(), g0, d_w.data_ptr(), d_dn.data_ptr(), d_ns.data_ptr(),
                               d_ws.data_ptr(), ws)
    assert ei.value.code == 1
    unfolder.set_complex128(False)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=g0, want_selected=False)
    w_sp, _, ns_sp, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, fG, grid, norm_mode=1)
    assert np.array_equal(res.n_selected, ns_sp) and rel_err(res.spectral_weight, w_sp) <= 1e-6


def test_complex128_rho_and_pauli_vectors(unfolder):
    """The projected variant on complex128 blocks: rho and Re Tr(rho sigma) are formed from the
    renormalised coefficients rounded to complex64 (their outputs are single precision in the ABI),
    so they must agree with the complex64 path fed the rounded coefficients to its own tolerance."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("mos2_8x8_spinor", scale=0.06)
    unfolder.verify_commens(sc.b_pc, sc.B_sc)
    grid = unfolder.set_energy_grid(*sc.energy_window())
    coeffs, gf, ener = synth.make_wavefunction(sc, 1)
    ener[2] = ener[1]
    ener = np.sort(ener)
    fG = sc.folding_G(1)
    nb = coeffs.shape[0]
    out_sp = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG, ikpt=1)
    rho_sp = unfolder.calc_rho(1)
    pv_sp = unfolder.pauli_vectors(1, fG.shape[0])
    unfolder.set_complex128(True)
    out_dp = unfolder.perform_unfolding(PwWavefunction(coeffs.astype(np.complex128), gf, ener), folding_G=fG, ikpt=2)
    rho_dp = unfolder.calc_rho(2)
    pv_dp = unfolder.pauli_vectors(2, fG.shape[0])
    unfolder.set_complex128(False)
    assert rel_err(out_dp.spectral_weight, out_sp.spectral_weight) <= 1e-6
    keys = {(int(f), int(i)) for f, i in zip(rho_sp.fold, rho_sp.iener)}
    assert keys == {(int(f), int(i)) for f, i in zip(rho_dp.fold, rho_dp.iener)} and len(keys) > 3
    for f, ie in keys:
        a, b = rho_sp.dense(f, ie, nb), rho_dp.dense(f, ie, nb)
        sure = (np.abs(a) >= 1.2e-5) | (np.abs(b) >= 1.2e-5)        # clear of the 1e-5 store threshold
        assert np.all(np.abs(a[sure] - b[sure]) <= 2e-6 * max(1.0, np.abs(a).max()))
    assert np.max(np.abs(pv_dp - pv_sp)) <= 5e-6 and np.abs(pv_sp).max() > 1e-3


@pytest.mark.parametrize("name,scale", [("graphene4x4", 0.1), ("mos2_8x8_spinor", 0.05)])
def test_double_precision_wavecar_file(unfolder, tmp_path, name, scale):
    """A WAVECAR_double (nprec = 45210, complex128 records -- the file a `kind_cplx_coeffs = dp` build of
    the reference reads, read_wavecar_mod.f90:363-377) through the file-level path: raw records
    uploaded as they lie, plane-wave list made on the device, against the dp oracle."""
    from bandup_b200.wavecar import Wavecar, unfold_wavecar, write_wavecar
    from oracle import oracle as O
    sc = synth.make_scenario(name, scale=scale, n_kpts=9)
    data = [synth.make_wavefunction(sc, iK) for iK in range(sc.n_sc_kpts)]
    c128 = [_c128(d[0], 11 + i) for i, d in enumerate(data)]
    path = tmp_path / "WAVECAR_double"
    write_wavecar(path, sc.A_sc, sc.ecut, list(sc.sc_kpts_frac), c128, [d[2] for d in data], extra_pad=1, double=True)
    wc = Wavecar(path)
    assert wc.double and wc.elem_bytes == 16
    assert np.array_equal(wc.coefficients(1), c128[0])
    got, grid = unfold_wavecar(unfolder, wc, sc.b_pc, sc.pc_kpts_cart, *sc.energy_window())
    host_g, _ = unfold_wavecar(unfolder, wc, sc.b_pc, sc.pc_kpts_cart, *sc.energy_window(), device_gvecs=False)
    assert np.array_equal(got.dN, host_g.dN)
    for iK in range(sc.n_sc_kpts):
        _, gf, ener = data[iK]
        _, dN, _, _ = O.unfold_kpt_dp(c128[iK], gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid)
        assert rel_err(got.dN[np.asarray(sc.folds[iK])], dN) <= TOL_DP
    unfolder.set_complex128(False)
