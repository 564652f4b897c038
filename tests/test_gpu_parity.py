"""GPU parity tests: every call goes through the C ABI (libbandup_gpu.so) and is compared with
the CPU oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star): selected
plane-wave indices bit-exact; spectral weights and delta_N within 1e-6 relative (abs floor
1e-12) of the exact_fp64-norm oracle."""
import ctypes as C

import numpy as np
import pytest

from bandup_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
ABS_FLOOR = 1e-12

SCEN = [("graphene4x4", 0.2), ("si333_vacancy", 0.25), ("mos2_8x8_spinor", 0.08), ("sige555", 0.05)]


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), ABS_FLOOR))) if a.size else 0.0


def oracle_kpt(sc, coeffs, gf, ener, fG, grid, smearing=0.0):
    from oracle import oracle as O
    gc = gf @ sc.B_sc
    w, dN, ns, _ = O.unfold_kpt(coeffs, gc, ener, sc.b_pc, fG, grid, smearing=smearing, norm_mode=1)
    sels = [O.select_coeffs(gc, f, sc.b_pc)[0] for f in fG]
    return w, dN, ns, sels


def setup(u, sc, smearing=None):
    M = u.verify_commens(sc.b_pc, sc.B_sc)
    assert np.array_equal(M, sc.M)
    return u.set_energy_grid(*sc.energy_window(), smearing_for_delta_function=smearing)


@pytest.mark.parametrize("name,scale", SCEN)
def test_unfold_matches_oracle(unfolder, name, scale):
    from bandup_b200.unfold import PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario(name, scale=scale)
    grid = setup(unfolder, sc)
    assert np.array_equal(grid, O.real_seq(*sc.energy_window()))
    for iK in sorted(set([0, sc.n_sc_kpts // 2, sc.n_sc_kpts - 1])):
        coeffs, gf, ener = synth.make_wavefunction(sc, iK)
        fG = sc.folding_G(iK)
        res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG, ikpt=iK + 1)
        w, dN, ns, sels = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
        assert np.array_equal(res.n_selected, ns)
        for a, b in zip(res.selected_coeff_indices, sels):
            assert np.array_equal(a, b)                     # bit-exact coset indices
        assert rel_err(res.spectral_weight, w) <= REL_TOL
        assert rel_err(res.dN, dN) <= REL_TOL
        # band norms before renormalisation (io_routines_mod.f90:1321-1324, fp64 accumulate)
        _, norms = O.renormalize(coeffs, 1)
        assert rel_err(res.band_norms, norms) <= REL_TOL
        assert not res.pckpt_cannot_be_parsed().any()


def test_all_cosets_sum_to_one(unfolder):
    """Sum over the det(M)=16 cosets of P is 1 for every band; n_selected adds up to n_pw."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("graphene4x4", scale=0.15)
    setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 2)
    g0 = np.array([[n1, n2, 0] for n1 in range(4) for n2 in range(4)], dtype=np.int32)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=g0)
    assert int(res.n_selected.sum()) == gf.shape[0]
    assert np.allclose(res.spectral_weight.sum(axis=0), 1.0, rtol=0, atol=1e-7)
    allsel = np.sort(np.concatenate(res.selected_coeff_indices))
    assert np.array_equal(allsel, np.arange(1, gf.shape[0] + 1))


def test_folds_sharing_a_coset_and_many_folds(unfolder):
    """Two pc-kpts whose folding vectors differ by a pc reciprocal vector get identical rows;
    40 folds exercise the shared-memory bins beyond a handful."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("si333_vacancy", scale=0.2)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    rng = np.random.default_rng(3)
    g0 = rng.integers(-4, 5, size=(40, 3)).astype(np.int32)
    # b_i = sum_j M_ji B_j: adding a column of M to g0 stays in the same coset
    g0[7] = g0[2] + sc.M[:, 0] - 2 * sc.M[:, 2]
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=g0)
    assert np.array_equal(res.spectral_weight[7], res.spectral_weight[2])
    assert np.array_equal(res.dN[7], res.dN[2])
    assert np.array_equal(res.selected_coeff_indices[7], res.selected_coeff_indices[2])
    fG = g0.astype(np.float64) @ sc.B_sc
    w, dN, ns, sels = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
    assert np.array_equal(res.n_selected, ns)
    for a, b in zip(res.selected_coeff_indices, sels):
        assert np.array_equal(a, b)
    assert rel_err(res.spectral_weight, w) <= REL_TOL
    assert rel_err(res.dN, dN) <= REL_TOL


def test_all_108_cosets_in_one_submit_two_passes(unfolder):
    """More than BANDUP_GPU_MAX_FOLDS distinct cosets (here every one of the det M = 108, plus 108
    repeats) are worked off in passes inside the same launches: rows, counts, selected lists and
    the projected variant behave as with one pass, and the weights of every band sum to one."""
    from bandup_b200.unfold import MAX_FOLDS, PwWavefunction
    sc = synth.make_scenario("si333_vacancy", scale=0.25)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    ener[2] = ener[1]
    ener = np.sort(ener)
    cube = np.array([[i, j, k] for i in range(6) for j in range(6) for k in range(6)], dtype=np.int32)
    lab = synth.coset_labels(cube, sc.M)
    n_distinct = len(np.unique(lab))
    assert n_distinct == 108 > MAX_FOLDS
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=cube, ikpt=4)
    fG = cube.astype(np.float64) @ sc.B_sc
    w, dN, ns, sels = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
    assert np.array_equal(res.n_selected, ns) and ns.min() > 0
    for a, b in zip(res.selected_coeff_indices, sels):
        assert np.array_equal(a, b)
    assert rel_err(res.spectral_weight, w) <= REL_TOL
    assert rel_err(res.dN, dN) <= REL_TOL
    _, first = np.unique(lab, return_index=True)
    assert np.allclose(res.spectral_weight[first].sum(axis=0), 1.0, atol=1e-7)     # all cosets: sum_k P = 1
    for f in range(cube.shape[0]):                                                  # repeats share rows
        assert np.array_equal(res.spectral_weight[f], res.spectral_weight[first[np.searchsorted(np.unique(lab), lab[f])]])
    # the unfolding-density operator of a fold in the second pass
    from oracle import oracle as O
    from oracle import oracle_np as ON
    rho = unfolder.calc_rho(4)
    f = int(np.sort(first)[100])            # the 101st distinct coset in submit order: second pass
    c_ren, _ = O.renormalize(coeffs, 1)
    ies = np.nonzero(dN[f] >= 1e-3)[0]
    assert len(ies) > 0
    for ie in ies[:4]:
        want = ON.calc_rho(c_ren, sels[f], ener, dN[f][ie], grid[ie], grid[1] - grid[0], sigma=ON.EPS_DP)
        got = rho.dense(f, int(ie) + 1, coeffs.shape[0]).astype(np.complex128)
        big = np.abs(want) >= 1.2e-5
        assert np.all(np.abs(got[big] - want[big]) <= 2e-6 * max(np.abs(want).max(), 1e-30) + 2e-7 * np.abs(want[big]))


def test_device_entry_point_with_more_than_64_cosets(unfolder):
    """bandup_gpu_unfold_device: two passes work; the one-byte slot table output is refused."""
    import torch
    from bandup_b200.unfold import BandupGpuError, LAYOUT_FORTRAN_INMEM
    sc = synth.make_scenario("si333_vacancy", scale=0.2)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 2)
    nb, n_pw, nsp = coeffs.shape
    cube = np.array([[i, j, k] for i in range(5) for j in range(4) for k in range(4)], dtype=np.int32)
    assert len(np.unique(synth.coset_labels(cube, sc.M))) > 64
    dev = torch.device("cuda:0")
    nf = cube.shape[0]
    d_c = torch.from_numpy(coeffs.view(np.float32).reshape(nb, n_pw * nsp, 2)).to(dev)
    d_g, d_e = torch.from_numpy(gf).to(dev), torch.from_numpy(ener).to(dev)
    d_w = torch.zeros((nf, nb), dtype=torch.float64, device=dev)
    d_dn = torch.zeros((nf, len(grid)), dtype=torch.float64, device=dev)
    d_ns = torch.zeros(nf, dtype=torch.int32, device=dev)
    d_slot = torch.zeros(n_pw, dtype=torch.uint8, device=dev)
    ws = unfolder.workspace_bytes(nsp, n_pw, nb, nf)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    args = (stream, nsp, n_pw, nb, LAYOUT_FORTRAN_INMEM, n_pw * nsp * 8, d_c.data_ptr(), d_g.data_ptr(),
            d_e.data_ptr(), cube, d_w.data_ptr(), d_dn.data_ptr(), d_ns.data_ptr())
    with pytest.raises(BandupGpuError) as ei:
        unfolder.unfold_device(*args, d_slot_out=d_slot.data_ptr(), d_workspace=d_ws.data_ptr(), workspace_bytes=ws)
    assert ei.value.code == 5
    unfolder.unfold_device(*args, d_workspace=d_ws.data_ptr(), workspace_bytes=ws)
    torch.cuda.synchronize()
    w, dN, ns, _ = oracle_kpt(sc, coeffs, gf, ener, cube.astype(np.float64) @ sc.B_sc, grid)
    assert np.array_equal(d_ns.cpu().numpy(), ns)
    assert rel_err(d_w.cpu().numpy(), w) <= REL_TOL and rel_err(d_dn.cpu().numpy(), dN) <= REL_TOL


@pytest.mark.parametrize("n_spinor", [1, 2])
@pytest.mark.parametrize("n_pw_trim", [0, 1])
def test_vasp_record_layout_and_odd_rows(unfolder, n_spinor, n_pw_trim):
    """Raw WAVECAR band records ([band][spinor][pw], stride nrecl) give the same numbers as the
    Fortran in-memory layout; odd n_pw makes every other row 8- but not 16-byte aligned."""
    from bandup_b200.unfold import LAYOUT_VASP_RECORD, PwWavefunction
    sc = synth.make_scenario("mos2_8x8_spinor" if n_spinor == 2 else "graphene4x4", scale=0.07)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 1)
    if (gf.shape[0] - n_pw_trim) % 2 == 0:
        n_pw_trim += 1                                      # force an odd plane-wave count
    if n_pw_trim:
        coeffs, gf = np.ascontiguousarray(coeffs[:, :-n_pw_trim]), np.ascontiguousarray(gf[:-n_pw_trim])
    nb, n_pw, _ = coeffs.shape
    assert n_pw % 2 == 1
    fG = sc.folding_G(1)
    ref = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG)
    w, dN, ns, sels = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
    assert rel_err(ref.spectral_weight, w) <= REL_TOL and rel_err(ref.dN, dN) <= REL_TOL
    for nrecl in (n_pw * n_spinor * 8, n_pw * n_spinor * 8 + 8 * 5):
        rec = np.zeros((nb, nrecl // 8), dtype=np.complex64)
        for isp in range(n_spinor):                         # spinor-up block then spinor-down block
            rec[:, isp * n_pw:(isp + 1) * n_pw] = coeffs[:, :, isp]
        wf = PwWavefunction(rec, gf, ener, n_pw=n_pw, n_bands=nb, n_spinor=n_spinor,
                            layout=LAYOUT_VASP_RECORD, band_stride_bytes=nrecl)
        res = unfolder.perform_unfolding(wf, folding_G=fG)
        assert np.array_equal(res.n_selected, ns)
        assert rel_err(res.spectral_weight, w) <= REL_TOL
        assert rel_err(res.dN, dN) <= REL_TOL
        assert np.allclose(res.spectral_weight, ref.spectral_weight, rtol=1e-7, atol=1e-15)  # fp32 8-term runs group differently per alignment


def test_gaussian_smearing(unfolder):
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("graphene4x4", scale=0.1)
    grid = setup(unfolder, sc, smearing=0.04)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    fG = sc.folding_G(0)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG)
    w, dN, ns, _ = oracle_kpt(sc, coeffs, gf, ener, fG, grid, smearing=0.04)
    assert rel_err(res.spectral_weight, w) <= REL_TOL
    assert rel_err(res.dN, dN) <= REL_TOL


def test_box_lambda_bin_edges(unfolder):
    """Energies exactly on bin edges take lambda = 1/2 in both neighbouring bins."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("graphene4x4", scale=0.08)
    grid = setup(unfolder, sc)
    coeffs, gf, _ = synth.make_wavefunction(sc, 0)
    nb = coeffs.shape[0]
    de = grid[1] - grid[0]
    ener = np.array([grid[3 + 2 * i] + 0.5 * de for i in range(nb)])
    ener[0] = grid[0] - 0.5 * de          # lower edge of the first bin
    ener[1] = grid[-1] + 0.5 * de         # upper edge of the last bin
    ener[2] = grid[10]                    # bin centre
    fG = sc.folding_G(0)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG)
    w, dN, _, _ = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
    assert rel_err(res.dN, dN) <= REL_TOL
    assert np.count_nonzero(res.dN[0]) == np.count_nonzero(dN[0])


def test_empty_selection_is_reported(unfolder):
    """A coset with no plane wave in the cutoff sphere: n_selected == 0, the reference's
    'pckpt cannot be parsed' branch (main_BandUP.f90:161-172); weights and delta_N are zero."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("si333_vacancy", scale=0.2)
    setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    keep = synth.coset_labels(gf, sc.M) != synth.coset_labels(gf, sc.M)[0]   # drop the G=0 coset
    coeffs, gf = np.ascontiguousarray(coeffs[:, keep]), np.ascontiguousarray(gf[keep])
    g0 = np.array([[0, 0, 0], [1, 0, 0]], dtype=np.int32)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=g0)
    assert res.n_selected[0] == 0 and res.n_selected[1] > 0
    assert res.pckpt_cannot_be_parsed().tolist() == [True, False]
    assert not res.spectral_weight[0].any() and not res.dN[0].any()
    assert len(res.selected_coeff_indices[0]) == 0


def test_zero_norm_band_is_left_unscaled(unfolder):
    """io_routines_mod.f90:1325: a band with s <= epsilon is not rescaled."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("graphene4x4", scale=0.08)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    coeffs[1] = 0.0
    coeffs[2] *= np.float32(1e-10)        # s ~ 1e-20 * O(n_pw) < epsilon
    fG = sc.folding_G(0)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG)
    w, dN, _, _ = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
    assert not res.spectral_weight[:, 1].any()
    assert res.band_norms[2] <= 2.3e-16
    assert rel_err(res.spectral_weight, w) <= REL_TOL
    assert rel_err(res.dN, dN) <= REL_TOL


def test_pipeline_double_buffering(unfolder):
    """Two SC-kpts in flight, fetched out of order; a third submit is refused until a fetch."""
    from bandup_b200.unfold import BandupGpuError, PwWavefunction
    sc = synth.make_scenario("graphene4x4", scale=0.1)
    grid = setup(unfolder, sc)
    data = [synth.make_wavefunction(sc, iK) for iK in range(3)]
    wfs = [PwWavefunction(*d) for d in data]
    unfolder.submit(10, wfs[0], folding_G=sc.folding_G(0))
    unfolder.submit(11, wfs[1], folding_G=sc.folding_G(1))
    with pytest.raises(BandupGpuError) as ei:
        unfolder.submit(12, wfs[2], folding_G=sc.folding_G(2))
    assert ei.value.code == 1
    unfolder._inflight.pop(12, None)
    r1 = unfolder.fetch(11)
    unfolder.submit(12, wfs[2], folding_G=sc.folding_G(2))
    r0 = unfolder.fetch(10)
    r2 = unfolder.fetch(12)
    for iK, r in enumerate([r0, r1, r2]):
        w, dN, ns, _ = oracle_kpt(sc, *data[iK], sc.folding_G(iK), grid)
        assert np.array_equal(r.n_selected, ns)
        assert rel_err(r.spectral_weight, w) <= REL_TOL and rel_err(r.dN, dN) <= REL_TOL
    with pytest.raises(BandupGpuError) as ei:
        unfolder.fetch_raw(99, 1, 1)
    assert ei.value.code == 8


def test_error_codes(unfolder):
    from bandup_b200.unfold import BandupGpuError, MAX_FOLDS, PwWavefunction
    sc = synth.make_scenario("graphene4x4", scale=0.08)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    wf = PwWavefunction(coeffs, gf, ener)
    with pytest.raises(BandupGpuError) as ei:               # set_cells / set_grid not called
        unfolder.submit(1, wf, g0=np.zeros((1, 3), np.int32))
    assert ei.value.code == 7
    unfolder._inflight.clear()
    a = synth.hex_cell(2.467, 15.0)
    with pytest.raises(BandupGpuError) as ei:               # verify_commens' stop
        unfolder.verify_commens(synth.rec_latt(a), synth.rec_latt(np.diag([4.0, 4.1, 1.0]) @ a))
    assert ei.value.code == 3
    setup(unfolder, sc)
    with pytest.raises(BandupGpuError) as ei:               # folding vector off the SC lattice
        unfolder.folding_g0(sc.folding_G(0) + np.array([3e-3, 0.0, 0.0]))
    assert ei.value.code == 4
    g0, res = unfolder.folding_g0(sc.folding_G(0) + np.array([4e-5, 0.0, 0.0]))
    assert res < 1e-3 and np.array_equal(g0, unfolder.folding_g0(sc.folding_G(0))[0])
    many = np.zeros((MAX_FOLDS + 1, 3), np.int32)
    many[:, 0] = np.arange(MAX_FOLDS + 1)                   # 4x4 cell: only 4 distinct cosets along B1
    unfolder.submit(1, wf, g0=many)                         # > MAX_FOLDS pc-kpts is fine ...
    r = unfolder.fetch(1)
    assert np.array_equal(r.spectral_weight[0], r.spectral_weight[4]) and r.spectral_weight.shape[0] == MAX_FOLDS + 1
    with pytest.raises(BandupGpuError) as ei:               # ... too many pc-kpts per submit is not
        unfolder.submit(1, wf, g0=np.zeros((5000, 3), np.int32))
    assert ei.value.code == 5
    unfolder._inflight.clear()
    lib = unfolder._lib
    assert lib.bandup_gpu_finalize(12345) == 6
    h = C.c_int(-1)
    assert lib.bandup_gpu_init(999, C.byref(h)) == 9 and h.value == -1


def test_device_resident_full_size_sige555(unfolder):
    """BASELINE config 5 at full size (3000 bands x ~100k plane waves, 2.4 GB): coefficients
    are generated in HBM by the counter-based hash, all bands go through K1-K3, and sampled
    bands are compared with the oracle fed with the same hash on the host.  Size-independent
    properties: sum over folds of P <= 1, delta_N sums to the in-window weight."""
    import torch
    from bandup_b200.unfold import LAYOUT_FORTRAN_INMEM
    from oracle import oracle as O
    sc = synth.make_scenario("sige555")
    grid = setup(unfolder, sc)
    iK = 3
    gf = sc.gvecs(iK)
    n_pw, nb = gf.shape[0], sc.n_bands
    assert 90_000 < n_pw < 115_000 and nb == 3000
    fG = sc.folding_G(iK)
    g0, _ = unfolder.folding_g0(fG)
    g0 = np.concatenate([g0, g0 + np.array([[1, 0, 0]], np.int32), g0 + np.array([[0, 2, 1]], np.int32)])
    nf = g0.shape[0]
    rng = np.random.default_rng(11)
    ener = np.sort(rng.uniform(-21.0, 6.0, nb))
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    d_c = torch.empty((nb, n_pw, 2), dtype=torch.float32, device=dev)
    unfolder.synth_fill_device(stream, d_c.data_ptr(), n_pw, nb, n_pw * 8, 20261024, iK)
    d_g = torch.from_numpy(gf).to(dev)
    d_e = torch.from_numpy(ener).to(dev)
    d_w = torch.empty((nf, nb), dtype=torch.float64, device=dev)
    d_dn = torch.empty((nf, len(grid)), dtype=torch.float64, device=dev)
    d_ns = torch.empty(nf, dtype=torch.int32, device=dev)
    d_norm = torch.empty(nb, dtype=torch.float64, device=dev)
    ws = unfolder.workspace_bytes(1, n_pw, nb, nf)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    unfolder.unfold_device(stream, 1, n_pw, nb, LAYOUT_FORTRAN_INMEM, n_pw * 8, d_c.data_ptr(),
                           d_g.data_ptr(), d_e.data_ptr(), g0, d_w.data_ptr(), d_dn.data_ptr(),
                           d_ns.data_ptr(), d_ws.data_ptr(), ws, d_norms=d_norm.data_ptr())
    torch.cuda.synchronize()
    w, dN, ns, norms = d_w.cpu().numpy(), d_dn.cpu().numpy(), d_ns.cpu().numpy(), d_norm.cpu().numpy()
    gc = gf @ sc.B_sc
    fG_all = g0.astype(np.float64) @ sc.B_sc
    sample = [0, 1, 777, 1500, 2998, 2999]
    for b in sample:
        row = O.synth_fill(n_pw, 1, 20261024, iK, band0=b).reshape(1, n_pw, 1)
        # generator check: the device block equals the host hash bit for bit
        host = torch.view_as_complex(d_c[b]).cpu().numpy()
        assert np.array_equal(host, row[0, :, 0])
        wo, _, nso, _ = O.unfold_kpt(row, gc, ener[b:b + 1], sc.b_pc, fG_all, grid, norm_mode=1)
        assert np.array_equal(ns, nso)
        assert rel_err(w[:, b], wo[:, 0]) <= REL_TOL
        _, n_o = O.renormalize(row, 1)
        assert rel_err(norms[b:b + 1], n_o) <= REL_TOL
    assert np.all(w.sum(axis=0) <= 1.0 + 1e-9) and np.all(w >= 0.0)
    de = grid[1] - grid[0]
    inside = (ener > grid[0] - 0.5 * de) & (ener < grid[-1] + 0.5 * de)
    assert np.allclose(dN.sum(axis=1), w[:, inside].sum(axis=1), rtol=1e-12)
    dN_o = O.delta_N(grid, ener, w[0])
    assert rel_err(dN[0], dN_o) <= REL_TOL


def test_sharded_job_whole_scenario(unfolder):
    """Every SC-kpt of a (shrunken) graphene job through ShardedUnfoldJob with the real GPU
    unfolder: double-buffered submit/fetch, rows assembled in pc-kpt order, against the oracle."""
    from bandup_b200.sharding import ShardedUnfoldJob
    from bandup_b200.unfold import PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario("graphene4x4", scale=0.08, n_kpts=15)
    grid = setup(unfolder, sc)
    data = {}

    def load(iK):
        coeffs, gf, ener = synth.make_wavefunction(sc, iK)
        data[iK] = (coeffs, gf, ener)
        return PwWavefunction(coeffs, gf, ener), sc.folding_G(iK), np.asarray(sc.folds[iK])

    costs = [sc.n_bands * sc.gvecs(iK).shape[0] for iK in range(sc.n_sc_kpts)]
    got = ShardedUnfoldJob(unfolder, sc.n_sc_kpts, costs, load).run()
    n_rows = sum(len(f) for f in sc.folds)
    assert got.row_ids.tolist() == list(range(n_rows))
    for iK in range(sc.n_sc_kpts):
        coeffs, gf, ener = data[iK]
        _, dN, _, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid, norm_mode=1)
        assert rel_err(got.dN[np.asarray(sc.folds[iK])], dN) <= REL_TOL


def test_device_batch_equals_per_kpt_calls(unfolder):
    """bandup_gpu_unfold_device_batch (one launch of each kernel for several SC-kpts with different
    n_pw / n_fold, one of them with two pc-kpts in the same coset) gives the same results as
    per-k-point bandup_gpu_unfold_device calls, and both agree with the oracle."""
    import torch
    from bandup_b200.unfold import LAYOUT_FORTRAN_INMEM
    sc = synth.make_scenario("si333_vacancy", scale=0.25)
    grid = setup(unfolder, sc)
    nener = len(grid)
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    kpts, host = [], []
    for iK in (0, 3, sc.n_sc_kpts - 1):
        coeffs, gf, ener = synth.make_wavefunction(sc, iK)
        fG = sc.folding_G(iK)
        if iK == 3:
            fG = np.concatenate([fG, fG[:1] + 2 * sc.b_pc[1]])      # same coset as its first fold
        g0, _ = unfolder.folding_g0(fG)
        nb, n_pw, _ = coeffs.shape
        nf = g0.shape[0]
        t = dict(c=torch.from_numpy(coeffs.view(np.float32).reshape(nb, n_pw, 2)).to(dev),
                 g=torch.from_numpy(gf).to(dev), e=torch.from_numpy(ener).to(dev))
        outs = [dict(w=torch.zeros((nf, nb), dtype=torch.float64, device=dev),
                     dn=torch.zeros((nf, nener), dtype=torch.float64, device=dev),
                     ns=torch.zeros(nf, dtype=torch.int32, device=dev),
                     nr=torch.zeros(nb, dtype=torch.float64, device=dev)) for _ in range(2)]
        host.append((coeffs, gf, ener, fG, t, outs, g0))
        kpts.append(dict(n_spinor=1, n_pw=n_pw, n_bands=nb, layout=LAYOUT_FORTRAN_INMEM, band_stride_bytes=n_pw * 8,
                         d_coeffs=t["c"].data_ptr(), d_g_frac=t["g"].data_ptr(), d_band_energies=t["e"].data_ptr(),
                         g0=g0, d_weights=outs[0]["w"].data_ptr(), d_dN=outs[0]["dn"].data_ptr(),
                         d_n_selected=outs[0]["ns"].data_ptr(), d_norms=outs[0]["nr"].data_ptr()))
    descs, keep = unfolder.make_batch(kpts)
    ws = unfolder.batch_workspace_bytes(descs)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    l0 = unfolder.launch_count()
    unfolder.unfold_device_batch(stream, descs, d_ws.data_ptr(), ws)
    assert unfolder.launch_count() - l0 == 4                        # K1, K2, K3 once each + one row-replication launch
    for coeffs, gf, ener, fG, t, outs, g0 in host:
        nb, n_pw, _ = coeffs.shape
        w1 = unfolder.workspace_bytes(1, n_pw, nb, g0.shape[0])
        d_w1 = torch.empty(w1, dtype=torch.uint8, device=dev)
        o = outs[1]
        unfolder.unfold_device(stream, 1, n_pw, nb, LAYOUT_FORTRAN_INMEM, n_pw * 8, t["c"].data_ptr(),
                               t["g"].data_ptr(), t["e"].data_ptr(), g0, o["w"].data_ptr(), o["dn"].data_ptr(),
                               o["ns"].data_ptr(), d_w1.data_ptr(), w1, d_norms=o["nr"].data_ptr())
        torch.cuda.synchronize()
        # same numbers; the tile size may differ between a batch and a single k-point, so the fp64
        # partial sums may be grouped differently (last-bit differences only)
        assert torch.equal(outs[0]["ns"], outs[1]["ns"])
        for key in ("w", "dn", "nr"):
            assert torch.allclose(outs[0][key], outs[1][key], rtol=1e-12, atol=1e-300), key
        w, dN, ns, _ = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
        assert np.array_equal(outs[0]["ns"].cpu().numpy(), ns)
        assert rel_err(outs[0]["w"].cpu().numpy(), w) <= REL_TOL
        assert rel_err(outs[0]["dn"].cpu().numpy(), dN) <= REL_TOL


@pytest.mark.parametrize("name,scale", [("graphene4x4", 0.1), ("mos2_8x8_spinor", 0.05)])
def test_unfold_from_wavecar_file(unfolder, tmp_path, name, scale):
    """File-level path: a synthetic WAVECAR in the reader's own record format, raw band records
    uploaded untouched (VASP_RECORD layout, stride nrecl), against the oracle fed with the
    de-interleaved coefficients."""
    from bandup_b200.wavecar import Wavecar, unfold_wavecar, write_wavecar
    from oracle import oracle as O
    sc = synth.make_scenario(name, scale=scale, n_kpts=9)
    data = [synth.make_wavefunction(sc, iK) for iK in range(sc.n_sc_kpts)]
    path = tmp_path / "WAVECAR"
    write_wavecar(path, sc.A_sc, sc.ecut, list(sc.sc_kpts_frac), [d[0] for d in data], [d[2] for d in data],
                  extra_pad=1)
    wc = Wavecar(path)
    got, grid = unfold_wavecar(unfolder, wc, sc.b_pc, sc.pc_kpts_cart, *sc.energy_window())
    host_g, _ = unfold_wavecar(unfolder, wc, sc.b_pc, sc.pc_kpts_cart, *sc.energy_window(), device_gvecs=False)
    assert np.array_equal(got.dN, host_g.dN)            # device-made and host-made plane-wave lists agree
    n_rows = sum(len(f) for f in sc.folds)
    assert got.row_ids.tolist() == list(range(n_rows))
    for iK in range(sc.n_sc_kpts):
        coeffs, gf, ener = data[iK]
        _, dN, _, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid, norm_mode=1)
        assert rel_err(got.dN[np.asarray(sc.folds[iK])], dN) <= REL_TOL


def test_spectral_function_and_dont_unfold(unfolder):
    """calc_spectral_function (Gaussian A(k,E)) on a resident SC-kpt, and -dont_unfold (P = 1)."""
    from bandup_b200.unfold import PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario("graphene4x4", scale=0.1)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    fG = sc.folding_G(0)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG, ikpt=4)
    for sd in (None, 0.08):
        sf = unfolder.calc_spectral_function(4, fG.shape[0], sd)
        for f in range(fG.shape[0]):
            want = O.spectral_function(grid, ener, res.spectral_weight[f], sd or 0.0)
            assert rel_err(sf[f], want) <= REL_TOL
    unfolder.set_perform_unfold(False)
    res1 = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG, ikpt=5)
    unfolder.set_perform_unfold(True)
    assert np.all(res1.spectral_weight == 1.0)
    assert rel_err(res1.dN[0], O.delta_N(grid, ener, np.ones(len(ener)))) <= REL_TOL
    assert np.array_equal(res1.n_selected, res.n_selected)


def test_deviation_from_faithful_fp32_norm_is_bounded(unfolder):
    """The reference's own renormalisation accumulates |C|^2 in single precision (a complex(sp)
    dot_product, io_routines_mod.f90:1322-1324), which by itself is 1e-6..5e-6 off.
    Against that 'faithful_fp32_sequential' restatement the GPU weights agree to the reference's
    own rounding noise, not to 1e-6: asserted loosely (5e-5) and tight (1e-6) against exact_fp64."""
    from bandup_b200.unfold import PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario("si333_vacancy", scale=0.3)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    fG = sc.folding_G(0)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG)
    gc = gf @ sc.B_sc
    w32, _, _, _ = O.unfold_kpt(coeffs, gc, ener, sc.b_pc, fG, grid, norm_mode=0)
    w64, _, _, _ = O.unfold_kpt(coeffs, gc, ener, sc.b_pc, fG, grid, norm_mode=1)
    assert rel_err(res.spectral_weight, w64) <= REL_TOL
    dev = rel_err(res.spectral_weight, w32)
    assert dev <= 5e-5
    print(f"max rel deviation from the fp32-norm restatement: {dev:.2e}; fp32 vs fp64 oracle: "
          f"{rel_err(w32, w64):.2e}")


def test_unfold_task_file_level(unfolder, tmp_path):
    """`bandup unfold` at file level (-no_symm_sckpts -no_symm_avg): prim_cell_lattice.in,
    supercell_lattice.in, KPOINTS_prim_cell.in, energy_info.in and a WAVECAR go in,
    unfolded_EBS_not-symmetry_averaged.dat comes out in the format `bandup plot` loads; the numbers
    are the oracle's (in memory to 1e-6; in the file to the 4 digits ES10.3 keeps)."""
    from bandup_b200 import bandup_files as BF
    from bandup_b200.wavecar import write_wavecar
    from oracle import oracle as O
    a = synth.hex_cell(2.467, 15.0)
    b = synth.rec_latt(a)
    segs = [([1 / 3, 1 / 3, 0.0], [0.0, 0.0, 0.0]), ([0.0, 0.0, 0.0], [0.5, 0.0, 0.0])]
    counts = [5, 4]
    kcart = np.concatenate([BF.kpts_line(np.array(s) @ b, np.array(e) @ b, n) for (s, e), n in zip(segs, counts)])
    sc = synth.Scenario("graphene4x4", a, np.diag([4, 4, 1]), 400.0 * 0.08, 6, 1, kcart, e_fermi=-1.25,
                        emin=-12.0, emax=6.0, dE=0.1)
    data = [synth.make_wavefunction(sc, iK) for iK in range(sc.n_sc_kpts)]
    BF.write_lattice(tmp_path / "prim_cell_lattice.in", a, "graphene pc")
    BF.write_lattice(tmp_path / "supercell_lattice.in", sc.A_sc, "graphene 4x4")
    BF.write_pckpts(tmp_path / "KPOINTS_prim_cell.in", segs, counts)
    BF.write_energy_info(tmp_path / "energy_info.in", sc.e_fermi, sc.emin, sc.emax, sc.dE)
    write_wavecar(tmp_path / "WAVECAR", sc.A_sc, sc.ecut, list(sc.sc_kpts_frac), [d[0] for d in data],
                  [d[2] for d in data])
    out = tmp_path / "unfolded_EBS_not-symmetry_averaged.dat"
    udo = tmp_path / "unfolding_density_operator_not-symm_avgd.dat"
    kpath, grid, dN = BF.unfold_task(unfolder, tmp_path / "WAVECAR", tmp_path / "prim_cell_lattice.in",
                                     tmp_path / "supercell_lattice.in", tmp_path / "KPOINTS_prim_cell.in",
                                     tmp_path / "energy_info.in", out_file=str(out), unf_dens_op_file=str(udo))
    assert np.array_equal(grid, O.real_seq(*sc.energy_window()))
    want = np.zeros_like(dN)
    for iK in range(sc.n_sc_kpts):
        coeffs, gf, ener = data[iK]
        _, d, _, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid, norm_mode=1)
        want[np.asarray(sc.folds[iK])] = d
    assert rel_err(dN, want) <= REL_TOL
    tab = BF.read_band_struc(out)
    assert tab.shape == (sum(counts) * len(grid), 3)
    assert np.allclose(tab[:, 2].reshape(sum(counts), len(grid)), want, rtol=6e-4, atol=1e-12)
    assert np.allclose(tab[:len(grid), 1], grid - sc.e_fermi, atol=5e-5)
    assert tab[-1, 0] > tab[0, 0] and np.all(np.diff(tab[::len(grid), 0]) >= -1e-12)
    # the rho file (-write_unf_dens_op): every block against the oracle's calc_rho, to the file's 4 digits
    from oracle import oracle_np as ON
    hdr, recs = BF.read_unf_dens_op(udo)
    assert hdr["nbands"] == sc.n_bands and hdr["nener"] == len(grid) and len(recs) > 5
    de = grid[1] - grid[0]
    for r in recs:
        row = r.pckpt_number - 1
        iK = next(i for i, f in enumerate(sc.folds) if row in f)
        assert r.folding_sckpt_number == iK + 1
        coeffs, gf, ener = data[iK]
        c_ren, _ = O.renormalize(coeffs, 1)
        sel, _ = O.select_coeffs(gf @ sc.B_sc, sc.pc_kpts_cart[row] - sc.sc_kpts_cart[iK], sc.b_pc)
        rho = ON.calc_rho(c_ren, sel, ener, want[row, r.iener - 1], grid[r.iener - 1], de)
        assert abs(r.unfolded_N - want[row, r.iener - 1]) <= 6e-4 * want[row, r.iener - 1]
        for i, j, v in zip(r.rows, r.cols, r.entries):
            assert i <= j and abs(v - (rho[i, j] if i != j else rho[i, j].real)) <= 6e-4 * abs(rho[i, j]) + 1e-7
        assert np.count_nonzero(np.abs(np.triu(rho)) >= 1.2e-4) <= len(r.entries) <= np.count_nonzero(np.abs(np.triu(rho)) >= 0.8e-4)


@pytest.mark.parametrize("name,scale", [("graphene4x4", 0.1), ("mos2_8x8_spinor", 0.05), ("si333_vacancy", 0.2)])
def test_wavecar_records_with_device_gvecs(unfolder, tmp_path, name, scale):
    """bandup_gpu_submit_kpt_vasp: header values + raw band records in; the plane-wave list, the
    scalar/spinor decision and the count check happen on the device.  The list must equal the
    reader's enumeration (oracle bo_gen_gvecs) element for element."""
    from bandup_b200.unfold import BandupGpuError
    from bandup_b200.wavecar import Wavecar, write_wavecar
    from oracle import oracle as O
    sc = synth.make_scenario(name, scale=scale, n_kpts=9)
    grid = setup(unfolder, sc)
    ks = [0, sc.n_sc_kpts - 1]
    data = [synth.make_wavefunction(sc, iK) for iK in ks]
    path = tmp_path / "WAVECAR"
    write_wavecar(path, sc.A_sc, sc.ecut, [sc.sc_kpts_frac[iK] for iK in ks], [d[0] for d in data],
                  [d[2] for d in data], extra_pad=2)
    wc = Wavecar(path)
    for i, iK in enumerate(ks):
        coeffs, gf, ener = data[i]
        nsp, npw = wc.submit_to(unfolder, 40 + i, i + 1, sc.folding_G(iK))
        assert (nsp, npw) == (sc.n_spinor, gf.shape[0])
        g_dev = unfolder.fetch_gvecs(40 + i, npw)
        g_ref, _ = O.gen_gvecs(sc.sc_kpts_frac[iK], sc.ecut, sc.B_sc)
        assert np.array_equal(g_dev, g_ref) and np.array_equal(g_dev, gf)
        res = unfolder.fetch(40 + i, want_selected=True)
        w, dN, ns, sels = oracle_kpt(sc, coeffs, gf, ener, sc.folding_G(iK), grid)
        assert np.array_equal(res.n_selected, ns)
        for a, b in zip(res.selected_coeff_indices, sels):
            assert np.array_equal(a, b)
        assert rel_err(res.spectral_weight, w) <= REL_TOL and rel_err(res.dN, dN) <= REL_TOL
    # a header whose plane-wave count no adjustment of 2m/hbar^2 can reproduce is the reader's error
    h = wc.kpt_header(1)
    raw = wc.read_band_records(1)
    with pytest.raises(BandupGpuError) as ei:
        unfolder.submit_vasp(50, raw, wc.nrecl, h.nplane - 7 * sc.n_spinor, wc.nbands, h.kpt_frac_coords, wc.encut,
                             h.band_energies, folding_G=sc.folding_G(ks[0]))
    assert ei.value.code == 1 and "plane waves" in str(ei.value)


def test_device_gvecs_full_size(unfolder):
    """~100k plane waves of a C5 k-point: device enumeration == the reader's, in order."""
    from oracle import oracle as O
    sc = synth.make_scenario("sige555")
    setup(unfolder, sc)
    iK = 5
    g_ref, _ = O.gen_gvecs(sc.sc_kpts_frac[iK], sc.ecut, sc.B_sc)
    nb = 3
    raw = np.zeros(nb * g_ref.shape[0] * 8, dtype=np.uint8)
    raw.view(np.complex64)[:] = 1.0
    nsp, npw = unfolder.submit_vasp(60, raw, g_ref.shape[0] * 8, g_ref.shape[0], nb, sc.sc_kpts_frac[iK], sc.ecut,
                                    np.array([-1.0, 0.0, 1.0]), folding_G=sc.folding_G(iK))
    assert (nsp, npw) == (1, g_ref.shape[0]) and npw > 90_000
    assert np.array_equal(unfolder.fetch_gvecs(60, npw), g_ref)
    res = unfolder.fetch(60)
    assert np.allclose(res.band_norms, npw) and np.all(res.n_selected > 0)


@pytest.mark.parametrize("n_pw,n_bands,n_spinor", [(1, 1, 1), (2, 3, 1), (3, 2, 2), (5, 1, 2), (1025, 2, 1)])
def test_tiny_and_ragged_blocks(unfolder, n_pw, n_bands, n_spinor):
    """Rows shorter than one 16-byte vector, rows of exactly one stage + 1, single bands."""
    from bandup_b200.unfold import PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario("graphene4x4", scale=0.3)
    grid = setup(unfolder, sc)
    gf = sc.gvecs(0)[:n_pw]
    rng = np.random.default_rng(n_pw * 7 + n_bands)
    coeffs = (rng.standard_normal((n_bands, n_pw, n_spinor)) + 1j * rng.standard_normal((n_bands, n_pw, n_spinor))
              ).astype(np.complex64)
    ener = np.sort(rng.uniform(-5, 2, n_bands))
    g0 = np.array([[0, 0, 0], [1, 0, 0]], dtype=np.int32)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=g0)
    fG = g0.astype(np.float64) @ sc.B_sc
    w, dN, ns, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, fG, grid, norm_mode=1)
    assert np.array_equal(res.n_selected, ns)
    assert rel_err(res.spectral_weight, w) <= REL_TOL and rel_err(res.dN, dN) <= REL_TOL
    _, norms = O.renormalize(coeffs, 1)
    assert rel_err(res.band_norms, norms) <= REL_TOL


def test_sixty_four_distinct_cosets(unfolder):
    """BANDUP_GPU_MAX_FOLDS distinct cosets in one pass: the shared-memory bins take 130 KB and the
    stage ring shrinks to what is left."""
    from bandup_b200.unfold import MAX_FOLDS, PwWavefunction
    sc = synth.make_scenario("si333_vacancy", scale=0.25)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 1)
    cube = np.array([[i, j, k] for i in range(5) for j in range(5) for k in range(5)], dtype=np.int32)
    lab = synth.coset_labels(cube, sc.M)
    _, first = np.unique(lab, return_index=True)
    g0 = cube[np.sort(first)[:MAX_FOLDS]]
    assert g0.shape[0] == MAX_FOLDS
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), g0=g0, want_selected=False)
    w, dN, ns, _ = oracle_kpt(sc, coeffs, gf, ener, g0.astype(np.float64) @ sc.B_sc, grid)
    assert np.array_equal(res.n_selected, ns)
    assert rel_err(res.spectral_weight, w) <= REL_TOL and rel_err(res.dN, dN) <= REL_TOL
    assert np.all(res.spectral_weight.sum(axis=0) <= 1.0 + 1e-7)


def test_random_cells_and_shapes(unfolder):
    """Randomised sweep: triclinic cells, integer SC matrices with negative entries and negative
    determinant, random pc k-points, both layouts, padded strides, scalar and spinor -- selection
    bit-exact and weights / delta_N within tolerance for every case."""
    from bandup_b200.unfold import LAYOUT_FORTRAN_INMEM, LAYOUT_VASP_RECORD, PwWavefunction
    rng = np.random.default_rng(20261020)
    done = 0
    while done < 16:
        a_pc = np.eye(3) * rng.uniform(2.5, 4.0, 3) + rng.uniform(-0.6, 0.6, (3, 3))
        M = rng.integers(-3, 4, (3, 3))
        det = int(round(np.linalg.det(M)))
        if abs(det) < 2 or abs(det) > 40 or abs(np.linalg.det(a_pc)) < 5.0:
            continue
        kf = rng.uniform(-0.5, 0.5, (4, 3))
        kf[0] = 0.0
        n_spinor = int(rng.integers(1, 3))
        sc = synth.Scenario(f"rand{done}", a_pc, M, float(rng.uniform(18.0, 40.0)), int(rng.integers(1, 12)),
                            n_spinor, kf @ synth.rec_latt(a_pc), e_fermi=float(rng.uniform(-2, 2)), emin=-6.0,
                            emax=4.0, dE=float(rng.choice([0.05, 0.1, 0.37])))
        Mgot = unfolder.verify_commens(sc.b_pc, sc.B_sc)
        # get_rec_latt divides by |V| (crystals_mod.f90:186-202), so a left-handed SC (det M < 0)
        # gets the inverted reciprocal basis and the reference's own formula returns -M: same lattice
        assert np.array_equal(Mgot, M if det > 0 else -M)
        grid = unfolder.set_energy_grid(*sc.energy_window())
        iK = int(rng.integers(0, sc.n_sc_kpts))
        coeffs, gf, ener = synth.make_wavefunction(sc, iK, seed=int(rng.integers(1, 10**6)))
        if gf.shape[0] < 4:
            continue
        ener = rng.uniform(sc.e_fermi - 7, sc.e_fermi + 5, sc.n_bands)        # unsorted band energies are legal
        fG = sc.folding_G(iK)
        nb, n_pw, _ = coeffs.shape
        if rng.random() < 0.5:
            nrecl = n_pw * n_spinor * 8 + 8 * int(rng.integers(0, 4))
            rec = np.zeros((nb, nrecl // 8), dtype=np.complex64)
            for isp in range(n_spinor):
                rec[:, isp * n_pw:(isp + 1) * n_pw] = coeffs[:, :, isp]
            wf = PwWavefunction(rec, gf, ener, n_pw=n_pw, n_bands=nb, n_spinor=n_spinor, layout=LAYOUT_VASP_RECORD,
                                band_stride_bytes=nrecl)
        else:
            wf = PwWavefunction(coeffs, gf, ener)
        res = unfolder.perform_unfolding(wf, folding_G=fG, ikpt=done + 1)
        w, dN, ns, sels = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
        assert np.array_equal(res.n_selected, ns), (done, M.tolist())
        for a, b in zip(res.selected_coeff_indices, sels):
            assert np.array_equal(a, b), (done, M.tolist())
        assert rel_err(res.spectral_weight, w) <= REL_TOL, done
        assert rel_err(res.dN, dN) <= REL_TOL, done
        done += 1


def test_full_size_si333_host_path(unfolder):
    """BASELINE config 2 at full size (500 bands x ~20k plane waves, det M = 108) through the host
    C ABI, every band against the oracle."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("si333_vacancy")
    grid = setup(unfolder, sc)
    iK = 7
    coeffs, gf, ener = synth.make_wavefunction(sc, iK)
    assert coeffs.shape[0] == 500 and 18_000 < gf.shape[0] < 22_000
    fG = sc.folding_G(iK)
    res = unfolder.perform_unfolding(PwWavefunction(coeffs, gf, ener), folding_G=fG)
    w, dN, ns, sels = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
    assert np.array_equal(res.n_selected, ns)
    for a, b in zip(res.selected_coeff_indices, sels):
        assert np.array_equal(a, b)
    assert rel_err(res.spectral_weight, w) <= REL_TOL and rel_err(res.dN, dN) <= REL_TOL


def test_device_resident_full_size_mos2_spinor(unfolder):
    """BASELINE config 3 at full size (1200 bands x ~79k plane waves x 2 spinor components, 1.5 GB,
    [pw][spinor] interleaved): sampled bands against the oracle, invariants on all."""
    import torch
    from bandup_b200.unfold import LAYOUT_FORTRAN_INMEM
    from oracle import oracle as O
    sc = synth.make_scenario("mos2_8x8_spinor")
    grid = setup(unfolder, sc)
    iK = 2
    gf = sc.gvecs(iK)
    n_pw, nb = gf.shape[0], sc.n_bands
    assert nb == 1200 and 70_000 < n_pw < 90_000
    g0, _ = unfolder.folding_g0(sc.folding_G(iK))
    nf = g0.shape[0]
    ener = np.sort(np.random.default_rng(5).uniform(-21.0, 6.0, nb))
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    row = 2 * n_pw
    d_c = torch.empty((nb, row, 2), dtype=torch.float32, device=dev)
    unfolder.synth_fill_device(stream, d_c.data_ptr(), row, nb, row * 8, 777, iK)
    d_g, d_e = torch.from_numpy(gf).to(dev), torch.from_numpy(ener).to(dev)
    d_w = torch.empty((nf, nb), dtype=torch.float64, device=dev)
    d_dn = torch.empty((nf, len(grid)), dtype=torch.float64, device=dev)
    d_ns = torch.empty(nf, dtype=torch.int32, device=dev)
    d_norm = torch.empty(nb, dtype=torch.float64, device=dev)
    ws = unfolder.workspace_bytes(2, n_pw, nb, nf)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    unfolder.unfold_device(stream, 2, n_pw, nb, LAYOUT_FORTRAN_INMEM, row * 8, d_c.data_ptr(), d_g.data_ptr(),
                           d_e.data_ptr(), g0, d_w.data_ptr(), d_dn.data_ptr(), d_ns.data_ptr(), d_ws.data_ptr(), ws,
                           d_norms=d_norm.data_ptr())
    torch.cuda.synchronize()
    w, dN, ns, norms = d_w.cpu().numpy(), d_dn.cpu().numpy(), d_ns.cpu().numpy(), d_norm.cpu().numpy()
    gc = gf @ sc.B_sc
    fG = g0.astype(np.float64) @ sc.B_sc
    for b in (0, 599, 1199):
        rowdata = O.synth_fill(row, 1, 777, iK, band0=b).reshape(1, n_pw, 2)     # [pw][spinor] interleaved
        wo, _, nso, _ = O.unfold_kpt(rowdata, gc, ener[b:b + 1], sc.b_pc, fG, grid, norm_mode=1)
        assert np.array_equal(ns, nso)
        assert rel_err(w[:, b], wo[:, 0]) <= REL_TOL
        _, n_o = O.renormalize(rowdata, 1)
        assert rel_err(norms[b:b + 1], n_o) <= REL_TOL
    assert np.all(w >= 0.0) and np.all(w.sum(axis=0) <= 1.0 + 1e-7)
    assert rel_err(dN[0], O.delta_N(grid, ener, w[0])) <= REL_TOL


def test_full_job_config1_graphene(unfolder, tmp_path):
    """BASELINE config 1 as a whole job: graphene 4x4, ENCUT 400 eV, 64 bands, every SC-kpt of the
    K-Gamma-M-K path from a synthetic WAVECAR (the reader's record format), all rows vs the oracle."""
    from bandup_b200.wavecar import Wavecar, unfold_wavecar, write_wavecar
    from oracle import oracle as O
    sc = synth.make_scenario("graphene4x4")
    assert sc.n_bands == 64 and sc.ecut == 400.0
    data = [synth.make_wavefunction(sc, iK) for iK in range(sc.n_sc_kpts)]
    assert 22_000 < data[0][1].shape[0] < 24_000
    write_wavecar(tmp_path / "WAVECAR", sc.A_sc, sc.ecut, list(sc.sc_kpts_frac), [d[0] for d in data],
                  [d[2] for d in data])
    got, grid = unfold_wavecar(unfolder, Wavecar(tmp_path / "WAVECAR"), sc.b_pc, sc.pc_kpts_cart, *sc.energy_window())
    assert got.row_ids.tolist() == list(range(sc.pc_kpts_cart.shape[0]))
    for iK in range(sc.n_sc_kpts):
        coeffs, gf, ener = data[iK]
        _, dN, _, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid, norm_mode=1)
        assert rel_err(got.dN[np.asarray(sc.folds[iK])], dN) <= REL_TOL


def test_fold_map_matches_float_vec_in_latt(unfolder):
    """K7: the geometric unfolding relations (which SC-kpt each pc-kpt folds onto) against a brute
    force with the oracle's vec_in_latt, including symmetry-equivalent point lists, skipped
    SC-kpts and the tolerance escalation of get_geom_unfolding_relations."""
    from bandup_b200.unfold import BandupGpuError
    from oracle import oracle as O
    sc = synth.make_scenario("si333_vacancy", scale=0.2, n_kpts=40)
    setup(unfolder, sc)
    pk = sc.pc_kpts_cart.copy()
    K = sc.sc_kpts_cart
    # exact folding first: must reproduce the scenario's own map
    fs, fe, tol = unfolder.get_geom_unfolding_relations(pk, K)
    want = np.empty(pk.shape[0], dtype=int)
    for iK, rows in enumerate(sc.folds):
        want[rows] = iK
    assert np.array_equal(fs, want) and np.array_equal(fe, want) and abs(tol - 1e-5) < 1e-18

    def brute(pk, eqv, skip, tol):
        out = np.full(pk.shape[0], -1)
        for i, k in enumerate(pk):
            for s in range(skip, len(eqv)):
                if any(O.vec_in_latt(k - e, sc.B_sc, tol) for e in eqv[s]):
                    out[i] = s
                    break
        return out
    # equivalent-point lists (each SC-kpt plus two images), the first two SC-kpts skipped, and
    # pc-kpts nudged off the lattice by up to 3e-5: the tolerance has to grow
    rng = np.random.default_rng(9)
    eqv = [np.stack([K[s], -K[s], K[s] + 0.37 * sc.B_sc[0]]) for s in range(K.shape[0])]
    pk2 = pk + rng.uniform(-1, 1, pk.shape) * 3e-5 / np.sqrt(3)
    skip = 2
    keep = np.array([want[i] >= skip or np.any([O.vec_in_latt(pk[i] + K[s], sc.B_sc, 1e-5) for s in range(skip, K.shape[0])])
                     for i in range(pk.shape[0])])
    pk2 = pk2[keep]
    fs2, fe2, tol2 = unfolder.get_geom_unfolding_relations(pk2, K, eqv_points=eqv, n_sckpts_to_skip=skip)
    assert tol2 > 1e-5 and tol2 < 4e-5
    assert np.array_equal(fs2, brute(pk2, eqv, skip, tol2))
    assert np.all(fs2 >= skip)
    # the attempt before must have left somebody unfolded
    assert np.any(brute(pk2, eqv, skip, tol2 - 0.05e-5) < 0)
    # a pc-kpt that folds nowhere: error, maps still filled
    bad = np.vstack([pk[:3], pk[3] + 0.013 * sc.B_sc[1]])
    with pytest.raises(BandupGpuError) as ei:
        unfolder.get_geom_unfolding_relations(bad, K)
    assert ei.value.code == 4


def test_pipeline_soak(unfolder):
    """300 SC-kpts of changing shape through the double-buffered submit/fetch pipeline (buffers grow
    and are reused, k-points fetched in and out of order, rho asked for now and then): every result
    must equal the first one computed for the same input, bit for bit (the path is deterministic),
    and the first ones are checked against the oracle."""
    from bandup_b200.unfold import PwWavefunction
    sc = synth.make_scenario("si333_vacancy", scale=0.12)
    grid = setup(unfolder, sc)
    inputs = []
    for iK in range(6):
        coeffs, gf, ener = synth.make_wavefunction(sc, iK)
        nb = 3 + 2 * iK
        inputs.append((PwWavefunction(np.ascontiguousarray(coeffs[:nb]), gf, ener[:nb]), sc.folding_G(iK), iK))
    first = {}
    rng = np.random.default_rng(0)
    pending = []
    for it in range(300):
        wf, fG, iK = inputs[int(rng.integers(0, len(inputs)))]
        ik = 1000 + it
        unfolder.submit(ik, wf, folding_G=fG)
        pending.append((ik, iK))
        if len(pending) == 2:
            j = int(rng.integers(0, 2))                     # fetch the older or the newer one
            ik_f, iK_f = pending.pop(j)
            res = unfolder.fetch(ik_f)
            if iK_f not in first:
                w, dN, ns, _ = oracle_kpt(sc, inputs[iK_f][0].pw_coeffs, inputs[iK_f][0].G_frac,
                                          inputs[iK_f][0].band_energies, inputs[iK_f][1], grid)
                assert np.array_equal(res.n_selected, ns)
                assert rel_err(res.spectral_weight, w) <= REL_TOL and rel_err(res.dN, dN) <= REL_TOL
                first[iK_f] = res
            else:
                ref = first[iK_f]
                assert np.array_equal(res.spectral_weight, ref.spectral_weight)
                assert np.array_equal(res.dN, ref.dN) and np.array_equal(res.band_norms, ref.band_norms)
            if it % 37 == 0:
                assert len(unfolder.calc_rho(ik_f).rho) > 0
    for ik_f, iK_f in pending:
        unfolder.fetch(ik_f)
    assert len(first) == len(inputs)


def test_host_register_caller_owned_buffer(unfolder):
    """bandup_gpu_host_register page-locks memory the caller allocated (a Fortran ALLOCATE, a NumPy
    array): same results as from pageable memory, and the upload runs at DMA rate."""
    import time
    from bandup_b200.unfold import PwWavefunction, host_register, host_unregister
    sc = synth.make_scenario("si333_vacancy", scale=0.6)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    coeffs = np.ascontiguousarray(coeffs)
    g0, _ = unfolder.folding_g0(sc.folding_G(0))
    wf = PwWavefunction(coeffs, gf, ener)
    ref = unfolder.perform_unfolding(wf, g0=g0, want_selected=False)

    def timed():
        t0 = time.perf_counter()
        for k in range(4):
            r = unfolder.perform_unfolding(wf, g0=g0, want_selected=False, ikpt=10 + k)
        return (time.perf_counter() - t0) / 4, r

    t_pageable, _ = timed()
    host_register(coeffs)
    try:
        t_locked, got = timed()
    finally:
        host_unregister(coeffs)
    assert np.array_equal(got.spectral_weight, ref.spectral_weight) and np.array_equal(got.dN, ref.dN)
    print(f"upload+unfold of {coeffs.nbytes / 1e6:.0f} MB: pageable {t_pageable * 1e3:.2f} ms, registered {t_locked * 1e3:.2f} ms")
    # no timing assertion: both go at about the PCIe rate (pageable through the staging ring), and a
    # 22 MB upload is too short to rank them reliably
    after = unfolder.perform_unfolding(wf, g0=g0, want_selected=False)        # pageable again: still fine
    assert np.array_equal(after.dN, ref.dN)


def test_maximum_pckpts_per_submit(unfolder):
    """BANDUP_GPU_MAX_PCKPTS pc-kpts in one submit (every coset of det M = 108 many times over):
    rows of pc-kpts sharing a coset are identical and match the oracle; one more is refused."""
    from bandup_b200._capi import CONST
    from bandup_b200.unfold import BandupGpuError, PwWavefunction
    n_max = CONST["BANDUP_GPU_MAX_PCKPTS"]
    sc = synth.make_scenario("si333_vacancy", scale=0.15)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 0)
    wf = PwWavefunction(coeffs, gf, ener)
    cube = np.array([[i, j, k] for i in range(6) for j in range(6) for k in range(6)], dtype=np.int32)
    g0 = cube[np.arange(n_max) % len(cube)]
    res = unfolder.perform_unfolding(wf, g0=g0, want_selected=False)
    assert res.spectral_weight.shape == (n_max, coeffs.shape[0]) and res.dN.shape == (n_max, len(grid))
    assert np.array_equal(res.spectral_weight[:len(cube)], res.spectral_weight[len(cube):2 * len(cube)])
    assert np.array_equal(res.dN[n_max - 1], res.dN[(n_max - 1) % len(cube)])
    pick = [0, 77, 215]
    w, dN, ns, _ = oracle_kpt(sc, coeffs, gf, ener, cube[pick].astype(np.float64) @ sc.B_sc, grid)
    assert np.array_equal(res.n_selected[pick], ns)
    assert rel_err(res.spectral_weight[pick], w) <= REL_TOL and rel_err(res.dN[pick], dN) <= REL_TOL
    with pytest.raises(BandupGpuError) as ei:
        unfolder.submit(7, wf, g0=np.concatenate([g0, g0[:1]]))
    assert ei.value.code == 5


@pytest.mark.parametrize("offset_bytes", [0, 8, 24])
def test_pageable_staging_handles_odd_sizes_and_misaligned_sources(unfolder, offset_bytes):
    """The pinned staging ring (streaming 64-byte stores, several copy threads, several chunks)
    with a block whose size is no multiple of 64 bytes and whose host address is only 8-byte
    aligned: same numbers as the page-locked upload of the same data."""
    from bandup_b200.unfold import PwWavefunction, pinned_empty, pinned_free
    sc = synth.make_scenario("si333_vacancy", scale=0.5)
    grid = setup(unfolder, sc)
    coeffs, gf, ener = synth.make_wavefunction(sc, 1)
    nb, n_pw, nsp = coeffs.shape
    nbytes = coeffs.nbytes
    assert nbytes > (8 << 20) and nbytes % 64 != 0          # > 8 chunks of >= 1 MB, ragged tail
    g0, _ = unfolder.folding_g0(sc.folding_G(1))
    pin = pinned_empty(nbytes)
    pin[:] = coeffs.view(np.uint8).reshape(-1)
    ref = unfolder.perform_unfolding(PwWavefunction(pin.view(np.complex64).reshape(nb, n_pw, nsp), gf, ener), g0=g0,
                                     want_selected=False)
    raw = np.empty(nbytes + 128, dtype=np.uint8)            # room for the alignment slack (< 64) plus the offset
    base = (-raw.ctypes.data) % 64 + offset_bytes            # 64-byte boundary + the requested offset
    view = raw[base:base + nbytes]
    view[:] = coeffs.view(np.uint8).reshape(-1)
    assert view.ctypes.data % 64 == offset_bytes
    got = unfolder.perform_unfolding(PwWavefunction(view.view(np.complex64).reshape(nb, n_pw, nsp), gf, ener), g0=g0,
                                     want_selected=False)
    assert np.array_equal(got.spectral_weight, ref.spectral_weight) and np.array_equal(got.dN, ref.dN)
    assert np.array_equal(got.n_selected, ref.n_selected)
    pinned_free(pin)


def test_pinned_blocks_chosen_by_measured_upload_rate(unfolder):
    """pinned_empty_near: n_keep + spares page-locked candidates, one timed upload each, the
    fastest kept (the VM hides which host socket a block lies behind); the kept blocks are ordinary
    page-locked buffers of the library's allocator."""
    from bandup_b200.unfold import _PINNED, pinned_empty_near, pinned_free
    live = len(_PINNED)
    nbytes = (24 << 20) + 40
    bufs, rates, kept = pinned_empty_near(nbytes, n_keep=2, spares=2, slice_mb=8)
    assert len(bufs) == 2 and len(rates) == 4 and kept == sorted(kept) and len(set(kept)) == 2
    assert all(r > 1.0 for r in rates) and min(rates[i] for i in kept) >= 0.98 * max(r for i, r in enumerate(rates) if i not in kept)
    assert all(b.size == nbytes and b.dtype == np.uint8 for b in bufs) and len(_PINNED) == live + 2
    bufs[0][:] = 7
    for b in bufs:
        pinned_free(b)
    one, r1, k1 = pinned_empty_near(1 << 20, spares=0)
    assert len(one) == 1 and r1 == [None] and k1 == [0]
    pinned_free(one[0])
    assert len(_PINNED) == live


@pytest.mark.parametrize("name,scale,n_kpts", [("graphene4x4", 0.08, 36), ("si333_vacancy", 0.12, 48)])
def test_device_batch_of_many_kpoints(unfolder, name, scale, n_kpts):
    """A whole (shrunken) job in ONE batch: dozens of SC k-points make K3 give every warp several
    energy bins (two or three for the graphene job, four for the Si one) and K2 mix the register and the
    shared-memory accumulators per k-point; every row against the oracle, and the second,
    identical call (descriptor block re-used on the device) gives bit-identical results."""
    import torch
    from bandup_b200.unfold import LAYOUT_FORTRAN_INMEM
    sc = synth.make_scenario(name, scale=scale, n_kpts=n_kpts)
    grid = setup(unfolder, sc)
    nener = len(grid)
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    assert sc.n_sc_kpts * ((nener + 7) // 8) > 4 * 148               # enough CTAs for several bins per warp
    kpts, host = [], []
    for iK in range(sc.n_sc_kpts):
        coeffs, gf, ener = synth.make_wavefunction(sc, iK)
        fG = sc.folding_G(iK)
        if iK == 1:                                                  # a k-point on the shared-memory bins (> 4 cosets)
            extra = np.array([[1, 0, 0], [0, 1, 0], [1, 1, 0], [2, 1, 0], [1, 2, 0]], dtype=np.float64) @ sc.B_sc
            fG = np.concatenate([fG, fG[:1] + extra])
        g0, _ = unfolder.folding_g0(fG)
        nb, n_pw, _ = coeffs.shape
        nf = g0.shape[0]
        t = dict(c=torch.from_numpy(coeffs.view(np.float32).reshape(nb, n_pw, 2)).to(dev),
                 g=torch.from_numpy(gf).to(dev), e=torch.from_numpy(ener).to(dev),
                 w=torch.zeros((nf, nb), dtype=torch.float64, device=dev),
                 dn=torch.zeros((nf, nener), dtype=torch.float64, device=dev),
                 ns=torch.zeros(nf, dtype=torch.int32, device=dev))
        host.append((coeffs, gf, ener, fG, t))
        kpts.append(dict(n_spinor=1, n_pw=n_pw, n_bands=nb, layout=LAYOUT_FORTRAN_INMEM, band_stride_bytes=n_pw * 8,
                         d_coeffs=t["c"].data_ptr(), d_g_frac=t["g"].data_ptr(), d_band_energies=t["e"].data_ptr(),
                         g0=g0, d_weights=t["w"].data_ptr(), d_dN=t["dn"].data_ptr(), d_n_selected=t["ns"].data_ptr()))
    descs, keep = unfolder.make_batch(kpts)
    ws = unfolder.batch_workspace_bytes(descs)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    unfolder.unfold_device_batch(stream, descs, d_ws.data_ptr(), ws)
    torch.cuda.synchronize()
    first = [(t["w"].clone(), t["dn"].clone()) for *_, t in host]
    unfolder.unfold_device_batch(stream, descs, d_ws.data_ptr(), ws)
    torch.cuda.synchronize()
    for (coeffs, gf, ener, fG, t), (w1, dn1) in zip(host, first):
        assert torch.equal(t["w"], w1) and torch.equal(t["dn"], dn1)
        w, dN, ns, _ = oracle_kpt(sc, coeffs, gf, ener, fG, grid)
        assert np.array_equal(t["ns"].cpu().numpy(), ns)
        assert rel_err(t["w"].cpu().numpy(), w) <= REL_TOL
        assert rel_err(t["dn"].cpu().numpy(), dN) <= REL_TOL


@pytest.mark.parametrize("n_spinor,layout_name", [(1, "LAYOUT_FORTRAN_INMEM"), (2, "LAYOUT_FORTRAN_INMEM"),
                                                  (2, "LAYOUT_VASP_RECORD")])
def test_device_batch_writes_stay_inside_their_buffers(unfolder, n_spinor, layout_name):
    """Guard bands (compute-sanitizer is not available on the GPU pool): inputs, outputs and the
    workspace of a batch are carved out of ONE arena filled with a pattern, 256 bytes apart; after
    the batch -- k-points with even and odd n_pw, 1 / 3 / 9 / 70 distinct cosets, i.e. register
    accumulators, shared-memory bins and a second pass, plus replicated rows -- every byte outside
    the outputs and the workspace (of exactly the size the library asked for) is what it was, every
    output element has been written, and the weights of a band over all cosets sum to one."""
    import torch
    from bandup_b200 import unfold as U
    layout = getattr(U, layout_name)
    sc = synth.make_scenario("si333_vacancy", scale=0.2)
    grid = setup(unfolder, sc)
    nener = len(grid)
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    cube = np.array([[i, j, k] for i in range(6) for j in range(6) for k in range(6)], dtype=np.int32)
    lab = synth.coset_labels(cube, sc.M)
    _, first = np.unique(lab, return_index=True)
    reps = cube[np.sort(first)]                                       # one vector per coset (108 of them)
    assert len(reps) == 108
    rng = np.random.default_rng(77)
    GUARD = 256
    plan, cursor = [], GUARD

    def carve(nbytes, align=16, shift=0):
        nonlocal cursor
        off = (cursor + align - 1) // align * align + shift
        cursor = off + nbytes + GUARD
        return off

    kpts_host = []
    for iK, (nb, trim, g0) in enumerate([(7, 0, reps[:1]), (5, 1, reps[3:6]), (6, 0, reps[10:19]), (4, 1, reps[:70]),
                                         (3, 0, np.concatenate([reps, reps[:5]]))]):
        gf = sc.gvecs(iK)
        if trim:
            gf = gf[:-1] if gf.shape[0] % 2 == 0 else gf
        else:
            gf = gf if gf.shape[0] % 2 == 0 else gf[:-1]
        n_pw = gf.shape[0]
        assert n_pw % 2 == trim
        row = n_pw * n_spinor
        coeffs = (rng.standard_normal((nb, row)) + 1j * rng.standard_normal((nb, row))).astype(np.complex64)
        ener = np.sort(rng.uniform(sc.e_fermi - 20, sc.e_fermi + 5, nb))
        nf = len(g0)
        # odd scalar rows make every other band start 8 bytes off a 16-byte boundary anyway; shift one block too
        o = dict(c=carve(coeffs.nbytes, shift=8 if iK == 2 else 0), g=carve(gf.nbytes), e=carve(ener.nbytes),
                 w=carve(nf * nb * 8), dn=carve(nf * nener * 8), ns=carve(nf * 4), nr=carve(nb * 8))
        kpts_host.append((coeffs, np.ascontiguousarray(gf), ener, np.ascontiguousarray(g0), o, nb, n_pw, nf))
    arena_bytes = cursor
    arena = torch.full((arena_bytes,), 0xA5, dtype=torch.uint8, device=dev)
    base = arena.data_ptr()
    assert base % 256 == 0
    writable = torch.zeros(arena_bytes, dtype=torch.bool, device=dev)
    kpts = []
    for coeffs, gf, ener, g0, o, nb, n_pw, nf in kpts_host:
        for key, arr in (("c", coeffs), ("g", gf), ("e", ener)):
            arena[o[key]:o[key] + arr.nbytes].copy_(torch.from_numpy(arr.view(np.uint8).reshape(-1)))
        for key, n in (("w", nf * nb * 8), ("dn", nf * nener * 8), ("ns", nf * 4), ("nr", nb * 8)):
            writable[o[key]:o[key] + n] = True
        kpts.append(dict(n_spinor=n_spinor, n_pw=n_pw, n_bands=nb, layout=layout, band_stride_bytes=n_pw * n_spinor * 8,
                         d_coeffs=base + o["c"], d_g_frac=base + o["g"], d_band_energies=base + o["e"], g0=g0,
                         d_weights=base + o["w"], d_dN=base + o["dn"], d_n_selected=base + o["ns"],
                         d_norms=base + o["nr"]))
    descs, keep = unfolder.make_batch(kpts)
    ws = unfolder.batch_workspace_bytes(descs)
    ws_arena = torch.full((GUARD + ws + GUARD,), 0x5A, dtype=torch.uint8, device=dev)
    before = arena.clone()
    for variant in (0, 1, 2):
        unfolder.set_option(U.CONST["BANDUP_GPU_OPT_K2_VARIANT"], variant)
        unfolder.unfold_device_batch(stream, descs, ws_arena.data_ptr() + GUARD, ws)
        torch.cuda.synchronize()
        assert bool(torch.all((arena == before) | writable)), f"variant {variant}: a write outside the output buffers"
        assert bool(torch.all(ws_arena[:GUARD] == 0x5A)) and bool(torch.all(ws_arena[GUARD + ws:] == 0x5A)), \
            f"variant {variant}: a write outside the workspace"
        host_arena = arena.cpu().numpy()
        for coeffs, gf, ener, g0, o, nb, n_pw, nf in kpts_host:
            w = host_arena[o["w"]:o["w"] + nf * nb * 8].view(np.float64).reshape(nf, nb)
            dn = host_arena[o["dn"]:o["dn"] + nf * nener * 8].view(np.float64)
            ns = host_arena[o["ns"]:o["ns"] + nf * 4].view(np.int32)
            nr = host_arena[o["nr"]:o["nr"] + nb * 8].view(np.float64)
            pattern = np.frombuffer(bytes([0xA5]) * 8, dtype=np.float64)[0]
            assert not np.any(w == pattern) and not np.any(dn == pattern) and not np.any(nr == pattern)
            assert np.all(ns >= 0) and np.all(ns <= n_pw)
            assert np.all(np.isfinite(w)) and np.all(w >= 0) and np.all(w <= 1 + 1e-9)
            if nf >= 108:
                assert int(ns[:108].sum()) == n_pw
                assert np.allclose(w[:108].sum(axis=0), 1.0, atol=1e-7)
                assert np.array_equal(w[108:], w[:5])                 # the replicated rows
        arena.copy_(before)
    unfolder.set_option(U.CONST["BANDUP_GPU_OPT_K2_VARIANT"], 0)
