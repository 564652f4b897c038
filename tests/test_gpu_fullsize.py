"""Full-size parity on BASELINE configs 3 and 5: EVERY band of one SC k-point against the oracle,
on k-points with an odd and with an even number of plane waves (odd n_pw: every other band row of
a scalar block starts 8 bytes off a 16-byte boundary, the head/tail peel of K2), through every way
into the kernels:

  * bandup_gpu_unfold_device_batch on HBM-resident coefficients -- with both accumulator layouts of
    K2 (shared-memory bins, register accumulators) and with the block shifted by 8 bytes;
  * the host path bandup_gpu_submit_kpt / fetch_kpt from page-locked memory, from pageable memory
    (the library's staging ring) and from raw WAVECAR band records ([band][spinor][pw], stride
    nrecl > row: VASP_RECORD with padding).

Tolerances as everywhere: n_selected exact; weights, norms and delta_N within 1e-6 relative (abs
floor 1e-12) of the exact-fp64-norm oracle.  The oracle runs a 2.4 GB k-point in well under a
second on the box's cores, so nothing is sampled."""
import numpy as np
import pytest

from bandup_b200 import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6
ABS_FLOOR = 1e-12
SEED = 20261020


def rel_err(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), ABS_FLOOR))) if a.size else 0.0


def _check(tag, got_w, got_dn, got_ns, got_norms, want):
    w, dN, ns, norms = want
    assert np.array_equal(got_ns, ns), tag
    assert got_w.shape == w.shape and rel_err(got_w, w) <= REL_TOL, (tag, rel_err(got_w, w))
    assert rel_err(got_dn, dN) <= REL_TOL, (tag, rel_err(got_dn, dN))
    if got_norms is not None:
        assert rel_err(got_norms, norms) <= REL_TOL, (tag, rel_err(got_norms, norms))


# (scenario, SC-kpt, parity of n_pw): C5 iK=1 has 100381 plane waves, iK=3 100468;
# C3 iK=0 has 79041 (and 7 pc-kpts: the bins layout), iK=4 78912, iK=5 78879 (2 pc-kpts each)
CASES = [("sige555", 1, 1), ("sige555", 3, 0), ("mos2_8x8_spinor", 0, 1), ("mos2_8x8_spinor", 4, 0),
         ("mos2_8x8_spinor", 5, 1)]


@pytest.mark.parametrize("name,iK,odd", CASES)
def test_every_band_full_size_all_entry_points(unfolder, name, iK, odd):
    import torch
    from bandup_b200._capi import CONST
    from bandup_b200.unfold import (LAYOUT_FORTRAN_INMEM, LAYOUT_VASP_RECORD, PwWavefunction, pinned_empty,
                                    pinned_free)
    from oracle import oracle as O
    sc = synth.make_scenario(name)
    assert np.array_equal(unfolder.verify_commens(sc.b_pc, sc.B_sc), sc.M)
    grid = unfolder.set_energy_grid(*sc.energy_window())
    gf = sc.gvecs(iK)
    n_pw, nb, nsp = gf.shape[0], sc.n_bands, sc.n_spinor
    assert n_pw % 2 == odd
    row = n_pw * nsp
    fG = sc.folding_G(iK)
    g0, _ = unfolder.folding_g0(fG)
    nf = g0.shape[0]
    rng = np.random.default_rng(100 + iK)
    ener = np.sort(rng.uniform(sc.e_fermi - 20.5, sc.e_fermi + 5.5, nb))
    block = O.synth_fill(row, nb, SEED, iK)                       # (nb, row) complex64, [pw][spinor] interleaved
    nbytes = block.nbytes

    # ---- device-resident batch entry point: both K2 accumulator layouts, aligned and shifted by 8 bytes
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    d_raw = torch.empty(nbytes + 16, dtype=torch.uint8, device=dev)
    d_g, d_e = torch.from_numpy(gf).to(dev), torch.from_numpy(ener).to(dev)
    results = {}
    for shift in (0, 8):
        d_raw[shift:shift + nbytes].copy_(torch.from_numpy(block.view(np.uint8).reshape(-1)))
        for variant in (1, 2):
            unfolder.set_option(CONST["BANDUP_GPU_OPT_K2_VARIANT"], variant)
            d_w = torch.zeros((nf, nb), dtype=torch.float64, device=dev)
            d_dn = torch.zeros((nf, len(grid)), dtype=torch.float64, device=dev)
            d_ns = torch.zeros(nf, dtype=torch.int32, device=dev)
            d_nr = torch.zeros(nb, dtype=torch.float64, device=dev)
            descs, keep = unfolder.make_batch([dict(
                n_spinor=nsp, n_pw=n_pw, n_bands=nb, layout=LAYOUT_FORTRAN_INMEM, band_stride_bytes=row * 8,
                d_coeffs=d_raw.data_ptr() + shift, d_g_frac=d_g.data_ptr(), d_band_energies=d_e.data_ptr(), g0=g0,
                d_weights=d_w.data_ptr(), d_dN=d_dn.data_ptr(), d_n_selected=d_ns.data_ptr(),
                d_norms=d_nr.data_ptr())])
            ws = unfolder.batch_workspace_bytes(descs)
            d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
            unfolder.unfold_device_batch(stream, descs, d_ws.data_ptr(), ws)
            torch.cuda.synchronize()
            results[f"device shift={shift} variant={variant}"] = (d_w.cpu().numpy(), d_dn.cpu().numpy(),
                                                                  d_ns.cpu().numpy(), d_nr.cpu().numpy())
            del d_ws
    unfolder.set_option(CONST["BANDUP_GPU_OPT_K2_VARIANT"], 0)
    del d_raw

    # ---- host path: page-locked, pageable, raw records with nrecl padding
    def fetch(tag, ikpt):
        r = unfolder.fetch(ikpt)
        results[tag] = (r.spectral_weight, r.dN, r.n_selected, r.band_norms)

    pin = pinned_empty(nbytes)
    pin[:] = block.view(np.uint8).reshape(-1)
    unfolder.submit(1, PwWavefunction(pin.view(np.complex64).reshape(nb, n_pw, nsp), gf, ener), g0=g0)
    fetch("host pinned", 1)
    pinned_free(pin)
    unfolder.submit(2, PwWavefunction(block.reshape(nb, n_pw, nsp), gf, ener), g0=g0)
    fetch("host pageable", 2)
    nrecl = row * 8 + 40                                          # band records padded as a WAVECAR's are
    rec = np.zeros((nb, nrecl), dtype=np.uint8)
    # [band][spinor][pw]: the spinor-up block then the spinor-down block (read_wavecar_mod.f90:570-577)
    blocked = np.ascontiguousarray(block.reshape(nb, n_pw, nsp).transpose(0, 2, 1)).reshape(nb, row)
    rec[:, :row * 8] = blocked.view(np.uint8)
    del blocked
    unfolder.submit(3, PwWavefunction(rec, gf, ener, n_pw=n_pw, n_bands=nb, n_spinor=nsp, layout=LAYOUT_VASP_RECORD,
                                      band_stride_bytes=nrecl), g0=g0)
    fetch("host VASP_RECORD nrecl-padded", 3)
    del rec

    # ---- the oracle, every band (renormalises its input in place: run it last)
    _, norms = O.renormalize(block.reshape(nb, n_pw, nsp), 1)
    w, dN, ns, _ = O.unfold_kpt(block.reshape(nb, n_pw, nsp), gf @ sc.B_sc, ener, sc.b_pc, fG, grid, norm_mode=1,
                                in_place=True)
    assert ns.min() > 0
    for tag, got in results.items():
        _check(f"{name} iK={iK} {tag}", *got, (w, dN, ns, norms))
    # size-independent properties on the full arrays
    assert np.all(w >= 0.0) and np.all(results["host pinned"][0].sum(axis=0) <= 1.0 + 1e-7)
    de = grid[1] - grid[0]
    inside = (ener > grid[0] - 0.5 * de) & (ener < grid[-1] + 0.5 * de)
    got_w, got_dn = results["host pinned"][0], results["host pinned"][1]
    assert np.allclose(got_dn.sum(axis=1), got_w[:, inside].sum(axis=1), rtol=1e-12)


def test_block_larger_than_4_GiB(unfolder):
    """Maximum sizes: one SC k-point whose coefficient block does not fit 32-bit byte offsets
    (5400 bands x 100381 plane waves x 8 B = 4.34 GB > 2^32; rows alternate between 16-byte aligned
    and 8 bytes off), through the device-resident batch entry and through the host path from
    pageable memory (staging ring, chunk offsets beyond 4 GiB).  Every band against the oracle."""
    import torch
    from bandup_b200.unfold import LAYOUT_FORTRAN_INMEM, PwWavefunction
    from oracle import oracle as O
    sc = synth.make_scenario("sige555")
    assert np.array_equal(unfolder.verify_commens(sc.b_pc, sc.B_sc), sc.M)
    grid = unfolder.set_energy_grid(*sc.energy_window())
    iK, nb = 1, 5400
    gf = sc.gvecs(iK)
    n_pw = gf.shape[0]
    assert n_pw % 2 == 1 and nb * n_pw * 8 > (1 << 32)
    fG = sc.folding_G(iK)
    g0, _ = unfolder.folding_g0(fG)
    nf = g0.shape[0]
    rng = np.random.default_rng(5)
    ener = np.sort(rng.uniform(sc.e_fermi - 20.5, sc.e_fermi + 5.5, nb))
    block = O.synth_fill(n_pw, nb, SEED, iK)
    nbytes = block.nbytes

    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream(dev).cuda_stream
    d_c = torch.from_numpy(block.view(np.uint8).reshape(-1)).to(dev)
    d_g, d_e = torch.from_numpy(gf).to(dev), torch.from_numpy(ener).to(dev)
    d_w = torch.zeros((nf, nb), dtype=torch.float64, device=dev)
    d_dn = torch.zeros((nf, len(grid)), dtype=torch.float64, device=dev)
    d_ns = torch.zeros(nf, dtype=torch.int32, device=dev)
    d_nr = torch.zeros(nb, dtype=torch.float64, device=dev)
    descs, keep = unfolder.make_batch([dict(
        n_spinor=1, n_pw=n_pw, n_bands=nb, layout=LAYOUT_FORTRAN_INMEM, band_stride_bytes=n_pw * 8,
        d_coeffs=d_c.data_ptr(), d_g_frac=d_g.data_ptr(), d_band_energies=d_e.data_ptr(), g0=g0,
        d_weights=d_w.data_ptr(), d_dN=d_dn.data_ptr(), d_n_selected=d_ns.data_ptr(), d_norms=d_nr.data_ptr())])
    ws = unfolder.batch_workspace_bytes(descs)
    d_ws = torch.empty(ws, dtype=torch.uint8, device=dev)
    unfolder.unfold_device_batch(stream, descs, d_ws.data_ptr(), ws)
    torch.cuda.synchronize()
    results = {"device": (d_w.cpu().numpy(), d_dn.cpu().numpy(), d_ns.cpu().numpy(), d_nr.cpu().numpy())}
    del d_c, d_ws

    unfolder.submit(1, PwWavefunction(block.reshape(nb, n_pw, 1), gf, ener), g0=g0)   # pageable NumPy memory
    r = unfolder.fetch(1)
    results["host pageable"] = (r.spectral_weight, r.dN, r.n_selected, r.band_norms)

    _, norms = O.renormalize(block.reshape(nb, n_pw, 1), 1)
    w, dN, ns, _ = O.unfold_kpt(block.reshape(nb, n_pw, 1), gf @ sc.B_sc, ener, sc.b_pc, fG, grid, norm_mode=1,
                                in_place=True)
    assert ns.min() > 0 and nbytes > (1 << 32)
    for tag, got in results.items():
        _check(f"4 GiB block, {tag}", *got, (w, dN, ns, norms))
    # the last band lies beyond the 4 GiB mark: its weights are real numbers, not a wrapped row's
    assert np.all(results["device"][0][:, -1] > 0)
