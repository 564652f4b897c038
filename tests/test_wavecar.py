"""WAVECAR container (CPU): the writer produces the record layout read_wavecar_mod.f90 parses, and
the parser recovers lattice, k-points, energies, plane-wave lists, spinor flag and coefficients."""
import numpy as np
import pytest

from bandup_b200 import synth
from bandup_b200.wavecar import Wavecar, fold_pckpts_onto_sckpts, write_wavecar


def _make(tmp_path, name, scale, n_k=3, extra_pad=0):
    sc = synth.make_scenario(name, scale=scale, n_kpts=9)
    ks = list(range(min(n_k, sc.n_sc_kpts)))
    data = [synth.make_wavefunction(sc, iK) for iK in ks]
    path = tmp_path / "WAVECAR"
    nrecl = write_wavecar(path, sc.A_sc, sc.ecut, [sc.sc_kpts_frac[iK] for iK in ks],
                          [d[0] for d in data], [d[2] for d in data], extra_pad=extra_pad)
    return sc, ks, data, path, nrecl


@pytest.mark.parametrize("name,scale,pad", [("graphene4x4", 0.1, 0), ("mos2_8x8_spinor", 0.05, 3),
                                            ("si333_vacancy", 0.2, 1)])
def test_roundtrip(tmp_path, name, scale, pad):
    sc, ks, data, path, nrecl = _make(tmp_path, name, scale, extra_pad=pad)
    wc = Wavecar(path)
    assert wc.nrecl == nrecl and wc.nspin == 1 and wc.nkpts == len(ks) and wc.nbands == sc.n_bands
    assert path.stat().st_size == nrecl * (2 + len(ks) * (1 + sc.n_bands))
    assert wc.encut == sc.ecut and np.array_equal(wc.A_matrix, sc.A_sc)
    assert np.allclose(wc.B_matrix, sc.B_sc, rtol=1e-14)
    for i, iK in enumerate(ks):
        coeffs, gf, ener = data[i]
        h = wc.kpt_header(i + 1)
        assert h.nplane == gf.shape[0] * sc.n_spinor
        assert np.array_equal(h.kpt_frac_coords, sc.sc_kpts_frac[iK])
        assert np.array_equal(h.band_energies, ener) and np.all(h.band_occupations == 1.0)
        g, nsp = wc.gvecs(i + 1)
        assert nsp == sc.n_spinor and np.array_equal(g, gf)
        assert np.array_equal(wc.coefficients(i + 1), coeffs)
        raw = wc.read_band_records(i + 1)
        assert raw.size == sc.n_bands * nrecl
        # band record = spin-up block then spin-down block, zero padded to nrecl
        rec0 = raw[:nrecl].view("<c8")
        assert np.array_equal(rec0[:gf.shape[0]], coeffs[0, :, 0])
        if sc.n_spinor == 2:
            assert np.array_equal(rec0[gf.shape[0]:2 * gf.shape[0]], coeffs[0, :, 1])
        assert not rec0[gf.shape[0] * sc.n_spinor:].any()
        wf = wc.wavefunction(i + 1)
        assert wf.band_stride_bytes == nrecl and wf.n_pw == gf.shape[0] and wf.n_spinor == sc.n_spinor
    with pytest.raises(IndexError):
        wc.kpt_header(len(ks) + 1)


def test_rejects_double_precision_and_garbage(tmp_path):
    p = tmp_path / "W"
    p.write_bytes(np.array([96.0, 1.0, 45210.0] + [0.0] * 21, dtype="<f8").tobytes())
    with pytest.raises(ValueError):
        Wavecar(p)
    p.write_bytes(np.array([96.0, 1.0, 45200.0] + [0.0] * 21, dtype="<f8").tobytes())
    with pytest.raises(ValueError):      # nkpts == 0: 'Unexpected file format'
        Wavecar(p)


def test_fold_map_matches_scenario():
    sc = synth.make_scenario("graphene4x4", scale=0.1, n_kpts=12)
    folds = fold_pckpts_onto_sckpts(sc.pc_kpts_cart, sc.sc_kpts_frac, sc.B_sc)
    assert folds == sc.folds
    with pytest.raises(ValueError):
        fold_pckpts_onto_sckpts(sc.pc_kpts_cart, sc.sc_kpts_frac[:1], sc.B_sc)


def test_second_spin_channel_record_indexing(tmp_path):
    """nspin = 2: the records of spin 2 follow all k-points of spin 1 (read_wavecar_mod.f90:409-414)."""
    sc = synth.make_scenario("graphene4x4", scale=0.06, n_kpts=6)
    ks = [0, 1]
    data = [synth.make_wavefunction(sc, iK) for iK in ks]
    path = tmp_path / "WAVECAR"
    nrecl = write_wavecar(path, sc.A_sc, sc.ecut, [sc.sc_kpts_frac[i] for i in ks], [d[0] for d in data],
                          [d[2] for d in data], nspin=2)
    wc = Wavecar(path)
    assert wc.nspin == 2 and path.stat().st_size == nrecl * (2 + 2 * 2 * (1 + sc.n_bands))
    assert wc.header_record(1, 2) == 2 + 2 * (1 + sc.n_bands)
    for ispin in (1, 2):
        for i in range(2):
            assert np.array_equal(wc.coefficients(i + 1, ispin), data[i][0])
            assert np.array_equal(wc.kpt_header(i + 1, ispin).band_energies, data[i][2])
    with pytest.raises(IndexError):
        wc.header_record(1, 3)
