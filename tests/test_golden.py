"""Oracle vs the golden vectors generated from the reference's own Python
(tests/golden/make_golden.py imports bandupy.unfolding_density_op / orbital_contributions)."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle_np as ON

GOLD = Path(__file__).resolve().parent / "golden"


def test_unf_dens_op_unfold_matches_reference_python():
    cases = json.loads((GOLD / "udo_golden.json").read_text())["cases"]
    assert len(cases) >= 10
    for c in cases:
        rho = np.array(c["rho_re"]) + 1j * np.array(c["rho_im"])
        O = np.array(c["O_re"]) + 1j * np.array(c["O_im"])
        N = c["unfolded_N"]
        got_clip = ON.unf_dens_op_unfold(c["rows"], c["cols"], rho, c["nbands"], N, O, True, True)
        got_raw = ON.unf_dens_op_unfold(c["rows"], c["cols"], rho, c["nbands"], N, O, True, False)
        assert abs(got_clip - c["unfold_clipped"]) <= 1e-12 * max(1.0, abs(c["unfold_clipped"]))
        assert abs(got_raw - c["unfold_times_N"]) <= 1e-12 * max(1.0, abs(c["unfold_times_N"]))
        assert 0.0 <= got_clip <= N
        # the upper-triangle helper used by the GPU tests agrees as well
        via_pairs = ON.unfold_rho_operator(c["rows"], c["cols"], rho, O, N, min_abs_rho=0.0, clip=True)
        assert abs(via_pairs - c["unfold_clipped"]) <= 1e-12 * max(1.0, abs(c["unfold_clipped"]))


def test_orbital_contribution_matrix_matches_reference_python():
    g = np.load(GOLD / "orb_contrib_golden.npz")
    meta = json.loads(str(g["meta"]))
    assert len(meta) == 3
    for i, m in enumerate(meta):
        O = ON.orb_contrib_matrix(g[f"c{i}_dual"], g[f"c{i}_proj"], g[f"c{i}_mask"], g[f"c{i}_ener"],
                                  m["max_band_dE"], m["min_abs"])
        ref = g[f"c{i}_O"]
        assert np.count_nonzero(O) == m["nnz"] == np.count_nonzero(ref)
        assert np.array_equal(O != 0, ref != 0)
        assert np.allclose(O, ref, rtol=1e-12, atol=1e-14)


def test_rho_text_file_is_what_the_reference_reader_parses(tmp_path):
    """bandup_files.write_unf_dens_op produces unfolding_density_operator_*.dat; the golden holds what
    the REFERENCE's read_unf_dens_ops parsed out of exactly that file.  Our writer must still
    produce the same bytes, and our reader restatement must parse the same operators."""
    from bandup_b200 import bandup_files as BF
    from bandup_b200 import synth
    inp = np.load(GOLD / "udo_file_inputs.npz")
    a = synth.hex_cell(2.467, 15.0)
    b = synth.rec_latt(a)
    B = synth.rec_latt(np.diag([4, 4, 1]) @ a)
    kp = BF.PcKptsPath([BF.kpts_line(np.array([1 / 3, 1 / 3, 0]) @ b, np.zeros(3), 3),
                        BF.kpts_line(np.zeros(3), np.array([0.5, 0, 0]) @ b, 2)], 0.0)
    ops = [(inp[f"op{i}_ie"], inp[f"op{i}_m1"], inp[f"op{i}_m2"], inp[f"op{i}_v"]) for i in range(5)]
    out = tmp_path / "udo.dat"
    BF.write_unf_dens_op(out, kp, inp["grid"], inp["dN"], ops, 12, B, b, inp["sck_of"], inp["sck"], e_fermi=-1.5)
    assert out.read_text() == (GOLD / "udo_file_fixture.dat").read_text()
    hdr, recs = BF.read_unf_dens_op(out)
    gold = json.loads((GOLD / "udo_file_golden.json").read_text())["records"]
    assert len(recs) == len(gold) > 0
    assert hdr["nbands"] == gold[0]["nbands"] and hdr["nener"] == gold[0]["nener"]
    assert abs(hdr["emin"] - gold[0]["emin"]) < 1e-12 and abs(hdr["emax"] - gold[0]["emax"]) < 1e-12
    for r, g in zip(recs, gold):
        assert (r.pckpt_number, r.iener, r.folding_sckpt_number) == (g["pckpt_number"], g["iener"],
                                                                     g["folding_sckpt_number"])
        assert r.energy == g["energy"] and r.unfolded_N == g["unfolded_N"]
        assert r.pckpt_coord_in_plot == g["pckpt_coord_in_plot"]
        # Hermitian completion with a real diagonal, as UnfDensOp.__init__ does
        dense = np.zeros((12, 12), dtype=complex)
        for i, j, v in zip(r.rows, r.cols, r.entries):
            if i == j:
                dense[i, j] += v.real
            else:
                dense[i, j] += v
                dense[j, i] += np.conj(v)
        want = np.zeros((12, 12), dtype=complex)
        want[g["hermitian_rows"], g["hermitian_cols"]] = np.array(g["hermitian_re"]) + 1j * np.array(g["hermitian_im"])
        assert np.array_equal(dense, want)


def test_orbital_choice_matches_the_reference_function():
    """formatted_orb_choice against what the reference's own function returned for the same 44 inputs
    (tests/golden/orb_choice_golden.json, made by importing bandupy.orbital_contributions), and
    ao_mask's combined atom/orbital index (KptInfo.combined_atom_orb_index: iat*n_orbs + iorb)."""
    import json
    from pathlib import Path
    from bandup_b200 import bandup_files as BF
    cases = json.loads((Path(__file__).parent / "golden" / "orb_choice_golden.json").read_text())
    assert len(cases) == 44
    for c in cases:
        assert BF.formatted_orb_choice(c["choice"], c["supported"]) == c["picked"], c["choice"]
    with pytest.raises(ValueError):
        BF.formatted_orb_choice("sp5", cases[0]["supported"])
    orbitals = cases[0]["supported"]
    mask = BF.ao_mask(4, orbitals, "p", [1, 3])
    want = np.zeros((4, len(orbitals)), dtype=np.uint8)
    want[[1, 3], 1:4] = 1
    assert np.array_equal(mask.reshape(4, len(orbitals)), want)
