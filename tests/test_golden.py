"""Oracle vs the golden vectors generated from the reference's own Python
(tests/golden/make_golden.py imports bandupy.unfolding_density_op / orbital_contributions)."""
import json
from pathlib import Path

import numpy as np

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
