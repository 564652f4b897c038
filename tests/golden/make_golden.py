#!/usr/bin/env python
"""Generates tests/golden/udo_golden.json and orb_contrib_golden.npz by IMPORTING THE REFERENCE's
own Python (bandupy.unfolding_density_op.UnfDensOp, bandupy.orbital_contributions.KptInfo) from
/root/reference.  Run in the build container only (the reference does not travel to the GPU box);
the vectors are committed.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

matplotlib is not installed here; bandupy.constants only calls mpl.get_backend() at import time, so a
two-attribute stub module is pre-seeded in sys.modules.  No reference source is modified or copied.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src/python_interface")
mpl = types.ModuleType("matplotlib")
mpl.get_backend = lambda: "agg"
mpl.use = lambda *a, **k: None
mpl.__version__ = "0.0"
sys.modules.setdefault("matplotlib", mpl)

from bandupy.unfolding_density_op import UnfDensOp  # noqa: E402

HERE = Path(__file__).resolve().parent
rng = np.random.default_rng(20261019)


def make_udo_cases():
    cases = []
    for case in range(12):
        nb = int(rng.integers(3, 40))
        n_ent = int(rng.integers(1, min(60, nb * (nb + 1) // 2)))
        pairs = set()
        while len(pairs) < n_ent:
            a, b = sorted(rng.integers(0, nb, size=2).tolist())
            pairs.add((a, b))                      # the rho file holds the upper triangle only
        pairs = sorted(pairs)
        rows = [p[0] for p in pairs]
        cols = [p[1] for p in pairs]
        vals = (rng.standard_normal(n_ent) + 1j * rng.standard_normal(n_ent)) * 0.3
        N = float(rng.uniform(0.001, 3.0))
        O = (rng.standard_normal((nb, nb)) + 1j * rng.standard_normal((nb, nb))) * 0.4
        O[rng.random((nb, nb)) < 0.5] = 0.0        # sparse like the thresholded contribution matrix
        udo = UnfDensOp(pckpt_number=1, pckpt_cart_coords=[0, 0, 0], pckpt_frac_coords_scrl=[0, 0, 0],
                        pckpt_frac_coords_pcrl=[0, 0, 0], pckpt_coord_in_plot=0.0, folding_sckpt_number=1,
                        folding_sckpt_cart_coords=[0, 0, 0], folding_sckpt_frac_coords_scrl=[0, 0, 0],
                        nbands=nb, nener_parent_grid=501, emin_parent_grid=-20.0, emax_parent_grid=5.0,
                        iener=7, energy=-19.7, unfolded_N=N, row_indices=rows, col_indices=cols,
                        entries=list(vals))
        from scipy.sparse import csr_matrix
        Osp = csr_matrix(O)
        clipped = udo.unfold(Osp, multiply_by_N=True, discard_imag=True, clip_interval=(0.0, N))
        raw = udo.unfold(Osp, multiply_by_N=True, discard_imag=True)
        plain = udo.unfold(Osp)                    # complex Tr(rho.O)
        cases.append(dict(nbands=nb, unfolded_N=N, rows=rows, cols=cols,
                          rho_re=vals.real.tolist(), rho_im=vals.imag.tolist(),
                          O_re=O.real.tolist(), O_im=O.imag.tolist(),
                          unfold_clipped=float(clipped), unfold_times_N=float(raw),
                          trace_rho_O=[float(np.real(plain)), float(np.imag(plain))],
                          trace_rho=[float(np.real(udo.trace)), float(np.imag(udo.trace))]))
    return cases


def make_orb_contrib_cases():
    from bandupy.orbital_contributions import KptInfo
    out = {}
    meta = []
    orbitals = ['s', 'py', 'pz', 'px', 'dxy', 'dyz', 'dz2', 'dxz', 'dx2']
    for case, (nb, nions) in enumerate([(12, 3), (30, 5), (24, 4)]):
        kp = KptInfo(number=1, frac_coords=[0, 0, 0], weight=1.0, nbands=nb, nions=nions, orbitals=orbitals,
                     create_ion_proj_norms_dicts=False)
        ener = np.sort(rng.uniform(-3.0, 3.0, nb))
        ener[1] = ener[0]                           # exact degeneracy
        ener[4] = ener[3] + 0.1                     # exactly on the max_band_dE edge (float compare)
        for ib in range(nb):
            kp.bands[ib]['ener'] = float(ener[ib])
        n_ao = nions * len(orbitals)
        proj = (rng.standard_normal((n_ao, nb)) + 1j * rng.standard_normal((n_ao, nb))) * 0.2
        dual = np.linalg.pinv(proj)
        kp.proj_matrix = proj
        kp.proj_matrix_dual = dual
        picks = [('all', None), ('p', [0, 2]), (['s', 'dz2'], [1])][case]
        O = kp.get_orbital_contribution_matrix(picked_orbitals=picks[0], selected_ion_indices=picks[1])
        from bandupy.orbital_contributions import formatted_orb_choice
        picked = formatted_orb_choice(picks[0], supported=orbitals)
        ions = list(range(nions)) if picks[1] is None else picks[1]
        mask = np.zeros(n_ao, dtype=np.uint8)
        for iat in ions:
            for orb in picked:
                mask[iat * len(orbitals) + orbitals.index(orb)] = 1
        out[f"c{case}_proj"] = proj
        out[f"c{case}_dual"] = dual
        out[f"c{case}_ener"] = ener
        out[f"c{case}_mask"] = mask
        out[f"c{case}_O"] = O.toarray()
        meta.append(dict(nbands=nb, nions=nions, n_orbs=len(orbitals), picked=picked, ions=ions,
                         max_band_dE=0.1, min_abs=1e-2, nnz=int(O.nnz)))
    out["meta"] = np.array(json.dumps(meta))
    return out


if __name__ == "__main__":
    (HERE / "udo_golden.json").write_text(json.dumps(
        {"source": "bandupy.unfolding_density_op.UnfDensOp (reference, imported unmodified)",
         "cases": make_udo_cases()}))
    np.savez_compressed(HERE / "orb_contrib_golden.npz", **make_orb_contrib_cases())
    print("wrote", HERE / "udo_golden.json", HERE / "orb_contrib_golden.npz")
