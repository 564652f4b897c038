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


def make_udo_file_golden():
    """Our writer (bandup_b200.bandup_files.write_unf_dens_op) -> the REFERENCE's reader
    (bandupy.unfolding_density_op.read_unf_dens_ops): what the reference parses out of our file."""
    sys.path.insert(0, str(HERE.parent.parent))
    from bandup_b200 import bandup_files as BF
    from bandup_b200 import synth
    from bandupy.unfolding_density_op import read_unf_dens_ops
    a = synth.hex_cell(2.467, 15.0)
    b = synth.rec_latt(a)
    B = synth.rec_latt(np.diag([4, 4, 1]) @ a)
    kp = BF.PcKptsPath([BF.kpts_line(np.array([1 / 3, 1 / 3, 0]) @ b, np.zeros(3), 3),
                        BF.kpts_line(np.zeros(3), np.array([0.5, 0, 0]) @ b, 2)], 0.0)
    grid = np.array([-3.0 + 0.05 * i for i in range(40)])
    nk, nb = 5, 12
    r = np.random.default_rng(77)
    dN = r.uniform(0.0, 2.0, (nk, grid.size))
    dN[r.random(dN.shape) < 0.5] = 1e-4                  # below min_allowed_dN: skipped by the writer
    ops = []
    for ipc in range(nk):
        ie_l, m1_l, m2_l, v_l = [], [], [], []
        for ie in sorted(r.choice(np.arange(1, grid.size + 1), size=4, replace=False)):
            bands = sorted(r.choice(np.arange(1, nb + 1), size=3, replace=False))
            for x in bands:
                for y in bands:
                    if y < x:
                        continue
                    v = complex(r.uniform(0.05, 0.6), 0.0) if x == y else complex(r.normal(0, 0.2), r.normal(0, 0.2))
                    if r.random() < 0.15:
                        v *= 1e-4                         # below the writer's |rho| >= 1e-4 filter
                    for (p, q, w) in ([(x, y, v)] if x == y else [(x, y, v), (y, x, np.conj(v))]):
                        ie_l.append(ie); m1_l.append(p); m2_l.append(q); v_l.append(w)
        ops.append((np.array(ie_l), np.array(m1_l), np.array(m2_l), np.array(v_l, dtype=np.complex64)))
    sck_of = [0, 1, 2, 2, 3]
    sck = r.uniform(-0.2, 0.2, (4, 3))
    out = HERE / "udo_file_fixture.dat"
    BF.write_unf_dens_op(out, kp, grid, dN, ops, nb, B, b, sck_of, sck, e_fermi=-1.5)
    parsed = read_unf_dens_ops(filename=str(out))
    recs = []
    for per_k in parsed:
        for u in per_k:
            m = u.csr_matrix.tocoo()
            recs.append(dict(pckpt_number=u.pckpt_number, iener=u.iener, energy=u.energy, unfolded_N=u.unfolded_N,
                             nbands=u.nbands, folding_sckpt_number=u.folding_sckpt_number,
                             pckpt_coord_in_plot=u.pckpt_coord_in_plot, nener=u.nener_parent_grid,
                             emin=u.emin_parent_grid, emax=u.emax_parent_grid,
                             hermitian_rows=m.row.tolist(), hermitian_cols=m.col.tolist(),
                             hermitian_re=m.data.real.tolist(), hermitian_im=m.data.imag.tolist()))
    np.savez_compressed(HERE / "udo_file_inputs.npz", grid=grid, dN=dN, sck_of=np.array(sck_of), sck=sck,
                        **{f"op{i}_{n}": arr for i, o in enumerate(ops) for n, arr in zip(("ie", "m1", "m2", "v"), o)})
    (HERE / "udo_file_golden.json").write_text(json.dumps(
        {"source": "bandupy.unfolding_density_op.read_unf_dens_ops (reference) reading udo_file_fixture.dat",
         "records": recs}))
    return len(recs)


if __name__ == "__main__":
    print("udo file records parsed by the reference reader:", make_udo_file_golden())
    (HERE / "udo_golden.json").write_text(json.dumps(
        {"source": "bandupy.unfolding_density_op.UnfDensOp (reference, imported unmodified)",
         "cases": make_udo_cases()}))
    np.savez_compressed(HERE / "orb_contrib_golden.npz", **make_orb_contrib_cases())
    print("wrote", HERE / "udo_golden.json", HERE / "orb_contrib_golden.npz")
