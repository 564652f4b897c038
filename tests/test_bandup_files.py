"""BandUP's text files either side of the path (CPU): parsers follow the reference readers, the
writer produces the column format `bandup plot` loads."""
import re

import numpy as np
import pytest

from bandup_b200 import bandup_files as BF
from bandup_b200 import synth

KPTS = """K-points along the directions K1-Gamma, Gamma-M1, M1-K1
 25 25 13
Line-mode 0.000000000
Reciprocal
   0.333333333  0.666666667  0.000000000    1 ! K
   0.000000000  0.000000000  0.000000000    1 ! Gamma

   0.000000000  0.000000000  0.000000000    1 ! Gamma
   0.500000000  0.500000000  0.000000000    1 ! M

   0.500000000  0.500000000  0.000000000    1 ! M
   0.333333333  0.666666667  0.000000000    1 ! K
"""

LATT = """C ! Graphene primitive cell
   2.467000000
   0.866025403  -0.500000000   0.000000000
   0.866025403   0.500000000   0.000000000
   0.000000000   0.000000000   6.080259424 # trailing comment
  2
Selective Dynamics
Cartesian
   0.288675135   0.000000000   3.040129712  F  F  F
"""


def test_lattice_and_kpath(tmp_path):
    (tmp_path / "prim_cell_lattice.in").write_text(LATT)
    a = BF.read_lattice(tmp_path / "prim_cell_lattice.in")
    assert np.allclose(a, 2.467 * np.array([[0.866025403, -0.5, 0], [0.866025403, 0.5, 0], [0, 0, 6.080259424]]))
    b = synth.rec_latt(a)
    (tmp_path / "KPOINTS_prim_cell.in").write_text(KPTS)
    kp = BF.read_pckpts(tmp_path / "KPOINTS_prim_cell.in", b)
    assert [d.shape[0] for d in kp.directions] == [25, 25, 13] and kp.zero_of_kpts_scale == 0.0
    K = np.array([1 / 3, 2 / 3, 0]) @ b
    assert np.allclose(kp.directions[0][0], K, atol=1e-8) and np.allclose(kp.directions[0][-1], 0.0)
    assert np.allclose(kp.directions[2][-1], K, atol=1e-8)
    # kpts_line: equally spaced, end points included
    d = kp.directions[1]
    assert np.allclose(np.diff(d, axis=0), (d[-1] - d[0]) / 24)
    # one count for all directions; cartesian with a0
    txt = KPTS.replace(" 25 25 13", " 7").replace("Reciprocal", "Cartesian 2.467")
    (tmp_path / "k2").write_text(txt)
    kp2 = BF.read_pckpts(tmp_path / "k2", b)
    assert [d.shape[0] for d in kp2.directions] == [7, 7, 7]
    assert np.allclose(kp2.directions[0][0], BF.TWOPI * np.array([1 / 3, 2 / 3, 0]) / 2.467, atol=1e-8)
    (tmp_path / "k3").write_text(KPTS.replace("Reciprocal", "Cartesian"))
    with pytest.raises(ValueError):
        BF.read_pckpts(tmp_path / "k3", b)
    BF.write_lattice(tmp_path / "l2", a)
    assert np.allclose(BF.read_lattice(tmp_path / "l2"), a, rtol=1e-15)


def test_energy_info(tmp_path):
    p = tmp_path / "energy_info.in"
    p.write_text(" -2.4543  ! E-Fermi, in eV\n-25.0    ! Emin - E-Fermi\n7.0     ! Emax\n0.05     ! Increment\n")
    ef, e0, e1, de = BF.read_energy_info(p)
    assert (ef, de) == (-2.4543, 0.05) and e0 == -25.0 + ef and e1 == 7.0 + ef
    p.write_text("# comment\n1.0\n5.0\n-20.0\nauto\n")
    ef, e0, e1, de = BF.read_energy_info(p)
    assert e0 == -19.0 and e1 == 6.0 and abs(de - 0.004 * 25.0) < 1e-15


def test_band_struc_format(tmp_path):
    b = synth.rec_latt(synth.hex_cell(2.467, 15.0))
    (tmp_path / "k").write_text(KPTS.replace(" 25 25 13", " 3 2 2"))
    kp = BF.read_pckpts(tmp_path / "k", b)
    grid = np.array([-1.0, -0.95, -0.9])
    n = kp.all_kpts.shape[0]
    dN = np.arange(n * 3, dtype=float).reshape(n, 3) * 1.23456e-3
    BF.write_band_struc(tmp_path / "out.dat", kp, grid, dN, e_fermi=0.5)
    lines = (tmp_path / "out.dat").read_text().splitlines()
    assert lines[0].startswith("#") and lines[1].startswith("#KptCoord")
    body = [ln for ln in lines if not ln.startswith("#")]
    assert len(body) == n * 3
    assert re.fullmatch(r" {2}0\.0000 {2} -1\.5000 {3}0\.000E\+00", body[0])      # '(2(f8.4,2X),ES10.3)'
    assert body[4].endswith(" 4.938E-03")
    tab = BF.read_band_struc(tmp_path / "out.dat")
    assert tab.shape == (n * 3, 3)
    # k outer / E inner; the k coordinate keeps growing across directions
    assert np.allclose(tab[:3, 1], grid - 0.5) and np.all(np.diff(tab[::3, 0]) >= 0)
    seg1 = np.linalg.norm(kp.directions[0][-1] - kp.directions[0][0])
    assert abs(tab[3 * 3, 0] - seg1) < 1e-4                                      # first k of direction 2
    sig = np.ones((n, 3, 3)) * 0.25
    BF.write_band_struc(tmp_path / "spin.dat", kp, grid, dN, sigma=sig, spin_proj=np.zeros((n, 3, 2)))
    assert BF.read_band_struc(tmp_path / "spin.dat").shape == (n * 3, 8)


def test_readers_against_the_reference_sample_plot_files(tmp_path):
    """The reference ships one set of real inputs with an output it produced
    (utils/sample_input_files_task_plot/; committed verbatim in tests/golden/sample_plot_golden.json
    by make_sample_plot_golden.py).  Its K-G-M-K path on that lattice, sampled 25/25/13 times, must
    land on the 61 distinct KptCoord values of the reference-made file to its 4 decimals, and the
    energy column is real_seq(-22, 7, 0.05): 581 values."""
    import json
    from pathlib import Path
    g = json.loads((Path(__file__).parent / "golden" / "sample_plot_golden.json").read_text())
    for name in ("prim_cell_lattice.in", "energy_info.in"):
        (tmp_path / name).write_text(g[name])
    kp_lines = g["KPOINTS_prim_cell.in"].splitlines()
    assert kp_lines[1].strip() == "Auto"
    A = BF.read_lattice(tmp_path / "prim_cell_lattice.in")
    assert np.allclose(A[0], [2.467 * 0.866025403, -2.467 * 0.5, 0.0])
    b = BF.rec_latt(A) if hasattr(BF, "rec_latt") else 2 * np.pi * np.linalg.inv(A).T
    kp_lines[1] = "25 25 13"
    (tmp_path / "KPOINTS_prim_cell.in").write_text("\n".join(kp_lines) + "\n")
    kpath = BF.read_pckpts(tmp_path / "KPOINTS_prim_cell.in", b)
    assert [len(d) for d in kpath.directions] == [25, 25, 13]
    coords, first = [], kpath.zero_of_kpts_scale
    for d in kpath.directions:                                   # write_band_struc's accumulation
        c = [first + float(np.linalg.norm(k - d[0])) for k in d]
        coords += c
        first = c[-1]
    got = sorted(set(np.round(coords, 4).tolist()))
    assert len(got) == len(g["kpt_coords"]) == 61
    assert np.allclose(got, g["kpt_coords"], atol=1.01e-4)
    e_fermi, e_start, e_end, de = BF.read_energy_info(tmp_path / "energy_info.in")
    assert (e_fermi, de) == (-2.4543, 0.05) and abs(e_end - (7.0 + e_fermi)) < 1e-12
    from oracle import oracle as O
    lo, hi, n = g["energies_min_max_n"]
    grid = O.real_seq(lo, hi, de)
    assert len(grid) == n == 581 and abs(grid[-1] - hi) < 1e-9
    # the file as shipped asks for the automatic scan: min_dk = 2e-5 A^-1 steps (io_routines_mod.f90:1163-1169)
    (tmp_path / "auto.in").write_text(g["KPOINTS_prim_cell.in"])
    auto = BF.read_pckpts(tmp_path / "auto.in", b)
    lens = [np.linalg.norm(d[-1] - d[0]) for d in auto.directions]
    assert [len(d) for d in auto.directions] == [int(np.ceil(1.0 + L / 2e-5)) for L in lens]
