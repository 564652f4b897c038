"""BandUP_gpu.x -- the C++ host above the C ABI (bandup_b200/host/bandup_unfold_gpu.cpp): the
reference's `bandup unfold` flow for -no_symm_sckpts -no_symm_avg with the reference's own
command-line flags and files, no Python in the process."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from bandup_b200 import bandup_files as BF
from bandup_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "bandup_b200" / "lib" / "BandUP_gpu.x"


def _job(tmp_path, spinor):
    from bandup_b200.wavecar import write_wavecar
    a = synth.hex_cell(3.16, 18.0) if spinor else synth.hex_cell(2.467, 15.0)
    b = synth.rec_latt(a)
    segs = [([1 / 3, 1 / 3, 0.0], [0.0, 0.0, 0.0]), ([0.0, 0.0, 0.0], [0.5, 0.0, 0.0])]
    counts = [5, 4]
    kcart = np.concatenate([BF.kpts_line(np.array(s) @ b, np.array(e) @ b, n) for (s, e), n in zip(segs, counts)])
    sc = synth.Scenario("job", a, np.diag([4, 4, 1]), 28.0, 7, 2 if spinor else 1, kcart, e_fermi=-1.25,
                        emin=-12.0, emax=6.0, dE=0.1)
    data = [synth.make_wavefunction(sc, iK) for iK in range(sc.n_sc_kpts)]
    for d in data:
        d[2][2] = d[2][1]
    BF.write_lattice(tmp_path / "prim_cell_lattice.in", a, "pc")
    BF.write_lattice(tmp_path / "supercell_lattice.in", sc.A_sc, "SC")
    BF.write_pckpts(tmp_path / "KPOINTS_prim_cell.in", segs, counts)
    BF.write_energy_info(tmp_path / "energy_info.in", sc.e_fermi, sc.emin, sc.emax, sc.dE)
    write_wavecar(tmp_path / "WAVECAR", sc.A_sc, sc.ecut, list(sc.sc_kpts_frac), [d[0] for d in data],
                  [d[2] for d in data], extra_pad=1)
    return sc, data, counts


def test_host_exe_is_built_and_has_no_cpu_fallback(gpu_lib, tmp_path):
    import torch
    assert EXE.exists()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _job(tmp_path, False)
    r = subprocess.run([str(EXE), "-no_symm_sckpts", "-no_symm_avg"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and "Stopping now." in r.stderr
    r = subprocess.run([str(EXE), "-bogus"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "unknown command-line option" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("spinor", [False, True])
def test_host_exe_reproduces_the_python_path_and_the_oracle(gpu_lib, unfolder, tmp_path, spinor):
    from oracle import oracle as O
    sc, data, counts = _job(tmp_path, spinor)
    r = subprocess.run([str(EXE), "-no_symm_sckpts", "-no_symm_avg", "-write_unf_dens_op", "-wf_file", "WAVECAR",
                        "-out_file_nosymm", "exe_EBS.dat", "-output_file_only_user_selec_direcs_unf_dens_op", "exe_udo.dat"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "pc-kpts unfolded" in r.stdout
    # the same job through the Python mirror
    kpath, grid, dN = BF.unfold_task(unfolder, tmp_path / "WAVECAR", tmp_path / "prim_cell_lattice.in",
                                     tmp_path / "supercell_lattice.in", tmp_path / "KPOINTS_prim_cell.in",
                                     tmp_path / "energy_info.in", out_file=str(tmp_path / "py_EBS.dat"),
                                     unf_dens_op_file=str(tmp_path / "py_udo.dat"))
    exe_tab, py_tab = BF.read_band_struc(tmp_path / "exe_EBS.dat"), BF.read_band_struc(tmp_path / "py_EBS.dat")
    assert exe_tab.shape == py_tab.shape == (sum(counts) * len(grid), 8 if spinor else 3)
    assert np.array_equal(exe_tab, py_tab)                     # same library, same numbers, same format
    body = lambda p: [ln for ln in Path(p).read_text().splitlines() if not ln.startswith("# File created")]
    assert body(tmp_path / "exe_EBS.dat")[1:] == body(tmp_path / "py_EBS.dat")[1:]
    _, exe_recs = BF.read_unf_dens_op(tmp_path / "exe_udo.dat")
    _, py_recs = BF.read_unf_dens_op(tmp_path / "py_udo.dat")
    assert len(exe_recs) == len(py_recs) > 3
    for x, y in zip(exe_recs, py_recs):
        assert (x.pckpt_number, x.iener, x.folding_sckpt_number) == (y.pckpt_number, y.iener, y.folding_sckpt_number)
        assert np.array_equal(x.rows, y.rows) and np.array_equal(x.cols, y.cols) and np.array_equal(x.entries, y.entries)
    # SC-kpts taken from a counter file (-ticket_file, the multi-rank load balancer): same output
    r = subprocess.run([str(EXE), "-no_symm_sckpts", "-no_symm_avg", "-wf_file", "WAVECAR", "-out_file_nosymm",
                        "dyn_EBS.dat", "-ticket_file", "tickets.bin"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert np.array_equal(BF.read_band_struc(tmp_path / "dyn_EBS.dat"), exe_tab)
    assert int(np.fromfile(tmp_path / "tickets.bin", dtype=np.int64)[0]) == sc.n_sc_kpts + 1   # every k-point once + the take that found none left
    # and against the oracle (to the file's 4 digits)
    for iK in range(sc.n_sc_kpts):
        coeffs, gf, ener = data[iK]
        _, d, _, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid, norm_mode=1)
        for f, row in enumerate(sc.folds[iK]):
            got = exe_tab[row * len(grid):(row + 1) * len(grid), 2]
            assert np.allclose(got, d[f], rtol=6e-4, atol=1e-12)


@pytest.mark.gpu
def test_second_spin_channel_of_a_spin_polarised_wavecar(gpu_lib, unfolder, tmp_path):
    """nspin = 2 with different coefficients and energies per channel: -spin_channel 2 (host exe)
    and spin_channel=2 (Python mirror) unfold the second channel -- equal to the oracle on that
    channel's data and different from channel 1."""
    from oracle import oracle as O
    from bandup_b200.wavecar import write_wavecar
    sc, data, counts = _job(tmp_path, False)
    rng = np.random.default_rng(5)
    data2 = []
    for c, gf, e in data:
        c2 = (c * (1.0 + 0.5 * rng.standard_normal(c.shape))).astype(np.complex64)
        data2.append((c2, gf, np.sort(e + rng.uniform(-0.4, 0.4, e.shape))))
    write_wavecar(tmp_path / "WAVECAR", sc.A_sc, sc.ecut, list(sc.sc_kpts_frac), [d[0] for d in data],
                  [d[2] for d in data], nspin=2, coeffs_spin2=[d[0] for d in data2],
                  energies_spin2=[d[2] for d in data2])
    tabs = {}
    for ch in (1, 2):
        r = subprocess.run([str(EXE), "-no_symm_sckpts", "-no_symm_avg", "-spin_channel", str(ch), "-out_file_nosymm",
                            f"exe{ch}.dat"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        tabs[ch] = BF.read_band_struc(tmp_path / f"exe{ch}.dat")
    assert not np.array_equal(tabs[1][:, 2], tabs[2][:, 2])
    kpath, grid, dN = BF.unfold_task(unfolder, tmp_path / "WAVECAR", tmp_path / "prim_cell_lattice.in",
                                     tmp_path / "supercell_lattice.in", tmp_path / "KPOINTS_prim_cell.in",
                                     tmp_path / "energy_info.in", out_file=str(tmp_path / "py2.dat"), spin_channel=2)
    assert np.array_equal(BF.read_band_struc(tmp_path / "py2.dat"), tabs[2])
    for ch, src in ((1, data), (2, data2)):
        for iK in range(sc.n_sc_kpts):
            coeffs, gf, ener = src[iK]
            _, d, _, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, sc.folding_G(iK), grid, norm_mode=1)
            for f, row in enumerate(sc.folds[iK]):
                got = tabs[ch][row * len(grid):(row + 1) * len(grid), 2]
                assert np.allclose(got, d[f], rtol=6e-4, atol=1e-12)


@pytest.mark.gpu
def test_exe_flags_skip_sckpts_and_dont_unfold(gpu_lib, unfolder, tmp_path):
    """-n_sckpts_to_skip: a hybrid-functional style WAVECAR whose first k-point repeats a later one
    with meaningless coefficients is unfolded from the later copy (cla_wrappers_mod.f90:258-268,
    band_unfolding_routines_mod.f90:311-315).  -dont_unfold: P = 1, so delta_N counts the SC bands
    of every bin (:1346-1351)."""
    from bandup_b200.wavecar import write_wavecar
    sc, data, counts = _job(tmp_path, False)
    base = [str(EXE), "-no_symm_sckpts", "-no_symm_avg"]
    r = subprocess.run(base + ["-out_file_nosymm", "ref.dat"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref = BF.read_band_struc(tmp_path / "ref.dat")
    rng = np.random.default_rng(9)
    junk = (rng.standard_normal(data[0][0].shape) + 1j * rng.standard_normal(data[0][0].shape)).astype(np.complex64)
    write_wavecar(tmp_path / "WAVECAR_hyb", sc.A_sc, sc.ecut, [sc.sc_kpts_frac[0]] + list(sc.sc_kpts_frac),
                  [junk] + [d[0] for d in data], [data[0][2]] + [d[2] for d in data], extra_pad=1)
    r = subprocess.run(base + ["-wf_file", "WAVECAR_hyb", "-n_sckpts_to_skip", "1", "-out_file_nosymm", "skip.dat"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(BF.read_band_struc(tmp_path / "skip.dat"), ref)
    r = subprocess.run(base + ["-wf_file", "WAVECAR_hyb", "-out_file_nosymm", "noskip.dat"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0 and not np.array_equal(BF.read_band_struc(tmp_path / "noskip.dat"), ref)
    _, grid, dN = BF.unfold_task(unfolder, tmp_path / "WAVECAR_hyb", tmp_path / "prim_cell_lattice.in",
                                 tmp_path / "supercell_lattice.in", tmp_path / "KPOINTS_prim_cell.in",
                                 tmp_path / "energy_info.in", out_file=str(tmp_path / "py_skip.dat"), n_sckpts_to_skip=1)
    assert np.array_equal(BF.read_band_struc(tmp_path / "py_skip.dat"), ref)
    # -dont_unfold
    r = subprocess.run(base + ["-dont_unfold", "-out_file_nosymm", "nounf.dat"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tab = BF.read_band_struc(tmp_path / "nounf.dat")
    de = grid[1] - grid[0]
    for iK in range(sc.n_sc_kpts):
        ener = data[iK][2]
        want = np.array([np.sum(np.abs(ener - E) < 0.5 * de - 1e-9) for E in grid], dtype=float)
        edge = np.array([np.any(np.abs(np.abs(ener - E) - 0.5 * de) <= 1e-9) for E in grid])
        for row in sc.folds[iK]:
            got = tab[row * len(grid):(row + 1) * len(grid), 2]
            assert np.allclose(got[~edge], want[~edge], atol=1e-9)


@pytest.mark.gpu
def test_exe_spin_projection_flags(gpu_lib, unfolder, tmp_path):
    """-normal_to_proj_plane / -origin_for_spin_proj_cart / -origin_for_spin_proj_rec change the two
    in-plane spin projection columns (calc_spin_projections, band_unfolding_routines_mod.f90:1266-1286)
    and nothing else; host exe and Python mirror agree, and the fractional origin wins."""
    sc, data, counts = _job(tmp_path, True)
    base = [str(EXE), "-no_symm_sckpts", "-no_symm_avg"]
    files = [tmp_path / n for n in ("WAVECAR", "prim_cell_lattice.in", "supercell_lattice.in", "KPOINTS_prim_cell.in",
                                    "energy_info.in")]
    r = subprocess.run(base + ["-out_file_nosymm", "def.dat"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    default = BF.read_band_struc(tmp_path / "def.dat")
    rec = np.array([0.25, -0.5, 0.0])
    cart = rec @ sc.B_sc
    r = subprocess.run(base + ["-out_file_nosymm", "a.dat", "-normal_to_proj_plane", "1, 0, 1", "-origin_for_spin_proj_cart",
                               "%.12f, %.12f, %.12f" % tuple(cart)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a = BF.read_band_struc(tmp_path / "a.dat")
    assert np.array_equal(a[:, :6], default[:, :6]) and not np.array_equal(a[:, 6:], default[:, 6:])
    BF.unfold_task(unfolder, *files, out_file=str(tmp_path / "py_a.dat"), normal_to_proj_plane=(1.0, 0.0, 1.0),
                   origin_for_spin_proj=cart)
    assert np.array_equal(BF.read_band_struc(tmp_path / "py_a.dat"), a)
    assert '_|_ to [ 1.000E+00,  0.000E+00,  1.000E+00].' in (tmp_path / "a.dat").read_text()
    r = subprocess.run(base + ["-out_file_nosymm", "b.dat", "-normal_to_proj_plane", "1, 0, 1", "-origin_for_spin_proj_rec",
                               "0.25, -0.5, 0", "-origin_for_spin_proj_cart", "9, 9, 9"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert np.array_equal(BF.read_band_struc(tmp_path / "b.dat"), a)
    BF.unfold_task(unfolder, *files, out_file=str(tmp_path / "py_b.dat"), normal_to_proj_plane=(1.0, 0.0, 1.0),
                   origin_for_spin_proj=(9.0, 9.0, 9.0), origin_for_spin_proj_rec=rec)
    assert np.array_equal(BF.read_band_struc(tmp_path / "py_b.dat"), a)
