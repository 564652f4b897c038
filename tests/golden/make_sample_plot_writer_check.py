#!/usr/bin/env python
"""The reference's own writer, EXECUTED (oracle/f90interp.py), against a file a compiled BandUP wrote.

utils/sample_input_files_task_plot/unfolded_EBS_symmetry-averaged.dat is the only output of a real
(compiled) BandUP run in the reference tree: 63 pc-kpts on K-G-M-K (25/25/13 per segment) x 581
energies.  Here `real_seq` and `write_band_struc` (lists_and_seqs_mod.f90:188-200,
io_routines_mod.f90:889-977) are run from the reference sources on that path, that energy window and
the file's own delta_N column, and every data line the interpreter prints is compared with the line
the compiled program printed: KptCoord (the per-direction accumulation and `norm`), E - E_Fermi and
the f8.4 / ES10.3 edit descriptors.  (The shipped file is from a BandUP version that looped over the
energies first; lines are matched by (k-point, energy), text for text.)

What it pins: the interpreter's arithmetic and formatted output on this routine against a compiled
build -- and with it the golden `unfolded_EBS` files of tests/golden/fortran_golden.npz, which the
same interpreted routine wrote.  Writes tests/golden/sample_plot_writer_check.json (counts and
SHA-256 of both line sets); tests/test_fortran_golden.py repeats the run where /root/reference exists.

    python tests/golden/make_sample_plot_writer_check.py
"""
import hashlib
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from oracle import f90interp as F  # noqa: E402
import make_fortran_golden as G  # noqa: E402

SAMPLE = Path("/root/reference/utils/sample_input_files_task_plot")
R8, I4 = np.float64, np.int32


def run():
    from bandup_b200 import bandup_files as BF
    gold = json.loads((Path(__file__).with_name("sample_plot_golden.json")).read_text())
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        (tmp / "prim_cell_lattice.in").write_text(gold["prim_cell_lattice.in"])
        kp = gold["KPOINTS_prim_cell.in"].splitlines()
        kp[1] = "25 25 13"               # the shipped file says "Auto"; its output has 25/25/13 points per segment
        (tmp / "KPOINTS_prim_cell.in").write_text("\n".join(kp) + "\n")
        A = BF.read_lattice(tmp / "prim_cell_lattice.in")
        it = G.interpreter()
        b = F.Var(np.zeros((3, 3)), None)
        it.call("get_rec_latt", np.asarray(A, dtype=R8), b)          # crystals_mod.f90:186-202
        kpath = BF.read_pckpts(tmp / "KPOINTS_prim_cell.in", b.value)
    dirs = [np.asarray(d, dtype=R8) for d in kpath.directions]
    n_k = sum(len(d) for d in dirs)
    ref_lines = (SAMPLE / "unfolded_EBS_symmetry-averaged.dat").read_text().splitlines()
    assert ref_lines[0].startswith("#KptCoord") and (len(ref_lines) - 1) % n_k == 0
    data = ref_lines[1:]
    nener = len(data) // n_k
    e_fermi = R8(-2.4543)
    de = R8(0.05)
    lo, hi = float(data[0].split()[1]), float(data[-1].split()[1])
    egrid = F.Var(None, G.decl(it, "real(kind=dp), dimension(:), allocatable :: return_list"))
    it.call("real_seq", R8(lo) + e_fermi, R8(hi) + e_fermi, de, egrid)
    egrid = egrid.value
    assert len(egrid) == nener
    dN = np.array([float(l.split()[2]) for l in data], dtype=R8).reshape(nener, n_k)     # energy-major in the file

    pck = it.new_struct("selected_pcbz_directions")
    pck["selec_pcbz_dir"] = G.struct_array(it, "needed_pcbz_dirs_for_ebs", len(dirs))
    out = it.new_struct("unfoldedquantitiesforoutput")
    out["pcbz_dir"] = G.struct_array(it, "list_of_pckpts_for_unfolded_quantities", len(dirs))
    k0 = 0
    for i, d in enumerate(dirs):
        pck["selec_pcbz_dir"][i]["needed_dir"] = G.struct_array(it, "list_of_trial_folding_pckpts", 1)
        pts = G.struct_array(it, "trial_folding_pckpt", len(d))
        uq = G.struct_array(it, "unfolded_quantities_for_given_pckpt", len(d))
        for j in range(len(d)):
            pts[j]["coords"][:] = d[j]
            uq[j]["dn"] = dN[:, k0 + j].copy()
        pck["selec_pcbz_dir"][i]["needed_dir"][0]["pckpt"] = pts
        out["pcbz_dir"][i]["pckpt"] = uq
        k0 += len(d)
    it.overrides["file_header_bandup"] = lambda: "# header"
    with tempfile.TemporaryDirectory() as tmp:
        f = Path(tmp) / "ebs.dat"
        it.call("write_band_struc", str(f), pck, egrid, out, ef=e_fermi, zero_of_kpts_scale=R8(kpath.zero_of_kpts_scale))
        got = f.read_text().splitlines()
    assert got[0] == "# header" and got[1].rstrip() == ref_lines[0].rstrip()
    mine = np.array(got[2:], dtype=object).reshape(n_k, nener).T.reshape(-1).tolist()     # k-major -> energy-major
    return mine, data, it.n_statements if hasattr(it, "n_statements") else None


def digest(lines):
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


def product_writer_lines(dN_energy_major):
    """The same file from the PRODUCT's writer (bandup_b200.bandup_files.write_band_struc, what
    unfold_task and BandUP_gpu.x print), data lines in the sample file's energy-major order.
    Needs only committed fixtures (no /root/reference)."""
    from bandup_b200 import bandup_files as BF, synth
    from oracle import oracle as O          # real_seq: the energy grid the library hands the product's writer
    gold = json.loads((Path(__file__).with_name("sample_plot_golden.json")).read_text())
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        (tmp / "prim_cell_lattice.in").write_text(gold["prim_cell_lattice.in"])
        kp = gold["KPOINTS_prim_cell.in"].splitlines()
        kp[1] = "25 25 13"
        (tmp / "KPOINTS_prim_cell.in").write_text("\n".join(kp) + "\n")
        A = BF.read_lattice(tmp / "prim_cell_lattice.in")
        kpath = BF.read_pckpts(tmp / "KPOINTS_prim_cell.in", synth.rec_latt(A))      # what unfold_task does
        nener, n_k = dN_energy_major.shape
        e_fermi = -2.4543
        lo, hi, _ = gold["energies_min_max_n"]
        egrid = O.real_seq(lo + e_fermi, hi + e_fermi, 0.05)
        assert len(egrid) == nener and sum(len(d) for d in kpath.directions) == n_k
        f = tmp / "ebs.dat"
        BF.write_band_struc(f, kpath, egrid, np.ascontiguousarray(dN_energy_major.T), e_fermi=e_fermi)
        got = f.read_text().splitlines()
    assert got[1].rstrip() == "#KptCoord #E-E_Fermi #delta_N"
    return np.array(got[2:], dtype=object).reshape(n_k, nener).T.reshape(-1).tolist()


if __name__ == "__main__":
    mine, ref, n_stmt = run()
    bad = [(i, a, b) for i, (a, b) in enumerate(zip(mine, ref)) if a != b]
    print(len(mine), "lines,", len(bad), "differ", bad[:5])
    # the file's delta_N column as a small fixture (sparse: 10 565 of 36 603 values are non-zero), so
    # that the product's writer can be held against the compiled program's file wherever the tests run
    d = np.array([float(l.split()[2]) for l in ref])
    idx = np.nonzero(d)[0].astype(np.int32)
    np.savez_compressed(Path(__file__).with_name("sample_plot_dN.npz"), idx=idx, val=d[idx].astype(np.float32),
                        shape=np.array([len(ref) // 63, 63], dtype=np.int32))
    dd = np.zeros(len(d))
    dd[idx] = d[idx].astype(np.float32).astype(np.float64)
    prod = product_writer_lines(dd.reshape(len(ref) // 63, 63))
    bad_p = [(i, a, b) for i, (a, b) in enumerate(zip(prod, ref)) if a != b]
    print("product writer:", len(prod), "lines,", len(bad_p), "differ", bad_p[:5])
    bad = bad + bad_p
    rec = {"data_lines": len(ref), "differing_lines": len(bad), "sha256_reference_file_lines": digest(ref),
           "sha256_interpreted_writer_lines": digest(mine), "sha256_product_writer_lines": digest(prod),
           "fortran_statements": n_stmt,
           "what": "write_band_struc + real_seq executed from the reference sources vs the compiled BandUP's own "
                   "sample output (utils/sample_input_files_task_plot), line for line"}
    Path(__file__).with_name("sample_plot_writer_check.json").write_text(json.dumps(rec, indent=1) + "\n")
    sys.exit(1 if bad else 0)
