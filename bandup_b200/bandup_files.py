"""BandUP's text input/output files on either side of the unfolding path -- just enough of the
reference's formats for a file-level drop-in of `bandup unfold` under its no-symmetry flags
(-no_symm_sckpts -no_symm_avg): the lattices, the pc k-point path, the energy window go in, the
`unfolded_EBS_*.dat` table comes out in the column format `bandup plot` reads with np.loadtxt
(python_interface/bandupy/plot.py:143-175).  Nothing here computes weights.

  prim_cell_lattice.in / supercell_lattice.in   POSCAR-style (read_vasp_files_mod.f90:63-200):
        comment / scaling factor / three lattice vectors; atoms, if any, are ignored here
  KPOINTS_prim_cell.in    VASP line-mode style (io_routines_mod.f90:981-1211): title / numbers of
        k-points per direction / 'Line-mode [zero_of_kpts_scale]' / 'Reciprocal' or
        'Cartesian a0' / pairs of start-end lines separated by blank lines
  energy_info.in          E_Fermi / Emin-E_F / Emax-E_F / dE or 'auto' (io_routines_mod.f90:334-388)
  unfolded_EBS_*.dat      '(2(f8.4,2X),ES10.3)' k-coordinate, E-E_F, delta_N; k outer, E inner,
        directions concatenated (write_band_struc, io_routines_mod.f90:889-977); with spin info
        five more ES10.3 columns (sigma_x, sigma_y, sigma_z, spin perp, spin para)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import synth

TWOPI = 2.0 * (4.0 * math.atan(1.0))


def read_lattice(path) -> np.ndarray:
    """Rows = lattice vectors in Angstrom (scaling factor applied)."""
    lines = Path(path).read_text().splitlines()
    scale = float(lines[1].split()[0])
    vecs = np.array([[float(x) for x in lines[2 + i].split("#")[0].split("!")[0].split()[:3]] for i in range(3)])
    return scale * vecs


def write_lattice(path, latt: np.ndarray, comment: str = "cell") -> None:
    a = np.asarray(latt, dtype=np.float64)
    rows = "\n".join("  %22.16f %22.16f %22.16f" % tuple(v) for v in a)
    Path(path).write_text(f"{comment}\n   1.0\n{rows}\n")


def kpts_line(kstart, kend, nk: int) -> np.ndarray:
    """lists_and_seqs_mod.f90:219-241: nk points from kstart to kend inclusive."""
    kstart, kend = np.asarray(kstart, dtype=np.float64), np.asarray(kend, dtype=np.float64)
    step = (kend - kstart) / float(nk - 1) if nk > 1 else np.zeros(3)
    return np.array([kstart + float(i) * step for i in range(nk)])


@dataclass
class PcKptsPath:
    directions: List[np.ndarray]      # per direction: (n_k, 3) Cartesian pc k-points
    zero_of_kpts_scale: float

    @property
    def all_kpts(self) -> np.ndarray:
        return np.concatenate(self.directions, axis=0)


MIN_DK = 2.0 * 1e-5     # min_dk (constants_and_types_mod.f90:65)


def read_pckpts(path, b_matrix_pc: np.ndarray) -> PcKptsPath:
    """read_pckpts_selected_by_user + kpts_line for every direction
    (define_pckpts_to_be_checked, band_unfolding_routines_mod.f90:486-547)."""
    lines = Path(path).read_text().splitlines()
    if len(lines) < 6:
        raise ValueError("ERROR reading input pc-kpts file. Please check the format.")
    n_line, mode_line, coord_line = lines[1].strip(), lines[2].strip(), lines[3].strip()
    try:
        counts = [int(x) for x in n_line.split()]
        if not counts:
            raise ValueError
    except ValueError:
        # no number on the second line: the reference scans the directions automatically in steps of
        # min_dk = 2 * default_tol_for_vec_equality = 2e-5 A^-1 (io_routines_mod.f90:1131-1170,
        # constants_and_types_mod.f90:65); with stop_if_GUR_fails = .TRUE. it then stops unless every
        # scanned point folds -- here that is the fold map's FOLDING_VECTOR error
        counts = None
    toks = mode_line.split()
    zero = 0.0
    if len(toks) > 1:
        try:
            zero = float(toks[1])
        except ValueError:
            zero = 0.0
    ctoks = coord_line.split()
    cartesian = ctoks[0][:1].upper() == "C"
    a0 = None
    if len(ctoks) > 1:
        try:
            a0 = float(ctoks[1])
        except ValueError:
            a0 = None
    if cartesian and a0 is None:
        try:
            a0 = float(lines[0].split()[0])           # deprecated: a0 on the first line
        except (ValueError, IndexError):
            raise ValueError('cartesian coordinates need the scaling parameter "a0" after the tag')
    pairs: List[Tuple[np.ndarray, np.ndarray]] = []
    body = [ln.split("!")[0].strip() for ln in lines[4:]]
    block: List[np.ndarray] = []
    for ln in body + [""]:
        if ln == "":
            if len(block) == 2:
                pairs.append((block[0], block[1]))
            elif len(block) != 0:
                raise ValueError("ERROR reading input pc-kpts file. Please check the format.")
            block = []
        else:
            block.append(np.array([float(x) for x in ln.split()[:3]]))
    if not pairs:
        raise ValueError("no pc-kpt direction found")
    if counts is not None and len(counts) < len(pairs):
        counts = [counts[0]] * len(pairs)
    b = np.asarray(b_matrix_pc, dtype=np.float64)
    dirs = []
    for idir, (ks, ke) in enumerate(pairs):
        if cartesian:
            ks_c, ke_c = TWOPI * ks / a0, TWOPI * ke / a0
        else:
            ks_c, ke_c = ks @ b, ke @ b
        nk = counts[idir] if counts is not None else int(np.ceil(1.0 + np.linalg.norm(ke_c - ks_c) / MIN_DK))
        if np.linalg.norm(ks_c - ke_c) < np.finfo(float).eps:
            nk = 1
        dirs.append(kpts_line(ks_c, ke_c, nk))
    return PcKptsPath(dirs, zero)


def write_pckpts(path, segments_frac: Sequence[Tuple[Sequence[float], Sequence[float]]], counts: Sequence[int],
                 title: str = "pc k-points", zero: float = 0.0) -> None:
    out = [title, " " + " ".join(str(int(c)) for c in counts), f"Line-mode {zero:.9f}", "Reciprocal"]
    for i, (a, b) in enumerate(segments_frac):
        out.append("   %.12f  %.12f  %.12f    1" % tuple(a))
        out.append("   %.12f  %.12f  %.12f    1" % tuple(b))
        if i + 1 < len(segments_frac):
            out.append("")
    Path(path).write_text("\n".join(out) + "\n")


def read_energy_info(path) -> Tuple[float, float, float, float]:
    """(e_fermi, E_start, E_end, delta_e) as read_energy_info_for_band_search builds them."""
    vals = []
    for ln in Path(path).read_text().splitlines():
        t = ln.strip()
        if not t or t.startswith("#"):
            continue
        vals.append(t.split("!")[0].split()[0])
        if len(vals) == 4:
            break
    if len(vals) < 4:
        raise ValueError("energy file needs E_Fermi, Emin-E_F, Emax-E_F and dE")
    e_fermi = float(vals[0])
    e_start, e_end = float(vals[1]) + e_fermi, float(vals[2]) + e_fermi
    if e_end < e_start:
        e_start, e_end = e_end, e_start
    try:
        de = abs(float(vals[3]))
    except ValueError:
        de = abs(0.4e-2 * (e_end - e_start))          # 'auto': 0.4 % of the window
    return e_fermi, e_start, e_end, de


def write_energy_info(path, e_fermi: float, emin: float, emax: float, dE) -> None:
    Path(path).write_text(f" {e_fermi}  ! E-Fermi, in eV\n{emin}    ! Emin - E-Fermi, in eV\n"
                          f"{emax}     ! Emax - E-Fermi, in eV\n{dE}     ! Increment, in eV\n")


def _es(x: float) -> str:
    """Fortran ES10.3: d.dddE+ee; a three-digit exponent drops the letter (d.ddd+eee) as the
    standard prescribes for Ew.d, and what does not fit the ten columns becomes asterisks."""
    s = "%.3E" % x
    mant, exp = s.split("E")
    if abs(int(exp)) > 99:
        s = "%s%+04d" % (mant, int(exp))
    return s.rjust(10) if len(s) <= 10 else "*" * 10


def write_band_struc(path, kpath: PcKptsPath, energy_grid: np.ndarray, dN: np.ndarray, e_fermi: float = 0.0,
                     sigma: Optional[np.ndarray] = None, spin_proj: Optional[np.ndarray] = None,
                     header: str = "# File created by bandup-b200 (GPU unfolding path)",
                     saxis: Sequence[float] = (0.0, 0.0, 1.0),
                     normal_to_proj_plane: Sequence[float] = (0.0, 0.0, 1.0)) -> None:
    """write_band_struc (io_routines_mod.f90:889-977).  dN: (n_pckpts_total, nener) in the order of
    kpath.all_kpts; sigma (n, nener, 3) and spin_proj (n, nener, 2) add the five spin columns."""
    grid = np.asarray(energy_grid, dtype=np.float64)
    with open(path, "w") as f:
        f.write(header.rstrip("\n") + "\n")
        if sigma is not None:
            # the three comment lines of the reference's spinor output (:921-929)
            f.write("# NB.: Unfolding of Pauli matrices eigenvs is a test feature.\n")
            f.write("# The quantization axis assumed was saxis = [%s, %s, %s]\n" % tuple(_es(v) for v in saxis))
            f.write('# The projection axis used for "#Spin _|_ k" is also _|_ to [%s, %s, %s].\n'
                    % tuple(_es(v) for v in normal_to_proj_plane))
            # '(3(A,1X), 2X, 5(A,1X))' (:930-933): the last label carries its own blank and is followed by 1X
            f.write("#KptCoord #E-E_Fermi #delta_N   #sigma_x    #sigma_y    #sigma_z    #Spin _|_ k #Spin // k  \n")
        else:
            f.write("#KptCoord #E-E_Fermi #delta_N \n")
        coord_first = kpath.zero_of_kpts_scale
        row = 0
        coord_k = coord_first
        for d in kpath.directions:
            first = d[0]
            for k in d:
                coord_k = coord_first + float(np.linalg.norm(k - first))
                for ie, E in enumerate(grid):
                    line = "%8.4f  %8.4f  %s" % (coord_k, E - e_fermi, _es(dN[row, ie]))
                    if sigma is not None:
                        cols = list(sigma[row, ie]) + (list(spin_proj[row, ie]) if spin_proj is not None else [0.0, 0.0])
                        line += "  " + "  ".join(_es(c) for c in cols) + "  "
                    f.write(line + "\n")
                row += 1
            coord_first = coord_k


def read_band_struc(path) -> np.ndarray:
    """The numeric columns, the way `bandup plot` loads them (plot.py:143-175: np.loadtxt)."""
    return np.loadtxt(path, comments="#")


SPDF = ['s', 'py', 'pz', 'px', 'dxy', 'dyz', 'dz2', 'dxz', 'dx2',
        'fy(3x2-y2)', 'fxyz', 'fyz2', 'fz3', 'fxz2', 'fz(x2-y2)', 'fx(x2-3y2)']


def formatted_orb_choice(orb_choices, supported: Sequence[str]) -> List[str]:
    """bandupy/orbital_contributions.py:40-80: 'all', 's'/'p'/'d'/'f', 'sp2', 'sp3', ... or explicit
    orbital names -> the picked orbitals in PROCAR order."""
    if orb_choices is None:
        orb_choices = 'all'
    if isinstance(orb_choices, (list, tuple)):
        picked = set()
        for o in orb_choices:
            picked.update(formatted_orb_choice(o, supported))
        return [o for o in SPDF if o in picked]
    table = {o: [o] for o in supported}
    table['all'] = list(supported)
    for general in ('s', 'p', 'd', 'f'):
        table[general] = [o for o in supported if general in o.lower()]
    table['sp2'] = table['spxy'] = table['spyx'] = ['s', 'px', 'py']
    table['spxz'] = table['spzx'] = ['s', 'px', 'pz']
    table['spyz'] = table['spzy'] = ['s', 'py', 'pz']
    table['sp3'] = table['sp2'] + ['pz']
    key = str(orb_choices).lower().strip()
    if key not in table:
        raise ValueError(f'ERROR: Invalid choice of orbital: "{orb_choices}".')
    return table[key]


def ao_mask(n_ions: int, orbitals: Sequence[str], picked_orbitals, selected_ion_indices=None) -> np.ndarray:
    """uint8 mask over the combined atom/orbital index a = iat*n_orbs_per_atom + iorb
    (KptInfo.combined_atom_orb_index, orbital_contributions.py:195-197)."""
    picked = formatted_orb_choice(picked_orbitals, orbitals)
    ions = range(n_ions) if selected_ion_indices is None else selected_ion_indices
    mask = np.zeros(n_ions * len(orbitals), dtype=np.uint8)
    for iat in ions:
        for orb in picked:
            mask[iat * len(orbitals) + list(orbitals).index(orb)] = 1
    return mask


def write_orb_dependent_ebs(path, kpath: PcKptsPath, energy_grid: np.ndarray, values: np.ndarray, e_fermi: float,
                            n_atoms_sc: int, picked_atoms: str, picked_orbitals: Sequence[str],
                            spin_channel: int = 1) -> None:
    """unfolded_EBS_orb_dependent.dat (orbital_contributions.py:473-488): '%.4f  %.4f   %.4f' lines,
    k outer / E inner, k coordinate as in the band-structure file, energies relative to E_F."""
    grid = np.asarray(energy_grid, dtype=np.float64)
    with open(path, "w") as f:
        f.write("# File created by bandup-b200 (GPU unfolding path)\n"
                "# Orbital/atom-decomposed unfolded band structure\n"
                f"# nAtomsSC: {n_atoms_sc}\n# PickedAtomIndicesSC: {picked_atoms}\n"
                f"# OrbitalProjector: {'+'.join(picked_orbitals)}\n# SpinChannel: {spin_channel}\n"
                "#KCoord #E-E_Fermi #{N*Unf[<Proj>]}(k,E)\n")
        coord_first = kpath.zero_of_kpts_scale
        coord_k = coord_first
        row = 0
        for d in kpath.directions:
            first = d[0]
            for k in d:
                coord_k = coord_first + float(np.linalg.norm(k - first))
                for ie, E in enumerate(grid):
                    f.write("%.4f  %.4f   %.4f\n" % (coord_k, E - e_fermi, values[row, ie]))
                row += 1
            coord_first = coord_k


def unfold_task(unfolder, wavecar_path, prim_cell_file, supercell_file, pckpts_file, energy_file,
                out_file: Optional[str] = None, spin_channel: int = 1, unf_dens_op_file: Optional[str] = None,
                normal_to_proj_plane: Sequence[float] = (0.0, 0.0, 1.0), origin_for_spin_proj=None,
                projections=None, orb_out_file: Optional[str] = None, n_sckpts_to_skip: int = 0,
                origin_for_spin_proj_rec=None):
    """`bandup unfold` at file level for -no_symm_sckpts -no_symm_avg: parse the reference's input
    files, run every SC-kpt of the WAVECAR through the GPU (wavecar.unfold_wavecar) and write
    unfolded_EBS_not-symmetry_averaged.dat -- for a spinor WAVECAR with the reference's five spin
    columns (sigma_x, sigma_y, sigma_z, spin perp / para to k; io_routines_mod.f90:950-960).
    unf_dens_op_file: the reference's -write_unf_dens_op (`bandup unfold --orbitals`), the
        unfolding-density operators in the text format `bandup projected-unfold` reads.
    projections + orb_out_file: `bandup projected-unfold` without the text round trip.
        projections = dict(load=callable(sc_kpt_number 1-based) -> (proj_matrix (n_ao, n_bands),
        proj_matrix_dual (n_bands, n_ao)) -- the 'direct'/'dual' arrays of the reference's
        orbital_projections/proj_matrices_iKpt_*_iSpin_*.npz --, n_ions, orbitals, picked_orbitals,
        selected_ion_indices); writes unfolded_EBS_orb_dependent.dat.
    Returns (kpath, energy_grid, dN)."""
    from .unfold import calc_spin_projections
    from .wavecar import Wavecar, unfold_wavecar
    a_pc = read_lattice(prim_cell_file)
    A_sc = read_lattice(supercell_file)
    b_pc = synth.rec_latt(a_pc)
    wc = Wavecar(wavecar_path)
    # the SC of the wavefunction file must be the one of supercell_lattice.in (main_BandUP.f90:57-70)
    if not np.allclose(A_sc, wc.A_matrix, rtol=0, atol=1e-4 * np.abs(A_sc).max()):
        raise ValueError("supercell_lattice.in and the WAVECAR disagree about the supercell")
    kpath = read_pckpts(pckpts_file, b_pc)
    all_k = kpath.all_kpts
    e_fermi, e_start, e_end, de = read_energy_info(energy_file)
    rho_by_row, sigma_by_row, orb_by_row = {}, {}, {}
    want_rho = bool(unf_dens_op_file)
    want_orb = projections is not None and bool(orb_out_file)
    file_kpt_of_job = {}

    def after_fetch(iK, res, row_ids):
        need_spin = res.n_spinor == 2
        if need_spin:
            spinor_kpts.add(iK)
        if not (want_rho or want_orb or need_spin):
            return
        ops = unfolder.calc_rho(iK)
        n_fold = len(row_ids)
        if want_rho:
            for f, row in enumerate(row_ids):
                s = ops.fold == f
                rho_by_row[int(row)] = (ops.iener[s], ops.m1[s], ops.m2[s], ops.rho[s])
        if need_spin:
            pv = unfolder.pauli_vectors(iK, n_fold)
            for f, row in enumerate(row_ids):
                sigma_by_row[int(row)] = pv[f]
        if want_orb:
            proj, dual = projections["load"](file_kpt_of_job[iK])
            mask = ao_mask(projections["n_ions"], projections["orbitals"], projections.get("picked_orbitals", "all"),
                           projections.get("selected_ion_indices"))
            unfolder.get_orbital_contribution_matrix(dual, proj, mask, energies_of_job[iK])
            vals = unfolder.unfold_operator(iK, n_fold, min_abs_rho=1e-4, clip=True)
            for f, row in enumerate(row_ids):
                orb_by_row[int(row)] = vals[f]

    # band energies and file k-point numbers of the job's k-points: the headers say (scalar vs
    # spinor is decided on the device when the k-point is submitted)
    spinor_kpts, energies_of_job = set(), {}
    sc_k = np.array([wc.kpt_header(i + 1, spin_channel).kpt_frac_coords for i in range(wc.nkpts)])
    unfolder.verify_commens(b_pc, wc.B_matrix)               # the fold map needs the SC lattice on the device
    fs, _, _ = unfolder.get_geom_unfolding_relations(all_k, sc_k @ wc.B_matrix, n_sckpts_to_skip=n_sckpts_to_skip)
    folds = [[] for _ in range(wc.nkpts)]
    for ik, s_ in enumerate(fs):
        folds[int(s_)].append(ik)
    used = [i for i in range(wc.nkpts) if folds[i]]
    for j, i in enumerate(used):
        h = wc.kpt_header(i + 1, spin_channel)
        energies_of_job[j] = h.band_energies
        file_kpt_of_job[j] = i + 1

    got, grid = unfold_wavecar(unfolder, wc, b_pc, all_k, e_start, e_end, de, ispin=spin_channel,
                               after_fetch=after_fetch, n_sckpts_to_skip=n_sckpts_to_skip)
    n = all_k.shape[0]
    if got.row_ids.tolist() != list(range(n)):
        # main_BandUP.f90:114: every requested pc-kpt must have been unfolded exactly once
        raise RuntimeError("not every pc-kpt was unfolded")
    if origin_for_spin_proj_rec is not None:
        # -origin_for_spin_proj_rec: fractional w.r.t. the SC reciprocal lattice, wins over the
        # Cartesian one (cla_wrappers_mod.f90:452-467, band_unfolding_routines_mod.f90:183-189)
        origin_for_spin_proj = np.asarray(origin_for_spin_proj_rec, dtype=np.float64) @ wc.B_matrix
    if out_file:
        sigma = spin_proj = None
        if spinor_kpts:
            sigma = np.stack([sigma_by_row[i] for i in range(n)])
            spin_proj = np.zeros((n, grid.size, 2))
            for i in range(n):
                if np.linalg.norm(all_k[i] - (np.zeros(3) if origin_for_spin_proj is None
                                              else np.asarray(origin_for_spin_proj))) < 1e-12:
                    continue                      # the direction of k is undefined at the origin
                for ie in np.nonzero(np.any(sigma[i] != 0.0, axis=1))[0]:
                    spin_proj[i, ie] = calc_spin_projections(sigma[i, ie], all_k[i], normal_to_proj_plane,
                                                             origin_for_spin_proj)
        write_band_struc(out_file, kpath, grid, got.dN, e_fermi, sigma=sigma, spin_proj=spin_proj,
                         normal_to_proj_plane=normal_to_proj_plane)
    if want_rho:
        write_unf_dens_op(unf_dens_op_file, kpath, grid, got.dN, [rho_by_row[i] for i in range(n)], wc.nbands,
                          wc.B_matrix, b_pc, [got.sckpt_of_row[i] for i in range(n)], got.sckpts_cart,
                          e_fermi=e_fermi, spin_channel=spin_channel)
    if want_orb:
        ions = projections.get("selected_ion_indices")
        write_orb_dependent_ebs(orb_out_file, kpath, grid, np.stack([orb_by_row[i] for i in range(n)]), e_fermi,
                                projections["n_ions"], "All" if ions is None else ",".join(str(i + 1) for i in ions),
                                formatted_orb_choice(projections.get("picked_orbitals", "all"), projections["orbitals"]),
                                spin_channel)
    return kpath, grid, got.dN


# ---- unfolding_density_operator_*.dat (write_unf_dens_op, io_routines_mod.f90:1637-1882) ---------

@dataclass
class UdoRecord:
    """One (pc-kpt, energy) block of the rho file."""
    pckpt_number: int               # 1-based, in output order
    iener: int                      # 1-based index into the energy grid
    energy: float
    unfolded_N: float
    rows: np.ndarray                # 0-based m1 (upper triangle, m1 <= m2)
    cols: np.ndarray
    entries: np.ndarray             # complex
    folding_sckpt_number: int = 1
    pckpt_coord_in_plot: float = 0.0


def write_unf_dens_op(path, kpath: PcKptsPath, energy_grid: np.ndarray, dN: np.ndarray, rho_ops, n_sc_bands: int,
                      B_matrix_SC: np.ndarray, b_matrix_pc: np.ndarray, sckpt_of_pckpt: Sequence[int],
                      sckpts_cart: np.ndarray, e_fermi: float = 0.0, spin_channel: int = 1,
                      min_allowed_dN: float = 1e-3) -> None:
    """The text file `bandup projected-unfold` parses (python_interface/bandupy/
    unfolding_density_op.py:143-235).  rho_ops[ipc] is a list of (iener1, m1, m2, rho) arrays with the
    STORED elements of that pc-kpt (1-based band indices, both triangles allowed); like the
    reference only the upper triangle with |rho| >= 1e-4 is written, next to GenSpecWeight = rho*N."""
    grid = np.asarray(energy_grid, dtype=np.float64)
    Binv_sc = np.linalg.inv(np.asarray(B_matrix_SC, dtype=np.float64))
    binv_pc = np.linalg.inv(np.asarray(b_matrix_pc, dtype=np.float64))
    bar = "#" * 88
    with open(path, "w") as f:
        f.write(bar + "\n# File created by bandup-b200 (GPU unfolding path)\n" + bar + "\n")
        f.write("# This file contains all information needed to unfold the expectation values of any\n"
                "# operator defined in the (k,E) space. In particular, it reports, for each point in\n"
                f"# the (k,E) grid satisfying N(k,E) >= {_es(min_allowed_dN)}, the following information:\n"
                "#     * Unfolding-density operators (UnfDensOp)\n"
                "#     * Generalized spectral weight matrices (GenSpecWeight)\n"
                "# Both UnfDensOp and GenSpecWeight are hermitian; only the upper-triangular matrix\n"
                "# elements are present. They are represented as (RealPart, ImagPart), and were\n"
                "# evaluated between SC states with band indices indicated by m1 and m2.\n#\n")
        f.write(f"# SpinChannel = {spin_channel}\n# nScBands = {n_sc_bands}\n")
        f.write(f"# emin = {grid.min() - e_fermi:.5f} eV\n# emax = {grid.max() - e_fermi:.5f} eV\n")
        f.write(f"# nEner = {grid.size}\n# OriginalEfermi = {e_fermi:.5f} eV\n")
        f.write("# The energies were shifted so that the Fermi level is at 0 eV.\n" + bar + "\n")
        ipc = 0
        coord_first = 0.0
        coord_k = 0.0
        for d in kpath.directions:
            first = d[0]
            for k in d:
                coord_k = coord_first + float(np.linalg.norm(k - first))
                isc = int(sckpt_of_pckpt[ipc])
                K = np.asarray(sckpts_cart[isc], dtype=np.float64)
                f.write("\n")
                f.write(f"# PcKptNumber = {ipc + 1}\n")
                f.write("#     PcKptCartesianCoords =  %.8f  %.8f  %.8f  (A^{-1})\n" % tuple(k))
                f.write("#     PcKptFractionalCoords =  %.8f  %.8f  %.8f  (w.r.t. SCRL basis)\n" % tuple(k @ Binv_sc))
                f.write("#     PcKptFractionalCoords =  %.8f  %.8f  %.8f  (w.r.t. PCRL basis)\n" % tuple(k @ binv_pc))
                f.write("#     PcKptLinearCoordsInBandPlot = %.4f  (A^{-1})\n" % coord_k)
                f.write(f"#     Folds into ScKptNumber = {isc + 1}\n")
                f.write("#         ScKptCartesianCoords =  %.8f  %.8f  %.8f  (A^{-1})\n" % tuple(K))
                f.write("#         ScKptFractionalCoords =  %.8f  %.8f  %.8f  (w.r.t. SCRL basis)\n#\n"
                        % tuple(K @ Binv_sc))
                iener_a, m1_a, m2_a, rho_a = rho_ops[ipc]
                for ie in np.unique(iener_a):
                    N = float(dN[ipc, ie - 1])
                    if N < min_allowed_dN:
                        continue
                    s = iener_a == ie
                    m1s, m2s, vals = m1_a[s], m2_a[s], rho_a[s]
                    trace = float(np.sum(vals[m1s == m2s].real))
                    tr = "%8.1E" % trace + ("" if abs(trace - 1.0) < 1e-1 else " (!=1)")
                    f.write("#     iEnerPC = %d  E = %.4f eV  N(k,E) = %s  Tr{UnfDensOp} = %s\n"
                            % (ie, grid[ie - 1], _es(N), tr.strip()))
                    f.write("#        m1      m2        GenSpecWeight[m1,m2]         UnfDensOp[m1,m2]\n")
                    order = np.lexsort((m2s, m1s))
                    for j in order:
                        if m2s[j] < m1s[j] or abs(np.complex64(vals[j])) < np.float32(1e-4):
                            continue
                        g = vals[j] * N
                        f.write("      %5d   %5d      (%s, %s)   (%s, %s)\n"
                                % (m1s[j], m2s[j], _es(g.real), _es(g.imag), _es(vals[j].real), _es(vals[j].imag)))
                ipc += 1
            coord_first = coord_k


def read_unf_dens_op(path) -> Tuple[dict, List[UdoRecord]]:
    """Restatement of read_unf_dens_ops (unfolding_density_op.py:143-235): header dict + one record
    per (pc-kpt, energy) block that has at least one matrix element line."""
    hdr: dict = {}
    recs: List[UdoRecord] = []
    lines = Path(path).read_text().splitlines()
    pck, coord, isc = 0, 0.0, 1
    cur = None
    in_rows = False
    for i, line in enumerate(lines):
        nxt = lines[i + 1] if i + 1 < len(lines) else None
        if "SpinChannel" in line:
            hdr["spin_channel"] = int(line.split("=")[-1])
        elif "nScBands" in line:
            hdr["nbands"] = int(line.split("=")[-1])
        elif "emin" in line:
            hdr["emin"] = float(line.split("=")[1].split()[0])
        elif "emax" in line:
            hdr["emax"] = float(line.split("=")[1].split()[0])
        elif "nEner" in line:
            hdr["nener"] = int(line.split("=")[1])
        elif "PcKptNumber" in line:
            pck = int(line.split("=")[-1])
        elif "PcKptLinearCoordsInBandPlot" in line:
            coord = float(line.split("=")[-1].split()[0])
        elif "Folds into ScKptNumber" in line:
            isc = int(line.split("=")[-1])
        elif "iEnerPC" in line:
            t = line.split()
            cur = dict(iener=int(t[3]), energy=float(t[6]), N=float(t[10]), rows=[], cols=[], vals=[])
        elif "m1" in line and "m2" in line and "GenSpecWeight" in line and "UnfDensOp" in line:
            in_rows = True
        elif in_rows:
            t = line.replace(")", "").replace("(", "").replace(",", "").split()
            r, c = int(t[0]) - 1, int(t[1]) - 1
            v = complex(float(t[4]), float(t[5]) if r != c else 0.0)
            cur["rows"].append(r)
            cur["cols"].append(c)
            cur["vals"].append(v)
            if nxt is None or nxt.startswith("#") or not nxt.strip():
                in_rows = False
                recs.append(UdoRecord(pck, cur["iener"], cur["energy"], cur["N"], np.array(cur["rows"]),
                                      np.array(cur["cols"]), np.array(cur["vals"]), isc, coord))
    return hdr, recs
