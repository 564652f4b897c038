#!/usr/bin/env python
"""End-to-end example on a B200: a synthetic graphene 4x4 supercell job written in BandUP's own
file formats (prim_cell_lattice.in, supercell_lattice.in, KPOINTS_prim_cell.in, energy_info.in,
WAVECAR), unfolded on the GPU, results in unfolded_EBS_not-symmetry_averaged.dat -- the file
`bandup plot` reads -- plus the unfolding-density operators for `bandup projected-unfold`.

    python examples/unfold_synthetic_graphene.py [workdir]

Equivalent BandUP call (reference):  bandup unfold -no_symm_sckpts -no_symm_avg --orbitals ...
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bandup_b200 import bandup_files as BF  # noqa: E402
from bandup_b200 import synth  # noqa: E402
from bandup_b200.unfold import GpuUnfolder  # noqa: E402
from bandup_b200.wavecar import write_wavecar  # noqa: E402

work = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(tempfile.mkdtemp(prefix="bandup_b200_"))
work.mkdir(parents=True, exist_ok=True)

# --- inputs, the way a VASP + `bandup kpts-sc-get` run leaves them --------------------------------
a = synth.hex_cell(2.467, 15.0)
b = synth.rec_latt(a)
segments = [([1 / 3, 1 / 3, 0.0], [0.0, 0.0, 0.0]), ([0.0, 0.0, 0.0], [0.5, 0.0, 0.0]),
            ([0.5, 0.0, 0.0], [1 / 3, 1 / 3, 0.0])]
counts = [11, 11, 8]
kcart = np.concatenate([BF.kpts_line(np.array(s) @ b, np.array(e) @ b, n) for (s, e), n in zip(segments, counts)])
sc = synth.Scenario("graphene4x4", a, np.diag([4, 4, 1]), 120.0, 32, 1, kcart, e_fermi=-2.45)
BF.write_lattice(work / "prim_cell_lattice.in", a, "graphene primitive cell")
BF.write_lattice(work / "supercell_lattice.in", sc.A_sc, "graphene 4x4 supercell")
BF.write_pckpts(work / "KPOINTS_prim_cell.in", segments, counts, title="K-Gamma-M-K")
BF.write_energy_info(work / "energy_info.in", sc.e_fermi, sc.emin, sc.emax, sc.dE)
data = [synth.make_wavefunction(sc, iK) for iK in range(sc.n_sc_kpts)]
write_wavecar(work / "WAVECAR", sc.A_sc, sc.ecut, list(sc.sc_kpts_frac), [d[0] for d in data], [d[2] for d in data])
print(f"inputs in {work}: {sc.n_sc_kpts} SC k-points, {sc.n_bands} bands, ~{data[0][1].shape[0]} plane waves")

# --- the unfolding itself ------------------------------------------------------------------------
with GpuUnfolder(0) as gpu:
    kpath, grid, dN = BF.unfold_task(
        gpu, work / "WAVECAR", work / "prim_cell_lattice.in", work / "supercell_lattice.in",
        work / "KPOINTS_prim_cell.in", work / "energy_info.in",
        out_file=str(work / "unfolded_EBS_not-symmetry_averaged.dat"),
        unf_dens_op_file=str(work / "unfolding_density_operator_not-symm_avgd.dat"))
    print(f"{gpu.launch_count()} kernel launches")

tab = BF.read_band_struc(work / "unfolded_EBS_not-symmetry_averaged.dat")
print(f"wrote {tab.shape[0]} rows (k, E-E_F, delta_N); sum of delta_N over the window per k-point: "
      f"{dN.sum(axis=1).min():.3f} .. {dN.sum(axis=1).max():.3f}")
