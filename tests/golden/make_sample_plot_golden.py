#!/usr/bin/env python
"""Golden numbers from the reference's own sample files for `bandup plot`
(/root/reference/utils/sample_input_files_task_plot/): the three inputs (a lattice, a pc k-point
path, an energy file) verbatim and, from the reference-made output unfolded_EBS_symmetry-averaged.dat,
the k-point linear coordinates and the energy values it contains.  They pin the readers and the
KptCoord accumulation of write_band_struc (io_routines_mod.f90:889-977).  Run in the build container."""
import json
from pathlib import Path
import numpy as np

src = Path("/root/reference/utils/sample_input_files_task_plot")
tab = np.loadtxt(src / "unfolded_EBS_symmetry-averaged.dat", comments="#")
out = {
    "prim_cell_lattice.in": (src / "prim_cell_lattice.in").read_text(),
    "KPOINTS_prim_cell.in": (src / "KPOINTS_prim_cell.in").read_text(),
    "energy_info.in": (src / "energy_info.in").read_text(),
    "kpt_coords": sorted(set(np.round(tab[:, 0], 4).tolist())),
    "energies_min_max_n": [float(tab[:, 1].min()), float(tab[:, 1].max()), int(len(set(np.round(tab[:, 1], 4).tolist())))],
}
Path(__file__).with_name("sample_plot_golden.json").write_text(json.dumps(out, indent=1))
print(len(out["kpt_coords"]), out["energies_min_max_n"])
