"""Golden vectors for the Fortran half of the path, produced by EXECUTING the reference's own
source files (oracle/f90interp.py interprets them where they lie under /root/reference/src; this
image has no Fortran compiler).  Run in the build container only:

    python tests/golden/make_fortran_golden.py          # writes tests/golden/fortran_golden.npz/.json

What runs, per case, all of it the reference's text:
  get_rec_latt, check_if_pc_and_SC_are_commensurate              (crystals_mod.f90:41-73,186-202)
  n_Gvecs_2b_generated, populate_wf_with_G_vec_coords             (read_wavecar_mod.f90:151-304)
  the band renormalisation block of read_wavefunction             (io_routines_mod.f90:1317-1330)
  real_seq                                                        (lists_and_seqs_mod.f90:188-200)
  select_coeffs_to_calc_spectral_weights -> vec_in_latt -> same_vector (band_unfolding_routines_mod.f90:550-612)
  perform_unfolding as a whole: spectral_weight_for_coeff, get_delta_Ns_for_EBS -> calc_delta_N_pckpt
      -> lambda -> integral_delta_x_minus_x0, calc_rho and its sparsification, get_SC_pauli_matrix_elmnts,
      trace_AB, calc_spin_projections                              (:615-651, :701-786, :1046-1510)
  get_delta_Ns_for_EBS with a smearing, calc_spectral_function     (:654-698, :760-786)
The inputs (cells, k-points, synthetic coefficients and band energies) are stored next to the outputs.
"""
from __future__ import annotations

import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import f90interp as F  # noqa: E402

REF = Path("/root/reference/src")
FILES = ["constants_and_types_mod.f90", "math_mod.f90", "lists_and_seqs_mod.f90", "crystals_mod.f90",
         "band_unfolding_routines_mod.f90", "io_routines_mod.f90", "vasp_related_modules/read_wavecar_mod.f90",
         "symmetry_mod.f90"]
R8, I4, LG = np.float64, np.int32, np.bool_


def interpreter(perform_unfold=True, write_unf_dens_op=True, dp_build=False):
    it = F.Interp()
    for f in FILES:
        it.load(REF / f)
    if dp_build:
        # the build the reference asks for when the WAVECAR is double precision: "change the variable
        # kind_cplx_coeffs from sp to dp ... and recompile" (read_wavecar_mod.f90:363-377,
        # constants_and_types_mod.f90:60) -- here: the parameter is given the value dp before anything
        # that depends on it is evaluated
        assert "kind_cplx_coeffs" not in it.mod_vals
        it.mod_vals["kind_cplx_coeffs"] = F.Var(I4(int(it.module_value("dp").value)), it.mod_decls["kind_cplx_coeffs"])
    it.overrides["time_now"] = lambda: R8(0.0)
    args = F.FStruct("comm_line_args", perform_unfold=LG(perform_unfold), write_unf_dens_op=LG(write_unf_dens_op),
                     normal_to_proj_plane=np.array([0.0, 0.0, 1.0]), saxis=np.array([0.0, 0.0, 1.0]),
                     no_symm_sckpts=LG(True),
                     origin_for_spin_proj_passed_in_rec=LG(False), origin_for_spin_proj_cartesian=np.zeros(3),
                     origin_for_spin_proj_rec=np.zeros(3), n_sckpts_to_skip=I4(0))
    it.set_module_variable("args", args)
    return it


def decl(it, text):
    return F.parse_declaration(text, it._eval_module_const)[0]


def struct_array(it, typename, n):
    a = np.empty(n, dtype=object)
    for i in range(n):
        a[i] = it.new_struct(typename)
    return a


def run_case(name, a_pc, M, kpt_frac, ecut, n_spinor, n_bands, pc_kpts_frac_sc, grid, seed, *,
             perform_unfold=True, zero_band=None, folding_noise=0.0, smearing=None, sf_std=None, dir_weights=None,
             wavecar=None, dp_build=False):
    """pc_kpts_frac_sc: the pc-kpts to unfold onto, in fractional SC reciprocal coordinates
    (K + an SC reciprocal lattice vector each).  With dir_weights it is a list of such lists, one
    per needed direction (the symmetry-equivalent copies of ONE pcbz direction, all of them folding
    onto this SC-kpt here), and get_delta_Ns_for_output averages over them with those weights;
    rows of the per-pc-kpt outputs are then numbered direction-major.
    wavecar: dict(other_kpt=[..], extra_pad=n, adjust_c=bool) -- the wavefunction does not come from
    populate_wf_with_G_vec_coords + synthetic arrays but from the reference's READER: a synthetic
    two-k-point WAVECAR is written (bandup_b200.wavecar.write_wavecar; this k-point is the second one)
    and read_wavecar itself parses it (records, spinor decision, the 2m/hbar^2 adjustment loop when
    adjust_c asks for a file with one plane wave more than the stock constant yields)."""
    it = interpreter(perform_unfold, dp_build=dp_build)
    rng = np.random.default_rng(seed)
    out = {}
    a_pc = np.asarray(a_pc, dtype=R8)
    A_sc = np.asarray(M, dtype=R8) @ a_pc
    b_pc, B_sc = F.Var(np.zeros((3, 3)), None), F.Var(np.zeros((3, 3)), None)
    it.call("get_rec_latt", a_pc, b_pc)
    it.call("get_rec_latt", A_sc, B_sc)
    b_pc, B_sc = b_pc.value, B_sc.value
    crystal_pc = F.FStruct("crystal_3d", rec_latt_vecs=b_pc, latt_vecs=a_pc)
    crystal_sc = F.FStruct("crystal_3d", rec_latt_vecs=B_sc, latt_vecs=A_sc)
    comm, Mout = F.Var(LG(False), None), F.Var(np.zeros((3, 3)), None)
    it.call("check_if_pc_and_sc_are_commensurate", comm, Mout, crystal_pc, crystal_sc)
    out.update(a_pc=a_pc, A_sc=A_sc, b_pc=b_pc, B_sc=B_sc, commensurate=np.array(bool(comm.value)), M=Mout.value)

    # ---- the plane-wave list as read_wavecar builds it -----------------------------------
    wf = it.new_struct("pw_wavefunction")
    kf = np.asarray(kpt_frac, dtype=R8)
    b1, b2, b3 = B_sc[0].copy(), B_sc[1].copy(), B_sc[2].copy()
    e_start, e_end, d_e = grid
    egrid = F.Var(None, decl(it, "real(kind=dp), dimension(:), allocatable :: return_list"))
    it.call("real_seq", R8(e_start), R8(e_end), R8(d_e), egrid)
    egrid = egrid.value

    def synthetic(g_frac_, k_):
        kg = (g_frac_ + k_) @ B_sc
        env = np.exp(-np.sum(kg * kg, axis=1) / (2.0 * 0.2624658 * ecut / 4.0))
        n_ = len(g_frac_)
        c_ = (rng.standard_normal((n_spinor, n_, n_bands)) + 1j * rng.standard_normal((n_spinor, n_, n_bands)))
        c_ *= env[None, :, None] * rng.uniform(0.2, 3.0, size=(1, 1, n_bands))
        c_ = c_.astype(np.complex128 if dp_build else np.complex64)
        if zero_band is not None:
            c_[:, :, zero_band] = 0
        e_ = np.sort(rng.uniform(e_start - 0.3, e_end + 0.3, size=n_bands))
        if n_bands >= 4:
            e_[1] = e_[0]                                     # an exact duplicate
            e_[2] = egrid[len(egrid) // 2] + 0.5 * (egrid[1] - egrid[0])   # exactly on a bin edge
            e_[3] = e_[2] + 1e-3                             # a close neighbour (shares bins: rho pairs)
            e_ = np.sort(e_)
        return c_, e_

    if wavecar is None:
        # ---- the plane-wave list as read_wavecar builds it ---------------------------------
        n_g = int(it.call("n_gvecs_2b_generated", kf, R8(ecut), b1, b2, b3))
        wf["n_pw"], wf["n_bands"], wf["n_spinor"] = I4(n_g), I4(n_bands), I4(n_spinor)
        it.call("populate_wf_with_g_vec_coords", wf, kf, R8(ecut), b1, b2, b3)
        g_frac = np.array([g["coord"] for g in wf["g_frac"]], dtype=np.int32)
        c, ener = synthetic(g_frac, kf)
        wf["pw_coeffs"] = c
        wf["band_energies"] = ener.astype(R8)
    else:
        # ---- a synthetic WAVECAR, parsed by the reference's reader ---------------------------------
        import tempfile
        from bandup_b200 import synth
        from bandup_b200.wavecar import write_wavecar
        c0 = float(it.module_value("vasp_2m_over_hbar_sqrd").value)
        c_file = c0
        if wavecar.get("adjust_c"):
            # put the cutoff a hair below one plane wave's kinetic energy: the stock constant leaves it
            # out, the constant + 1e-10 takes it in -- the file holds one plane wave more than the
            # reader first expects, and its adjustment loop has to find that (:459-472)
            wide = synth.gen_gvecs(kf, ecut * 1.05, B_sc, c0)
            kg = (wide + kf) @ B_sc
            et = np.sum(kg * kg, axis=1) / c0
            target = np.sort(et[et < ecut])[-3]
            ecut = float(target * (1.0 - 3e-10))
            c_file = c0 + 1e-10
        kpts = [np.asarray(wavecar["other_kpt"], dtype=R8), kf]
        lists = [synth.gen_gvecs(k_, ecut, B_sc, c_file) for k_ in kpts]
        data = [synthetic(g_, k_) for g_, k_ in zip(lists, kpts)]
        with tempfile.TemporaryDirectory() as tmp:
            path = Path(tmp) / "WAVECAR"
            nrecl = write_wavecar(path, A_sc, ecut, kpts, [np.ascontiguousarray(np.transpose(d[0], (2, 1, 0))) for d in data],
                                  [d[1] for d in data], extra_pad=int(wavecar.get("extra_pad", 0)), double=dp_build)
            out["wavecar_bytes"] = np.frombuffer(path.read_bytes(), dtype=np.uint8).copy()
            out["wavecar_nrecl"] = np.array(nrecl)
            out["wavecar_ikpt"] = np.array(2)
            it.overrides["available_io_unit"] = lambda *a: I4(60 + len(it.units))
            wf["i_spin"] = I4(1)
            iost = F.Var(I4(-7), None)
            it.call("read_wavecar", wf, str(path), I4(2), iostat=iost)
            assert int(iost.value) == 0 and not it.units
        if wavecar.get("adjust_c"):
            assert len(lists[1]) == len(synth.gen_gvecs(kf, ecut, B_sc, c0)) + 1
        n_g = int(wf["n_pw"])
        assert n_g == len(lists[1]) and int(wf["n_spinor"]) == n_spinor and int(wf["n_bands"]) == n_bands
        g_frac = np.array([g["coord"] for g in wf["g_frac"]], dtype=np.int32)
        c, ener = wf["pw_coeffs"].copy(), wf["band_energies"].copy()
        out.update(reader_A_matrix=wf["a_matrix"].copy(), reader_B_matrix=wf["b_matrix"].copy(),
                   reader_encut=np.array(wf["encut"]), reader_vcell=np.array(wf["vcell"]),
                   reader_occupations=wf["band_occupations"].copy(), reader_kpt_frac=wf["kpt_frac_coords"].copy(),
                   reader_n_spin=np.array(wf["n_spin"]), reader_is_spinor=np.array(bool(wf["is_spinor"])))
    g_cart = np.array([g["coord"] for g in wf["g_cart"]], dtype=R8)
    out.update(kpt_frac=kf, ecut=np.array(ecut), g_frac=g_frac, g_cart=g_cart)
    out.update(coeffs_in=c.copy(), band_energies=ener, energy_grid=egrid)

    # ---- the renormalisation block of read_wavefunction -------------------------------------
    it.exec_statements(REF / "io_routines_mod.f90", 1317, 1330, {
        "wf": (wf, None),
        "read_coefficients": (LG(True), "logical :: read_coefficients"),
        "iband": (None, "integer :: iband"), "i": (None, "integer :: i"),
        "inner_prod": (None, _inner_prod_decl(it)),
    })
    out["coeffs_renormalised"] = wf["pw_coeffs"].copy()

    # ---- geometric unfolding relations for ONE SC-kpt, one pcbz direction, its needed directions ----
    K_cart = kf @ B_sc
    dirs = pc_kpts_frac_sc if dir_weights is not None else [pc_kpts_frac_sc]
    n_dirs, n_k = len(dirs), len(dirs[0])
    n_fold = n_dirs * n_k
    gur = it.new_struct("geom_unfolding_relations_for_each_sckpt")
    gur["list_of_sckpts"] = struct_array(it, "vec3d", 1)
    gur["list_of_sckpts"][0]["coord"][:] = K_cart
    gur["sckpt"] = struct_array(it, "selected_pcbz_directions", 1)
    gur["sckpt"][0]["selec_pcbz_dir"] = struct_array(it, "needed_pcbz_dirs_for_ebs", 1)
    gur["sckpt"][0]["selec_pcbz_dir"][0]["needed_dir"] = struct_array(it, "list_of_trial_folding_pckpts", n_dirs)
    delta_N = it.new_struct("unfoldedquantities")
    delta_N["selec_pcbz_dir"] = struct_array(it, "unfoldedquantities_info_for_needed_pcbz_dirs", 1)
    delta_N["selec_pcbz_dir"][0]["needed_dir"] = struct_array(it, "list_of_pckpts_for_unfolded_quantities", n_dirs)
    folding_G = np.zeros((n_fold, 3))
    pck_all, uq_all = [], []
    for d, kks in enumerate(dirs):
        pck = struct_array(it, "trial_folding_pckpt", n_k)
        gur["sckpt"][0]["selec_pcbz_dir"][0]["needed_dir"][d]["pckpt"] = pck
        uq = struct_array(it, "unfolded_quantities_for_given_pckpt", n_k)
        delta_N["selec_pcbz_dir"][0]["needed_dir"][d]["pckpt"] = uq
        for f, kk in enumerate(kks):
            k_cart = np.asarray(kk, dtype=R8) @ B_sc + folding_noise * rng.standard_normal(3)
            pck[f]["scoords"][:] = k_cart
            pck[f]["coords"][:] = k_cart
            pck[f]["folds"] = LG(True)
            folding_G[d * n_k + f] = k_cart - K_cart
            pck_all.append(pck[f])
            uq_all.append(uq[f])
    out["pckpts_cart"] = np.array([p["scoords"] for p in pck_all])
    out["K_cart"] = K_cart
    times = it.new_struct("timekeeping")
    nener = len(egrid)
    sel_all, w_all, dn_all = [], [], []
    rho_rows = []          # (row, iener, m1, m2, re, im), 1-based like the reference except the row
    sigma = np.zeros((n_fold, nener, 3))
    proj = np.zeros((n_fold, nener, 2))
    for f in range(n_fold):
        ci = gur["current_index"]
        ci["i_sckpt"], ci["i_selec_pcbz_dir"] = I4(1), I4(1)
        ci["i_needed_dirs"], ci["ipc_kpt"] = I4(f // n_k + 1), I4(f % n_k + 1)
        sel = F.Var(None, decl(it, "integer, dimension(:), allocatable :: selected_coeff_indices"))
        it.call("select_coeffs_to_calc_spectral_weights", sel, wf, crystal_pc, gur)
        if sel.value is None:
            sel_all.append(np.zeros(0, dtype=np.int32))
            w_all.append(np.full(n_bands, np.nan))
            dn_all.append(np.full(nener, np.nan))
            continue
        sel_all.append(sel.value.copy())
        it.call("perform_unfolding", delta_N, times, gur, wf, sel.value, egrid)
        q = uq_all[f]
        dn_all.append(q["dn"].copy())
        # the weights are a local of perform_unfolding: re-run its weight statements (:1335-1345)
        sw = it.exec_statements(REF / "band_unfolding_routines_mod.f90", 1335, 1351, {
            "wf": (wf, None), "times": (times, None),
            "selected_coeff_indices": (sel.value, decl(it, "integer, dimension(:), allocatable :: selected_coeff_indices")),
            "spectral_weight": (np.zeros(n_bands), decl(it, "real(kind=dp), dimension(:), allocatable :: spectral_weight")),
            "n_wf_components": (I4(n_spinor), "integer :: n_wf_components"),
            "i_spinor": (None, "integer :: i_spinor"),
        })["spectral_weight"].value
        w_all.append(sw.copy())
        rho_rows += rho_entries(f, q["rhos"])
        if n_spinor == 2:
            sigma[f] = q["sigma"]
            proj[f, :, 0] = q["spin_proj_perp"]
            proj[f, :, 1] = q["spin_proj_para"]
    out["folding_G"] = folding_G
    out["selected"] = np.concatenate(sel_all) if sel_all else np.zeros(0, dtype=np.int32)
    out["n_selected"] = np.array([len(s) for s in sel_all], dtype=np.int32)
    out["weights"] = np.array(w_all)
    out["delta_N"] = np.array(dn_all)
    out["rho"] = np.array(rho_rows, dtype=R8).reshape(-1, 6)
    if n_spinor == 2:
        out["sigma"] = sigma
        out["spin_proj"] = proj
    if dir_weights is not None:
        # ---- get_delta_Ns_for_output: the average over the needed directions (:789-1043) ----------
        all_dirs = struct_array(it, "irr_bz_directions", 1)
        all_dirs[0]["irr_dir"] = struct_array(it, "bz_direction", n_dirs)
        for d, wgt in enumerate(dir_weights):
            all_dirs[0]["irr_dir"][d]["weight"] = R8(wgt)
        only_sel = it.new_struct("unfoldedquantitiesforoutput")
        avgd = it.new_struct("unfoldedquantitiesforoutput")
        it.call("get_delta_ns_for_output", only_sel, avgd, delta_N, all_dirs, gur["sckpt"][0], times)
        pk = avgd["pcbz_dir"][0]["pckpt"]
        out["dir_weights"] = np.asarray(dir_weights, dtype=R8)
        out["n_dirs"] = np.array(n_dirs)
        out["avg_delta_N"] = np.array([p["dn"] for p in pk])
        avg_rows = []
        for k, p in enumerate(pk):
            avg_rows += rho_entries(k, p["rhos"], by_position=True)
        out["avg_rho"] = np.array(avg_rows, dtype=R8).reshape(-1, 6)
        if n_spinor == 2:
            out["avg_sigma"] = np.array([p["sigma"] for p in pk])
            out["avg_spin_proj"] = np.array([np.stack([p["spin_proj_perp"], p["spin_proj_para"]], axis=1) for p in pk])
        # ---- write_band_struc: the unfolded_EBS_*.dat file of the averaged quantities (io_routines_mod.f90:889-977)
        import tempfile
        it.overrides["file_header_bandup"] = lambda: "# File created by BandUP (interpreted for a golden vector)"
        with tempfile.TemporaryDirectory() as tmp:
            ebs = Path(tmp) / "unfolded_EBS_symmetry-averaged.dat"
            it.call("write_band_struc", str(ebs), gur["sckpt"][0], egrid, avgd, ef=R8(0.37), zero_of_kpts_scale=R8(0.25))
            assert not it.units
            out["ebs_file"] = np.frombuffer(ebs.read_bytes(), dtype=np.uint8).copy()
        out["ebs_e_fermi"], out["ebs_zero_of_kpts_scale"] = np.array(0.37), np.array(0.25)
        # what "only the selected directions" keeps is the first needed direction, unchanged
        assert all(np.array_equal(a["dn"], b["dn"]) for a, b in
                   zip(only_sel["pcbz_dir"][0]["pckpt"], delta_N["selec_pcbz_dir"][0]["needed_dir"][0]["pckpt"]))
    if smearing is not None:
        dn_s = []
        for f in range(n_fold):
            v = F.Var(None, decl(it, "real(kind=dp), dimension(:), allocatable :: delta_n_pckpt"))
            it.call("get_delta_ns_for_ebs", v, egrid, wf["band_energies"], np.nan_to_num(out["weights"][f]),
                    smearing_for_delta_function=R8(smearing))
            dn_s.append(v.value.copy())
        out["smearing"] = np.array(smearing)
        out["delta_N_smeared"] = np.array(dn_s)
    if sf_std is not None:
        v = F.Var(None, decl(it, "real(kind=dp), dimension(:), allocatable :: sf_at_pckpt"))
        it.call("calc_spectral_function", v, egrid, wf["band_energies"], np.nan_to_num(out["weights"][0]), R8(sf_std))
        out["sf_std"] = np.array(sf_std)
        out["spectral_function"] = v.value.copy()
    out["perform_unfold"] = np.array(perform_unfold)
    out["dp_build"] = np.array(dp_build)
    out["n_spinor"] = np.array(n_spinor)
    print(f"{name}: n_pw {n_g}, n_fold {n_fold}, selected {out['n_selected'].tolist()}, rho entries {len(rho_rows)}, "
          f"{it.n_statements} Fortran statements executed", flush=True)
    return out, it.n_statements


def run_gur_case(name, a_pc, M, sckpts_frac, fold_onto, g_shifts, n_skip, seed, noise=0.0):
    """get_geom_unfolding_relations (band_unfolding_routines_mod.f90:137-483, with get_star,
    reduce_point_to_bz, point_is_in_bz, pt_eqv_by_point_group_symop behind it) under -no_symm_sckpts:
    which SC-kpt every pc-kpt folds onto.  pc-kpt i = SC-kpt fold_onto[i] + the SC reciprocal lattice
    vector g_shifts[i] (+ noise, so that the wrapper has to raise its tolerance)."""
    it = interpreter()
    rng = np.random.default_rng(seed)
    args = it.mod_vals["args"].value
    args["n_sckpts_to_skip"] = I4(n_skip)
    a_pc = np.asarray(a_pc, dtype=R8)
    A_sc = np.asarray(M, dtype=R8) @ a_pc
    b_pc, B_sc = F.Var(np.zeros((3, 3)), None), F.Var(np.zeros((3, 3)), None)
    it.call("get_rec_latt", a_pc, b_pc)
    it.call("get_rec_latt", A_sc, B_sc)
    b_pc, B_sc = b_pc.value, B_sc.value
    crystal_pc, crystal_sc = it.new_struct("crystal_3d"), it.new_struct("crystal_3d")
    crystal_pc["latt_vecs"][:], crystal_pc["rec_latt_vecs"][:] = a_pc, b_pc
    crystal_sc["latt_vecs"][:], crystal_sc["rec_latt_vecs"][:] = A_sc, B_sc
    sck = np.asarray(sckpts_frac, dtype=R8) @ B_sc
    pck = np.array([sck[j] + np.asarray(g, dtype=R8) @ B_sc for j, g in zip(fold_onto, g_shifts)])
    pck = pck + noise * rng.standard_normal(pck.shape)
    n_k = len(pck)
    list_of_sckpts = struct_array(it, "vec3d", len(sck))
    for s_, v in zip(list_of_sckpts, sck):
        s_["coord"][:] = v
    to_check = it.new_struct("selected_pcbz_directions")
    to_check["selec_pcbz_dir"] = struct_array(it, "needed_pcbz_dirs_for_ebs", 1)
    to_check["selec_pcbz_dir"][0]["needed_dir"] = struct_array(it, "list_of_trial_folding_pckpts", 1)
    pts = struct_array(it, "trial_folding_pckpt", n_k)
    to_check["selec_pcbz_dir"][0]["needed_dir"][0]["pckpt"] = pts
    for p_, v in zip(pts, pck):
        p_["coords"][:] = v
    gur = it.new_struct("geom_unfolding_relations_for_each_sckpt")
    it.call("get_geom_unfolding_relations", gur, list_of_sckpts, to_check, crystal_pc, crystal_sc)
    folds = np.array([[bool(gur["sckpt"][s_]["selec_pcbz_dir"][0]["needed_dir"][0]["pckpt"][k]["folds"])
                       for k in range(n_k)] for s_ in range(len(sck))])
    fold_sckpt = np.array([int(np.nonzero(folds[:, k])[0][0]) if folds[:, k].any() else -1 for k in range(n_k)],
                          dtype=np.int32)
    attempts = it.call_counts["get_gur_not_public"]
    scoords = np.array([gur["sckpt"][fold_sckpt[k]]["selec_pcbz_dir"][0]["needed_dir"][0]["pckpt"][k]["scoords"]
                        if fold_sckpt[k] >= 0 else np.full(3, np.nan) for k in range(n_k)])
    out = dict(a_pc=a_pc, A_sc=A_sc, b_pc=b_pc, B_sc=B_sc, sckpts_cart=sck, pckpts_cart=pck, folds=folds,
               fold_sckpt=fold_sckpt, n_attempts=np.array(attempts), n_sckpts_to_skip=np.array(n_skip),
               n_pckpts=np.array(int(gur["n_pckpts"])), n_folding_pckpts=np.array(int(gur["n_folding_pckpts"])),
               sckpt_used=np.array(gur["sckpt_used_for_unfolding"], dtype=bool), scoords=scoords)
    print(f"{name}: {n_k} pc-kpts onto {len(sck)} SC-kpts, fold map {fold_sckpt.tolist()}, {attempts} attempt(s), "
          f"{it.n_statements} Fortran statements executed", flush=True)
    return out, it.n_statements


def rho_entries(row, rhos, by_position=False):
    """(row, iener, m1, m2, re, im) of the stored elements of an array of UnfoldDensityOpContainer, in the
    order the reference appended them.  by_position: the averaged operators sit at index iener of an
    array over the whole energy grid (:868-871); unused slots have no elements."""
    rows = []
    if rhos is None:
        return rows
    for pos, r in enumerate(rhos):
        if r["rho"] is None:
            continue
        ie = int(r["iener_in_full_pc_egrid"])
        if by_position:
            assert ie == pos + 1
        for v, bi in zip(r["rho"], r["band_indices"]):
            rows.append((row, ie, int(bi["indices"][0]), int(bi["indices"][1]), float(v.real), float(v.imag)))
    return rows


def _inner_prod_decl(it):
    """inner_prod as read_wavefunction declares it (the declaration is looked up in the source so
    that its kind is the reference's, not a guess)."""
    for no, s in it.files[str(REF / "io_routines_mod.f90")]:
        if F.is_declaration(s) and "inner_prod" in s and 1200 < no < 1317:
            for d in F.parse_declaration(s, it._eval_module_const):
                if d.name == "inner_prod":
                    return d
    raise RuntimeError("declaration of inner_prod not found")


tri = [[3.1, 0.0, 0.0], [0.4, 2.9, 0.0], [0.3, -0.2, 3.4]]            # triclinic pc (Angstrom)
hexa = [[3.16, 0.0, 0.0], [-1.58, 2.7366, 0.0], [0.0, 0.0, 9.0]]
fcc = [[0.0, 2.7, 2.7], [2.7, 0.0, 2.7], [2.7, 2.7, 0.0]]
SPECS = [
    # name, a_pc, M, K (frac SC), ecut, n_spinor, n_bands, pc-kpts (frac SC rec. coords), grid, seed, extras
    ("triclinic_scalar", tri, [[2, 0, 0], [1, 2, 0], [0, 1, 2]], [0.25, 0.1, -0.3], 58.0, 1, 6,
     [[0.25, 0.1, -0.3], [1.25, 0.1, -0.3], [0.25, 1.1, 0.7], [-0.75, 2.1, -1.3]], (-3.0, 2.0, 0.25), 11,
     dict(smearing=0.2, sf_std=0.15)),
    ("hex_spinor", hexa, [[2, 1, 0], [-1, 1, 0], [0, 0, 1]], [1.0 / 3.0, 0.0, 0.0], 43.0, 2, 5,
     [[1.0 / 3.0, 0.0, 0.0], [4.0 / 3.0, 1.0, 0.0]], (-2.0, 2.0, 0.5), 12, {}),
    ("fcc_negative_det", fcc, [[1, -1, 1], [-1, 1, 1], [1, 1, -2]], [0.0, 0.5, 0.5], 56.0, 1, 5,
     [[0.0, 0.5, 0.5], [1.0, 0.5, 1.5], [0.0, 1.5, 0.5]], (-2.0, 2.0, 0.5), 13, dict(zero_band=2)),
    ("tolerance_ladder", tri, [[2, 0, 0], [0, 1, 0], [0, 0, 1]], [0.1, 0.2, 0.3], 60.0, 1, 4,
     [[1.1, 0.2, 0.3]], (-1.0, 1.0, 0.5), 14, dict(folding_noise=1.2e-5)),
    ("dont_unfold", fcc, [[2, 0, 0], [0, 2, 0], [0, 0, 1]], [0.5, 0.0, 0.0], 40.0, 1, 5,
     [[0.5, 0.0, 0.0], [1.5, 1.0, 0.0]], (-2.0, 2.0, 0.5), 15, dict(perform_unfold=False)),
    # three needed directions x two pc-kpts each, averaged by get_delta_Ns_for_output
    ("symm_avg_scalar", tri, [[2, 0, 0], [1, 2, 0], [0, 1, 2]], [0.2, -0.1, 0.4], 45.0, 1, 6,
     [[[0.2, -0.1, 0.4], [1.2, -0.1, 0.4]], [[0.2, 0.9, 1.4], [-0.8, 1.9, 0.4]], [[1.2, 0.9, 0.4], [0.2, -0.1, 1.4]]],
     (-3.0, 2.0, 0.5), 16, dict(dir_weights=[0.5, 0.3, 0.2])),
    ("symm_avg_spinor", hexa, [[2, 1, 0], [-1, 1, 0], [0, 0, 1]], [0.0, 0.25, 0.0], 36.0, 2, 5,
     [[[0.0, 0.25, 0.0], [1.0, 0.25, 0.0]], [[1.0, 1.25, 0.0], [0.0, 1.25, 1.0]]],
     (-2.0, 2.0, 0.5), 17, dict(dir_weights=[0.75, 0.25])),
    # the wavefunction comes out of the reference's own WAVECAR reader (a synthetic two-k-point file)
    ("wavecar_scalar", tri, [[2, 0, 0], [1, 2, 0], [0, 1, 2]], [0.3, 0.15, -0.2], 40.0, 1, 5,
     [[0.3, 0.15, -0.2], [1.3, 1.15, -0.2]], (-2.0, 2.0, 0.5), 18,
     dict(wavecar=dict(other_kpt=[0.0, 0.0, 0.0], extra_pad=3))),
    ("wavecar_spinor_adjusted_cutoff", hexa, [[2, 1, 0], [-1, 1, 0], [0, 0, 1]], [0.21, 0.13, 0.05], 36.0, 2, 4,
     [[0.21, 0.13, 0.05], [1.21, 0.13, 0.05]], (-2.0, 2.0, 0.5), 19,
     dict(wavecar=dict(other_kpt=[0.5, 0.0, 0.0], extra_pad=0, adjust_c=True))),
    # the reference built with kind_cplx_coeffs = dp: complex128 coefficients, a WAVECAR_double
    ("dp_build_scalar", fcc, [[2, 0, 0], [0, 2, 0], [0, 0, 2]], [0.25, 0.25, 0.0], 44.0, 1, 5,
     [[0.25, 0.25, 0.0], [1.25, 0.25, 1.0], [0.25, 1.25, 0.0]], (-2.0, 2.0, 0.5), 20, dict(dp_build=True)),
    ("dp_build_wavecar_double_spinor", hexa, [[2, 1, 0], [-1, 1, 0], [0, 0, 1]], [0.1, 0.3, 0.0], 33.0, 2, 4,
     [[0.1, 0.3, 0.0], [1.1, 1.3, 0.0]], (-2.0, 2.0, 0.5), 21,
     dict(dp_build=True, wavecar=dict(other_kpt=[0.0, 0.0, 0.5], extra_pad=1))),
]


GUR_SPECS = [
    # name, a_pc, M, SC-kpts (frac), SC-kpt each pc-kpt folds onto, its SC reciprocal lattice shift, n_sckpts_to_skip, seed, noise
    ("gur_no_symmetry", tri, [[2, 0, 0], [1, 2, 0], [0, 1, 2]],
     [[0.0, 0.0, 0.0], [0.5, 0.0, 0.0], [0.25, 0.1, -0.3], [0.5, 0.5, 0.5], [0.1, 0.2, 0.3]],
     [2, 0, 4, 1, 3, 2, 4, 0], [[0, 0, 0], [1, 0, 0], [0, 1, 1], [-1, 2, 0], [1, 1, 1], [2, -1, 0], [0, 0, -2], [0, 3, 1]],
     0, 31, 0.0),
    # the first SC-kpt is skipped (-n_sckpts_to_skip 1) although a duplicate of it comes later in the list;
    # one pc-kpt is off its lattice point by ~2e-5, so the wrapper raises the tolerance
    ("gur_skip_and_tolerance", fcc, [[2, 0, 0], [0, 2, 0], [0, 0, 1]],
     [[0.5, 0.0, 0.0], [0.0, 0.5, 0.0], [0.5, 0.0, 0.0], [0.25, 0.25, 0.0]],
     [2, 1, 3, 2, 1], [[0, 0, 0], [1, 0, 1], [0, 1, 0], [1, 1, 0], [0, 0, 1]], 1, 32, 1.0e-5),
]


def main():
    t0 = time.time()
    cases = {}
    n_stmt = 0
    only = sys.argv[1:]
    for spec in GUR_SPECS:
        if only and spec[0] not in only:
            continue
        out, n = run_gur_case(*spec)
        n_stmt += n
        for k, v in out.items():
            cases[f"{spec[0]}/{k}"] = v
    for name, a_pc, M, K, ecut, nsp, nb, pck, grid, seed, extra in SPECS:
        if only and name not in only:
            continue
        out, n = run_case(name, a_pc, M, K, ecut, nsp, nb, pck, grid, seed, **extra)
        n_stmt += n
        for k, v in out.items():
            cases[f"{name}/{k}"] = v
    if only:
        return
    here = Path(__file__).resolve().parent
    np.savez_compressed(here / "fortran_golden.npz", **cases)
    manifest = {
        "source": "the reference's Fortran sources executed by oracle/f90interp.py (an interpreter; no Fortran "
                  "compiler exists in the build image)",
        "script": "tests/golden/make_fortran_golden.py",
        "reference_files_sha256": {f: hashlib.sha256((REF / f).read_bytes()).hexdigest() for f in FILES},
        "fortran_statements_executed": n_stmt,
        "cases": [s[0] for s in SPECS],
        "gur_cases": [s[0] for s in GUR_SPECS],
        "seconds": round(time.time() - t0, 1),
    }
    (here / "fortran_golden.json").write_text(json.dumps(manifest, indent=1) + "\n")
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
