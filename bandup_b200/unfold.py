"""Host-side mirror of the reference's `band_unfolding` module for the hot path.

Names, argument meaning and error behaviour follow the reference
(/root/reference/src/band_unfolding_routines_mod.f90 and main_BandUP.f90:131-177)
so that parity tests read like calls into the Fortran; every number is produced
by libbandup_gpu.so through the C ABI (include/bandup_gpu.h).  Nothing here
computes on the CPU and nothing here imports the oracle.

Array conventions (NumPy, C order):
  * lattices are (3,3) with ROW i = vector i, like `crystal%rec_latt_vecs(i,:)`;
  * pw_coeffs has shape (n_bands, n_pw, n_spinor) complex64, which is byte for
    byte the Fortran array pw_coeffs(n_spinor, n_pw, n_bands)
    (constants_and_types_mod.f90:151);
  * G_frac is (n_pw, 3) int32 == wf%G_frac(:)%coord(1:3).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import BandupGpuError, CONST, c_f64p, c_i32p

LAYOUT_FORTRAN_INMEM = CONST["BANDUP_GPU_LAYOUT_FORTRAN_INMEM"]
LAYOUT_VASP_RECORD = CONST["BANDUP_GPU_LAYOUT_VASP_RECORD"]
MAX_FOLDS = CONST["BANDUP_GPU_MAX_FOLDS"]
K_COSET = CONST["BANDUP_GPU_K_COSET"]
K_WEIGHTS = CONST["BANDUP_GPU_K_WEIGHTS"]
K_FINALIZE = CONST["BANDUP_GPU_K_FINALIZE"]
K_DELTA_N = CONST["BANDUP_GPU_K_DELTA_N"]
K_RHO = CONST["BANDUP_GPU_K_RHO"]
K_SYMM_AVG = CONST["BANDUP_GPU_K_SYMM_AVG"]
# `if(symm_avg_dNs(iener) < 1E-3) cycle` (band_unfolding_routines_mod.f90:906): a default-real literal
MIN_DN_SYMM_AVG = float(np.float32(1e-3))


def _fortran_order(latt: np.ndarray) -> np.ndarray:
    """(3,3) row=vector matrix -> the 9 doubles as Fortran stores latt(i,j)."""
    a = np.asarray(latt, dtype=np.float64)
    if a.shape != (3, 3):
        raise ValueError("lattice must be 3x3")
    return np.ascontiguousarray(a.T).ravel()


@dataclass
class PwWavefunction:
    """type(pw_wavefunction) (constants_and_types_mod.f90:140-153), hot-path fields only."""
    pw_coeffs: np.ndarray                 # complex64 (n_bands, n_pw, n_spinor) or a raw record buffer;
    #                                       complex128 for a `kind_cplx_coeffs = dp` build (OPT_COMPLEX128)
    G_frac: np.ndarray                    # int32 (n_pw, 3)
    band_energies: np.ndarray             # float64 (n_bands,)
    kpt_frac_coords: Optional[np.ndarray] = None
    n_pw: int = 0
    n_bands: int = 0
    n_spinor: int = 1
    layout: int = LAYOUT_FORTRAN_INMEM
    band_stride_bytes: int = 0            # 0: densely packed rows

    def __post_init__(self):
        self.G_frac = np.ascontiguousarray(self.G_frac, dtype=np.int32)
        self.band_energies = np.ascontiguousarray(self.band_energies, dtype=np.float64)
        if self.layout == LAYOUT_FORTRAN_INMEM and self.pw_coeffs.ndim == 3:
            c = self.pw_coeffs
            if c.dtype == np.complex128:
                c = np.ascontiguousarray(c)
            elif c.dtype != np.complex64 or not c.flags.c_contiguous:
                c = np.ascontiguousarray(c, dtype=np.complex64)
            self.pw_coeffs = c
            self.n_bands, self.n_pw, self.n_spinor = c.shape
        if not self.n_pw:
            self.n_pw = int(self.G_frac.shape[0])
        if not self.n_bands:
            self.n_bands = int(self.band_energies.shape[0])
        if not self.band_stride_bytes:
            dp = isinstance(self.pw_coeffs, np.ndarray) and self.pw_coeffs.dtype == np.complex128
            self.band_stride_bytes = self.n_pw * self.n_spinor * (16 if dp else 8)
        if self.G_frac.shape != (self.n_pw, 3):
            raise ValueError("G_frac must be (n_pw, 3)")
        if self.band_energies.shape != (self.n_bands,):
            raise ValueError("band_energies must be (n_bands,)")


@dataclass
class UnfoldedKpt:
    """What perform_unfolding leaves behind for the pc-kpts folding onto one SC-kpt."""
    spectral_weight: np.ndarray           # (n_fold, n_bands)   P_mK(k)
    dN: np.ndarray                        # (n_fold, nener)     delta_N(k, E)
    n_selected: np.ndarray                # (n_fold,)           size(selected_coeff_indices)
    band_norms: np.ndarray                # (n_bands,)          sum_G |C|^2 before renormalisation
    selected_coeff_indices: Optional[List[np.ndarray]] = None   # 1-based, ascending
    n_spinor: int = 1

    def pckpt_cannot_be_parsed(self) -> np.ndarray:
        """main_BandUP.f90:161-172: an empty selection is the reference's fatal branch."""
        return self.n_selected == 0


@dataclass
class UnfoldDensityOps:
    """rhos(:) of every pc-kpt of one SC-kpt as perform_unfolding stores them
    (band_unfolding_routines_mod.f90:1404-1427): COO, reference append order."""
    n_energies: int               # sum over pc-kpts of size(rhos)
    fold: np.ndarray              # 0-based pc-kpt index within the SC-kpt's fold list
    iener: np.ndarray             # iener_in_full_pc_egrid (1-based)
    m1: np.ndarray                # band_indices (1-based)
    m2: np.ndarray
    rho: np.ndarray               # complex64

    def dense(self, fold: int, iener: int, n_bands: int) -> np.ndarray:
        sel = (self.fold == fold) & (self.iener == iener)
        out = np.zeros((n_bands, n_bands), dtype=np.complex64)
        out[self.m1[sel] - 1, self.m2[sel] - 1] = self.rho[sel]
        return out

    def energies_of(self, fold: int) -> np.ndarray:
        return np.unique(self.iener[self.fold == fold])


class GpuUnfolder:
    """One device context; single caller thread (like the serial part of BandUP_main)."""

    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        h = C.c_int(-1)
        rc = self._lib.bandup_gpu_init(int(device), C.byref(h))
        if rc != 0:
            raise BandupGpuError(rc, self._lib.bandup_gpu_last_error(-1).decode())
        self.handle = h.value
        self.device = device
        self.nener = 0
        self.energy_grid: Optional[np.ndarray] = None
        self.M: Optional[np.ndarray] = None
        self._inflight = {}

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "handle", -1) >= 0:
            self._lib.bandup_gpu_finalize(self.handle)
            self.handle = -1

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        _capi.check(self.handle, rc)

    # -- run-wide set-up ----------------------------------------------------
    def verify_commens(self, rec_latt_pc: np.ndarray, rec_latt_SC: np.ndarray,
                       tol: float = 0.0) -> np.ndarray:
        """check_if_pc_and_SC_are_commensurate (crystals_mod.f90:41-73) as used by
        verify_commens; returns int_M with A_i = sum_j M_ij a_j.  Raises
        BandupGpuError(NOT_COMMENSURATE) where the reference prints and stops."""
        b = _fortran_order(rec_latt_pc)
        B = _fortran_order(rec_latt_SC)
        M = np.zeros(9, dtype=np.int32)
        self._check(self._lib.bandup_gpu_set_cells(
            self.handle, B.ctypes.data_as(c_f64p), b.ctypes.data_as(c_f64p), float(tol),
            M.ctypes.data_as(c_i32p)))
        self.M = M.reshape(3, 3).T.copy()  # Fortran order -> row-major
        return self.M

    def set_energy_grid(self, E_start: float, E_end: float, delta_e: float,
                        smearing_for_delta_function: Optional[float] = None) -> np.ndarray:
        """real_seq (lists_and_seqs_mod.f90:188-200) + sigma of get_delta_Ns_for_EBS (:774-777)."""
        n = C.c_int32(0)
        sm = -1.0 if smearing_for_delta_function is None else abs(float(smearing_for_delta_function))
        if smearing_for_delta_function is not None and sm == 0.0:
            sm = 1e-300  # present but zero: max(epsilon, 0) = epsilon
        self._check(self._lib.bandup_gpu_set_grid(self.handle, float(E_start), float(E_end),
                                                  float(delta_e), sm, C.byref(n)))
        self.nener = n.value
        grid = np.zeros(self.nener, dtype=np.float64)
        self._check(self._lib.bandup_gpu_get_grid(self.handle, grid.ctypes.data_as(c_f64p)))
        self.energy_grid = grid
        return grid

    def folding_g0(self, folding_G: np.ndarray):
        """folding_G = Scoords(k) - K (Cartesian) -> integer SC Miller triple + residual."""
        fg = np.ascontiguousarray(np.atleast_2d(folding_G), dtype=np.float64)
        g0 = np.zeros((fg.shape[0], 3), dtype=np.int32)
        res = C.c_double(0.0)
        self._check(self._lib.bandup_gpu_folding_g0(self.handle, fg.ctypes.data_as(c_f64p), fg.shape[0],
                                                    g0.ctypes.data_as(c_i32p), C.byref(res)))
        return g0, res.value

    def get_geom_unfolding_relations(self, pckpts_cart: np.ndarray, sckpts_cart: np.ndarray,
                                     eqv_points: Optional[Sequence[np.ndarray]] = None, n_sckpts_to_skip: int = 0):
        """get_geom_unfolding_relations (band_unfolding_routines_mod.f90:430-483): for every pc-kpt the
        SC-kpt it folds onto.  eqv_points[s] are the Cartesian points equivalent to SC-kpt s under the
        SC's symmetry operations (None: no symmetry, each SC-kpt alone).  Returns
        (fold_sckpt, fold_eqv, tol_used); raises BandupGpuError(FOLDING_VECTOR) when a pc-kpt folds
        onto none."""
        pk = np.ascontiguousarray(np.atleast_2d(pckpts_cart), dtype=np.float64)
        sk = np.ascontiguousarray(np.atleast_2d(sckpts_cart), dtype=np.float64)
        if eqv_points is None:
            eqv_points = [sk[i:i + 1] for i in range(sk.shape[0])]
        off = np.zeros(sk.shape[0] + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(e) for e in eqv_points])
        eq = np.ascontiguousarray(np.concatenate([np.atleast_2d(e) for e in eqv_points]), dtype=np.float64)
        fs = np.full(pk.shape[0], -1, dtype=np.int32)
        fe = np.full(pk.shape[0], -1, dtype=np.int32)
        tol = C.c_double(0.0)
        self._check(self._lib.bandup_gpu_fold_map(
            self.handle, pk.shape[0], pk.ctypes.data_as(c_f64p), sk.shape[0], off.ctypes.data_as(c_i32p),
            eq.ctypes.data_as(c_f64p), int(n_sckpts_to_skip), fs.ctypes.data_as(c_i32p), fe.ctypes.data_as(c_i32p),
            C.byref(tol)))
        return fs, fe, tol.value

    # -- per SC k-point -----------------------------------------------------
    def submit(self, ikpt: int, wf: PwWavefunction, folding_G: Optional[np.ndarray] = None,
               g0: Optional[np.ndarray] = None) -> None:
        """Asynchronous: upload + kernels for all pc-kpts folding onto SC-kpt `ikpt`."""
        if g0 is None:
            g0, _ = self.folding_g0(folding_G)
        g0 = np.ascontiguousarray(np.atleast_2d(g0), dtype=np.int32)
        n_fold = g0.shape[0]
        coeffs = wf.pw_coeffs
        ptr = coeffs.ctypes.data if isinstance(coeffs, np.ndarray) else int(coeffs)
        self._check(self._lib.bandup_gpu_submit_kpt(
            self.handle, int(ikpt), wf.n_spinor, wf.n_pw, wf.n_bands, wf.layout,
            int(wf.band_stride_bytes), C.c_void_p(ptr), wf.G_frac.ctypes.data_as(c_i32p),
            wf.band_energies.ctypes.data_as(c_f64p), n_fold, g0.ctypes.data_as(c_i32p)))
        # keep the host arrays alive until fetch (the upload is asynchronous)
        self._inflight[int(ikpt)] = (wf, g0, n_fold)

    def submit_vasp(self, ikpt: int, band_records: np.ndarray, nrecl: int, nplane: int, n_bands: int,
                    kpt_frac_coords: np.ndarray, ecut: float, band_energies: np.ndarray,
                    folding_G: Optional[np.ndarray] = None, g0: Optional[np.ndarray] = None,
                    two_m_over_hbar_sqrd: float = 0.0):
        """Raw WAVECAR band records in; plane-wave list, scalar/spinor decision and the reader's
        2m/hbar^2 adjustment done on the device (read_wavecar_mod.f90:213-304, 446-472); otherwise
        like submit().  Returns (n_spinor, n_pw)."""
        if g0 is None:
            g0, _ = self.folding_g0(folding_G)
        g0 = np.ascontiguousarray(np.atleast_2d(g0), dtype=np.int32)
        k = np.ascontiguousarray(kpt_frac_coords, dtype=np.float64)
        e = np.ascontiguousarray(band_energies, dtype=np.float64)
        nsp, npw = C.c_int(0), C.c_int(0)
        self._check(self._lib.bandup_gpu_submit_kpt_vasp(
            self.handle, int(ikpt), int(nplane), int(n_bands), int(nrecl), C.c_void_p(band_records.ctypes.data),
            k.ctypes.data_as(c_f64p), float(ecut), float(two_m_over_hbar_sqrd), e.ctypes.data_as(c_f64p),
            g0.shape[0], g0.ctypes.data_as(c_i32p), C.byref(nsp), C.byref(npw)))
        wf = PwWavefunction(band_records, np.zeros((npw.value, 3), np.int32), e, n_pw=npw.value, n_bands=n_bands,
                            n_spinor=nsp.value, layout=LAYOUT_VASP_RECORD, band_stride_bytes=nrecl)
        self._inflight[int(ikpt)] = (wf, g0, g0.shape[0])
        return nsp.value, npw.value

    def fetch_gvecs(self, ikpt: int, n_pw: int) -> np.ndarray:
        g = np.zeros((n_pw, 3), dtype=np.int32)
        self._check(self._lib.bandup_gpu_fetch_gvecs(self.handle, int(ikpt), g.ctypes.data_as(c_i32p), n_pw))
        return g

    def fetch_raw(self, ikpt: int, n_fold: int, n_bands: int):
        """bandup_gpu_fetch_kpt without the bookkeeping of submit(): weights, dN, n_selected."""
        w = np.zeros((n_fold, n_bands), dtype=np.float64)
        dN = np.zeros((n_fold, max(self.nener, 1)), dtype=np.float64)
        ns = np.zeros(n_fold, dtype=np.int32)
        self._check(self._lib.bandup_gpu_fetch_kpt(
            self.handle, int(ikpt), w.ctypes.data_as(c_f64p), dN.ctypes.data_as(c_f64p),
            ns.ctypes.data_as(c_i32p), None, None, 0))
        self._inflight.pop(int(ikpt), None)
        return w, dN, ns

    def fetch(self, ikpt: int, want_selected: bool = False) -> UnfoldedKpt:
        wf, g0, n_fold = self._inflight.pop(int(ikpt))
        w = np.zeros((n_fold, wf.n_bands), dtype=np.float64)
        dN = np.zeros((n_fold, self.nener), dtype=np.float64)
        ns = np.zeros(n_fold, dtype=np.int32)
        norms = np.zeros(wf.n_bands, dtype=np.float64)
        sel = None
        cap = 0
        if want_selected:
            cap = n_fold * wf.n_pw
            sel = np.zeros(cap, dtype=np.int32)
        self._check(self._lib.bandup_gpu_fetch_kpt(
            self.handle, int(ikpt), w.ctypes.data_as(c_f64p), dN.ctypes.data_as(c_f64p),
            ns.ctypes.data_as(c_i32p), norms.ctypes.data_as(c_f64p),
            sel.ctypes.data_as(c_i32p) if sel is not None else None, cap))
        lists = None
        if sel is not None:
            off = np.concatenate([[0], np.cumsum(ns)])
            lists = [sel[off[f]:off[f + 1]].copy() for f in range(n_fold)]
        return UnfoldedKpt(w, dN, ns, norms, lists, n_spinor=wf.n_spinor)

    def perform_unfolding(self, wf: PwWavefunction, folding_G: Optional[np.ndarray] = None,
                          g0: Optional[np.ndarray] = None, ikpt: int = 1,
                          want_selected: bool = True) -> UnfoldedKpt:
        """The body of the reference's `if(pckpt_folds)` block (main_BandUP.f90:145-172) for
        every pc-kpt folding onto this SC-kpt: renormalise -> select -> weights -> delta_N."""
        self.submit(ikpt, wf, folding_G=folding_G, g0=g0)
        return self.fetch(ikpt, want_selected=want_selected)

    def select_coeffs_to_calc_spectral_weights(self, wf: PwWavefunction,
                                               folding_G: np.ndarray) -> List[np.ndarray]:
        """band_unfolding_routines_mod.f90:550-612; returns selected_coeff_indices (1-based,
        ascending) per pc-kpt.  An empty list is the reference's 'not allocated'."""
        return self.perform_unfolding(wf, folding_G, want_selected=True).selected_coeff_indices

    def set_perform_unfold(self, perform_unfold: bool) -> None:
        """args%perform_unfold; False == the reference's -dont_unfold (every weight is 1)."""
        self._check(self._lib.bandup_gpu_set_perform_unfold(self.handle, int(bool(perform_unfold))))

    def calc_spectral_function(self, ikpt: int, n_fold: int, std_dev: Optional[float] = None) -> np.ndarray:
        """calc_spectral_function (band_unfolding_routines_mod.f90:654-698) for a fetched, still
        resident SC-kpt: A(k, E_i) with a Gaussian delta (default std_dev 0.025 eV)."""
        out = np.zeros((n_fold, self.nener), dtype=np.float64)
        self._check(self._lib.bandup_gpu_spectral_function(
            self.handle, int(ikpt), -1.0 if std_dev is None else float(std_dev), out.ctypes.data_as(c_f64p)))
        return out

    # -- projected variant ------------------------------------------------------
    def calc_rho(self, ikpt: int, min_dN: float = 1e-3) -> "UnfoldDensityOps":
        """The energy selection + calc_rho loop of perform_unfolding
        (band_unfolding_routines_mod.f90:1372-1428, :1046-1161) for every pc-kpt of an SC-kpt that
        has been fetched and is still resident.  Returns the stored elements (|rho| >= 1e-5, both
        triangles) in the reference's append order."""
        ne, nn = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.bandup_gpu_rho_count(self.handle, int(ikpt), float(min_dN), C.byref(ne),
                                                   C.byref(nn)))
        n = nn.value
        fold = np.zeros(n, dtype=np.int32)
        iener = np.zeros(n, dtype=np.int32)
        m1 = np.zeros(n, dtype=np.int32)
        m2 = np.zeros(n, dtype=np.int32)
        re = np.zeros(n, dtype=np.float32)
        im = np.zeros(n, dtype=np.float32)
        f32p = C.POINTER(C.c_float)
        self._check(self._lib.bandup_gpu_fetch_rho(
            self.handle, int(ikpt), n, fold.ctypes.data_as(c_i32p), iener.ctypes.data_as(c_i32p),
            m1.ctypes.data_as(c_i32p), m2.ctypes.data_as(c_i32p), re.ctypes.data_as(f32p),
            im.ctypes.data_as(f32p)))
        return UnfoldDensityOps(ne.value, fold, iener, m1, m2, (re + 1j * im).astype(np.complex64))

    # -- symmetry averaging over the needed directions ---------------------------
    def symm_average(self, weights: np.ndarray, per_dir: np.ndarray) -> np.ndarray:
        """The dense averaging loops of get_delta_Ns_for_output (band_unfolding_routines_mod.f90:826-853
        delta_N; :984-1036 spin_proj_perp / spin_proj_para / sigma): sum_d weights[d] * per_dir[d],
        accumulated in direction order in dp.  per_dir: (n_dirs, ...) -> (...)."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        x = np.ascontiguousarray(per_dir, dtype=np.float64)
        if x.ndim < 1 or x.shape[0] != w.shape[0]:
            raise ValueError("per_dir must have one leading slice per direction")
        out = np.zeros(x.shape[1:], dtype=np.float64)
        self._check(self._lib.bandup_gpu_symm_average(self.handle, w.shape[0], out.size,
                                                      w.ctypes.data_as(c_f64p), x.ctypes.data_as(c_f64p),
                                                      out.ctypes.data_as(c_f64p)))
        return out

    def symm_average_rho(self, weights: np.ndarray, per_dir: "list[UnfoldDensityOps]", avg_dN: np.ndarray,
                         n_bands: int, min_dN: float = MIN_DN_SYMM_AVG) -> "UnfoldDensityOps":
        """The symmetry average of the unfolding-density operators (band_unfolding_routines_mod.f90:
        857-973).  per_dir[d]: the operators of direction d with `fold` = index of the pc-kpt along
        the direction, in (fold, iener, m1, m2) order; avg_dN (n_pckpts, nener) the averaged delta_N.
        Returns the averaged operators in the reference's element order."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        dn = np.ascontiguousarray(avg_dN, dtype=np.float64)
        if dn.ndim != 2 or len(per_dir) != w.shape[0]:
            raise ValueError("avg_dN must be (n_pckpts, nener) and per_dir one entry per direction")
        off = np.zeros(len(per_dir) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(r.rho) for r in per_dir])
        n = int(off[-1])
        cat = lambda name, dt: np.ascontiguousarray(  # noqa: E731
            np.concatenate([np.asarray(getattr(r, name)) for r in per_dir]) if n else np.zeros(0), dtype=dt)
        fold, iener, m1, m2 = (cat(a, np.int32) for a in ("fold", "iener", "m1", "m2"))
        rho = cat("rho", np.complex64)
        re, im = np.ascontiguousarray(rho.real), np.ascontiguousarray(rho.imag)
        outs = [np.zeros(max(n, 1), dtype=np.int32) for _ in range(4)]
        ore, oim = np.zeros(max(n, 1), dtype=np.float32), np.zeros(max(n, 1), dtype=np.float32)
        f32p = C.POINTER(C.c_float)
        n_out = C.c_int64(0)
        self._check(self._lib.bandup_gpu_symm_average_rho(
            self.handle, w.shape[0], dn.shape[0], dn.shape[1], int(n_bands), w.ctypes.data_as(c_f64p),
            dn.ctypes.data_as(c_f64p), float(min_dN), off.ctypes.data_as(C.POINTER(C.c_int64)),
            *[a.ctypes.data_as(c_i32p) for a in (fold, iener, m1, m2)], re.ctypes.data_as(f32p),
            im.ctypes.data_as(f32p), n, C.byref(n_out), *[a.ctypes.data_as(c_i32p) for a in outs],
            ore.ctypes.data_as(f32p), oim.ctypes.data_as(f32p)))
        m = n_out.value
        o_fold, o_iener = outs[0][:m].copy(), outs[1][:m].copy()
        n_ops = int(np.unique(o_fold.astype(np.int64) * dn.shape[1] + o_iener).size)
        return UnfoldDensityOps(n_ops, o_fold, o_iener, outs[2][:m].copy(), outs[3][:m].copy(),
                                (ore[:m] + 1j * oim[:m]).astype(np.complex64))

    def pauli_vectors(self, ikpt: int, n_fold: int) -> np.ndarray:
        """unf_pauli_vec = Re Tr(rho . sigma_i) for every (pc-kpt, energy) with a rho
        (get_SC_pauli_matrix_elmnts + band_unfolding_routines_mod.f90:1459-1481); (n_fold, nener, 3)."""
        out = np.zeros((n_fold, self.nener, 3), dtype=np.float64)
        self._check(self._lib.bandup_gpu_pauli_vectors(self.handle, int(ikpt), out.ctypes.data_as(c_f64p)))
        return out

    def get_orbital_contribution_matrix(self, proj_matrix_dual: np.ndarray, proj_matrix: np.ndarray,
                                        ao_mask: np.ndarray, band_energies: np.ndarray,
                                        max_band_dE: float = 0.1, min_abs: float = 1e-2) -> np.ndarray:
        """KptInfo.get_orbital_contribution_matrix, use_dual=True
        (bandupy/orbital_contributions.py:304-362).  ao_mask selects the combined atom/orbital
        indices iat*n_orbs_per_atom + iorb.  The result also becomes the current operator."""
        dual = np.ascontiguousarray(proj_matrix_dual, dtype=np.complex128)
        proj = np.ascontiguousarray(proj_matrix, dtype=np.complex128)
        nb, n_ao = dual.shape
        if proj.shape != (n_ao, nb):
            raise ValueError("proj_matrix must be (n_ao, n_bands) and its dual (n_bands, n_ao)")
        mask = np.ascontiguousarray(ao_mask, dtype=np.uint8)
        ener = np.ascontiguousarray(band_energies, dtype=np.float64)
        O = np.zeros((nb, nb), dtype=np.complex128)
        self._check(self._lib.bandup_gpu_orb_contrib_matrix(
            self.handle, nb, n_ao, dual.ctypes.data_as(c_f64p), proj.ctypes.data_as(c_f64p),
            mask.ctypes.data_as(C.POINTER(C.c_uint8)), ener.ctypes.data_as(c_f64p), float(max_band_dE),
            float(min_abs), O.ctypes.data_as(c_f64p)))
        return O

    def set_operator(self, general_operator: np.ndarray) -> None:
        O = np.ascontiguousarray(general_operator, dtype=np.complex128)
        if O.ndim != 2 or O.shape[0] != O.shape[1]:
            raise ValueError("Incompatible matrix shapes!")
        self._check(self._lib.bandup_gpu_set_operator(self.handle, O.shape[0], O.ctypes.data_as(c_f64p)))

    def unfold_operator(self, ikpt: int, n_fold: int, min_abs_rho: float = 1e-5,
                        clip: bool = True) -> np.ndarray:
        """UnfDensOp.unfold(O, multiply_by_N=True, discard_imag=True, clip_interval=(0, N)) for every
        (pc-kpt, energy) of SC-kpt `ikpt` (bandupy/unfolding_density_op.py:121-141 as called from
        orbital_contributions.py:444-458); 0 where no operator exists."""
        out = np.zeros((n_fold, self.nener), dtype=np.float64)
        self._check(self._lib.bandup_gpu_unfold_operator(self.handle, int(ikpt), float(min_abs_rho),
                                                         int(bool(clip)), out.ctypes.data_as(c_f64p)))
        return out

    # -- device-resident path (bench, torch interop) -------------------------
    def workspace_bytes(self, n_spinor: int, n_pw: int, n_bands: int, n_fold: int) -> int:
        return int(self._lib.bandup_gpu_workspace_bytes(self.handle, n_spinor, n_pw, n_bands, n_fold))

    def unfold_device(self, stream: int, n_spinor: int, n_pw: int, n_bands: int, layout: int,
                      band_stride_bytes: int, d_coeffs: int, d_g_frac: int, d_band_energies: int,
                      g0: np.ndarray, d_weights: int, d_dN: int, d_n_selected: int,
                      d_workspace: int, workspace_bytes: int, d_norms: int = 0,
                      d_slot_out: int = 0) -> None:
        g0 = np.ascontiguousarray(np.atleast_2d(g0), dtype=np.int32)
        self._check(self._lib.bandup_gpu_unfold_device(
            self.handle, C.c_void_p(stream), n_spinor, n_pw, n_bands, layout, int(band_stride_bytes),
            C.c_void_p(d_coeffs), C.c_void_p(d_g_frac), C.c_void_p(d_band_energies), g0.shape[0],
            g0.ctypes.data_as(c_i32p), C.c_void_p(d_weights), C.c_void_p(d_dN),
            C.c_void_p(d_n_selected), C.c_void_p(d_norms or None), C.c_void_p(d_slot_out or None),
            C.c_void_p(d_workspace), int(workspace_bytes)))

    def make_batch(self, kpts: Sequence[dict]):
        """Array of bandup_gpu_kpt_desc from dicts with the keyword names of unfold_device
        (n_spinor, n_pw, n_bands, layout, band_stride_bytes, d_coeffs, d_g_frac, d_band_energies, g0,
        d_weights, d_dN, d_n_selected[, d_norms]).  Returns (descs, keepalive)."""
        arr = (_capi.KptDesc * len(kpts))()
        keep = []
        for d, k in zip(arr, kpts):
            g0 = np.ascontiguousarray(np.atleast_2d(k["g0"]), dtype=np.int32)
            keep.append(g0)
            d.n_spinor, d.n_pw, d.n_bands = int(k["n_spinor"]), int(k["n_pw"]), int(k["n_bands"])
            d.layout, d.n_fold, d.reserved = int(k["layout"]), g0.shape[0], 0
            d.band_stride_bytes = int(k["band_stride_bytes"])
            d.d_coeffs, d.d_g_frac, d.d_band_energies = k["d_coeffs"], k["d_g_frac"], k["d_band_energies"]
            d.g0 = g0.ctypes.data
            d.d_weights, d.d_dN, d.d_n_selected = k["d_weights"], k["d_dN"], k["d_n_selected"]
            d.d_norms = k.get("d_norms") or None
        return arr, keep

    def batch_workspace_bytes(self, descs) -> int:
        return int(self._lib.bandup_gpu_batch_workspace_bytes(self.handle, len(descs), descs))

    def unfold_device_batch(self, stream: int, descs, d_workspace: int, workspace_bytes: int) -> None:
        """All k-points of `descs` with one launch per kernel (K1, K2, K3)."""
        self._check(self._lib.bandup_gpu_unfold_device_batch(
            self.handle, C.c_void_p(stream), len(descs), descs, C.c_void_p(d_workspace), int(workspace_bytes)))

    def synth_fill_device(self, stream: int, d_coeffs: int, row_elems: int, n_bands: int,
                          band_stride_bytes: int, seed: int, ikpt: int) -> None:
        self._check(self._lib.bandup_gpu_synth_fill_device(
            self.handle, C.c_void_p(stream), C.c_void_p(d_coeffs), int(row_elems), int(n_bands),
            int(band_stride_bytes), C.c_uint64(seed), C.c_uint64(ikpt)))

    # -- multi-GPU gather (NCCL inside the library) ------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL id made by rank 0; hand it to the other ranks, then comm_init everywhere."""
        lib = _capi.load()
        buf = C.create_string_buffer(CONST["BANDUP_GPU_COMM_ID_BYTES"])
        rc = lib.bandup_gpu_comm_unique_id(buf)
        if rc != 0:
            raise BandupGpuError(rc, lib.bandup_gpu_last_error(-1).decode())
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes) -> None:
        buf = C.create_string_buffer(unique_id, CONST["BANDUP_GPU_COMM_ID_BYTES"])
        self._check(self._lib.bandup_gpu_comm_init(self.handle, int(n_ranks), int(rank), buf))

    def comm_finalize(self) -> None:
        self._check(self._lib.bandup_gpu_comm_finalize(self.handle))

    def gather_dN(self, dN_local: np.ndarray, row_ids_local: np.ndarray):
        """ncclAllGather of the disjoint delta_N rows of all ranks; returns (dN_all, row_ids_all)
        in ascending global row id on every rank."""
        return self.gather_rows(np.asarray(dN_local, dtype=np.float64).reshape(-1, self.nener), row_ids_local)

    def gather_rows(self, rows_local: np.ndarray, row_ids_local: np.ndarray):
        """The same collective for rows of any shape (sigma: (n, nener, 3)); returns
        (rows_all, row_ids_all) in ascending global row id on every rank."""
        r = np.ascontiguousarray(rows_local, dtype=np.float64)
        ids = np.ascontiguousarray(row_ids_local, dtype=np.int64).reshape(-1)
        if r.ndim < 2 or r.shape[0] != ids.shape[0]:
            raise ValueError("rows_local must be (n_rows, ...) with one row id per row")
        shape = r.shape[1:]
        row_len = int(np.prod(shape))
        d = r.reshape(-1, row_len)
        i64p = C.POINTER(C.c_int64)
        tot = C.c_int64(0)
        self._check(self._lib.bandup_gpu_gather_rows(self.handle, row_len, d.ctypes.data_as(c_f64p),
                                                     ids.ctypes.data_as(i64p), ids.shape[0], None, None, 0,
                                                     C.byref(tot)))
        out = np.zeros((tot.value, row_len), dtype=np.float64)
        oid = np.zeros(tot.value, dtype=np.int64)
        self._check(self._lib.bandup_gpu_gather_rows(self.handle, row_len, d.ctypes.data_as(c_f64p),
                                                     ids.ctypes.data_as(i64p), ids.shape[0],
                                                     out.ctypes.data_as(c_f64p), oid.ctypes.data_as(i64p),
                                                     tot.value, C.byref(tot)))
        return out.reshape((tot.value,) + shape), oid

    def allgather_device(self, stream: int, d_send: int, d_recv: int, count: int) -> None:
        """ncclAllGather of `count` doubles per rank between device buffers on the caller's stream
        (the device-resident form of gather_dN; asynchronous)."""
        self._check(self._lib.bandup_gpu_allgather_device(self.handle, C.c_void_p(stream), C.c_void_p(d_send),
                                                          C.c_void_p(d_recv), int(count)))

    # -- instrumentation ------------------------------------------------------
    def set_profiling(self, enabled: bool, kernels=None) -> None:
        """Bracket launches with CUDA events; `kernels` (ids) restricts it to those kernels."""
        if enabled and kernels is not None:
            mask = 0
            for k in kernels:
                mask |= 1 << int(k)
            self._check(self._lib.bandup_gpu_set_profiling_mask(self.handle, mask))
        else:
            self._check(self._lib.bandup_gpu_set_profiling(self.handle, int(bool(enabled))))

    def reset_kernel_times(self) -> None:
        self._check(self._lib.bandup_gpu_reset_kernel_times(self.handle))

    def kernel_time(self, which: int):
        ms = C.c_double(0.0)
        n = C.c_int64(0)
        self._check(self._lib.bandup_gpu_kernel_time(self.handle, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self._lib.bandup_gpu_launch_count(self.handle))

    def set_tuning(self, chunk_vecs: int = 0, ctas_per_sm: int = 0, n_stages: int = 0) -> None:
        self._check(self._lib.bandup_gpu_set_tuning(self.handle, int(chunk_vecs), int(ctas_per_sm),
                                                    int(n_stages)))

    def set_option(self, option: int, value: int) -> None:
        """bandup_gpu_set_option: CONST['BANDUP_GPU_OPT_K2_VARIANT'] / ['BANDUP_GPU_OPT_DEPENDENT_LAUNCH'] /
        ['BANDUP_GPU_OPT_COMPLEX128']."""
        self._check(self._lib.bandup_gpu_set_option(self.handle, int(option), int(value)))

    def set_complex128(self, enabled: bool) -> None:
        """The reference's `kind_cplx_coeffs = dp` build (constants_and_types_mod.f90:60): from now on
        every coefficient block handed to this unfolder is complex128."""
        self.set_option(CONST["BANDUP_GPU_OPT_COMPLEX128"], int(bool(enabled)))


def calc_spin_projections(pauli_vector, cart_coords_kpt, normal_to_proj_plane, origin=None):
    """calc_spin_projections (band_unfolding_routines_mod.f90:1266-1286; BETA in the reference):
    components of the unfolded Pauli vector perpendicular / parallel to the k-point direction
    within the plane of normal `normal_to_proj_plane`.  Plain 3-vector algebra on the host, as in
    the reference.  Returns (spin_proj_perp, spin_proj_para)."""
    def versor(v):
        v = np.asarray(v, dtype=np.float64)
        return v / np.linalg.norm(v)
    o = np.zeros(3) if origin is None else np.asarray(origin, dtype=np.float64)
    para = versor(np.asarray(cart_coords_kpt, dtype=np.float64) - o)
    perp = versor(np.cross(versor(normal_to_proj_plane), para))
    pv = np.asarray(pauli_vector, dtype=np.float64)
    return float(np.dot(pv, perp)), float(np.dot(pv, para))


_PINNED = {}   # data address -> base pointer of live bandup_gpu_pinned_alloc blocks


def pinned_empty(nbytes: int) -> np.ndarray:
    """uint8 NumPy view of page-locked host memory (bandup_gpu_pinned_alloc); release it with
    pinned_free once no upload from it is in flight."""
    lib = _capi.load()
    p = C.c_void_p()
    rc = lib.bandup_gpu_pinned_alloc(int(nbytes), C.byref(p))
    if rc != 0:
        raise BandupGpuError(rc, lib.bandup_gpu_last_error(-1).decode())
    buf = (C.c_uint8 * int(nbytes)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)
    _PINNED[arr.ctypes.data] = p.value
    return arr


def pick_fastest(rates: Sequence[float], n_keep: int) -> List[int]:
    """Indices of the n_keep candidates to keep: anything within 2 % of the best counts as equally
    near and then the lower index wins (no churn on a box whose memory is all alike); the rest by rate."""
    fast = max(rates) * 0.98
    order = sorted(range(len(rates)), key=lambda i: (0, i) if rates[i] >= fast else (1, -rates[i]))
    return sorted(order[:n_keep])


def pinned_empty_near(nbytes: int, n_keep: int = 1, spares: int = 2, device: int = 0, slice_mb: int = 256):
    """n_keep page-locked buffers of nbytes each, chosen by MEASUREMENT among n_keep + spares
    candidates: the ones a host->device copy reads fastest.  Returns (buffers, GB/s of every
    candidate, indices kept).

    Why: the GPU boxes are virtual machines that show ONE NUMA node (numa_node = -1, `lscpu`: one
    socket) while their memory spans the host's sockets, so there is nothing to bind to -- yet a block
    that happens to lie behind the far socket uploads at 42-48 GB/s instead of the link's 55.5
    (same box, same GPU, one process in three: gpurun_out/r02x_bench*.json).  The copy engine can
    tell the blocks apart even though sysfs cannot.  torch only provides the device scratch, the
    stream and the events of the timing."""
    if int(nbytes) < 1:
        raise ValueError("pinned_empty_near needs nbytes >= 1")
    n_keep, spares = max(1, int(n_keep)), max(0, int(spares))
    cands = []
    try:
        for i in range(n_keep + spares):
            try:
                cands.append(pinned_empty(nbytes))
            except BandupGpuError:
                if i < n_keep:
                    raise
                break                # a spare the host has no room for: choose among what there is
        if len(cands) == n_keep:
            return cands, [None] * n_keep, list(range(n_keep))
        rates = _upload_rates(cands, int(nbytes), device, min(int(nbytes), max(1, int(slice_mb)) << 20))
    except BaseException:
        for h in cands:          # nothing is handed out: give the candidates back
            pinned_free(h)
        raise
    keep = pick_fastest(rates, n_keep)
    for i in range(len(cands)):
        if i not in keep:
            pinned_free(cands[i])
    return [cands[i] for i in keep], [round(r, 2) for r in rates], keep


def _upload_rates(cands, nbytes: int, device: int, step: int) -> List[float]:
    """GB/s of one host->device copy of every candidate block, in slices of `step` bytes."""
    import torch
    with torch.cuda.device(device):
        scratch = torch.empty(step, dtype=torch.uint8, device=f"cuda:{device}")
        stream = torch.cuda.Stream(device=device)
        rates = []
        for h in cands:
            src = torch.from_numpy(h)
            with torch.cuda.stream(stream):
                scratch.copy_(src[:step], non_blocking=True)          # warm the mapping
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for off in range(0, int(nbytes) - step + 1, step):
                    scratch.copy_(src[off:off + step], non_blocking=True)
                b.record()
            stream.synchronize()
            moved = ((int(nbytes) - step) // step + 1) * step
            rates.append(moved / (a.elapsed_time(b) * 1e-3) / 1e9)
        del scratch
    return rates


def host_register(arr: np.ndarray) -> None:
    """Page-lock a NumPy array the caller owns, in place (bandup_gpu_host_register), so that
    submits from it are asynchronous DMA; host_unregister(arr) before it is freed."""
    lib = _capi.load()
    if not arr.flags["C_CONTIGUOUS"]:
        raise ValueError("host_register needs a C-contiguous array")
    rc = lib.bandup_gpu_host_register(C.c_void_p(arr.ctypes.data), int(arr.nbytes))
    if rc != 0:
        raise BandupGpuError(rc, lib.bandup_gpu_last_error(-1).decode())


def host_unregister(arr: np.ndarray) -> None:
    lib = _capi.load()
    rc = lib.bandup_gpu_host_unregister(C.c_void_p(arr.ctypes.data))
    if rc != 0:
        raise BandupGpuError(rc, lib.bandup_gpu_last_error(-1).decode())


_POOL = []     # idle page-locked buffers kept for reuse (cudaMallocHost costs ~1 s per GB)


def pinned_pool_get(nbytes: int) -> np.ndarray:
    """A page-locked buffer of at least nbytes from the reuse pool (allocated on a miss)."""
    best = None
    for i, a in enumerate(_POOL):
        if a.size >= nbytes and (best is None or a.size < _POOL[best].size):
            best = i
    if best is not None:
        return _POOL.pop(best)
    return pinned_empty(nbytes)


def pinned_pool_put(arr: np.ndarray) -> None:
    """Hands a buffer back for reuse (no upload from it may be in flight)."""
    _POOL.append(arr)


def pinned_pool_release() -> None:
    """Frees every idle buffer of the pool."""
    while _POOL:
        pinned_free(_POOL.pop())


def pinned_free(arr: np.ndarray) -> None:
    """bandup_gpu_pinned_free for a buffer from pinned_empty (the array must not be used afterwards)."""
    base = _PINNED.pop(arr.ctypes.data, None)
    if base is not None:
        _capi.load().bandup_gpu_pinned_free(C.c_void_p(base))
