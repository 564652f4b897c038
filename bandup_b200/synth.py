"""Synthetic inputs for the unfolding hot path (lattices, k-points, G lists, coefficients).

The reference ships no wavefunction files (its tutorials need VASP/QE/ABINIT runs), so
tests and bench.py synthesise inputs shaped like BASELINE.json's configurations.  This
module is input plumbing only: it never computes weights or delta_N.

Geometry follows the reference's conventions:
  * reciprocal lattice rows B_i = 2 pi (A^-1)^T            (crystals_mod.f90:186-202)
  * A_i = sum_j M_ij a_j, so b_i = sum_j M_ji B_j          (crystals_mod.f90:45-47)
  * the plane-wave list of an SC k-point is enumerated exactly as the VASP reader
    regenerates it (read_wavecar_mod.f90:255-304): ig3 slowest, ig1 fastest, each
    index running 0..nbmax then -nbmax..-1, kept when |(K+G).B|^2 / c < ENCUT, with
    c the float32-rounded literal 0.262465831 (read_wavecar_mod.f90:107).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

VASP_2M_OVER_HBAR_SQRD = float(np.float32(0.262465831))


def rec_latt(latt: np.ndarray) -> np.ndarray:
    """Rows b_i = 2 pi (a_j x a_k) / |a_1 . (a_2 x a_3)|."""
    a = np.asarray(latt, dtype=np.float64)
    vol = abs(np.dot(a[0], np.cross(a[1], a[2])))
    return 2.0 * np.pi * np.array([np.cross(a[1], a[2]), np.cross(a[2], a[0]), np.cross(a[0], a[1])]) / vol


def _max_nb(B: np.ndarray, ecut: float, c: float) -> Tuple[int, int, int]:
    """Search box of the G enumeration (read_wavecar_mod.f90:151-210)."""
    b1, b2, b3 = B
    nrm = np.linalg.norm

    def ang(u, v):
        return np.arccos(np.dot(u, v) / (nrm(u) * nrm(v)))

    r = np.sqrt(ecut * c)
    p12, p13, p23 = ang(b1, b2), ang(b1, b3), ang(b2, b3)
    s = np.cos(ang(b3, np.cross(b1, b2)))
    A = (int(r / (nrm(b1) * abs(np.sin(p12))) + 1), int(r / (nrm(b2) * abs(np.sin(p12))) + 1),
         int(r / (nrm(b3) * abs(s)) + 1))
    s = np.cos(ang(b2, np.cross(b1, b3)))
    Bq = (int(r / (nrm(b1) * abs(np.sin(p13))) + 1), int(r / (nrm(b2) * abs(s)) + 1),
          int(r / (nrm(b3) * abs(np.sin(p13))) + 1))
    s = np.cos(ang(b1, np.cross(b2, b3)))
    Cq = (int(r / (nrm(b1) * abs(s)) + 1), int(r / (nrm(b2) * abs(np.sin(p23))) + 1),
          int(r / (nrm(b3) * abs(np.sin(p23))) + 1))
    return tuple(max(A[i], Bq[i], Cq[i]) for i in range(3))


def gen_gvecs(kfrac: np.ndarray, ecut: float, B: np.ndarray,
              c: float = VASP_2M_OVER_HBAR_SQRD) -> np.ndarray:
    """int32 (n_pw, 3) Miller triples of the plane waves of SC k-point `kfrac`, reader order."""
    nb = _max_nb(B, ecut, c)

    def seq(n):
        return np.concatenate([np.arange(0, n + 1), np.arange(-n, 0)])

    s1, s2, s3 = seq(nb[0]), seq(nb[1]), seq(nb[2])
    # ig3 slowest ... ig1 fastest
    g3, g2, g1 = np.meshgrid(s3, s2, s1, indexing="ij")
    g = np.stack([g1.ravel(), g2.ravel(), g3.ravel()], axis=1)
    kf = np.asarray(kfrac, dtype=np.float64)
    sk = ((kf[0] + g[:, 0])[:, None] * B[0] + (kf[1] + g[:, 1])[:, None] * B[1]) + \
         (kf[2] + g[:, 2])[:, None] * B[2]
    gtot = np.sqrt(sk[:, 0] ** 2 + sk[:, 1] ** 2 + sk[:, 2] ** 2)
    etot = gtot ** 2 / c
    return np.ascontiguousarray(g[etot < ecut], dtype=np.int32)


@dataclass
class Scenario:
    """Geometry of one unfolding job under the reference's no-symmetry semantics
    (-no_symm_sckpts / -no_symm_avg: Scoords = coords, folding_G = k - K)."""
    name: str
    a_pc: np.ndarray
    M: np.ndarray                       # int (3,3): A_SC = M . a_pc
    ecut: float
    n_bands: int
    n_spinor: int
    pc_kpts_cart: np.ndarray            # (n_k, 3)
    e_fermi: float = 0.0
    emin: float = -20.0
    emax: float = 5.0
    dE: float = 0.05
    A_sc: np.ndarray = field(init=False)
    b_pc: np.ndarray = field(init=False)
    B_sc: np.ndarray = field(init=False)
    sc_kpts_frac: np.ndarray = field(init=False)   # (n_K, 3) in B_sc
    sc_kpts_cart: np.ndarray = field(init=False)
    folds: List[List[int]] = field(init=False)     # per K: indices into pc_kpts_cart

    def __post_init__(self):
        self.a_pc = np.asarray(self.a_pc, dtype=np.float64)
        self.M = np.asarray(self.M, dtype=np.int64)
        self.A_sc = self.M.astype(np.float64) @ self.a_pc
        self.b_pc = rec_latt(self.a_pc)
        self.B_sc = rec_latt(self.A_sc)
        Binv = np.linalg.inv(self.B_sc)
        kf = self.pc_kpts_cart @ Binv                       # SC-fractional coords of every pc-kpt
        red = kf - np.floor(kf + 1e-9)                      # K = k mod B_sc, in [0,1)
        red[np.abs(red) < 1e-9] = 0.0
        keys = np.round(red, 7)
        uniq: Dict[tuple, int] = {}
        Ks: List[np.ndarray] = []
        folds: List[List[int]] = []
        for ik, key in enumerate(map(tuple, keys)):
            if key not in uniq:
                uniq[key] = len(Ks)
                Ks.append(red[ik])
                folds.append([])
            folds[uniq[key]].append(ik)
        self.sc_kpts_frac = np.array(Ks)
        self.sc_kpts_cart = self.sc_kpts_frac @ self.B_sc
        self.folds = folds

    @property
    def det_M(self) -> int:
        return int(round(abs(np.linalg.det(self.M))))

    @property
    def n_sc_kpts(self) -> int:
        return len(self.folds)

    def folding_G(self, iK: int) -> np.ndarray:
        """(n_fold, 3) Cartesian k - K for the pc-kpts folding onto SC-kpt iK."""
        return self.pc_kpts_cart[self.folds[iK]] - self.sc_kpts_cart[iK]

    def gvecs(self, iK: int) -> np.ndarray:
        return gen_gvecs(self.sc_kpts_frac[iK], self.ecut, self.B_sc)

    def energy_window(self) -> Tuple[float, float, float]:
        """E_start, E_end, delta_e as read_energy_info_for_band_search builds them
        (io_routines_mod.f90:361-371)."""
        return self.emin + self.e_fermi, self.emax + self.e_fermi, abs(self.dE)


def kpath(points_frac: List[List[float]], b: np.ndarray, n_per_segment: int) -> np.ndarray:
    """Straight segments between pc-fractional points, endpoints included once."""
    pts = [np.asarray(p, dtype=np.float64) @ b for p in points_frac]
    out = []
    for i in range(len(pts) - 1):
        t = np.linspace(0.0, 1.0, n_per_segment, endpoint=(i == len(pts) - 2))
        out.append(pts[i][None, :] + t[:, None] * (pts[i + 1] - pts[i])[None, :])
    return np.concatenate(out, axis=0)


def hex_cell(a: float, c: float) -> np.ndarray:
    return np.array([[np.sqrt(3) / 2 * a, -0.5 * a, 0.0], [np.sqrt(3) / 2 * a, 0.5 * a, 0.0], [0.0, 0.0, c]])


def fcc_cell(a: float) -> np.ndarray:
    return 0.5 * a * np.array([[0.0, 1.0, 1.0], [1.0, 0.0, 1.0], [1.0, 1.0, 0.0]])


FCC_TO_CUBIC = np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1]])


def make_scenario(name: str, scale: float = 1.0, n_kpts: Optional[int] = None) -> Scenario:
    """BASELINE.json configurations.  `scale` < 1 shrinks ENCUT and the band count so the
    oracle finishes in seconds; geometry (M, cosets, k-path) is unchanged."""
    if name == "graphene4x4":          # config 1
        a = hex_cell(2.467, 15.0)
        b = rec_latt(a)
        nk = n_kpts or 30
        k = kpath([[1 / 3, 1 / 3, 0], [0, 0, 0], [0.5, 0, 0], [1 / 3, 1 / 3, 0]], b, max(2, nk // 3))
        return Scenario(name, a, np.diag([4, 4, 1]), 400.0 * scale, max(4, int(64 * scale)), 1, k)
    if name == "si333_vacancy":        # configs 2 and 4
        a = fcc_cell(5.43)
        b = rec_latt(a)
        nk = n_kpts or 100
        k = kpath([[0.5, 0.5, 0.5], [0, 0, 0], [0.5, 0, 0.5]], b, max(2, nk // 2))
        return Scenario(name, a, 3 * FCC_TO_CUBIC, 160.0 * scale, max(4, int(500 * scale)), 1, k)
    if name == "mos2_8x8_spinor":      # config 3
        a = hex_cell(3.16, 18.0)
        b = rec_latt(a)
        nk = n_kpts or 50
        k = kpath([[0, 0, 0], [0.5, 0, 0], [1 / 3, 1 / 3, 0], [0, 0, 0]], b, max(2, nk // 3))
        return Scenario(name, a, np.diag([8, 8, 1]), 230.0 * scale, max(4, int(1200 * scale)), 2, k)
    if name == "sige555":              # config 5
        a = fcc_cell(5.54)
        b = rec_latt(a)
        nk = n_kpts or 200
        k = kpath([[0.5, 0.5, 0.5], [0, 0, 0], [0.5, 0, 0.5]], b, max(2, nk // 2))
        return Scenario(name, a, 5 * FCC_TO_CUBIC, 163.0 * scale, max(4, int(3000 * scale)), 1, k)
    raise KeyError(name)


def coset_labels(g_frac: np.ndarray, M: np.ndarray) -> np.ndarray:
    """A small integer label per plane wave, equal for waves in the same coset of the pc
    reciprocal lattice (used only to bias the synthetic coefficients)."""
    Mt = M.T.astype(np.int64)
    adj = np.round(np.linalg.inv(Mt) * np.linalg.det(Mt)).astype(np.int64)
    D = int(round(abs(np.linalg.det(Mt))))
    key = (g_frac.astype(np.int64) @ adj) % D
    _, lab = np.unique(key, axis=0, return_inverse=True)
    return lab.reshape(-1)


def make_wavefunction(sc: Scenario, iK: int, seed: int = 20261019):
    """Coefficients (n_bands, n_pw, n_spinor) complex64, G_frac, band energies for SC-kpt iK.

    C = (N(0,1) + i N(0,1)) * exp(-|K+G|^2 / (2 c E_w)), E_w = ENCUT/4, times a per-band
    Dirichlet(0.3) coset bias so the weights are far from uniform; NOT normalised, so the
    renormalisation is exercised.  Energies: sorted uniform in [E_F-20, E_F+6] with exact
    duplicates and values placed exactly on bin edges / centres (tests lambda = 1/2)."""
    rng = np.random.default_rng(seed + 1000 * iK)
    g = sc.gvecs(iK)
    n_pw = g.shape[0]
    kg = (sc.sc_kpts_frac[iK][None, :] + g) @ sc.B_sc
    env = np.exp(-np.sum(kg * kg, axis=1) / (2.0 * VASP_2M_OVER_HBAR_SQRD * sc.ecut / 4.0))
    lab = coset_labels(g, sc.M)
    ncos = int(lab.max()) + 1
    bias = rng.dirichlet(np.full(ncos, 0.3), size=sc.n_bands)           # (nb, ncos)
    amp = np.sqrt(bias[:, lab] * ncos)                                  # (nb, n_pw)
    z = rng.standard_normal((sc.n_bands, n_pw, sc.n_spinor, 2)).astype(np.float32)
    coeffs = (z[..., 0] + 1j * z[..., 1]).astype(np.complex64)
    coeffs *= (amp * env[None, :])[:, :, None].astype(np.float32)
    scale = rng.uniform(0.3, 3.0, size=sc.n_bands).astype(np.float32)   # un-normalise
    coeffs *= scale[:, None, None]
    E0, E1, dE = sc.energy_window()
    ener = np.sort(rng.uniform(sc.e_fermi - 20.0, sc.e_fermi + 6.0, size=sc.n_bands))
    if sc.n_bands >= 8:
        ener[1] = ener[0]                                               # exact duplicate
        ener[3] = E0 + 7 * dE + 0.5 * dE                                # on a bin edge
        ener[5] = E0 + 11 * dE                                          # on a bin centre
        ener[6] = (E0 + 12 * dE) - 0.5 * (dE)                           # edge from above
        ener = np.sort(ener)
    return np.ascontiguousarray(coeffs), g, np.ascontiguousarray(ener)
