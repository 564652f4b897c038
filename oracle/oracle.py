"""ctypes front end of the C oracle (oracle/bandup_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by bandup_b200.
Parity status: pinned against the reference's Fortran as executed by oracle/f90interp.py
(tests/golden/fortran_golden.npz, tests/test_fortran_golden.py); see the header of bandup_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libbandup_oracle.so"

_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "bandup_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True, env=env,
                       capture_output=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.bo_lambda.restype = C.c_double
        L.bo_lambda.argtypes = [C.c_double] * 4
        L.bo_integral_delta.restype = C.c_double
        L.bo_integral_delta.argtypes = [C.c_double] * 4
        L.bo_real_seq.restype = C.c_int64
        L.bo_real_seq.argtypes = [C.c_double, C.c_double, C.c_double, _f64p]
        L.bo_select_coeffs.restype = C.c_int64
        L.bo_select_coeffs.argtypes = [_f64p, C.c_int64, _f64p, _f64p, _i32p, _f64p]
        L.bo_select_coeffs_int.restype = C.c_int64
        L.bo_select_coeffs_int.argtypes = [_i32p, C.c_int64, _i32p, _i32p, _i32p]
        L.bo_gen_gvecs.restype = C.c_int64
        L.bo_gen_gvecs.argtypes = [_f64p, C.c_double, _f64p, C.c_double, _i32p, _f64p, C.c_int64]
        L.bo_renormalize.restype = None
        L.bo_renormalize.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, _f64p]
        L.bo_spectral_weights.restype = None
        L.bo_spectral_weights.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, _i32p, C.c_int64, _f64p]
        L.bo_delta_N.restype = None
        L.bo_delta_N.argtypes = [_f64p, C.c_int64, _f64p, _f64p, C.c_int, C.c_double, _f64p]
        L.bo_spectral_function.restype = None
        L.bo_spectral_function.argtypes = [_f64p, C.c_int64, _f64p, _f64p, C.c_int, C.c_double, _f64p]
        L.bo_commensurate.restype = C.c_int
        L.bo_commensurate.argtypes = [_f64p, _f64p, _f64p, _i32p, C.c_double]
        L.bo_vec_in_latt.restype = C.c_int
        L.bo_vec_in_latt.argtypes = [_f64p, _f64p, C.c_double]
        L.bo_rec_latt.restype = None
        L.bo_rec_latt.argtypes = [_f64p, _f64p]
        L.bo_calc_rho.restype = C.c_int
        L.bo_calc_rho.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, _i32p, C.c_int64, _f64p,
                                  C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.bo_unfold_kpt.restype = C.c_int
        L.bo_unfold_kpt.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, _f64p, _f64p, _f64p, _f64p,
                                    C.c_int, _f64p, C.c_int64, C.c_double, C.c_int, _f64p, _f64p, _i32p,
                                    _f64p]
        L.bo_unfold_kpt_dp.restype = C.c_int
        L.bo_unfold_kpt_dp.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, _f64p, _f64p, _f64p, _f64p,
                                       C.c_int, _f64p, C.c_int64, C.c_double, _f64p, _f64p, _i32p, _f64p]
        L.bo_synth_fill.restype = None
        L.bo_synth_fill.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64]
        L.bo_synth_fill_rows.restype = None
        L.bo_synth_fill_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        L.bo_symm_average.restype = None
        L.bo_symm_average.argtypes = [_f64p, C.c_int, C.c_int64, _f64p, _f64p]
        _f32p = C.POINTER(C.c_float)
        L.bo_symm_average_rho.restype = C.c_int64
        L.bo_symm_average_rho.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _f64p, C.c_double, _i64p] + \
            [_i32p] * 4 + [_f32p] * 2 + [_i32p] * 4 + [_f32p] * 2
        L.bo_num_threads.restype = C.c_int
        L.bo_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def num_threads() -> int:
    return int(lib().bo_num_threads())


def set_num_threads(n: int) -> None:
    lib().bo_set_num_threads(int(n))


def rec_latt(latt):
    a = _f64(latt)
    out = np.zeros((3, 3))
    lib().bo_rec_latt(_p(a, _f64p), _p(out, _f64p))
    return out


def commensurate(b_pc, B_sc, tol=0.0):
    b, B = _f64(b_pc), _f64(B_sc)
    M = np.zeros((3, 3))
    iM = np.zeros((3, 3), dtype=np.int32)
    ok = lib().bo_commensurate(_p(b, _f64p), _p(B, _f64p), _p(M, _f64p), _p(iM, _i32p), float(tol))
    return bool(ok), M, iM


def vec_in_latt(vec, latt, tol=1e-5):
    v, l = _f64(vec), _f64(latt)
    return bool(lib().bo_vec_in_latt(_p(v, _f64p), _p(l, _f64p), float(tol)))


def real_seq(first, last, incr):
    n = lib().bo_real_seq(float(first), float(last), float(incr), None)
    out = np.zeros(n)
    lib().bo_real_seq(float(first), float(last), float(incr), _p(out, _f64p))
    return out


def gen_gvecs(kfrac, ecut, B, c=0.0):
    k, Bm = _f64(kfrac), _f64(B)
    n = lib().bo_gen_gvecs(_p(k, _f64p), float(ecut), _p(Bm, _f64p), float(c), None, None, 0)
    gf = np.zeros((n, 3), dtype=np.int32)
    gc = np.zeros((n, 3))
    lib().bo_gen_gvecs(_p(k, _f64p), float(ecut), _p(Bm, _f64p), float(c), _p(gf, _i32p), _p(gc, _f64p), n)
    return gf, gc


def select_coeffs(G_cart, folding_G, b_pc):
    """1-based ascending indices and the tolerance the escalation loop ended on."""
    g, fg, b = _f64(G_cart), _f64(folding_G), _f64(b_pc)
    sel = np.zeros(g.shape[0], dtype=np.int32)
    tol = C.c_double(0.0)
    n = lib().bo_select_coeffs(_p(g, _f64p), g.shape[0], _p(fg, _f64p), _p(b, _f64p), _p(sel, _i32p),
                               C.byref(tol))
    return sel[:n].copy(), tol.value


def select_coeffs_int(g_frac, g0, M):
    g = np.ascontiguousarray(g_frac, dtype=np.int32)
    g0 = np.ascontiguousarray(g0, dtype=np.int32)
    Mi = np.ascontiguousarray(M, dtype=np.int32)
    sel = np.zeros(g.shape[0], dtype=np.int32)
    n = lib().bo_select_coeffs_int(_p(g, _i32p), g.shape[0], _p(g0, _i32p), _p(Mi, _i32p), _p(sel, _i32p))
    return sel[:n].copy()


def renormalize(coeffs, norm_mode=1):
    """Returns (renormalised copy, norms). coeffs: (n_bands, n_pw, n_spinor) complex64."""
    c = np.array(coeffs, dtype=np.complex64, order="C", copy=True)
    nb, npw, nsp = c.shape
    norms = np.zeros(nb)
    lib().bo_renormalize(c.ctypes.data, nsp, npw, nb, int(norm_mode), _p(norms, _f64p))
    return c, norms


def spectral_weights(coeffs, sel):
    c = np.ascontiguousarray(coeffs, dtype=np.complex64)
    nb, npw, nsp = c.shape
    s = np.ascontiguousarray(sel, dtype=np.int32)
    w = np.zeros(nb)
    lib().bo_spectral_weights(c.ctypes.data, nsp, npw, nb, _p(s, _i32p), s.shape[0], _p(w, _f64p))
    return w


def delta_N(energies, sc_ener, weights, smearing=0.0):
    e, s, w = _f64(energies), _f64(sc_ener), _f64(weights)
    dN = np.zeros(e.shape[0])
    lib().bo_delta_N(_p(e, _f64p), e.shape[0], _p(s, _f64p), _p(w, _f64p), s.shape[0], float(smearing),
                     _p(dN, _f64p))
    return dN


def spectral_function(energies, sc_ener, weights, std_dev=0.0):
    e, s, w = _f64(energies), _f64(sc_ener), _f64(weights)
    sf = np.zeros(e.shape[0])
    lib().bo_spectral_function(_p(e, _f64p), e.shape[0], _p(s, _f64p), _p(w, _f64p), s.shape[0], float(std_dev),
                               _p(sf, _f64p))
    return sf


def lam(pc_ener, sc_ener, delta_e, sigma):
    return float(lib().bo_lambda(float(pc_ener), float(sc_ener), float(delta_e), float(sigma)))


def calc_rho(coeffs, sel, sc_ener, delta_N_val, pc_ener, delta_e, sigma=0.0):
    c = np.ascontiguousarray(coeffs, dtype=np.complex64)
    nb, npw, nsp = c.shape
    s = np.ascontiguousarray(sel, dtype=np.int32)
    e = _f64(sc_ener)
    rho = np.zeros((nb, nb), dtype=np.complex64)
    rc = lib().bo_calc_rho(c.ctypes.data, nsp, npw, nb, _p(s, _i32p), s.shape[0], _p(e, _f64p),
                           float(delta_N_val), float(pc_ener), float(delta_e), float(sigma), rho.ctypes.data)
    if rc:
        raise ZeroDivisionError("calc_rho: delta_N = 0")
    return rho


def calc_rho_dont_unfold(sc_ener, delta_N_val, pc_ener, delta_e, sigma=0.0):
    """The -dont_unfold branch of calc_rho (band_unfolding_routines_mod.f90:1136-1154)."""
    e = _f64(sc_ener)
    nb = e.shape[0]
    rho = np.zeros((nb, nb), dtype=np.complex64)
    L = lib()
    L.bo_calc_rho_dont_unfold.restype = C.c_int
    L.bo_calc_rho_dont_unfold.argtypes = [C.c_int, _f64p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
    if L.bo_calc_rho_dont_unfold(nb, _p(e, _f64p), float(delta_N_val), float(pc_ener), float(delta_e), float(sigma),
                                 rho.ctypes.data):
        raise ZeroDivisionError("calc_rho: delta_N = 0")
    return rho


def unfold_kpt(coeffs, G_cart, sc_ener, b_pc, folding_G, energies, smearing=0.0, norm_mode=1,
               in_place=False):
    """The reference's per-SC-kpt sequence.  Returns weights, dN, n_selected, times[3]."""
    c = coeffs if in_place else np.array(coeffs, dtype=np.complex64, order="C", copy=True)
    nb, npw, nsp = c.shape
    g, e, b = _f64(G_cart), _f64(sc_ener), _f64(b_pc)
    fg = _f64(np.atleast_2d(folding_G))
    en = _f64(energies)
    nf = fg.shape[0]
    w = np.zeros((nf, nb))
    dN = np.zeros((nf, en.shape[0]))
    ns = np.zeros(nf, dtype=np.int32)
    times = np.zeros(3)
    lib().bo_unfold_kpt(c.ctypes.data, nsp, npw, nb, _p(g, _f64p), _p(e, _f64p), _p(b, _f64p), _p(fg, _f64p),
                        nf, _p(en, _f64p), en.shape[0], float(smearing), int(norm_mode), _p(w, _f64p),
                        _p(dN, _f64p), _p(ns, _i32p), _p(times, _f64p))
    return w, dN, ns, times


def unfold_kpt_dp(coeffs, G_cart, sc_ener, b_pc, folding_G, energies, smearing=0.0):
    """The same sequence as a `kind_cplx_coeffs = dp` build of the reference runs it
    (constants_and_types_mod.f90:60): complex128 coefficients, double-precision dot_product, abs and
    rescale.  Returns weights, dN, n_selected, norms (before renormalisation)."""
    c = np.array(coeffs, dtype=np.complex128, order="C", copy=True)
    nb, npw, nsp = c.shape
    g, e, b = _f64(G_cart), _f64(sc_ener), _f64(b_pc)
    fg = _f64(np.atleast_2d(folding_G))
    en = _f64(energies)
    nf = fg.shape[0]
    w = np.zeros((nf, nb))
    dN = np.zeros((nf, en.shape[0]))
    ns = np.zeros(nf, dtype=np.int32)
    norms = np.zeros(nb)
    lib().bo_unfold_kpt_dp(c.ctypes.data, nsp, npw, nb, _p(g, _f64p), _p(e, _f64p), _p(b, _f64p), _p(fg, _f64p),
                           nf, _p(en, _f64p), en.shape[0], float(smearing), _p(w, _f64p), _p(dN, _f64p),
                           _p(ns, _i32p), _p(norms, _f64p))
    return w, dN, ns, norms


def synth_fill(row_elems, n_bands, seed, ikpt, band0=0, out=None):
    """Rows band0..band0+n_bands-1 of the counter-based synthetic coefficient block."""
    c = np.zeros((n_bands, row_elems), dtype=np.complex64) if out is None else out
    lib().bo_synth_fill_rows(c.ctypes.data, int(row_elems), int(band0), int(n_bands), C.c_uint64(seed),
                             C.c_uint64(ikpt))
    return c


def symm_average(weights, per_dir):
    """out = sum_d weights[d] * per_dir[d], accumulated in direction order (dp)."""
    w = _f64(weights)
    x = _f64(per_dir)
    out = np.zeros(x.shape[1:])
    lib().bo_symm_average(_p(w, _f64p), w.shape[0], out.size, _p(x, _f64p), _p(out, _f64p))
    return out


def symm_average_rho(weights, dir_off, pckpt, iener, m1, m2, rho, avg_dN, n_bands, min_dN=None):
    """Symmetry-averaged unfolding-density operators; see bo_symm_average_rho.  rho complex64.
    Returns (pckpt, iener, m1, m2, rho) arrays of the averaged entries."""
    if min_dN is None:
        min_dN = float(np.float32(1e-3))
    w = _f64(weights)
    off = np.ascontiguousarray(dir_off, dtype=np.int64)
    ints = [np.ascontiguousarray(a, dtype=np.int32) for a in (pckpt, iener, m1, m2)]
    r = np.ascontiguousarray(rho, dtype=np.complex64)
    re, im = np.ascontiguousarray(r.real), np.ascontiguousarray(r.imag)
    dn = _f64(avg_dN)
    n = int(off[-1])
    outs = [np.zeros(max(n, 1), dtype=np.int32) for _ in range(4)]
    ore, oim = np.zeros(max(n, 1), dtype=np.float32), np.zeros(max(n, 1), dtype=np.float32)
    f32p = C.POINTER(C.c_float)
    m = lib().bo_symm_average_rho(_p(w, _f64p), w.shape[0], dn.shape[1], int(n_bands), _p(dn, _f64p),
                                  float(min_dN), _p(off, _i64p), *[_p(a, _i32p) for a in ints],
                                  _p(re, f32p), _p(im, f32p), *[_p(a, _i32p) for a in outs],
                                  _p(ore, f32p), _p(oim, f32p))
    return tuple(a[:m].copy() for a in outs) + ((ore[:m] + 1j * oim[:m]).astype(np.complex64),)
