"""VASP WAVECAR container: the on-disk layout the reference's reader consumes
(src/vasp_related_modules/read_wavecar_mod.f90), parsed so that the raw band records can be handed
to the GPU untouched (layout BANDUP_GPU_LAYOUT_VASP_RECORD, stride nrecl) -- no host de-interleave.

Direct-access records of `nrecl` bytes (read_wavecar_mod.f90:112-127):
  rec 1   nrecl, nspin, nprec                       3 doubles; nprec = 45200 <=> complex64 (:356-377)
  rec 2   nkpts, nbands, ENCUT, a1(3), a2(3), a3(3) 12 doubles                              (:386)
  then, per spin, per k-point:
    1 header record  nplane, k(3), {Re E, Im E, occ} x nbands   doubles                    (:420-427)
    nbands records   complex64[nplane]; for a spinor nplane = 2 n_G, spin-up block first   (:446-452, :570-577)
each zero-padded to nrecl.  The plane-wave list is NOT stored: it is regenerated from k, ENCUT and
the lattice (:255-304), with a small adaptive tweak of 2m/hbar^2 when the count is off (:459-535).

This module is input plumbing (and the writer the tests use to synthesise files); it computes no
weights.  The reference's Fortran reader stays what BandUP itself uses (north_star: readers kept).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import synth

NPREC_COMPLEX64 = 45200
NPREC_COMPLEX128 = 45210      # WAVECAR_double: needs a `kind_cplx_coeffs = dp` build of the reference (:363-377)
READ_THREADS = int(os.environ.get("BANDUP_GPU_READ_THREADS", str(min(16, os.cpu_count() or 8))))


_POOL_EX = None


def _read_pool() -> ThreadPoolExecutor:
    global _POOL_EX
    if _POOL_EX is None:
        _POOL_EX = ThreadPoolExecutor(READ_THREADS, thread_name_prefix="wavecar-read")
    return _POOL_EX


def _pread_full(fd: int, mv: memoryview, offset: int) -> int:
    """preadv until `mv` is full (or EOF); returns the number of bytes read."""
    done = 0
    while done < len(mv):
        k = os.preadv(fd, [mv[done:]], offset + done)
        if k <= 0:
            break
        done += k
    return done


@dataclass
class KptHeader:
    nplane: int                 # as stored: n_G * n_spinor
    kpt_frac_coords: np.ndarray
    band_energies: np.ndarray   # Re(cener)
    band_occupations: np.ndarray
    first_band_record: int      # 0-based record index of band 1


def write_wavecar(path, A: np.ndarray, ecut: float, kpts_frac: Sequence[np.ndarray],
                  coeffs: Sequence[np.ndarray], energies: Sequence[np.ndarray],
                  occupations: Optional[Sequence[np.ndarray]] = None, nspin: int = 1,
                  extra_pad: int = 0, coeffs_spin2: Optional[Sequence[np.ndarray]] = None,
                  energies_spin2: Optional[Sequence[np.ndarray]] = None, double: bool = False) -> int:
    """Writes a WAVECAR (single precision; `double`: complex128 records, nprec = 45210).  coeffs[ik] is
    (n_bands, n_pw, n_spinor) in the plane-wave order of synth.gen_gvecs (== the reader's enumeration).
    With nspin = 2 the second spin channel repeats the first unless coeffs_spin2 / energies_spin2 are
    given.  Returns nrecl."""
    A = np.asarray(A, dtype=np.float64)
    nk = len(kpts_frac)
    nb = coeffs[0].shape[0]
    max_plane = max(c.shape[1] * c.shape[2] for c in coeffs)
    esz, ctype = (16, np.complex128) if double else (8, np.complex64)
    nrecl = max(esz * max_plane, 8 * (4 + 3 * nb), 8 * 12) + esz * int(extra_pad)
    with open(path, "wb") as f:
        def put(rec_bytes: bytes):
            assert len(rec_bytes) <= nrecl
            f.write(rec_bytes + b"\0" * (nrecl - len(rec_bytes)))

        put(np.array([nrecl, nspin, NPREC_COMPLEX128 if double else NPREC_COMPLEX64], dtype="<f8").tobytes())
        put(np.concatenate([[nk, nb, ecut], A.reshape(9)]).astype("<f8").tobytes())
        for _spin in range(nspin):
            c_src = coeffs_spin2 if (_spin == 1 and coeffs_spin2 is not None) else coeffs
            e_src = energies_spin2 if (_spin == 1 and energies_spin2 is not None) else energies
            for ik in range(nk):
                c = np.ascontiguousarray(c_src[ik], dtype=ctype)
                if c.shape[0] != nb:
                    raise ValueError("every k-point must have the same number of bands")
                n_pw, nsp = c.shape[1], c.shape[2]
                occ = np.ones(nb) if occupations is None else np.asarray(occupations[ik], dtype=np.float64)
                ener = np.asarray(e_src[ik], dtype=np.float64)
                hdr = np.empty(4 + 3 * nb, dtype="<f8")
                hdr[0] = n_pw * nsp
                hdr[1:4] = np.asarray(kpts_frac[ik], dtype=np.float64)
                hdr[4::3] = ener
                hdr[5::3] = 0.0
                hdr[6::3] = occ
                put(hdr.tobytes())
                for ib in range(nb):
                    # spinor-up block, then spinor-down block
                    put(np.ascontiguousarray(c[ib].T).astype("<c16" if double else "<c8").tobytes())
    return nrecl


class Wavecar:
    """Read-only view of a WAVECAR file."""

    def __init__(self, path):
        self.path = Path(path)
        with open(self.path, "rb") as f:
            head = np.frombuffer(f.read(24), dtype="<f8")
            self.nrecl = int(round(head[0]))
            self.nspin = int(round(head[1]))
            self.nprec = int(round(head[2]))
            if self.nprec not in (NPREC_COMPLEX64, NPREC_COMPLEX128):
                raise ValueError(f"unexpected precision tag {self.nprec} in the WAVECAR header")
            # read_wavecar_mod.f90:360-377: a WAVECAR_double needs a dp build of the reference; here the
            # unfolder is switched to complex128 coefficients (GpuUnfolder.set_complex128) by unfold_wavecar
            self.double = self.nprec == NPREC_COMPLEX128
            self.elem_bytes = 16 if self.double else 8
            if self.nrecl < 96 or self.nrecl % 8:
                raise ValueError(f"unexpected record length {self.nrecl}")
            f.seek(self.nrecl)
            rec2 = np.frombuffer(f.read(96), dtype="<f8")
        self.nkpts = int(round(rec2[0]))
        self.nbands = int(round(rec2[1]))
        self.encut = float(rec2[2])
        if self.nkpts == 0:
            raise ValueError("ERROR (read_wavecar): Unexpected file format")
        self.A_matrix = rec2[3:12].reshape(3, 3).copy()
        self.B_matrix = synth.rec_latt(self.A_matrix)
        self._gcache = {}
        self._hcache = {}

    def header_record(self, ikpt: int, ispin: int = 1) -> int:
        """0-based record index of the k-point header (ikpt, ispin 1-based) (:409-414)."""
        if not (1 <= ikpt <= self.nkpts) or not (1 <= ispin <= self.nspin):
            raise IndexError("ikpt / ispin out of range")
        return 2 + (ispin - 1) * self.nkpts * (1 + self.nbands) + (ikpt - 1) * (1 + self.nbands)

    def kpt_header(self, ikpt: int, ispin: int = 1) -> KptHeader:
        key = (ikpt, ispin)
        if key not in self._hcache:
            rec = self.header_record(ikpt, ispin)
            with open(self.path, "rb") as f:
                f.seek(rec * self.nrecl)
                h = np.frombuffer(f.read(8 * (4 + 3 * self.nbands)), dtype="<f8")
            self._hcache[key] = KptHeader(int(round(h[0])), h[1:4].copy(), h[4::3].copy(), h[6::3].copy(), rec + 1)
        return self._hcache[key]

    def gvecs(self, ikpt: int, ispin: int = 1) -> Tuple[np.ndarray, int]:
        """(G_frac int32 (n_G, 3), n_spinor): the reader's regeneration incl. its adaptive tweak of
        2m/hbar^2 (read_wavecar_mod.f90:440-535)."""
        key = (ikpt, ispin)
        if key in self._gcache:
            return self._gcache[key]
        h = self.kpt_header(ikpt, ispin)
        c = synth.VASP_2M_OVER_HBAR_SQRD
        g = synth.gen_gvecs(h.kpt_frac_coords, self.encut, self.B_matrix, c)
        nplane = h.nplane
        n_spinor = 1
        if g.shape[0] and int(round(nplane / g.shape[0])) == 2:
            n_spinor = 2
            nplane //= 2
        step = 1e-10 if nplane >= g.shape[0] else -1e-10
        i = 0
        while g.shape[0] != nplane and abs(i * step) <= 1000 * 1e-10:
            i += 1
            g = synth.gen_gvecs(h.kpt_frac_coords, self.encut, self.B_matrix, c + i * step)
        if g.shape[0] != nplane:
            # :519-531 'ERROR (read_wavecar): The number of plane-waves ... does not match'
            raise ValueError(f"WAVECAR holds {nplane} plane waves, {g.shape[0]} were generated")
        self._gcache[key] = (g, n_spinor)
        return g, n_spinor

    def band_block_bytes(self) -> int:
        return self.nbands * self.nrecl

    def read_band_records(self, ikpt: int, ispin: int = 1, out: Optional[np.ndarray] = None) -> np.ndarray:
        """The nbands raw coefficient records of a k-point, contiguous on disk, as uint8
        [nbands * nrecl] -- read straight into `out` (e.g. page-locked memory from
        bandup_b200.unfold.pinned_empty) so the upload is one DMA."""
        n = self.band_block_bytes()
        if out is None:
            out = np.empty(n, dtype=np.uint8)
        if out.dtype != np.uint8 or out.size < n:
            raise ValueError("out must be a uint8 buffer of at least nbands*nrecl bytes")
        first = self.header_record(ikpt, ispin) + 1
        base = first * self.nrecl
        mv = memoryview(out[:n])
        fd = os.open(self.path, os.O_RDONLY)
        try:
            # big blocks (GBs per k-point) are read by several threads: preadv releases the GIL, and
            # one thread copying out of the page cache is far slower than the PCIe upload that follows
            n_thr = min(READ_THREADS, max(1, n // (4 << 20)))
            if n_thr <= 1:
                got = _pread_full(fd, mv, base)
            else:
                step = -(-n // n_thr)
                step += (-step) % 4096
                parts = [(lo, min(lo + step, n)) for lo in range(0, n, step)]
                got = sum(_read_pool().map(lambda p: _pread_full(fd, mv[p[0]:p[1]], base + p[0]), parts))
        finally:
            os.close(fd)
        if got != n:
            raise IOError(f"short read: {got} of {n} bytes")
        return out[:n]

    def coefficients(self, ikpt: int, ispin: int = 1) -> np.ndarray:
        """(n_bands, n_pw, n_spinor) complex64 as read_wavecar leaves wf%pw_coeffs -- a HOST
        de-interleave used only by tests; the GPU path consumes read_band_records directly."""
        g, nsp = self.gvecs(ikpt, ispin)
        raw = self.read_band_records(ikpt, ispin).reshape(self.nbands, self.nrecl)
        c = raw[:, :self.elem_bytes * g.shape[0] * nsp].copy().view("<c16" if self.double else "<c8").reshape(
            self.nbands, nsp, g.shape[0])
        return np.ascontiguousarray(c.transpose(0, 2, 1))

    def submit_to(self, unfolder, job_ikpt: int, ikpt: int, folding_G: np.ndarray, ispin: int = 1,
                  out: Optional[np.ndarray] = None, raw: Optional[np.ndarray] = None):
        """Header + raw records of file k-point `ikpt` (1-based) straight to the device
        (GpuUnfolder.submit_vasp): no plane-wave enumeration, no de-interleave on the host.
        `raw`: the records if they have been read already (read_band_records)."""
        h = self.kpt_header(ikpt, ispin)
        if raw is None:
            raw = self.read_band_records(ikpt, ispin, out=out)
        return unfolder.submit_vasp(job_ikpt, raw, self.nrecl, h.nplane, self.nbands, h.kpt_frac_coords, self.encut,
                                    h.band_energies, folding_G=folding_G)

    def wavefunction(self, ikpt: int, ispin: int = 1, out: Optional[np.ndarray] = None):
        """bandup_b200.unfold.PwWavefunction over the raw records (VASP_RECORD layout)."""
        from .unfold import LAYOUT_VASP_RECORD, PwWavefunction
        h = self.kpt_header(ikpt, ispin)
        g, nsp = self.gvecs(ikpt, ispin)
        raw = self.read_band_records(ikpt, ispin, out=out)
        return PwWavefunction(raw, g, h.band_energies, kpt_frac_coords=h.kpt_frac_coords, n_pw=g.shape[0],
                              n_bands=self.nbands, n_spinor=nsp, layout=LAYOUT_VASP_RECORD,
                              band_stride_bytes=self.nrecl)


def fold_pckpts_onto_sckpts(pc_kpts_cart: np.ndarray, sc_kpts_frac: np.ndarray, B_sc: np.ndarray,
                            tol: float = 1e-5) -> List[List[int]]:
    """For every SC-kpt of the file, the pc-kpts k with k - K on the SC reciprocal lattice (the
    geometric unfolding relation under -no_symm_sckpts, band_unfolding_routines_mod.f90:192-204,
    317-324).  Raises when a pc-kpt folds onto none (main_BandUP.f90:114)."""
    Binv = np.linalg.inv(B_sc)
    K = np.asarray(sc_kpts_frac, dtype=np.float64) @ B_sc
    folds: List[List[int]] = [[] for _ in range(K.shape[0])]
    for ik, k in enumerate(np.asarray(pc_kpts_cart, dtype=np.float64)):
        frac = (k[None, :] - K) @ Binv
        resid = np.abs((frac - np.rint(frac)) @ B_sc).max(axis=1)
        hit = np.nonzero(resid <= tol)[0]
        if hit.size == 0:
            raise ValueError(f"pc-kpt {ik} does not fold onto any SC-kpt of the wavefunction file")
        folds[int(hit[0])].append(ik)
    return folds


def unfold_wavecar(unfolder, wavecar: Wavecar, rec_latt_pc: np.ndarray, pc_kpts_cart: np.ndarray,
                   E_start: float, E_end: float, delta_e: float, ispin: int = 1, rank: int = 0,
                   world_size: int = 1, device=None, group=None, after_fetch=None, device_gvecs: bool = True,
                   n_sckpts_to_skip: int = 0, reader: str = "mmap"):
    """The reference's main loop (main_BandUP.f90:131-177) over a WAVECAR under its no-symmetry
    flags (-no_symm_sckpts -no_symm_avg): every SC-kpt of the file, all pc-kpts folding onto it in
    one pass, the raw band records uploaded as they are while the previous k-point computes.
    reader = "mmap" (default): the file is mapped and the mapping handed to the library as it is -- its
    staging ring copies 16 MB chunks out of the page cache while the previous chunks are on the wire,
    so a k-point's read and its upload overlap each other (C5: 52 GB/s of file data against 38-39 for
    "pread"); "pread": the records of k-point j+1 are read (threaded preadv) into a page-locked
    buffer by a helper thread while j uploads.
    Returns sharding.GatheredDeltaN (rows in pc-kpt order) and the energy grid."""
    from .sharding import ShardedUnfoldJob
    from .unfold import pinned_pool_get, pinned_pool_put
    unfolder.verify_commens(rec_latt_pc, wavecar.B_matrix)
    grid = unfolder.set_energy_grid(E_start, E_end, delta_e)
    if hasattr(unfolder, "set_complex128"):
        unfolder.set_complex128(wavecar.double)      # a WAVECAR_double: the reference's dp build
    sc_k = np.array([wavecar.kpt_header(i + 1, ispin).kpt_frac_coords for i in range(wavecar.nkpts)])
    if hasattr(unfolder, "get_geom_unfolding_relations"):     # K7 on the device
        # -n_sckpts_to_skip: the first k-points of a hybrid-functional WAVECAR carry no bands
        # (cla_wrappers_mod.f90:258-268, band_unfolding_routines_mod.f90:311-315)
        fs, _, _ = unfolder.get_geom_unfolding_relations(pc_kpts_cart, sc_k @ wavecar.B_matrix,
                                                         n_sckpts_to_skip=n_sckpts_to_skip)
        folds = [[] for _ in range(wavecar.nkpts)]
        for ik, s in enumerate(fs):
            folds[int(s)].append(ik)
    else:
        folds = fold_pckpts_onto_sckpts(pc_kpts_cart, sc_k, wavecar.B_matrix)
    used = [i for i in range(wavecar.nkpts) if folds[i]]       # SC-kpts no pc-kpt folds onto are skipped
    costs = [float(wavecar.nbands) * wavecar.kpt_header(i + 1, ispin).nplane for i in used]
    # two buffers more than the pipeline depth: k-points j-1 and j-2 may be in flight and j+1 is being
    # read ahead while j is submitted; a buffer is refilled only after its k-point was fetched
    # (taken from / returned to a reuse pool: page-locking memory costs about a second per GB)
    use_mmap = reader == "mmap" and device_gvecs
    mm = np.memmap(wavecar.path, dtype=np.uint8, mode="r") if use_mmap else None
    bufs = [] if use_mmap else [pinned_pool_get(wavecar.band_block_bytes()) for _ in range(min(4, max(1, len(used))))]
    K_cart = sc_k @ wavecar.B_matrix
    n_loaded = [0]

    def load(j):
        i = used[j]
        fG = np.asarray(pc_kpts_cart)[folds[i]] - K_cart[i]
        if use_mmap:       # no copy here: the library stages the mapped (pageable) records chunk by chunk
            base = (wavecar.header_record(i + 1, ispin) + 1) * wavecar.nrecl
            raw = mm[base:base + wavecar.band_block_bytes()]

            def submit_mapped(u, job_ikpt, folding_G):
                wavecar.submit_to(u, job_ikpt, i + 1, folding_G, ispin, raw=raw)
            return submit_mapped, fG, np.asarray(folds[i])
        # the buffers rotate with THIS rank's loads (loads are sequential), not with the k-point index
        buf = bufs[n_loaded[0] % len(bufs)]
        n_loaded[0] += 1
        if device_gvecs:   # header + raw records only; the plane-wave list is made on the device
            raw = wavecar.read_band_records(i + 1, ispin, out=buf)      # file -> page-locked memory, here

            def submit(u, job_ikpt, folding_G):
                wavecar.submit_to(u, job_ikpt, i + 1, folding_G, ispin, raw=raw)
            return submit, fG, np.asarray(folds[i])
        return wavecar.wavefunction(i + 1, ispin, out=buf), fG, np.asarray(folds[i])

    # load() only touches the file and its own buffer, so the next k-point is parsed by a helper
    # thread while this one uploads and computes (file parsing overlaps compute)
    job = ShardedUnfoldJob(unfolder, len(used), costs, load, rank=rank, world_size=world_size, depth=2,
                           after_fetch=after_fetch, prefetch=not use_mmap)
    try:
        got = job.run(device=device, group=group)
    finally:
        for b in bufs:     # every k-point has been fetched (or the job failed): no upload is in flight
            pinned_pool_put(b)
    got.sckpt_of_row = {int(r): used[j] for j in range(len(used)) for r in folds[used[j]]}
    got.sckpts_cart = K_cart
    return got, grid
