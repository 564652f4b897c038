#!/usr/bin/env python
"""bench.py -- the unfolding hot path on B200: coefficient GB/s against the HBM roofline,
SC-kpts/s end to end, and the CPU restatement of the reference timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload sige555|mos2_8x8_spinor|si333_vacancy|graphene4x4]
                    [--kpts-per-step 0]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path (K1 coset
classification -> K2 fused norm + masked |C|^2 reduction -> K3 finalisation + delta_N) over one batch
of `kpts-per-step` SC k-points PER GPU (weak scaling: SC k-points are independent units;
the only exchange is the gather of the disjoint delta_N rows, done with NCCL when N > 1).

  value      coeff GB/s = sum over ranks of n_bands*n_pw*n_spinor*8 bytes / device time,
             inputs resident in HBM (each SC-kpt block is 2.4 GB >> 126 MB L2: no flush needed)
  e2e        the same metric through the host C ABI (bandup_gpu_submit_kpt/fetch_kpt) with
             pinned HOST buffers: H2D of every coefficient block and D2H of weights/delta_N
             inside the timed region (double-buffered, PCIe-bound)
  roofline   K2 (k2_weights_reduce) alone: algorithmic bytes per launch / mean launch time,
             measured live with CUDA events on the launching stream, vs MEASURED_PEAKS.json
  e2e_file   (N = 1) the reference's real loop body, file -> delta_N: a synthetic WAVECAR in the
             page cache read record by record into page-locked buffers, uploaded as it lies
             (bandup_gpu_submit_kpt_vasp, plane-wave list made on the device), next to the CPU
             restatement reading the same file into its coefficient array and unfolding it
  cpu_baseline  oracle/ (C + OpenMP restatement of the reference's Fortran loops; the Fortran
             itself cannot be compiled in this image) on a bounded sample of the same workload

--impl reference times that CPU restatement alone (rank 0 only), same metric and config.

The device-timed region lasts at least --min-seconds (0.5 s): `steps` in the JSON line is the number
of steps actually timed (>= --steps).  With N > 1 the delta_N rows are gathered ONCE per job, at the
end of the timed region, by the library's own collective (bandup_gpu_allgather_device over the
communicator made with bandup_gpu_comm_unique_id / bandup_gpu_comm_init).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SEED = 20261019 + 5
METRIC = "coeff_GB_per_s"
UNIT = "GB/s"

WORKLOADS = {
    "sige555": "C5: 1000-atom Si-Ge 5x5x5 supercell, 3000 bands, ~100k PWs, det M=500, 200 pc k-points (L-G-X) folding onto 138 SC k-points",
    "mos2_8x8_spinor": "C3: MoS2 8x8 monolayer, spinor, 1200 bands, ~80k PWs, det M=64, 48 pc k-points (G-M-K-G) folding onto 28 SC k-points",
    "si333_vacancy": "C2: Si 3x3x3 vacancy supercell, 500 bands, ~20k PWs, det M=108, 100 pc k-points (L-G-X) folding onto 99 SC k-points",
    "graphene4x4": "C1: graphene 4x4, 64 bands, ~23k PWs, det M=16, 30 pc k-points (K-G-M-K) folding onto 23 SC k-points",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sige555", choices=sorted(WORKLOADS))
    ap.add_argument("--kpts-per-step", type=int, default=0,
                    help="SC k-points per GPU per step; 0: as many of the configuration's k-points as fit --batch-gb "
                         "(C5: 16 of 138, C1: all 23)")
    ap.add_argument("--batch-gb", type=float, default=40.0,
                    help="HBM budget of one batch's coefficient blocks when --kpts-per-step is 0")
    ap.add_argument("--complex128", action="store_true",
                    help="the reference's kind_cplx_coeffs = dp build: the same blocks as complex128 (16-byte elements); "
                         "no CPU baseline / file leg")
    ap.add_argument("--kpt-offset", type=int, default=0, help="first SC k-point of rank 0's shard (A/B runs on other parts of the path)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: min(steps, 8)")
    ap.add_argument("--min-seconds", type=float, default=0.5,
                    help="the device-timed region runs at least this long (more steps than --steps if need be)")
    ap.add_argument("--no-e2e-file", action="store_true", help="skip the file-level end-to-end measurement")
    ap.add_argument("--file-kpts", type=int, default=4, help="SC k-points in the synthetic WAVECAR of e2e_file")
    ap.add_argument("--k2-variant", type=int, default=0, help="0 auto, 1 shared-memory bins, 2 register accumulators")
    ap.add_argument("--dependent-launch", type=int, default=3,
                    help="programmatic dependent launch: bit 0 = K2 behind K1, bit 1 = K3 behind K2 (3 = both, 0 = none)")
    ap.add_argument("--no-gpu-probe", action="store_true", help="do not measure which GPUs share a host uplink (N < visible GPUs)")
    ap.add_argument("--gpu-order", default="", help="CUDA device of local rank i, comma separated (default: spread over the PCIe tree)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin the rank to the CPU cores of its GPU's NUMA node")
    ap.add_argument("--no-pinned-probe", action="store_true",
                    help="e2e: take the first page-locked blocks instead of the fastest of four candidates")
    ap.add_argument("--tuning", default="", help="chunk_vecs,ctas_per_sm,n_stages of K2 (0 = default), for sweeps")
    ap.add_argument("--per-kpt-launch", action="store_true",
                    help="launch K1..K3 once per SC k-point instead of once per step (batch)")
    return ap.parse_args()


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock, power and throttle reasons of one GPU while the timed regions run.  NVML is polled
    in-process every ~2 ms (the device-timed region of the default run lasts only ~15 ms, far
    below nvidia-smi's 200 ms period); nvidia-smi is the fallback when pynvml is missing."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int, pci_bus_id: str = ""):
        self.index = index
        self.pci_bus_id = pci_bus_id   # NVML numbers GPUs in PCI order, CUDA need not: prefer the bus id
        self.samples = []          # (t, sm_mhz, power_w, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self.thread = None
        self.source = None
        self.proc = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.pci_bus_id:
                try:
                    h = pynvml.nvmlDeviceGetHandleByPciBusId(self.pci_bus_id.encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append((time.perf_counter(),
                                             float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons_fn(h))))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml, 2 ms period"
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)

            def pump():
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for line in self.proc.stdout:
                    f = [x.strip() for x in line.split(",")]
                    try:
                        mask = sum(self.REASONS[n] for n, v in zip(names, f[3:7]) if v.lower().startswith("active"))
                        self.samples.append((time.perf_counter(), float(f[0]), float(f[2]), mask))
                        self.max_mhz = float(f[1])
                    except (ValueError, IndexError):
                        continue
            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
            self.source = "nvidia-smi, 200 ms period"
        except Exception:
            self.proc = None

    def stop(self):
        self._stop.set()
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self, t0: float, t1: float) -> dict:
        """Median SM clock and the throttle reasons seen between t0 and t1 (perf_counter times)."""
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        inside = [x for x in self.samples if t0 <= x[0] <= t1]
        note = None
        if not inside:     # region shorter than the sampling period: the nearest samples on either side
            inside = sorted(self.samples, key=lambda x: min(abs(x[0] - t0), abs(x[0] - t1)))[:2]
            note = "region shorter than the sampling period; nearest samples used"
        mask = 0
        for x in inside:
            mask |= x[3]
        out = {"sm_mhz": float(np.median([x[1] for x in inside])), "sm_max_mhz": self.max_mhz,
               "power_w_max": max(x[2] for x in inside), "samples": len(inside),
               "reasons": sorted(n for n, bit in self.REASONS.items() if mask & bit), "source": self.source}
        if note:
            out["note"] = note
        return out


def make_energies(sc, rng):
    return np.sort(rng.uniform(sc.e_fermi - 20.5, sc.e_fermi + 5.5, size=sc.n_bands))

# --------------------------------------------------------------------------- file-level e2e
def file_dir(need_bytes):
    """A directory with room for the synthetic WAVECAR ($TMPDIR, /tmp, /dev/shm), or None."""
    import shutil
    import tempfile
    for d in (os.environ.get("TMPDIR"), tempfile.gettempdir(), "/dev/shm"):
        try:
            if d and os.path.isdir(d) and shutil.disk_usage(d).free > need_bytes + (2 << 30):
                return d
        except OSError:
            continue
    return None


def write_synthetic_wavecar(path, sc, k_indices, energies):
    """A single-precision WAVECAR (read_wavecar_mod.f90:112-127, 356-452) of the workload's own
    SC k-points, coefficients from the counter-based generator (oracle.synth_fill: the same numbers
    the device-resident arm generates in HBM), one write per k-point.  Scalar or spinor; returns
    (nrecl, coefficient bytes in the file)."""
    from oracle import oracle as O
    nb, nsp = sc.n_bands, sc.n_spinor
    gvs = [sc.gvecs(iK) for iK in k_indices]
    nrecl = max(max(g.shape[0] for g in gvs) * nsp * 8, 8 * (4 + 3 * nb), 96)
    coeff_bytes = 0
    with open(path, "wb") as f:
        def put(arr):
            b = np.ascontiguousarray(arr, dtype="<f8").tobytes()
            f.write(b + b"\0" * (nrecl - len(b)))
        put([nrecl, 1, 45200])
        put(np.concatenate([[len(k_indices), nb, sc.ecut], np.asarray(sc.A_sc, dtype=np.float64).reshape(9)]))
        for iK, gf, ener in zip(k_indices, gvs, energies):
            n_pw = gf.shape[0]
            row = n_pw * nsp
            hdr = np.zeros(4 + 3 * nb)
            hdr[0] = row
            hdr[1:4] = sc.sc_kpts_frac[iK]
            hdr[4::3] = ener
            hdr[6::3] = 1.0
            put(hdr)
            rec = np.empty((nb, nrecl // 8), dtype=np.complex64)
            rec[:, row:] = 0
            block = O.synth_fill(row, nb, SEED, iK)                  # [band][pw][spinor]
            if nsp == 1:
                rec[:, :row] = block
            else:       # records hold the spinor-up block, then the spinor-down block (:570-577)
                rec[:, :row] = block.reshape(nb, n_pw, 2).transpose(0, 2, 1).reshape(nb, row)
            f.write(memoryview(rec.view(np.uint8).reshape(-1)))
            coeff_bytes += nb * row * 8
            del rec, block
    return nrecl, coeff_bytes


def cpu_read_kpt(wc, ikpt, out, n_pw, nsp, threads):
    """What read_wavecar does for one k-point (read_wavecar_mod.f90:537-583): every band record of
    the file copied into the coefficient array wf%pw_coeffs(n_spinor, n_pw, n_bands) -- here with
    `threads` readers at once (the reference reads with one).  out: (n_bands, n_pw, n_spinor) c64."""
    from concurrent.futures import ThreadPoolExecutor
    nb, nrecl = wc.nbands, wc.nrecl
    base = (wc.header_record(ikpt) + 1) * nrecl
    row_bytes = n_pw * nsp * 8
    flat = out.reshape(nb, n_pw * nsp)
    fd = os.open(wc.path, os.O_RDONLY)
    try:
        def part(lo_hi):
            lo, hi = lo_hi
            tmp = np.empty(n_pw * nsp, dtype=np.complex64) if nsp == 2 else None
            for b in range(lo, hi):
                dst = tmp if nsp == 2 else flat[b]
                mv = memoryview(dst.view(np.uint8))
                done = 0
                while done < row_bytes:
                    k = os.preadv(fd, [mv[done:]], base + b * nrecl + done)
                    if k <= 0:
                        raise IOError("short read")
                    done += k
                if nsp == 2:    # up block, down block -> (pw, spinor) interleaved, as the reader stores it
                    flat[b] = tmp.reshape(2, n_pw).T.reshape(-1)
        step = -(-nb // threads)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(part, [(lo, min(lo + step, nb)) for lo in range(0, nb, step)]))
    finally:
        os.close(fd)


def cpu_file_sample(sc, path, k_indices, repeats, threads=None):
    """The reference's loop body from the file: read_wavecar of every k-point of the WAVECAR, then
    renormalise + select + weights + delta_N (main_BandUP.f90:145-172).  Returns (GB/s of coefficient
    data, kpts/s, cores, seconds reading, seconds computing)."""
    from bandup_b200.wavecar import Wavecar
    from oracle import oracle as O
    O.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = O.num_threads()
    grid = O.real_seq(*sc.energy_window())
    wc = Wavecar(path)
    tot_bytes, t_read, t_comp, done = 0, 0.0, 0.0, 0
    for _ in range(repeats):
        for j, iK in enumerate(k_indices):
            h = wc.kpt_header(j + 1)
            gf = sc.gvecs(iK)
            n_pw = gf.shape[0]
            assert h.nplane == n_pw * sc.n_spinor
            gc = gf @ sc.B_sc
            fG = sc.folding_G(iK)
            t0 = time.perf_counter()
            block = np.empty((sc.n_bands, n_pw, sc.n_spinor), dtype=np.complex64)      # the reader's allocate
            cpu_read_kpt(wc, j + 1, block, n_pw, sc.n_spinor, min(cores, 16))
            t1 = time.perf_counter()
            O.unfold_kpt(block, gc, h.band_energies, sc.b_pc, fG, grid, norm_mode=0, in_place=True)
            t2 = time.perf_counter()
            t_read += t1 - t0
            t_comp += t2 - t1
            tot_bytes += block.nbytes
            done += 1
            del block
    tot = t_read + t_comp
    return tot_bytes / tot / 1e9, done / tot, cores, t_read, t_comp


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(sc, k_indices, seconds_budget, max_kpts, threads=None):
    """Times oracle.unfold_kpt (renormalise + select + weights + delta_N: the reference's loop
    body, main_BandUP.f90:145-172) on full-size SC k-points of the workload.  Returns
    (GB/s, kpts/s, cores, n_kpts_done, per-phase seconds)."""
    from oracle import oracle as O
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the
    # reference's OpenMP loops are meant to run on every core)
    O.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = O.num_threads()
    grid = O.real_seq(*sc.energy_window())
    rng = np.random.default_rng(SEED)
    tot_bytes, tot_t, done = 0, 0.0, 0
    phases = np.zeros(3)
    block = None
    for iK in k_indices[:max_kpts]:
        gf = sc.gvecs(iK)
        n_pw = gf.shape[0]
        row = n_pw * sc.n_spinor
        if block is None or block.shape[1] != row:
            block = np.empty((sc.n_bands, row), dtype=np.complex64)
        O.synth_fill(row, sc.n_bands, SEED, iK, out=block)
        ener = make_energies(sc, rng)
        gc = gf @ sc.B_sc
        fG = sc.folding_G(iK)
        t0 = time.perf_counter()
        _, _, _, times = O.unfold_kpt(block.reshape(sc.n_bands, n_pw, sc.n_spinor), gc, ener, sc.b_pc, fG, grid,
                                      norm_mode=0, in_place=True)
        dt = time.perf_counter() - t0
        tot_t += dt
        phases += times
        tot_bytes += sc.n_bands * row * 8
        done += 1
        if tot_t >= seconds_budget:
            break
    return tot_bytes / tot_t / 1e9, done / tot_t, cores, done, phases


def batch_kpts(args, sc):
    """SC k-points per GPU per step (--kpts-per-step, else what fits --batch-gb)."""
    esz = 16 if args.complex128 else 8
    n = args.kpts_per_step
    if n <= 0:
        # one batch = as much of the job as ~40 GB of the 180 GB of HBM holds (small cells: the whole
        # job; the 1000-atom cell: 16 SC-kpts).  K1/K3 and K2's ramp and drain are paid once per batch:
        # 4 / 8 / 16 SC-kpts of C5 per launch run at 7174 / 7251 / 7293 GB/s whole step on the same board
        per_kpt = sc.n_bands * sc.gvecs(0).shape[0] * sc.n_spinor * esz
        n = max(1, min(sc.n_sc_kpts, int(args.batch_gb * 1e9 // per_kpt)))
    return n


def workload_config(args, sc, world, kps, n_pw, n_fold, nener):
    """The `config` object of the JSON line: the workload and how the B200 arm runs it.  The reference
    arm prints the same object for the same arguments (it times a bounded sample of that workload)."""
    esz = 16 if args.complex128 else 8
    step_bytes = sum(sc.n_bands * n * sc.n_spinor * esz for n in n_pw)
    return {"workload": WORKLOADS[args.workload] + (" [complex128 coefficients]" if args.complex128 else ""),
            "kpts_per_step_per_gpu": kps, "n_bands": sc.n_bands,
            "n_spinor": sc.n_spinor, "n_pw": [int(n) for n in n_pw],
            "n_fold": [int(n) for n in n_fold], "nener": int(nener),
            "launches": "per SC k-point" if args.per_kpt_launch else
            "one launch of K1, K2, K3 per step (bandup_gpu_unfold_device_batch)",
            "l2": (f"inputs larger than L2: every step streams {step_bytes / 1e9:.2f} GB of coefficients "
                   "(L2: 126 MB), read with evict_first" if step_bytes > 2 * 126e6 else
                   f"WARNING: a step streams only {step_bytes / 1e6:.0f} MB (L2: 126 MB); raise --kpts-per-step"),
            "k2_variant": {0: "auto", 1: "shared-memory bins", 2: "register accumulators"}[args.k2_variant],
            "dependent_launch": args.dependent_launch,
            "parallelism": f"SC k-points sharded over {world} GPU(s); delta_N rows gathered once per job by "
                           "bandup_gpu_allgather_device (ncclAllGather inside libbandup_gpu.so)"
            if world > 1 else "1 GPU"}


def run_reference(args, sc, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    ks = list(range(sc.n_sc_kpts))
    n_w, n_s = max(args.warmup, 0), max(args.steps, 1)
    # bounded sample: each step is ONE full-size SC k-point of the workload (its own batch of
    # kpts_per_step would take minutes on the host)
    if n_w:
        cpu_sample(sc, ks, 1e9, min(n_w, 1))
    gbs, kps, cores, done, phases = cpu_sample(sc, ks[1:], 1e9, n_s)
    e2e_file = None
    if not args.no_e2e_file:
        e2e_file = reference_file_leg(args, sc)
    # the B200 arm's configuration for the same arguments (rank 0's batch)
    kps_b = batch_kpts(args, sc)
    my_k = [(args.kpt_offset + j) % sc.n_sc_kpts for j in range(kps_b)]
    line = {
        "metric": METRIC, "value": gbs, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": n_s, "warmup": n_w, "ms_per_step": 1e3 / kps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c64 in, f32 |C|^2, f64 accumulate",
        "data": "synthetic",
        "config": workload_config(args, sc, world, kps_b, [sc.gvecs(iK).shape[0] for iK in my_k],
                                  [len(sc.folding_G(iK)) for iK in my_k], len(O.real_seq(*sc.energy_window()))),
        "sc_kpts_per_s": kps,
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{done} full-size SC k-points of the workload, one per step (a whole batch per "
                                   "step would take minutes on the host); C+OpenMP restatement of "
                                   "the reference loops (Fortran unbuildable here: no Fortran compiler)",
                         "phase_seconds": {"renormalise": phases[0], "select+weights": phases[1],
                                           "delta_N": phases[2]}},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "e2e_file": e2e_file,
    }
    print(json.dumps(line), flush=True)


def file_kpts(args, sc):
    """The SC k-points of the e2e_file WAVECAR: the first --file-kpts of the workload, fewer when a
    k-point block is large (at most ~10 GB of file)."""
    per_k = sc.n_bands * sc.gvecs(0).shape[0] * sc.n_spinor * 8
    n = max(1, min(args.file_kpts if per_k > (1 << 30) else max(args.file_kpts, int(2e9 // per_k)),
                   sc.n_sc_kpts, max(1, int(10.5e9 // per_k))))
    return list(range(n))


def reference_file_leg(args, sc):
    """e2e_file of the CPU arm: the port reads a synthetic WAVECAR (page cache warm) k-point by
    k-point into its coefficient array and unfolds it."""
    import tempfile
    ks = file_kpts(args, sc)
    rng = np.random.default_rng(SEED)
    eners = [make_energies(sc, rng) for _ in ks]
    need = sum(sc.n_bands * sc.gvecs(iK).shape[0] * sc.n_spinor * 8 for iK in ks)
    d = file_dir(need)
    if d is None:
        return {"unavailable": "no directory with room for the synthetic WAVECAR"}
    tmp = tempfile.mkdtemp(prefix="bandup_bench_", dir=d)
    path = os.path.join(tmp, "WAVECAR")
    try:
        nrecl, coeff_bytes = write_synthetic_wavecar(path, sc, ks, eners)
        cpu_file_sample(sc, path, ks[:1], 1)                       # warm-up (page cache, thread pools)
        gbs, kps, cores, t_read, t_comp = cpu_file_sample(sc, path, ks, 1)
        return {"value": gbs, "unit": UNIT, "sc_kpts_per_s": kps, "cores": cores, "sc_kpts": len(ks),
                "file_bytes": os.path.getsize(path), "coeff_bytes": coeff_bytes, "dir": d, "page_cache": "warm",
                "seconds": {"read_wavecar": t_read, "renormalise+unfold": t_comp},
                "what": "read_wavecar of every k-point of a synthetic WAVECAR (16 reader threads) + "
                        "renormalise, select, weights, delta_N: the reference's loop body (main_BandUP.f90:145-172)"}
    finally:
        try:
            os.remove(path)
            os.rmdir(tmp)
        except OSError:
            pass


# --------------------------------------------------------------------------- GPU arm
def gpu_file_leg(args, sc, u):
    """e2e_file of the GPU arm: the reference's loop body from the file -- every SC k-point of a
    synthetic WAVECAR (page cache warm) through bandup_b200.wavecar.unfold_wavecar: raw band
    records read by a helper thread into page-locked buffers, uploaded as they lie
    (bandup_gpu_submit_kpt_vasp), plane-wave list and folding map made on the device, results
    fetched to the host.  The first k-point is checked against the oracle (outside the timing)."""
    import tempfile
    from bandup_b200.wavecar import Wavecar, unfold_wavecar
    from oracle import oracle as O
    ks = file_kpts(args, sc)
    rng = np.random.default_rng(SEED)
    eners = [make_energies(sc, rng) for _ in ks]
    need = sum(sc.n_bands * sc.gvecs(iK).shape[0] * sc.n_spinor * 8 for iK in ks)
    d = file_dir(need)
    if d is None:
        return {"unavailable": "no directory with room for the synthetic WAVECAR"}
    tmp = tempfile.mkdtemp(prefix="bandup_bench_", dir=d)
    path = os.path.join(tmp, "WAVECAR")
    try:
        nrecl, coeff_bytes = write_synthetic_wavecar(path, sc, ks, eners)
        wc = Wavecar(path)
        pc = np.concatenate([sc.pc_kpts_cart[sc.folds[i]] for i in ks])
        window = sc.energy_window()
        reps = 3
        timings = {}
        for reader in ("pread", "mmap"):
            got, grid = unfold_wavecar(u, wc, sc.b_pc, pc, *window, reader=reader)   # warm-up: page cache, pinned pool
            t0 = time.perf_counter()
            for _ in range(reps):
                got, grid = unfold_wavecar(u, wc, sc.b_pc, pc, *window, reader=reader)
            timings[reader] = time.perf_counter() - t0
        dt = timings["mmap"]                                              # unfold_wavecar's default reader is the headline
        # parity of what came back (first k-point of the file), outside the timed region
        iK = ks[0]
        gf = sc.gvecs(iK)
        blk = O.synth_fill(gf.shape[0] * sc.n_spinor, sc.n_bands, SEED, iK).reshape(sc.n_bands, gf.shape[0], sc.n_spinor)
        _, dN, ns, _ = O.unfold_kpt(blk, gf @ sc.B_sc, eners[0], sc.b_pc, sc.folding_G(iK), grid, norm_mode=1,
                                    in_place=True)
        mine = got.dN[:len(sc.folds[iK])]
        err = float(np.max(np.abs(mine - dN) / np.maximum(np.abs(dN), 1e-12)))
        if err > 1e-6:
            raise AssertionError(f"bench e2e_file: delta_N differs from the oracle (max rel err {err:.2e})")
        return {"value": reps * coeff_bytes / dt / 1e9, "unit": UNIT, "sc_kpts_per_s": reps * len(ks) / dt,
                "sc_kpts": len(ks), "passes": reps, "file_bytes": os.path.getsize(path), "coeff_bytes": coeff_bytes,
                "dir": d, "page_cache": "warm", "checked_vs_oracle_max_rel_err": err,
                "by_reader": {k: reps * coeff_bytes / v / 1e9 for k, v in timings.items()},
                "what": "synthetic WAVECAR -> bandup_b200.wavecar.unfold_wavecar (file mapped, raw records staged by the "
                        "library's page-locked ring as they lie, bandup_gpu_submit_kpt_vasp) -> delta_N on the host; "
                        "by_reader: the same with threaded preadv into page-locked buffers"}
    finally:
        from bandup_b200.unfold import pinned_pool_release
        pinned_pool_release()
        try:
            os.remove(path)
            os.rmdir(tmp)
        except OSError:
            pass


def run_b200(args, sc, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from bandup_b200 import build as _build
    if rank == 0:
        _build.build()
    # Which GPU this local rank drives.  End to end the path is bound by host->device copies, and GPUs
    # behind one PCIe switch share its uplink, so fewer ranks than GPUs are spread over the uplinks
    # instead of filling the box in index order: rank 0 measures which GPUs slow each other down
    # (bandup_b200.numa.probe_h2d_order, in a child process) and tells the others over gloo, before
    # any rank binds to a device.  --gpu-order overrides; with every GPU in use it is index order.
    from bandup_b200.numa import device_for_local_rank, probed_gpu_order
    if args.gpu_order:
        os.environ["BANDUP_B200_GPU_ORDER"] = args.gpu_order
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at the VERSION and
        # WARN levels (showVersion in NCCL's init.cc; the level may also come from /etc/nccl.conf);
        # an explicit INFO/TRACE from the caller is respected
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("cpu:gloo,cuda:nccl")      # NCCL binds to the current device at its first use
        probe = [None]
        if not args.gpu_order and not args.no_gpu_probe and local_world < torch.cuda.device_count():
            if rank == 0:
                probe[0] = probed_gpu_order()
            dist.broadcast_object_list(probe, src=0, device=torch.device("cpu"))      # over gloo: no GPU is bound yet
            if probe[0] and sorted(probe[0].get("order", [])) == list(range(torch.cuda.device_count())):
                os.environ["BANDUP_B200_GPU_ORDER"] = ",".join(str(i) for i in probe[0]["order"])
        devinfo = device_for_local_rank(local_rank, local_world)
        if probe[0]:
            devinfo["source"] = probe[0]["source"]
            devinfo["h2d_prefix_total_GBps"] = probe[0].get("prefix_total_GBps")
            devinfo["h2d_solo_GBps"] = probe[0]["solo_GBps"]
        cuda_idx = devinfo["device"]
        torch.cuda.set_device(cuda_idx)
        dist.all_reduce(torch.zeros(1, device=torch.device("cuda", cuda_idx)))      # first NCCL use: binds the device
        torch.cuda.synchronize()
    else:
        _build.build()
        devinfo = device_for_local_rank(local_rank, local_world)
        cuda_idx = devinfo["device"]
    from bandup_b200.unfold import (GpuUnfolder, PwWavefunction, LAYOUT_FORTRAN_INMEM, pinned_empty, pinned_empty_near, pinned_free, K_COSET,
                                    K_WEIGHTS, K_DELTA_N)

    from bandup_b200._capi import CONST
    torch.cuda.set_device(cuda_idx)
    dev = torch.device("cuda", cuda_idx)
    numa = {"bound": False}
    all_cpus = os.sched_getaffinity(0)
    if not args.no_numa_bind:
        from bandup_b200.numa import bind_to_gpu_node
        numa = bind_to_gpu_node(cuda_idx)      # before any page-locked allocation (first touch)
    u = GpuUnfolder(cuda_idx)
    u.verify_commens(sc.b_pc, sc.B_sc)
    grid = u.set_energy_grid(*sc.energy_window())
    if args.tuning:
        u.set_tuning(*[int(x) for x in args.tuning.split(",")])
    if args.k2_variant:
        u.set_option(CONST["BANDUP_GPU_OPT_K2_VARIANT"], args.k2_variant)
    if args.dependent_launch != 3:
        u.set_option(CONST["BANDUP_GPU_OPT_DEPENDENT_LAUNCH"], args.dependent_launch)
    if world > 1:
        # the library's own communicator (ncclCommInitRank inside libbandup_gpu.so): rank 0 makes the
        # id, torch.distributed carries its 128 bytes to the others (MPI_Bcast in a Fortran host)
        box = [GpuUnfolder.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, device=torch.device("cpu"))
        u.comm_init(world, rank, box[0])
    nener = len(grid)
    nb, nsp = sc.n_bands, sc.n_spinor
    esz = 16 if args.complex128 else 8          # bytes per coefficient
    if args.complex128:
        u.set_complex128(True)
        args.no_cpu_baseline = args.no_e2e_file = True
    kps = batch_kpts(args, sc)
    rng = np.random.default_rng(SEED + rank)
    stream = torch.cuda.current_stream(dev).cuda_stream

    # this rank's shard of SC k-points (contiguous blocks of the K list)
    my_k = [(args.kpt_offset + rank * kps + j) % sc.n_sc_kpts for j in range(kps)]
    items = []
    for iK in my_k:
        gf = sc.gvecs(iK)
        n_pw = gf.shape[0]
        g0, _ = u.folding_g0(sc.folding_G(iK))
        nf = g0.shape[0]
        ener = make_energies(sc, rng)
        d_c = torch.empty((nb, n_pw * nsp, 2), dtype=torch.float32, device=dev)
        u.synth_fill_device(stream, d_c.data_ptr(), n_pw * nsp, nb, n_pw * nsp * 8, SEED, iK)
        if args.complex128:     # the same numbers widened to complex128
            d_c = torch.view_as_real(torch.view_as_complex(d_c).to(torch.complex128)).contiguous()
        ws = u.workspace_bytes(nsp, n_pw, nb, nf)
        items.append(dict(
            iK=iK, n_pw=n_pw, nf=nf, g0=g0, gf=gf, ener=ener, d_c=d_c,
            d_g=torch.from_numpy(gf).to(dev), d_e=torch.from_numpy(ener).to(dev),
            d_w=torch.empty((nf, nb), dtype=torch.float64, device=dev),
            d_ns=torch.empty(nf, dtype=torch.int32, device=dev),
            d_ws=torch.empty(ws, dtype=torch.uint8, device=dev), ws=ws))
    nf_max = max(len(f) for f in sc.folds)
    # delta_N rows of the step, padded to nf_max rows per SC-kpt for the gather
    d_dn = torch.zeros((kps, nf_max, nener), dtype=torch.float64, device=dev)
    d_dn_all = torch.empty((world, kps, nf_max, nener), dtype=torch.float64, device=dev) if world > 1 else None
    torch.cuda.synchronize()
    step_bytes = sum(nb * it["n_pw"] * nsp * esz for it in items)

    # one batch descriptor per step: K1, K2, K3 are each launched once for all kps k-points
    descs, _keep = u.make_batch([dict(
        n_spinor=nsp, n_pw=it["n_pw"], n_bands=nb, layout=LAYOUT_FORTRAN_INMEM,
        band_stride_bytes=it["n_pw"] * nsp * esz, d_coeffs=it["d_c"].data_ptr(), d_g_frac=it["d_g"].data_ptr(),
        d_band_energies=it["d_e"].data_ptr(), g0=it["g0"], d_weights=it["d_w"].data_ptr(),
        d_dN=d_dn[j].data_ptr(), d_n_selected=it["d_ns"].data_ptr()) for j, it in enumerate(items)])
    ws_batch = u.batch_workspace_bytes(descs)
    d_ws_batch = torch.empty(ws_batch, dtype=torch.uint8, device=dev)

    def step():
        if args.per_kpt_launch:
            for j, it in enumerate(items):
                u.unfold_device(stream, nsp, it["n_pw"], nb, LAYOUT_FORTRAN_INMEM, it["n_pw"] * nsp * esz,
                                it["d_c"].data_ptr(), it["d_g"].data_ptr(), it["d_e"].data_ptr(), it["g0"],
                                it["d_w"].data_ptr(), d_dn[j].data_ptr(), it["d_ns"].data_ptr(),
                                it["d_ws"].data_ptr(), it["ws"])
        else:
            u.unfold_device_batch(stream, descs, d_ws_batch.data_ptr(), ws_batch)

    def gather():
        # the path's only exchange, ONCE per job (after the last SC-kpt, main_BandUP.f90:187-191 is
        # where the gathered table is consumed): the ranks' disjoint delta_N rows, by the library's
        # collective on the same stream as the kernels
        if world > 1:
            u.allgather_device(stream, d_dn.data_ptr(), d_dn_all.data_ptr(), d_dn.numel())

    _bar = torch.zeros(1, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(_bar)       # NCCL on this rank's device
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing -------------------------------------------------
    n_warm = max(3, args.warmup)          # never fewer than three untimed steps before the timed region
    for _ in range(n_warm):
        step()
    gather()
    barrier()
    # calibration: the timed region lasts at least --min-seconds whatever --steps says
    ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ec0.record()
    for _ in range(3):
        step()
    ec1.record()
    torch.cuda.synchronize()
    est_ms = max_over_ranks(ec0.elapsed_time(ec1) / 3.0)
    n_steps = max(args.steps, int(np.ceil(1.1 * args.min_seconds * 1e3 / max(est_ms, 1e-3))))
    # per-kernel breakdown in its own untimed pass: an event pair between back-to-back launches
    # costs device time and serialises the dependent-launch chain
    u.set_profiling(True)
    u.reset_kernel_times()
    n_breakdown = 3
    for _ in range(n_breakdown):
        step()
    torch.cuda.synchronize()
    per_kernel = {name: u.kernel_time(k)[0] / n_breakdown for k, name in
                  ((K_COSET, "k1_coset_classify"), (K_WEIGHTS, "k2_weights_reduce"),
                   (K_DELTA_N, "k3_finalize_delta_n"))}
    u.set_profiling(False)
    u.reset_kernel_times()
    launches0 = u.launch_count()
    bus = (devinfo.get("pci") or [""] * (cuda_idx + 1))[cuda_idx]
    sampler = ClockSampler(cuda_idx, bus)
    if rank == 0:
        sampler.start()
    # (A) the whole job: n_steps steps, then the one gather of the delta_N rows; no event brackets
    # inside, K1 -> K2 -> K3 chained by programmatic dependent launch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_dev0 = time.perf_counter()
    e0.record()
    for _ in range(n_steps):
        step()
    gather()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    gpu_launches = u.launch_count() - launches0
    # (B) the roofline kernel alone, live: the same steps again with every K2 launch bracketed by
    # CUDA events on its stream (brackets would serialise the chain, hence not inside (A))
    u.set_profiling(True, kernels=[K_WEIGHTS])
    u.reset_kernel_times()
    barrier()
    for _ in range(n_steps):
        step()
    barrier()
    t_dev1 = time.perf_counter()
    k2_ms, k2_n = u.kernel_time(K_WEIGHTS)
    all_ms = sum(per_kernel.values())
    u.set_profiling(False)
    # every rank streams its own shard; shard sizes differ only through n_pw(K) (a few per cent)
    if world > 1:
        t = torch.tensor([float(step_bytes)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        total_bytes = float(t.item()) * n_steps
    else:
        total_bytes = float(step_bytes) * n_steps
    value = total_bytes / (ms * 1e-3) / 1e9
    kpts_per_s = kps * world * n_steps / (ms * 1e-3)
    k2_bytes = step_bytes / kps if args.per_kpt_launch else step_bytes
    peak, peak_src = measured_peak_gbs()
    achieved = k2_bytes / (k2_ms / max(k2_n, 1) * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k2_weights_reduce", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "bytes_per_launch": k2_bytes, "ms_per_launch": k2_ms / max(k2_n, 1), "launches_timed": int(k2_n),
                "measured": "CUDA events around every K2 launch on its stream, in a second pass of the same "
                            f"{n_steps} steps right after the timed region (event brackets between the kernels "
                            "would serialise the dependent-launch chain the whole-step value is timed with)",
                "k2_share_of_kernel_time": per_kernel["k2_weights_reduce"] / all_ms if all_ms else None,
                "kernel_ms_per_step": per_kernel,
                "kernel_ms_per_step_source": f"separate untimed pass of {n_breakdown} steps, every launch bracketed"}
    ncu_traffic = ROOT / "profiles" / "k2_traffic.json"
    if ncu_traffic.exists():
        try:
            t = json.loads(ncu_traffic.read_text())
            digest = (ROOT / "bandup_b200" / "lib" / "libbandup_gpu.digest").read_text().strip()
            # the capture must be of THIS library (source digest) and of the same launch shape (same
            # workload, same k-points per launch); otherwise it is dropped, not reported stale
            if t.get("lib_digest") != digest:
                roofline["traffic_dropped"] = "profiles/k2_traffic.json was captured with a different build of libbandup_gpu.so"
            elif t.get("workload") == args.workload and abs(t["dram_bytes_per_launch"] / k2_bytes - 1.0) < 0.05:
                roofline["traffic"] = t["dram_bytes_per_launch"]
                roofline["traffic_source"] = t.get("source")
            else:
                roofline["traffic_dropped"] = "profiles/k2_traffic.json is of another workload or launch shape"
        except Exception:
            pass

    # ---- end to end through the host C ABI ----------------------------------------
    e2e = None
    t_e2e = (0.0, 0.0)
    if not args.no_e2e:
        n_e2e = args.e2e_steps or min(args.steps, 8)
        host = []
        # two distinct pinned blocks, cycled (2 x 2.4 GB per rank), the two of four candidates that a
        # host->device copy reads fastest: the boxes are VMs with a hidden host NUMA layout
        # (bandup_b200.unfold.pinned_empty_near)
        src_items = items[:2]
        blk_max = max(nb * it["n_pw"] * nsp * esz for it in src_items)
        blocks, cand_rates, cand_kept = pinned_empty_near(blk_max, n_keep=len(src_items),
                                                          spares=0 if args.no_pinned_probe else (2 if world == 1 else 1),
                                                          device=cuda_idx)
        for h_full, it in zip(blocks, src_items):
            nbytes = nb * it["n_pw"] * nsp * esz
            h = h_full[:nbytes]
            torch.from_numpy(h).copy_(it["d_c"].view(torch.uint8).reshape(-1))
            host.append((h, it))
        torch.cuda.synchronize()
        ctype = np.complex128 if args.complex128 else np.complex64
        wfs = [PwWavefunction(h.view(ctype).reshape(nb, it["n_pw"], nsp), it["gf"], it["ener"])
               for h, it in host]
        seq = [(i % len(host)) for i in range(kps)]
        h2d = sum(nb * host[i][1]["n_pw"] * nsp * esz + host[i][1]["n_pw"] * 12 + nb * 8 for i in seq)
        d2h = sum(host[i][1]["nf"] * (nb + nener) * 8 + host[i][1]["nf"] * 4 + nb * 8 for i in seq)

        def e2e_step(base):
            # double-buffered: SC-kpt j+1 uploads while j computes (bandup_gpu_pipeline_depth = 2)
            u.submit(base, wfs[seq[0]], g0=host[seq[0]][1]["g0"])
            out = None
            for j in range(kps):
                if j + 1 < kps:
                    u.submit(base + j + 1, wfs[seq[j + 1]], g0=host[seq[j + 1]][1]["g0"])
                out = u.fetch(base + j)
            return out

        e2e_step(0)
        # The timed job: kps * n_e2e SC-kpts per GPU in total, handed out one at a time from a
        # counter all ranks share (bandup_b200.sharding.KptTickets), so a rank behind a slower
        # host link takes fewer k-points instead of holding the others back -- SC-kpts are
        # independent and the result does not depend on who computed which.
        from bandup_b200.sharding import KptTickets
        n_total = kps * n_e2e * world
        tickets = KptTickets.from_default_group(n_total, job="bench_e2e")
        blk_bytes = [nb * host[i][1]["n_pw"] * nsp * esz for i in range(len(host))]
        blk_h2d = [b + host[i][1]["n_pw"] * 12 + nb * 8 for i, b in enumerate(blk_bytes)]
        blk_d2h = [host[i][1]["nf"] * (nb + nener) * 8 + host[i][1]["nf"] * 4 + nb * 8 for i in range(len(host))]
        mine = 0
        last, last_blk = None, 0
        barrier()
        t0 = time.perf_counter()
        cur = tickets.take()
        if cur is not None:
            u.submit(10_000 + cur, wfs[0], g0=host[0][1]["g0"])
        while cur is not None:
            nxt = tickets.take()
            if nxt is not None:
                b = (mine + 1) % len(host)
                u.submit(10_000 + nxt, wfs[b], g0=host[b][1]["g0"])
            last, last_blk = u.fetch(10_000 + cur), mine % len(host)
            mine += 1
            cur = nxt
        torch.cuda.synchronize()
        t_e2e = (t0, time.perf_counter())
        dt_mine = t_e2e[1] - t0
        dt = max_over_ranks(dt_mine)
        # outside the timed region: what the last fetch returned must be, bit for bit, what the
        # device-resident entry point computes for the same block (same kernels, same tile plan)
        e2e_checked = None
        if last is not None:
            it = host[last_blk][1]
            c_w = torch.zeros((it["nf"], nb), dtype=torch.float64, device=dev)
            c_dn = torch.zeros((it["nf"], nener), dtype=torch.float64, device=dev)
            c_ns = torch.zeros(it["nf"], dtype=torch.int32, device=dev)
            c_nr = torch.zeros(nb, dtype=torch.float64, device=dev)
            u.unfold_device(stream, nsp, it["n_pw"], nb, LAYOUT_FORTRAN_INMEM, it["n_pw"] * nsp * esz,
                            it["d_c"].data_ptr(), it["d_g"].data_ptr(), it["d_e"].data_ptr(), it["g0"],
                            c_w.data_ptr(), c_dn.data_ptr(), c_ns.data_ptr(), it["d_ws"].data_ptr(), it["ws"],
                            d_norms=c_nr.data_ptr())
            torch.cuda.synchronize()
            same = (np.array_equal(last.spectral_weight, c_w.cpu().numpy()) and np.array_equal(last.dN, c_dn.cpu().numpy())
                    and np.array_equal(last.n_selected, c_ns.cpu().numpy())
                    and np.array_equal(last.band_norms, c_nr.cpu().numpy()))
            if not same:
                raise AssertionError("bench e2e: the host path's last fetched result differs from the "
                                     "device-resident result of the same SC k-point")
            e2e_checked = bool(same and float(last.spectral_weight.sum()) > 0.0 and int(last.n_selected.min()) > 0)
        n0, n1 = (mine + 1) // 2, mine // 2          # k-points taken from host block 0 / 1
        sums = torch.tensor([n0 * blk_bytes[0] + n1 * blk_bytes[-1], n0 * blk_h2d[0] + n1 * blk_h2d[-1],
                             n0 * blk_d2h[0] + n1 * blk_d2h[-1]], dtype=torch.float64, device=dev)
        counts = [mine]
        my_h2d = n0 * blk_h2d[0] + n1 * blk_h2d[-1]
        h2d_rates = [my_h2d / dt_mine / 1e9]
        gpu_of_rank = [cuda_idx]
        checked_all = e2e_checked
        if world > 1:
            dist.all_reduce(sums)
            cnt = torch.zeros((world, 4), dtype=torch.float64, device=dev)
            cnt[rank, 0], cnt[rank, 1], cnt[rank, 2] = mine, h2d_rates[0], cuda_idx
            cnt[rank, 3] = 1.0 if (e2e_checked or last is None) else 0.0
            dist.all_reduce(cnt)
            rows = cnt.tolist()
            counts = [int(r[0]) for r in rows]
            h2d_rates = [float(r[1]) for r in rows]
            gpu_of_rank = [int(r[2]) for r in rows]
            checked_all = all(r[3] == 1.0 for r in rows)
        e2e_bytes, h2d_tot, d2h_tot = (float(v) for v in sums.tolist())
        e2e = {"value": e2e_bytes / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d_tot / n_e2e / world), "d2h_bytes_per_step": int(d2h_tot / n_e2e / world),
               "sc_kpts_per_s": n_total / dt, "steps": n_e2e,
               "schedule": "SC-kpts pulled one at a time from a counter shared by all ranks "
                           "(bandup_b200.sharding.KptTickets); bytes per step = per-GPU average",
               "kpts_per_rank": counts,
               "h2d_GBps_per_rank": [round(v, 2) for v in h2d_rates],
               "gpu_of_rank": gpu_of_rank, "gpu_order": devinfo.get("order"), "gpu_order_source": devinfo.get("source"),
               "h2d_probe_prefix_total_GBps": devinfo.get("h2d_prefix_total_GBps"), "h2d_solo_GBps": devinfo.get("h2d_solo_GBps"),
               "pci_of_gpu": devinfo.get("pci"),
               "e2e_checked": checked_all,
               "e2e_check": "every rank's last fetched weights / delta_N / n_selected / norms == the device-resident "
                            "entry point's for the same block, bit for bit (outside the timed region)",
               "timer": "host monotonic clock between device synchronisations, max over ranks",
               "bound": "cudaMemcpyAsync H2D of the coefficient blocks over the host PCIe fabric", "numa": numa,
               "pinned_block_probe": {"candidates_GBps": cand_rates, "kept": cand_kept,
                                      "what": "rank 0's page-locked candidate blocks, one timed upload each; the fastest are "
                                              "the job's host buffers (the VM hides the host's NUMA layout)"}}
        if world == 1:
            # the drop-in caller's case: the same call from PAGEABLE memory (a Fortran ALLOCATE), which
            # the library moves through its pinned staging ring; informational, not the headline
            try:
                pg = np.empty_like(host[0][0])
                pg[:] = host[0][0]
                wf_pg = PwWavefunction(pg.view(ctype).reshape(nb, host[0][1]["n_pw"], nsp), host[0][1]["gf"],
                                       host[0][1]["ener"])
                u.submit(20_000, wf_pg, g0=host[0][1]["g0"])
                u.fetch(20_000)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n_pg = 3
                for r_ in range(n_pg):
                    u.submit(20_001 + r_, wf_pg, g0=host[0][1]["g0"])
                    u.fetch(20_001 + r_)
                torch.cuda.synchronize()
                e2e["pageable_host_buffers"] = {"value": n_pg * blk_bytes[0] / (time.perf_counter() - t0) / 1e9,
                                                "unit": UNIT, "sc_kpts": n_pg}
            except Exception as exc:      # never lose the bench line over the extra figure
                e2e["pageable_host_buffers"] = {"error": repr(exc)}
    e2e_file = None
    if world == 1 and not args.no_e2e and not args.no_e2e_file:
        for h, _ in (host if e2e is not None else []):
            pinned_free(h)
        host = []
        try:
            e2e_file = gpu_file_leg(args, sc, u)
        except Exception as exc:          # never lose the bench line over the extra figure
            e2e_file = {"error": repr(exc)}
    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(t_dev0, t_dev1)          # the device-timed region (`value`, `roofline`)
        if e2e is not None:
            e2e["clocks"] = sampler.summary(*t_e2e)

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        os.sched_setaffinity(0, all_cpus)        # the CPU baseline gets every host core again
        gbs, ckps, cores, done, phases = cpu_sample(sc, my_k, args.cpu_seconds, min(kps, 8))
        # the reference author's habit: OMP_NUM_THREADS = n_cores/2 (tutorial run script)
        half = max(1, cores // 2)
        gbs_half, _, _, _, _ = cpu_sample(sc, my_k, args.cpu_seconds / 4, 1, threads=half)
        cpu = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sc_kpts_per_s": ckps,
               "half_cores": {"cores": half, "value": gbs_half},
               "sample": f"{done} full-size SC k-point(s) of the step's batch; C+OpenMP restatement of the "
                         "reference loops (renormalise, select, weights, delta_N)",
               "phase_seconds": {"renormalise": phases[0], "select+weights": phases[1], "delta_N": phases[2]}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": n_steps,
            "steps_requested": args.steps, "timed_region_ms": ms,
            "warmup": n_warm, "ms_per_step": ms / n_steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "c128 in, f64 |C|^2, f64 accumulate" if args.complex128 else "c64 in, f32 |C|^2, f64 accumulate",
            "data": "synthetic",
            "config": workload_config(args, sc, world, kps, [it["n_pw"] for it in items], [it["nf"] for it in items], nener),
            "sc_kpts_per_s": kpts_per_s, "gpu_launches": int(gpu_launches),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_file": e2e_file, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    u.close()
    if world > 1:
        barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from bandup_b200 import synth
    sc = synth.make_scenario(args.workload)
    if args.impl == "reference":
        run_reference(args, sc, rank, world)
        return
    run_b200(args, sc, rank, world, local_rank)


if __name__ == "__main__":
    main()
