"""The C ABI from a plain C host (tests/abi_driver.c): the call sequence of the patched BandUP main
loop, compiled with gcc and linked against libbandup_gpu.so -- no Python/torch in the process."""
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from bandup_b200 import synth

ROOT = Path(__file__).resolve().parent.parent


def _build_driver(tmp_path, gpu_lib):
    exe = tmp_path / "abi_driver"
    libdir = ROOT / "bandup_b200" / "lib"
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests" / "abi_driver.c"),
           "-L", str(libdir), "-lbandup_gpu", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    import os
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(cmd, check=True, capture_output=True, env=env)
    return exe


def _write_input(path, sc, ks):
    from oracle import oracle as O
    grid = O.real_seq(*sc.energy_window())
    data = []
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", len(ks), len(grid)))
        f.write(np.ascontiguousarray(sc.B_sc.T).tobytes())      # Fortran memory order
        f.write(np.ascontiguousarray(sc.b_pc.T).tobytes())
        f.write(np.array(sc.energy_window(), dtype="<f8").tobytes())
        for iK in ks:
            coeffs, gf, ener = synth.make_wavefunction(sc, iK)
            fG = sc.folding_G(iK)
            nb, n_pw, nsp = coeffs.shape
            f.write(struct.pack("<iiii", nsp, n_pw, nb, fG.shape[0]))
            f.write(np.ascontiguousarray(fG, dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(gf, dtype="<i4").tobytes())
            f.write(np.ascontiguousarray(ener, dtype="<f8").tobytes())
            f.write(np.ascontiguousarray(coeffs, dtype="<c8").tobytes())
            data.append((coeffs, gf, ener, fG))
    return grid, data


def test_c_driver_compiles_links_and_refuses_without_gpu(tmp_path, gpu_lib):
    import torch
    exe = _build_driver(tmp_path, gpu_lib)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the run is covered by the gpu test")
    sc = synth.make_scenario("graphene4x4", scale=0.05, n_kpts=6)
    _write_input(tmp_path / "in.bin", sc, [0])
    r = subprocess.run([str(exe), str(tmp_path / "in.bin")], capture_output=True, text=True)
    assert r.returncode == 10 + 9                                # BANDUP_GPU_ERR_NO_DEVICE: no CPU fallback
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_matches_oracle(tmp_path, gpu_lib):
    from oracle import oracle as O
    exe = _build_driver(tmp_path, gpu_lib)
    sc = synth.make_scenario("si333_vacancy", scale=0.2, n_kpts=10)
    ks = [0, 1, 2, 3, 4]
    grid, data = _write_input(tmp_path / "in.bin", sc, ks)
    r = subprocess.run([str(exe), str(tmp_path / "in.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0].split()[1:] == [str(v) for v in sc.M.T.reshape(-1)]     # int_M in Fortran order
    got = {}
    for ln in lines[1:]:
        t = ln.split()
        if t[0] != "K":
            continue
        key = (int(t[1]), int(t[3]))
        if t[4] == "nsel":
            got.setdefault(key, {})["nsel"] = int(t[5])
            got[key]["W"] = np.array(t[7:], dtype=float)
        else:
            got.setdefault(key, {})["dN"] = np.array(t[5:], dtype=float)
    for i, (coeffs, gf, ener, fG) in enumerate(data):
        w, dN, ns, _ = O.unfold_kpt(coeffs, gf @ sc.B_sc, ener, sc.b_pc, fG, grid, norm_mode=1)
        for f in range(fG.shape[0]):
            g = got[(i + 1, f)]
            assert g["nsel"] == ns[f]
            assert np.max(np.abs(g["W"] - w[f]) / np.maximum(np.abs(w[f]), 1e-12)) <= 1e-6
            assert np.max(np.abs(g["dN"] - dN[f]) / np.maximum(np.abs(dN[f]), 1e-12)) <= 1e-6
    assert int(lines[-1].split()[1]) == 3 * len(ks)              # K1, K2, K3 per submit
