"""The host staging pipeline of the pageable upload path (bandup_b200/csrc/stage_pipeline.h: pageable
block -> pieces copied by a thread pool into a ring of page-locked chunks -> DMA) against a software
wire, on the CPU: every byte arrives in order for ragged sizes, misaligned sources, rings of one to
four buffers, one to four threads, a wire slower than the copiers, back-to-back blocks on one ring,
and a failing send aborts the block without hanging the pool (tests/stage_pipeline_driver.cpp).
The same code moves real coefficient blocks in the `-m gpu` tests (pageable submits)."""
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_staging_pipeline_against_a_software_wire(tmp_path):
    exe = tmp_path / "stage_pipeline_driver"
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-pthread",
                    str(ROOT / "tests" / "stage_pipeline_driver.cpp"), "-o", str(exe)], check=True, capture_output=True, env=env)
    r = subprocess.run([str(exe), "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.strip().endswith("0 failures") and int(r.stdout.split()[0]) > 500
