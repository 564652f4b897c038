"""Every `file:line[-line]` citation of the reference in the public header, the design documents,
the kernels, the host mirrors and the oracle must point inside an existing reference file.  Runs
only where /root/reference is mounted (the build container); skipped elsewhere."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

FILES = ["include/bandup_gpu.h", "DESIGN.md", "INTEGRATION.md", "fortran/bandup_gpu_mod.f90",
         "oracle/bandup_oracle.c", "oracle/oracle_np.py", "bandup_b200/unfold.py", "bandup_b200/wavecar.py",
         "bandup_b200/bandup_files.py", "bandup_b200/sharding.py", "bandup_b200/host/bandup_unfold_gpu.cpp"] + \
        [str(p.relative_to(ROOT)) for p in (ROOT / "bandup_b200" / "csrc").glob("*.cu*")]

PAT = re.compile(r"([A-Za-z_][\w/.-]*\.(?:f90|py))[:`]*:(\d+)(?:-(\d+))?")


def _index():
    idx = {}
    for p in REF.rglob("*"):
        if p.suffix in (".f90", ".py") and p.is_file():
            idx.setdefault(p.name, []).append(p)
    return idx


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is not mounted here")
def test_reference_citations_point_into_existing_files():
    idx = _index()
    n = 0
    bad = []
    for rel in FILES:
        text = (ROOT / rel).read_text()
        for m in PAT.finditer(text):
            name, lo, hi = Path(m.group(1)).name, int(m.group(2)), int(m.group(3) or m.group(2))
            if name not in idx:
                continue                      # a file of this repository (e.g. tests/..py:12), not a citation
            n += 1
            n_lines = max(len(p.read_text(errors="replace").splitlines()) for p in idx[name])
            if not (1 <= lo <= hi <= n_lines):
                bad.append(f"{rel}: {m.group(0)} (file has {n_lines} lines)")
    assert n > 150, n
    assert not bad, "\n".join(bad)
