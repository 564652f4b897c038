"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/bandup_gpu.h
declares, refuses to run without a device (no CPU fallback), and the Fortran ISO_C_BINDING shim
names exactly those symbols with matching argument counts."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_prototypes():
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "bandup_gpu.h").read_text(), flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(bandup_gpu_\w+)\s*\(([^)]*)\)\s*;", text):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


def test_library_exports_every_header_symbol(gpu_lib):
    from bandup_b200 import _capi
    protos = _header_prototypes()
    assert len(protos) >= 20
    assert sorted(protos) == _capi.header_symbols()
    for name in protos:
        assert hasattr(gpu_lib, name), f"libbandup_gpu.so does not export {name}"
    assert gpu_lib.bandup_gpu_abi_version() == _capi.CONST["BANDUP_GPU_ABI_VERSION"]


def test_ctypes_binding_covers_header(gpu_lib):
    """bandup_b200/_capi.py declares argtypes for every header function with the right arity."""
    protos = _header_prototypes()
    for name, n in protos.items():
        fn = getattr(gpu_lib, name)
        assert fn.argtypes is not None, f"{name} has no ctypes signature"
        assert len(fn.argtypes) == n, f"{name}: header has {n} args, ctypes binding {len(fn.argtypes)}"


def test_no_cpu_fallback_without_device(gpu_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_int(7)
    rc = gpu_lib.bandup_gpu_init(0, C.byref(h))
    assert rc == 9 and h.value == -1                      # BANDUP_GPU_ERR_NO_DEVICE
    assert b"no CPU fallback" in gpu_lib.bandup_gpu_last_error(-1)
    from bandup_b200.unfold import BandupGpuError, GpuUnfolder
    with pytest.raises(BandupGpuError):
        GpuUnfolder(0)


def test_product_never_imports_oracle():
    for p in (ROOT / "bandup_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{p} imports the oracle"
    for p in (ROOT / "bandup_b200" / "csrc").iterdir():
        assert "oracle/" not in p.read_text().replace("oracle/bandup_oracle.c", ""), p


def test_fortran_shim_matches_header():
    f90 = (ROOT / "fortran" / "bandup_gpu_mod.f90").read_text()
    protos = _header_prototypes()
    bound = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*&?\s*\n?\s*bind\(c,\s*name='(\w+)'\)", f90, re.S):
        args = [a for a in re.sub(r"[&\s]", "", m.group(2)).split(",") if a]
        assert m.group(1) == m.group(3)
        bound[m.group(3)] = len(args)
    assert len(bound) >= 18
    for name, n in bound.items():
        assert name in protos, f"Fortran shim binds {name}, which the header does not declare"
        assert protos[name] == n, f"{name}: header has {protos[name]} args, Fortran interface {n}"
    # the drop-in entry points must all be bound
    for name in ["bandup_gpu_init", "bandup_gpu_set_cells", "bandup_gpu_set_grid", "bandup_gpu_folding_g0",
                 "bandup_gpu_submit_kpt", "bandup_gpu_fetch_kpt", "bandup_gpu_finalize",
                 "bandup_gpu_rho_count", "bandup_gpu_fetch_rho"]:
        assert name in bound, name
    consts = dict(re.findall(r"integer\(c_int\), parameter :: (\w+) = (\d+)", f90))
    from bandup_b200 import _capi
    for k, v in consts.items():
        assert _capi.CONST[k] == int(v), k
