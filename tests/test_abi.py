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


def _c_args(name):
    """[(base type, is_pointer)] of a header prototype."""
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "bandup_gpu.h").read_text(), flags=re.S)
    m = re.search(r"\b" + name + r"\s*\(([^)]*)\)\s*;", text)
    out = []
    for a in m.group(1).split(","):
        a = a.strip()
        if a in ("", "void"):
            continue
        ptr = a.count("*")
        a = re.sub(r"\bconst\b|\*", " ", a)
        toks = a.split()
        base = " ".join(toks[:-1]) if len(toks) > 1 else toks[0]    # drop the parameter name
        out.append((base.strip(), ptr))
    return out


def _fortran_interfaces():
    """{c name: [(arg, fortran type text, has value attribute)]} of every bind(c) interface."""
    f90 = (ROOT / "fortran" / "bandup_gpu_mod.f90").read_text()
    f90 = re.sub(r"&\s*\n\s*", " ", f90)                            # join continuation lines
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\(([^)]*)\)\s*bind\(c,\s*name='(\w+)'\)(.*?)end function \1", f90, re.S):
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        decl = {}
        for line in m.group(4).splitlines():
            line = line.split("!")[0].strip()
            if "::" not in line:
                continue
            left, right = line.split("::", 1)
            for var in re.split(r",(?![^()]*\))", right):
                v = re.sub(r"\(.*\)", "", var).strip()
                decl[v] = (left.strip(), "(" in var)
        out[m.group(3)] = [(a,) + decl[a] for a in args]
    return out


C_TO_F = {"int": "integer(c_int)", "unsigned": "integer(c_int)", "int32_t": "integer(c_int32_t)",
          "int64_t": "integer(c_int64_t)", "size_t": "integer(c_size_t)", "double": "real(c_double)",
          "float": "real(c_float)", "uint8_t": "integer(c_int8_t)", "uint64_t": "integer(c_int64_t)"}


def test_fortran_shim_argument_types_and_passing_convention():
    """The shim cannot be compiled here (no Fortran front end), so every dummy argument is checked
    against the C prototype: scalars the C side takes by value carry `value` and the matching kind;
    pointers are either by-reference dummies of the matching kind (scalars or assumed-size arrays)
    or `type(c_ptr), value`."""
    n_checked = 0
    for name, fargs in _fortran_interfaces().items():
        cargs = _c_args(name)
        assert len(cargs) == len(fargs), name
        for (cbase, cptr), (fname, ftype, is_array) in zip(cargs, fargs):
            has_value = bool(re.search(r"\bvalue\b", ftype))
            kind = ftype.split(",")[0].replace(" ", "")
            where = f"{name}({fname}): C `{cbase}{' *' if cptr else ''}` vs Fortran `{ftype}`"
            if kind == "type(c_ptr)":
                # a C pointer handed over as an address (T*), or a pointer returned through T**
                assert (cptr == 1 and has_value) or (cptr == 2 and not has_value), where
            elif not cptr:
                assert has_value and not is_array, where             # by-value scalar
                assert kind == C_TO_F[cbase].replace(" ", ""), where
            else:
                assert not has_value, where                          # by reference: Fortran passes the address
                if cbase in ("void", "char"):
                    continue
                if cbase == "bandup_gpu_kpt_desc":
                    continue
                assert kind == C_TO_F[cbase].replace(" ", ""), where
            n_checked += 1
    assert n_checked > 150


def test_ctypes_binding_argument_types(gpu_lib):
    """Width and kind of every ctypes argument / return type against the C prototype."""
    byval = {"int": C.c_int, "unsigned": C.c_uint, "int32_t": C.c_int32, "int64_t": C.c_int64,
             "size_t": C.c_size_t, "double": C.c_double, "uint64_t": C.c_uint64}
    pointee = {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double,
               "float": C.c_float, "uint8_t": C.c_uint8}
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "bandup_gpu.h").read_text(), flags=re.S)
    n = 0
    for name in _header_prototypes():
        fn = getattr(gpu_lib, name)
        ret = re.search(r"(?:^|\n)((?:const\s+)?[\w]+(?:\s*\*)?)\s*\b" + name + r"\s*\(", text).group(1).strip()
        ret = re.sub(r"\s*\*", " *", ret)
        want_ret = {"int": C.c_int, "size_t": C.c_size_t, "int64_t": C.c_int64, "const char *": C.c_char_p}[ret]
        assert fn.restype is want_ret, f"{name}: returns {ret}, ctypes says {fn.restype}"
        for i, ((cbase, cptr), at) in enumerate(zip(_c_args(name), fn.argtypes)):
            where = f"{name} arg {i}: C `{cbase}{'*' * cptr}` vs ctypes {at}"
            if not cptr:
                assert at is byval[cbase], where
            elif at in (C.c_void_p, C.c_char_p):
                pass                                                  # untyped address
            else:
                assert hasattr(at, "_type_"), where                   # POINTER(T)
                if cptr == 2:
                    assert at._type_ is C.c_void_p, where
                elif cbase in pointee:
                    assert C.sizeof(at._type_) == C.sizeof(pointee[cbase]), where
                    assert (at._type_ in (C.c_float, C.c_double)) == (cbase in ("float", "double")), where
            n += 1
    assert n > 200
