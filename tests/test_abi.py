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


def _crack_shim():
    """The shim as parsed by NumPy's Fortran front end (numpy.f2py.crackfortran: a real free-form
    Fortran 90 parser, independent of the regular expressions used above)."""
    import contextlib
    import io
    cf = pytest.importorskip("numpy.f2py.crackfortran")
    cf.verbose = 0
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
        blocks = cf.crackfortran([str(ROOT / "fortran" / "bandup_gpu_mod.f90")])
    return blocks


def test_fortran_shim_parses_as_fortran():
    """No Fortran compiler here, but a Fortran parser: the module must come out as one module with
    one interface block and the two contained procedures, every bind(c) function with a typed
    result and every dummy argument declared with the kind and passing convention the C prototype
    asks for."""
    blocks = _crack_shim()
    assert [b["block"] for b in blocks] == ["module"] and blocks[0]["name"] == "bandup_gpu"
    body = blocks[0]["body"]
    assert [b["block"] for b in body] == ["interface", "function", "subroutine"]
    assert [b["name"] for b in body[1:]] == ["bandup_gpu_error_message", "bandup_gpu_unfold_kpt_f"]
    protos = _header_prototypes()
    lower = {k.lower(): k for k in protos}                 # Fortran names are case-insensitive
    kinds = {"int": ("integer", "c_int"), "unsigned": ("integer", "c_int"), "int32_t": ("integer", "c_int32_t"),
             "int64_t": ("integer", "c_int64_t"), "size_t": ("integer", "c_size_t"),
             "double": ("real", "c_double"), "float": ("real", "c_float"), "uint8_t": ("integer", "c_int8_t"),
             "uint64_t": ("integer", "c_int64_t")}
    n_args = 0
    funcs = body[0]["body"]
    assert len(funcs) == len({f["name"] for f in funcs}) >= 40
    for f in funcs:
        assert f["block"] == "function", f["name"]
        cname = lower[f["name"]]
        res = f["vars"][f["name"]]
        assert res["typespec"] in ("integer", "type"), cname      # int / size_t / int64_t status, or const char*
        cargs = _c_args(cname)
        assert len(cargs) == len(f["args"]) == protos[cname], cname
        for (cbase, cptr), a in zip(cargs, f["args"]):
            v = f["vars"][a]
            assert "typespec" in v, f"{cname}: dummy argument {a} is not declared"
            by_value = "value" in v.get("attrspec", [])
            where = f"{cname}({a}): C `{cbase}{'*' * cptr}` vs {v}"
            if v["typespec"] == "type":
                assert v["typename"] == "c_ptr", where
                assert (cptr == 1 and by_value) or (cptr == 2 and not by_value), where
            elif cptr == 0:
                assert by_value and "dimension" not in v, where
                assert (v["typespec"], v["kindselector"]["kind"]) == kinds[cbase], where
            else:
                assert not by_value, where
                if cbase in kinds:
                    assert (v["typespec"], v["kindselector"]["kind"]) == kinds[cbase], where
            n_args += 1
    assert n_args > 150


def test_fortran_interface_bodies_import_what_they_use():
    """An interface body does not see the host's `use iso_c_binding`: every c_* kind or type a
    body names must be on its own `import` line (a classic compile error otherwise)."""
    f90 = re.sub(r"&\s*\n\s*", " ", (ROOT / "fortran" / "bandup_gpu_mod.f90").read_text())
    iface = f90[f90.index("\ninterface"):f90.index("\nend interface")]
    n = 0
    for m in re.finditer(r"function\s+(\w+)\s*\([^)]*\)\s*bind\(c,[^)]*\)(.*?)end function \1", iface, re.S):
        text = "\n".join(l.split("!")[0] for l in m.group(2).splitlines())
        imp = re.search(r"^\s*import\s+(?:::)?\s*(.*)$", text, re.M)
        assert imp, f"{m.group(1)} has no import statement"
        imported = {t.strip().lower() for t in imp.group(1).split(",")}
        used = {t.lower() for t in re.findall(r"\bc_\w+", text.replace(imp.group(0), ""))}
        assert used <= imported, f"{m.group(1)} uses {sorted(used - imported)} without importing them"
        assert imported <= used, f"{m.group(1)} imports {sorted(imported - used)} and never uses them"
        n += 1
    assert n >= 40


def test_fortran_shim_source_form():
    """Free-form limits a compiler enforces: lines of at most 132 characters, at most 39
    continuation lines per statement (Fortran 95; 255 from 2003 on), balanced block ends."""
    lines = (ROOT / "fortran" / "bandup_gpu_mod.f90").read_text().splitlines()
    assert max(len(l) for l in lines) <= 132
    run = 0
    for l in lines:
        code = l.split("!")[0].rstrip()
        run = run + 1 if code.endswith("&") else 0
        assert run <= 39
    code = "\n".join(l.split("!")[0] for l in lines).lower()
    for opener, closer in [(r"^\s*function\s+\w+", r"^\s*end function"), (r"^\s*subroutine\s+\w+", r"^\s*end subroutine"),
                           (r"^\s*interface\s*$", r"^\s*end interface"), (r"^\s*module\s+\w+", r"^\s*end module"),
                           (r"\bthen\s*$", r"^\s*end\s*if\b"), (r"^\s*do\b", r"^\s*end\s*do\b")]:
        assert len(re.findall(opener, code, re.M)) == len(re.findall(closer, code, re.M)), opener
