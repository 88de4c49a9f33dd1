"""Static ABI check of the Fortran drop-in modules (nemo-fmi-devel_b200/shim/*.F90) against include/nemo_fct.h.  There is no
Fortran compiler in the image, so the INTERFACE blocks are parsed here: every BIND(C) function must exist in the header with the
same number of arguments, in the same order, passed the same way (VALUE for C scalars and handles, by reference for C pointers,
matching base types), and TYPE, BIND(C) :: nemo_fct_domain must list the fields of struct nemo_fct_domain in the same order.
When a Fortran compiler is present, the modules are also syntax-checked against stub modules (skipped otherwise)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nemo_fct.h")
SHIMS = [os.path.join(ROOT, "nemo-fmi-devel_b200", "shim", n) for n in ("traadv_fct_gpu.F90", "traadv_mus_gpu.F90")]


def _c_source():
    s = open(HEADER).read()
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    s = re.sub(r"//[^\n]*", " ", s)
    return s


def _c_kind(param):
    """how one C parameter is passed: (by_value, base)"""
    p = " ".join(param.replace("const", " ").split())
    if p in ("void", ""):
        return None
    stars = p.count("*")
    base = re.sub(r"[\*\s]+\w*$", "", p) if stars else p.rsplit(" ", 1)[0]
    base = base.replace("*", "").strip()
    base = {"unsigned char": "char", "long long": "longlong", "unsigned long long": "longlong"}.get(base, base)
    if base == "nemo_fct_handle":
        return (stars == 0, "handle")
    return (stars == 0, base)


def c_prototypes():
    out = {}
    for m in re.finditer(r"\b(int|long long|const char \*|void)\s*(nemo_\w+)\s*\(([^;{]*?)\)\s*;", _c_source(), flags=re.S):
        params = [x.strip() for x in m.group(3).split(",")]
        out[m.group(2)] = [k for k in (_c_kind(x) for x in params) if k is not None]
    return out


def c_struct_fields(name):
    m = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), _c_source(), flags=re.S)
    fields = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for item in decl.split(None, 1)[1].split(","):
            fields.append(re.sub(r"\[.*\]", "", item).strip())
    return fields


def _fortran_lines(path):
    """free-form source with comments stripped, continuation lines joined, preprocessor lines dropped"""
    out, cur = [], ""
    for raw in open(path):
        line = raw.split("!")[0].rstrip()
        if line.lstrip().startswith("#") or not line.strip():
            continue
        line = line.strip()
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1].rstrip() + " "
            continue
        out.append(cur + line)
        cur = ""
    return out


def fortran_interfaces(path):
    """{c_name: [(by_value, base), ...]} of every BIND(C) function of the INTERFACE blocks"""
    lines = _fortran_lines(path)
    out, i = {}, 0
    while i < len(lines):
        m = re.match(r"(?:INTEGER\(C_INT\)\s+|TYPE\(C_PTR\)\s+)?FUNCTION\s+(\w+)\s*\(([^)]*)\)\s*BIND\(C,\s*NAME='(\w+)'\)", lines[i], flags=re.I)
        if not m:
            i += 1
            continue
        fname, dummies, cname = m.group(1), [a.strip().lower() for a in m.group(2).split(",") if a.strip()], m.group(3)
        decl = {}
        i += 1
        while not re.match(r"END FUNCTION", lines[i], flags=re.I):
            if "::" in lines[i] and not lines[i].upper().startswith("IMPORT"):
                spec, names = lines[i].split("::", 1)
                up = spec.upper()
                base = ("handle" if "TYPE(C_PTR)" in up else "int" if "INTEGER(C_INT)" in up else "double" if "REAL(C_DOUBLE)" in up
                        else "char" if "CHARACTER" in up else "size_t" if "C_SIZE_T" in up
                        else re.search(r"TYPE\((\w+)\)", spec, flags=re.I).group(1))
                for n in names.split(","):
                    decl[n.strip().lower()] = ("VALUE" in up, base)
            i += 1
        assert fname.lower() == cname.lower()
        out[cname] = [decl[d] for d in dummies]
    return out


def _compatible(f, c):
    fv, fb = f
    cv, cb = c
    if cb == "handle":                                   # nemo_fct_handle by value <-> TYPE(C_PTR), VALUE ; nemo_fct_handle * <-> TYPE(C_PTR)
        return fb == "handle" and fv == cv
    if cv:                                               # C scalar by value
        return fv and fb == cb
    if fb == "handle":                                   # any C pointer may be bound as TYPE(C_PTR), VALUE
        return fv
    return (not fv) and (fb == cb or (cb == "void" and fb in ("char", "double", "int")))


@pytest.mark.parametrize("shim", SHIMS, ids=[os.path.basename(p) for p in SHIMS])
def test_every_bound_function_matches_its_c_prototype(shim):
    protos = c_prototypes()
    bound = fortran_interfaces(shim)
    assert bound, "no BIND(C) interface found in " + shim
    for cname, fargs in bound.items():
        assert cname in protos, "%s binds %s, which include/nemo_fct.h does not declare" % (os.path.basename(shim), cname)
        cargs = protos[cname]
        assert len(fargs) == len(cargs), (cname, len(fargs), len(cargs))
        for n, (f, c) in enumerate(zip(fargs, cargs)):
            assert _compatible(f, c), "%s argument %d: Fortran %s vs C %s" % (cname, n + 1, f, c)


def test_the_parser_sees_what_it_should():
    protos = c_prototypes()
    assert protos["nemo_tra_adv_fct"] == [(True, "handle"), (True, "int"), (True, "int"), (False, "char"), (True, "double")] + \
        [(False, "double")] * 6 + [(True, "int")] * 3
    assert protos["nemo_fct_create"] == [(False, "nemo_fct_domain"), (True, "int"), (False, "handle")]
    bound = fortran_interfaces(SHIMS[0])
    assert {"nemo_fct_create", "nemo_fct_set_domain_arrays", "nemo_fct_set_e3t", "nemo_tra_adv_fct", "nemo_interp_4th_cpt",
            "nemo_fct_comm_unique_id", "nemo_fct_comm_init", "nemo_fct_last_error"} <= set(bound)
    assert not _compatible((False, "int"), (True, "int")) and not _compatible((True, "double"), (False, "double"))


def test_domain_descriptor_has_the_c_field_order():
    cf = c_struct_fields("nemo_fct_domain")
    lines = _fortran_lines(SHIMS[0])
    a = next(i for i, l in enumerate(lines) if re.match(r"TYPE,\s*BIND\(C\)\s*::\s*nemo_fct_domain", l, flags=re.I))
    b = next(i for i in range(a, len(lines)) if re.match(r"END TYPE", lines[i], flags=re.I))
    ff = []
    for l in lines[a + 1:b]:
        assert l.upper().startswith("INTEGER(C_INT)"), l          # the C struct holds nothing but int
        ff += [re.sub(r"\(.*\)", "", n).strip() for n in l.split("::", 1)[1].split(",")]
    assert [x.lower() for x in ff] == [x.lower() for x in cf]
    src = open(SHIMS[0]).read()
    jp = int(re.search(r"#define NEMO_FCT_JPMAXNGH (\d+)", open(HEADER).read()).group(1))
    assert re.search(r"isendto\(%d\)" % jp, src)
    nid = int(re.search(r"#define NEMO_FCT_UNIQUE_ID_BYTES (\d+)", open(HEADER).read()).group(1))
    assert "DIMENSION(%d)" % nid in src


def test_mpi_symbols_are_visible_under_key_mpp_mpi():
    """lib_mpp is PRIVATE (lib_mpp.F90:63,116): MPI_BCAST / MPI_CHARACTER need the module's own mpif.h"""
    src = open(SHIMS[0]).read()
    assert "MPI_BCAST" in src.upper()
    guard = re.search(r"#if defined key_mpp_mpi\s*\n(?:\s*!.*\n)*\s*INCLUDE 'mpif.h'\s*\n#endif", src)
    assert guard, "INCLUDE 'mpif.h' under key_mpp_mpi is missing"
    assert src.index("IMPLICIT NONE") < guard.start() < src.upper().index("CONTAINS")


@pytest.mark.skipif(not any(shutil.which(c) for c in ("gfortran", "flang", "ifx", "nvfortran")), reason="no Fortran compiler in this image")
def test_shims_pass_a_syntax_check_with_stub_modules(tmp_path):
    fc = next(c for c in ("gfortran", "flang", "ifx", "nvfortran") if shutil.which(c))
    stubs = os.path.join(ROOT, "oracle", "_ref_recipe", "stubs.F90")
    cmd = [fc, "-cpp", "-fsyntax-only", "-J", str(tmp_path), stubs] + SHIMS
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-4000:]
