"""The C-ABI shared library: loads without a GPU, exports every symbol include/nemo_fct.h declares, mirrors the
struct layout, and FAILS LOUDLY (no CPU fallback) when a compute entry point is used without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "nemo_fct.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nemo_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(N):
    names = declared_functions()
    assert len(names) >= 25
    L = C.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "declared in nemo_fct.h but not exported: " + n
    assert sorted(N.ABI_SYMBOLS) == names, "python binding and header disagree"
    assert N.lib().nemo_fct_abi_version() == 1


def test_struct_layout_matches_header(N):
    src = open(os.path.join(ROOT, "include", "nemo_fct.h")).read()
    body = re.search(r"typedef struct nemo_fct_domain \{(.*?)\} nemo_fct_domain;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        assert decl.startswith("int ")
        for v in decl[4:].split(","):
            fields.append(v.strip().split("[")[0])
    assert fields == [n for n, _ in N.Domain._fields_]
    assert C.sizeof(N.Domain) == 4 * (len(fields) - 1 + N.JPMAXNGH)


def test_no_cpu_fallback_create_fails_without_gpu(N):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    dom = N.mpp_init(20, 16, 5, 0)
    with pytest.raises(N.NemoFctError, match="no usable CUDA device|CPU"):
        N.FctContext(dom, 0)


def test_create_rejects_inconsistent_domain(N):
    dom = N.mpp_init(40, 30, 5, 4, 2, 2, 3)
    dom.nimpp += 1                     # a host whose mpp_init disagrees with the library's decomposition
    h = C.c_void_p()
    rc = N.lib().nemo_fct_create(C.byref(dom), 0, C.byref(h))
    assert rc != 0 and b"nimpp" in N.lib().nemo_fct_last_error()


def test_null_handle_is_an_error_not_a_crash(N):
    L = N.lib()
    a = np.zeros(8)
    p = a.ctypes.data_as(C.c_void_p)
    assert L.nemo_tra_adv_fct(None, 1, 1, b"TRA", 1.0, p, p, p, p, p, p, 1, 2, 2) != 0
    assert L.nemo_fct_set_e3t(None, p, p, p, 0) != 0
    assert L.nemo_fct_synchronize(None) != 0
    assert b"NULL" in L.nemo_fct_last_error()


def test_oracle_is_not_linked_into_the_product(N):
    """the shipped library must not contain or depend on the oracle"""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for sym in ("tra_adv_fct", "nonosc", "interp_4th_cpt", "oce_world_run", "lbc_lnk_generic", "mpp_lnk_generic", "mpp_init"):
        assert sym not in exported
    ldd = subprocess.run(["ldd", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd


def test_header_is_plain_c_and_the_c_example_links(N, tmp_path):
    """include/nemo_fct.h is C99 (a Fortran / C host binds it, no C++ or torch types), examples/fct_step.c builds against the
    shared library with nothing but gcc, and on a machine without a GPU it stops at nemo_fct_create with a message."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "nemo_fct.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    exe = str(tmp_path / "fct_step")
    libdir = os.path.dirname(N.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "fct_step.c"), "-L", libdir, "-lnemo_fct", "-Wl,-rpath," + libdir, "-o", exe])
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU path" in r.stderr, (r.returncode, r.stderr)
