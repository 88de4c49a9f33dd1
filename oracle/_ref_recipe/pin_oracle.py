"""Pin the oracle against the reference's own Fortran (see README.md): dump inputs, run oracle/_ref/fct_ref_driver (the unmodified
traadv_fct.F90), compare pta with oracle/traadv_fct.c bit for bit, and write tests/golden/fortran_pin.json when it agrees."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "fct_ref_driver")
    if not os.path.exists(exe):
        print("pin_oracle: %s not built (no gfortran or no NEMO tree): no compiled-binary pin (the source-text pin is tests/test_cpu_reference_exec.py)" % exe)
        return 0
    import helpers as H
    from oracle import oracle as O
    cases, ok_all = [], True
    for (G, GJ, K, jperio, kjpt, h, v, lin, isf) in ((30, 22, 11, 0, 2, 2, 2, False, False), (34, 27, 9, 4, 2, 4, 4, False, False),
                                                    (36, 24, 12, 6, 3, 4, 2, True, True), (40, 30, 10, 1, 2, 2, 4, True, False)):
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt=kjpt, seed=7000 + jperio, ln_linssh=lin, ln_isfcav=isf)
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v, ln_linssh=lin, ln_isfcav=isf, key_mpp_mpi=False)
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
            with open(fin, "wb") as f:
                np.array([G, GJ, K, jperio, kjpt, h, v, int(lin), int(isf), 0], np.int32).tofile(f)
                np.array([gf["p2dt"]], np.float64).tofile(f)
                for k in ("tmask", "umask", "vmask", "wmask", "e3t_b", "e3t_n", "e3t_a", "e1e2t", "r1_e1e2t"):
                    np.ascontiguousarray(gf[k], np.float64).tofile(f)
                for k in ("mikt", "mbkt"):
                    np.ascontiguousarray(gf[k], np.int32).tofile(f)
                for k in ("pun", "pvn", "pwn", "ptb", "ptn", "pta"):
                    np.ascontiguousarray(gf[k], np.float64).tofile(f)
            subprocess.check_call([exe, fin, fout])
            got = np.fromfile(fout, np.float64).reshape(ref.shape)
        inner = (slice(None), slice(0, K - 1), slice(1, -1), slice(1, -1))
        same = bool(np.array_equal(got[inner], ref[inner]))
        rel = float(np.max(np.abs(got[inner] - ref[inner]) / np.maximum(np.abs(ref[inner]), 1e-300)))
        print("jperio=%d h%d/v%d linssh=%s isfcav=%s: %s (max rel diff %.3e)" % (jperio, h, v, lin, isf, "BIT-IDENTICAL" if same else "DIFFERS", rel))
        cases.append(dict(G=G, GJ=GJ, K=K, jperio=jperio, kjpt=kjpt, h=h, v=v, ln_linssh=lin, ln_isfcav=isf, identical=same, max_rel=rel))
        ok_all = ok_all and same
    if ok_all:
        with open(os.path.join(ROOT, "tests", "golden", "fortran_pin.json"), "w") as f:
            json.dump({"what": "oracle/traadv_fct.c == unmodified traadv_fct.F90 (gfortran, arch-linux_gfortran.fcm flags)", "cases": cases}, f, indent=1)
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
