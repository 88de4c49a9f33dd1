"""BASELINE config C3 (ORCA025-like 1442x1207x75, T+S, FCT 4th order + compact vertical, T-pivot fold; the headline benchmark) cannot be
run whole through the translated reference (about 6 hours); this samples it.  The full-size ORACLE result is computed (4x2 ranks,
~48 GB of host memory, half a minute) and, for a few 48 x 40-column windows cut out of the actual global fields, the REFERENCE's
tra_adv_fct text is executed on the window as a closed domain.  tra_adv_fct at a cell depends on inputs within 3 cells of it (and on
the whole column), so 6 cells inside the window the artificial boundary is not felt: there the reference's result must equal the
full-size oracle's, bit for bit.  The same for the northernmost 16 rows over the whole width, run as a jperio = 4 domain of their own
(east-west seam and T-pivot fold at the real size).  tests/test_gpu_full_size.py holds the CUDA path to the same full-size oracle result.

    python tests/golden/c3_windows.py          # writes the "c3_windows" entry of tests/golden/ref_exec_pins.json (8 minutes)"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H                      # noqa: E402
import golden_cases as GC                # noqa: E402
from oracle import oracle as O           # noqa: E402
from oracle import ref_exec as R         # noqa: E402

WINDOWS = [(100, 150), (700, 600), (1300, 1100), (400, 1150), (1000, 30)]     # 0-based (i0, j0) corners
WI, WJ, MARGIN = 48, 40, 6
BAND = 16                               # the northernmost rows, full width: east-west cyclic seam and T-pivot fold at the real size


def run():
    BF = importlib.import_module("nemo-fmi-devel_b200.bench_fields")
    G, GJ, K, jperio, kjpt, h, v, rdt = BF.CONFIGS["orca025"]
    gf = H.global_bench_fields(O, G, GJ, K, jperio, kjpt, rn_rdt=rdt, cfl=0.25)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 4, 2, kjpt, h, v, poison=False)
    res = {}
    for (i0, j0) in WINDOWS:
        g = {k: np.ascontiguousarray(a[..., j0:j0 + WJ, i0:i0 + WI]) for k, a in gf.items() if isinstance(a, np.ndarray) and a.ndim >= 2}
        g["p2dt"] = gf["p2dt"]
        out = R.tra_adv_fct(g, WI, WJ, K, kjpt, h, v, False, False, R.reference_lbc(0, WI, WJ))
        a = np.ascontiguousarray(out[..., MARGIN:-MARGIN, MARGIN:-MARGIN])
        b = np.ascontiguousarray(ref[..., j0 + MARGIN:j0 + WJ - MARGIN, i0 + MARGIN:i0 + WI - MARGIN])
        res["%d_%d" % (i0, j0)] = {"equal_to_full_size_oracle": bool(np.array_equal(a.view(np.uint64), b.view(np.uint64))),
                                   "interior_sha256": GC.digest(a),
                                   "changed": bool(not np.array_equal(a, g["pta"][..., MARGIN:-MARGIN, MARGIN:-MARGIN]))}
    # the north-fold band: the last BAND rows of the actual fields over the whole width, as a jperio = 4 domain of its own (cyclic
    # east-west, T-pivot fold in the north, an artificial closed boundary in the south): MARGIN rows north of that boundary the
    # reference's result must again be the full-size oracle's -- across the seam and under the fold
    g = {k: np.ascontiguousarray(a[..., GJ - BAND:, :]) for k, a in gf.items() if isinstance(a, np.ndarray) and a.ndim >= 2}
    g["p2dt"] = gf["p2dt"]
    out = R.tra_adv_fct(g, G, BAND, K, kjpt, h, v, False, False, R.reference_lbc(jperio, G, BAND))
    a = np.ascontiguousarray(out[..., MARGIN:, :])
    b = np.ascontiguousarray(ref[..., GJ - BAND + MARGIN:, :])
    res["fold_band_last_%d_rows" % BAND] = {"equal_to_full_size_oracle": bool(np.array_equal(a.view(np.uint64), b.view(np.uint64))),
                                            "interior_sha256": GC.digest(a), "changed": bool(not np.array_equal(a, g["pta"][..., MARGIN:, :]))}
    return {"config": "orca025 (C3) BENCH fields, FCT h4/v4; windows %dx%d columns x %d levels cut out of the global fields, run as closed "
                      "domains by the reference's text, interiors %d cells inside the window compared with the full-size oracle (4x2 ranks)"
                      % (WI, WJ, K, MARGIN), "input_sha256": GC.input_hash(gf, {}), "windows": res}


if __name__ == "__main__":
    assert R.available(), "the reference tree is not on this machine"
    out = run()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_exec_pins.json")
    pins = json.load(open(path))
    pins["c3_windows"] = out
    with open(path, "w") as f:
        json.dump(pins, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1))
