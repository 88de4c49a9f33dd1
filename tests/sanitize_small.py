"""Small all-schedules run for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tests/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H          # noqa: E402
import nemo_fct_b200 as N    # noqa: E402
from oracle import oracle as O   # noqa: E402

G, GJ, K = 76, 45, 7
ok = True
for jperio, (h, v) in ((4, (4, 4)), (0, (2, 2))):
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=7)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
    for schedule in (0, 1, 2, 3):
        got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 1, 1, 2, h, v, schedule=schedule)
        same = bool(np.array_equal(got, ref))
        print("jperio", jperio, "h/v", h, v, "schedule", schedule, "bit-identical" if same else "MISMATCH", flush=True)
        ok = ok and same
    got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 2, 2, 2, h, v, schedule=2)
    print("jperio", jperio, "2x2 in-process group", bool(np.array_equal(got, ref)), flush=True)
    ok = ok and bool(np.array_equal(got, ref))
sys.exit(0 if ok else 1)
