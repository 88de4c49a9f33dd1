"""Regenerate tests/golden/vectors.json: golden vectors of the ORACLE for the seeded cases of tests/golden_cases.py
(sha256 of the bytes of every output array -- the parity bar is bit-exactness -- plus a small block of sample values).

Run from the repository root in the build container:  python tests/golden/make_vectors.py
(The reference Fortran cannot be compiled in this image and ships no golden vectors for these routines; the file pins the
oracle's own arithmetic, see the docstring of tests/golden_cases.py.)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O      # noqa: E402
import golden_cases as GC           # noqa: E402

meta = {}
for name in sorted(GC.CASES):
    gf, extra = GC.inputs(O, name)
    out = GC.run_oracle(O, name, gf, extra)
    meta[name] = {"input_sha256": GC.input_hash(gf, extra), "shape": [GC.KJPT, GC.K, GC.GJ, GC.G], "outputs": {}}
    for k, v in sorted(out.items()):
        assert np.isfinite(v).all(), (name, k)
        meta[name]["outputs"][k] = {"sha256": GC.digest(v), "sum": float(v.sum()), "sample": GC.sample(v)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json")
with open(path, "w") as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print(path, os.path.getsize(path), "bytes,", len(meta), "cases")
