"""Regenerate tests/golden/bench_partitions.json from the reference tree (run in the build container only).

The reference ships, for its BENCH configuration, the best processor grids and the resulting local domain sizes
(tests/BENCH/EXPREF/best_jpni_jpnj_eorca{1,025,12}; one line per core count:
 ``nb_cores N ( jpni x jpnj ), nb_points P ( jpimax x jpjmax )``).  They are the only golden data the reference
holds on this path: they pin mpp_basic_decomposition (src/OCE/LBC/mppini.F90:723-724) of the oracle and of the
product.  Global sizes come from namelist_cfg_orca*_like (nn_isize/nn_jsize).
"""
import json
import os
import re

REF = "/root/reference/tests/BENCH/EXPREF"
GLOBAL = {"eorca1": (362, 332, 6), "eorca025": (1442, 1207, 4), "eorca12": (4322, 3147, 4)}   # jpiglo, jpjglo, jperio
KEEP = 160   # lines per table

out = {}
for name, (gi, gj, jperio) in GLOBAL.items():
    rows = []
    with open(os.path.join(REF, "best_jpni_jpnj_" + name)) as f:
        for line in f:
            m = re.search(r"nb_cores\s+(\d+)\s+\(\s*(\d+)\s+x\s+(\d+)\s*\),\s+nb_points\s+(\d+)\s+\(\s*(\d+)\s+x\s+(\d+)\s*\)", line)
            if m:
                rows.append([int(x) for x in m.groups()])
            if len(rows) >= KEEP:
                break
    out[name] = {"jpiglo": gi, "jpjglo": gj, "jperio": jperio, "columns": ["nb_cores", "jpni", "jpnj", "nb_points", "jpimax", "jpjmax"], "rows": rows}
with open(os.path.join(os.path.dirname(__file__), "bench_partitions.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print({k: len(v["rows"]) for k, v in out.items()})
