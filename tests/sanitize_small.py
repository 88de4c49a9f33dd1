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
    for schedule in (0, 1, 2, 3, 4):
        got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 1, 1, 2, h, v, schedule=schedule)
        same = bool(np.array_equal(got, ref))
        print("jperio", jperio, "h/v", h, v, "schedule", schedule, "bit-identical" if same else "MISMATCH", flush=True)
        ok = ok and same
    for schedule in (2, 4):                      # 4 in a group: persistent k_fct_fused beside the frame chain
        got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 2, 2, 2, h, v, schedule=schedule)
        print("jperio", jperio, "2x2 in-process group schedule", schedule, bool(np.array_equal(got, ref)), flush=True)
        ok = ok and bool(np.array_equal(got, ref))
# widened rows: MUSCL (all three structures, in-process group), centred scheme, tra_nxt
import golden_cases as GC    # noqa: E402
for jperio in (4, 0):
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=8)
    mx = H.mus_extra_fields(O, gf, G, GJ, K, jperio, seed=8, runoff=True)
    ref, _ = H.oracle_mus(O, gf, mx, G, GJ, K, jperio, 1, 1, 2)
    for schedule in (0, 1, 2):
        _, loc = H.device_mus(N, gf, mx, G, GJ, K, jperio, 1, 1, 2, schedule=schedule)
        same = bool(np.array_equal(loc[0], ref))
        print("tra_adv_mus jperio", jperio, "schedule", schedule, "bit-identical" if same else "MISMATCH", flush=True)
        ok = ok and same
    got, _ = H.device_mus(N, gf, mx, G, GJ, K, jperio, 2, 2, 2, schedule=2)
    inner = (slice(None), slice(0, K - 1), slice(1, -1), slice(1, -1))
    same = bool(np.array_equal(got[inner], ref[inner]))
    print("tra_adv_mus jperio", jperio, "2x2 in-process group", same, flush=True)
    ok = ok and same
    for (h, v) in ((2, 2), (4, 4)):
        ref, _ = H.oracle_cen(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
        _, loc = H.device_cen(N, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
        same = bool(np.array_equal(loc[0], ref))
        print("tra_adv_cen jperio", jperio, "h/v", h, v, "bit-identical" if same else "MISMATCH", flush=True)
        ok = ok and same
for name in ("nxt_vvl_jperio4", "nxt_fix_jperio1"):
    gf, extra = GC.inputs(O, name)
    ref = GC.run_oracle(O, name, gf, extra)
    got = GC.run_device(N, name, gf, extra, 2)
    same = all(bool(np.array_equal(got[k], ref[k])) for k in ref)
    print("tra_nxt", name, "bit-identical" if same else "MISMATCH", flush=True)
    ok = ok and same
# diagnostic reductions: glob_sum (double-double) and stp_ctl (extrema + locations): warp shuffles + shared memory
import torch                 # noqa: E402
gf = H.random_fields(O, G, GJ, K, 4, kjpt=2, seed=9)
rng = np.random.default_rng(9)
sshn = np.ascontiguousarray(rng.standard_normal((GJ, G)))
un = np.ascontiguousarray(rng.standard_normal((K, GJ, G)))
tsn = np.ascontiguousarray(np.stack([10.0 + rng.standard_normal((K, GJ, G)), 35.0 + rng.standard_normal((K, GJ, G))]))
w = O.World(G, GJ, K, 4)
ref_ctl = O.stp_ctl(w.doms[0], sshn, un, tsn, gf["tmask"])
ref_sum = O.glob_sum(w, [np.ascontiguousarray(gf["ptb"][0] * gf["e3t_n"])], [np.ascontiguousarray(gf["tmask_i"])])[0]
w.close()
ctx = N.FctContext(N.mpp_init(G, GJ, K, 4, 1, 1, 1), 0)
ctx.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], gf["mbkt"])
dev = torch.device("cuda:0")
got = ctx.stp_ctl(1, torch.from_numpy(sshn).to(dev), torch.from_numpy(un).to(dev), torch.from_numpy(tsn).to(dev))
same = all(got[k] == ref_ctl[k] for k in ("zmax", "ih", "iu", "is1", "is2", "nan_found", "kindic"))
print("stp_ctl", "bit-identical" if same else "MISMATCH", flush=True)
ok = ok and same
gsum = ctx.glob_sum("san", [torch.from_numpy(np.ascontiguousarray(gf["ptb"][0])).to(dev)], torch.from_numpy(np.ascontiguousarray(gf["tmask_i"])).to(dev),
                    w3d=torch.from_numpy(np.ascontiguousarray(gf["e3t_n"])).to(dev))[0]
print("glob_sum", "bit-identical" if gsum == ref_sum else "MISMATCH", flush=True)
ok = ok and gsum == ref_sum
ctx.close()
sys.exit(0 if ok else 1)
