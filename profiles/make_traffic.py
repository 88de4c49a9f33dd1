"""Regenerate profiles/ncu_traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and time of EVERY kernel of one
bench.py step, from one ncu pass over the default bench command.  Run on the GPU box (gpurun), in the same call as the bench:

    python profiles/make_traffic.py [--workload orca025] [--out gpurun_out/ncu_traffic.json]

then copy the result to profiles/ncu_traffic.json.  bench.py reads it for roofline.traffic and refuses entries whose kernel set
does not match the kernels it launched (stale file).  The launch list (per-launch time, cold cache, serialised) is written
next to it as <out>.launches.csv."""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = [("k_fct_fused", "fct_fused"), ("k_interp_4th_cpt_tiled", "interp_4th_cpt"), ("k_interp_4th_cpt", "interp_4th_cpt"),
         ("k_fct_low_antidiff_tma", "fct_low_antidiff_inner"), ("k_fct_low_antidiff_inner", "fct_low_antidiff"),
         ("k_fct_low_antidiff", "fct_low_antidiff"), ("k_fct_nonosc_final", "fct_nonosc_final"), ("k_fct_laplacian", "fct_laplacian"),
         ("k_fct_betas", "fct_betas"), ("k_fct_limit", "fct_limit"), ("k_fct_final", "fct_final"), ("k_lbc_pack", "lbc_pack"),
         ("k_lbc_unpack", "lbc_fill_unpack"), ("k_lbc_fill", "lbc_fill_unpack")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="orca025")
    ap.add_argument("--schedule", type=int, default=4)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ncu_traffic.json"))
    a = ap.parse_args()
    log = a.out + ".launches.csv"
    cmd = ["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-c", "2000",
           "--csv", "--log-file", log, sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(a.steps), "--warmup", "3", "--no-cpu-baseline",
           "--no-e2e", "--workload", a.workload, "--schedule", str(a.schedule)]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    rows = [r for r in csv.reader(open(log)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iid = hdr.index("ID")
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        per[(r[iid], r[ik])][r[im]] = float(r[iv].replace(",", ""))
    launches = collections.defaultdict(list)
    for (_, kname), m in per.items():
        for pat, nm in NAMES:
            if pat in kname:
                launches[nm].append(m)
                break
    nsteps = a.steps + 3                                                        # warm-up steps launch the same kernels
    kernels, times, counts = {}, {}, {}
    for nm, ms in launches.items():
        kernels[nm] = sum(m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0) for m in ms) / nsteps
        times[nm] = sum(m.get("gpu__time_duration.sum", 0.0) for m in ms) / nsteps
        counts[nm] = len(ms) / nsteps
    sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip() or os.environ.get("GIT_SHA", "unknown")
    key = "%s:schedule%d" % (a.workload, a.schedule)
    out = {}
    committed = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(committed):
        out = json.load(open(committed))
    out[key] = {"source": "profiles/make_traffic.py: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                          "over `bench.py --steps %d --warmup 3` (B200, N=1); per STEP: bytes and ns summed over the launches of each kernel" % a.steps,
                "git": sha, "kernels": kernels, "time_ns_per_step_cold_serialised": times, "launches_per_step": counts}
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps({key: out[key]}, indent=1))


if __name__ == "__main__":
    main()
