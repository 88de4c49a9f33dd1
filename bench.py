#!/usr/bin/env python
"""bench.py -- FCT tracer-advection step (tra_adv_fct + its lbc_lnk halo / north-fold exchanges) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              bench.py --gpus N --steps K --warmup W

One "step" = one tra_adv_fct call over all kjpt tracers of the workload, halo exchanges included.
Metric (BASELINE.json / BASELINE.md par. 3):  Mpts/s = jpiglo*jpjglo*jpk / t_step / 1e6  (whole job, all GPUs).
  value    : inputs resident in HBM, CUDA events on the launching stream, max over ranks, barrier + sync both sides
  e2e      : the same call through the host-pointer C-ABI entry point (nemo_fct_set_e3t + nemo_tra_adv_fct) with
             pinned HOST buffers: H2D of e3t_b/n/a, pun, pvn, pwn, ptb, ptn, pta and D2H of pta inside the timed region
  roofline : HBM bound; whole-step algorithmic bytes B_alg = N*(32*kjpt+56) over the step time, plus the same
             accounting per kernel from CUDA events recorded around every launch inside the timed region
  cpu_baseline / --impl reference : the C oracle (line-faithful restatement of the reference's CPU tra_adv_fct;
             the Fortran itself cannot be built here: no Fortran compiler) on the host cores, threads as MPI ranks.
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FCT tracer-advection Mpts/s per step"
UNIT = "Mpts/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    """HBM GB/s denominator: driver-measured MEASURED_PEAKS.json, else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (cpu_baseline of our arm, and the whole of --impl reference)
# ----------------------------------------------------------------------------------------------------------------
CPU_PARTITION = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2), 16: (4, 4)}   # tests/BENCH/EXPREF/best_jpni_jpnj_eorca025 (16: 4x4)
CPU_CORES_MAX = 16            # pinned: the boxes of the BENCH and SCALE runs have different core counts
CPU_SAMPLE_PTS = 30.0e6       # points of the j-slab one CPU step works on (~0.5 s on 16 cores)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_shape(cfg):
    """The bounded sample both CPU legs time: same jpiglo, jpk, jperio, tracers and scheme, jpjglo cut to a j-slab"""
    G, GJ, K, jperio, kjpt, h, v, rdt = cfg
    gj = GJ if G * GJ * K <= CPU_SAMPLE_PTS else max(16, int(CPU_SAMPLE_PTS / (G * K)))
    return G, gj, K


def cpu_reference_run(cfg, steps, warmup):
    """Time the oracle's tra_adv_fct (threads standing in for MPI ranks, halo exchange through the restated
    mpp_lnk, ln_nnogather = .TRUE. as the reference's default) on the bounded j-slab.  Returns (Mpts/s, info dict)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    from oracle import oracle as O
    G, GJ, K, jperio, kjpt, h, v, rdt = cfg
    cores = max(c for c in CPU_PARTITION if c <= min(host_cores(), CPU_CORES_MAX))
    jpni, jpnj = CPU_PARTITION[cores]
    _, gj, _ = cpu_sample_shape(cfg)
    gf = H.global_bench_fields(O, G, gj, K, jperio, kjpt, rn_rdt=rdt)
    w = O.World(G, gj, K, jperio, jpni, jpnj, ln_nnogather=True)
    loc = {k: w.scatter(gf[k]) for k in H.DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
    for r, d in enumerate(w.doms):
        d.set_fields(*[loc[k][r] for k in H.DOM_KEYS])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        w.tra_adv_fct(gf["p2dt"], loc["pun"], loc["pvn"], loc["pwn"], loc["ptb"], loc["ptn"], loc["pta"], kjpt, h, v)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    w.close()
    npts = G * gj * K
    ms = 1e3 * sum(times) / len(times)
    info = {"cores": cores, "kind": "port",
            "sample": "j-slab %dx%dx%d of %dx%dx%d (same jpiglo/jpk/jperio=%d, kjpt=%d, FCT h%d/v%d), %dx%d subdomains = %d host threads as MPI "
                      "ranks (pinned to <= %d), C restatement of tra_adv_fct with the reference's per-call work arrays "
                      "(gcc -O3 -ffp-contract=off, as arch-linux_gfortran.fcm: -O3, no fast-math), not the Fortran binary" %
                      (G, gj, K, G, GJ, K, jperio, kjpt, h, v, jpni, jpnj, cores, CPU_CORES_MAX),
            "ms_per_step": ms}
    return npts / (ms * 1e-3) / 1e6, info


def run_reference(args, cfg, rank):
    """--impl reference: the CPU arm on OUR arm's config / metric / unit; each step is the bounded j-slab sample"""
    if rank != 0:
        return
    BF = importlib.import_module("nemo-fmi-devel_b200.bench_fields")
    val, info = cpu_reference_run(cfg, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, cfg, args.gpus, BF.best_partition(args.gpus), default_schedule(args), args.scheme),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def default_schedule(args):
    return 4 if args.schedule is None else args.schedule


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs local to its GPU (sysfs local_cpulist of the PCI device), so that the pinned
    host buffers of the e2e path are first-touched on the GPU's own NUMA node.  Round 1 measured 19 GB/s per GPU at N = 8 with
    all eight ranks on node 0.  Returns a short description for the JSON line; never fails the run."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (dom, bus, getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0))
        with open(path + "local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        node = None
        try:
            with open(path + "numa_node") as f:
                node = int(f.read().strip())
        except Exception:
            pass
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"numa_node": node, "cpus": len(allowed)}
        return {"numa_node": node, "cpus": 0, "note": "local cpulist outside the allowed set: affinity unchanged"}
    except Exception as ex:
        return {"note": "affinity unchanged: %r" % (ex,)}


def workload_config(name, cfg, n_gpus, part, schedule, scheme):
    """identical for both arms (the driver compares them): the GPU arm runs the whole workload, the CPU legs a j-slab of it"""
    G, GJ, K, jperio, kjpt, h, v, rdt = cfg
    sg, sgj, sk = cpu_sample_shape(cfg)
    return {"workload": "%s: tests/BENCH-style %dx%dx%d, jperio=%d, kjpt=%d, FCT h%d/v%d, z-star (ln_linssh=F)" %
                        (name, G, GJ, K, jperio, kjpt, h, v),
            "jpni_x_jpnj": "%dx%d" % part, "n_gpus": n_gpus, "schedule": schedule, "scheme": scheme,
            "cache": "inputs larger than L2 (working set >> 126 MB), no explicit flush",
            "cpu_sample": "the CPU legs (cpu_baseline, --impl reference) time a bounded j-slab %dx%dx%d of this workload on <= %d host "
                          "threads; the GPU arm runs all of it on jpni_x_jpnj GPUs" % (sg, sgj, sk, CPU_CORES_MAX)}


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="orca025")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--schedule", type=int, default=None)
    ap.add_argument("--scheme", default="fct", choices=["fct", "mus", "nxt"],
                    help="fct (default, the BASELINE.json metric); mus / nxt time the widened rows tra_adv_mus / tra_nxt "
                         "on the same workload (device-resident value + roofline only)")
    args = ap.parse_args()
    if args.scheme != "fct":
        args.no_e2e = args.no_cpu_baseline = True

    BF = importlib.import_module("nemo-fmi-devel_b200.bench_fields")
    cfg = BF.CONFIGS[args.workload]
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import nemo_fct_b200 as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the FCT path has no CPU fallback (use --impl reference for the CPU arm)")
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity0 = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)                 # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    G, GJ, K, jperio, kjpt, h, v, rdt = cfg
    part = BF.best_partition(world)
    dom = N.mpp_init(G, GJ, K, jperio, part[0], part[1], rank + 1)
    ctx = N.FctContext(dom, local_rank)
    sched = default_schedule(args)
    ctx.set_schedule(sched)
    if world > 1:
        idt = torch.zeros(N.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(N.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)
    # a dedicated (non-default) torch stream: the library launches on it, torch events time it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)

    # ---- synthetic BENCH fields, generated on the device; halos through the product's own lbc_lnk --------------
    def lbc(trip):
        flat = []
        for t, nat, sgn in trip:
            flat += [t, nat, sgn]
        ctx.lbc_lnk_multi("bench", *flat)

    f = BF.bench_fields(dom, kjpt, lbc, rdt, device=dev)
    torch.cuda.synchronize()
    ctx.set_domain_arrays(*[f[k].cpu().numpy() for k in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")],
                          ln_linssh=False, ln_isfcav=False)
    for k in ("umask", "vmask", "wmask"):
        f[k] = None
    ctx.set_e3t(f["e3t_b"], f["e3t_n"], f["e3t_a"])          # device-resident, borrowed
    pta0 = f["pta"].clone()
    npts_global = G * GJ * K
    npts_local = dom.jpi * dom.jpj * dom.jpk

    if args.scheme == "mus":                                  # uniform BENCH grid: e1e2u = e1e2v = e1e2t, e3u = e3v = e3w = e3t (zco)
        r1 = f["r1_e1e2t"].cpu().numpy()
        ctx.set_mus_metrics(r1, r1)
        ctx.set_e3uvw(f["e3t_n"], f["e3t_n"], f["e3t_n"])
    forcing = N.NxtForcing(atfp=0.1) if args.scheme == "nxt" else None

    def step():
        if args.scheme == "fct":
            ctx.tra_adv_fct(1, 1, "TRA", f["p2dt"], f["pun"], f["pvn"], f["pwn"], f["ptb"], f["ptn"], f["pta"], kjpt, h, v)
        elif args.scheme == "mus":
            ctx.tra_adv_mus(1, 1, "TRC", f["p2dt"], f["pun"], f["pvn"], f["pwn"], f["ptb"], f["pta"], kjpt)
        else:                                                 # leap-frog + Asselin (tra_nxt_vvl) + the two lbc_lnk of tra_nxt
            ctx.tra_nxt(2, 1, False, rdt, "TRA", forcing, f["ptb"], f["ptn"], f["pta"], kjpt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    f["pta"].copy_(pta0)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.set_profiling(True)
    n0 = N.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = N.launch_count() - n0
    ms = e0.elapsed_time(e1) / args.steps
    prof = ctx.profile()
    ctx.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = npts_global / (ms_max * 1e-3) / 1e6
    if not bool(torch.isfinite(f["pta"]).all().item()):
        raise SystemExit("bench.py: non-finite result")
    # decomposition-invariant checksum: wrapping int64 sum of the bit patterns of pta over the GLOBAL interior
    # (2:jpiglo-1, 2:jpjglo-1, all levels and tracers), every global point taken from the one rank that owns it.  The same
    # workload gives the same number at N = 1, 2, 4, 8 and for every schedule (the step is bit-reproducible).
    i0 = max(dom.nldi, 2 - dom.nimpp + 1); i1 = min(dom.nlei, G - 1 - dom.nimpp + 1)
    j0 = max(dom.nldj, 2 - dom.njmpp + 1); j1 = min(dom.nlej, GJ - 1 - dom.njmpp + 1)
    bits = f["pta"][..., j0 - 1:j1, i0 - 1:i1].contiguous().view(torch.int64)
    csum = bits.sum(dtype=torch.int64).reshape(1)
    cnt = torch.tensor([bits.numel()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(csum); dist.all_reduce(cnt)
    checksum = {"int64_sum_of_pta_bits_global_interior": int(csum.item()), "points": int(cnt.item()),
                "expected_points": (G - 2) * (GJ - 2) * K * kjpt, "steps_applied": args.steps}

    # ---- roofline ---------------------------------------------------------------------------------------------------
    peak, peak_src = measured_peak()
    b_alg = BF.algorithmic_bytes(npts_local, kjpt)            # this GPU's share
    scope = "whole step per GPU: B_alg = N_local*(32*kjpt+56) bytes over the step time (all kernels + exchanges)"
    if args.scheme == "mus":      # per tracer-point r ptb, pta, w pta; per point r pun pvn pwn e3u_n e3v_n e3w_n e3t_n tmask
        b_alg = npts_local * (24 * kjpt + 64)
        scope = "whole step per GPU: B_alg = N_local*(24*kjpt+64) bytes (tra_adv_mus) over the step time"
    elif args.scheme == "nxt":    # per tracer-point r ptb ptn pta, w ptb ptn; per point r e3t_b e3t_n e3t_a
        b_alg = npts_local * (40 * kjpt + 24)
        scope = "whole step per GPU: B_alg = N_local*(40*kjpt+24) bytes (tra_nxt_vvl) over the step time"
    n3 = npts_local
    # per-kernel algorithmic bytes: the arrays each kernel must read/write once ("kernels").  In the fused
    # schedules the frame kernels run on thin bands on a side stream, overlapped with the inner kernels: their event
    # spans include waiting for SMs, so only their time is listed.
    fused = sched >= 1
    a3 = n3 * 8                                                                        # one 3-D array
    kbytes = {
        "interp_4th_cpt": a3 * (2 * kjpt) + a3 * 2,                                    # r ptn; w ztw; r wmask, zwt
        "fct_low_antidiff_inner": a3 * ((3 + 5 + (1 if v == 4 else 0)) * kjpt) + a3 * 7,   # r ptb ptn pta [ztw]; w pta zwi zwx zwy zwz; r pun pvn pwn e3t_b/n/a tmask
        "fct_fused": a3 * ((3 + 1 + (1 if v == 4 else 0)) * kjpt) + a3 * 7,            # r ptb ptn pta [ztw]; w pta; r pun pvn pwn e3t_b/n/a tmask
        "fct_nonosc_final": a3 * ((6 + 1) * kjpt) + a3 * 2,                            # r ptb zwi zwx zwy zwz pta; w pta; r tmask e3t_n
        "mus_inner": a3 * (3 * kjpt) + a3 * 11,                                        # r ptb pta; w pta; r pun pvn pwn e3u e3v e3w e3t tmask umask vmask wmask
        "tra_nxt": a3 * (5 * kjpt) + a3 * 3,
        "mus_hflux": a3 * (3 * kjpt) + a3 * 6,                                         # r ptb; w fx fy; r pun pvn e3u e3v umask vmask
        "mus_trend": a3 * (5 * kjpt) + a3 * 5,                                         # r fx fy pta ptb; w pta; r tmask pwn e3w wmask e3t
    }
    if not fused:
        kbytes.update({
            "fct_laplacian": a3 * (1 * kjpt + 2 * kjpt) + a3 * 2,                      # r ptn, umask, vmask; w zltu, zltv
            "fct_low_antidiff": a3 * ((3 + 5 + (2 if h == 4 else 0) + (1 if v == 4 else 0)) * kjpt) + a3 * 8,
            "fct_betas": a3 * ((5 + 2) * kjpt) + a3 * 2,                               # r ptb zwi zwx zwy zwz; w zbetup zbetdo; r tmask e3t_n
            "fct_limit": a3 * ((5 + 3) * kjpt),                                        # r zbetup zbetdo zwx zwy zwz; w zwx zwy zwz
            "fct_final": a3 * ((4 + 1) * kjpt) + a3 * 1,                               # r zwx zwy zwz pta; w pta; r e3t_n
        })
    kern = {}
    for name, (tot, calls) in prof.items():
        avg = tot / max(calls, 1)
        ent = {"ms": round(avg, 5), "launches_per_step": calls / args.steps}
        if fused and name not in kbytes:
            ent["stream"] = "side (boundary frame, overlapped with the inner kernels)"
        if name in kbytes:
            ent["alg_bytes"] = kbytes[name]
            ent["gbs"] = round(kbytes[name] / (avg * 1e-3) / 1e9, 1)
            ent["frac"] = round(ent["gbs"] / peak, 4)
        kern[name] = ent
    dominant = max((k for k in kern if k in kbytes), key=lambda k: kern[k]["ms"] * kern[k]["launches_per_step"], default=None)
    achieved = b_alg / (ms_max * 1e-3) / 1e9
    # measured DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch) from the committed `ncu --set full`
    # capture of this workload / schedule at N = 1 (profiles/ncu_traffic.json), summed over the kernels of one step
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as tf:
            key = "%s:schedule%d" % (args.workload, sched) if args.scheme == "fct" else "%s:%s:schedule%d" % (args.workload, args.scheme, sched)
            ent = json.load(tf).get(key)
        # stale-file guard: the capture must hold exactly the big kernels this run launched
        big = {k for k in kern if k in kbytes}
        if ent and world == 1 and big and big <= set(ent["kernels"]):
            traffic = int(sum(ent["kernels"].values()))
            traffic_src = "profiles/ncu_traffic.json (%s, git %s)" % (key, ent.get("git", "round 1"))
            for name in kern:
                if name in ent["kernels"]:
                    kern[name]["traffic"] = int(ent["kernels"][name])
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "scope": scope,
                "alg_bytes_per_step": b_alg, "dominant_kernel": dominant, "kernels": kern}

    # ---- e2e: host-pointer entry points, pinned host buffers ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        names3 = ["e3t_b", "e3t_n", "e3t_a", "pun", "pvn", "pwn"]
        names4 = ["ptb", "ptn", "pta"]
        host = {}
        for k in names3 + names4:
            host[k] = torch.empty(f[k].shape, dtype=torch.float64, pin_memory=True)
            host[k].copy_(f[k] if k != "pta" else pta0)
        hn = {k: host[k].numpy() for k in host}
        h2d = sum(hn[k].nbytes for k in hn)
        d2h = hn["pta"].nbytes
        n_e2e = max(2, min(args.steps, 5))

        def e2e_step():
            ctx.set_e3t(hn["e3t_b"], hn["e3t_n"], hn["e3t_a"])                 # e3t varies every step under vvl
            ctx.tra_adv_fct(1, 1, "TRA", f["p2dt"], hn["pun"], hn["pvn"], hn["pwn"], hn["ptb"], hn["ptn"], hn["pta"], kjpt, h, v)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        ems = 1e3 * (time.perf_counter() - t0) / n_e2e
        te = torch.tensor([ems], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ems = float(te.item())
        e2e = {"value": round(npts_global / (ems * 1e-3) / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": round(ems, 3), "steps": n_e2e,
               "path": "nemo_fct_set_e3t + nemo_tra_adv_fct (host pointers, pinned), synchronous call; inside it upload | step | download are pipelined over tracer batches on three streams", "host_affinity": numa}
        ctx.set_e3t(f["e3t_b"], f["e3t_n"], f["e3t_a"])

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, affinity0)                                     # the CPU leg uses the cores the process was given
            cval, info = cpu_reference_run(cfg, 3, 1)
            cpu = {"value": round(cval, 3), "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
        except Exception as ex:                                                    # never lose the GPU numbers
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        metric = METRIC if args.scheme == "fct" else {"mus": "MUSCL tracer-advection Mpts/s per step", "nxt": "tra_nxt (Asselin filter + swap) Mpts/s per step"}[args.scheme]
        line = {"metric": metric, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_max, 4), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, cfg, world, part, sched, args.scheme),
                "tracer_mpts_per_s": round(value * kjpt, 2), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks, "checksum": checksum}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
