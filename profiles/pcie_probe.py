"""Host link probe for the end-to-end numbers: pinned host <-> device copy bandwidth of every rank, alone and with all ranks
copying at once.  Run under torchrun on the GPU box (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/pcie_probe.py

Rank 0 prints one JSON line.  The e2e leg of bench.py moves 12.5 GB up and 2.1 GB down per ORCA025 step, whatever N is: what N
GPUs can pull from the host at the same time bounds it from below."""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist


def copy_gbs(dst, src, stream, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        dst.copy_(src, non_blocking=True)
        stream.synchronize()
        e0.record(stream)
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record(stream)
    stream.synchronize()
    return src.numel() * src.element_size() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    nbytes = 512 << 20
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    host2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host2.fill_(2)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev2 = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}

    def barrier():
        if world > 1:
            dist.barrier()

    # alone: one rank at a time
    alone_up, alone_dn = 0.0, 0.0
    for r in range(world):
        barrier()
        if r == rank:
            alone_up = copy_gbs(dev, host, s_up, 4)
            alone_dn = copy_gbs(host2, dev2, s_dn, 4)
    barrier()
    # all ranks at once, one direction
    all_up = copy_gbs(dev, host, s_up, 8)
    barrier()
    all_dn = copy_gbs(host2, dev2, s_dn, 8)
    barrier()
    # all ranks, both directions at once (what the pipelined entry point does)
    t0 = time.perf_counter()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(s_up); e[2].record(s_dn)
    with torch.cuda.stream(s_up):
        for _ in range(8):
            dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s_dn):
        for _ in range(8):
            host2.copy_(dev2, non_blocking=True)
    e[1].record(s_up); e[3].record(s_dn)
    torch.cuda.synchronize()
    duplex_up = nbytes * 8 / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9
    duplex_dn = nbytes * 8 / (e[2].elapsed_time(e[3]) * 1e-3) / 1e9
    mine = [alone_up, alone_dn, all_up, all_dn, duplex_up, duplex_dn]
    if world > 1:
        out = [None] * world
        dist.all_gather_object(out, mine)
    else:
        out = [mine]
    if rank == 0:
        keys = ["alone_h2d", "alone_d2h", "all_h2d", "all_d2h", "duplex_h2d", "duplex_d2h"]
        for i, k in enumerate(keys):
            res[k + "_gbs_per_rank"] = [round(o[i], 1) for o in out]
            res[k + "_gbs_sum"] = round(sum(o[i] for o in out), 1)
        try:
            res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        except Exception as ex:  # noqa: BLE001
            res["topo"] = "unavailable: %s" % ex
        res["n_gpus"] = world
        res["cpus"] = len(os.sched_getaffinity(0))
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
