"""N > 1 host logic without a GPU: two processes (gloo, world_size 2, 127.0.0.1) execute the product's compiled
lbc_lnk plan -- the exact per-peer cell lists the CUDA pack / ncclSend / ncclRecv / unpack path uses -- on numpy
fields and must reproduce the oracle's mpp_lnk (threads as MPI ranks) bit for bit on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _exchange(N, doms, rank, field, nat, sgn, nlev, pval=0.0):
    """what nemo_lbc_lnk_multi does on the device, with torch.distributed p2p in place of NCCL"""
    world = len(doms)
    me = doms[rank]
    flat = field.reshape(nlev, -1)
    reqs, recvbufs = [], {}
    for p in range(world):
        if p == rank:
            continue
        _, src_on_me, _ = N.lbc_plan(doms[p], nat, rank)            # pack: cells peer p wants from me, message order
        if len(src_on_me):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(flat[:, src_on_me])), p))
        dst, _, sp = N.lbc_plan(me, nat, p)
        if len(dst):
            recvbufs[p] = (torch.empty((nlev, len(dst)), dtype=torch.float64), dst, sp)
            reqs.append(dist.irecv(recvbufs[p][0], p))
    dst, src, sp = N.lbc_plan(me, nat, rank)                        # cells whose source is on this rank
    local = flat[:, src].copy() if len(dst) else None
    for r in reqs:
        r.wait()
    fdst, _, fsp = N.lbc_plan(me, nat, -1)
    flat[:, fdst] = pval * (sgn ** fsp)
    if local is not None:
        flat[:, dst] = local * (sgn ** sp)
    for p, (buf, d, s) in recvbufs.items():
        flat[:, d] = buf.numpy() * (sgn ** s)


def _worker(rank, world, port, cases, out):
    sys.path.insert(0, ROOT)
    import nemo_fct_b200 as N
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ok = True
    for (G, GJ, jperio, ni, nj) in cases:
        doms = [N.mpp_init(G, GJ, 4, jperio, ni, nj, r + 1) for r in range(world)]
        w = O.World(G, GJ, 4, jperio, ni, nj, ln_nnogather=False)
        for nat, sgn in [("T", 1.0), ("U", -1.0), ("V", -1.0), ("W", 1.0)]:
            rng = np.random.default_rng(G + 7 * jperio + ord(nat))      # same stream on both ranks
            fields = [rng.standard_normal((4, d.jpj, d.jpi)) for d in doms]
            ref = [f.copy() for f in fields]
            w.lbc_lnk([ref], nat, [sgn])                                # the oracle runs all ranks as threads
            mine = fields[rank].copy()
            _exchange(N, doms, rank, mine, nat, sgn, 4)
            ok = ok and np.array_equal(mine, ref[rank])
        w.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def test_two_rank_exchange_over_gloo_matches_oracle():
    cases = [(22, 17, 0, 2, 1), (22, 17, 1, 2, 1), (22, 17, 4, 2, 1), (22, 17, 6, 2, 1), (22, 17, 4, 1, 2), (22, 17, 7, 1, 2)]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cases, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
