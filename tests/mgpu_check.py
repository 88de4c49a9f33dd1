"""Multi-GPU parity check (one process per GPU, NCCL halo / north-fold exchange).  Launch on an N-GPU box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py
Every rank builds the same global random inputs on the CPU, runs the ORACLE mono-domain once, runs the product on its
own subdomain (jpni x jpnj = best partition for N) and compares its interior with the oracle bit for bit, for
closed / cyclic / T-pivot / F-pivot domains and every kernel schedule."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H          # noqa: E402
import nemo_fct_b200 as N    # noqa: E402
from oracle import oracle as O   # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    part = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}[world]
    G, GJ, K, kjpt = 130, 96, 12, 2
    ok_all = True
    only_fct = os.environ.get("MGPU_ONLY_FCT", "0") == "1"          # tests/test_gpu_multi.py: the tra_adv_fct matrix only
    only_diag = os.environ.get("MGPU_ONLY_DIAG", "0") == "1"        # the collective diagnostics only (glob_sum, stp_ctl)
    for jperio in (() if only_diag else (0, 1, 4, 6)):
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt=kjpt, seed=50 + jperio)
        for (h, v) in ((2, 2), (4, 4)):
            ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v)
            for schedule in (0, 2, 4):
                w = O.World(G, GJ, K, jperio, part[0], part[1])
                loc = {k: w.scatter(gf[k])[rank] for k in H.DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
                od = w.doms[rank]
                dom = N.mpp_init(G, GJ, K, jperio, part[0], part[1], rank + 1)
                ctx = N.FctContext(dom, lr)
                ctx.set_schedule(schedule)
                if world > 1:
                    idt = torch.zeros(N.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
                    if rank == 0:
                        idt.copy_(torch.frombuffer(bytearray(N.comm_unique_id()), dtype=torch.uint8))
                    dist.broadcast(idt, 0)
                    ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)
                ctx.set_domain_arrays(loc["tmask"], loc["umask"], loc["vmask"], loc["wmask"], loc["e1e2t"], loc["r1_e1e2t"],
                                      loc["mikt"], loc["mbkt"], False, False)
                ctx.set_e3t(loc["e3t_b"], loc["e3t_n"], loc["e3t_a"])
                t = {k: torch.from_numpy(loc[k]).to(dev) for k in ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
                ctx.tra_adv_fct(1, 1, "TRA", gf["p2dt"], t["pun"], t["pvn"], t["pwn"], t["ptb"], t["ptn"], t["pta"], kjpt, h, v)
                ctx.synchronize()
                got = t["pta"].cpu().numpy()
                j0, i0 = od.njmpp - 1, od.nimpp - 1
                want = ref[..., j0 + od.nldj - 1:j0 + od.nlej, i0 + od.nldi - 1:i0 + od.nlei]
                mine = got[..., od.nldj - 1:od.nlej, od.nldi - 1:od.nlei]
                ok = bool(np.array_equal(mine, want))
                nex, nby = ctx.comm_report()
                flag = torch.tensor([1 if ok else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if rank == 0:
                    print("jperio=%d h%d/v%d schedule=%d %dx%d: %s (rank0: %d exchanges, %d bytes sent)" %
                          (jperio, h, v, schedule, part[0], part[1], "BIT-IDENTICAL" if flag.item() else "MISMATCH", nex, nby), flush=True)
                ok_all = ok_all and bool(flag.item())
                ctx.close()
                w.close()
    # ---- widened rows over NCCL: tra_adv_mus (both structures) and tra_nxt (Asselin + swap, Euler swap) --------------
    def make_ctx(jperio, loc, schedule):
        dom = N.mpp_init(G, GJ, K, jperio, part[0], part[1], rank + 1)
        ctx = N.FctContext(dom, lr)
        ctx.set_schedule(schedule)
        if world > 1:
            idt = torch.zeros(N.UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(N.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            ctx.comm_init(bytes(idt.cpu().numpy().tobytes()), world, rank)
        ctx.set_domain_arrays(loc["tmask"], loc["umask"], loc["vmask"], loc["wmask"], loc["e1e2t"], loc["r1_e1e2t"],
                              loc["mikt"], loc["mbkt"], False, False)
        ctx.set_e3t(loc["e3t_b"], loc["e3t_n"], loc["e3t_a"])
        return ctx

    def agree(ok, what):
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("%s %dx%d: %s" % (what, part[0], part[1], "BIT-IDENTICAL" if flag.item() else "MISMATCH"), flush=True)
        return bool(flag.item())

    # ---- collective diagnostics over NCCL: glob_sum (MPI_SUMDD) and stp_ctl with ln_ctl (mpp_max / mpp_maxloc) ---------
    from test_gpu_glob_sum import _tmask_i
    for jperio in (() if only_fct else (0, 4, 6)):
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt=kjpt, seed=90 + jperio)
        rng = np.random.default_rng(11)
        big = np.ascontiguousarray(gf["ptb"][0] * (1.0 + 1e8 * rng.standard_normal(gf["ptb"][0].shape)))    # heavy cancellation
        cvol = np.ascontiguousarray(gf["e1e2t"][None] * gf["e3t_n"] * gf["tmask"])
        sshn = np.ascontiguousarray(0.5 * rng.standard_normal((GJ, G)) * gf["tmask"][0])
        un = np.ascontiguousarray(0.3 * rng.standard_normal((K, GJ, G)) * gf["umask"])
        tem = (10.0 + rng.standard_normal((K, GJ, G))) * gf["tmask"]
        sal = (35.0 + rng.standard_normal((K, GJ, G))) * gf["tmask"]
        wet = np.argwhere(gf["tmask"][:, 2:-2, 2:-2] == 1.0)
        kk, jj, ii = wet[len(wet) // 2]
        un[kk, jj + 2, ii + 2] = -12.0                                  # one rank holds it, every rank must report it
        tsn = np.ascontiguousarray(np.stack([tem, sal]))
        w1 = O.World(G, GJ, K, jperio)
        want_sum = O.glob_sum(w1, [np.ascontiguousarray(big * cvol)], [np.ascontiguousarray(gf["tmask_i"])])[0]
        want_ctl = O.stp_ctl(w1.doms[0], sshn, un, tsn, gf["tmask"])
        w1.close()
        w = O.World(G, GJ, K, jperio, part[0], part[1])
        loc = {k: w.scatter(gf[k])[rank] for k in H.DOM_KEYS}
        ctx = make_ctx(jperio, loc, 4)
        ti = _tmask_i(ctx.dom, loc["tmask"])
        td = lambda a: torch.from_numpy(np.ascontiguousarray(w.scatter(a)[rank])).to(dev)   # noqa: E731
        got_sum = ctx.glob_sum("mgpu", [td(big)], torch.from_numpy(ti).to(dev), w3d=td(cvol))[0]
        ok_all = agree(got_sum == want_sum, "glob_sum over NCCL jperio=%d (%r)" % (jperio, got_sum)) and ok_all
        l_ts = torch.from_numpy(np.ascontiguousarray(np.stack([w.scatter(tsn[0])[rank], w.scatter(tsn[1])[rank]]))).to(dev)
        got = ctx.stp_ctl(5, td(sshn), td(un), l_ts, collective=True)
        ok = (got["zmax"] == want_ctl["zmax"] and got["iu"] == want_ctl["iu"] and got["kindic"] == -3 and got["nan_found"] == 0
              and got["is1"] == want_ctl["is1"] and got["is2"] == want_ctl["is2"] and got["ih"] == want_ctl["ih"])
        ok_all = agree(ok, "stp_ctl (ln_ctl) over NCCL jperio=%d |U|max=%g at %s" % (jperio, got["zmax"][1], got["iu"])) and ok_all
        ctx.close()
        w.close()

    for jperio in (() if (only_fct or only_diag) else (0, 1, 4, 6)):
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt=kjpt, seed=70 + jperio)
        mx = H.mus_extra_fields(O, gf, G, GJ, K, jperio, seed=70 + jperio)
        _, refl = H.oracle_mus(O, gf, mx, G, GJ, K, jperio, part[0], part[1], kjpt)
        w = O.World(G, GJ, K, jperio, part[0], part[1])
        loc = {k: w.scatter(gf[k])[rank] for k in H.DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
        lx = {k: w.scatter(mx[k])[rank] for k in H.MUS_KEYS}
        for schedule in (0, 1, 2):
            ctx = make_ctx(jperio, loc, schedule)
            ctx.set_mus_metrics(lx["r1_e1e2u"], lx["r1_e1e2v"])
            ctx.set_e3uvw(lx["e3u_n"], lx["e3v_n"], lx["e3w_n"])
            t = {k: torch.from_numpy(loc[k]).to(dev) for k in ("pun", "pvn", "pwn", "ptb", "pta")}
            ctx.tra_adv_mus(1, 1, "TRC", gf["p2dt"], t["pun"], t["pvn"], t["pwn"], t["ptb"], t["pta"], kjpt)
            ctx.synchronize()
            ok_all = agree(np.array_equal(t["pta"].cpu().numpy(), refl[rank]), "tra_adv_mus jperio=%d schedule=%d" % (jperio, schedule)) and ok_all
            ctx.close()
        for (cdtype, l_euler) in (("TRA", False), ("TRC", True)):
            oloc = {k: w.scatter(gf[k]) for k in H.DOM_KEYS + ("ptb", "ptn", "pta")}
            for r, d in enumerate(w.doms):
                d.set_fields(*[oloc[k][r] for k in H.DOM_KEYS], ln_linssh=False)
            t = {k: torch.from_numpy(oloc[k][rank].copy()).to(dev) for k in ("ptb", "ptn", "pta")}
            w.tra_nxt(3, 1, l_euler, 900.0, cdtype, [O.NxtForcing(atfp=0.1) for _ in w.doms], oloc["ptb"], oloc["ptn"], oloc["pta"], kjpt)
            ctx = make_ctx(jperio, loc, 2)
            ctx.tra_nxt(3, 1, l_euler, 900.0, cdtype, N.NxtForcing(atfp=0.1), t["ptb"], t["ptn"], t["pta"], kjpt)
            ctx.synchronize()
            ok = all(np.array_equal(t[k].cpu().numpy(), oloc[k][rank]) for k in ("ptb", "ptn", "pta"))
            ok_all = agree(ok, "tra_nxt %s euler=%d jperio=%d" % (cdtype, l_euler, jperio)) and ok_all
            ctx.close()
        w.close()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok_all else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
