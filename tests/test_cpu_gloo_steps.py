"""N > 1 on the CPU, whole steps: two processes (gloo, world_size 2, 127.0.0.1), each owning one subdomain, run a complete
tra_adv_fct step (4th order + compact vertical, exchanges X1..X4), a complete tra_adv_mus step and a complete tra_nxt with
  * the product's column kernels compiled for the host from the same source (tests/emu), and
  * the product's compiled lbc_lnk plans executed over torch.distributed p2p (what pack / ncclSend / ncclRecv / unpack do),
in the order run_fct / run_mus / run_nxt launch them, and must reproduce the oracle's rank (threads as MPI ranks) bit for bit,
including the halos and the north-fold rows."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
K, KJPT = 8, 2


def _worker(rank, world, port, cases, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import nemo_fct_b200 as N
    from oracle import oracle as O
    import helpers as H
    import emu_api
    from test_cpu_gloo_exchange import _exchange
    L = emu_api.load()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ok = True
    for (G, GJ, jperio, ni, nj) in cases:
        doms = [N.mpp_init(G, GJ, K, jperio, ni, nj, r + 1) for r in range(world)]
        gf = H.random_fields(O, G, GJ, K, jperio, KJPT, seed=500 + jperio)           # same stream on both ranks
        mx = H.mus_extra_fields(O, gf, G, GJ, K, jperio, seed=500 + jperio)
        w = O.World(G, GJ, K, jperio, ni, nj)
        me = w.doms[rank]
        loc = {k: w.scatter(gf[k])[rank] for k in H.DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
        lmx = {k: w.scatter(mx[k])[rank] for k in H.MUS_KEYS}
        loc["p2dt"] = gf["p2dt"]
        jpi, jpj = me.jpi, me.jpj
        nlev = K * KJPT

        def lbc(fields_nat_sgn):
            for a, nat, sgn in fields_nat_sgn:
                _exchange(N, doms, rank, a, nat, sgn, nlev)

        # ---- tra_adv_fct, reference pass structure with X1..X4 (run_fct, schedule 0): 4th order + compact vertical ----
        refg, refl_fct, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, ni, nj, KJPT, 4, 4)
        got, _ = emu_api.fct_step(L, loc, KJPT, 4, 4, False, False, 2, lbc)
        ok = ok and bool(np.array_equal(got, refl_fct[rank]))
        if jpi >= 20 and jpj >= 20:
            # the fused schedule as run_fct launches it with real neighbours: band split, TMA tiles where jpi is even
            got, plan = emu_api.fct_step_fused(L, loc, KJPT, 4, 4, False, False, 1, lbc, me.npolj != 0, True, want_split=True, tma=True)
            ok = ok and bool(np.array_equal(got, refl_fct[rank]))
        # ---- tra_adv_mus, reference structure (run_mus, schedule 0) ----
        _, refl = H.oracle_mus(O, gf, mx, G, GJ, K, jperio, ni, nj, KJPT)
        pta = loc["pta"].copy()
        shp = pta.shape
        zwx, zwy, fx, fy = (np.zeros(shp) for _ in range(4))
        emu_api.mus(L, 0, (1, jpi - 1, 1, jpj - 1), 2, loc, lmx, None, pta, zwx, zwy, fx, fy, False, False, KJPT)
        lbc([(zwx, "U", -1.0), (zwy, "V", -1.0)])
        emu_api.mus(L, 1, (2, jpi - 1, 2, jpj - 1), 2, loc, lmx, None, pta, zwx, zwy, fx, fy, False, False, KJPT)
        lbc([(fx, "U", -1.0), (fy, "V", -1.0)])
        emu_api.mus(L, 3, (2, jpi - 1, 2, jpj - 1), 2, loc, lmx, None, pta, zwx, zwy, fx, fy, False, False, KJPT)
        ok = ok and bool(np.array_equal(pta, refl[rank]))
        # ---- tra_nxt: lbc on pta, tra_nxt_vvl, lbc on ptb, ptn, pta (run_nxt) ----
        oloc = {k: w.scatter(gf[k]) for k in H.DOM_KEYS + ("ptb", "ptn", "pta")}
        for r, d in enumerate(w.doms):
            d.set_fields(*[oloc[k][r] for k in H.DOM_KEYS], ln_linssh=False)
        atfp, rdt, r1_rau0 = 0.1, 900.0, 1.0 / 1026.0
        tb, tn, ta = (oloc[k][rank].copy() for k in ("ptb", "ptn", "pta"))
        w.tra_nxt(3, 1, False, rdt, "TRA", [O.NxtForcing(atfp=atfp, r1_rau0=r1_rau0) for _ in w.doms],
                  oloc["ptb"], oloc["ptn"], oloc["pta"], KJPT)
        lbc([(ta, "T", 1.0)])
        iflags = (C.c_int * 5)(0, 0, 0, 0, 0)
        tab = (C.c_void_p * 6)(None, None, None, None, None, None)
        rc = L.emu_nxt(1, 0, jpi, jpj, K, KJPT, C.c_double(atfp), C.c_double(atfp * rdt), C.c_double(atfp * rdt * r1_rau0), iflags,
                       emu_api.p(loc["e3t_b"]), emu_api.p(loc["e3t_n"]), emu_api.p(loc["e3t_a"]), emu_api.p(loc["mikt"]),
                       emu_api.p(tb), emu_api.p(tn), emu_api.p(ta), None, None, tab, *([None] * 12))
        ok = ok and rc == 0
        lbc([(tb, "T", 1.0), (tn, "T", 1.0), (ta, "T", 1.0)])
        for got, k in ((tb, "ptb"), (tn, "ptn"), (ta, "pta")):
            ok = ok and bool(np.array_equal(got, oloc[k][rank]))
        w.close()
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def test_two_rank_fct_muscl_and_tra_nxt_steps_over_gloo():
    import emu_api
    from test_cpu_gloo_exchange import _free_port
    emu_api.load()                                              # build libemu.so once, before the workers start
    cases = [(30, 24, 0, 2, 1), (30, 24, 4, 2, 1), (30, 24, 6, 1, 2), (30, 24, 1, 2, 1), (44, 24, 4, 2, 1), (26, 42, 6, 1, 2)]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cases, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert out.get(timeout=5) == 1
