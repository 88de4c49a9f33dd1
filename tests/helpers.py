"""Shared test helpers: build inputs with the ORACLE's lbc_lnk (CPU), run the oracle, run the product."""
import importlib

import numpy as np
import torch

BF = importlib.import_module("nemo-fmi-devel_b200.bench_fields")

DOM_KEYS = ("tmask", "umask", "vmask", "wmask", "e3t_b", "e3t_n", "e3t_a", "e1e2t", "r1_e1e2t", "mikt", "mbkt")


def to_np(f):
    return {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in f.items()}


def global_bench_fields(O, jpiglo, jpjglo, jpk, jperio, kjpt, rn_rdt=3600.0, **kw):
    """BENCH-style fields on the whole domain, halos set by the oracle's single-domain lbc_lnk. numpy dict."""
    w = O.World(jpiglo, jpjglo, jpk, jperio, 1, 1)

    def lbc(trip):
        w.lbc_lnk([[t.numpy()] for t, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

    f = to_np(BF.bench_fields(w.doms[0], kjpt, lbc, rn_rdt, **kw))
    w.close()
    return f


def random_fields(O, jpiglo, jpjglo, jpk, jperio, kjpt, seed=0, land=0.15, cfl=0.3, ln_linssh=False, ln_isfcav=False):
    """Seeded random global inputs with land patches and partial columns, made lbc-consistent by the oracle's
    single-domain lbc_lnk (so that any decomposition sees valid halos, A.2 of SURVEY.md)."""
    rng = np.random.default_rng(seed)
    w = O.World(jpiglo, jpjglo, jpk, jperio, 1, 1)
    d = w.doms[0]
    jpi, jpj = d.jpi, d.jpj

    def lbc(arrs, nat, sgn):
        w.lbc_lnk([[a] for a in arrs], nat, sgn)

    k_bot = rng.integers(1, jpk, size=(jpj, jpi)).astype(np.float64)          # 1..jpk-1
    k_bot[rng.random((jpj, jpi)) < land] = 0.0
    k_bot[rng.random((jpj, jpi)) < 0.5] = jpk - 1                             # many full columns
    k_top = np.minimum(1.0, k_bot)
    if ln_isfcav:                                                             # some ice-shelf cavities
        cav = (rng.random((jpj, jpi)) < 0.2) & (k_bot >= 4)
        k_top[cav] = rng.integers(2, 4, size=int(cav.sum()))
    kb3, kt3 = k_bot[None].copy(), k_top[None].copy()
    lbc([kb3, kt3], "TT", [1.0, 1.0])
    k_bot_i = kb3[0].astype(np.int32)
    k_top_i = kt3[0].astype(np.int32)
    m = w.dom_msk([np.ascontiguousarray(k_top_i)], [np.ascontiguousarray(k_bot_i)])[0]
    tmask, umask, vmask, wmask = m["tmask"], m["umask"], m["vmask"], m["wmask"]
    e1e2t = (1.0e5 * (1.0 + 0.1 * rng.random((jpj, jpi)))) ** 2
    e1e2t3 = e1e2t[None].copy(); lbc([e1e2t3], "T", [1.0]); e1e2t = np.ascontiguousarray(e1e2t3[0])
    e1e2t[e1e2t == 0.0] = 1.0e10                                              # closed edges: keep metrics non-zero
    r1_e1e2t = 1.0 / e1e2t
    e3 = 10.0 + 50.0 * rng.random((jpk, 1, 1))
    e3t_n = e3 * (1.0 + 0.01 * rng.random((jpk, jpj, jpi)))
    if ln_linssh:
        e3t_b = e3t_n.copy(); e3t_a = e3t_n.copy()
    else:
        e3t_b = e3t_n * (1.0 + 1e-3 * rng.standard_normal((jpk, jpj, jpi)))
        e3t_a = e3t_n * (1.0 + 1e-3 * rng.standard_normal((jpk, jpj, jpi)))
    lbc([e3t_b, e3t_n, e3t_a], "TTT", [1.0, 1.0, 1.0])
    for a in (e3t_b, e3t_n, e3t_a):
        a[a == 0.0] = 30.0
    p2dt = 2.0 * 3600.0
    ptb = (rng.random((kjpt, jpk, jpj, jpi)) * 10.0 + 5.0) * tmask[None]
    ptb[:, :, :, : jpi // 2] += 20.0 * tmask[None][:, :, :, : jpi // 2]       # a front, so the limiter works
    ptn = ptb * (1.0 + 0.05 * rng.standard_normal((kjpt, jpk, jpj, jpi)))
    pta = 1e-6 * rng.standard_normal((kjpt, jpk, jpj, jpi)) * tmask[None]
    lbc([ptb.reshape(-1, jpj, jpi), ptn.reshape(-1, jpj, jpi), pta.reshape(-1, jpj, jpi)], "TTT", [1.0, 1.0, 1.0])
    umax = cfl * 1.0e5 / p2dt
    pun = rng.uniform(-umax, umax, (jpk, jpj, jpi)) * 1.0e5 * e3t_n * umask
    pvn = rng.uniform(-umax, umax, (jpk, jpj, jpi)) * 1.0e5 * e3t_n * vmask
    pwn = rng.uniform(-1.0, 1.0, (jpk, jpj, jpi)) * cfl * e1e2t[None] * e3t_n / p2dt * wmask
    pun[-1] = 0.0; pvn[-1] = 0.0; pwn[-1] = 0.0
    lbc([pun, pvn, pwn], "UVW", [-1.0, -1.0, 1.0])
    w.close()
    return dict(tmask=tmask, umask=umask, vmask=vmask, wmask=wmask, e1e2t=e1e2t, r1_e1e2t=r1_e1e2t,
                mikt=m["mikt"], mbkt=m["mbkt"], tmask_i=m["tmask_i"], e3t_b=e3t_b, e3t_n=e3t_n, e3t_a=e3t_a,
                pun=pun, pvn=pvn, pwn=pwn, ptb=ptb, ptn=ptn, pta=pta, p2dt=p2dt)


def oracle_fct(O, gf, jpiglo, jpjglo, jpk, jperio, jpni, jpnj, kjpt, h, v, ln_linssh=False, ln_isfcav=False,
               ln_nnogather=True, key_mpp_mpi=True, capture=None, poison=True):
    """Run the oracle's tra_adv_fct on global fields ``gf`` decomposed jpni x jpnj; returns the global pta
    (assembled from the local interiors) and the per-rank local pta list."""
    w = O.World(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, ln_nnogather=ln_nnogather, key_mpp_mpi=key_mpp_mpi)
    loc = {k: w.scatter(gf[k]) for k in DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
    for r, d in enumerate(w.doms):
        d.set_fields(*[loc[k][r] for k in DOM_KEYS], ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        if capture is not None:
            cap = {n: np.full(d.shape3, np.nan) for n in capture["names"]}
            capture.setdefault("out", []).append(cap)
            d.set_dbg(capture["jn"], **cap)
    O.lib().oracle_poison_workspace(int(poison))
    w.tra_adv_fct(gf["p2dt"], loc["pun"], loc["pvn"], loc["pwn"], loc["ptb"], loc["ptn"], loc["pta"], kjpt, h, v)
    O.lib().oracle_poison_workspace(0)
    glob = w.gather(loc["pta"], gf["pta"].copy())
    doms = w.doms
    w.close()
    return glob, loc["pta"], doms


def device_fct(N, gf, jpiglo, jpjglo, jpk, jperio, jpni, jpnj, kjpt, h, v, ln_linssh=False, ln_isfcav=False,
               host_path=False, schedule=0, nsteps=1, kernels=None, arith=0):
    """Run the product on cuda:0 through the C ABI: single subdomain (jpni = jpnj = 1) or an in-process group of
    jpni x jpnj subdomains on one GPU.  Returns (global pta, list of local pta).  kernels: list that receives the names of
    the kernels subdomain 0 launched (per-kernel profiling of the library)"""
    from oracle import oracle as O   # only for scatter/gather bookkeeping of the test itself
    w = O.World(jpiglo, jpjglo, jpk, jperio, jpni, jpnj)
    loc = {k: w.scatter(gf[k]) for k in DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
    n = jpni * jpnj
    dev = torch.device("cuda:0")
    if n == 1:
        dom = N.mpp_init(jpiglo, jpjglo, jpk, jperio, 1, 1, 1)
        ctx = N.FctContext(dom, 0)
        ctxs = [ctx]
    else:
        grp = N.LocalGroup(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, 0)
        ctxs = grp.ctx
    ctxs[0].set_schedule(schedule)
    if arith:
        ctxs[0].set_arithmetic(arith)
    if kernels is not None:
        ctxs[0].set_profiling(True)
    for r, c in enumerate(ctxs):
        c.set_domain_arrays(loc["tmask"][r], loc["umask"][r], loc["vmask"][r], loc["wmask"][r], loc["e1e2t"][r],
                            loc["r1_e1e2t"][r], loc["mikt"][r], loc["mbkt"][r], ln_linssh, ln_isfcav)
        c.set_e3t(loc["e3t_b"][r], loc["e3t_n"][r], loc["e3t_a"][r])
    if host_path:
        assert n == 1
        pta = loc["pta"][0].copy()
        ctxs[0].tra_adv_fct(1, 1, "TRA", gf["p2dt"], loc["pun"][0], loc["pvn"][0], loc["pwn"][0], loc["ptb"][0],
                            loc["ptn"][0], pta, kjpt, h, v)
        out = [pta]
    else:
        t = {k: [torch.from_numpy(a).to(dev) for a in loc[k]] for k in ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
        for _ in range(nsteps):
            if n == 1:
                ctxs[0].tra_adv_fct(1, 1, "TRA", gf["p2dt"], t["pun"][0], t["pvn"][0], t["pwn"][0], t["ptb"][0],
                                    t["ptn"][0], t["pta"][0], kjpt, h, v)
                ctxs[0].synchronize()
            else:
                grp.tra_adv_fct(1, 1, "TRA", gf["p2dt"], t["pun"], t["pvn"], t["pwn"], t["ptb"], t["ptn"], t["pta"],
                                kjpt, h, v)
                grp.synchronize()
        out = [a.cpu().numpy() for a in t["pta"]]
    if kernels is not None:
        kernels.extend(ctxs[0].profile().keys())
    glob = w.gather(out, gf["pta"].copy())
    w.close()
    for c in ctxs:
        c.close()
    return glob, out


def max_rel_diff(a, b):
    """max |a-b| / max(|b|, tiny*scale) -- the metric of BASELINE.md par. 5"""
    scale = np.abs(b).max()
    den = np.maximum(np.abs(b), 1e-30 * max(scale, 1e-300))
    return float((np.abs(a - b) / den).max())


# ---- MUSCL (tra_adv_mus) and tra_nxt helpers ---------------------------------------------------------------------------
MUS_KEYS = ("r1_e1e2u", "r1_e1e2v", "e3u_n", "e3v_n", "e3w_n")


def mus_extra_fields(O, gf, jpiglo, jpjglo, jpk, jperio, seed=1, runoff=False):
    """Extra module arrays tra_adv_mus reads (global, lbc-consistent): r1_e1e2u/v, e3u_n/e3v_n/e3w_n, and for
    ld_msc_ups a river-mouth mask rnfmsk (jpj,jpi) + rnfmsk_z (jpk)."""
    rng = np.random.default_rng(1000 + seed)
    w = O.World(jpiglo, jpjglo, jpk, jperio, 1, 1)
    d = w.doms[0]
    jpi, jpj = d.jpi, d.jpj

    def lbc(arrs, nat, sgn):
        w.lbc_lnk([[a] for a in arrs], nat, sgn)

    e1e2u = (gf["e1e2t"] * (1.0 + 0.05 * rng.random((jpj, jpi))))[None].copy()
    e1e2v = (gf["e1e2t"] * (1.0 + 0.05 * rng.random((jpj, jpi))))[None].copy()
    lbc([e1e2u, e1e2v], "UV", [1.0, 1.0])
    e1e2u[e1e2u == 0.0] = 1.0e10
    e1e2v[e1e2v == 0.0] = 1.0e10
    e3u_n = gf["e3t_n"] * (1.0 + 0.01 * rng.standard_normal((jpk, jpj, jpi)))
    e3v_n = gf["e3t_n"] * (1.0 + 0.01 * rng.standard_normal((jpk, jpj, jpi)))
    e3w_n = gf["e3t_n"] * (1.0 + 0.01 * rng.standard_normal((jpk, jpj, jpi)))
    lbc([e3u_n, e3v_n, e3w_n], "UVW", [1.0, 1.0, 1.0])
    for a in (e3u_n, e3v_n, e3w_n):
        a[a == 0.0] = 30.0
    out = dict(r1_e1e2u=np.ascontiguousarray(1.0 / e1e2u[0]), r1_e1e2v=np.ascontiguousarray(1.0 / e1e2v[0]),
               e3u_n=e3u_n, e3v_n=e3v_n, e3w_n=e3w_n)
    if runoff:
        msk = (rng.random((jpj, jpi)) < 0.2).astype(np.float64) * rng.choice([0.5, 1.0], size=(jpj, jpi))
        m3 = msk[None].copy(); lbc([m3], "T", [1.0])
        out["rnfmsk"] = np.ascontiguousarray(m3[0])
        z = np.zeros(jpk); z[: max(2, jpk // 3)] = 0.5; z[0] = 1.0
        out["rnfmsk_z"] = z
    w.close()
    return out


def oracle_mus(O, gf, mx, jpiglo, jpjglo, jpk, jperio, jpni, jpnj, kjpt, ln_linssh=False, ln_isfcav=False,
               ld_msc_ups=False, xind_zero=False, key_mpp_mpi=True):
    """oracle tra_adv_mus on the global fields decomposed jpni x jpnj; returns (global pta, local list)."""
    w = O.World(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, key_mpp_mpi=key_mpp_mpi)
    loc = {k: w.scatter(gf[k]) for k in DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "pta")}
    lx = {k: w.scatter(mx[k]) for k in MUS_KEYS}
    rn = w.scatter(mx["rnfmsk"]) if ld_msc_ups else None
    xind = []
    for r, d in enumerate(w.doms):
        d.set_fields(*[loc[k][r] for k in DOM_KEYS], ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        d.set_mus_fields(*[lx[k][r] for k in MUS_KEYS])
        xi = d.mus_xind(ld_msc_ups, rn[r] if ld_msc_ups else None, mx["rnfmsk_z"] if ld_msc_ups else None)
        if xind_zero:
            xi[:] = 0.0
        xind.append(xi)
    w.tra_adv_mus(gf["p2dt"], loc["pun"], loc["pvn"], loc["pwn"], loc["ptb"], loc["pta"], kjpt, xind)
    glob = w.gather(loc["pta"], gf["pta"].copy())
    w.close()
    return glob, loc["pta"]


def device_mus(N, gf, mx, jpiglo, jpjglo, jpk, jperio, jpni, jpnj, kjpt, ln_linssh=False, ln_isfcav=False,
               ld_msc_ups=False, schedule=0, host_path=False):
    """the product's tra_adv_mus on cuda:0 through the C ABI (single subdomain or in-process group)."""
    from oracle import oracle as O   # scatter/gather bookkeeping of the test itself
    w = O.World(jpiglo, jpjglo, jpk, jperio, jpni, jpnj)
    loc = {k: w.scatter(gf[k]) for k in DOM_KEYS + ("pun", "pvn", "pwn", "ptb", "pta")}
    lx = {k: w.scatter(mx[k]) for k in MUS_KEYS}
    rn = w.scatter(mx["rnfmsk"]) if ld_msc_ups else None
    n = jpni * jpnj
    dev = torch.device("cuda:0")
    if n == 1:
        ctxs = [N.FctContext(N.mpp_init(jpiglo, jpjglo, jpk, jperio, 1, 1, 1), 0)]
    else:
        grp = N.LocalGroup(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, 0)
        ctxs = grp.ctx
    ctxs[0].set_schedule(schedule)
    for r, c in enumerate(ctxs):
        c.set_domain_arrays(loc["tmask"][r], loc["umask"][r], loc["vmask"][r], loc["wmask"][r], loc["e1e2t"][r],
                            loc["r1_e1e2t"][r], loc["mikt"][r], loc["mbkt"][r], ln_linssh, ln_isfcav)
        c.set_e3t(loc["e3t_b"][r], loc["e3t_n"][r], loc["e3t_a"][r])
        c.set_mus_metrics(lx["r1_e1e2u"][r], lx["r1_e1e2v"][r])
        c.set_e3uvw(lx["e3u_n"][r], lx["e3v_n"][r], lx["e3w_n"][r])
        c.set_mus_upstream(ld_msc_ups, rn[r] if ld_msc_ups else None, mx["rnfmsk_z"] if ld_msc_ups else None)
    if host_path:
        assert n == 1
        pta = loc["pta"][0].copy()
        ctxs[0].tra_adv_mus(1, 1, "TRC", gf["p2dt"], loc["pun"][0], loc["pvn"][0], loc["pwn"][0], loc["ptb"][0], pta, kjpt)
        out = [pta]
    else:
        t = {k: [torch.from_numpy(a).to(dev) for a in loc[k]] for k in ("pun", "pvn", "pwn", "ptb", "pta")}
        if n == 1:
            ctxs[0].tra_adv_mus(1, 1, "TRC", gf["p2dt"], t["pun"][0], t["pvn"][0], t["pwn"][0], t["ptb"][0], t["pta"][0], kjpt)
            ctxs[0].synchronize()
        else:
            grp.tra_adv_mus(1, 1, "TRC", gf["p2dt"], t["pun"], t["pvn"], t["pwn"], t["ptb"], t["pta"], kjpt)
            grp.synchronize()
        out = [a.cpu().numpy() for a in t["pta"]]
    glob = w.gather(out, gf["pta"].copy())
    w.close()
    for c in ctxs:
        c.close()
    return glob, out


# ---- tra_adv_cen helpers ---------------------------------------------------------------------------------------------
def oracle_cen(O, gf, jpiglo, jpjglo, jpk, jperio, jpni, jpnj, kjpt, h, v, ln_linssh=False, ln_isfcav=False, poison=False):
    w = O.World(jpiglo, jpjglo, jpk, jperio, jpni, jpnj)
    loc = {k: w.scatter(gf[k]) for k in DOM_KEYS + ("pun", "pvn", "pwn", "ptn", "pta")}
    for r, d in enumerate(w.doms):
        d.set_fields(*[loc[k][r] for k in DOM_KEYS], ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    O.lib().oracle_poison_workspace(int(poison))
    w.tra_adv_cen(loc["pun"], loc["pvn"], loc["pwn"], loc["ptn"], loc["pta"], kjpt, h, v)
    O.lib().oracle_poison_workspace(0)
    glob = w.gather(loc["pta"], gf["pta"].copy())
    w.close()
    return glob, loc["pta"]


def device_cen(N, gf, jpiglo, jpjglo, jpk, jperio, jpni, jpnj, kjpt, h, v, ln_linssh=False, ln_isfcav=False):
    from oracle import oracle as O
    w = O.World(jpiglo, jpjglo, jpk, jperio, jpni, jpnj)
    loc = {k: w.scatter(gf[k]) for k in DOM_KEYS + ("pun", "pvn", "pwn", "ptn", "pta")}
    n = jpni * jpnj
    dev = torch.device("cuda:0")
    if n == 1:
        ctxs = [N.FctContext(N.mpp_init(jpiglo, jpjglo, jpk, jperio, 1, 1, 1), 0)]
    else:
        grp = N.LocalGroup(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, 0)
        ctxs = grp.ctx
    for r, c in enumerate(ctxs):
        c.set_domain_arrays(loc["tmask"][r], loc["umask"][r], loc["vmask"][r], loc["wmask"][r], loc["e1e2t"][r],
                            loc["r1_e1e2t"][r], loc["mikt"][r], loc["mbkt"][r], ln_linssh, ln_isfcav)
        c.set_e3t(loc["e3t_b"][r], loc["e3t_n"][r], loc["e3t_a"][r])
    t = {k: [torch.from_numpy(a).to(dev) for a in loc[k]] for k in ("pun", "pvn", "pwn", "ptn", "pta")}
    if n == 1:
        ctxs[0].tra_adv_cen(1, 1, "TRA", t["pun"][0], t["pvn"][0], t["pwn"][0], t["ptn"][0], t["pta"][0], kjpt, h, v)
        ctxs[0].synchronize()
    else:
        grp.tra_adv_cen(1, 1, "TRA", t["pun"], t["pvn"], t["pwn"], t["ptn"], t["pta"], kjpt, h, v)
        grp.synchronize()
    out = [a.cpu().numpy() for a in t["pta"]]
    glob = w.gather(out, gf["pta"].copy())
    w.close()
    for c in ctxs:
        c.close()
    return glob, out
