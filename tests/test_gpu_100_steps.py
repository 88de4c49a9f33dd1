"""100 model-like steps (BASELINE.json north_star: 'bounded drift over 100 steps, global tracer content conserved to
round-off').  The same host-side time stepping drives the CPU oracle and the CUDA path (host-pointer C-ABI entry points,
`nemo_tra_adv_fct` + `nemo_lbc_lnk_multi`), Euler forward as NEMO's first step (neuler = 0, traadv.F90:95):
tb = tn ;  tn <- tn + p2dt * adv_trend ;  lbc_lnk(tn).
Asserted: (1) drift device-vs-oracle after 100 steps: ZERO (bit-identical; documented bar 1e-12 per step);
(2) with a discretely non-divergent flow the volume integral of each tracer is conserved to round-off (double-double
glob_sum, lib_fortran.F90:300-332) and (3) the FCT limiter keeps every tracer inside its initial range."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _nondivergent_setup(O, G, GJ, K, jperio, kjpt, seed):
    """streamfunction transports: pun = d(psi)/dj, pvn = -d(psi)/di with integer psi (exact differences) => the
    horizontal divergence vanishes exactly, pwn = 0, steady e3t: volume and tracer content are conserved."""
    rng = np.random.default_rng(seed)
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=kjpt, seed=seed, land=0.0, cfl=0.2)
    w = O.World(G, GJ, K, jperio)
    full = np.full((GJ, G), K - 1, np.int32); top = np.ones((GJ, G), np.int32)
    kb = full.astype(np.float64)[None].copy(); w.lbc_lnk([[kb]], "T", [1.0])
    kt = np.minimum(1.0, kb)
    m = w.dom_msk([np.ascontiguousarray(kt[0].astype(np.int32))], [np.ascontiguousarray(kb[0].astype(np.int32))])[0]
    for k in ("tmask", "umask", "vmask", "wmask", "mikt", "mbkt", "tmask_i"):
        gf[k] = m[k]
    gf["e1e2t"] = np.full((GJ, G), 1.0e10); gf["r1_e1e2t"] = 1.0 / gf["e1e2t"]
    e3 = np.full((K, GJ, G), 64.0)
    gf["e3t_b"] = e3.copy(); gf["e3t_n"] = e3.copy(); gf["e3t_a"] = e3.copy()
    jj, ii = np.meshgrid(np.arange(GJ), np.arange(G), indexing="ij")
    psi = np.zeros((K, GJ, G))
    for k in range(K - 1):
        # psi = 0 on the closed south / north rows (0-based 0 and GJ-2, GJ-1): no flow through the land faces
        psi[k] = np.round(40.0 * np.sin(2 * np.pi * (ii / (G - 2))) * np.sin(np.pi * np.minimum(jj, GJ - 2) / (GJ - 2)) * (1.0 + 0.1 * k))
    scale = 0.2 * 1.0e5 * 64.0 * 1.0e5 / gf["p2dt"] / 40.0          # |u| dt / dx <~ 0.2
    scale = 8.0 * 2.0 ** np.floor(np.log2(scale))                   # power of two: differences stay exact; Courant ~ 0.15
    pun = np.zeros_like(psi); pvn = np.zeros_like(psi)
    pun[:, 1:, :] = (psi[:, 1:, :] - psi[:, :-1, :]) * scale
    pvn[:, :, 1:] = -(psi[:, :, 1:] - psi[:, :, :-1]) * scale
    pun *= gf["umask"]; pvn *= gf["vmask"]
    w.lbc_lnk([[pun], [pvn]], "UV", [-1.0, -1.0])
    gf["pun"], gf["pvn"], gf["pwn"] = pun, pvn, np.zeros_like(psi)
    t0 = (10.0 + 5.0 * np.sin(2 * np.pi * 3 * ii / G)[None] * np.ones((kjpt, K, 1, 1)) + (jj > GJ // 2)[None, None] * 8.0) * gf["tmask"][None]
    t0 = np.ascontiguousarray(t0 + 0.0 * rng.random(t0.shape))
    w.lbc_lnk([[t0.reshape(-1, GJ, G)]], "T", [1.0])
    gf["ptb"] = t0.copy(); gf["ptn"] = t0.copy(); gf["pta"] = np.zeros_like(t0)
    w.close()
    return gf


def _run(step_fct, lbc_t, gf, nsteps, kjpt):
    tn = gf["ptn"].copy()
    tb = tn.copy()
    for _ in range(nsteps):
        ta = np.zeros_like(tn)
        step_fct(tb, tn, ta)
        new = np.ascontiguousarray(tn + gf["p2dt"] * ta) * gf["tmask"][None]     # Euler forward: tb = tn
        lbc_t(new)
        tb, tn = new.copy(), new
    return tb, tn


@pytest.mark.parametrize("case", ["nondivergent_cyclic", "bench_fold"])
def test_100_steps_no_drift_conservation_bounds(N, O, case):
    kjpt, K = 2, 8
    if case == "nondivergent_cyclic":
        G, GJ, jperio, h, v = 48, 36, 1, 4, 4
        gf = _nondivergent_setup(O, G, GJ, K, jperio, kjpt, seed=1)
    else:
        G, GJ, jperio, h, v = 50, 38, 4, 4, 4
        gf = H.global_bench_fields(O, G, GJ, K, jperio, kjpt, cfl=0.2)
    nsteps = 100
    # ---- oracle
    w = O.World(G, GJ, K, jperio)
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS])

    def o_step(tb, tn, ta):
        w.tra_adv_fct(gf["p2dt"], [gf["pun"]], [gf["pvn"]], [gf["pwn"]], [tb], [tn], [ta], kjpt, h, v)

    def o_lbc(a):
        w.lbc_lnk([[a.reshape(-1, GJ, G)]], "T", [1.0])

    otb, otn = _run(o_step, o_lbc, gf, nsteps, kjpt)
    # ---- device, host-pointer entry points
    dom = N.mpp_init(G, GJ, K, jperio)
    ctx = N.FctContext(dom, 0)
    ctx.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], gf["mbkt"])
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])

    def d_step(tb, tn, ta):
        ctx.tra_adv_fct(1, 1, "TRA", gf["p2dt"], gf["pun"], gf["pvn"], gf["pwn"], tb, tn, ta, kjpt, h, v)

    def d_lbc(a):
        ctx.lbc_lnk("tranxt", a.reshape(-1, GJ, G), "T", 1.0)

    dtb, dtn = _run(d_step, d_lbc, gf, nsteps, kjpt)
    ctx.close()
    assert np.isfinite(otn).all()
    assert H.max_rel_diff(dtn, otn) <= 1e-12 * nsteps          # the stated drift bound
    assert np.array_equal(dtn, otn) and np.array_equal(dtb, otb), "drift after 100 steps is not zero"
    assert not np.array_equal(otn, gf["ptn"])
    if case == "nondivergent_cyclic":
        vol = gf["e1e2t"][None] * gf["e3t_n"]
        for n in range(kjpt):
            c0 = O.glob_sum(w, [np.ascontiguousarray(vol * gf["ptn"][n])], [np.ascontiguousarray(gf["tmask_i"])])[0]
            c1 = O.glob_sum(w, [np.ascontiguousarray(vol * dtn[n])], [np.ascontiguousarray(gf["tmask_i"])])[0]
            assert abs(c1 - c0) <= 1e-11 * abs(c0), (c0, c1)
            wet = gf["tmask"] == 1.0
            lo, hi = gf["ptn"][n][wet].min(), gf["ptn"][n][wet].max()
            assert dtn[n][wet].min() >= lo - 1e-9 and dtn[n][wet].max() <= hi + 1e-9, (lo, hi, dtn[n][wet].min(), dtn[n][wet].max())
    w.close()
