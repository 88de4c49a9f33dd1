"""Small seeded cases whose ORACLE outputs are committed as golden vectors (tests/golden/vectors.json: sha256 of every
output array + a block of sample values).

The reference holds no golden vectors for this path and cannot be compiled here; these vectors are generated from the oracle
(tests/golden/make_vectors.py, script committed) AND are what the reference's own source text produces when executed by
translation (tests/golden/ref_exec_pins.json, tests/test_cpu_reference_exec.py).  The CPU suite checks that the oracle still
reproduces them bit for bit, the GPU suite compares the CUDA path with the committed arrays directly, with no oracle in the loop.
Every case records a sha256 of its inputs so that a change of the input generator (numpy RNG stream) is told apart from a
change of the arithmetic."""
import hashlib

import numpy as np

import helpers as H

G, GJ, K, KJPT = 48, 40, 7, 2          # a size at which every fused device schedule engages (even jpi: TMA tiles)

#: name -> (routine, jperio, options)
CASES = {
    "fct_h2v2_jperio0": ("fct", 0, dict(h=2, v=2)),
    "fct_h4v4_jperio4": ("fct", 4, dict(h=4, v=4)),
    "fct_h4v2_jperio6_linssh_isfcav": ("fct", 6, dict(h=4, v=2, ln_linssh=True, ln_isfcav=True)),
    "mus_jperio4": ("mus", 4, dict()),
    "mus_jperio1_upstream_linssh": ("mus", 1, dict(ld_msc_ups=True, ln_linssh=True)),
    "cen_h2v4_jperio6": ("cen", 6, dict(h=2, v=4)),
    "cen_h4v2_jperio0": ("cen", 0, dict(h=4, v=2)),
    "nxt_vvl_jperio4": ("nxt", 4, dict(ln_linssh=False)),
    "nxt_fix_jperio1": ("nxt", 1, dict(ln_linssh=True)),
}
NXT = dict(atfp=0.1, rdt=900.0, r1_rau0=1.0 / 1026.0)


def seed_of(name):
    return 1000 + sorted(CASES).index(name)


def inputs(O, name):
    """(gf, extra) for a case: the seeded global fields and the routine-specific extras."""
    kind, jperio, opt = CASES[name]
    seed = seed_of(name)
    gf = H.random_fields(O, G, GJ, K, jperio, KJPT, seed=seed, ln_linssh=opt.get("ln_linssh", False),
                         ln_isfcav=opt.get("ln_isfcav", False))
    extra = {}
    if kind == "mus":
        extra = H.mus_extra_fields(O, gf, G, GJ, K, jperio, seed=seed, runoff=True)
    if kind == "nxt":
        rng = np.random.default_rng(seed + 7)
        w = O.World(G, GJ, K, jperio, 1, 1)
        f2 = {k: rng.standard_normal((1, GJ, G)) * 1e-4 for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf")}
        w.lbc_lnk([[a] for a in f2.values()], "T" * 6, [1.0] * 6)
        sbc = rng.standard_normal((KJPT, GJ, G)) * 1e-5
        sbc_b = rng.standard_normal((KJPT, GJ, G)) * 1e-5
        w.lbc_lnk([[sbc], [sbc_b]], "TT", [1.0, 1.0])
        w.close()
        extra = {k: np.ascontiguousarray(v[0]) for k, v in f2.items()}
        extra["sbc"], extra["sbc_b"] = sbc, sbc_b
    return gf, extra


def digest(a):
    """sha256 of the array's bytes: bit-exactness is the bar, so a hash is a complete golden vector"""
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample(a):
    """a small block of values kept next to the hash, to see HOW a mismatch looks"""
    return np.ascontiguousarray(a[..., 1:3, 5:9, 7:11]).ravel().tolist()


def input_hash(gf, extra):
    h = hashlib.sha256()
    for d in (gf, extra):
        for k in sorted(d):
            v = d[k]
            if isinstance(v, np.ndarray):
                h.update(k.encode()); h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()


def run_oracle(O, name, gf, extra):
    """dict of output arrays of the oracle for this case (mono-domain)"""
    kind, jperio, opt = CASES[name]
    if kind == "fct":
        out, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, KJPT, opt["h"], opt["v"], ln_linssh=opt.get("ln_linssh", False),
                                 ln_isfcav=opt.get("ln_isfcav", False))
        return {"pta": out}
    if kind == "mus":
        out, _ = H.oracle_mus(O, gf, extra, G, GJ, K, jperio, 1, 1, KJPT, ln_linssh=opt.get("ln_linssh", False),
                              ld_msc_ups=opt.get("ld_msc_ups", False))
        return {"pta": out}
    if kind == "cen":
        out, _ = H.oracle_cen(O, gf, G, GJ, K, jperio, 1, 1, KJPT, opt["h"], opt["v"])
        return {"pta": out}
    w = O.World(G, GJ, K, jperio, 1, 1)
    d = w.doms[0]
    d.set_fields(*[gf[k] for k in H.DOM_KEYS], ln_linssh=opt["ln_linssh"])
    tb, tn, ta = gf["ptb"].copy(), gf["ptn"].copy(), gf["pta"].copy()
    forc = O.NxtForcing(atfp=NXT["atfp"], r1_rau0=NXT["r1_rau0"],
                        **{k: extra[k] for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf")})
    w.tra_nxt(5, 1, False, NXT["rdt"], "TRA", [forc], [tb], [tn], [ta], KJPT, [extra["sbc"]], [extra["sbc_b"]])
    w.close()
    return {"ptb": tb, "ptn": tn, "pta": ta}


def run_device(N, name, gf, extra, schedule):
    """the same case through the C ABI on cuda:0"""
    import torch
    kind, jperio, opt = CASES[name]
    lin, isf = opt.get("ln_linssh", False), opt.get("ln_isfcav", False)
    if kind == "fct":
        out, _ = H.device_fct(N, gf, G, GJ, K, jperio, 1, 1, KJPT, opt["h"], opt["v"], ln_linssh=lin, ln_isfcav=isf, schedule=schedule)
        return {"pta": out}
    if kind == "mus":
        out, _ = H.device_mus(N, gf, extra, G, GJ, K, jperio, 1, 1, KJPT, ln_linssh=lin, ld_msc_ups=opt.get("ld_msc_ups", False),
                              schedule=schedule)
        return {"pta": out}
    if kind == "cen":
        out, _ = H.device_cen(N, gf, G, GJ, K, jperio, 1, 1, KJPT, opt["h"], opt["v"])
        return {"pta": out}
    ctx = N.FctContext(N.mpp_init(G, GJ, K, jperio, 1, 1, 1), 0)
    ctx.set_domain_arrays(*[gf[k] for k in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")], lin, isf)
    ctx.set_e3t(gf["e3t_b"], gf["e3t_n"], gf["e3t_a"])
    t = {k: torch.from_numpy(gf[k].copy()).cuda() for k in ("ptb", "ptn", "pta")}
    forc = N.NxtForcing(atfp=NXT["atfp"], r1_rau0=NXT["r1_rau0"],
                        **{k: torch.from_numpy(extra[k]).cuda() for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf")})
    ctx.tra_nxt(5, 1, False, NXT["rdt"], "TRA", forc, t["ptb"], t["ptn"], t["pta"], KJPT,
                torch.from_numpy(extra["sbc"]).cuda(), torch.from_numpy(extra["sbc_b"]).cuda())
    ctx.synchronize()
    out = {k: v.cpu().numpy() for k, v in t.items()}
    ctx.close()
    return out
