"""Committed golden vectors (tests/golden/vectors.json, made by tests/golden/make_vectors.py from the oracle):
CPU: the oracle still reproduces them bit for bit; GPU: the CUDA path reproduces them, no oracle in the loop."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
import golden_cases as GC

META = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


def _case(name):
    gf, extra = GC.inputs(O, name)
    assert GC.input_hash(gf, extra) == META[name]["input_sha256"], \
        "the seeded input generator no longer produces the inputs the golden vectors were made from (numpy RNG stream?)"
    return gf, extra


def _check(name, out):
    gold = META[name]["outputs"]
    assert sorted(out) == sorted(gold)
    for k, v in out.items():
        if GC.digest(v) != gold[k]["sha256"]:
            diff = np.abs(np.asarray(GC.sample(v)) - np.asarray(gold[k]["sample"])).max()
            raise AssertionError("%s/%s differs from the golden vector (sum %r vs %r, max |diff| on the sample block %r)"
                                 % (name, k, float(v.sum()), gold[k]["sum"], diff))


def test_golden_covers_every_case():
    assert sorted(META) == sorted(GC.CASES)


@pytest.mark.parametrize("name", sorted(GC.CASES))
def test_oracle_reproduces_golden(name):
    gf, extra = _case(name)
    _check(name, GC.run_oracle(O, name, gf, extra))


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [0, 2])
@pytest.mark.parametrize("name", sorted(GC.CASES))
def test_device_reproduces_golden(name, schedule):
    """whole arrays: the result on the interior, halos and level jpk returned as received (exchanged halos for tra_nxt)"""
    N = importlib.import_module("nemo-fmi-devel_b200")
    gf, extra = _case(name)
    _check(name, GC.run_device(N, name, gf, extra, schedule))


@pytest.mark.parametrize("name", [n for n in sorted(GC.CASES) if GC.CASES[n][0] in ("fct", "mus")])
def test_emulated_default_schedule_reproduces_golden(name):
    """the default device schedule (fused, TMA tiles / differences in place) run on the CPU from the product's own kernel
    sources and column sets (tests/emu) against the committed vectors: what the GPU case above will see, minus the GPU"""
    import emu_api
    L = emu_api.load()
    kind, jperio, opt = GC.CASES[name]
    gf, extra = _case(name)
    lin, isf = opt.get("ln_linssh", False), opt.get("ln_isfcav", False)
    fold = jperio in (3, 4, 5, 6)
    w = O.World(GC.G, GC.GJ, GC.K, jperio, 1, 1)

    def lbc(trip):
        w.lbc_lnk([[a.reshape(-1, GC.GJ, GC.G)] for a, _, _ in trip], "".join(n for _, n, _ in trip), [s for _, _, s in trip])

    if kind == "fct":
        pta, plan = emu_api.fct_step_fused(L, gf, GC.KJPT, opt["h"], opt["v"], lin, isf, 1, lbc, fold, True, tma=True)
        assert plan["used_tma"]
    else:
        xind = None
        if opt.get("ld_msc_ups"):
            w.doms[0].set_fields(*[gf[k] for k in ("tmask", "umask", "vmask", "wmask", "e3t_b", "e3t_n", "e3t_a", "e1e2t", "r1_e1e2t",
                                                   "mikt", "mbkt")])
            xind = w.doms[0].mus_xind(True, extra["rnfmsk"], extra["rnfmsk_z"])
        plan = emu_api.mus_plan(L, 0, GC.G, GC.GJ, fold)
        pta = gf["pta"].copy()
        zwx, zwy, fx, fy = (np.zeros(pta.shape) for _ in range(4))
        for region, kern in (("inner", 2), ("grad", 0)):
            for rc in plan[region]:
                emu_api.mus(L, kern, rc, 1, gf, extra, xind, pta, zwx, zwy, fx, fy, lin, isf, GC.KJPT)
        lbc([(zwx, "U", -1.0), (zwy, "V", -1.0)])
        for rc in plan["hflux"]:
            emu_api.mus(L, 1, rc, 1, gf, extra, xind, pta, zwx, zwy, fx, fy, lin, isf, GC.KJPT)
        lbc([(fx, "U", -1.0), (fy, "V", -1.0)])
        for rc in plan["trend"]:
            emu_api.mus(L, 3, rc, 1, gf, extra, xind, pta, zwx, zwy, fx, fy, lin, isf, GC.KJPT)
    w.close()
    _check(name, {"pta": pta})
