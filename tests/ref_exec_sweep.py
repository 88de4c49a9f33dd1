"""Randomised sweep of the oracle against the reference's own source text (oracle/f90exec.py, oracle/ref_exec.py) -- a script, not
a test (needs the reference tree; minutes of translated Fortran):

    python tests/ref_exec_sweep.py [--seconds 420]

Part 1: tra_adv_fct on random small domains (every jperio 0-7, 2nd / 4th order, ln_linssh / ln_isfcav, 1-3 tracers, land fraction up
to 0.9, CFL up to 0.9, jpk down to 3).  Part 2: the multi-rank exchange (mpp_lnk + mpp_nfd, gather and no-gather) on random layouts
up to 5 x 3 ranks, every nature, random sign.  Part 3: tra_adv_mus (with and without the upstream indicator) and tra_adv_cen (2nd order
horizontal, 2nd / compact vertical).  Round 2 (seeds 12345 and 999): 6230 FCT cases, 9480 exchanges and 5677 MUSCL / centred cases, 0 mismatches (bit for bit)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H                      # noqa: E402
from oracle import oracle as O           # noqa: E402
from oracle import ref_exec as R         # noqa: E402


def same(a, b):
    return np.array_equal(a.view(np.uint64), b.view(np.uint64))


def sweep_fct(seconds, rng):
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < seconds and bad < 5:
        G, GJ, K, kjpt = int(rng.integers(8, 30)), int(rng.integers(8, 24)), int(rng.integers(3, 10)), int(rng.integers(1, 4))
        jperio, h, v = int(rng.integers(0, 8)), int(rng.choice([2, 4])), int(rng.choice([2, 4]))
        lin, isf = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        land, cfl, seed = float(rng.choice([0.0, 0.15, 0.5, 0.9])), float(rng.choice([0.05, 0.3, 0.9])), int(rng.integers(0, 10 ** 6))
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt=kjpt, seed=seed, land=land, cfl=cfl, ln_linssh=lin, ln_isfcav=isf)
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, kjpt, h, v, ln_linssh=lin, ln_isfcav=isf)
        got = R.tra_adv_fct(gf, G, GJ, K, kjpt, h, v, lin, isf, R.reference_lbc(jperio, G, GJ))
        n += 1
        if not same(got, ref):
            bad += 1
            print("MISMATCH", dict(G=G, GJ=GJ, K=K, kjpt=kjpt, jperio=jperio, h=h, v=v, lin=lin, isf=isf, land=land, cfl=cfl, seed=seed), flush=True)
    print("tra_adv_fct: %d cases, %d mismatches, %.0f s" % (n, bad, time.time() - t0))
    return bad


def sweep_exchange(seconds, rng):
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < seconds and bad < 5:
        G, GJ, K = int(rng.integers(12, 40)), int(rng.integers(12, 30)), int(rng.integers(1, 3))
        jperio, ni, nj, nog = int(rng.integers(0, 8)), int(rng.integers(1, 6)), int(rng.integers(1, 4)), bool(rng.integers(0, 2))
        try:
            w = O.World(G, GJ, K, jperio, ni, nj, ln_nnogather=nog)
        except ValueError:
            continue                                   # a layout the reference refuses
        mw = R.MppWorld(w.doms, nog)
        glob = rng.standard_normal((K, GJ, G))
        for nat in "TUVWF":
            sgn = float(rng.choice([1.0, -1.0]))
            a = w.scatter(glob)
            b = [x.copy() for x in a]
            w.lbc_lnk([a], nat, [sgn])
            mw.lbc_lnk(b, nat, sgn)
            n += 1
            if not all(same(x, y) for x, y in zip(a, b)):
                bad += 1
                print("MISMATCH", G, GJ, K, jperio, ni, nj, nog, nat, sgn, flush=True)
        w.close()
    print("mpp_lnk: %d exchanges, %d mismatches, %.0f s" % (n, bad, time.time() - t0))
    return bad


def sweep_mus_cen(seconds, rng):
    t0, n, bad = time.time(), 0, 0
    while time.time() - t0 < seconds and bad < 5:
        G, GJ, K, kjpt = int(rng.integers(8, 26)), int(rng.integers(8, 22)), int(rng.integers(3, 9)), int(rng.integers(1, 4))
        jperio, lin, isf, ups = int(rng.integers(0, 8)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        land, seed = float(rng.choice([0.0, 0.15, 0.6])), int(rng.integers(0, 10 ** 6))
        gf = H.random_fields(O, G, GJ, K, jperio, kjpt=kjpt, seed=seed, land=land, ln_linssh=lin, ln_isfcav=isf)
        mx = H.mus_extra_fields(O, gf, G, GJ, K, jperio, seed=seed, runoff=True)
        lbc = R.reference_lbc(jperio, G, GJ)
        ref, _ = H.oracle_mus(O, gf, mx, G, GJ, K, jperio, 1, 1, kjpt, ln_linssh=lin, ln_isfcav=isf, ld_msc_ups=ups)
        n += 1
        if not same(R.tra_adv_mus(gf, mx, G, GJ, K, kjpt, lin, isf, ups, lbc), ref):
            bad += 1
            print("MUS MISMATCH", G, GJ, K, kjpt, jperio, lin, isf, ups, land, seed, flush=True)
        if isf:
            continue                                   # tra_adv_cen is run without cavities here
        v = int(rng.choice([2, 4]))
        ref, _ = H.oracle_cen(O, gf, G, GJ, K, jperio, 1, 1, kjpt, 2, v)
        n += 1
        if not same(R.tra_adv_cen(gf, G, GJ, K, kjpt, 2, v, False, False, lbc), ref):
            bad += 1
            print("CEN MISMATCH", G, GJ, K, kjpt, jperio, v, seed, flush=True)
    print("tra_adv_mus / tra_adv_cen: %d cases, %d mismatches, %.0f s" % (n, bad, time.time() - t0))
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=420.0)
    ap.add_argument("--seed", type=int, default=12345)
    a = ap.parse_args()
    assert R.available(), "the reference tree is not on this machine"
    rng = np.random.default_rng(a.seed)
    sys.exit(1 if sweep_fct(a.seconds * 0.5, rng) + sweep_exchange(a.seconds * 0.3, rng) + sweep_mus_cen(a.seconds * 0.2, rng) else 0)
