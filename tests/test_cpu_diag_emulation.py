"""The diagnostic reductions of the product (csrc/glob_sum.cu, csrc/stp_ctl.cu) compiled for the host (tests/emu: one host thread
per CUDA thread, warp shuffles and shared memory emulated) against the oracle: glob_sum's double-double sum (lib_fortran_generic.h90:
32-65) and stp_ctl's extrema with first-occurrence locations (stpctl.F90:115-124, 162-165) -- without a GPU."""
import numpy as np
import pytest

from oracle import oracle as O
import helpers as H
import emu_api


@pytest.fixture(scope="module")
def emu():
    return emu_api.load()


@pytest.mark.parametrize("jperio", [0, 4])
def test_glob_sum_kernels_equal_the_oracle_whatever_the_grid(emu, jperio):
    G, GJ, K = 37, 29, 7
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=300 + jperio)
    rng = np.random.default_rng(2)
    cvol = np.ascontiguousarray(gf["e1e2t"][None] * gf["e3t_n"] * gf["tmask"])
    fields = [np.ascontiguousarray(gf["ptb"][jn] * (1.0 + 1e8 * rng.standard_normal(gf["ptb"][jn].shape))) for jn in range(2)]
    ti = np.ascontiguousarray(gf["tmask_i"])
    w = O.World(G, GJ, K, jperio)
    want = [O.glob_sum(w, [np.ascontiguousarray(f * cvol)], [ti])[0] for f in fields]
    w.close()
    for nblk in (1, 3, 8):
        got = emu_api.glob_sum(emu, fields, ti, w3d=cvol, nblk=nblk)
        assert [g[0] for g in got] == want, nblk
    plain = [float(np.sum(f * cvol * ti[None])) for f in fields]
    assert plain != want                                   # the plain fp64 sum does not get there: the pair matters
    got2d = emu_api.glob_sum(emu, [np.ascontiguousarray(gf["e1e2t"])], ti, nblk=2)     # a 2-D field, no weight
    w = O.World(G, GJ, K, jperio)
    assert got2d[0][0] == O.glob_sum(w, [np.ascontiguousarray(gf["e1e2t"])], [ti])[0]
    w.close()


@pytest.mark.parametrize("case", ["clean", "ties", "nan", "all_land"])
def test_stp_ctl_kernels_equal_the_oracle(emu, case):
    G, GJ, K, jperio = 35, 27, 6, 4
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=321)
    rng = np.random.default_rng(4)
    tmask = np.ascontiguousarray(gf["tmask"] if case != "all_land" else np.zeros_like(gf["tmask"]))
    sshn = np.ascontiguousarray(rng.standard_normal((GJ, G)))
    un = np.ascontiguousarray(rng.standard_normal((K, GJ, G)))
    tem = np.ascontiguousarray(10.0 + rng.standard_normal((K, GJ, G)))
    sal = np.ascontiguousarray(35.0 + rng.standard_normal((K, GJ, G)))
    wet = np.argwhere(gf["tmask"] == 1.0)
    a, b = wet[len(wet) // 4], wet[3 * len(wet) // 4]
    if case == "ties":
        sal[tuple(a)] = sal[tuple(b)] = 44.0; sal[tuple(wet[3])] = sal[tuple(wet[-3])] = 1.0
        un[tuple(a)] = 9.0; un[tuple(b)] = -9.0
        sshn[a[1], a[2]] = -5.0; sshn[b[1], b[2]] = 5.0
    if case == "nan":
        un[tuple(a)] = np.nan; sal[tuple(b)] = np.nan; tem[tuple(a)] = np.nan
    w = O.World(G, GJ, K, jperio)
    ref = O.stp_ctl(w.doms[0], sshn, un, np.ascontiguousarray(np.stack([tem, sal])), tmask)
    w.close()
    for nblk in (1, 5):
        vals, idx, flags = emu_api.stp_ctl(emu, sshn, un, tem, sal, tmask, nblk=nblk)
        none = -np.finfo(np.float64).max
        zmax = [vals[0] if idx[0] >= 0 else none, vals[1] if idx[1] >= 0 else none, -vals[2] if idx[2] >= 0 else none,
                vals[3] if idx[3] >= 0 else none, -vals[4] if idx[4] >= 0 else none, vals[5] if idx[5] >= 0 else none]
        assert zmax == ref["zmax"], (case, nblk)

        def loc(l, nd):
            if l < 0:
                return [0, 0, 0][:nd]
            return [int(l % G) + 1, int((l // G) % GJ) + 1, int(l // (G * GJ)) + 1][:nd]
        assert loc(idx[0], 2) == ref["ih"] and loc(idx[1], 3) == ref["iu"] and loc(idx[2], 3) == ref["is1"] and loc(idx[3], 3) == ref["is2"]
        assert (flags & 1) == ref["nan_found"] == int(case == "nan")
        assert bool(flags & 2) == (case != "all_land")
