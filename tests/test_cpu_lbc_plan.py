"""The product's compiled lbc_lnk gather plan (host C++, layout.cpp) executed in numpy must reproduce the oracle's
mpp_lnk / lbc_nfd / mpp_nfd restatement bit for bit -- every cell of every rank, halos and corners included --
for all jperio, grid-point types, fold signs and processor grids (mpp_lnk_generic.h90:85-336,
lbc_nfd_generic.h90:66-163, mpp_nfd_generic.h90:222-298)."""
import numpy as np
import pytest


def run_plan(N, doms, fields, nat, sgn, pval=0.0):
    out = [f.copy() for f in fields]
    for r, d in enumerate(doms):
        flat = out[r].reshape(out[r].shape[0], -1)
        dst, _, sp = N.lbc_plan(d, nat, -1)
        flat[:, dst] = pval * (sgn ** sp)
        for p in range(len(doms)):
            dst, src, sp = N.lbc_plan(d, nat, p)
            if len(dst):
                flat[:, dst] = fields[p].reshape(fields[p].shape[0], -1)[:, src] * (sgn ** sp)
    return out


LAYOUTS = [(1, 1), (2, 1), (1, 2), (2, 2), (3, 2), (4, 2), (3, 1)]


@pytest.mark.parametrize("jperio", range(8))
@pytest.mark.parametrize("size", [(20, 17), (21, 16), (12, 11)])
def test_plan_equals_oracle_gather_path_on_arbitrary_fields(N, O, jperio, size):
    """ln_nnogather = .FALSE. (MPI_ALLGATHER path): identical for ANY input, odd jpiglo included."""
    G, GJ = size
    rng = np.random.default_rng(jperio * 100 + G)
    ncase = 0
    for ni, nj in LAYOUTS:
        try:
            w = O.World(G, GJ, 3, jperio, ni, nj, ln_nnogather=False)
        except ValueError:
            continue
        try:
            doms = [N.mpp_init(G, GJ, 3, jperio, ni, nj, r + 1) for r in range(ni * nj)]
        except N.NemoFctError:
            w.close(); continue
        for nat, sgn in [("T", 1.0), ("U", -1.0), ("V", -1.0), ("W", 1.0), ("F", -1.0), ("T", -1.0), ("F", 1.0), ("U", 1.0)]:
            fields = [rng.standard_normal((2, d.jpj, d.jpi)) for d in doms]
            ref = [f.copy() for f in fields]
            w.lbc_lnk([ref], nat, [sgn])
            got = run_plan(N, doms, fields, nat, sgn)
            for r in range(len(doms)):
                assert np.array_equal(ref[r], got[r]), (size, jperio, (ni, nj), nat, sgn, r)
            ncase += 1
        w.close()
    assert ncase > 0


@pytest.mark.parametrize("jperio", [3, 4, 5, 6])
def test_nogather_path_agrees_on_fold_consistent_fields(N, O, jperio):
    """ln_nnogather = .TRUE. (BENCH default) differs from the gather path only on fold-INconsistent data
    (chap_misc.tex:314-316 requires both to agree); on fields that already satisfy the fold (as every model field
    does after its own lbc_lnk) the plan, the no-gather path and the gather path coincide.  T/U/V/W only: for
    F-points the reference's no-gather path additionally writes (1|nlci, nlcj-1), which lbc_nfd never does."""
    G, GJ = 24, 19
    rng = np.random.default_rng(jperio)
    w1 = O.World(G, GJ, 3, jperio, 1, 1)
    for ni, nj in [(2, 1), (2, 2), (3, 2), (4, 2)]:
        doms = [N.mpp_init(G, GJ, 3, jperio, ni, nj, r + 1) for r in range(ni * nj)]
        for nat, sgn in [("T", 1.0), ("U", -1.0), ("V", -1.0), ("W", 1.0)]:
            glob = rng.standard_normal((2, GJ, G))
            w1.lbc_lnk([[glob]], nat, [sgn])                       # make it fold-consistent
            res = []
            for nogather in (True, False):
                w = O.World(G, GJ, 3, jperio, ni, nj, ln_nnogather=nogather)
                loc = w.scatter(glob)
                for a in loc:                                       # spoil the halos: the exchange must rebuild them
                    a[:, 0, :] = 9.0; a[:, -1, :] = 9.0; a[:, :, 0] = 9.0; a[:, :, -1] = 9.0
                spoiled = [a.copy() for a in loc]
                w.lbc_lnk([loc], nat, [sgn])
                res.append(loc)
                w.close()
            got = run_plan(N, doms, spoiled, nat, sgn)
            for r in range(len(doms)):
                assert np.array_equal(res[0][r], res[1][r]), ("oracle nogather vs gather", jperio, ni, nj, nat, r)
                assert np.array_equal(res[1][r], got[r]), ("plan", jperio, ni, nj, nat, r)
    w1.close()


def test_pval_land_value_and_f_points(N, O):
    """closed boundaries take pval (lbc_lnk_generic.h90:76-78); F-points keep their west/south edge (:89,:96)"""
    G, GJ = 14, 12
    rng = np.random.default_rng(1)
    w = O.World(G, GJ, 2, 0, 1, 1)
    d = N.mpp_init(G, GJ, 2, 0, 1, 1, 1)
    f = rng.standard_normal((2, GJ, G))
    got = run_plan(N, [d], [f], "F", 1.0, pval=0.0)[0]
    assert np.array_equal(got[:, :-1, 0], f[:, :-1, 0]) and np.all(got[:, :, -1] == 0.0) and np.all(got[:, -1, :] == 0.0)
    dst, _, _ = N.lbc_plan(d, "T", -1)
    assert len(dst) == 2 * G + 2 * GJ - 4                       # the whole rim is land for T-points
    w.close()
