"""stp_ctl's extrema / NaN guard on the device (csrc/stp_ctl.cu, nemo_stp_ctl_dev) against the oracle's restatement of
src/OCE/stpctl.F90:115-124 (zmax), :149-156 (the condition), :162-165 (MAXLOC / MINLOC), :184 (kindic)."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

G, GJ, K = 61, 47, 9


def _state(O, jperio, seed, case):
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=seed)
    rng = np.random.default_rng(seed)
    tmask = gf["tmask"]
    sshn = 0.5 * rng.standard_normal((GJ, G)) * tmask[0]
    un = 0.3 * rng.standard_normal((K, GJ, G)) * gf["umask"]
    tem = (10.0 + 5.0 * rng.standard_normal((K, GJ, G))) * tmask
    sal = (35.0 + 0.5 * rng.standard_normal((K, GJ, G))) * tmask
    wet = np.argwhere(tmask == 1.0)
    k, j, i = wet[len(wet) // 3]
    if case == "ties":                                     # MAXLOC / MINLOC take the first occurrence in array order
        k2, j2, i2 = wet[2 * len(wet) // 3]
        sal[k, j, i] = sal[k2, j2, i2] = 41.0
        sal[wet[5][0], wet[5][1], wet[5][2]] = sal[wet[-5][0], wet[-5][1], wet[-5][2]] = 3.0
        un[k, j, i] = -7.0; un[k2, j2, i2] = 7.0
        sshn[j, i] = 4.0; sshn[j2, i2] = -4.0
    elif case == "ssh":
        sshn[j, i] = -23.0
    elif case == "u":
        un[k, j, i] = -11.5
    elif case == "s_neg":
        sal[k, j, i] = -0.25
    elif case == "s_big":
        sal[k, j, i] = 100.0
    elif case == "nan_u":
        un[k, j, i] = np.nan
    elif case == "nan_s":
        sal[k, j, i] = np.nan
    elif case == "nan_on_land":                            # a NaN salinity where tmask == 0 is not looked at
        dry = np.argwhere(tmask == 0.0)
        sal[dry[0][0], dry[0][1], dry[0][2]] = np.nan
    tsn = np.ascontiguousarray(np.stack([tem, sal]))
    return gf, np.ascontiguousarray(sshn), np.ascontiguousarray(un), tsn


def _ctx(N, gf, dom):
    ctx = N.FctContext(dom, 0)
    ctx.set_domain_arrays(gf["tmask"], gf["umask"], gf["vmask"], gf["wmask"], gf["e1e2t"], gf["r1_e1e2t"], gf["mikt"], gf["mbkt"])
    return ctx


@pytest.mark.parametrize("case", ["clean", "ties", "ssh", "u", "s_neg", "s_big", "nan_u", "nan_s", "nan_on_land"])
def test_stp_ctl_equals_the_oracle(N, O, case):
    jperio = 4
    gf, sshn, un, tsn = _state(O, jperio, 910, case)
    w = O.World(G, GJ, K, jperio)
    ref = O.stp_ctl(w.doms[0], sshn, un, tsn, gf["tmask"])
    w.close()
    dev = torch.device("cuda:0")
    ctx = _ctx(N, gf, N.mpp_init(G, GJ, K, jperio, 1, 1, 1))
    n0 = N.launch_count()
    got = ctx.stp_ctl(7, torch.from_numpy(sshn).to(dev), torch.from_numpy(un).to(dev), torch.from_numpy(tsn).to(dev))
    assert N.launch_count() == n0 + 2
    ctx.close()
    for key in ("zmax", "ih", "iu", "is1", "is2", "nan_found", "kindic"):
        assert got[key] == ref[key], (case, key, got[key], ref[key])
    assert (got["kindic"] == -3) == (case not in ("clean", "ties", "nan_on_land"))
    if got["kindic"]:
        assert "stp_ctl" in got["message"] and "kt=       7" in got["message"]
    # an independent numpy reading of the same lines
    wet = gf["tmask"] == 1.0
    if case in ("clean", "ties"):
        assert got["zmax"][0] == np.abs(sshn).max() and got["zmax"][1] == np.abs(un).max()
        assert got["zmax"][2] == -tsn[1][wet].min() and got["zmax"][3] == tsn[1][wet].max()
        assert got["zmax"][4] == -tsn[0][wet].min() and got["zmax"][5] == tsn[0][wet].max()
        kk, jj, ii = np.unravel_index(np.argmax(np.abs(un)), un.shape)
        assert got["iu"] == [ii + 1, jj + 1, kk + 1]
        s = np.where(wet, tsn[1], np.inf)
        kk, jj, ii = np.unravel_index(np.argmin(s), s.shape)
        assert got["is1"] == [ii + 1, jj + 1, kk + 1]


def test_stp_ctl_on_a_subdomain_reports_global_indices(N, O):
    jperio, jpni, jpnj = 4, 2, 2
    gf, sshn, un, tsn = _state(O, jperio, 911, "u")
    w = O.World(G, GJ, K, jperio, jpni, jpnj)
    dev = torch.device("cuda:0")
    loc = {n: w.scatter(gf[n]) for n in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")}
    l_ssh, l_un, l_tem, l_sal = w.scatter(sshn), w.scatter(un), w.scatter(tsn[0]), w.scatter(tsn[1])
    for r in range(jpni * jpnj):
        l_ts = np.ascontiguousarray(np.stack([l_tem[r], l_sal[r]]))
        ref = O.stp_ctl(w.doms[r], l_ssh[r], l_un[r], l_ts, loc["tmask"][r])
        ctx = N.FctContext(N.mpp_init(G, GJ, K, jperio, jpni, jpnj, r + 1), 0)
        ctx.set_domain_arrays(*[loc[n][r] for n in ("tmask", "umask", "vmask", "wmask", "e1e2t", "r1_e1e2t", "mikt", "mbkt")])
        got = ctx.stp_ctl(3, torch.from_numpy(l_ssh[r]).to(dev), torch.from_numpy(l_un[r]).to(dev), torch.from_numpy(l_ts).to(dev))
        ctx.close()
        for key in ("zmax", "ih", "iu", "is1", "is2", "nan_found", "kindic"):
            assert got[key] == ref[key], (r, key, got[key], ref[key])
        if got["kindic"]:                                   # the rank(s) holding the 11.5 m/s point name the same global cell
            kk, jj, ii = np.unravel_index(np.argmax(np.abs(un)), un.shape)
            assert got["iu"] == [ii + 1, jj + 1, kk + 1]
    w.close()


def test_stp_ctl_refuses_bad_calls(N, O):
    gf, sshn, un, tsn = _state(O, 0, 912, "clean")
    dev = torch.device("cuda:0")
    ctx = N.FctContext(N.mpp_init(G, GJ, K, 0, 1, 1, 1), 0)
    with pytest.raises(N.NemoFctError, match="set_domain_arrays"):
        ctx.stp_ctl(1, torch.from_numpy(sshn).to(dev), torch.from_numpy(un).to(dev), torch.from_numpy(tsn).to(dev))
    ctx.close()
