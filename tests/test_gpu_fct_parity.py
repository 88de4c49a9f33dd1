"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): max relative difference <= 1e-12 after one step.  The product is compiled without
FMA contraction and keeps the reference's operation order, so the expectation asserted here is stronger:
BIT-IDENTICAL interior pta (np.array_equal).  The 1e-12 tolerance is kept as the documented bar.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-12

SMALL = dict(jpiglo=30, jpjglo=22, jpk=11)


@pytest.mark.parametrize("jperio", [0, 1, 4, 6, 7, 2, 3, 5])
@pytest.mark.parametrize("hv", [(2, 2), (4, 2), (2, 4), (4, 4)])
def test_random_fields_single_domain_bit_exact(N, O, jperio, hv):
    h, v = hv
    g = SMALL
    gf = H.random_fields(O, g["jpiglo"], g["jpjglo"], g["jpk"], jperio, kjpt=3, seed=10 * jperio + h + v)
    ref, _, _ = H.oracle_fct(O, gf, g["jpiglo"], g["jpjglo"], g["jpk"], jperio, 1, 1, 3, h, v)
    got, _ = H.device_fct(N, gf, g["jpiglo"], g["jpjglo"], g["jpk"], jperio, 1, 1, 3, h, v)
    assert not np.isnan(ref).any()
    assert H.max_rel_diff(got, ref) <= TOL
    assert np.array_equal(got, ref), "device pta differs from the oracle bitwise"
    assert not np.array_equal(ref, gf["pta"]), "the step did nothing"


@pytest.mark.parametrize("flags", [(True, False), (True, True), (False, True)])
def test_linssh_isfcav_variants(N, O, flags):
    ln_linssh, ln_isfcav = flags
    g = SMALL
    gf = H.random_fields(O, g["jpiglo"], g["jpjglo"], g["jpk"], 4, kjpt=2, seed=77, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    for h, v in [(2, 2), (4, 4)]:
        if v == 4 and ln_isfcav:
            continue   # forbidden by the reference: 4th-order compact with ice-shelf cavities (traadv.F90:249-252)
        ref, _, _ = H.oracle_fct(O, gf, g["jpiglo"], g["jpjglo"], g["jpk"], 4, 1, 1, 2, h, v, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        got, _ = H.device_fct(N, gf, g["jpiglo"], g["jpjglo"], g["jpk"], 4, 1, 1, 2, h, v, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        assert np.array_equal(got, ref)


def test_bench128_closed_T_S(N, O):
    """BASELINE config C1 (tests/BENCH 128x128x75, T+S, FCT 2/2, closed) at full size, BENCH fields."""
    gf = H.global_bench_fields(O, 128, 128, 75, 0, 2)
    ref, _, _ = H.oracle_fct(O, gf, 128, 128, 75, 0, 1, 1, 2, 2, 2)
    got, _ = H.device_fct(N, gf, 128, 128, 75, 0, 1, 1, 2, 2, 2)
    assert H.max_rel_diff(got, ref) <= TOL
    assert np.array_equal(got, ref)


def test_host_pointer_entry_point(N, O):
    """nemo_tra_adv_fct with HOST buffers (what the Fortran shim calls) == device-resident path == oracle.  The call is
    pipelined over batches of tracers (upload | step | download on three streams): 2, 5 and 7 tracers, pageable and
    page-locked (nemo_fct_host_register) host arrays, the one-kernel schedule on a grid large enough for it"""
    g = SMALL
    gf = H.random_fields(O, g["jpiglo"], g["jpjglo"], g["jpk"], 6, kjpt=2, seed=5)
    ref, _, _ = H.oracle_fct(O, gf, g["jpiglo"], g["jpjglo"], g["jpk"], 6, 1, 1, 2, 4, 4)
    got, _ = H.device_fct(N, gf, g["jpiglo"], g["jpjglo"], g["jpk"], 6, 1, 1, 2, 4, 4, host_path=True)
    assert np.array_equal(got, ref)
    # (96 x 130 and 76 x 214: K2 has >= 96 rows, so every batch is further pipelined over 2 / 4 row slabs of the one-kernel step)
    for kjpt, G, GJ, K in ((5, 64, 50, 9), (7, 30, 22, 11), (3, 96, 130, 9), (2, 76, 214, 7)):
        gf = H.random_fields(O, G, GJ, K, 4, kjpt=kjpt, seed=50 + kjpt)
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 4, 1, 1, kjpt, 4, 4)
        got, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, kjpt, 4, 4, host_path=True, schedule=4)
        assert np.array_equal(got, ref), kjpt
        for k in ("pun", "pvn", "pwn", "ptb", "ptn"):
            N.host_register(gf[k])
        got, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, kjpt, 4, 4, host_path=True, schedule=4)
        for k in ("pun", "pvn", "pwn", "ptb", "ptn"):
            N.host_unregister(gf[k])
        assert np.array_equal(got, ref), kjpt
    gf = H.random_fields(O, 90, 150, 8, 0, kjpt=2, seed=61)                          # closed domain, 2nd order, slabs
    ref, _, _ = H.oracle_fct(O, gf, 90, 150, 8, 0, 1, 1, 2, 2, 2)
    got, _ = H.device_fct(N, gf, 90, 150, 8, 0, 1, 1, 2, 2, 2, host_path=True, schedule=4)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("layout", [(2, 1), (1, 2), (2, 2), (3, 2), (4, 2)])
@pytest.mark.parametrize("jperio", [0, 1, 4, 6])
def test_decomposition_invariance_in_process_group(N, O, layout, jperio):
    """jpni x jpnj subdomains on one GPU (in-process communicator) give the mono-domain result bit for bit --
    the property SETTE checks for the reference (stpctl.F90:134,192)."""
    jpni, jpnj = layout
    G, GJ, K = 34, 27, 9
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=3)
    mono, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, 4, 4)
    got, _ = H.device_fct(N, gf, G, GJ, K, jperio, jpni, jpnj, 2, 4, 4)
    assert np.array_equal(got, mono)


# ---- schedule 1: fused inner region + boundary frame on a side stream ------------------------------------------
@pytest.mark.parametrize("jperio", [0, 1, 4, 6, 7, 3])
@pytest.mark.parametrize("hv", [(2, 2), (4, 2), (2, 4), (4, 4)])
def test_schedule1_single_domain_bit_exact(N, O, jperio, hv):
    h, v = hv
    G, GJ, K = 75, 44, 11
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=100 + 10 * jperio + h + v)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
    got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 1, 1, 2, h, v, schedule=1)
    bad = np.argwhere(got != ref)
    assert np.array_equal(got, ref), "first mismatches (jn,k,j,i): %s of %d" % (bad[:6].tolist(), len(bad))


@pytest.mark.parametrize("flags", [(True, False), (True, True), (False, True)])
def test_schedule1_linssh_isfcav(N, O, flags):
    ln_linssh, ln_isfcav = flags
    G, GJ, K = 48, 40, 9
    gf = H.random_fields(O, G, GJ, K, 4, kjpt=2, seed=78, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
    for h, v in [(2, 2), (4, 2)]:
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 4, 1, 1, 2, h, v, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        got, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, 2, h, v, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav, schedule=1)
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("layout", [(2, 1), (2, 2), (3, 2)])
@pytest.mark.parametrize("jperio", [0, 4, 6])
def test_schedule1_decomposition_invariance(N, O, layout, jperio):
    jpni, jpnj = layout
    G, GJ, K = 80, 58, 7
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=4)
    mono, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, 4, 4)
    got, _ = H.device_fct(N, gf, G, GJ, K, jperio, jpni, jpnj, 2, 4, 4, schedule=1)
    assert np.array_equal(got, mono)


def test_schedule1_repeated_steps_and_non_product_masks(N, O):
    """(a) three consecutive steps on the same context (persistent buffers, stream/event reuse) equal three oracle
    steps; (b) masks that are NOT plain tmask products make the inner kernel read umask/vmask/wmask arrays."""
    G, GJ, K = 60, 41, 8
    gf = H.random_fields(O, G, GJ, K, 1, kjpt=2, seed=9)
    ref = gf
    for _ in range(3):
        out, _, _ = H.oracle_fct(O, ref, G, GJ, K, 1, 1, 1, 2, 4, 4)
        ref = dict(ref); ref["pta"] = out
    got, _ = H.device_fct(N, gf, G, GJ, K, 1, 1, 1, 2, 4, 4, schedule=1, nsteps=3)
    assert np.array_equal(got, ref["pta"])
    gf2 = dict(gf); gf2["umask"] = gf["umask"].copy(); gf2["umask"][2, 10:14, 20:30] = 0.0   # e.g. a closed strait
    ref2, _, _ = H.oracle_fct(O, gf2, G, GJ, K, 1, 1, 1, 2, 4, 4)
    got2, _ = H.device_fct(N, gf2, G, GJ, K, 1, 1, 1, 2, 4, 4, schedule=1)
    assert np.array_equal(got2, ref2) and not np.array_equal(ref2, ref["pta"])


# ---- schedule 2: the inner P1-P5 kernel with TMA-staged tiles (even jpi) ------------------------------------------
@pytest.mark.parametrize("jperio", [0, 1, 4, 6])
@pytest.mark.parametrize("hv", [(2, 2), (4, 2), (2, 4), (4, 4)])
def test_schedule2_tma_tiles_bit_exact(N, O, jperio, hv):
    h, v = hv
    G, GJ, K = 76, 45, 11                       # even jpiglo: TMA path; tiles overhang the inner rectangle
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=200 + 10 * jperio + h + v)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
    for schedule in (2, 3):
        got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 1, 1, 2, h, v, schedule=schedule)
        bad = np.argwhere(got != ref)
        assert np.array_equal(got, ref), "schedule %d: first mismatches (jn,k,j,i): %s of %d" % (schedule, bad[:6].tolist(), len(bad))


def test_schedule2_variants(N, O):
    """linssh / isfcav, non-product masks (array-reading template), odd jpi (falls back to the cp.async kernel),
    in-process 2x2 group, 3 steps."""
    G, GJ, K = 64, 50, 9
    for ln_linssh, ln_isfcav in [(True, False), (True, True)]:
        gf = H.random_fields(O, G, GJ, K, 4, kjpt=2, seed=31, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 4, 1, 1, 2, 4, 2, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        got, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, 2, 4, 2, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav, schedule=2)
        assert np.array_equal(got, ref)
    gf = H.random_fields(O, G, GJ, K, 1, kjpt=3, seed=32)
    gf["vmask"] = gf["vmask"].copy(); gf["vmask"][1, 20:24, 10:30] = 0.0
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 1, 1, 1, 3, 4, 4)
    got, _ = H.device_fct(N, gf, G, GJ, K, 1, 1, 1, 3, 4, 4, schedule=2)
    assert np.array_equal(got, ref)
    gf = H.random_fields(O, 63, GJ, K, 6, kjpt=2, seed=33)
    ref, _, _ = H.oracle_fct(O, gf, 63, GJ, K, 6, 1, 1, 2, 4, 4)
    got, _ = H.device_fct(N, gf, 63, GJ, K, 6, 1, 1, 2, 4, 4, schedule=2)
    assert np.array_equal(got, ref)
    gf = H.random_fields(O, 90, 70, K, 4, kjpt=2, seed=34)
    ref = gf
    for _ in range(3):
        out, _, _ = H.oracle_fct(O, ref, 90, 70, K, 4, 1, 1, 2, 4, 4)
        ref = dict(ref); ref["pta"] = out
    got, _ = H.device_fct(N, gf, 90, 70, K, 4, 2, 2, 2, 4, 4, schedule=2, nsteps=3)
    assert np.array_equal(got, ref["pta"])
