"""GPU parity of schedule 4 -- the WHOLE FCT step of the inner region in ONE kernel (k_fct_fused: P1-P8 from the input fields,
no intermediate array in HBM; frame chain with X1..X4 on the side stream) -- against the CPU oracle, bit for bit.
Reference: src/OCE/TRA/traadv_fct.F90:54-327 (tra_adv_fct) + :330-428 (nonosc)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _check(got, ref, what=""):
    bad = np.argwhere(got != ref)
    assert np.array_equal(got, ref), "%s: first mismatches (jn,k,j,i): %s of %d" % (what, bad[:6].tolist(), len(bad))


@pytest.mark.parametrize("jperio", [0, 1, 4, 6])
@pytest.mark.parametrize("hv", [(2, 2), (4, 2), (2, 4), (4, 4)])
def test_one_kernel_bit_exact(N, O, jperio, hv):
    h, v = hv
    G, GJ, K = 76, 45, 11                       # even jpiglo: TMA path; 3 x 4 tiles overhanging the output rectangle
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=400 + 10 * jperio + h + v)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
    ran = []
    got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 1, 1, 2, h, v, schedule=4, kernels=ran)
    _check(got, ref, "schedule 4")
    assert "fct_fused" in ran and "fct_nonosc_final" not in ran and "fct_low_antidiff_inner" not in ran, ran
    assert not np.array_equal(ref, gf["pta"])


def test_one_kernel_options_and_fallbacks(N, O):
    """linssh / isfcav; three tracers; non-product masks and odd jpi (both fall back to the three-kernel schedule);
    in-process 2x2 group over 3 steps (fold partners in different subdomains)."""
    G, GJ, K = 64, 50, 9
    for ln_linssh, ln_isfcav in [(True, False), (True, True)]:
        gf = H.random_fields(O, G, GJ, K, 4, kjpt=2, seed=41, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 4, 1, 1, 2, 4, 2, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav)
        got, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, 2, 4, 2, ln_linssh=ln_linssh, ln_isfcav=ln_isfcav, schedule=4)
        _check(got, ref, "linssh=%s isfcav=%s" % (ln_linssh, ln_isfcav))
    gf = H.random_fields(O, G, GJ, K, 1, kjpt=3, seed=42)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 1, 1, 1, 3, 4, 4)
    got, _ = H.device_fct(N, gf, G, GJ, K, 1, 1, 1, 3, 4, 4, schedule=4)
    _check(got, ref, "kjpt=3")
    gf["vmask"] = gf["vmask"].copy(); gf["vmask"][1, 20:24, 10:30] = 0.0
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 1, 1, 1, 3, 4, 4)
    got, _ = H.device_fct(N, gf, G, GJ, K, 1, 1, 1, 3, 4, 4, schedule=4)
    _check(got, ref, "non-product masks")
    gf = H.random_fields(O, 63, GJ, K, 6, kjpt=2, seed=43)
    ref, _, _ = H.oracle_fct(O, gf, 63, GJ, K, 6, 1, 1, 2, 4, 4)
    got, _ = H.device_fct(N, gf, 63, GJ, K, 6, 1, 1, 2, 4, 4, schedule=4)
    _check(got, ref, "odd jpi")
    gf = H.random_fields(O, 90, 70, K, 4, kjpt=2, seed=44)
    ref = gf
    for _ in range(3):
        out, _, _ = H.oracle_fct(O, ref, 90, 70, K, 4, 1, 1, 2, 4, 4)
        ref = dict(ref); ref["pta"] = out
    got, _ = H.device_fct(N, gf, 90, 70, K, 4, 2, 2, 2, 4, 4, schedule=4, nsteps=3)
    _check(got, ref["pta"], "2x2 group, 3 steps")


@pytest.mark.parametrize("jperio,hv", [(4, (4, 4)), (0, (2, 2))])
def test_one_kernel_production_depth_with_jk_chunks(N, O, jperio, hv):
    """jpk = 75: with few tiles the launcher splits the jk loop across grid z (each chunk restarts the three-stage pipeline
    two levels above its first output level)"""
    h, v = hv
    G, GJ, K = 96, 58, 75
    gf = H.random_fields(O, G, GJ, K, jperio, kjpt=2, seed=450 + jperio)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, jperio, 1, 1, 2, h, v)
    got, _ = H.device_fct(N, gf, G, GJ, K, jperio, 1, 1, 2, h, v, schedule=4)
    _check(got, ref, "jpk=75")


def test_one_kernel_orca2_pisces_shape_26_tracers(N, O):
    """BASELINE config C5: ORCA2_ICE_PISCES-shaped 182x149x31, T/S + 24 passive tracers in one call (trcadv.F90:127),
    2nd order FCT, jperio = 4 -- and the same 26 tracers with the 4th-order / compact scheme"""
    G, GJ, K, kjpt = 182, 149, 31, 26
    gf = H.random_fields(O, G, GJ, K, 4, kjpt=kjpt, seed=460)
    for (h, v) in ((2, 2), (4, 4)):
        ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 4, 1, 1, kjpt, h, v)
        for schedule in (4, 2, 0):
            got, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, kjpt, h, v, schedule=schedule)
            _check(got, ref, "C5 h%d v%d schedule %d" % (h, v, schedule))


def test_inlined_division_is_the_ieee_division(N):
    """div_rn of the fused kernel (the compiler's fast path written inline, fct_fused_kernel.cuh) == x / y bit for bit on
    2^24 operand pairs in each of 8 classes: ordinary, the FCT ranges (1e-15 .. 1e40), tiny numerators, zeros, subnormals,
    huge, Inf / NaN, anything"""
    for seed in (1, 2, 3):
        assert N.selftest_division(1 << 24, seed) == 0


def test_fast_arithmetic_is_opt_in_and_close(N, O):
    """nemo_fct_set_arithmetic(FAST): divisions as multiplications by a refined reciprocal (<= 1 ulp each).  Not bit-identical
    (that is why STRICT is the default): the absolute difference stays at round-off of the field scale, while the per-point
    relative difference exceeds the 1e-12 bar of BASELINE.json on near-zero trends (measured 5e-11 at ORCA025)."""
    G, GJ, K = 76, 45, 11
    gf = H.random_fields(O, G, GJ, K, 4, kjpt=2, seed=470)
    ref, _, _ = H.oracle_fct(O, gf, G, GJ, K, 4, 1, 1, 2, 4, 4)
    strict, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, 2, 4, 4, schedule=4)
    fast, _ = H.device_fct(N, gf, G, GJ, K, 4, 1, 1, 2, 4, 4, schedule=4, arith=1)
    assert np.array_equal(strict, ref)
    assert not np.array_equal(fast, ref)
    assert np.abs(fast - ref).max() <= 1e-13 * np.abs(ref).max()
    with pytest.raises(N.NemoFctError, match="arithmetic"):
        N.FctContext(N.mpp_init(G, GJ, K, 4, 1, 1, 1), 0).set_arithmetic(7)
