"""A second, independent statement of tra_adv_fct + nonosc in vectorised numpy (whole-array slices instead of the
oracle's triple loops), written from src/OCE/TRA/traadv_fct.F90:113-297 and :354-426 for a mono-domain with closed or
E-W cyclic boundaries and the 2nd-order vertical scheme.  Test infrastructure: two restatements in different styles and
languages that agree bit for bit leave little room for an indexing slip in either (the reference itself cannot be run here).

Arrays are (jpk, jpj, jpi) / (kjpt, jpk, jpj, jpi), i.e. the Fortran memory image; a Fortran index ji is numpy index ji-1.
Elementwise numpy arithmetic is IEEE fp64 without contraction, so keeping the reference's operation order gives identical bits."""
import numpy as np


def lbc(a, cd_nat, psgn, jperio):
    """lbc_lnk_generic.h90:80-106 for jperio 0 (closed) / 1 (E-W cyclic), mono-domain; psgn is irrelevant without a fold"""
    if jperio == 1:
        a[..., 0] = a[..., -2]
        a[..., -1] = a[..., 1]
    else:
        if cd_nat != "F":
            a[..., 0] = 0.0
        a[..., -1] = 0.0
    if cd_nat != "F":
        a[..., 0, :] = 0.0
    a[..., -1, :] = 0.0


def sign_half(x):
    """0.5 + SIGN(0.5, x) with the key_nosignedzero override: x >= 0 -> +"""
    return 0.5 + np.where(x >= 0.0, 0.5, -0.5)


def tra_adv_fct(f, kjpt, kn_fct_h, jperio, ln_linssh=False, capture=None):
    """capture: optional list that receives, per tracer, what the limiter + final trend start from (after X2, :280):
    dict(zwi, zwx, zwy, zwz, pta) -- used to test the fused limiter kernel in isolation"""
    tmask, umask, vmask, wmask = f["tmask"], f["umask"], f["vmask"], f["wmask"]
    e3t_b, e3t_n, e3t_a, r1, e12 = f["e3t_b"], f["e3t_n"], f["e3t_a"], f["r1_e1e2t"][None], f["e1e2t"][None]
    pun, pvn, pwn, p2dt = f["pun"], f["pvn"], f["pwn"], f["p2dt"]
    pta_all = f["pta"].copy()
    shp = tmask.shape
    zwi = np.zeros(shp); zwx = np.zeros(shp); zwy = np.zeros(shp); zwz = np.zeros(shp)
    I = (slice(0, -1), slice(1, -1), slice(1, -1))          # (1:jpkm1, 2:jpjm1, 2:jpim1)
    r1_6 = 1.0 / 6.0
    for jn in range(kjpt):
        ptb, ptn, pta = f["ptb"][jn], f["ptn"][jn], pta_all[jn]
        # upstream fluxes (:123-156)
        zfp = pun + np.abs(pun); zfm = pun - np.abs(pun)
        zwx[:-1, :-1, :-1] = (0.5 * (zfp[:, :, :-1] * ptb[:, :, :-1] + zfm[:, :, :-1] * ptb[:, :, 1:]))[:-1, :-1]
        zfp = pvn + np.abs(pvn); zfm = pvn - np.abs(pvn)
        zwy[:-1, :-1, :-1] = (0.5 * (zfp[:, :-1, :] * ptb[:, :-1, :] + zfm[:, :-1, :] * ptb[:, 1:, :]))[:-1, :, :-1]
        zfp = pwn + np.abs(pwn); zfm = pwn - np.abs(pwn)
        zwz[1:-1] = (0.5 * (zfp[1:-1] * ptb[1:-1] + zfm[1:-1] * ptb[:-2])) * wmask[1:-1]
        if ln_linssh:
            zwz[0] = pwn[0] * ptb[0]
        # trend and low-order field (:158-170)
        ztra = -((zwx[:-1, 1:-1, 1:-1] - zwx[:-1, 1:-1, :-2]) + zwy[:-1, 1:-1, 1:-1] - zwy[:-1, :-2, 1:-1]
                 + zwz[:-1, 1:-1, 1:-1] - zwz[1:, 1:-1, 1:-1]) * r1[:, 1:-1, 1:-1]
        pta[I] = pta[I] + ztra / e3t_n[I] * tmask[I]
        zwi[I] = (e3t_b[I] * ptb[I] + p2dt * ztra) / e3t_a[I] * tmask[I]
        # anti-diffusive fluxes (:180-278)
        if kn_fct_h == 2:
            zwx[:-1, :-1, :-1] = (0.5 * pun[:, :, :-1] * (ptn[:, :, :-1] + ptn[:, :, 1:]))[:-1, :-1] - zwx[:-1, :-1, :-1]
            zwy[:-1, :-1, :-1] = (0.5 * pvn[:, :-1, :] * (ptn[:, :-1, :] + ptn[:, 1:, :]))[:-1, :, :-1] - zwy[:-1, :-1, :-1]
        else:
            ztu = np.zeros(shp); ztv = np.zeros(shp); zltu = np.zeros(shp); zltv = np.zeros(shp)
            ztu[:-1, :-1, :-1] = ((ptn[:, :, 1:] - ptn[:, :, :-1]) * umask[:, :, :-1])[:-1, :-1]
            ztv[:-1, :-1, :-1] = ((ptn[:, 1:, :] - ptn[:, :-1, :]) * vmask[:, :-1, :])[:-1, :, :-1]
            zltu[I] = (ztu[:-1, 1:-1, 1:-1] + ztu[:-1, 1:-1, :-2]) * r1_6           # '+' as in the reference (:204)
            zltv[I] = (ztv[:-1, 1:-1, 1:-1] + ztv[:-1, :-2, 1:-1]) * r1_6
            lbc(zltu, "T", 1.0, jperio); lbc(zltv, "T", 1.0, jperio)
            zC2t_u = ptn[:-1, :-1, :-1] + ptn[:-1, :-1, 1:]
            zC2t_v = ptn[:-1, :-1, :-1] + ptn[:-1, 1:, :-1]
            zwx[:-1, :-1, :-1] = 0.5 * pun[:-1, :-1, :-1] * (zC2t_u + zltu[:-1, :-1, :-1] - zltu[:-1, :-1, 1:]) - zwx[:-1, :-1, :-1]
            zwy[:-1, :-1, :-1] = 0.5 * pvn[:-1, :-1, :-1] * (zC2t_v + zltv[:-1, :-1, :-1] - zltv[:-1, 1:, :-1]) - zwy[:-1, :-1, :-1]
        zwz[1:-1, 1:-1, 1:-1] = ((pwn[1:-1] * 0.5 * (ptn[1:-1] + ptn[:-2]) - zwz[1:-1]) * wmask[1:-1])[:, 1:-1, 1:-1]
        if ln_linssh:
            zwz[0] = 0.0
        lbc(zwi, "T", 1.0, jperio); lbc(zwx, "U", -1.0, jperio); lbc(zwy, "V", -1.0, jperio); lbc(zwz, "W", 1.0, jperio)
        if capture is not None:
            capture.append(dict(zwi=zwi.copy(), zwx=zwx.copy(), zwy=zwy.copy(), zwz=zwz.copy(), pta=pta.copy()))
        nonosc(ptb, zwx, zwy, zwz, zwi, p2dt, tmask, e3t_n, e12, jperio)
        # final trend (:288-297)
        pta[I] = pta[I] - ((zwx[:-1, 1:-1, 1:-1] - zwx[:-1, 1:-1, :-2]) + zwy[:-1, 1:-1, 1:-1] - zwy[:-1, :-2, 1:-1]
                           + zwz[:-1, 1:-1, 1:-1] - zwz[1:, 1:-1, 1:-1]) * r1[:, 1:-1, 1:-1] / e3t_n[I]
    return pta_all


def nonosc(pbef, paa, pbb, pcc, paft, p2dt, tmask, e3t_n, e12, jperio):
    """traadv_fct.F90:354-426, in place on paa, pbb, pcc"""
    zbig, zrtrn = 1.0e40, 1.0e-15
    shp = pbef.shape
    zbetup = np.zeros(shp); zbetdo = np.zeros(shp)
    zbup = np.maximum(pbef * tmask - zbig * (1.0 - tmask), paft * tmask - zbig * (1.0 - tmask))
    zbdo = np.minimum(pbef * tmask + zbig * (1.0 - tmask), paft * tmask + zbig * (1.0 - tmask))
    C = (slice(0, -1), slice(1, -1), slice(1, -1))
    W = (slice(0, -1), slice(1, -1), slice(0, -2)); E = (slice(0, -1), slice(1, -1), slice(2, None))
    S = (slice(0, -1), slice(0, -2), slice(1, -1)); Nn = (slice(0, -1), slice(2, None), slice(1, -1))
    D = (slice(1, None), slice(1, -1), slice(1, -1))        # jk+1
    up_km1 = np.concatenate([zbup[:1], zbup[:-2]], axis=0)[:, 1:-1, 1:-1]      # ikm1 = MAX(jk-1, 1)
    do_km1 = np.concatenate([zbdo[:1], zbdo[:-2]], axis=0)[:, 1:-1, 1:-1]
    zup = zbup[C]
    for x in (zbup[W], zbup[E], zbup[S], zbup[Nn], up_km1, zbup[D]):
        zup = np.maximum(zup, x)
    zdo = zbdo[C]
    for x in (zbdo[W], zbdo[E], zbdo[S], zbdo[Nn], do_km1, zbdo[D]):
        zdo = np.minimum(zdo, x)
    p, m = (lambda x: np.maximum(0.0, x)), (lambda x: np.minimum(0.0, x))
    zpos = p(paa[W]) - m(paa[C]) + p(pbb[S]) - m(pbb[C]) + p(pcc[D]) - m(pcc[C])
    zneg = p(paa[C]) - m(paa[W]) + p(pbb[C]) - m(pbb[S]) + p(pcc[C]) - m(pcc[D])
    zbt = e12[:, 1:-1, 1:-1] * e3t_n[C] / p2dt
    zbetup[C] = (zup - paft[C]) / (zpos + zrtrn) * zbt
    zbetdo[C] = (paft[C] - zdo) / (zneg + zrtrn) * zbt
    lbc(zbetup, "T", 1.0, jperio); lbc(zbetdo, "T", 1.0, jperio)
    one = 1.0
    zau = np.minimum(np.minimum(one, zbetdo[C]), zbetup[E]); zbu = np.minimum(np.minimum(one, zbetup[C]), zbetdo[E])
    zcu = sign_half(paa[C])
    new_aa = paa[C] * (zcu * zau + (1.0 - zcu) * zbu)
    zav = np.minimum(np.minimum(one, zbetdo[C]), zbetup[Nn]); zbv = np.minimum(np.minimum(one, zbetup[C]), zbetdo[Nn])
    zcv = sign_half(pbb[C])
    new_bb = pbb[C] * (zcv * zav + (1.0 - zcv) * zbv)
    za = np.minimum(np.minimum(one, zbetdo[D]), zbetup[C]); zb = np.minimum(np.minimum(one, zbetup[D]), zbetdo[C])
    zc = sign_half(pcc[D])
    new_cc = pcc[D] * (zc * za + (1.0 - zc) * zb)
    paa[C] = new_aa; pbb[C] = new_bb; pcc[D] = new_cc
    lbc(paa, "U", -1.0, jperio); lbc(pbb, "V", -1.0, jperio)
