"""ref_exec -- ORACLE-SIDE TEST INFRASTRUCTURE: runs the reference's own tra_adv_fct / nonosc / interp_4th_cpt source
(src/OCE/TRA/traadv_fct.F90, read from the reference tree at run time, never copied) through oracle/f90exec.py on numpy arrays.
Only usable where the reference tree exists (this container); the GPU box sees the golden vectors generated from it
(tests/golden/ref_exec_*.npz, oracle/_ref_recipe/make_ref_exec_golden.py)."""
import os

import numpy as np

from . import f90exec

REF_ROOT = os.environ.get("NEMO_REFERENCE_ROOT", "/root/reference")
DOM_ARRAYS = ("tmask", "umask", "vmask", "wmask", "e3t_b", "e3t_n", "e3t_a", "e1e2t", "r1_e1e2t")
DOM_INT_ARRAYS = ("mikt", "mbkt")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "src", "OCE", "TRA", "traadv_fct.F90"))


def _read(*rel):
    with open(os.path.join(REF_ROOT, *rel)) as f:
        return f.read()


def F(a):
    """Fortran view a(jpi,jpj[,jpk[,kjpt]]) of a C-ordered array [kjpt][jpk][jpj][jpi]: same memory"""
    return np.transpose(a)


class RefDomain:
    """the module variables the tracer routines read (dom_oce, par_oce, in_out_manager, trd_oce, diaptr) for ONE mono-processor
    domain, plus lbc_lnk / lbc_lnk_multi.  `lbc` = callable(list of (C-ordered array, nature, sign)) doing the exchange in place:
    reference_lbc() below, i.e. the reference's own lbc_lnk / lbc_nfd text (lbc_lnk_multi is a loop over its fields,
    lbc_lnk_multi_generic.h90)."""

    def __init__(self, gf, jpi, jpj, jpk, ln_linssh, ln_isfcav, lbc, sign_mode="nosignedzero", undef=np.nan):
        ns = {}
        # what an automatic (stack) array holds before it is defined: NaN shows every read of an undefined element that reaches
        # the result; 0.0 is one value the memory may hold (what the oracle's calloc'ed work arrays hold)
        ns["f_alloc"] = lambda shape, integer=False: (np.zeros(tuple(int(n) for n in shape), dtype=np.int32, order="F") if integer
                                                      else np.full(tuple(int(n) for n in shape), undef, order="F"))
        ns.update(jpi=jpi, jpj=jpj, jpk=jpk, jpim1=jpi - 1, jpjm1=jpj - 1, jpkm1=jpk - 1, ln_linssh=bool(ln_linssh), ln_isfcav=bool(ln_isfcav),
                  lwp=False, numout=6, l_trdtra=False, l_trdtrc=False, ln_diaptr=False, iom_use=lambda name: False)
        for k in DOM_ARRAYS:
            ns[k] = F(np.ascontiguousarray(gf[k]))
        for k in DOM_INT_ARRAYS:
            ns[k] = F(np.ascontiguousarray(gf[k], dtype=np.int32))

        def lbc_lnk_multi(cdname, *triples):
            assert len(triples) % 3 == 0
            lbc([(np.transpose(triples[i]), triples[i + 1], float(triples[i + 2])) for i in range(0, len(triples), 3)])

        def lbc_lnk(cdname, a, nat, sgn):
            lbc([(np.transpose(a), nat, float(sgn))])
        ns.update(lbc_lnk_multi=lbc_lnk_multi, lbc_lnk=lbc_lnk)
        self.ns = ns
        self.defines = f90exec.cpp_defines(_read("src", "OCE", "vectopt_loop_substitute.h90"))      # fs_2, fs_jpim1 (no key_vectopt_loop)
        if sign_mode == "nosignedzero":
            # the reference's gfortran arch file builds with -Dkey_nosignedzero (arch/arch-linux_gfortran.fcm:46): SIGN is then
            # lib_fortran's SIGN_SCALAR (lib_fortran.F90:339-351), taken here from the reference's own text
            f90exec.load(_read("src", "OCE", "lib_fortran.F90"), ns, only=("sign_scalar",), defined=("key_nosignedzero",))
            scalar = ns["sign_scalar"]
            ns["f_sign"] = lambda a, b: (np.where(np.asarray(b) >= 0.0, np.abs(a), -np.abs(a))
                                         if isinstance(a, np.ndarray) or isinstance(b, np.ndarray) else scalar(a, b))
        else:
            assert sign_mode == "ieee"

    def load(self, *rel, only=None):
        text = _read(*rel)
        f90exec.module_parameters(text, self.ns, self.defines)
        return f90exec.load(text, self.ns, arrays=DOM_ARRAYS, int_arrays=DOM_INT_ARRAYS, defines=self.defines, only=only)


def reference_lbc(jperio, jpi, jpj):
    """lbc_lnk of the reference itself for ONE mono-processor domain, from its text: lbc_lnk_generic.h90 (the build without
    key_mpp_mpi, lbclnk.F90:59-170, 3-D variant) calling lbc_nfd_generic.h90 (lbcnfd.F90, 3-D variant) for the north folds.  The
    scalars it reads are those mpp_init sets for jpni = jpnj = 1 (mppini.F90:88-101: npolj = jperio for the folds, l_Iperio /
    l_Jperio from the cyclic types)."""
    ns = dict(jpi=jpi, jpj=jpj, jpim1=jpi - 1, jpjm1=jpj - 1, jpiglo=jpi, jpjglo=jpj, jpni=1, jpnj=1, nlci=jpi, nlcj=jpj, jperio=jperio,
              npolj=jperio if jperio in (3, 4, 5, 6) else 0, l_iperio=jperio in (1, 4, 6, 7), l_jperio=jperio in (2, 7))
    nfd = f90exec.cpp(_read("src", "OCE", "LBC", "lbc_nfd_generic.h90"), defined=("DIM_3d",), macros={"ROUTINE_NFD": "lbc_nfd_3d"})
    lnk = f90exec.cpp(_read("src", "OCE", "LBC", "lbc_lnk_generic.h90"), defined=("DIM_3d",), macros={"ROUTINE_LNK": "lbc_lnk_3d"})
    f90exec.load(nfd, ns)
    f90exec.load(lnk, ns)
    ns["lbc_nfd"] = ns["lbc_nfd_3d"]                                   # INTERFACE lbc_nfd (lbcnfd.F90:27-31)

    def lbc(items):
        for a, nat, sgn in items:                                      # a: C-ordered [...][jpj][jpi]: 2-D, 3-D or 4-D
            a3 = a.reshape((-1,) + a.shape[-2:])                       # (a 4-D field is a 3-D field with jpk * kjpt levels)
            assert np.shares_memory(a3, a)
            ns["lbc_lnk_3d"]("ref_exec", np.transpose(a3), nat, float(sgn))
    return lbc


def tra_adv_fct(gf, jpi, jpj, jpk, kjpt, kn_fct_h, kn_fct_v, ln_linssh, ln_isfcav, lbc, cdtype="TRA", sign_mode="nosignedzero", trends=None):
    """the reference's tra_adv_fct on the C-ordered fields of `gf` (helpers.random_fields); returns the new pta.  `trends`: a dict
    that switches l_trdtra on and receives what the routine hands to trd_tra (traadv_fct.F90:96-112, 172-176, 299-308): C-ordered
    copies of ztrdx / ztrdy / ztrdz under the keys (jn, 'x' | 'y' | 'z')."""
    dom = RefDomain(gf, jpi, jpj, jpk, ln_linssh, ln_isfcav, lbc, sign_mode)
    if trends is not None:
        dom.ns.update(l_trdtra=True, l_trdtrc=True, jptra_xad=1, jptra_yad=2, jptra_zad=3)
        dom.ns["trd_tra"] = lambda kt, cd, jn, ktrd, ptrd, pu=None, ptra=None: trends.__setitem__(
            (int(jn), "xyz"[int(ktrd) - 1]), np.ascontiguousarray(np.transpose(ptrd)).copy())
    dom.load("src", "OCE", "TRA", "traadv_fct.F90", only=("tra_adv_fct", "nonosc", "interp_4th_cpt"))
    pta = np.array(gf["pta"], copy=True)
    args = [F(np.ascontiguousarray(gf[k])) for k in ("pun", "pvn", "pwn", "ptb", "ptn")]
    dom.ns["tra_adv_fct"](1, 1, cdtype, float(gf["p2dt"]), args[0], args[1], args[2], args[3], args[4], F(pta), kjpt, kn_fct_h, kn_fct_v)
    return pta


MUS_ARRAYS = ("r1_e1e2u", "r1_e1e2v", "e3u_n", "e3v_n", "e3w_n", "rnfmsk", "rnfmsk_z")
NXT_ARRAYS = ("tsb", "tsn", "tsa", "emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf", "h_rnf", "qsr_hc", "qsr_hc_b", "rnf_tsc", "rnf_tsc_b",
              "risf_tsc", "risf_tsc_b", "r1_hisf_tbl", "ralpha", "sbc_tsc", "sbc_tsc_b")
NXT_INT_ARRAYS = ("nk_rnf", "misfkt", "misfkb")


def tra_adv_mus(gf, mx, jpi, jpj, jpk, kjpt, ln_linssh, ln_isfcav, ld_msc_ups, lbc, cdtype="TRA"):
    """the reference's tra_adv_mus (src/OCE/TRA/traadv_mus.F90) called at kt = kit000, so that it builds its own upstream
    indicator xind from rnfmsk / rnfmsk_z (:98-113); returns the new pta"""
    dom = RefDomain(gf, jpi, jpj, jpk, ln_linssh, ln_isfcav, lbc)
    for k in ("r1_e1e2u", "r1_e1e2v", "e3u_n", "e3v_n", "e3w_n"):
        dom.ns[k] = F(np.ascontiguousarray(mx[k]))
    if ld_msc_ups:
        dom.ns["rnfmsk"] = F(np.ascontiguousarray(mx["rnfmsk"]))
        dom.ns["rnfmsk_z"] = np.ascontiguousarray(mx["rnfmsk_z"])
    text = _read("src", "OCE", "TRA", "traadv_mus.F90")
    f90exec.module_parameters(text, dom.ns, dom.defines)
    f90exec.load(text, dom.ns, arrays=DOM_ARRAYS + MUS_ARRAYS, int_arrays=DOM_INT_ARRAYS, defines=dom.defines, only=("tra_adv_mus",))
    pta = np.array(gf["pta"], copy=True)
    args = [F(np.ascontiguousarray(gf[k])) for k in ("pun", "pvn", "pwn", "ptb")]
    dom.ns["tra_adv_mus"](1, 1, cdtype, float(gf["p2dt"]), args[0], args[1], args[2], args[3], F(pta), kjpt, bool(ld_msc_ups))
    return pta


def tra_adv_cen(gf, jpi, jpj, jpk, kjpt, kn_cen_h, kn_cen_v, ln_linssh, ln_isfcav, lbc, cdtype="TRA", undef=np.nan):
    """the reference's tra_adv_cen (src/OCE/TRA/traadv_cen.F90; its compact vertical scheme calls the reference's
    interp_4th_cpt of traadv_fct.F90); returns the new pta"""
    dom = RefDomain(gf, jpi, jpj, jpk, ln_linssh, ln_isfcav, lbc, undef=undef)
    dom.load("src", "OCE", "TRA", "traadv_fct.F90", only=("interp_4th_cpt",))
    dom.load("src", "OCE", "TRA", "traadv_cen.F90", only=("tra_adv_cen",))
    pta = np.array(gf["pta"], copy=True)
    args = [F(np.ascontiguousarray(gf[k])) for k in ("pun", "pvn", "pwn", "ptn")]
    dom.ns["tra_adv_cen"](1, 1, cdtype, args[0], args[1], args[2], args[3], F(pta), kjpt, kn_cen_h, kn_cen_v)
    return pta


def tra_nxt(gf, extra, jpi, jpj, jpk, kt, nit000, neuler, rdt, atfp, r1_rau0, ln_linssh, lbc, flags=None):
    """the reference's tra_nxt driver (src/OCE/TRA/tranxt.F90:65-177: lbc_lnk on tsa, tra_nxt_fix or tra_nxt_vvl, lbc_lnk on tsb, tsn,
    tsa) on tsb / tsn / tsa = the case's ptb / ptn / pta (jpts = 2); `flags` (ln_traqsr, ln_rnf, ln_isf, ln_rnf_depth, nksr) switch on
    the forcings whose arrays `extra` then holds; no BDY, AGRIF or trend diagnostics; returns (tsb, tsn, tsa)"""
    dom = RefDomain(gf, jpi, jpj, jpk, ln_linssh, False, lbc)
    ts = {k: np.array(gf[s], copy=True) for k, s in (("tsb", "ptb"), ("tsn", "ptn"), ("tsa", "pta"))}
    ns = dom.ns
    for k, a in ts.items():
        ns[k] = F(a)
    ns.update(jpts=2, jp_tem=1, jp_sal=2, nit000=nit000, neuler=neuler, rdt=float(rdt), r2dt=0.0, atfp=float(atfp), r1_rau0=float(r1_rau0),
              ln_timing=False, ln_bdy=False, ln_traldf_iso=False, ln_ctl=False, ln_traqsr=False, ln_rnf=False, ln_isf=False,
              ln_rnf_depth=False, nksr=0)
    for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf"):
        ns[k] = F(np.ascontiguousarray(extra[k]))
    ns["sbc_tsc"], ns["sbc_tsc_b"] = F(np.ascontiguousarray(extra["sbc"])), F(np.ascontiguousarray(extra["sbc_b"]))
    # the optional forcings of tra_nxt_vvl (tranxt.F90:262-343): solar penetration, runoffs (surface or over depth), ice shelves
    ns.update({k: (int(v) if k == "nksr" else bool(v)) for k, v in (flags or {}).items()})
    for k in ("h_rnf", "r1_hisf_tbl", "ralpha", "rnf_tsc", "rnf_tsc_b", "risf_tsc", "risf_tsc_b", "qsr_hc", "qsr_hc_b"):
        if k in extra:
            ns[k] = F(np.ascontiguousarray(extra[k]))
    for k in NXT_INT_ARRAYS:
        if k in extra:
            ns[k] = F(np.ascontiguousarray(extra[k], dtype=np.int32))
    text = _read("src", "OCE", "TRA", "tranxt.F90")
    f90exec.load(text, ns, arrays=DOM_ARRAYS + NXT_ARRAYS, int_arrays=DOM_INT_ARRAYS + NXT_INT_ARRAYS, defines=dom.defines, module_vars=("r2dt",))
    ns["tra_nxt"](kt)
    return ts["tsb"], ts["tsn"], ts["tsa"]


def tra_adv(gf, vel, jpi, jpj, jpk, kt, nit000, neuler, rdt, nn_fct_h, nn_fct_v, ln_linssh, ln_isfcav, lbc):
    """the reference's tra_adv driver (src/OCE/TRA/traadv.F90:77-175) on the FCT branch: r2dt rule (:88-90), effective transports
    from un, vn, wn (:93-118; no Stokes drift, z-tilde, eiv or mle additions), then ITS call of tra_adv_fct (:150) on tsb, tsn, tsa =
    the case's ptb, ptn, pta.  vel = dict(e2u, e1v, e3u_n, e3v_n, un, vn, wn).  Returns (tsa, r2dt)."""
    dom = RefDomain(gf, jpi, jpj, jpk, ln_linssh, ln_isfcav, lbc)
    ns = dom.ns
    ts = {k: np.array(gf[s], copy=True) for k, s in (("tsb", "ptb"), ("tsn", "ptn"), ("tsa", "pta"))}
    for k, a in ts.items():
        ns[k] = F(a)
    for k in ("e2u", "e1v", "e3u_n", "e3v_n", "un", "vn", "wn"):
        ns[k] = F(np.ascontiguousarray(vel[k]))
    ns.update(jpts=ts["tsa"].shape[0], jp_tem=1, jp_sal=2, nit000=nit000, neuler=neuler, rdt=float(rdt), r2dt=0.0, ln_timing=False,
              ln_wave=False, ln_sdw=False, ln_vvl_ztilde=False, ln_vvl_layer=False, ln_ldfeiv=False, ln_traldf_triad=False, ln_mle=False,
              ln_ctl=False, nn_fct_h=nn_fct_h, nn_fct_v=nn_fct_v, iom_put=lambda *a: None)
    dom.load("src", "OCE", "TRA", "traadv_fct.F90", only=("tra_adv_fct", "nonosc", "interp_4th_cpt"))
    text = _read("src", "OCE", "TRA", "traadv.F90")
    f90exec.module_parameters(text, ns, dom.defines)                   # np_CEN, np_FCT, ... (traadv.F90:58-63)
    f90exec.load(text, ns, arrays=DOM_ARRAYS + NXT_ARRAYS + ("e2u", "e1v", "e3u_n", "e3v_n", "un", "vn", "wn"), int_arrays=DOM_INT_ARRAYS,
                 defines=dom.defines, only=("tra_adv",), module_vars=("r2dt",))
    ns["nadv"] = ns["np_fct"]
    ns["tra_adv"](kt)
    return ts["tsa"], ns["r2dt"]


def trc_adv(gf, vel, trb, trn, tra, jpi, jpj, jpk, kt, nittrc000, r2dttrc, nn_fct_h, nn_fct_v, ln_linssh, ln_isfcav, lbc):
    """the reference's trc_adv driver (src/TOP/TRP/trcadv.F90:70-145) on the FCT branch: effective transports (:88-101), then ITS call
    tra_adv_fct( kt, nittrc000, 'TRC', r2dttrc, zun, zvn, zwn, trb, trn, tra, jptra, nn_fct_h, nn_fct_v ) (:127).  trb, trn, tra:
    C-ordered [jptra][jpk][jpj][jpi]; returns the new tra."""
    dom = RefDomain(gf, jpi, jpj, jpk, ln_linssh, ln_isfcav, lbc)
    ns = dom.ns
    out = np.array(tra, copy=True)
    ns.update(trb=F(np.ascontiguousarray(trb)), trn=F(np.ascontiguousarray(trn)), tra=F(out), jptra=out.shape[0])
    for k in ("e2u", "e1v", "e3u_n", "e3v_n", "un", "vn", "wn"):
        ns[k] = F(np.ascontiguousarray(vel[k]))
    ns.update(nittrc000=nittrc000, r2dttrc=float(r2dttrc), ln_timing=False, l_offline=False, ln_wave=False, ln_sdw=False, ln_vvl_ztilde=False,
              ln_vvl_layer=False, ln_ldfeiv=False, ln_traldf_triad=False, ln_mle=False, ln_ctl=False, nn_fct_h=nn_fct_h, nn_fct_v=nn_fct_v)
    dom.load("src", "OCE", "TRA", "traadv_fct.F90", only=("tra_adv_fct", "nonosc", "interp_4th_cpt"))
    text = _read("src", "TOP", "TRP", "trcadv.F90")
    f90exec.module_parameters(text, ns, dom.defines, defined=("key_top",))        # np_CEN, np_FCT, ... (trcadv.F90:52-57)
    f90exec.load(text, ns, arrays=DOM_ARRAYS + ("trb", "trn", "tra", "e2u", "e1v", "e3u_n", "e3v_n", "un", "vn", "wn"), int_arrays=DOM_INT_ARRAYS,
                 defines=dom.defines, only=("trc_adv",), defined=("key_top",))
    ns["nadv"] = ns["np_fct"]
    ns["trc_adv"](kt)
    return out


def mpp_basic_decomposition(jpiglo, jpjglo, jperio, knbi, knbj):
    """the reference's mpp_basic_decomposition (src/OCE/LBC/mppini.F90:695-798, nn_hls = 1) from its text:
    (kimax, kjmax, kimppt, kjmppt, klci, klcj), tables shaped (knbj, knbi) like the product's; raises if the reference calls ctl_stop"""
    def stop(*a):
        raise ValueError("ctl_stop: " + " ".join(str(x) for x in a))
    ns = dict(jpiglo=jpiglo, jpjglo=jpjglo, jperio=jperio, nn_hls=1, ctmp1="", ctmp2="", ctl_stop=stop)
    f90exec.load(_read("src", "OCE", "LBC", "mppini.F90"), ns, only=("mpp_basic_decomposition",), defined=("key_mpp_mpi",))
    tabs = [np.zeros((knbi, knbj), np.int32, order="F") for _ in range(4)]
    out = ns["mpp_basic_decomposition"](knbi, knbj, 0, 0, *tabs)
    return (out["kimax"], out["kjmax"]) + tuple(np.ascontiguousarray(t.T) for t in tabs)


def glob_sum_3d(ptab, tmask_i):
    """the reference's glob_sum for a 3-D field on ONE domain, from its text: lib_fortran_generic.h90:30-65 (GLOBSUM_CODE, DIM_3d,
    OPERATION_GLOBSUM => masked by tmask_i) accumulating with DDPDD (lib_fortran.F90:300-332; the (sum, error) pair is a COMPLEX);
    mpp_sum is the identity on one rank.  ptab: C-ordered [jpk][jpj][jpi]; returns (REAL(ctmp), AIMAG(ctmp) is not returned by the
    reference) -- here the fp64 result only."""
    ns = dict(tmask_i=F(np.ascontiguousarray(tmask_i)), mpp_sum=lambda cdname, c: None)
    f90exec.load(_read("src", "OCE", "lib_fortran.F90"), ns, only=("ddpdd",))
    gen = f90exec.cpp(_read("src", "OCE", "lib_fortran_generic.h90"), defined=("GLOBSUM_CODE", "DIM_3d", "OPERATION_GLOBSUM"),
                      macros={"FUNCTION_GLOBSUM": "glob_sum_3d"})
    f90exec.load(gen, ns, arrays=("tmask_i",), only=("glob_sum_3d",))
    return ns["glob_sum_3d"]("ref_exec", F(np.ascontiguousarray(ptab)))


def decomposition_tables(jpiglo, jpjglo, jperio, jpni, jpnj):
    """the per-rank tables mpp_init derives from mpp_basic_decomposition for an all-ocean layout (mppini.F90:255-300: rank r holds
    subdomain (ii, ij) with r = (ij-1)*jpni + ii - 1), built from the REFERENCE's mpp_basic_decomposition text: dict of Fortran-
    shaped arrays nimppt, njmppt, nlcit, nlcjt (jpnij) and nfiimpp, nfilcit, nfipproc (jpni, jpnj)"""
    kimax, kjmax, kimppt, kjmppt, klci, klcj = mpp_basic_decomposition(jpiglo, jpjglo, jperio, jpni, jpnj)
    t = dict(jpimax=int(kimax), jpjmax=int(kjmax))
    for name, tab in (("nimppt", kimppt), ("njmppt", kjmppt), ("nlcit", klci), ("nlcjt", klcj)):
        t[name] = np.ascontiguousarray(tab.reshape(-1), dtype=np.int32)            # (knbj, knbi) C order == rank order
    t["nfiimpp"] = np.asfortranarray(kimppt.T.astype(np.int32))
    t["nfilcit"] = np.asfortranarray(klci.T.astype(np.int32))
    t["nfipproc"] = np.asfortranarray(np.arange(jpni * jpnj, dtype=np.int32).reshape(jpnj, jpni).T)
    return t


def mpp_init_nfdcom(jpiglo, jpjglo, jperio, jpni, jpnj, narea, nlci, nldi, nlei):
    """the reference's mpp_init_nfdcom (src/OCE/LBC/mppini.F90:1180-1240) for rank `narea` (1-based) from its text: the no-gather
    fold partners (nsndto, isendto) and nfsloop / nfeloop"""
    t = decomposition_tables(jpiglo, jpjglo, jperio, jpni, jpnj)
    ns = dict(t, jpiglo=jpiglo, jpni=jpni, jpnj=jpnj, narea=narea, njmpp=int(t["njmppt"][narea - 1]), nlci=nlci, nldi=nldi, nlei=nlei,
              isendto=np.zeros(3, np.int32), nsndto=0, nfsloop=0, nfeloop=0, l_north_nogather=False)
    f90exec.load(_read("src", "OCE", "LBC", "mppini.F90"), ns, arrays=("nimppt", "njmppt", "nlcit", "nfiimpp", "nfilcit", "nfipproc", "isendto"),
                 only=("mpp_init_nfdcom",), defined=("key_mpp_mpi",), module_vars=("nsndto", "nfsloop", "nfeloop", "l_north_nogather"))
    ns["mpp_init_nfdcom"]()
    return ns["nsndto"], [int(x) for x in ns["isendto"][:ns["nsndto"]]], ns["nfsloop"], ns["nfeloop"]


# ---- the multi-rank exchange: mpp_lnk / mpp_nfd of the reference, one thread per MPI rank --------------------------------------------
import queue
import threading


class _Mpi:
    """two-sided messages between rank threads: a FIFO per (source, destination, tag); MPI_ALLGATHER over the northern ranks"""

    def __init__(self):
        self.lock = threading.Lock()
        self.box = {}
        self.gather = {}

    def _q(self, key):
        with self.lock:
            return self.box.setdefault(key, queue.Queue())

    def send(self, src, dst, tag, data):
        self._q((src, dst, tag)).put(np.array(data, copy=True))

    def recv(self, src, dst, tag):
        return self._q((src, dst, tag)).get(timeout=60)


class MppWorld:
    """jpni x jpnj ranks running the REFERENCE's mpp_lnk_3d (mpp_lnk_generic.h90) with its north fold (mpp_nfd_generic.h90 ->
    lbc_nfd_generic.h90 on the gathered rows, or lbc_nfd_nogather_generic.h90), each in its own thread and namespace; mppsend / mpprecv /
    MPI_ALLGATHER are emulated.  The decomposition scalars of every rank (what mpp_init leaves in dom_oce / lib_mpp) are taken from the
    oracle's mpp_init (`doms`: sizes and fold partners are pinned to the reference separately, see mpp_basic_decomposition and
    mpp_init_nfdcom above); nothing of the oracle's exchange code is involved."""

    def __init__(self, doms, ln_nnogather):
        self.doms, self.n = doms, len(doms)
        d0 = doms[0]
        self.mpi = _Mpi()
        top = [r for r, d in enumerate(doms) if d.njmpp == max(x.njmpp for x in doms)]
        tabs = dict(nimppt=np.array([d.nimpp for d in doms], np.int32), nlcit=np.array([d.nlci for d in doms], np.int32),
                    nldit=np.array([d.nldi for d in doms], np.int32), nleit=np.array([d.nlei for d in doms], np.int32))
        nfipproc = np.asfortranarray(np.arange(self.n, dtype=np.int32).reshape(d0.jpnj, d0.jpni).T)
        nfiimpp = np.asfortranarray(tabs["nimppt"].reshape(d0.jpnj, d0.jpni).T)
        lnk = f90exec.cpp(_read("src", "OCE", "LBC", "mpp_lnk_generic.h90"), defined=("DIM_3d",), macros={"ROUTINE_LNK": "mpp_lnk_3d"})
        nfd = f90exec.cpp(_read("src", "OCE", "LBC", "mpp_nfd_generic.h90"), defined=("DIM_3d",), macros={"ROUTINE_NFD": "mpp_nfd_3d"})
        fold = {nd: f90exec.cpp(_read("src", "OCE", "LBC", "lbc_nfd_generic.h90"), defined=("DIM_%dd" % nd,), macros={"ROUTINE_NFD": "lbc_nfd_%dd" % nd})
                for nd in (3, 4)}
        nog = f90exec.cpp(_read("src", "OCE", "LBC", "lbc_nfd_nogather_generic.h90"), defined=("DIM_3d",), macros={"ROUTINE_NFD": "lbc_nfd_nogather_3d"})
        self.ns = []
        for r, d in enumerate(doms):
            ns = dict(jpi=d.jpi, jpj=d.jpj, jpim1=d.jpi - 1, jpjm1=d.jpj - 1, jpiglo=d.jpiglo, jpjglo=d.jpjglo, jpni=d.jpni, jpnj=d.jpnj,
                      jpnij=d.jpnij, jpimax=d.jpimax, jpjmax=d.jpjmax, jperio=d.jperio, narea=d.narea, nproc=d.nproc, nimpp=d.nimpp,
                      njmpp=d.njmpp, nlci=d.nlci, nlcj=d.nlcj, nldi=d.nldi, nlei=d.nlei, nldj=d.nldj, nlej=d.nlej, nbondi=d.nbondi,
                      nbondj=d.nbondj, noea=d.noea, nowe=d.nowe, noso=d.noso, nono=d.nono, npolj=d.npolj, l_iperio=bool(d.l_Iperio),
                      l_jperio=bool(d.l_Jperio), nsndto=d.nsndto, isendto=np.array(list(d.isendto) + [0, 0, 0], np.int32)[:3],
                      nfsloop=d.nfsloop, nfeloop=d.nfeloop, nn_hls=1, nreci=2, nrecj=2, numcom=0, ln_timing=False, l_isend=False,
                      jpmaxngh=3, mpi_status_size=1, mpi_double_precision=0, l_north_nogather=bool(ln_nnogather), ncom_stp=1, nit000=1,
                      ln_rstart=False, l_full_nf_update=True, ndim_rank_north=len(top), nrank_north=np.array(top, np.int32), ncomm_north=0,
                      nfipproc=nfipproc, nfiimpp=nfiimpp, **tabs)
            me = r

            def mppsend(ktyp, pmess, kbytes, kdest, md_req=None, me=me):
                self.mpi.send(me, int(kdest), int(ktyp), np.asarray(pmess).reshape(-1, order="F")[:int(kbytes)])

            def mpprecv(ktyp, pmess, kbytes, ksource, me=me):
                data = self.mpi.recv(int(ksource), me, int(ktyp))
                assert data.size == int(kbytes)
                flat = pmess if pmess.ndim == 1 else pmess.reshape(-1, order="F")
                assert np.shares_memory(flat, pmess)
                flat[:int(kbytes)] = data

            def mpi_allgather(sbuf, scount, stype, rbuf, rcount, rtype, comm, ierr, me=me, top=top):
                for q in top:                                          # every northern rank sends its rows to every northern rank
                    self.mpi.send(me, q, 99, np.asarray(sbuf).reshape(-1, order="F")[:int(scount)])
                flat = rbuf.reshape(-1, order="F")
                assert np.shares_memory(flat, rbuf)
                for n, q in enumerate(top):
                    flat[n * int(rcount):(n + 1) * int(rcount)] = self.mpi.recv(q, me, 99)

            ns.update(mppsend=mppsend, mpprecv=mpprecv, mpi_allgather=mpi_allgather, mpi_wait=lambda *a: None, tic_tac=lambda *a: None,
                      mpp_report=lambda *a, **k: None)
            arrs = ("nimppt", "nlcit", "nldit", "nleit", "nfipproc", "nfiimpp", "isendto", "nrank_north")
            for nd in (3, 4):
                f90exec.load(fold[nd], ns, int_arrays=arrs)
            f90exec.load(nog, ns, int_arrays=arrs)
            f90exec.load(nfd, ns, int_arrays=arrs, module_vars=("l_full_nf_update",))
            f90exec.load(lnk, ns, int_arrays=arrs)
            ns["lbc_nfd"] = lambda ptab, nat, sgn, ns=ns: ns["lbc_nfd_%dd" % ptab.ndim](ptab, nat, sgn)      # INTERFACE lbc_nfd
            ns["lbc_nfd_nogather"] = lambda ptab, ptab2, nat, sgn, ns=ns: ns["lbc_nfd_nogather_3d"](ptab, ptab2, nat, sgn)
            ns["mpp_nfd"] = ns["mpp_nfd_3d"]
            self.ns.append(ns)

    def lbc_lnk(self, locs, nat, sgn):
        """mpp_lnk_3d on every rank at once; locs: list over ranks of C-ordered [jpk][jpj][jpi] arrays, updated in place"""
        err = []

        def run(r):
            try:
                self.ns[r]["mpp_lnk_3d"]("ref_exec", np.transpose(locs[r]), nat, float(sgn))
            except Exception as e:        # noqa: BLE001
                import traceback
                err.append((r, traceback.format_exc()))
        th = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(self.n)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
        if err or any(t.is_alive() for t in th):
            raise RuntimeError("mpp_lnk failed or hung: %s" % (err[:1] or "timeout"))

    def rank_lbc(self, r):
        """the exchange as rank r's tracer routine calls it (to be used from that rank's own thread, all ranks running)"""
        def lbc(items):
            for a, nat, sgn in items:
                a3 = a.reshape((-1,) + a.shape[-2:])
                assert np.shares_memory(a3, a)
                self.ns[r]["mpp_lnk_3d"]("ref_exec", np.transpose(a3), nat, float(sgn))
        return lbc


def tra_adv_fct_mpp(world, gf, kjpt, kn_fct_h, kn_fct_v, ln_linssh, ln_isfcav, ln_nnogather=True):
    """the reference's tra_adv_fct on EVERY rank of `world` (an oracle World, used for the decomposition scalars and to scatter the
    global fields of `gf`), the ranks exchanging their halos through the reference's mpp_lnk / mpp_nfd (MppWorld): what `mpirun -np
    jpnij nemo` does for this routine.  Returns the list of local pta arrays."""
    mw = MppWorld(world.doms, ln_nnogather)
    loc = {k: world.scatter(gf[k]) for k in DOM_ARRAYS + DOM_INT_ARRAYS + ("pun", "pvn", "pwn", "ptb", "ptn", "pta")}
    out, err = [None] * mw.n, []

    def run(r):
        try:
            d = world.doms[r]
            g = {k: loc[k][r] for k in loc}
            g["p2dt"] = gf["p2dt"]
            out[r] = tra_adv_fct(g, d.jpi, d.jpj, d.jpk, kjpt, kn_fct_h, kn_fct_v, ln_linssh, ln_isfcav, mw.rank_lbc(r))
        except Exception:                 # noqa: BLE001
            import traceback
            err.append((r, traceback.format_exc()))
    th = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(mw.n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=240)
    if err or any(t.is_alive() for t in th):
        raise RuntimeError("tra_adv_fct_mpp failed or hung: %s" % (err[:1] or "timeout"))
    return out


def dom_msk(doms, k_top, k_bot, ln_nnogather=True):
    """the reference's dom_msk (src/OCE/DOM/dommsk.F90:66-290, free-slip: rn_shlat = 0, no BDY mask file) on every rank of a layout,
    from its text: tmask from k_top / k_bot, lbc_lnk (mono-processor reference lbc_lnk, or the reference's mpp_lnk on emulated ranks),
    umask / vmask / fmask, wmask, ssmask, tmask_h with the north-fold half row, tmask_i.  k_top, k_bot: lists over ranks of int32
    (jpj, jpi).  Returns a list over ranks of dict(tmask, umask, vmask, wmask, tmask_i) as C-ordered arrays."""
    n = len(doms)
    mw = MppWorld(doms, ln_nnogather) if n > 1 else None
    res, err = [None] * n, []

    def run(r):
        try:
            d = doms[r]
            jpi, jpj, jpk = d.jpi, d.jpj, d.jpk
            lbc = mw.rank_lbc(r) if mw else reference_lbc(d.jperio, jpi, jpj)
            a3 = lambda: np.full((jpi, jpj, jpk), np.nan, order="F")       # noqa: E731  (module arrays: ALLOCATEd, undefined)
            a2 = lambda: np.full((jpi, jpj), np.nan, order="F")            # noqa: E731
            ns = dict(jpi=jpi, jpj=jpj, jpk=jpk, jpim1=jpi - 1, jpjm1=jpj - 1, jpkm1=jpk - 1, jpiglo=d.jpiglo, jpjglo=d.jpjglo, jperio=d.jperio,
                      nlci=d.nlci, nlcj=d.nlcj, nlej=d.nlej, nn_hls=1, lwp=False, lwm=False, numout=6, numond=7, numnam_ref=1, numnam_cfg=2,
                      rn_shlat=0.0, ln_vorlat=False, ln_bdy=False, ln_mask_file=False, ios=0,
                      mig=np.arange(1, jpi + 1, dtype=np.int32) + d.nimpp - 1, mjg=np.arange(1, jpj + 1, dtype=np.int32) + d.njmpp - 1,
                      tmask=a3(), umask=a3(), vmask=a3(), fmask=a3(), wmask=a3(), wumask=a3(), wvmask=a3(), ssmask=a2(), ssumask=a2(),
                      ssvmask=a2(), tmask_h=a2(), tmask_i=a2(), tpol=np.full(d.jpiglo, np.nan), fpol=np.full(d.jpiglo, np.nan),
                      ctl_nam=lambda *a: None, usr_def_fmask=lambda *a: None, cn_cfg="none", nn_cfg=0)

            def lbc_lnk(cdname, a, nat, sgn):
                lbc([(np.transpose(a), nat, float(sgn))])

            def lbc_lnk_multi(cdname, *t):
                lbc([(np.transpose(t[i]), t[i + 1], float(t[i + 2])) for i in range(0, len(t), 3)])
            ns.update(lbc_lnk=lbc_lnk, lbc_lnk_multi=lbc_lnk_multi)
            defines = f90exec.cpp_defines(_read("src", "OCE", "vectopt_loop_substitute.h90"))
            f90exec.load(_read("src", "OCE", "DOM", "dommsk.F90"), ns, defines=defines, only=("dom_msk",), int_arrays=("mig", "mjg"),
                         arrays=("tmask", "umask", "vmask", "fmask", "wmask", "wumask", "wvmask", "ssmask", "ssumask", "ssvmask", "tmask_h",
                                 "tmask_i", "tpol", "fpol", "bdytmask"))
            ns["dom_msk"](np.transpose(np.ascontiguousarray(k_top[r], np.int32)), np.transpose(np.ascontiguousarray(k_bot[r], np.int32)))
            res[r] = {k: np.ascontiguousarray(np.transpose(ns[k])) for k in ("tmask", "umask", "vmask", "wmask", "tmask_i")}
        except Exception:                 # noqa: BLE001
            import traceback
            err.append((r, traceback.format_exc()))
    th = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    if err or any(t.is_alive() for t in th):
        raise RuntimeError("dom_msk failed or hung: %s" % (err[:1] or "timeout"))
    return res


MPP_INIT_SCALARS = ("jpi", "jpj", "jpk", "jpim1", "jpjm1", "jpkm1", "jpij", "jpnij", "jpimax", "jpjmax", "nreci", "nrecj", "l_iperio", "l_jperio",
                    "noso", "nowe", "noea", "nono", "nlci", "nldi", "nlei", "nlcj", "nldj", "nlej", "nbondi", "nbondj", "nimpp", "njmpp", "npolj",
                    "nproc", "nsndto", "nfsloop", "nfeloop", "l_north_nogather", "ndim_rank_north")
MPP_INIT_TABLES = ("nfiimpp", "nfipproc", "nfilcit", "nimppt", "ibonit", "nlcit", "nlcjt", "njmppt", "ibonjt", "nldit", "nldjt", "nleit", "nlejt",
                   "isendto", "nrank_north")


def mpp_init(jpiglo, jpjglo, jpkglo, jperio, jpni, jpnj, narea, ln_nnogather=True):
    """the reference's mpp_init (src/OCE/LBC/mppini.F90:110-692) for rank `narea` (1-based) of an all-ocean jpni x jpnj layout, from
    its text (with ITS mpp_basic_decomposition and mpp_init_nfdcom): the decomposition scalars it leaves in dom_oce / lib_mpp -- sizes,
    offsets, inner bounds, neighbour ranks, boundary flags, north-fold type, fold partners -- as a dict.  Stubbed: the search for the
    best partition and for land subdomains (every subdomain holds ocean), the north communicator, IOIPSL, all printing."""
    import types
    def stop(*a):
        raise ValueError("ctl_stop: " + " ".join(str(x) for x in a))
    ns = dict(jpiglo=jpiglo, jpjglo=jpjglo, jpkglo=jpkglo, jperio=jperio, jpni=jpni, jpnj=jpnj, mppsize=jpni * jpnj, narea=narea, nn_hls=1,
              lwp=False, ln_ctl=False, sn_cfctl=types.SimpleNamespace(l_layout=False), ln_read_cfg=False, ln_bdy=False, ln_mask_file=False,
              ln_nnogather=bool(ln_nnogather), numnam_ref=1, numnam_cfg=2, numbot=-1, numbdy=-1, numout=6, ctmp1="", ctmp2="", ctmp3="", ctmp4="",
              ctl_stop=stop, ctl_warn=lambda *a: None, ctl_nam=lambda *a: None, mpp_sum=lambda *a: None, mpp_init_ioipsl=lambda: None,
              isendto=np.zeros(3, np.int32), ndim_rank_north=0, nrank_north=np.zeros(jpni, np.int32))
    for k in MPP_INIT_SCALARS:
        ns.setdefault(k, 0)
    ns["mpp_init_bestpartition"] = lambda knbij, knbi=None, knbj=None, knbcnt=None, ldlist=None: {1: jpni, 2: jpnj, 3: 0}
    ns["mpp_init_isoce"] = lambda knbi, knbj, ldisoce: ldisoce.__setitem__(Ellipsis, 1)

    def mpp_ini_north():                                               # lib_mpp.F90: the ranks of the northern row, all ocean here
        ns["ndim_rank_north"] = jpni
        ns["nrank_north"] = np.arange((jpnj - 1) * jpni, jpnj * jpni, dtype=np.int32)
    ns["mpp_ini_north"] = mpp_ini_north
    f90exec.load(_read("src", "OCE", "LBC", "mppini.F90"), ns, int_arrays=MPP_INIT_TABLES, defined=("key_mpp_mpi",),
                 only=("mpp_init", "mpp_basic_decomposition", "mpp_init_nfdcom"), module_vars=MPP_INIT_SCALARS + MPP_INIT_TABLES)
    ns["mpp_init"]()
    out = {k: ns[k] for k in MPP_INIT_SCALARS}
    out["isendto"] = [int(x) for x in ns["isendto"][:ns["nsndto"]]]
    return out


def trc_nxt(gf, trb, trn, tra, sbc_trc, sbc_trc_b, extra, jpi, jpj, jpk, kt, nittrc000, neuler, rdttrc, atfp, r1_rau0, ln_linssh, lbc,
            ln_top_euler=False):
    """the reference's trc_nxt driver (src/TOP/TRP/trcnxt.F90:56-183) from its text: lbc_lnk on tra, then the Euler swap (trn = tra,
    trb = trn) or tra_nxt_fix / tra_nxt_vvl of tranxt.F90 with cdtype = 'TRC' (no solar, runoff or ice-shelf terms) + lbc_lnk on
    trb, trn, tra.  Arrays C-ordered [jptra][jpk][jpj][jpi]; returns (trb, trn, tra)."""
    dom = RefDomain(gf, jpi, jpj, jpk, ln_linssh, False, lbc)
    ns = dom.ns
    t = {"trb": np.array(trb, copy=True), "trn": np.array(trn, copy=True), "tra": np.array(tra, copy=True)}
    for k, a in t.items():
        ns[k] = F(a)
    ns.update(jptra=t["tra"].shape[0], jp_tem=1, jp_sal=2, nittrc000=nittrc000, neuler=neuler, rdttrc=float(rdttrc), r2dttrc=2.0 * float(rdttrc),
              atfp=float(atfp), r1_rau0=float(r1_rau0), ln_timing=False, ln_bdy=False, ln_traldf_iso=False, ln_ctl=False, l_offline=False,
              ln_top_euler=bool(ln_top_euler), ln_traqsr=False, ln_rnf=False, ln_isf=False, ln_rnf_depth=False, nksr=0,
              sbc_trc=F(np.ascontiguousarray(sbc_trc)), sbc_trc_b=F(np.ascontiguousarray(sbc_trc_b)))
    for k in ("emp_b", "emp", "fwfisf_b", "fwfisf", "rnf_b", "rnf"):
        ns[k] = F(np.ascontiguousarray(extra[k]))
    arrays = DOM_ARRAYS + NXT_ARRAYS + ("trb", "trn", "tra", "sbc_trc", "sbc_trc_b")
    f90exec.load(_read("src", "OCE", "TRA", "tranxt.F90"), ns, arrays=arrays, int_arrays=DOM_INT_ARRAYS + NXT_INT_ARRAYS, defines=dom.defines,
                 only=("tra_nxt_fix", "tra_nxt_vvl"))
    f90exec.load(_read("src", "TOP", "TRP", "trcnxt.F90"), ns, arrays=arrays, int_arrays=DOM_INT_ARRAYS + NXT_INT_ARRAYS, defines=dom.defines,
                 only=("trc_nxt",), defined=("key_top",))
    ns["trc_nxt"](kt)
    return t["trb"], t["trn"], t["tra"]


def stp_ctl(dom, sshn, un, tsn, tmask, umask, kt=1):
    """the reference's stp_ctl (src/OCE/stpctl.F90:58-196) on one subdomain from its text, in the branch without ln_ctl: the extrema
    zmax(1:6), the error condition (kindic = -3 and ctl_stop) and, when it fires, the MAXLOC / MINLOC positions.  The files it
    would write (time.step, run.stat, output.abort) are not opened (lwm = .false.).  dom: the decomposition scalars (nimpp, njmpp,
    narea); arrays C-ordered.  Returns dict(zmax[6], kindic, and ih, iu, is1, is2 when kindic = -3)."""
    import types
    stopped = []
    ssmask = np.ascontiguousarray(np.max(tmask, axis=0))               # dommsk.F90: ssmask = MAXVAL( tmask, DIM=3 )
    ns = dict(sshn=F(np.ascontiguousarray(sshn)), un=F(np.ascontiguousarray(un)), tsn=F(np.ascontiguousarray(tsn)), tmask=F(np.ascontiguousarray(tmask)),
              umask=F(np.ascontiguousarray(umask)), ssmask=F(ssmask), jp_tem=1, jp_sal=2, nit000=kt, nitend=kt + 10, lwp=False, lwm=False,
              ln_ctl=False, lk_mpp=True, ll_wd=False, ln_zad_aimp=False, nstop=0, numout=6, numstp=7, numrun=8, narea=dom.narea,
              nimpp=dom.nimpp, njmpp=dom.njmpp, sn_cfctl=types.SimpleNamespace(ptimincr=1, l_runstat=False),
              ctl_stop=lambda *a: stopped.append(a), dia_wri_state=lambda *a: None, **{"ctmp%d" % i: "" for i in range(1, 11)})
    f90exec.load(_read("src", "OCE", "stpctl.F90"), ns, arrays=("sshn", "un", "tsn", "tmask", "umask", "ssmask", "wi", "cu_adv", "wmask"),
                 only=("stp_ctl",), module_vars=("zmax", "ih", "iu", "is1", "is2", "nstop"))
    out = ns["stp_ctl"](kt, 0)
    res = dict(zmax=[float(x) for x in ns["zmax"][:6]], kindic=int(out["kindic"]), stopped=bool(stopped))
    if res["kindic"]:
        res.update(ih=[int(x) for x in ns["ih"]], iu=[int(x) for x in ns["iu"]], is1=[int(x) for x in ns["is1"]], is2=[int(x) for x in ns["is2"]])
    return res


def tra_adv_mus_mpp(world, gf, mx, kjpt, ln_linssh, ln_isfcav, ld_msc_ups, ln_nnogather=True):
    """the reference's tra_adv_mus on EVERY rank of `world`, halos through the reference's mpp_lnk / mpp_nfd on emulated MPI ranks
    (as tra_adv_fct_mpp).  Returns the list of local pta arrays."""
    mw = MppWorld(world.doms, ln_nnogather)
    loc = {k: world.scatter(gf[k]) for k in DOM_ARRAYS + DOM_INT_ARRAYS + ("pun", "pvn", "pwn", "ptb", "pta")}
    lx = {k: world.scatter(mx[k]) for k in ("r1_e1e2u", "r1_e1e2v", "e3u_n", "e3v_n", "e3w_n") + (("rnfmsk",) if ld_msc_ups else ())}
    out, err = [None] * mw.n, []

    def run(r):
        try:
            d = world.doms[r]
            g = {k: loc[k][r] for k in loc}
            g["p2dt"] = gf["p2dt"]
            m = {k: lx[k][r] for k in lx}
            if ld_msc_ups:
                m["rnfmsk_z"] = mx["rnfmsk_z"]
            out[r] = tra_adv_mus(g, m, d.jpi, d.jpj, d.jpk, kjpt, ln_linssh, ln_isfcav, ld_msc_ups, mw.rank_lbc(r))
        except Exception:                 # noqa: BLE001
            import traceback
            err.append((r, traceback.format_exc()))
    th = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(mw.n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=240)
    if err or any(t.is_alive() for t in th):
        raise RuntimeError("tra_adv_mus_mpp failed or hung: %s" % (err[:1] or "timeout"))
    return out
