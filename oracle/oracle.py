"""ctypes front-end of the CPU ORACLE (test infrastructure only -- see oracle/nemo_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The oracle is a loop-for-loop C restatement of the reference's tra_adv_fct / nonosc / interp_4th_cpt
(src/OCE/TRA/traadv_fct.F90), lbc_lnk / mpp_lnk / lbc_nfd / mpp_nfd (src/OCE/LBC/*.h90) and mpp_init
(src/OCE/LBC/mppini.F90).  PARITY PIN: see nemo_oracle.h (the reference's source text executed by translation, oracle/f90exec.py + ref_exec.py).

Arrays are numpy float64 in Fortran layout seen from C order: shape (jpk, jpj, jpi) [or (kjpt, jpk, jpj, jpi)],
i.e. ji fastest, exactly the memory image of REAL(wp) a(jpi,jpj,jpk[,kjpt]).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INFO_FIELDS = ("jpi jpj jpk jpiglo jpjglo jpni jpnj jpnij jpimax jpjmax jperio narea nproc nimpp njmpp "
               "nlci nlcj nldi nlei nldj nlej nbondi nbondj noea nowe noso nono npolj l_Iperio l_Jperio "
               "nsndto isendto1 isendto2 isendto3 nfsloop nfeloop").split()

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.mpp_init.restype = C.c_void_p
        L.mpp_init.argtypes = [C.c_int] * 8
        L.mpp_finalize.argtypes = [C.c_void_p]
        L.oce_world_dom.restype = C.c_void_p
        L.oce_world_dom.argtypes = [C.c_void_p, C.c_int]
        L.oce_dom_info.argtypes = [C.c_void_p, IP]
        L.oce_dom_set_fields.argtypes = [C.c_void_p] + [C.c_void_p] * 11 + [C.c_int, C.c_int]
        L.oce_dom_set_dbg.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.oce_world_tra_adv_fct.argtypes = [C.c_void_p, C.c_double] + [C.POINTER(C.c_void_p)] * 6 + [C.c_int] * 3
        L.oce_world_lbc_lnk.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_void_p)), C.c_char_p, DP, C.c_int]
        L.oce_world_dom_msk.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 9
        L.interp_4th_cpt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mpp_basic_decomposition.argtypes = [C.c_int] * 5 + [IP] * 6
        L.oracle_poison_workspace.argtypes = [C.c_int]
        L.glob_sum_local.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, DP]
        L.ddpdd.argtypes = [DP, DP]
        L.stp_ctl_local.argtypes = [C.c_void_p] * 5 + [DP, IP, IP, IP, IP, IP, IP]
        L.sign_nosignedzero.restype = C.c_double
        L.sign_nosignedzero.argtypes = [C.c_double, C.c_double]
        L.tra_adv_transports.argtypes = [C.c_void_p] * 11
        L.oce_dom_set_mus_fields.argtypes = [C.c_void_p] * 6
        L.oce_dom_set_diag.argtypes = [C.c_void_p] * 4
        L.oce_world_tra_adv_cen.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 5 + [C.c_int] * 3
        L.tra_adv_mus_xind.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oce_world_tra_adv_mus.argtypes = [C.c_void_p, C.c_double] + [C.POINTER(C.c_void_p)] * 5 + [C.c_int, C.POINTER(C.c_void_p)]
        L.oce_world_tra_nxt.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_char_p]
                                        + [C.POINTER(C.c_void_p)] * 6 + [C.c_int])
        _LIB = L
    return _LIB


def _ptr(a):
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _ptr_table(arrs):
    t = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        t[i] = None if a is None else a.ctypes.data
    return t


class NxtForcing(C.Structure):
    """oce_nxt_forcing: module variables read by tra_nxt_vvl (tranxt.F90:262-343); arrays kept alive in _keep."""
    _fields_ = ([("atfp", C.c_double), ("r1_rau0", C.c_double)]
                + [(n, C.c_int) for n in "ln_traqsr ln_rnf ln_isf ln_rnf_depth nksr".split()]
                + [(n, C.c_void_p) for n in ("emp_b emp fwfisf_b fwfisf rnf_b rnf qsr_hc qsr_hc_b nk_rnf h_rnf rnf_tsc "
                                             "rnf_tsc_b misfkt misfkb risf_tsc risf_tsc_b r1_hisf_tbl ralpha").split()])

    def __init__(self, atfp=0.1, r1_rau0=1.0 / 1026.0, **kw):
        super().__init__()
        self.atfp, self.r1_rau0 = atfp, r1_rau0
        self._keep = {}
        for k, v in kw.items():
            if v is None:
                continue
            if isinstance(v, np.ndarray):
                self._keep[k] = v
                setattr(self, k, v.ctypes.data)
            else:
                setattr(self, k, int(v))


class Dom:
    """One subdomain: decomposition scalars (attributes named as in dom_oce.F90) + bound module arrays."""

    def __init__(self, handle):
        self.h = handle
        out = (C.c_int * len(INFO_FIELDS))()
        lib().oce_dom_info(handle, out)
        for k, v in zip(INFO_FIELDS, out):
            setattr(self, k, int(v))
        self.isendto = [self.isendto1, self.isendto2, self.isendto3][: self.nsndto]
        self._keep = {}

    @property
    def shape3(self):
        return (self.jpk, self.jpj, self.jpi)

    def set_fields(self, tmask, umask, vmask, wmask, e3t_b, e3t_n, e3t_a, e1e2t, r1_e1e2t, mikt, mbkt,
                   ln_linssh=False, ln_isfcav=False):
        arrs = [tmask, umask, vmask, wmask, e3t_b, e3t_n, e3t_a, e1e2t, r1_e1e2t, mikt, mbkt]
        assert mikt.dtype == np.int32 and mbkt.dtype == np.int32
        self._keep["fields"] = arrs
        lib().oce_dom_set_fields(self.h, *[_ptr(a) for a in arrs], int(ln_linssh), int(ln_isfcav))

    def set_diag(self, trdx=None, trdy=None, trdz=None):
        """l_trd / l_hst / l_ptr hooks of tra_adv_fct: (kjpt,jpk,jpj,jpi) arrays receiving ztrdx, ztrdy (= zptry), ztrdz"""
        self._keep["diag"] = (trdx, trdy, trdz)
        lib().oce_dom_set_diag(self.h, *[None if a is None else _ptr(a) for a in (trdx, trdy, trdz)])

    def set_mus_fields(self, r1_e1e2u, r1_e1e2v, e3u_n, e3v_n, e3w_n):
        arrs = [r1_e1e2u, r1_e1e2v, e3u_n, e3v_n, e3w_n]
        self._keep["mus"] = arrs
        lib().oce_dom_set_mus_fields(self.h, *[_ptr(a) for a in arrs])

    def mus_xind(self, ld_msc_ups=False, rnfmsk=None, rnfmsk_z=None):
        """xind of traadv_mus.F90:99-113 (needs tmask bound when ld_msc_ups)."""
        xind = np.empty(self.shape3)
        lib().tra_adv_mus_xind(self.h, int(ld_msc_ups), None if rnfmsk is None else _ptr(rnfmsk),
                               None if rnfmsk_z is None else _ptr(rnfmsk_z), _ptr(xind))
        return xind

    def set_dbg(self, jn, **arrs):
        names = "zwi zwx zwy zwz zbetup zbetdo paa pbb pcc ztw zltu zltv".split()
        tab = _ptr_table([arrs.get(n) for n in names])
        self._keep["dbg"] = arrs
        lib().oce_dom_set_dbg(self.h, jn, tab)


class World:
    """mpp_init(...) result: the set of subdomains of one run (threads stand in for MPI ranks)."""

    def __init__(self, jpiglo, jpjglo, jpk, jperio, jpni=1, jpnj=1, ln_nnogather=True, key_mpp_mpi=True):
        self.h = lib().mpp_init(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, int(ln_nnogather), int(key_mpp_mpi))
        if not self.h:
            raise ValueError("mpp_init: layout not runnable by the reference (more than jpmaxngh no-gather fold partners)")
        self.jpiglo, self.jpjglo, self.jpk, self.jperio = jpiglo, jpjglo, jpk, jperio
        self.jpni, self.jpnj, self.jpnij = jpni, jpnj, jpni * jpnj
        self.doms = [Dom(lib().oce_world_dom(self.h, r)) for r in range(self.jpnij)]

    def close(self):
        if self.h:
            lib().mpp_finalize(self.h)
            self.h = None

    # ---- global <-> local (chap_LBC.tex:311-318: T_g(i+nimpp-1, j+njmpp-1, k) = T_l(i,j,k)) ----
    def scatter(self, glob):
        """global (..., jpjglo, jpiglo) -> list of contiguous local copies (..., jpj, jpi), halos included."""
        out = []
        for d in self.doms:
            j0, i0 = d.njmpp - 1, d.nimpp - 1
            out.append(np.array(glob[..., j0:j0 + d.jpj, i0:i0 + d.jpi], order="C", copy=True))
        return out

    def gather(self, locs, glob=None):
        """assemble the global array from local INTERIORS (nldj:nlej, nldi:nlei)."""
        if glob is None:
            glob = np.full(locs[0].shape[:-2] + (self.jpjglo, self.jpiglo), np.nan, dtype=locs[0].dtype)
        for d, a in zip(self.doms, locs):
            j0, i0 = d.njmpp - 1, d.nimpp - 1
            glob[..., j0 + d.nldj - 1:j0 + d.nlej, i0 + d.nldi - 1:i0 + d.nlei] = a[..., d.nldj - 1:d.nlej, d.nldi - 1:d.nlei]
        return glob

    # ---- collective calls ("mpirun -np jpnij") ----
    def lbc_lnk(self, fields, cd_nat, psgn):
        """fields: list over fields of list over ranks of arrays (ipk, jpj, jpi) (or (jpj,jpi) for 2D)."""
        nfld = len(fields)
        ipk = 1 if fields[0][0].ndim == 2 else int(np.prod(fields[0][0].shape[:-2]))
        per_rank = [_ptr_table([fields[f][r] for f in range(nfld)]) for r in range(self.jpnij)]
        tabs = (C.POINTER(C.c_void_p) * self.jpnij)(*[C.cast(t, C.POINTER(C.c_void_p)) for t in per_rank])
        sg = (C.c_double * nfld)(*psgn)
        lib().oce_world_lbc_lnk(self.h, nfld, tabs, cd_nat.encode(), sg, ipk)

    def tra_adv_fct(self, p2dt, pun, pvn, pwn, ptb, ptn, pta, kjpt, kn_fct_h, kn_fct_v):
        """each argument: list over ranks of arrays; pta is updated in place."""
        L = lib()
        L.oce_world_tra_adv_fct(self.h, p2dt, _ptr_table(pun), _ptr_table(pvn), _ptr_table(pwn), _ptr_table(ptb),
                                _ptr_table(ptn), _ptr_table(pta), kjpt, kn_fct_h, kn_fct_v)

    def tra_adv_cen(self, pun, pvn, pwn, ptn, pta, kjpt, kn_cen_h, kn_cen_v):
        """tra_adv_cen on every subdomain; pta updated in place."""
        lib().oce_world_tra_adv_cen(self.h, _ptr_table(pun), _ptr_table(pvn), _ptr_table(pwn), _ptr_table(ptn),
                                    _ptr_table(pta), kjpt, kn_cen_h, kn_cen_v)

    def tra_adv_mus(self, p2dt, pun, pvn, pwn, ptb, pta, kjpt, xind):
        """tra_adv_mus on every subdomain; pta updated in place; xind: list over ranks."""
        lib().oce_world_tra_adv_mus(self.h, p2dt, _ptr_table(pun), _ptr_table(pvn), _ptr_table(pwn), _ptr_table(ptb),
                                    _ptr_table(pta), kjpt, _ptr_table(xind))

    def tra_nxt(self, kt, kit000, l_euler, rdt, cdtype, forcing, ptb, ptn, pta, kjpt, sbc=None, sbc_b=None):
        """tra_nxt ('TRA') / trc_nxt ('TRC') on every subdomain; forcing: list over ranks of NxtForcing."""
        ft = (C.c_void_p * self.jpnij)(*[C.addressof(f) for f in forcing])
        lib().oce_world_tra_nxt(self.h, kt, kit000, int(l_euler), rdt, cdtype.encode(), ft, _ptr_table(ptb),
                                _ptr_table(ptn), _ptr_table(pta), None if sbc is None else _ptr_table(sbc),
                                None if sbc_b is None else _ptr_table(sbc_b), kjpt)

    def dom_msk(self, k_top, k_bot):
        """returns per-rank dict(tmask, umask, vmask, wmask, tmask_i, mikt, mbkt)."""
        res = []
        for d in self.doms:
            res.append(dict(tmask=np.zeros(d.shape3), umask=np.zeros(d.shape3), vmask=np.zeros(d.shape3),
                            wmask=np.zeros(d.shape3), tmask_i=np.zeros((d.jpj, d.jpi)),
                            mikt=np.zeros((d.jpj, d.jpi), np.int32), mbkt=np.zeros((d.jpj, d.jpi), np.int32)))
        tabs = [_ptr_table(k_top), _ptr_table(k_bot)] + [
            _ptr_table([r[n] for r in res]) for n in ("tmask", "umask", "vmask", "wmask", "tmask_i", "mikt", "mbkt")]
        lib().oce_world_dom_msk(self.h, *tabs)
        return res


def stp_ctl(dom, sshn, un, tsn, tmask):
    """Extrema test of stp_ctl (stpctl.F90:115-124, 149-166, 184) on one subdomain.  tsn: (2, jpk, jpj, jpi) = (tem, sal).
    Returns dict(zmax[6], ih, iu, is1, is2, nan_found, kindic)."""
    zmax = (C.c_double * 6)()
    ih, iu, is1, is2 = (C.c_int * 2)(), (C.c_int * 3)(), (C.c_int * 3)(), (C.c_int * 3)()
    nanf, kindic = C.c_int(0), C.c_int(0)
    lib().stp_ctl_local(dom.h, _ptr(sshn), _ptr(un), _ptr(tsn), _ptr(tmask), zmax, ih, iu, is1, is2, C.byref(nanf), C.byref(kindic))
    return dict(zmax=list(zmax), ih=list(ih), iu=list(iu), is1=list(is1), is2=list(is2), nan_found=nanf.value, kindic=kindic.value)


def glob_sum(world, locs, tmask_i):
    """glob_sum (lib_fortran_generic.h90:32-65): per-rank double-double partial sums combined with DDPDD
    (the MPI_SUMDD reduction, lib_mpp.F90:1158-1186).  Returns the fp64 REAL part."""
    L = lib()
    tot = (C.c_double * 2)(0.0, 0.0)
    for d, a, m in zip(world.doms, locs, tmask_i):
        part = (C.c_double * 2)(0.0, 0.0)
        ipk = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
        L.glob_sum_local(_ptr(a), _ptr(m), d.jpi, d.jpj, ipk, part)
        L.ddpdd(part, tot)
    return tot[0], tot[1]
