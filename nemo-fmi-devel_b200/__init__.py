"""nemo-fmi-devel_b200 -- host-side mirror of the NEMO interfaces on the FCT tracer-advection path.

This package is a thin ctypes binding of ``libnemo_fct.so`` (C ABI in ``include/nemo_fct.h``), keeping the
reference's routine names, argument order and meaning so that tests read like calls into NEMO:

    mpp_init            src/OCE/LBC/mppini.F90:110           -> :func:`mpp_init`
    tra_adv_fct         src/OCE/TRA/traadv_fct.F90:54        -> :meth:`FctContext.tra_adv_fct`
    interp_4th_cpt      src/OCE/TRA/traadv_fct.F90:517       -> :meth:`FctContext.interp_4th_cpt`
    lbc_lnk_multi       src/OCE/LBC/lbc_lnk_multi_generic.h90:16 -> :meth:`FctContext.lbc_lnk_multi`
    tra_adv transports  src/OCE/TRA/traadv.F90:100-124       -> :meth:`FctContext.tra_adv_transports`
    tra_adv / trc_adv   src/OCE/TRA/traadv.F90:77, src/TOP/TRP/trcadv.F90:70 -> :meth:`FctContext.tra_adv`, :meth:`FctContext.trc_adv`
    tra_adv_mus         src/OCE/TRA/traadv_mus.F90:55        -> :meth:`FctContext.tra_adv_mus`
    tra_adv_cen         src/OCE/TRA/traadv_cen.F90:46        -> :meth:`FctContext.tra_adv_cen`
    tra_nxt / trc_nxt   src/OCE/TRA/tranxt.F90:65, src/TOP/TRP/trcnxt.F90:56 -> :meth:`FctContext.tra_nxt`
    l_trd/l_hst/l_ptr   src/OCE/TRA/traadv_fct.F90:96-112    -> :meth:`FctContext.set_trend_diag`

Arrays are fp64 with the Fortran memory image ``a(jpi,jpj,jpk[,kjpt])``, i.e. C-order shape ``([kjpt,] jpk, jpj, jpi)``.
numpy arrays are passed as HOST pointers (the library copies H2D/D2H, as a Fortran host would see it); CUDA
``torch`` tensors are passed as DEVICE pointers (device-resident path).  There is no CPU implementation here:
without the CUDA library the import of the binding fails loudly, and without a GPU every compute call raises.

(The directory name carries a hyphen as the project layout prescribes; import it with
``importlib.import_module("nemo-fmi-devel_b200")`` or through the ``nemo_fct_b200`` alias module at the repo root.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnemo_fct.so")
JPMAXNGH = 3
UNIQUE_ID_BYTES = 128


class NemoFctError(RuntimeError):
    """Raised for every non-zero return of the C ABI (the Fortran shim would CALL ctl_stop('STOP', msg))."""


class Domain(C.Structure):
    """struct nemo_fct_domain: the decomposition scalars of par_oce.F90 / dom_oce.F90 (mppini.F90:548-580)."""
    _fields_ = [(n, C.c_int) for n in (
        "jpiglo jpjglo jpk jperio jpni jpnj narea jpi jpj jpimax jpjmax nimpp njmpp nlci nlcj nldi nlei nldj nlej "
        "nbondi nbondj noea nowe noso nono npolj l_Iperio l_Jperio nsndto").split()] + [
        ("isendto", C.c_int * JPMAXNGH), ("key_mpp_mpi", C.c_int)]

    @property
    def shape3(self):
        return (self.jpk, self.jpj, self.jpi)

    @property
    def nproc(self):
        return self.narea - 1

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "isendto"}
        d["isendto"] = list(self.isendto)[: self.nsndto]
        return d


_lib = None


def build(verbose=False):
    """Compile libnemo_fct.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], stdout=out)
    return LIB_PATH


def lib():
    """The loaded C-ABI library.  Fails loudly if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  There is no CPU fallback for the FCT path.")
    L = C.CDLL(LIB_PATH)
    vp, ip, dp, i, d = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.c_double
    L.nemo_fct_last_error.restype = C.c_char_p
    L.nemo_fct_launch_count.restype = C.c_longlong
    sig = {
        "nemo_mpp_init": [i] * 8 + [C.POINTER(Domain)],
        "nemo_mpp_basic_decomposition": [i] * 5 + [ip] * 6,
        "nemo_lbc_plan_query": [C.POINTER(Domain), C.c_char, i, ip, ip, ip],
        "nemo_fct_create": [C.POINTER(Domain), i, C.POINTER(vp)],
        "nemo_fct_destroy": [vp],
        "nemo_fct_set_domain_arrays": [vp] + [vp] * 8 + [i, i],
        "nemo_fct_set_e3t": [vp, vp, vp, vp, i],
        "nemo_fct_set_stream": [vp, vp],
        "nemo_fct_synchronize": [vp],
        "nemo_fct_comm_unique_id": [vp],
        "nemo_fct_comm_init": [vp, vp, i, i],
        "nemo_fct_comm_init_local": [C.POINTER(vp), i],
        "nemo_tra_adv_fct": [vp, i, i, C.c_char_p, d] + [vp] * 6 + [i, i, i],
        "nemo_tra_adv_fct_dev": [vp, i, i, C.c_char_p, d] + [vp] * 6 + [i, i, i],
        "nemo_group_tra_adv_fct_dev": [C.POINTER(vp), i, i, i, C.c_char_p, d] + [C.POINTER(vp)] * 6 + [i, i, i],
        "nemo_interp_4th_cpt": [vp, vp, vp],
        "nemo_interp_4th_cpt_dev": [vp, vp, vp],
        "nemo_tra_adv_transports_dev": [vp] * 11,
        "nemo_tra_adv_dev": [vp, i, i, i, d] + [vp] * 10 + [i, i, i],
        "nemo_trc_adv_dev": [vp, i, i, d] + [vp] * 3 + [i, i, i],
        "nemo_fct_set_mus_metrics": [vp, vp, vp],
        "nemo_fct_set_e3uvw": [vp, vp, vp, vp, i],
        "nemo_fct_set_mus_upstream": [vp, i, vp, vp],
        "nemo_tra_adv_mus": [vp, i, i, C.c_char_p, d] + [vp] * 5 + [i],
        "nemo_tra_adv_mus_dev": [vp, i, i, C.c_char_p, d] + [vp] * 5 + [i],
        "nemo_group_tra_adv_mus_dev": [C.POINTER(vp), i, i, i, C.c_char_p, d] + [C.POINTER(vp)] * 5 + [i],
        "nemo_tra_adv_cen_dev": [vp, i, i, C.c_char_p] + [vp] * 5 + [i, i, i],
        "nemo_group_tra_adv_cen_dev": [C.POINTER(vp), i, i, i, C.c_char_p] + [C.POINTER(vp)] * 5 + [i, i, i],
        "nemo_fct_set_trend_diag": [vp, vp, vp, vp],
        "nemo_tra_nxt_dev": [vp, i, i, i, d, C.c_char_p, vp, vp, vp, vp, vp, vp, i],
        "nemo_group_tra_nxt_dev": [C.POINTER(vp), i, i, i, i, d, C.c_char_p] + [C.POINTER(vp)] * 6 + [i],
        "nemo_lbc_lnk_multi": [vp, C.c_char_p, i, C.POINTER(vp), C.c_char_p, dp, i, i, d],
        "nemo_lbc_lnk_multi_dev": [vp, C.c_char_p, i, C.POINTER(vp), C.c_char_p, dp, i, i, d],
        "nemo_group_lbc_lnk_multi_dev": [C.POINTER(vp), i, C.c_char_p, i, C.POINTER(C.POINTER(vp)), C.c_char_p, dp, i, i, d],
        "nemo_fct_comm_report": [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)],
        "nemo_fct_set_schedule": [vp, i],
        "nemo_fct_set_arithmetic": [vp, i],
        "nemo_fct_declare_transport_options": [vp, i, i, i, i],
        "nemo_fct_host_register": [vp, C.c_size_t],
        "nemo_fct_host_unregister": [vp],
        "nemo_glob_sum_dev": [vp, C.c_char_p, i, C.POINTER(vp), vp, vp, i, dp],
        "nemo_group_glob_sum_dev": [C.POINTER(vp), i, C.c_char_p, i, C.POINTER(C.POINTER(vp)), C.POINTER(vp), C.POINTER(vp), i, dp],
        "nemo_stp_ctl_dev": [vp, i, vp, vp, vp, i, vp],
        "nemo_fct_set_profiling": [vp, i],
        "nemo_fct_profile_read": [vp, i, C.c_char_p, i, dp, C.POINTER(C.c_longlong)],
        "nemo_fct_abi_version": [],
        "nemo_fct_selftest_division": [i, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)          # AttributeError here == a declared symbol is not exported
        fn.argtypes = argtypes
        if name not in ("nemo_fct_last_error", "nemo_fct_launch_count"):
            fn.restype = C.c_int
    _lib = L
    return L


#: every symbol include/nemo_fct.h declares (checked by the CPU test-suite against the built library)
ABI_SYMBOLS = (
    "nemo_mpp_init nemo_mpp_basic_decomposition nemo_lbc_plan_query nemo_fct_create nemo_fct_destroy "
    "nemo_fct_set_domain_arrays nemo_fct_set_e3t nemo_fct_set_stream nemo_fct_synchronize nemo_fct_comm_unique_id "
    "nemo_fct_comm_init nemo_fct_comm_init_local nemo_tra_adv_fct nemo_tra_adv_fct_dev nemo_group_tra_adv_fct_dev "
    "nemo_interp_4th_cpt nemo_interp_4th_cpt_dev nemo_tra_adv_transports_dev nemo_tra_adv_dev nemo_trc_adv_dev nemo_lbc_lnk_multi "
    "nemo_fct_set_mus_metrics nemo_fct_set_e3uvw nemo_fct_set_mus_upstream nemo_tra_adv_mus nemo_tra_adv_mus_dev "
    "nemo_group_tra_adv_mus_dev nemo_tra_nxt_dev nemo_group_tra_nxt_dev nemo_tra_adv_cen_dev nemo_group_tra_adv_cen_dev "
    "nemo_fct_set_trend_diag "
    "nemo_lbc_lnk_multi_dev nemo_group_lbc_lnk_multi_dev nemo_fct_last_error nemo_fct_abi_version "
    "nemo_fct_launch_count nemo_fct_comm_report nemo_fct_set_schedule nemo_fct_set_profiling "
    "nemo_fct_profile_read nemo_fct_selftest_division nemo_fct_set_arithmetic "
    "nemo_glob_sum_dev nemo_group_glob_sum_dev nemo_fct_declare_transport_options "
    "nemo_fct_host_register nemo_fct_host_unregister nemo_stp_ctl_dev").split()


class StpCtlResult(C.Structure):
    """nemo_stp_ctl_result: zmax(1:6) and the locations of stp_ctl (stpctl.F90:115-124, 162-165), kindic (:184)"""
    _fields_ = [("zmax", C.c_double * 6), ("ih", C.c_int * 2), ("iu", C.c_int * 3), ("is1", C.c_int * 3), ("is2", C.c_int * 3),
                ("nan_found", C.c_int), ("kindic", C.c_int)]


class NxtForcing(C.Structure):
    """nemo_nxt_forcing: the module variables tra_nxt_vvl reads (tranxt.F90:262-343).  Keyword arguments: scalars
    (atfp, r1_rau0, ln_traqsr, ln_rnf, ln_isf, ln_rnf_depth, nksr) or CUDA tensors (emp_b, emp, ..., ralpha)."""
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
            if hasattr(v, "data_ptr"):
                if not v.is_cuda:
                    raise ValueError(f"NxtForcing.{k}: device tensors only")
                self._keep[k] = v
                setattr(self, k, v.data_ptr())
            else:
                setattr(self, k, int(v))


def _check(rc):
    if rc != 0:
        raise NemoFctError(lib().nemo_fct_last_error().decode())


def _is_dev(a):
    return hasattr(a, "is_cuda") and a.is_cuda


def _ptr(a):
    """raw address of a numpy array (host) or torch tensor (host or device); must be contiguous fp64/int32"""
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "arrays must be contiguous"
        return C.c_void_p(a.data_ptr())
    assert a.flags["C_CONTIGUOUS"], "arrays must be contiguous"
    return C.c_void_p(a.ctypes.data)


def _f64(a, what):
    dt = str(a.dtype)
    if "float64" not in dt:
        raise TypeError(f"{what}: REAL(wp) arrays are float64, got {dt}")
    return a


def selftest_division(n=1 << 22, seed=1, device=0):
    """the fused kernel's inlined IEEE division against the compiler's x / y on n operand pairs per class; returns the number
    of results whose bit patterns differ (must be 0)"""
    nbad = C.c_longlong(-1)
    _check(lib().nemo_fct_selftest_division(device, n, seed, C.byref(nbad)))
    return int(nbad.value)


def host_register(a):
    """page-lock a host numpy array for the host-pointer entry points (cudaHostRegister)"""
    _check(lib().nemo_fct_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes))


def host_unregister(a):
    _check(lib().nemo_fct_host_unregister(a.ctypes.data_as(C.c_void_p)))


def launch_count():
    """CUDA kernels launched by the library in this process so far."""
    return int(lib().nemo_fct_launch_count())


def mpp_init(jpiglo, jpjglo, jpk, jperio, jpni=1, jpnj=1, narea=1, key_mpp_mpi=True):
    """mpp_init (mppini.F90:110-692) for rank ``narea`` (1-based): returns the filled :class:`Domain`."""
    dom = Domain()
    _check(lib().nemo_mpp_init(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, narea, int(key_mpp_mpi), C.byref(dom)))
    return dom


def mpp_basic_decomposition(jpiglo, jpjglo, jperio, jpni, jpnj):
    """mpp_basic_decomposition (mppini.F90:695-798): (jpimax, jpjmax, nimppt, njmppt, nlcit, nlcjt), tables (jpnj, jpni)."""
    n = jpni * jpnj
    tabs = [np.zeros(n, np.int32) for _ in range(4)]
    jpimax, jpjmax = C.c_int(), C.c_int()
    ip = C.POINTER(C.c_int)
    _check(lib().nemo_mpp_basic_decomposition(jpiglo, jpjglo, jperio, jpni, jpnj, C.byref(jpimax), C.byref(jpjmax),
                                              *[t.ctypes.data_as(ip) for t in tabs]))
    return (jpimax.value, jpjmax.value) + tuple(t.reshape(jpnj, jpni) for t in tabs)


def lbc_plan(dom, cd_nat, peer):
    """Compiled lbc_lnk gather plan of rank ``dom.narea`` for type ``cd_nat``: cells received from rank ``peer``
    (0-based; -1 = land-value fills) as arrays (dst_index, src_index, sgn_power) of local indices (i-1)+(j-1)*jpi."""
    L = lib()
    n = L.nemo_lbc_plan_query(C.byref(dom), cd_nat.encode(), peer, None, None, None)
    if n < 0:
        _check(1)
    dst, src, sp = (np.zeros(n, np.int32) for _ in range(3))
    ip = C.POINTER(C.c_int)
    if n:
        L.nemo_lbc_plan_query(C.byref(dom), cd_nat.encode(), peer, dst.ctypes.data_as(ip), src.ctypes.data_as(ip),
                              sp.ctypes.data_as(ip))
    return dst, src, sp


def comm_unique_id():
    """ncclUniqueId (128 bytes) to broadcast from rank 0 (mynode/MPI_Init analogue, lib_mpp.F90:197-331)."""
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(lib().nemo_fct_comm_unique_id(buf))
    return buf.raw


class FctContext:
    """One subdomain on one GPU: the device-resident state behind ``tra_adv_fct`` (created once, like nemo_alloc)."""

    def __init__(self, dom, device=-1):
        self.dom = dom
        self._h = C.c_void_p()
        _check(lib().nemo_fct_create(C.byref(dom), device, C.byref(self._h)))
        self._keep = {}

    # -- life cycle -------------------------------------------------------------------------------------------
    def close(self):
        if self._h:
            lib().nemo_fct_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_domain_arrays(self, tmask, umask, vmask, wmask, e1e2t, r1_e1e2t, mikt, mbkt, ln_linssh=False, ln_isfcav=False):
        """dom_oce.F90 module arrays (host numpy): masks (jpk,jpj,jpi), metrics (jpj,jpi), mikt/mbkt int32 (jpj,jpi)."""
        for a in (tmask, umask, vmask, wmask):
            assert tuple(a.shape) == self.dom.shape3, (a.shape, self.dom.shape3)
        mikt = np.ascontiguousarray(mikt, np.int32)
        mbkt = np.ascontiguousarray(mbkt, np.int32)
        arrs = [np.ascontiguousarray(_f64(a, "set_domain_arrays")) for a in (tmask, umask, vmask, wmask, e1e2t, r1_e1e2t)]
        _check(lib().nemo_fct_set_domain_arrays(self._h, *[_ptr(a) for a in arrs], _ptr(mikt), _ptr(mbkt),
                                                int(ln_linssh), int(ln_isfcav)))

    def set_e3t(self, e3t_b, e3t_n, e3t_a):
        """e3t_b/n/a (jpk,jpj,jpi): numpy -> copied to the device; CUDA tensors -> borrowed in place."""
        dev = _is_dev(e3t_n)
        for a in (e3t_b, e3t_n, e3t_a):
            assert tuple(a.shape) == self.dom.shape3 and _is_dev(a) == dev
            _f64(a, "set_e3t")
        if dev:
            self._keep["e3t"] = (e3t_b, e3t_n, e3t_a)
        _check(lib().nemo_fct_set_e3t(self._h, _ptr(e3t_b), _ptr(e3t_n), _ptr(e3t_a), int(dev)))

    def set_stream(self, stream_ptr):
        """run on the caller's CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``); None = own stream"""
        _check(lib().nemo_fct_set_stream(self._h, C.c_void_p(stream_ptr) if stream_ptr else None))

    def synchronize(self):
        _check(lib().nemo_fct_synchronize(self._h))

    def set_profiling(self, on):
        """CUDA-event timing of every kernel launch of this context (read with :meth:`profile`)"""
        _check(lib().nemo_fct_set_profiling(self._h, int(on)))

    def profile(self):
        """{kernel name: (total ms, launches)} since set_profiling(True)"""
        stride, mx = 32, 24
        names = C.create_string_buffer(stride * mx)
        ms = (C.c_double * mx)()
        calls = (C.c_longlong * mx)()
        n = lib().nemo_fct_profile_read(self._h, mx, names, stride, ms, calls)
        if n < 0:
            _check(1)
        return {names.raw[k * stride:(k + 1) * stride].split(b"\0")[0].decode(): (ms[k], int(calls[k])) for k in range(n)}

    def set_schedule(self, schedule):
        _check(lib().nemo_fct_set_schedule(self._h, schedule))

    def set_arithmetic(self, mode):
        """0 = STRICT (IEEE, bit-identical to the CPU restatement; default), 1 = FAST (relaxed divisions in schedule 4)"""
        _check(lib().nemo_fct_set_arithmetic(self._h, mode))

    def comm_init(self, unique_id, nranks, rank):
        _check(lib().nemo_fct_comm_init(self._h, unique_id, nranks, rank))

    def comm_report(self):
        n, b = C.c_longlong(), C.c_longlong()
        _check(lib().nemo_fct_comm_report(self._h, C.byref(n), C.byref(b)))
        return n.value, b.value

    # -- the hot path -----------------------------------------------------------------------------------------
    def tra_adv_fct(self, kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, ptn, pta, kjpt, kn_fct_h, kn_fct_v):
        """tra_adv_fct( kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, ptn, pta, kjpt, kn_fct_h, kn_fct_v )
        (traadv_fct.F90:54-55).  pta is updated in place on (2:jpim1, 2:jpjm1, 1:jpkm1, :)."""
        s3, s4 = self.dom.shape3, (kjpt,) + self.dom.shape3
        dev = _is_dev(pta)
        for a, shp in ((pun, s3), (pvn, s3), (pwn, s3), (ptb, s4), (ptn, s4), (pta, s4)):
            if tuple(a.shape) != shp:
                raise ValueError(f"tra_adv_fct: array of shape {tuple(a.shape)}, expected {shp}")
            if _is_dev(a) != dev:
                raise ValueError("tra_adv_fct: all arrays must live on the same side (all host or all device)")
            _f64(a, "tra_adv_fct")
        fn = lib().nemo_tra_adv_fct_dev if dev else lib().nemo_tra_adv_fct
        _check(fn(self._h, kt, kit000, cdtype.encode(), float(p2dt), _ptr(pun), _ptr(pvn), _ptr(pwn), _ptr(ptb),
                  _ptr(ptn), _ptr(pta), kjpt, kn_fct_h, kn_fct_v))

    def interp_4th_cpt(self, pt_in, pt_out):
        """interp_4th_cpt( pt_in, pt_out ) (traadv_fct.F90:517): pt_out defined on (2:jpim1,2:jpjm1,2:jpkm1)."""
        dev = _is_dev(pt_in)
        assert tuple(pt_in.shape) == self.dom.shape3 and tuple(pt_out.shape) == self.dom.shape3
        fn = lib().nemo_interp_4th_cpt_dev if dev else lib().nemo_interp_4th_cpt
        _check(fn(self._h, _ptr(_f64(pt_in, "interp_4th_cpt")), _ptr(_f64(pt_out, "interp_4th_cpt"))))

    def tra_adv_transports(self, e2u, e1v, e3u_n, e3v_n, un, vn, wn, zun, zvn, zwn):
        """zun = e2u*e3u_n*un, zvn = e1v*e3v_n*vn, zwn = e1e2t*wn (traadv.F90:100-124); device tensors only."""
        arrs = (e2u, e1v, e3u_n, e3v_n, un, vn, wn, zun, zvn, zwn)
        if not all(_is_dev(a) for a in arrs):
            raise ValueError("tra_adv_transports: device tensors only")
        _check(lib().nemo_tra_adv_transports_dev(self._h, *[_ptr(a) for a in arrs]))

    def tra_adv(self, kt, nit000, neuler, rdt, e2u, e1v, e3u_n, e3v_n, un, vn, wn, tsb, tsn, tsa, jpts, nn_fct_h, nn_fct_v):
        """tra_adv( kt ) of traadv.F90:77 with nadv = np_FCT, module state passed explicitly; device tensors only."""
        arrs = (e2u, e1v, e3u_n, e3v_n, un, vn, wn, tsb, tsn, tsa)
        if not all(_is_dev(a) for a in arrs):
            raise ValueError("tra_adv: device tensors only")
        _check(lib().nemo_tra_adv_dev(self._h, kt, nit000, neuler, float(rdt), *[_ptr(a) for a in arrs], jpts, nn_fct_h, nn_fct_v))

    def declare_transport_options(self, ln_wave_sdw=False, ln_vvl_ztilde_or_layer=False, ln_ldfeiv=False, ln_mle=False):
        """namelist switches of the host that add to the Eulerian transports in tra_adv (traadv.F90:103-129); tra_adv / trc_adv
        refuse to run with any of them set (they would silently drop the addition)"""
        _check(lib().nemo_fct_declare_transport_options(self._h, int(ln_wave_sdw), int(ln_vvl_ztilde_or_layer), int(ln_ldfeiv), int(ln_mle)))

    def trc_adv(self, kt, nittrc000, r2dttrc, trb, trn, tra, jptra, nn_fct_h, nn_fct_v):
        """trc_adv( kt ) of trcadv.F90:70 with nadv = np_FCT; reuses the transports built by the last tra_adv call."""
        if not all(_is_dev(a) for a in (trb, trn, tra)):
            raise ValueError("trc_adv: device tensors only")
        _check(lib().nemo_trc_adv_dev(self._h, kt, nittrc000, float(r2dttrc), _ptr(trb), _ptr(trn), _ptr(tra), jptra, nn_fct_h, nn_fct_v))

    # -- MUSCL (the scheme of the TOP namelists) and the time-stepping swap -----------------------------------------
    def set_mus_metrics(self, r1_e1e2u, r1_e1e2v):
        """r1_e1e2u, r1_e1e2v (jpj,jpi) host numpy (dom_oce.F90:118)"""
        arrs = [np.ascontiguousarray(_f64(a, "set_mus_metrics")) for a in (r1_e1e2u, r1_e1e2v)]
        for a in arrs:
            assert tuple(a.shape) == (self.dom.jpj, self.dom.jpi)
        _check(lib().nemo_fct_set_mus_metrics(self._h, _ptr(arrs[0]), _ptr(arrs[1])))

    def set_e3uvw(self, e3u_n, e3v_n, e3w_n):
        """e3u_n, e3v_n, e3w_n (jpk,jpj,jpi): numpy -> copied to the device; CUDA tensors -> borrowed in place."""
        dev = _is_dev(e3u_n)
        for a in (e3u_n, e3v_n, e3w_n):
            assert tuple(a.shape) == self.dom.shape3 and _is_dev(a) == dev
            _f64(a, "set_e3uvw")
        if dev:
            self._keep["e3uvw"] = (e3u_n, e3v_n, e3w_n)
        _check(lib().nemo_fct_set_e3uvw(self._h, _ptr(e3u_n), _ptr(e3v_n), _ptr(e3w_n), int(dev)))

    def set_mus_upstream(self, ld_msc_ups, rnfmsk=None, rnfmsk_z=None):
        """xind of traadv_mus.F90:99-113 from host rnfmsk (jpj,jpi), rnfmsk_z (jpk); ld_msc_ups = False -> xind = 1"""
        if not ld_msc_ups:
            _check(lib().nemo_fct_set_mus_upstream(self._h, 0, None, None))
            return
        a = np.ascontiguousarray(_f64(rnfmsk, "set_mus_upstream"))
        z = np.ascontiguousarray(_f64(rnfmsk_z, "set_mus_upstream"))
        assert tuple(a.shape) == (self.dom.jpj, self.dom.jpi) and z.shape == (self.dom.jpk,)
        _check(lib().nemo_fct_set_mus_upstream(self._h, 1, _ptr(a), _ptr(z)))

    def tra_adv_mus(self, kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, pta, kjpt):
        """tra_adv_mus( kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, pta, kjpt, ld_msc_ups ) (traadv_mus.F90:55-56);
        ld_msc_ups is the state set by :meth:`set_mus_upstream`.  pta updated on (2:jpim1, 2:jpjm1, 1:jpkm1, :)."""
        s3, s4 = self.dom.shape3, (kjpt,) + self.dom.shape3
        dev = _is_dev(pta)
        for a, shp in ((pun, s3), (pvn, s3), (pwn, s3), (ptb, s4), (pta, s4)):
            if tuple(a.shape) != shp:
                raise ValueError(f"tra_adv_mus: array of shape {tuple(a.shape)}, expected {shp}")
            if _is_dev(a) != dev:
                raise ValueError("tra_adv_mus: all arrays must live on the same side (all host or all device)")
            _f64(a, "tra_adv_mus")
        fn = lib().nemo_tra_adv_mus_dev if dev else lib().nemo_tra_adv_mus
        _check(fn(self._h, kt, kit000, cdtype.encode(), float(p2dt), _ptr(pun), _ptr(pvn), _ptr(pwn), _ptr(ptb), _ptr(pta), kjpt))

    def tra_adv_cen(self, kt, kit000, cdtype, pun, pvn, pwn, ptn, pta, kjpt, kn_cen_h, kn_cen_v):
        """tra_adv_cen( kt, kit000, cdtype, pun, pvn, pwn, ptn, pta, kjpt, kn_cen_h, kn_cen_v ) (traadv_cen.F90:46-47);
        device tensors only."""
        s3, s4 = self.dom.shape3, (kjpt,) + self.dom.shape3
        for a, shp in ((pun, s3), (pvn, s3), (pwn, s3), (ptn, s4), (pta, s4)):
            if not _is_dev(a) or tuple(a.shape) != shp:
                raise ValueError(f"tra_adv_cen: device tensor of shape {shp} expected, got {tuple(a.shape)}")
            _f64(a, "tra_adv_cen")
        _check(lib().nemo_tra_adv_cen_dev(self._h, kt, kit000, cdtype.encode(), _ptr(pun), _ptr(pvn), _ptr(pwn), _ptr(ptn),
                                          _ptr(pta), kjpt, kn_cen_h, kn_cen_v))

    def set_trend_diag(self, ztrdx=None, ztrdy=None, ztrdz=None):
        """l_trd / l_hst / l_ptr hooks of tra_adv_fct (traadv_fct.F90:96-112): device tensors (kjpt,jpk,jpj,jpi) that
        receive ztrdx, ztrdy (= zptry), ztrdz at every following tra_adv_fct call; no arguments = hooks off."""
        arrs = (ztrdx, ztrdy, ztrdz)
        if any(a is not None for a in arrs):
            for a in arrs:
                if a is None or not _is_dev(a):
                    raise ValueError("set_trend_diag: three device tensors or none")
                _f64(a, "set_trend_diag")
        self._keep["diag"] = arrs
        _check(lib().nemo_fct_set_trend_diag(self._h, *[None if a is None else _ptr(a) for a in arrs]))

    def tra_nxt(self, kt, nit000, l_euler, rdt, cdtype, forcing, ptb, ptn, pta, kjpt, sbc_tc=None, sbc_tc_b=None):
        """tra_nxt( kt ) (tranxt.F90:65) for cdtype 'TRA' / trc_nxt( kt ) (trcnxt.F90:56) for 'TRC', module state passed
        explicitly; device tensors only.  forcing: :class:`NxtForcing` (may be None for an Euler step)."""
        s4 = (kjpt,) + self.dom.shape3
        for a in (ptb, ptn, pta):
            if not _is_dev(a) or tuple(a.shape) != s4:
                raise ValueError("tra_nxt: device tensors of shape (kjpt, jpk, jpj, jpi) only")
            _f64(a, "tra_nxt")
        _check(lib().nemo_tra_nxt_dev(self._h, kt, nit000, int(l_euler), float(rdt), cdtype.encode(),
                                      C.byref(forcing) if forcing is not None else None, _ptr(ptb), _ptr(ptn), _ptr(pta),
                                      _ptr(sbc_tc) if sbc_tc is not None else None,
                                      _ptr(sbc_tc_b) if sbc_tc_b is not None else None, kjpt))

    def lbc_lnk_multi(self, cdname, *triplets, pval=None):
        """lbc_lnk_multi( cdname, pt1, cdna1, psgn1 [, pt2, cdna2, psgn2, ...] [, pval] )
        (lbc_lnk_multi_generic.h90:16-29).  Fields (ipk,jpj,jpi) or (jpj,jpi), all with the same ipk."""
        assert len(triplets) % 3 == 0 and triplets
        fields, nats, sgns = triplets[0::3], triplets[1::3], triplets[2::3]
        nfld = len(fields)
        jpj, jpi = self.dom.jpj, self.dom.jpi
        ipk = None
        for a in fields:
            if tuple(a.shape[-2:]) != (jpj, jpi):
                raise ValueError(f"lbc_lnk: field of shape {tuple(a.shape)}")
            k = int(np.prod(a.shape[:-2])) if len(a.shape) > 2 else 1
            if ipk is not None and k != ipk:
                raise ValueError("lbc_lnk_multi: all fields must have the same number of levels")
            ipk = k
            _f64(a, "lbc_lnk")
        dev = _is_dev(fields[0])
        tab = (C.c_void_p * nfld)(*[_ptr(a) for a in fields])
        sg = (C.c_double * nfld)(*[float(s) for s in sgns])
        fn = lib().nemo_lbc_lnk_multi_dev if dev else lib().nemo_lbc_lnk_multi
        _check(fn(self._h, cdname.encode(), nfld, tab, "".join(nats).encode(), sg, ipk, int(pval is not None),
                  float(pval or 0.0)))

    def glob_sum(self, cdname, fields, tmask_i, w3d=None):
        """glob_sum( cdname, ptab ) (lib_fortran_generic.h90:32-65) for each device field (ipk,jpj,jpi) of `fields`: masked sum in
        double-double over all ranks of the communicator; w3d (same shape) is multiplied in first.  Returns a list of floats."""
        nfld = len(fields)
        ipk = int(np.prod(fields[0].shape[:-2])) if len(fields[0].shape) > 2 else 1
        tab = (C.c_void_p * nfld)(*[_ptr(_f64(a, "glob_sum")) for a in fields])
        out = (C.c_double * nfld)()
        _check(lib().nemo_glob_sum_dev(self._h, cdname.encode(), nfld, tab, None if w3d is None else _ptr(w3d), _ptr(tmask_i), ipk, out))
        return list(out)

    def stp_ctl(self, kt, sshn, un, tsn, collective=False):
        """stp_ctl( kt, kindic ) (stpctl.F90:115-186) on device tensors sshn (jpj,jpi), un (jpk,jpj,jpi), tsn (2,jpk,jpj,jpi) = (tem, sal).
        Returns dict(zmax, ih, iu, is1, is2, nan_found, kindic, message)."""
        for a in (sshn, un, tsn):
            _f64(a, "stp_ctl")
        r = StpCtlResult()
        _check(lib().nemo_stp_ctl_dev(self._h, int(kt), _ptr(sshn), _ptr(un), _ptr(tsn), int(bool(collective)), C.byref(r)))
        return dict(zmax=list(r.zmax), ih=list(r.ih), iu=list(r.iu), is1=list(r.is1), is2=list(r.is2), nan_found=r.nan_found,
                    kindic=r.kindic, message=lib().nemo_fct_last_error().decode() if r.kindic else "")

    def lbc_lnk(self, cdname, pt, cd_nat, psgn, pval=None):
        """lbc_lnk( cdname, ptab, cd_nat, psgn [, pval] ) (lbclnk.F90:21-29)"""
        self.lbc_lnk_multi(cdname, pt, cd_nat, psgn, pval=pval)


class LocalGroup:
    """All jpni*jpnj subdomains of a layout inside ONE process on ONE GPU (decomposition tests on a single
    device).  Collective calls take per-rank lists of device tensors."""

    def __init__(self, jpiglo, jpjglo, jpk, jperio, jpni, jpnj, device=0):
        self.n = jpni * jpnj
        self.doms = [mpp_init(jpiglo, jpjglo, jpk, jperio, jpni, jpnj, r + 1) for r in range(self.n)]
        self.ctx = [FctContext(d, device) for d in self.doms]
        self._hs = (C.c_void_p * self.n)(*[c._h for c in self.ctx])
        _check(lib().nemo_fct_comm_init_local(self._hs, self.n))

    def close(self):
        for c in self.ctx:
            c.close()

    def _tab(self, arrs):
        return (C.c_void_p * self.n)(*[_ptr(a) for a in arrs])

    def tra_adv_fct(self, kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, ptn, pta, kjpt, kn_fct_h, kn_fct_v):
        for lst in (pun, pvn, pwn, ptb, ptn, pta):
            assert len(lst) == self.n and all(_is_dev(a) for a in lst)
        _check(lib().nemo_group_tra_adv_fct_dev(self._hs, self.n, kt, kit000, cdtype.encode(), float(p2dt),
                                                self._tab(pun), self._tab(pvn), self._tab(pwn), self._tab(ptb),
                                                self._tab(ptn), self._tab(pta), kjpt, kn_fct_h, kn_fct_v))

    def tra_adv_mus(self, kt, kit000, cdtype, p2dt, pun, pvn, pwn, ptb, pta, kjpt):
        for lst in (pun, pvn, pwn, ptb, pta):
            assert len(lst) == self.n and all(_is_dev(a) for a in lst)
        _check(lib().nemo_group_tra_adv_mus_dev(self._hs, self.n, kt, kit000, cdtype.encode(), float(p2dt),
                                                self._tab(pun), self._tab(pvn), self._tab(pwn), self._tab(ptb),
                                                self._tab(pta), kjpt))

    def tra_adv_cen(self, kt, kit000, cdtype, pun, pvn, pwn, ptn, pta, kjpt, kn_cen_h, kn_cen_v):
        for lst in (pun, pvn, pwn, ptn, pta):
            assert len(lst) == self.n and all(_is_dev(a) for a in lst)
        _check(lib().nemo_group_tra_adv_cen_dev(self._hs, self.n, kt, kit000, cdtype.encode(), self._tab(pun), self._tab(pvn),
                                                self._tab(pwn), self._tab(ptn), self._tab(pta), kjpt, kn_cen_h, kn_cen_v))

    def tra_nxt(self, kt, nit000, l_euler, rdt, cdtype, forcing, ptb, ptn, pta, kjpt, sbc_tc=None, sbc_tc_b=None):
        """forcing: list over ranks of NxtForcing (or None for an Euler step)"""
        for lst in (ptb, ptn, pta):
            assert len(lst) == self.n and all(_is_dev(a) for a in lst)
        ft = None if forcing is None else (C.c_void_p * self.n)(*[C.addressof(f) for f in forcing])
        _check(lib().nemo_group_tra_nxt_dev(self._hs, self.n, kt, nit000, int(l_euler), float(rdt), cdtype.encode(), ft,
                                            self._tab(ptb), self._tab(ptn), self._tab(pta),
                                            None if sbc_tc is None else self._tab(sbc_tc),
                                            None if sbc_tc_b is None else self._tab(sbc_tc_b), kjpt))

    def lbc_lnk_multi(self, cdname, fields, nats, sgns, pval=None):
        """fields[f][rank]: device tensors (ipk,jpj,jpi)"""
        nfld = len(fields)
        ipk = int(np.prod(fields[0][0].shape[:-2])) if len(fields[0][0].shape) > 2 else 1
        per_rank = [(C.c_void_p * nfld)(*[_ptr(fields[f][r]) for f in range(nfld)]) for r in range(self.n)]
        tabs = (C.POINTER(C.c_void_p) * self.n)(*[C.cast(t, C.POINTER(C.c_void_p)) for t in per_rank])
        sg = (C.c_double * nfld)(*[float(s) for s in sgns])
        _check(lib().nemo_group_lbc_lnk_multi_dev(self._hs, self.n, cdname.encode(), nfld, tabs, "".join(nats).encode(),
                                                  sg, ipk, int(pval is not None), float(pval or 0.0)))

    def glob_sum(self, cdname, fields, tmask_i, w3d=None):
        """fields[f][rank], tmask_i[rank], w3d[rank] (or None): the same sums as FctContext.glob_sum over the in-process group"""
        nfld = len(fields)
        ipk = int(np.prod(fields[0][0].shape[:-2])) if len(fields[0][0].shape) > 2 else 1
        per_rank = [(C.c_void_p * nfld)(*[_ptr(fields[f][r]) for f in range(nfld)]) for r in range(self.n)]
        tabs = (C.POINTER(C.c_void_p) * self.n)(*[C.cast(t, C.POINTER(C.c_void_p)) for t in per_rank])
        out = (C.c_double * nfld)()
        _check(lib().nemo_group_glob_sum_dev(self._hs, self.n, cdname.encode(), nfld, tabs, None if w3d is None else self._tab(w3d),
                                             self._tab(tmask_i), ipk, out))
        return list(out)

    def synchronize(self):
        self.ctx[0].synchronize()
