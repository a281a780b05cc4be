"""ctypes front-end of the CPU oracle (oracle/surfdisp_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (pysurfinv_b200) never imports this module.

The oracle restates /root/reference/fast_surf_src/{fast_surf,init,calcul,flat1,surfa}.f (and, with
``sibling=1``, senskernel-1.0/src/SURF_PERTURB/*.f, the program that wrote TEST1/*).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleOpts(C.Structure):
    _fields_ = [
        ("precision", C.c_int), ("sibling", C.c_int), ("nmode", C.c_int), ("ndiv", C.c_int),
        ("ndiv_cap_r", C.c_int), ("ndiv_cap_l", C.c_int), ("neville_cap", C.c_int),
        ("stale_mmax", C.c_int), ("atten", C.c_int), ("flat", C.c_int),
        ("dc", C.c_double), ("fact", C.c_double), ("t_base", C.c_double),
    ]


class OracleCounters(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in
                ("sweeps_R", "steps_R", "sweeps_L", "steps_L", "sub_U", "flat_layers", "scan_evals",
                 "polish_evals")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    so = os.path.join(_HERE, "libsurfdisp_oracle.so")
    src = os.path.join(_HERE, "surfdisp_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libsurfdisp_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.surfdisp_oracle_default_opts.argtypes = [C.POINTER(OracleOpts)]
        L.surfdisp_oracle.argtypes = [C.POINTER(OracleOpts), C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_int,
                                      dp, dp, dp, dp, dp, ip, C.POINTER(OracleCounters)]
        L.surfdisp_oracle.restype = C.c_int
        L.surfdisp_oracle_batch.argtypes = [C.POINTER(OracleOpts), C.c_int, C.c_int, C.c_int, ip, dp,
                                            C.c_int, dp, dp, dp, ip, ip, C.POINTER(OracleCounters), C.c_int]
        L.surfdisp_oracle_batch.restype = C.c_int
        L.surfdisp_oracle_partials.argtypes = [C.POINTER(OracleOpts), C.c_int, dp, dp, dp, dp, dp, C.c_int, dp, dp, dp, dp, dp]
        L.surfdisp_oracle_partials.restype = C.c_int
        _LIB = L
    return _LIB


def make_opts(**kw):
    o = OracleOpts()
    lib().surfdisp_oracle_default_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError("unknown oracle option %r" % k)
        setattr(o, k, v)
    return o


def sibling_opts(**kw):
    """Options reproducing senskernel-1.0 SURF_PERTURB (real*8, multi-mode) -- the TEST1 generator."""
    base = dict(precision=2, sibling=1, nmode=2, ndiv_cap_r=999, ndiv_cap_l=999, neville_cap=5000)
    base.update(kw)
    return make_opts(**base)


def _dp(x):
    return x.ctypes.data_as(C.POINTER(C.c_double))


def forward(kind, vp, vs, rho, h, qsinv, periods, opts=None, counters=None):
    """One model.  Returns dict(c, u, ratio, cvar: [nmode, nper]; imax: [nmode]; status)."""
    o = opts if opts is not None else make_opts()
    a = np.ascontiguousarray(vp, dtype=np.float64)
    b = np.ascontiguousarray(vs, dtype=np.float64)
    r = np.ascontiguousarray(rho, dtype=np.float64)
    d = np.ascontiguousarray(h, dtype=np.float64)
    q = np.ascontiguousarray(qsinv, dtype=np.float64)
    per = np.ascontiguousarray(periods, dtype=np.float64)
    n, K, nm = len(b), len(per), o.nmode
    c = np.zeros((nm, K)); u = np.zeros((nm, K)); rt = np.zeros((nm, K)); cv = np.zeros((nm, K))
    imax = np.zeros(nm, dtype=np.int32)
    st = lib().surfdisp_oracle(C.byref(o), int(kind), n, _dp(a), _dp(b), _dp(r), _dp(d), _dp(q), K, _dp(per),
                               _dp(c), _dp(u), _dp(rt), _dp(cv), imax.ctypes.data_as(C.POINTER(C.c_int)),
                               C.byref(counters) if counters is not None else None)
    if st < 0:
        raise ValueError("oracle: bad arguments")
    return dict(c=c, u=u, ratio=rt, cvar=cv, imax=imax, status=st)


def forward_batch(kind, layers, nlay, periods, opts=None, nthreads=1, counters=None):
    """layers: [5, M, Lmax] (vp, vs, rho, h, 1/Qs).  Returns c[M,K], u[M,K], nfound[M], status[M]."""
    o = opts if opts is not None else make_opts()
    lay = np.ascontiguousarray(layers, dtype=np.float64)
    assert lay.ndim == 3 and lay.shape[0] == 5
    M, lmax = lay.shape[1], lay.shape[2]
    nl = np.ascontiguousarray(nlay, dtype=np.int32)
    per = np.ascontiguousarray(periods, dtype=np.float64)
    K = len(per)
    c = np.zeros((M, K)); u = np.zeros((M, K))
    nf = np.zeros(M, dtype=np.int32); st = np.zeros(M, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    rc = lib().surfdisp_oracle_batch(C.byref(o), int(kind), M, lmax, nl.ctypes.data_as(ip), _dp(lay), K,
                                     _dp(per), _dp(c), _dp(u), nf.ctypes.data_as(ip), st.ctypes.data_as(ip),
                                     C.byref(counters) if counters is not None else None, int(nthreads))
    if rc != 0:
        raise ValueError("oracle batch: bad arguments")
    return c, u, nf, st


def partials(vp, vs, rho, h, qsinv, periods, opts=None):
    """Rayleigh fundamental-mode partial derivatives of REIGEN (surfa.f:1130-1135, 1179-1185, 1202-1208):
    dict(c [K], dcda, dcdb, dcdr [K, n]) with respect to Vp, Vs, density of the layers of each period's
    attenuation-corrected, flattened model."""
    o = opts if opts is not None else make_opts()
    a = np.ascontiguousarray(vp, dtype=np.float64); b = np.ascontiguousarray(vs, dtype=np.float64)
    r = np.ascontiguousarray(rho, dtype=np.float64); d = np.ascontiguousarray(h, dtype=np.float64)
    q = np.ascontiguousarray(qsinv, dtype=np.float64); per = np.ascontiguousarray(periods, dtype=np.float64)
    n, K = len(b), len(per)
    c = np.zeros(K); da = np.zeros((K, n)); db = np.zeros((K, n)); dr = np.zeros((K, n))
    st = lib().surfdisp_oracle_partials(C.byref(o), n, _dp(a), _dp(b), _dp(r), _dp(d), _dp(q), K, _dp(per), _dp(c), _dp(da), _dp(db), _dp(dr))
    if st < 0:
        raise ValueError("oracle: bad arguments")
    return dict(c=c, dcda=da, dcdb=db, dcdr=dr, status=st)
