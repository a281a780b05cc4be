"""TEST ONLY: builds and wraps tests/hostmirror/hostmirror.cpp (g++ build of the CUDA kernels' per-lane math)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libhostmirror.so")
        src = os.path.join(_HERE, "hostmirror.cpp")
        hdr = os.path.join(_HERE, "..", "..", "pysurfinv_b200", "csrc", "surfdisp_core.cuh")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so, src, "-lm"])
        L = C.CDLL(so)
        fp = C.POINTER(C.c_float)
        L.hm_forward.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp, C.c_int, fp, C.c_float, C.c_float,
                                 C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp, fp,
                                 C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_longlong)]
        L.hm_forward.restype = C.c_int
        L.hm_partials.argtypes = [C.c_int, fp, fp, fp, fp, fp, C.c_int, fp, fp, fp, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_float, fp, fp, fp]
        L.hm_partials.restype = None
        _LIB = L
    return _LIB


def partials(vp, vs, rho, h, qsinv, periods, c, ratio, ndiv=5, ndiv_cap=99):
    """REIGEN partial derivatives of the kernels' per-lane code (host build): dcda, dcdb, dcdr [K, n]."""
    f = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    a, b, r, d, q, per, cc, rt = f(vp), f(vs), f(rho), f(h), f(qsinv), f(periods), f(c), f(ratio)
    K, n = len(per), len(b)
    out = [np.zeros((K, n), np.float32) for _ in range(3)]
    p = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    lib().hm_partials(n, p(a), p(b), p(r), p(d), p(q), K, p(per), p(cc), p(rt), 1.0, 1, 1, ndiv, ndiv_cap, 4.0, p(out[0]), p(out[1]), p(out[2]))
    return dict(dcda=out[0], dcdb=out[1], dcdr=out[2])


def forward(kind, vp, vs, rho, h, qsinv, periods, G=4, stale=1, ndiv=5, ndiv_cap=None, exact_scan=0):
    f = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    a, b, r, d, q, per = f(vp), f(vs), f(rho), f(h), f(qsinv), f(periods)
    K = len(per)
    c = np.zeros(K, np.float32); u = np.zeros(K, np.float32); rt = np.zeros(K, np.float32)
    p = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    sw = C.c_longlong(0)
    rounds = (C.c_longlong * 6)(0, 0, 0, 0, 0, 0)
    if ndiv_cap is None:
        ndiv_cap = 99 if kind == 2 else 999
    nf = lib().hm_forward(G, kind, len(b), p(a), p(b), p(r), p(d), p(q), K, p(per), 0.01, 4.0, 1.0, 1, 1, stale,
                          ndiv, ndiv_cap, p(c), p(u), p(rt), C.byref(sw), exact_scan, rounds)
    return dict(c=c, u=u, ratio=rt, nfound=nf, sweeps=sw.value, rounds=rounds[0], slow_periods=rounds[1], windows=rounds[2], windows_ok=rounds[3], direct=rounds[4], coarse_events=rounds[5])
