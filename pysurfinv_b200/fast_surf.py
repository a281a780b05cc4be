"""Drop-in for the f2py module ``pySurfInv.fast_surf`` (reference fast_surf_src/fast_surf.pyf:6-19).

    (ur0, ul0, cr0, cl0) = fast_surf.fast_surf(nlay, ilvry, Vp, Vs, rho, h, qsinv, per, nper)

is the call made at reference models.py:27 and senskernel.py:188.  Copy (or symlink) this file as
``pySurfInv/fast_surf.py`` next to the reference's Python sources, in place of the compiled
``fast_surf*.so`` (compile_fast_surf.sh:7), and the Monte-Carlo loop runs on the B200 kernels.

Differences from the f2py module, all documented in INTEGRATION.md:
* outputs are fully defined (zeros beyond the found prefix / for the other wave type) instead of
  uninitialised memory;
* the call is re-entrant (no COMMON blocks) and ``ndiv`` does not decay across calls.
"""
import ctypes as C

import numpy as np

from . import api

_NPER = 200


def fast_surf(n_layer0, kind0, a_ref0, b_ref0, rho_ref0, d_ref0, qs_ref0, cvper, ncvper):
    n = int(n_layer0)
    f32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64)[:n], dtype=np.float32)  # f2py casts to real*4
    a, b, rho, d, qs = (f32(x) for x in (a_ref0, b_ref0, rho_ref0, d_ref0, qs_ref0))
    for x in (a, b, rho, d, qs):
        if x.size != n:
            raise ValueError("fast_surf: layer arrays must have n_layer0 = %d entries" % n)
    per = np.ascontiguousarray(cvper, dtype=np.float32)
    if per.size != _NPER:
        raise ValueError("fast_surf: cvper must have %d entries (fast_surf.pyf:14)" % _NPER)
    ur0 = np.zeros(_NPER, np.float32); ul0 = np.zeros(_NPER, np.float32)
    cr0 = np.zeros(_NPER, np.float32); cl0 = np.zeros(_NPER, np.float32)
    L = api.load_library()
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    L.fast_surf_(C.byref(C.c_int(n)), C.byref(C.c_int(int(kind0))), fp(a), fp(b), fp(rho), fp(d), fp(qs), fp(per),
                 C.byref(C.c_int(int(ncvper))), fp(ur0), fp(ul0), fp(cr0), fp(cl0))
    return ur0, ul0, cr0, cl0
