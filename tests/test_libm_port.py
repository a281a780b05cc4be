"""CPU tier: sd_libm.cuh (the glibc logf / powf restatement used by the model-preparation kernel) against the
host libm, bit for bit, over every float32 of the argument ranges the path can produce: radii 6371 - depth
for depths down to 3000 km, radius ratios r_i / r_(i+1) and a / r just above 1, and r / a just below 1
(flat1.f:44-68).  The oracle calls the host libm; the CUDA kernel runs the same header."""
import ctypes as C

import pytest

from tests.hostmirror import mirror as HM


@pytest.mark.parametrize("lo,hi", [(3371.0, 6371.5), (0.5, 2.0), (1.0e-3, 1.1e-3), (100.0, 100.5)])
def test_logf_powf_match_host_libm(lo, hi):
    L = HM.lib()
    L.hm_libm_check.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_longlong)]
    L.hm_libm_check.restype = None
    out = (C.c_longlong * 4)()
    L.hm_libm_check(lo, hi, out)
    assert out[0] > 1000
    assert (out[1], out[2], out[3]) == (0, 0, 0), "mismatches logf/powf2.275/powf5: %s of %d" % (list(out)[1:], out[0])
