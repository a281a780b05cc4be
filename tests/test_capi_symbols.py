"""The CUDA C-ABI library builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every
symbol include/surfdisp_b200.h declares.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "surfdisp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\(", src)
    return sorted({n for n in names if n.startswith("surfdisp_") or n == "fast_surf_"})


def test_library_exports_every_declared_symbol():
    from pysurfinv_b200 import build
    so = build.build()
    lib = ctypes.CDLL(so)
    names = _declared()
    assert "fast_surf_" in names and "surfdisp_batch" in names and len(names) >= 9
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_argument_validation_without_gpu():
    from pysurfinv_b200 import api
    L = api.load_library()
    assert L.surfdisp_version().startswith(b"surfdisp_b200")
    o = api.default_opts()
    assert abs(o.dc - 0.01) < 1e-9 and o.fact == 4.0 and o.ndiv == 5 and o.ndiv_cap_rayleigh == 99 and o.stale_mmax == 1
    assert L.surfdisp_workspace_bytes(0, 1, 1) == 0
    assert L.surfdisp_workspace_bytes(1000, 77, 40) >= 1000 * 8 * 80 * 4
    per = (ctypes.c_float * 4)(8, 10, 12, 14)
    # bad kind / bad sizes are rejected before any CUDA call
    assert L.surfdisp_batch(None, 3, 1, 8, None, None, 4, per, None, None, None, None, None, 0, None) == -1
    assert L.surfdisp_batch(None, 2, 1, 8, None, None, 0, per, None, None, None, None, None, 0, None) == -1
    assert L.surfdisp_batch(None, 2, 0, 8, None, None, 4, per, None, None, None, None, None, 0, None) == 0  # empty batch
    assert L.surfdisp_misfit_batch(5, 1, 4, None, None, per, per, None, None, None, None) == -1
    # pipelined host path: device block = inputs + outputs + workspace; argument checks before any CUDA call
    ws = L.surfdisp_workspace_bytes(1000, 77, 40)
    pb = L.surfdisp_pipelined_bytes(1000, 77, 40)
    assert pb >= ws + 5 * 1000 * 77 * 4 + 2 * 1000 * 40 * 4 + 3 * 1000 * 4
    assert L.surfdisp_pipelined_bytes(0, 1, 1) == 0
    assert L.surfdisp_host_batch_pipelined(None, 2, 1, 8, None, None, 4, per, None, None, None, None, None, 0, 4, None, None) == -1
    assert L.surfdisp_host_batch_pipelined(None, 2, 0, 8, None, None, 4, per, None, None, None, None, None, 0, 4, None, None) == 0


def test_product_fails_loudly_without_gpu():
    import torch
    from pysurfinv_b200 import api
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.SurfdispError):
        api.DispersionSolver()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pysurfinv_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "libsurfdisp_oracle" not in txt, f
