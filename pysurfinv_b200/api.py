"""Host side of the B200 dispersion solver: ctypes over the C ABI (include/surfdisp_b200.h), torch for
device buffers and streams only.

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsurfdisp_b200.so")

KIND_LOVE = 1
KIND_RAYLEIGH = 2
MAX_PERIODS = 200

F_NO_ROOT_FIRST = 1
F_NO_ROOT_AT_K = 2
F_ROOT_ABOVE_HS = 4
F_SCAN_LIMIT = 8


class SurfdispOpts(C.Structure):
    """Mirror of ``SurfdispOpts`` in include/surfdisp_b200.h (defaults = reference init.f:25,43-58)."""
    _fields_ = [("dc", C.c_float), ("fact", C.c_float), ("t_base", C.c_float), ("ndiv", C.c_int),
                ("ndiv_cap_rayleigh", C.c_int), ("ndiv_cap_love", C.c_int), ("atten", C.c_int),
                ("flatten", C.c_int), ("stale_mmax", C.c_int), ("compute_group", C.c_int), ("exact_scan", C.c_int),
                ("group_f64", C.c_int)]


class SurfdispMcState(C.Structure):
    """Mirror of ``SurfdispMcState`` in include/surfdisp_b200.h (one Monte-Carlo step of an ensemble of chains)."""
    _fields_ = [("n_chains", C.c_int), ("n_params", C.c_int), ("n_periods", C.c_int), ("n_layers_max", C.c_int),
                ("kind", C.c_int), ("misfit_mode", C.c_int), ("chain_len", C.c_int), ("chains_per_point", C.c_int),
                ("track_steps", C.c_int), ("pad_", C.c_int), ("seed", C.c_ulonglong),
                ("cur", C.c_void_p), ("prop", C.c_void_p), ("chi0", C.c_void_p), ("status", C.c_void_p),
                ("accepted", C.c_void_p), ("init_mask", C.c_void_p), ("misfit", C.c_void_p), ("track", C.c_void_p),
                ("step", C.c_void_p), ("bounds", C.c_void_p), ("obs", C.c_void_p), ("isig", C.c_void_p), ("use", C.c_void_p),
                ("layers", C.c_void_p), ("n_layers", C.c_void_p), ("c_pred", C.c_void_p), ("nfound", C.c_void_p),
                ("flags", C.c_void_p), ("c_cur", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class SurfdispError(RuntimeError):
    pass


_lib = None


def load_library():
    """Loads libsurfdisp_b200.so; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SurfdispError("CUDA library %s not built: run `python -m pysurfinv_b200.build`" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.surfdisp_default_opts.argtypes = [C.POINTER(SurfdispOpts)]
    L.surfdisp_default_opts.restype = None
    L.surfdisp_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.surfdisp_workspace_bytes.restype = C.c_size_t
    L.surfdisp_batch.argtypes = [C.POINTER(SurfdispOpts), C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, fp,
                                 vp, vp, vp, vp, vp, C.c_size_t, vp]
    L.surfdisp_batch.restype = C.c_int
    L.surfdisp_batch_hinted.argtypes = [C.POINTER(SurfdispOpts), C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, fp, vp,
                                        vp, vp, vp, vp, vp, C.c_size_t, vp]
    L.surfdisp_batch_hinted.restype = C.c_int
    L.surfdisp_partials_batch.argtypes = [C.POINTER(SurfdispOpts), C.c_int, C.c_int, vp, vp, C.c_int, fp, vp, vp, vp, vp, vp, vp,
                                          vp, C.c_size_t, vp]
    L.surfdisp_partials_batch.restype = C.c_int
    L.surfdisp_misfit_batch.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, fp, fp, C.POINTER(C.c_ubyte), fp, vp, vp]
    L.surfdisp_misfit_batch.restype = C.c_int
    L.surfdisp_host_batch.argtypes = [C.POINTER(SurfdispOpts), C.c_int, C.c_int, C.c_int, C.c_int, ip, fp,
                                      C.c_int, fp, fp, fp, ip, ip]
    L.surfdisp_host_batch.restype = C.c_int
    L.surfdisp_pipelined_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.surfdisp_pipelined_bytes.restype = C.c_size_t
    L.surfdisp_host_batch_pipelined.argtypes = [C.POINTER(SurfdispOpts), C.c_int, C.c_int, C.c_int, ip, fp, C.c_int, fp,
                                                fp, fp, ip, ip, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    L.surfdisp_host_batch_pipelined.restype = C.c_int
    L.fast_surf_.argtypes = [ip, ip, fp, fp, fp, fp, fp, fp, ip, fp, fp, fp, fp]
    L.fast_surf_.restype = None
    L.surfdisp_read_counters.argtypes = [vp, C.POINTER(C.c_ulonglong), vp]
    L.surfdisp_read_counters.restype = C.c_int
    L.surfdisp_batch_profiled.argtypes = L.surfdisp_batch.argtypes + [fp]
    L.surfdisp_batch_profiled.restype = C.c_int
    L.surfdisp_measure_peaks.argtypes = [C.POINTER(C.c_double)]
    L.surfdisp_measure_peaks.restype = C.c_int
    from . import stack as _stack
    L.surfdisp_build_stacks.argtypes = [C.POINTER(_stack.StackTemplateC), C.c_int, vp, C.c_int, vp, vp, vp]
    L.surfdisp_build_stacks.restype = C.c_int
    ub = C.POINTER(C.c_ubyte)
    L.surfdisp_check_priors.argtypes = [C.POINTER(_stack.StackTemplateC), C.c_int, vp, vp, vp]
    L.surfdisp_check_priors.restype = C.c_int
    L.surfdisp_mc_propose.argtypes = [C.POINTER(_stack.StackTemplateC), C.c_int, fp, fp, fp, vp, vp, vp, vp,
                                      C.c_ulonglong, C.c_uint, vp]
    L.surfdisp_mc_propose.restype = C.c_int
    L.surfdisp_mc_accept.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_ulonglong, C.c_uint, vp]
    L.surfdisp_mc_accept.restype = C.c_int
    L.surfdisp_params_pipelined_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.surfdisp_params_pipelined_bytes.restype = C.c_size_t
    L.surfdisp_host_params_pipelined.argtypes = [C.POINTER(SurfdispOpts), C.POINTER(_stack.StackTemplateC), C.c_int, C.c_int, C.c_int,
                                                 fp, C.c_int, fp, fp, fp, ip, ip, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    L.surfdisp_host_params_pipelined.restype = C.c_int
    L.surfdisp_mc_step.argtypes = [C.POINTER(SurfdispOpts), C.POINTER(_stack.StackTemplateC), C.POINTER(SurfdispMcState), fp, vp]
    L.surfdisp_mc_step.restype = C.c_int
    L.surfdisp_host_release.argtypes = []
    L.surfdisp_host_release.restype = None
    L.surfdisp_set_split_min_models.argtypes = [C.c_int]
    L.surfdisp_set_split_min_models.restype = None
    L.surfdisp_version.restype = C.c_char_p
    L.surfdisp_last_cuda_error.restype = C.c_char_p
    _lib = L
    return L


def default_opts(**kw):
    o = SurfdispOpts()
    load_library().surfdisp_default_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError("unknown option %r" % k)
        setattr(o, k, v)
    return o


def _check(rc, what):
    if rc != 0:
        names = {-1: "invalid argument", -2: "workspace too small / out of memory", -3: "CUDA error"}
        msg = "%s failed: %s" % (what, names.get(rc, rc))
        if rc == -3:
            msg += " (%s)" % load_library().surfdisp_last_cuda_error().decode()
        raise SurfdispError(msg)


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class DispersionSolver:
    """Batched Rayleigh/Love phase + group velocity on one GPU.

    ``layers`` follows the argument order of ``fast_surf.fast_surf`` (reference fast_surf.pyf:6-19):
    [5, M, Lmax] = (Vp, Vs, rho, h, 1/Qs), top layer first, last layer = half-space.
    """

    def __init__(self, device=None, opts=None):
        import torch
        if not torch.cuda.is_available():
            raise SurfdispError("no CUDA device visible: the dispersion solver has no CPU path")
        self.torch = torch
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.lib = load_library()
        self.opts = opts if opts is not None else default_opts()
        self._ws = None
        self._pinned = {}

    # -- device path -------------------------------------------------------------------------------
    def workspace(self, M, lmax, K):
        need = int(self.lib.surfdisp_workspace_bytes(M, lmax, K))
        if self._ws is None or self._ws.numel() < need:
            self._ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        return self._ws

    def measure_peaks(self):
        """(FP32 FMA TFLOP/s, FP64 FMA TFLOP/s, MUFU ex2 Top/s) measured with register-resident chains."""
        arr = (C.c_double * 3)()
        with self.torch.cuda.device(self.device):
            _check(self.lib.surfdisp_measure_peaks(arr), "surfdisp_measure_peaks")
        return tuple(float(x) for x in arr)

    def forward(self, layers, nlay, periods, kind=KIND_RAYLEIGH, group=True, out=None, kernel_ms=None, hint=None):
        """layers: float32 device tensor [5, M, Lmax]; nlay: int32 device tensor [M]; periods: host
        sequence.  Returns dict of device tensors c[M,K], u[M,K], nfound[M], flags[M].  Asynchronous on
        the current torch stream.  hint: optional float32 device tensor [M, K], phase velocities of a nearby model
        per model (surfdisp_batch_hinted): fewer sweeps, same results."""
        torch = self.torch
        if layers.dtype != torch.float32 or layers.dim() != 3 or layers.shape[0] != 5 or not layers.is_contiguous():
            raise ValueError("layers must be a contiguous float32 tensor [5, M, Lmax]")
        if nlay.dtype != torch.int32 or not nlay.is_contiguous() or nlay.numel() != layers.shape[1]:
            raise ValueError("nlay must be a contiguous int32 tensor [M]")
        if layers.device != self.device or nlay.device != self.device:
            raise ValueError("inputs must live on %s" % self.device)
        per = np.ascontiguousarray(periods, dtype=np.float32)
        M, lmax, K = int(layers.shape[1]), int(layers.shape[2]), int(per.size)
        if K < 1 or K > MAX_PERIODS:
            raise ValueError("1 <= len(periods) <= %d" % MAX_PERIODS)
        if out is None:
            out = dict(c=torch.empty((M, K), dtype=torch.float32, device=self.device),
                       u=torch.empty((M, K), dtype=torch.float32, device=self.device) if group else None,
                       nfound=torch.empty(M, dtype=torch.int32, device=self.device),
                       flags=torch.empty(M, dtype=torch.int32, device=self.device))
        if M == 0:
            return out
        ws = self.workspace(M, lmax, K)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            args = [C.byref(self.opts), int(kind), M, lmax, nlay.data_ptr(), layers.data_ptr(),
                    K, _fptr(per), out["c"].data_ptr(),
                    out["u"].data_ptr() if (group and out.get("u") is not None) else None,
                    out["nfound"].data_ptr(), out["flags"].data_ptr(), ws.data_ptr(), ws.numel(), stream]
            if hint is not None:
                if hint.dtype != torch.float32 or tuple(hint.shape) != (M, K) or not hint.is_contiguous() or hint.device != self.device:
                    raise ValueError("hint must be a contiguous float32 device tensor [M, K]")
                rc = self.lib.surfdisp_batch_hinted(*args[:8], hint.data_ptr(), *args[8:])
            elif kernel_ms is None:
                rc = self.lib.surfdisp_batch(*args)
            else:  # list that receives [prep_ms, phase1_ms, phase2_ms]; synchronises
                ms = (C.c_float * 3)()
                rc = self.lib.surfdisp_batch_profiled(*args, ms)
                kernel_ms[:] = [float(x) for x in ms]
        _check(rc, "surfdisp_batch")
        return out

    def partials(self, layers, nlay, periods):
        """Rayleigh phase velocities and REIGEN's partial derivatives (surfa.f:1130-1135, 1179-1185, 1202-1208): dict of
        device tensors c [M, K], dcda / dcdb / dcdr [M, K, Lmax] (with respect to Vp, Vs, density of the layers of each
        period's attenuation-corrected, flattened model), nfound, flags."""
        torch = self.torch
        per = np.ascontiguousarray(periods, dtype=np.float32)
        M, lmax, K = int(layers.shape[1]), int(layers.shape[2]), int(per.size)
        out = dict(c=torch.empty((M, K), dtype=torch.float32, device=self.device),
                   nfound=torch.empty(M, dtype=torch.int32, device=self.device), flags=torch.empty(M, dtype=torch.int32, device=self.device))
        for k in ("dcda", "dcdb", "dcdr"):
            out[k] = torch.empty((M, K, lmax), dtype=torch.float32, device=self.device)
        if M == 0:
            return out
        ws = self.workspace(M, lmax, K)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.surfdisp_partials_batch(C.byref(self.opts), M, lmax, nlay.data_ptr(), layers.data_ptr(), K, _fptr(per),
                                                  out["c"].data_ptr(), out["dcda"].data_ptr(), out["dcdb"].data_ptr(), out["dcdr"].data_ptr(),
                                                  out["nfound"].data_ptr(), out["flags"].data_ptr(), ws.data_ptr(), ws.numel(), stream)
        _check(rc, "surfdisp_partials_batch")
        return out

    def counters(self):
        """(layer_steps, sweeps, u_sublayers, models) executed by the last forward()."""
        arr = (C.c_ulonglong * 4)()
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        _check(self.lib.surfdisp_read_counters(self._ws.data_ptr(), arr, stream), "surfdisp_read_counters")
        return tuple(int(x) for x in arr)

    def misfit(self, c_pred, nfound, obs, sigma, mask=None, periods=None, mode=0):
        """Per-model (misfit, chiSqr, L) like Point.misfit (mode 0) / PointCascadia.misfit (mode 1)."""
        torch = self.torch
        M, K = int(c_pred.shape[0]), int(c_pred.shape[1])
        o = np.ascontiguousarray(obs, dtype=np.float32)
        s = np.ascontiguousarray(sigma, dtype=np.float32)
        if o.size != K or s.size != K:
            raise ValueError("obs/sigma must have one entry per period")
        mk = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        pr = None if periods is None else np.ascontiguousarray(periods, dtype=np.float32)
        out = torch.empty((M, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.surfdisp_misfit_batch(int(mode), M, K, c_pred.data_ptr(), nfound.data_ptr(), _fptr(o),
                                                _fptr(s), None if mk is None else mk.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                                None if pr is None else _fptr(pr), out.data_ptr(), stream)
        _check(rc, "surfdisp_misfit_batch")
        return out

    def build_stacks(self, template, params, lmax=None, out=None):
        """Device-side model assembly (reference models.py:72-102 + layers.py, for M models at once).
        template: pysurfinv_b200.stack.StackTemplate; params: float32 device tensor [M, template.nparams]
        (column order = order of the free parameters in the setting, like MCinv._brownians).
        Returns (layers [5, M, lmax], nlay [M]) device tensors ready for forward()."""
        torch = self.torch
        P = template.nparams
        if params.dtype != torch.float32 or params.dim() != 2 or params.shape[1] != P or not params.is_contiguous():
            raise ValueError("params must be a contiguous float32 tensor [M, %d]" % P)
        if params.device != self.device:
            raise ValueError("params must live on %s" % self.device)
        M = int(params.shape[0])
        lmax = int(lmax) if lmax is not None else template.max_layers()
        if out is None:
            out = (torch.empty((5, M, lmax), dtype=torch.float32, device=self.device),
                   torch.empty(M, dtype=torch.int32, device=self.device))
        tc = template.to_c()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.surfdisp_build_stacks(C.byref(tc), M, params.data_ptr() if P else None, lmax,
                                                out[0].data_ptr(), out[1].data_ptr(), stream)
        _check(rc, "surfdisp_build_stacks")
        return out

    def check_priors(self, template, params):
        """SURFDISP P_* bits of the prior rules each model violates (CascadiaPrism / CascadiaContinent / CascadiaOcean
        .isgood, reference models.py:294-360, 385-523, 571-677; all rules are evaluated); int32 device tensor [M]."""
        torch = self.torch
        M = int(params.shape[0])
        out = torch.empty(M, dtype=torch.int32, device=self.device)
        tc = template.to_c()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.surfdisp_check_priors(C.byref(tc), M, params.data_ptr() if template.nparams else None,
                                                out.data_ptr(), stream)
        _check(rc, "surfdisp_check_priors")
        return out

    def mc_propose(self, template, cur, seed, step_index, reset_mask=None, out=None, status=None):
        """One proposal per chain (MCinv.perturb / reset, reference models.py:190-219, brownian.py:17-27)."""
        torch = self.torch
        lo, hi, st = template.bounds()
        M, P = int(cur.shape[0]), template.nparams
        if out is None:
            out = torch.empty_like(cur)
        tc = template.to_c()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.surfdisp_mc_propose(C.byref(tc), M, _fptr(lo), _fptr(hi), _fptr(st), cur.data_ptr(),
                                              None if reset_mask is None else reset_mask.data_ptr(), out.data_ptr(),
                                              None if status is None else status.data_ptr(), int(seed), int(step_index), stream)
        _check(rc, "surfdisp_mc_propose")
        return out

    def mc_accept(self, chi1, prop, chi0, cur, seed, step_index, force_mask=None, accepted=None):
        """Metropolis rule of reference point.py:34-37; chi0 and cur are updated in place where accepted."""
        torch = self.torch
        M, P = int(cur.shape[0]), int(cur.shape[1])
        if accepted is None:
            accepted = torch.empty(M, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.surfdisp_mc_accept(M, P, chi1.data_ptr(), prop.data_ptr(), chi0.data_ptr(), cur.data_ptr(),
                                             None if force_mask is None else force_mask.data_ptr(), accepted.data_ptr(),
                                             int(seed), int(step_index), stream)
        _check(rc, "surfdisp_mc_accept")
        return accepted

    # -- host path (what a reference-side caller uses): pinned staging, H2D, solve, D2H --------------
    def _pin(self, key, shape, dtype):
        t = self._pinned.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self.torch.empty(shape, dtype=dtype, pin_memory=True)
            self._pinned[key] = t
        return t

    def forward_host(self, layers, nlay, periods, kind=KIND_RAYLEIGH, group=True, chunks=None):
        """numpy in, numpy out.  Inputs are staged through pinned memory, copied to the GPU, solved and
        copied back; returns dict(c, u, nfound, flags) of numpy arrays the caller owns (copies of the pinned
        staging buffers: a later call does not change them)."""
        torch = self.torch
        lay = np.ascontiguousarray(layers, dtype=np.float32)
        nl = np.ascontiguousarray(nlay, dtype=np.int32)
        hl = self._pin("lay", lay.shape, torch.float32)
        hn = self._pin("nl", nl.shape, torch.int32)
        hl.numpy()[...] = lay
        hn.numpy()[...] = nl
        r = self.forward_pinned(hl, hn, periods, kind, group, chunks)
        return {k: (None if v is None else v.copy()) for k, v in r.items()}

    def forward_pinned(self, hl, hn, periods, kind=KIND_RAYLEIGH, group=True, chunks=None):
        """Same as forward_host but the caller already holds pinned host tensors (hl float32 [5][M][L], hn int32 [M]).
        One call of surfdisp_host_batch_pipelined: the batch is cut into `chunks` chunks (default 8 from 65536
        models on); a chunk is copied host->device while the previous one is prepared and its first period
        searched, the later periods run as one launch over the whole batch, the group velocities are computed
        chunk by chunk and copied out under the next chunk.  Batches above 2^21 models go through the same call
        in pieces of 2^21 (bounded device memory).

        The returned arrays are VIEWS of pinned result buffers that this solver keeps and reuses: the next
        forward_pinned / forward_host call with the same shapes overwrites them (zero-copy hand-over for callers
        that consume a result before asking for the next one; forward_host returns copies)."""
        torch = self.torch
        M, lmax = int(hl.shape[1]), int(hl.shape[2])
        per = np.ascontiguousarray(periods, dtype=np.float32)
        K = int(per.size)
        hc = self._pin("c", (M, K), torch.float32)
        hf = self._pin("nf", (M,), torch.int32)
        hg = self._pin("fl", (M,), torch.int32)
        hu = self._pin("u", (M, K), torch.float32) if group else None
        if M == 0:
            return dict(c=hc.numpy(), u=None if hu is None else hu.numpy(), nfound=hf.numpy(), flags=hg.numpy())
        piece = 1 << 21
        compute = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        copy = self._copy_stream
        ip = C.POINTER(C.c_int)

        def run(lay_t, nl_t, a, b):
            m = b - a
            nch = chunks if chunks is not None else (8 if m >= (1 << 16) else 1)
            nch = max(1, min(int(nch), m))
            need = int(self.lib.surfdisp_pipelined_bytes(m, lmax, K))
            if getattr(self, "_pipe_buf", None) is None or self._pipe_buf.numel() < need:
                self._pipe_buf = None
                self._pipe_buf = torch.empty(need, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                rc = self.lib.surfdisp_host_batch_pipelined(
                    C.byref(self.opts), int(kind), m, lmax, C.cast(C.c_void_p(nl_t.data_ptr()), ip),
                    C.cast(C.c_void_p(lay_t.data_ptr()), C.POINTER(C.c_float)), K, _fptr(per),
                    C.cast(C.c_void_p(hc[a:b].data_ptr()), C.POINTER(C.c_float)),
                    C.cast(C.c_void_p(hu[a:b].data_ptr()), C.POINTER(C.c_float)) if group else None,
                    C.cast(C.c_void_p(hf[a:b].data_ptr()), ip), C.cast(C.c_void_p(hg[a:b].data_ptr()), ip),
                    C.c_void_p(self._pipe_buf.data_ptr()), C.c_size_t(self._pipe_buf.numel()), nch,
                    C.c_void_p(compute.cuda_stream), C.c_void_p(copy.cuda_stream))
            _check(rc, "surfdisp_host_batch_pipelined")

        if M <= piece:
            run(hl, hn, 0, M)
        else:
            for a in range(0, M, piece):
                b = min(M, a + piece)
                st = self._pin("lay_piece", (5, b - a, lmax), torch.float32)
                st.copy_(hl[:, a:b])
                run(st, hn[a:b], a, b)
        return dict(c=hc.numpy(), u=None if hu is None else hu.numpy(), nfound=hf.numpy(), flags=hg.numpy())


def _forward_params_pinned(self, template, h_params, periods, kind=KIND_RAYLEIGH, group=True, chunks=None, lmax=None):
    """Model1D.forward for a batch (reference models.py:93-121), host buffers in, host buffers out: h_params pinned
    float32 [M][P] parameter vectors of `template`; one surfdisp_host_params_pipelined call copies them (4 P bytes
    per model), assembles the stacks on the device, solves and copies c, U, nfound, flags back.  Returns views of
    pinned result buffers the solver reuses (see forward_pinned)."""
    torch = self.torch
    M, P = int(h_params.shape[0]), template.nparams
    if h_params.dtype != torch.float32 or h_params.dim() != 2 or int(h_params.shape[1]) != P or not h_params.is_contiguous():
        raise ValueError("params must be a contiguous float32 tensor [M, %d]" % P)
    per = np.ascontiguousarray(periods, dtype=np.float32)
    K = int(per.size)
    lmax = int(lmax) if lmax is not None else template.max_layers()
    hc = self._pin("pc", (M, K), torch.float32)
    hf = self._pin("pnf", (M,), torch.int32)
    hg = self._pin("pfl", (M,), torch.int32)
    hu = self._pin("pu", (M, K), torch.float32) if group else None
    if M == 0:
        return dict(c=hc.numpy(), u=None if hu is None else hu.numpy(), nfound=hf.numpy(), flags=hg.numpy())
    compute = torch.cuda.current_stream(self.device)
    if getattr(self, "_copy_stream", None) is None:
        self._copy_stream = torch.cuda.Stream(self.device)
    nch = chunks if chunks is not None else (8 if M >= (1 << 16) else 1)
    need = int(self.lib.surfdisp_params_pipelined_bytes(M, P, lmax, K))
    if getattr(self, "_pipe_buf", None) is None or self._pipe_buf.numel() < need:
        self._pipe_buf = None
        self._pipe_buf = torch.empty(need, dtype=torch.uint8, device=self.device)
    tc = template.to_c()
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_float)
    with torch.cuda.device(self.device):
        rc = self.lib.surfdisp_host_params_pipelined(
            C.byref(self.opts), C.byref(tc), int(kind), M, lmax, C.cast(C.c_void_p(h_params.data_ptr()), fp), K, _fptr(per),
            C.cast(C.c_void_p(hc.data_ptr()), fp), C.cast(C.c_void_p(hu.data_ptr()), fp) if group else None,
            C.cast(C.c_void_p(hf.data_ptr()), ip), C.cast(C.c_void_p(hg.data_ptr()), ip),
            C.c_void_p(self._pipe_buf.data_ptr()), C.c_size_t(self._pipe_buf.numel()), max(1, min(int(nch), M)),
            C.c_void_p(compute.cuda_stream), C.c_void_p(self._copy_stream.cuda_stream))
    _check(rc, "surfdisp_host_params_pipelined")
    return dict(c=hc.numpy(), u=None if hu is None else hu.numpy(), nfound=hf.numpy(), flags=hg.numpy())


DispersionSolver.forward_params_pinned = _forward_params_pinned


def host_batch(layers, nlay, periods, kind=KIND_RAYLEIGH, group=True, opts=None, device=0):
    """Pure C-ABI host call (surfdisp_host_batch): no torch involved.  numpy in, numpy out."""
    L = load_library()
    lay = np.ascontiguousarray(layers, dtype=np.float32)
    nl = np.ascontiguousarray(nlay, dtype=np.int32)
    per = np.ascontiguousarray(periods, dtype=np.float32)
    M, lmax, K = lay.shape[1], lay.shape[2], per.size
    c = np.zeros((M, K), np.float32)
    u = np.zeros((M, K), np.float32) if group else None
    nf = np.zeros(M, np.int32)
    fl = np.zeros(M, np.int32)
    ip = C.POINTER(C.c_int)
    o = opts if opts is not None else default_opts()
    rc = L.surfdisp_host_batch(C.byref(o), int(device), int(kind), M, lmax, nl.ctypes.data_as(ip), _fptr(lay), K,
                               _fptr(per), _fptr(c), None if u is None else _fptr(u), nf.ctypes.data_as(ip),
                               fl.ctypes.data_as(ip))
    _check(rc, "surfdisp_host_batch")
    return dict(c=c, u=u, nfound=nf, flags=fl)
