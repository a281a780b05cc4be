"""Ensemble Monte-Carlo inversion on the GPU: the loop of reference point.py:40-85 (Point.MCinv) run for M
independent chains at once, every stage on the device (surfdisp_mc_step: proposal + model assembly in one kernel,
the dispersion solver, misfit + Metropolis rule + state update + track row in one kernel).  After a first eager
step the launches of one step are captured in a CUDA graph and replayed: the host only replays; nothing is copied
back until the track is saved.  The chains of several points (grid nodes, model3D.py:50-57) run side by side.

The reference runs runN steps as runN/chainL sub-chains in a process pool (point.py:87-105); here the sub-chains
are the rows of the ensemble.  ``save_npz`` writes, per sub-chain, the file layout Point.MCinv writes
(point.py:78-85: mcTrack rows = [misfit, L, accepted, free parameters...], setting, obs, invMeta), so the
reference's PostPoint / Model3D tools read GPU-produced chains unchanged.
"""
import os

import numpy as np

from . import api
from .stack import Param, _is_spec

# layer.prop['LayerName'] of the reference's layer classes (layers.py:142,161,194,209,224,242,270,292): the keys
# Model1D.toYML (models.py:53-62) writes.  'LandSediment' / 'LandCrust' are not keys of the reference's own
# layerClassDict (layers.py:552-568), so a file written with them cannot be read back by buildModel1D; for those
# two the setting keeps the type key the user gave (documented deviation in favour of a readable file).
_LAYER_NAME = {"Sediment": "Sediment", "Crust": "Crust", "Mantle": "OceanMantle", "OceanMantle": "OceanMantle",
               "OceanWater": "OceanWater", "OceanSediment": "OceanSediment", "OceanCrust": "OceanCrust",
               "OceanSedimentCascadia": "OceanSedimentCascadia", "ReferenceMantle": "ReferenceMantle",
               "OceanMantleHybrid": "OceanMantleHybrid"}


def setting_to_yml(setting):
    """The dict Model1D.toYML() (reference models.py:53-62) returns for a model built from `setting`: every free
    parameter as [v, vmin, vmax, step] (BrownianVar), fixed specs as plain numbers, 'Info' last."""
    def conv(v):
        if _is_spec(v):
            if v[1] in ("fixed", "total"):
                return v[0]
            p = Param(v, "")
            return [p.v0, p.vmin, p.vmax, p.step]
        if isinstance(v, (list, tuple)):
            return [conv(x) for x in v]
        return v
    out = {}
    for name, parm in setting.items():
        if name == "Info":
            continue
        d = {k: conv(v) for k, v in dict(parm).items()}
        if name == "OceanWater":
            d["Vs"] = 0          # layers.py:195
        out[_LAYER_NAME.get(name, name)] = d
    out["Info"] = dict(setting.get("Info", {}))
    return out


def write_point_npz(path, track, setting, obs, pid, chain_length):
    """The per-point file Point.MCinvMP writes (reference point.py:112-123) and PostPoint / Model3D.loadInvDir read
    (point.py:139-149, model3D.py:36-57): the sub-chains concatenated in order.
    track: [n_subchains, steps, 3 + P] rows [misfit, L, accepted, parameters] (Model1D._dump, models.py:243-245)."""
    tr = np.asarray(track, dtype=np.float64)
    mc = tr.reshape(-1, tr.shape[-1])
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, mcTrack=mc, setting=setting_to_yml(setting),
                        obs={k: (v if np.ma.isMaskedArray(v) else np.asarray(v)) for k, v in obs.items()},
                        invMeta={"pid": pid, "chainL": int(chain_length)})
    return path


class ChainEnsemble:
    """n_points x chains_per_point chains.  obs / sigma / mask: [K] (one point) or [n_points, K]; start: None (the
    template's start values), [P] or [n_points, P]; bounds: None (the template's) or (lo, hi, step) arrays
    [n_points, P] (per-point boxes, e.g. a sediment thickness from a local map, models.py:262-292).
    chain_length: sub-chain length (chainL of point.py:32): every chain_length steps a sub-chain starts; the chains
    flagged in ``init`` (default: the first chain of every point, point.py:101 `i==0`) start from the start model,
    the others from a uniform redraw (point.py:47-52)."""

    def __init__(self, solver, template, periods, obs, sigma, n_chains, seed=42, mask=None, misfit_mode=0, kind=api.KIND_RAYLEIGH,
                 lmax=None, n_points=1, chain_length=0, start=None, bounds=None, init=None, track_steps=0, use_graph=True):
        torch = solver.torch
        self.solver, self.template, self.torch = solver, template, torch
        self.periods = np.ascontiguousarray(periods, dtype=np.float32)
        K = self.periods.size
        self.n_points, self.cpp = int(n_points), int(n_chains)
        self.M, self.P, self.seed, self.K = self.n_points * self.cpp, template.nparams, int(seed), K
        self.misfit_mode, self.kind, self.chain_length = int(misfit_mode), int(kind), int(chain_length)
        dev = solver.device
        bc = lambda x, dt: np.array(np.broadcast_to(np.asarray(x, dtype=dt), (self.n_points, K)))   # (writable copies)
        self.obs, self.sigma = bc(obs, np.float32), bc(sigma, np.float32)
        self.mask = None if mask is None else bc(mask, np.uint8)
        use = np.ones((self.n_points, K), np.uint8) if mask is None else (self.mask != 0).astype(np.uint8)
        if np.any(use.sum(axis=1) == 0):
            raise ValueError("All observations are masked???")          # point.py:356
        lo, hi, st = template.bounds() if bounds is None else bounds
        # (array stride of the stacks: the most layers any model inside the boxes can have)
        self.lmax = int(lmax) if lmax is not None else template.max_layers(lo, hi)
        bnd = np.zeros((self.n_points, 3, 64), np.float32)
        for i, b in enumerate((lo, hi, st)):
            bnd[:, i, :self.P] = np.broadcast_to(np.asarray(b, np.float32), (self.n_points, self.P))
        if not (np.all(bnd[:, 1, :self.P] > bnd[:, 0, :self.P]) and np.all(bnd[:, 2, :self.P] > 0)):
            raise ValueError("bounds: vmax > vmin and step > 0 required")
        sv = template.start_values() if start is None else np.asarray(start, np.float32)
        sv = np.broadcast_to(sv, (self.n_points, self.P))
        self.cur = torch.from_numpy(np.repeat(sv, self.cpp, axis=0).copy()).to(dev).contiguous()
        self.prop = torch.empty_like(self.cur)
        self.chi0 = torch.full((self.M,), 88888.0, dtype=torch.float32, device=dev)
        self.accepted = torch.zeros(self.M, dtype=torch.uint8, device=dev)
        self.status = torch.zeros(self.M, dtype=torch.int32, device=dev)
        if init is None:
            init = np.zeros((self.n_points, self.cpp), np.uint8); init[:, 0] = 1
        self.init = torch.from_numpy(np.ascontiguousarray(init, np.uint8).reshape(-1)).to(dev)
        self.misfit3 = torch.empty((self.M, 3), dtype=torch.float32, device=dev)
        self.d_bounds = torch.from_numpy(bnd).to(dev)
        self.d_obs = torch.from_numpy(self.obs).to(dev)
        self.d_isig = torch.from_numpy((1.0 / self.sigma).astype(np.float32)).to(dev)
        self.d_use = torch.from_numpy(use).to(dev)
        self.d_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.stacks = (torch.empty((5, self.M, self.lmax), dtype=torch.float32, device=dev),
                       torch.empty(self.M, dtype=torch.int32, device=dev))
        self.c_pred = torch.empty((self.M, K), dtype=torch.float32, device=dev)
        self.c_cur = torch.zeros((self.M, K), dtype=torch.float32, device=dev)    # curves of the current models: search hints
        self.nfound = torch.empty(self.M, dtype=torch.int32, device=dev)
        self.flags = torch.empty(self.M, dtype=torch.int32, device=dev)
        self.ws = torch.empty(int(solver.lib.surfdisp_workspace_bytes(self.M, self.lmax, K)), dtype=torch.uint8, device=dev)
        self.track_steps = int(track_steps)
        self.track = torch.zeros((max(self.track_steps, 1), self.M, 3 + self.P), dtype=torch.float32, device=dev)
        self.step_index = 0
        self.use_graph = bool(use_graph)
        self._graph = None
        self._tc = template.to_c()
        self._state = self._make_state()

    def _make_state(self):
        s = api.SurfdispMcState()
        s.n_chains, s.n_params, s.n_periods, s.n_layers_max = self.M, self.P, self.K, self.lmax
        s.kind, s.misfit_mode, s.chain_len, s.chains_per_point = self.kind, self.misfit_mode, self.chain_length, self.cpp
        s.track_steps, s.seed = self.track_steps, self.seed
        p = lambda t: t.data_ptr()
        s.cur, s.prop, s.chi0, s.status, s.accepted = p(self.cur), p(self.prop), p(self.chi0), p(self.status), p(self.accepted)
        s.init_mask, s.misfit, s.track, s.step = p(self.init), p(self.misfit3), p(self.track), p(self.d_step)
        s.bounds, s.obs, s.isig, s.use = p(self.d_bounds), p(self.d_obs), p(self.d_isig), p(self.d_use)
        s.layers, s.n_layers, s.c_pred, s.nfound, s.flags = p(self.stacks[0]), p(self.stacks[1]), p(self.c_pred), p(self.nfound), p(self.flags)
        s.c_cur = p(self.c_cur)
        s.workspace, s.workspace_bytes = p(self.ws), self.ws.numel()
        return s

    def _launch(self):
        s = self.solver
        import ctypes as C
        with self.torch.cuda.device(s.device):
            stream = self.torch.cuda.current_stream(s.device).cuda_stream
            rc = s.lib.surfdisp_mc_step(C.byref(s.opts), C.byref(self._tc), C.byref(self._state),
                                        self.periods.ctypes.data_as(C.POINTER(C.c_float)), stream)
        api._check(rc, "surfdisp_mc_step")

    def step(self):
        """One Monte-Carlo step of every chain (the first one eagerly; from the second on as a graph replay)."""
        torch = self.torch
        if not self.use_graph or self.step_index == 0:
            self._launch()
        else:
            if self._graph is None:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(self.solver.device)
                side.wait_stream(torch.cuda.current_stream(self.solver.device))
                with torch.cuda.stream(side):
                    # capture does not execute: the step counter and the state are untouched by it
                    with torch.cuda.graph(g, stream=side):
                        self._launch()
                torch.cuda.current_stream(self.solver.device).wait_stream(side)
                self._graph = g
            self._graph.replay()
        self.step_index += 1

    def run(self, n_steps):
        """n_steps steps of every chain; raises like the reference (models.py:216-219) if a chain found no admissible
        model within 1000 perturbations + 10000 uniform redraws."""
        if self.track_steps and self.step_index + n_steps > self.track_steps:
            raise ValueError("track buffer holds %d steps" % self.track_steps)
        for _ in range(n_steps):
            self.step()
        if int(self.status.min().item()) < 0:
            raise RuntimeError("Error: Cound not find a good model through reset.")
        return self

    def mc_track(self):
        """[M, steps, 3 + P] numpy array (steps recorded so far)."""
        n = min(self.step_index, self.track_steps)
        return self.track[:n].permute(1, 0, 2).contiguous().cpu().numpy()

    def best_misfit(self):
        """Per point: smallest misfit any of its chains has recorded, device tensor [n_points]."""
        n = min(self.step_index, self.track_steps)
        return self.track[:n, :, 0].amin(dim=0).view(self.n_points, self.cpp).amin(dim=1)

    def save_point_npz(self, outdir, pid, setting, chain_length=None, point=0):
        """<outdir>/<pid>.npz: all sub-chains of one point merged like Point.MCinvMP does (point.py:112-123)."""
        tr = self.mc_track()[point * self.cpp:(point + 1) * self.cpp]
        obs = {"T": self.periods, "c": self.obs[point], "uncer": self.sigma[point]}
        if self.mask is not None:
            obs["c"] = np.ma.masked_array(self.obs[point], mask=(self.mask[point] == 0))
        return write_point_npz(os.path.join(outdir, "%s.npz" % pid), tr, setting, obs, pid, chain_length or tr.shape[1])

    def save_npz(self, outdir, pid, setting, chain_length=None):
        """One file per chain, in the layout of Point.MCinv (point.py:78-85)."""
        os.makedirs(outdir, exist_ok=True)
        tr = self.mc_track()
        paths = []
        for i in range(self.M):
            pt = i // self.cpp
            obs = {"T": self.periods, "c": self.obs[pt], "uncer": self.sigma[pt]}
            name = "tmp_%03d_%s" % (i, pid)
            path = os.path.join(outdir, name + ".npz")
            np.savez_compressed(path, mcTrack=tr[i].astype(np.float64), setting=setting_to_yml(setting), obs=obs,
                                invMeta={"pid": name, "chainL": chain_length or tr.shape[1]})
            paths.append(path)
        return paths
