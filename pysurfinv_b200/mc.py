"""Ensemble Monte-Carlo inversion on the GPU: the loop of reference point.py:40-85 (Point.MCinv) run for M
independent chains at once, every stage on the device -- proposal (surfdisp_mc_propose), model assembly
(surfdisp_build_stacks), dispersion (surfdisp_batch), misfit (surfdisp_misfit_batch), Metropolis rule
(surfdisp_mc_accept).  The host only launches; nothing is copied back until the track is saved.

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
    def __init__(self, solver, template, periods, obs, sigma, n_chains, seed=42, mask=None, misfit_mode=0, kind=api.KIND_RAYLEIGH,
                 lmax=None):
        torch = solver.torch
        self.solver, self.template, self.torch = solver, template, torch
        self.periods = np.ascontiguousarray(periods, dtype=np.float32)
        self.obs = np.ascontiguousarray(obs, dtype=np.float32)
        self.sigma = np.ascontiguousarray(sigma, dtype=np.float32)
        self.mask, self.misfit_mode, self.kind = mask, misfit_mode, kind
        self.M, self.P, self.seed = int(n_chains), template.nparams, int(seed)
        self.lmax = int(lmax) if lmax is not None else template.max_layers()
        dev = solver.device
        self.cur = torch.from_numpy(np.tile(template.start_values(), (self.M, 1))).to(dev).contiguous()
        self.prop = torch.empty_like(self.cur)
        self.chi0 = torch.full((self.M,), 88888.0, dtype=torch.float32, device=dev)
        self.accepted = torch.empty(self.M, dtype=torch.uint8, device=dev)
        self.status = torch.empty(self.M, dtype=torch.int32, device=dev)
        self.stacks = (torch.empty((5, self.M, self.lmax), dtype=torch.float32, device=dev),
                       torch.empty(self.M, dtype=torch.int32, device=dev))
        self.fwd = None
        self.step_index = 0
        self.track = []

    def _evaluate(self, params):
        s = self.solver
        s.build_stacks(self.template, params, lmax=self.lmax, out=self.stacks)
        self.fwd = s.forward(self.stacks[0], self.stacks[1], self.periods, kind=self.kind, group=False, out=self.fwd)
        return s.misfit(self.fwd["c"], self.fwd["nfound"], self.obs, self.sigma, mask=self.mask, periods=self.periods,
                        mode=self.misfit_mode)

    def step(self, restart=False, first=False, record=True):
        """One Monte-Carlo step of every chain.  first: evaluate the start model as it is (point.py:47-50, unless
        it violates the priors: then it is perturbed first); restart: uniform redraw (point.py:52)."""
        torch, s = self.torch, self.solver
        k = self.step_index
        force = None
        if first or restart:
            force = torch.ones(self.M, dtype=torch.uint8, device=s.device)
        if first:
            bad = s.check_priors(self.template, self.cur) & self.template.prior_mask
            s.mc_propose(self.template, self.cur, self.seed, k, out=self.prop, status=self.status)
            self.prop = torch.where((bad != 0)[:, None], self.prop, self.cur).contiguous()
        else:
            s.mc_propose(self.template, self.cur, self.seed, k, reset_mask=force if restart else None, out=self.prop,
                         status=self.status)
        m3 = self._evaluate(self.prop)          # [M, 3] = (misfit, chiSqr, L)
        chi1 = m3[:, 1].contiguous()
        s.mc_accept(chi1, self.prop, self.chi0, self.cur, self.seed, k, force_mask=force, accepted=self.accepted)
        if record:
            # rows as Model1D._dump writes them (models.py:243-245): [misfit, L, accepted, parameters of the PROPOSAL]
            self.track.append(torch.cat([m3[:, 0:1], m3[:, 2:3], self.accepted[:, None].float(), self.prop], dim=1))
        self.step_index += 1
        return m3

    def run(self, chain_length, restart_first=False):
        """chain_length steps per chain; the first step takes the start model (or a uniform redraw)."""
        for i in range(chain_length):
            self.step(first=(i == 0 and not restart_first), restart=(i == 0 and restart_first))
        return self

    def mc_track(self):
        """[M, steps, 3 + P] numpy array."""
        return self.torch.stack(self.track, dim=1).cpu().numpy()

    def save_point_npz(self, outdir, pid, setting, chain_length=None):
        """<outdir>/<pid>.npz: all sub-chains of this point merged like Point.MCinvMP does (point.py:112-123)."""
        tr = self.mc_track()
        obs = {"T": self.periods, "c": self.obs, "uncer": self.sigma}
        return write_point_npz(os.path.join(outdir, "%s.npz" % pid), tr, setting, obs, pid, chain_length or tr.shape[1])

    def save_npz(self, outdir, pid, setting, chain_length=None):
        """One file per chain, in the layout of Point.MCinv (point.py:78-85)."""
        os.makedirs(outdir, exist_ok=True)
        tr = self.mc_track()
        obs = {"T": self.periods, "c": self.obs, "uncer": self.sigma}
        paths = []
        for i in range(self.M):
            name = "tmp_%03d_%s" % (i, pid)
            path = os.path.join(outdir, name + ".npz")
            np.savez_compressed(path, mcTrack=tr[i].astype(np.float64), setting=dict(setting), obs=obs,
                                invMeta={"pid": name, "chainL": chain_length or tr.shape[1]})
            paths.append(path)
        return paths
