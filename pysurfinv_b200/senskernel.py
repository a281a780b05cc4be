"""Sensitivity kernels of the Rayleigh phase velocity, the quantity SensKernelPert (reference senskernel.py:130-158)
obtains by finite differences -- (c(1.001 v) - c(0.999 v)) / 0.2 / H per layer and period, 2 n + 1 forward solves per
model -- here from the analytic partial derivatives REIGEN computes alongside the group velocity (surfa.f:1130-1135,
1179-1185, 1202-1208; ``DispersionSolver.partials``), chained back through the attenuation correction
(calcul.f:121-127) and the earth flattening (flat1.f:33-62) to the input layer values.
"""
import numpy as np

R0 = 6371.0


def flatten_factors(h, kind=2):
    """Velocity and density factors of flat1.f:33-62 per layer (the last layer as the half-space)."""
    h = np.asarray(h, dtype=np.float64)
    n = len(h)
    pwr = 2.275 if kind == 2 else 5.0
    top = R0 - np.concatenate([[0.0], np.cumsum(h)[:-1]])        # radius of the top of every layer
    bot = R0 - np.cumsum(h)
    fv, fr = np.empty(n), np.empty(n)
    lg = np.log(top[:-1] / bot[:-1])
    fv[:-1] = (1.0 / bot[:-1] - 1.0 / top[:-1]) * R0 / lg
    fr[:-1] = (top[:-1] ** pwr - bot[:-1] ** pwr) / (lg * R0 ** pwr * pwr)
    fv[-1] = R0 / top[-1]
    fr[-1] = (top[-1] / R0) ** pwr
    return fv, fr


def input_kernels(part, vp, vs, rho, h, qsinv, periods, t_base=1.0):
    """dc/dVs, dc/dVp, dc/drho of the INPUT layer values [K, n] from REIGEN's partials of the prepared model.
    part: dict(dcda, dcdb, dcdr) [K, n] numpy.  Layers below the layer-dropping depth of a period were treated as part of
    the half-space by the solver (surfa.f:854-866): their kernels are zero, and the half-space factors belong to the
    deepest layer kept -- which the solver flattened as a regular layer, so regular factors are used throughout except
    for the model's last layer."""
    vp, vs, rho, h, q = (np.asarray(x, dtype=np.float64) for x in (vp, vs, rho, h, qsinv))
    fv, fr = flatten_factors(h)
    K = len(periods)
    dvs, dvp, drho = (np.zeros((K, len(h))) for _ in range(3))
    for k, T in enumerate(periods):
        qsq = q * np.log(t_base / T) / np.pi
        safe_vp = np.where(vp > 0, vp, 1.0)
        qpq = qsq * (4.0 / 3.0) * vs ** 2 / safe_vp ** 2
        dvs[k] = part["dcdb"][k] * (1.0 + qsq) * fv + part["dcda"][k] * fv * qsq * (8.0 / 3.0) * vs / safe_vp
        dvp[k] = part["dcda"][k] * fv * (1.0 - qpq)
        drho[k] = part["dcdr"][k] * fr
    return dvs, dvp, drho


class SensKernel:
    """kernel['Vs'][period, layer] and kernel['Vp'] in the units of SensKernelPert (senskernel.py:146-158):
    relative perturbation per km, (vH - vL) / 0.2 / H with vH, vL the phase velocities of the model perturbed by
    +-0.1 % in that layer = dc/dv * v * 0.01 / H."""

    def __init__(self, solver, vp, vs, rho, h, qsinv, periods):
        import torch
        vp, vs, rho, h, qsinv = (np.asarray(x, dtype=np.float64) for x in (vp, vs, rho, h, qsinv))
        keep = h > 1e-3                                     # senskernel.py:183 (same filter as models.py:20)
        keep[-1] = True
        lay = np.stack([vp[keep], vs[keep], rho[keep], h[keep], qsinv[keep]]).astype(np.float32)[:, None, :]
        nl = np.array([lay.shape[2]], np.int32)
        out = solver.partials(torch.from_numpy(np.ascontiguousarray(lay)).to(solver.device), torch.from_numpy(nl).to(solver.device), periods)
        part = {k: out[k][0].cpu().numpy().astype(np.float64) for k in ("dcda", "dcdb", "dcdr")}
        self.periods = np.asarray(periods, dtype=np.float64)
        self.c = out["c"][0].cpu().numpy()
        self.nfound = int(out["nfound"][0])
        self.partials = part
        f32 = lambda x: x[keep].astype(np.float32).astype(np.float64)
        dvs, dvp, drho = input_kernels(part, f32(vp), f32(vs), f32(rho), f32(h), f32(qsinv), self.periods)
        H = h[keep]
        self.dcdvs, self.dcdvp, self.dcdrho = dvs, dvp, drho
        self.kernel = {"Vs": dvs * vs[keep] * 0.01 / H, "Vp": dvp * vp[keep] * 0.01 / H}
