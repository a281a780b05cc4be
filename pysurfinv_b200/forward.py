"""Host-side mirror of the reference consumers of the boundary (models.py:11-33, point.py:15-37)."""
import numpy as np

from . import fast_surf as _fs


def cal_forward(in_profile, wavetype="Ray", periods=(5, 10, 20, 40, 60, 80)):
    """Same contract as reference ``models._calForward`` (models.py:11-33): ``in_profile`` is the
    float array [6, n] of rows (h, Vs, Vp, rho, qs, qp); layers with h <= 1e-3 km are dropped; returns
    the phase velocities at ``periods`` or ``None`` when any is < 0.01 (root not found).
    For 'Love' the Love phase velocities are returned (the reference returns the unset Rayleigh
    array there)."""
    if wavetype == "Ray":
        ilvry = 2
    elif wavetype == "Love":
        ilvry = 1
    else:
        raise ValueError("Wrong surface wave type: %s!" % wavetype)
    prof = np.asarray(in_profile, dtype=np.float64)
    ind = np.where(prof[0] > 1e-3)[0]
    h, vs, vp, rho, qs, _qp = prof[:, ind]
    nper = len(periods)
    per = np.zeros(200, dtype=np.float64)
    per[:nper] = periods
    ur0, ul0, cr0, cl0 = _fs.fast_surf(h.size, ilvry, vp, vs, rho, h, 1.0 / qs, per, nper)
    c = cr0 if ilvry == 2 else cl0
    if np.any(c[:nper] < 0.01):
        return None
    return c[:nper].copy()


def misfit(c_obs, c_pred, uncer):
    """Host mirror of ``Point.misfit`` (point.py:15-31) for one model; the batched GPU version is
    ``DispersionSolver.misfit``."""
    if c_pred is None:
        return 88888, 88888, 0
    c_obs = np.ma.masked_array(c_obs) if not np.ma.isMaskedArray(c_obs) else c_obs
    n = c_obs.count()
    chi = (((c_obs - c_pred) / uncer) ** 2).sum()
    mis = np.sqrt(chi / n)
    chi = chi if chi < 50 else np.sqrt(chi * 50.0)
    return mis, chi, np.exp(-0.5 * chi)


def misfit_cascadia(c_obs, c_pred, uncer, periods):
    """Host mirror of ``PointCascadia.misfit`` (point.py:337-366): mean of the two period bands (T <= 40 s, T > 40 s)."""
    if c_pred is None:
        return 88888, 88888, 0
    T = np.asarray(periods, dtype=np.float64)
    c_obs = np.ma.masked_array(c_obs) if not np.ma.isMaskedArray(c_obs) else c_obs
    n = c_obs.count()
    bias = (c_obs - c_pred) / uncer
    b1, b2 = bias[T <= 40], bias[T > 40]
    e1, e2 = (b1.count() == 0), (b2.count() == 0)
    if not e1 and not e2:
        chi = ((b1 ** 2).mean() + (b2 ** 2).mean()) / 2 * n
    elif e1 and not e2:
        chi = (b2 ** 2).mean() * n
    elif not e1 and e2:
        chi = (b1 ** 2).mean() * n
    else:
        raise ValueError("All observations are masked???")
    mis = np.sqrt(chi / n)
    chi = chi if chi < 50 else np.sqrt(chi * 50.0)
    return mis, chi, np.exp(-0.5 * chi)


def accept(chi0, chi1, rnd):
    """Metropolis rule of point.py:34-37 with the uniform draw passed in."""
    if chi1 < chi0:
        return True
    return rnd > 1 - np.exp(-(chi1 - chi0) / 2)
