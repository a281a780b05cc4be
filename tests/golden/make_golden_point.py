"""Regenerates tests/golden/point_reference.json by RUNNING the reference's own point.py / models.py classes in the
build container (third-party plotting / geodesy modules stubbed like in make_golden_layers.py):

* misfit      -- Point.misfit (point.py:15-31) and PointCascadia.misfit (point.py:337-366) on given predicted curves
                 (the model's forward() is replaced by a function returning the stored prediction: the arithmetic
                 that produces the fixtures is the reference's);
* priors      -- CascadiaContinent.isgood (models.py:385-523) and CascadiaOcean.isgood (models.py:571-677) verdicts on
                 perturbed settings.  CascadiaOcean.isgood calls scipy.signal.cwt / scipy.signal.ricker, which SciPy
                 removed in 1.15 (this image has 1.18): the two functions are restated here from SciPy's published
                 1.14 implementation (scipy/signal/_wavelets.py) and injected into scipy.signal before the call;
* postpoint   -- a merged per-point chain file written by pysurfinv_b200.mc.write_point_npz is fed through
                 PostPointCascadia.__init__ (point.py:139-171): Markov-chain fill, minimum-misfit model, acceptance
                 threshold, average model and its misfit.  pySurfInv.fast_surf is served by the CPU oracle here
                 (test infrastructure; no Fortran compiler exists in the image).
Build container only (needs /root/reference).
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "point_reference.json")
REF = "/root/reference"

OCEAN_SETTING = {
    "OceanWater": {"H": 2.5},
    "OceanSedimentCascadia": {"H": [1.0, "rel_pos", 100, 0.1]},
    "OceanCrust": {"H": 7, "Vs": [3.25, 3.94]},
    "OceanMantle": {"BottomDepth": 200, "Vs": [[4.35, "abs", 0.3, 0.02], [4.15, "abs", 0.3, 0.02], [4.25, "abs", 0.3, 0.02],
                                                [4.45, "abs", 0.3, 0.02], [4.6, "abs", 0.3, 0.02]]},
    "Info": {"modelType": "CascadiaOcean", "refLayer": True, "topo": -2.5},
}
PERIODS = [10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80]          # point.py:400


def ricker(points, a):
    """scipy.signal.ricker as published in SciPy 1.14 (_wavelets.py): Mexican-hat wavelet."""
    A = 2 / (np.sqrt(3 * a) * (np.pi ** 0.25))
    wsq = a ** 2
    vec = np.arange(0, points) - (points - 1.0) / 2
    xsq = vec ** 2
    mod = (1 - xsq / wsq)
    gauss = np.exp(-xsq / (2 * wsq))
    return A * mod * gauss


def cwt(data, wavelet, widths, dtype=None, **kwargs):
    """scipy.signal.cwt as published in SciPy 1.14 (_wavelets.py)."""
    if dtype is None:
        dtype = np.complex128 if np.asarray(wavelet(1, widths[0], **kwargs)).dtype.char in "FDG" else np.float64
    output = np.empty((len(widths), len(data)), dtype=dtype)
    for ind, width in enumerate(widths):
        N = np.min([10 * width, len(data)])
        wavelet_data = np.conj(wavelet(N, width, **kwargs)[::-1])
        output[ind] = np.convolve(data, wavelet_data, mode="same")
    return output


def import_point():
    import make_golden_layers as G
    brown, layers, utils, models = G.import_reference()
    sys.modules["Triforce.obspyPlus"].randString = lambda n: "x" * n
    import scipy.signal
    if not hasattr(scipy.signal, "cwt"):
        scipy.signal.cwt, scipy.signal.ricker = cwt, ricker
    # CascadiaOcean.isgood (models.py:586) calls np.where on a Python bool (it compares two LISTS, `grp` is not an
    # array there): NumPy 1.x -- the reference's era -- answered with atleast_1d(cond).nonzero() (a deprecation
    # warning since 1.17), NumPy 2 raises.  The NumPy 1.x behaviour is restored for 0-d conditions.
    _where = np.where

    def where_np1(cond, *args):
        if not args and np.ndim(cond) == 0:
            return np.atleast_1d(cond).nonzero()
        return _where(cond, *args)
    np.where = where_np1
    # pySurfInv.fast_surf served by the CPU oracle (float32 fast_surf semantics), same signature as fast_surf.pyf:6-19
    from oracle import oracle as O
    fs = types.ModuleType("pySurfInv.fast_surf")

    def fast_surf(nlay, kind, vp, vs, rho, h, qsinv, per, nper):
        f32 = lambda x: np.asarray(x, dtype=np.float32).astype(np.float64)
        r = O.forward(kind, f32(vp), f32(vs), f32(rho), f32(h), f32(qsinv), f32(per[:nper]), opts=O.make_opts(precision=0))
        out = [np.zeros(200, np.float32) for _ in range(4)]      # ur0, ul0, cr0, cl0
        nf = int(r["imax"][0])
        (out[2] if kind == 2 else out[3])[:nf] = r["c"][0][:nf]
        (out[0] if kind == 2 else out[1])[:nf] = r["u"][0][:nf]
        return tuple(out)
    fs.fast_surf = fast_surf
    sys.modules["pySurfInv.fast_surf"] = fs
    sys.modules["pySurfInv"].fast_surf = fs
    point = G._load("pySurfInv.point", os.path.join(REF, "point.py"))
    return brown, layers, utils, models, point


class _FixedForward:
    def __init__(self, cp):
        self.cp = cp

    def forward(self, periods=None):
        return self.cp


def misfit_cases(point):
    rng = np.random.default_rng(11)
    T = np.array(PERIODS, float)
    out = []
    for case in range(12):
        K = len(T)
        cO = 3.5 + 0.4 * (T - 10) / 70 + rng.normal(0, 0.01, K)
        unc = rng.uniform(0.005, 0.016, K)
        scale = [0.3, 1.0, 3.0, 10.0][case % 4]           # small chi, chi near 50, soft-clipped chi
        cP = cO + scale * unc * rng.normal(0, 1, K)
        mask = np.zeros(K, bool)
        if case % 3 == 1:
            mask[rng.choice(K, 4, replace=False)] = True
        if case == 10:
            mask[T <= 40] = True                           # only the long band left (point.py:352-353)
        if case == 11:
            mask[T > 40] = True
        fail = (case == 9)
        obs_c = np.ma.masked_array(cO, mask=mask) if mask.any() else cO
        res = {}
        for cls, name in ((point.Point, "point"), (point.PointCascadia, "cascadia")):
            p = object.__new__(cls)
            p.obs = {"T": list(T), "c": obs_c, "uncer": unc}
            m = p.misfit(model=_FixedForward(None if fail else cP))
            res[name] = [float(x) for x in m]
        out.append({"T": T.tolist(), "obs": cO.tolist(), "mask": mask.tolist(), "uncer": unc.tolist(),
                    "pred": None if fail else cP.tolist(), **res})
    return out


def _fixed(setting, values):
    """setting with the free parameters replaced by `values` (all fixed)."""
    from pysurfinv_b200.stack import _is_spec
    it = iter(values)
    s2 = {}
    for name, parm in setting.items():
        if name == "Info":
            s2[name] = dict(parm); continue
        d = {}
        for k, v in parm.items():
            if _is_spec(v):
                d[k] = [float(next(it)), "fixed"] if v[1] not in ("fixed", "total") else v
            elif isinstance(v, list):
                d[k] = [([float(next(it)), "fixed"] if (_is_spec(x) and x[1] not in ("fixed", "total")) else
                         (x if _is_spec(x) else [x, "fixed"])) for x in v]
            else:
                d[k] = v
        s2[name] = d
    return s2


def prior_cases(models):
    rng = np.random.default_rng(23)
    out = {"continent": [], "ocean": []}
    base = {"Sediment": {"H": 2.0, "Vs": [1.2, 2.0]}, "Crust": {"H": 30.0, "Vs": [3.3, 3.5, 3.7, 3.9]},
            "Mantle": {"BottomDepth": 200.0, "Vs": [4.4, 4.3, 4.5, 4.4, 4.6]}, "Info": {"modelType": "MCInv"}}
    fix = lambda v: [[float(x), "fixed"] for x in v]
    for _ in range(60):
        s = json.loads(json.dumps(base))
        s["Sediment"]["H"] = float(rng.uniform(0.5, 4.0))
        s["Sediment"]["Vs"] = fix(rng.uniform(0.9, 2.6, 2))
        s["Crust"]["H"] = float(rng.uniform(18.0, 45.0))
        s["Crust"]["Vs"] = fix(np.sort(rng.uniform(3.1, 4.1, 4)) + rng.normal(0, 0.08, 4))
        s["Mantle"]["Vs"] = fix(rng.uniform(4.0, 5.0, 5))
        mod = models.buildModel1D(s)
        out["continent"].append({"setting": s, "isgood": bool(models.CascadiaContinent.isgood(mod))})
    for _ in range(200):
        s = json.loads(json.dumps(OCEAN_SETTING))
        s["OceanWater"]["H"] = float(rng.uniform(0.5, 4.0)); s["Info"]["topo"] = -s["OceanWater"]["H"]
        s["OceanSedimentCascadia"]["H"] = float(rng.uniform(0.02, 2.0))
        s["OceanCrust"]["Vs"] = fix(np.sort(rng.uniform(3.0, 4.1, 2)) if rng.random() < 0.8 else rng.uniform(3.0, 4.1, 2))
        # smooth, mostly monotone mantle profiles with the occasional bump (the hybrid-parameterisation rules of
        # models.py:613-637 reject most random draws)
        v0 = rng.uniform(4.0, 4.5); dv = np.cumsum(rng.uniform(-0.02, 0.12, 4) * (1 if rng.random() < 0.7 else rng.choice([-1, 1], 4)))
        coef = np.concatenate([[v0], v0 + dv])
        if _ >= 120:       # strongly oscillating profiles: the oscillation and wavelet rules (models.py:603-611, 627-635)
            coef = rng.uniform(3.6, 4.9, 5 + int(rng.integers(0, 3)))
        s["OceanMantle"]["Vs"] = fix(coef)
        mod = models.buildModel1D(s)
        out["ocean"].append({"setting": s, "isgood": bool(mod.isgood())})
    return out


def synthetic_track(template, n_sub=4, steps=30, seed=5):
    """Deterministic stand-in for a GPU-produced ensemble track [n_sub, steps, 3 + P]."""
    rng = np.random.default_rng(seed)
    lo, hi, st = template.bounds()
    P = template.nparams
    tr = np.zeros((n_sub, steps, 3 + P))
    for i in range(n_sub):
        cur = template.start_values().astype(float)
        for k in range(steps):
            prop = np.clip(cur + rng.normal(0, 1, P) * st, lo + 1e-6, hi - 1e-6)
            mis = float(rng.uniform(0.6, 3.0))
            acc = 1.0 if (k == 0 or rng.random() < 0.4) else 0.0
            tr[i, k] = np.concatenate([[mis, np.exp(-0.5 * mis * mis * 18), acc], prop])
            if acc:
                cur = prop
    return tr


def postpoint_case(point):
    import tempfile
    from pysurfinv_b200 import mc, stack
    t = stack.StackTemplate(OCEAN_SETTING)
    tr = synthetic_track(t)
    obs = {"T": PERIODS, "c": np.linspace(3.57, 3.90, len(PERIODS)), "uncer": np.full(len(PERIODS), 0.01)}
    with tempfile.TemporaryDirectory() as d:
        path = mc.write_point_npz(os.path.join(d, "-127.0_46.0.npz"), tr, OCEAN_SETTING, obs, "-127.0_46.0", tr.shape[1])
        pp = point.PostPointCascadia(path)
    return {"n_sub": int(tr.shape[0]), "steps": int(tr.shape[1]), "seed": 5, "N": int(pp.N), "thres": float(pp.thres),
            "acc_final": int(pp.accFinal.sum()), "min_params": [float(x) for x in pp.minMod._brownians()],
            "avg_params": [float(x) for x in pp.avgMod._brownians()], "avg_misfit": float(pp.avgMod.misfit),
            "avg_L": float(pp.avgMod.L), "min_misfit": float(pp.minMod.misfit),
            "mcparas_row7": [float(x) for x in pp.MCparas[7]], "invMeta": {k: (v if isinstance(v, str) else int(v)) for k, v in pp.invMeta.items()},
            "avg_pred": [float(x) for x in pp.avgMod.forward(PERIODS)]}


def main():
    brown, layers, utils, models, point = import_point()
    out = {"misfit": misfit_cases(point), "priors": prior_cases(models), "postpoint": postpoint_case(point),
           "ocean_setting": OCEAN_SETTING, "periods": PERIODS}
    with open(OUT, "w") as f:
        json.dump(out, f)
    pr = out["priors"]
    print("wrote", OUT, len(out["misfit"]), "misfit cases;", sum(c["isgood"] for c in pr["continent"]), "/", len(pr["continent"]),
          "continent good;", sum(c["isgood"] for c in pr["ocean"]), "/", len(pr["ocean"]), "ocean good; postpoint thres",
          out["postpoint"]["thres"], "acc", out["postpoint"]["acc_final"])


if __name__ == "__main__":
    main()
