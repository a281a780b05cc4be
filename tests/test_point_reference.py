"""CPU tier: the callers on either side of the solver against fixtures produced by RUNNING the reference's own
point.py / models.py classes (tests/golden/make_golden_point.py -> point_reference.json): misfit (Point /
PointCascadia), the prior rules of CascadiaContinent and CascadiaOcean, and the merged per-point chain file as
PostPointCascadia reads it.  The GPU kernels are held to the same fixtures in tests/test_gpu_mc.py."""
import json
import os

import numpy as np
import pytest

from oracle import model_builder as MB
from pysurfinv_b200 import forward as F, mc, stack as S

GOLD = os.path.join(os.path.dirname(__file__), "golden", "point_reference.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_host_misfit_mirrors_match_reference(gold):
    for c in gold["misfit"]:
        obs = np.ma.masked_array(c["obs"], mask=c["mask"]) if any(c["mask"]) else np.array(c["obs"])
        pred = None if c["pred"] is None else np.array(c["pred"])
        np.testing.assert_allclose(np.array(F.misfit(obs, pred, np.array(c["uncer"])), dtype=float), c["point"], rtol=1e-12)
        np.testing.assert_allclose(np.array(F.misfit_cascadia(obs, pred, np.array(c["uncer"]), c["T"]), dtype=float),
                                   c["cascadia"], rtol=1e-12)


def test_prior_restatements_match_reference_verdicts(gold):
    for c in gold["priors"]["continent"]:
        t = S.StackTemplate(c["setting"])
        assert ((MB.priors(t, np.zeros(0)) & S.PRIOR_CONTINENT) == 0) == c["isgood"]
    fired = 0
    for c in gold["priors"]["ocean"]:
        t = S.StackTemplate(c["setting"])
        b = MB.priors_ocean(t, np.zeros(0))
        fired |= b
        assert ((b & S.PRIOR_OCEAN) == 0) == c["isgood"]
    # the fixture set exercises every rule that can fire (the first-pair rule cannot: water / constant sediment on top)
    assert fired & (S.PRIOR_OCEAN & ~S.P_FIRSTPAIR) == (S.PRIOR_OCEAN & ~S.P_FIRSTPAIR)


def _synthetic_track(template, n_sub, steps, seed):
    # the generator of make_golden_point.synthetic_track (kept in step with it: the golden values depend on it)
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


def test_merged_point_file_reads_like_postpoint(gold, tmp_path):
    """write_point_npz -> the file layout of Point.MCinvMP (point.py:112-123); the reader logic of
    PostPoint.__init__ (point.py:139-171) restated on the loaded arrays must give what the reference's own
    PostPointCascadia gave on the same file (golden): threshold, accepted count, minimum and average model."""
    g = gold["postpoint"]
    setting = gold["ocean_setting"]
    t = S.StackTemplate(setting)
    tr = _synthetic_track(t, g["n_sub"], g["steps"], g["seed"])
    obs = {"T": gold["periods"], "c": np.linspace(3.57, 3.90, len(gold["periods"])), "uncer": np.full(len(gold["periods"]), 0.01)}
    path = mc.write_point_npz(str(tmp_path / "-127.0_46.0.npz"), tr, setting, obs, "-127.0_46.0", tr.shape[1])
    f = np.load(path, allow_pickle=True)
    MC, st, ob, meta = f["mcTrack"], f["setting"][()], f["obs"][()], f["invMeta"][()]
    assert MC.shape == (g["N"], 3 + t.nparams) and meta == {"pid": "-127.0_46.0", "chainL": g["steps"]}
    assert list(st.keys()) == ["OceanWater", "OceanSedimentCascadia", "OceanCrust", "OceanMantle", "Info"]
    assert st["OceanWater"] == {"H": 2.5, "Vs": 0} and st["OceanSedimentCascadia"]["H"] == [1.0, 0.0, 2.0, 0.1]
    mis, acc, par = MC[:, 0], MC[:, 2], MC[:, 3:].copy()
    for i in range(len(mis)):              # point.py:153-158
        if acc[i]:
            last = i
        else:
            par[i] = par[last]
    np.testing.assert_allclose(par[7], g["mcparas_row7"], rtol=0, atol=1e-12)
    imin = np.nanargmin(mis)
    np.testing.assert_allclose(par[imin], g["min_params"], atol=1e-12)
    thres = max(mis[imin] * 2, mis[imin] + 0.5)        # point.py:307-309
    assert abs(thres - g["thres"]) < 1e-12 and int((mis < thres).sum()) == g["acc_final"]
    avg = par[mis < thres].mean(axis=0)
    np.testing.assert_allclose(avg, g["avg_params"], atol=1e-12)
    # the average model's prediction and misfit (point.py:170-171) through the model-assembly restatement + the oracle
    from oracle import oracle as O
    h, vs, vp, rho, qs = MB.build_one(t, avg)
    k = h > 1e-3
    f32 = lambda x: np.asarray(x, np.float32).astype(np.float64)
    r = O.forward(2, f32(vp[k]), f32(vs[k]), f32(rho[k]), f32(h[k]), f32(1.0 / qs[k]), f32(gold["periods"]), opts=O.make_opts(precision=0))
    np.testing.assert_allclose(r["c"][0], g["avg_pred"], atol=2e-6)
    m = F.misfit_cascadia(obs["c"], np.array(g["avg_pred"]), obs["uncer"], gold["periods"])
    assert abs(m[0] - g["avg_misfit"]) < 1e-9 and abs(m[2] - g["avg_L"]) < 1e-12


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree only exists in the build container")
def test_reference_postpoint_loads_our_file(gold, tmp_path):
    """The reference's own PostPointCascadia (imported with the plotting modules stubbed) reads a file written by
    write_point_npz (build container only; the GPU box has no reference tree)."""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import make_golden_point as G; b, l, u, m, p = G.import_point(); "
            "import numpy as np; from pysurfinv_b200 import mc, stack; t = stack.StackTemplate(G.OCEAN_SETTING); "
            "tr = G.synthetic_track(t); obs = {'T': G.PERIODS, 'c': np.linspace(3.57, 3.90, 18), 'uncer': np.full(18, 0.01)}; "
            "path = mc.write_point_npz(%r, tr, G.OCEAN_SETTING, obs, 'x', tr.shape[1]); pp = p.PostPointCascadia(path); "
            "print('THRES %%.12f %%d' %% (pp.thres, pp.accFinal.sum()))") % (os.path.join(os.path.dirname(__file__), "golden"),
                                                                          str(tmp_path / "x.npz"))
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("THRES")][0].split()
    assert abs(float(line[1]) - gold["postpoint"]["thres"]) < 1e-9 and int(line[2]) == gold["postpoint"]["acc_final"]


def test_thermal_mantle_restatement_matches_reference_classes():
    """oracle.model_builder.hybrid_mantle (HSCM + OceanSeisRitz + OceanSeisRuan + not-a-knot spline) against stacks the
    reference's own OceanMantleHybrid produced (tests/golden/make_golden_thermal.py), and the ocean prior verdicts."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "thermal_reference.json")) as f:
        gold = json.load(f)
    for c in gold["stacks"]:
        t = S.StackTemplate(c["setting"])
        h, vs, vp, rho, qs = MB.build_one(t, np.zeros(0))
        assert len(h) == len(c["h"])
        for mine, key in ((h, "h"), (vs, "vs"), (vp, "vp"), (rho, "rho"), (qs, "qs")):
            np.testing.assert_allclose(mine, np.array(c[key]), rtol=1e-12, atol=1e-12)
        assert ((MB.priors_ocean(t, np.zeros(0)) & S.PRIOR_OCEAN) == 0) == c["isgood"]
