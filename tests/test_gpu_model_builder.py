"""GPU tier: the device-side model builder (surfdisp_build_stacks) against the numpy restatement of the
reference's model assembly (oracle/model_builder.py, itself pinned to fixtures produced by the reference's own
classes, tests/golden/layers_reference.json) -- and the whole chain parameters -> stacks -> dispersion."""
import json
import os

import numpy as np
import pytest

from oracle import model_builder as MB
from oracle import oracle as O
from pysurfinv_b200 import stack as S
from pysurfinv_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "layers_reference.json")

CONTINENTAL = {"Sediment": {"H": [2.0, "abs_pos", 1.5, 0.1], "Vs": [1.5, 0.8, 2.6, 0.05]},
               "Crust": {"H": [30.0, "abs", 12.0, 1.0], "Vs": [[3.3, "rel", 10, 0.02], [3.5, "rel", 10, 0.02],
                                                             [3.7, "rel", 10, 0.02], [3.9, "rel", 10, 0.02]]},
               "Mantle": {"BottomDepth": 200.0, "Vs": [[4.4, "abs", 0.3, 0.02], [4.3, "abs", 0.3, 0.02],
                                                      [4.5, "abs", 0.3, 0.02], [4.4, "abs", 0.3, 0.02],
                                                      [4.6, "abs", 0.3, 0.02]]},
               "Info": {"refLayer": True}}
OCEANIC = {"OceanWater": {"H": [2.7, "abs_pos", 1.0, 0.1]}, "OceanSedimentCascadia": {"H": [0.35, 0.05, 1.0, 0.05]},
           "OceanCrust": {"H": [7.0, "abs", 2.5, 0.2], "Vs": [[3.25, "rel", 10, 0.02], [3.94, "rel", 10, 0.02]]},
           "OceanMantle": {"BottomDepth": 200.0, "Vs": [[4.4, "abs", 0.3, 0.02], [4.2, "abs", 0.3, 0.02],
                                                         [4.1, "abs", 0.3, 0.02], [4.3, "abs", 0.3, 0.02]]},
           "Info": {"refLayer": True, "topo": -2.7}}


@pytest.fixture(scope="module")
def solver():
    import torch
    from pysurfinv_b200 import api
    assert torch.cuda.is_available()
    return api.DispersionSolver("cuda:0")


def _random_params(t, M, seed):
    lo, hi, _ = t.bounds()
    rng = np.random.default_rng(seed)
    return (lo + (hi - lo) * rng.random((M, t.nparams))).astype(np.float32)


def _compare(solver, t, params, lmax):
    import torch
    lay_d, nl_d = solver.build_stacks(t, torch.from_numpy(params).cuda(), lmax=lmax)
    torch.cuda.synchronize()
    lay, nl = lay_d.cpu().numpy(), nl_d.cpu().numpy()
    ref, nref = MB.build_stacks(t, params, lmax)
    assert np.array_equal(nl, nref)
    # float32 roundings of identical float64 arithmetic: at most one ulp apart
    np.testing.assert_allclose(lay, ref, rtol=2e-7, atol=1e-7)
    beyond = np.arange(lmax)[None, :] >= nl[:, None]
    assert np.all(lay[:, beyond] == 0)
    return lay_d, nl_d, lay, nl


@pytest.mark.parametrize("setting,lmax", [(CONTINENTAL, 142), (OCEANIC, 96)])
def test_builder_matches_model_assembly_oracle(solver, setting, lmax):
    t = S.StackTemplate(setting)
    params = _random_params(t, 4096, seed=5)
    _compare(solver, t, params, lmax)


def test_builder_reproduces_reference_fixtures(solver):
    """The fixtures come from running the reference's own layer classes (make_golden_layers.py)."""
    import torch
    with open(GOLD) as f:
        gold = json.load(f)
    for st in gold["stacks"]:
        t = S.StackTemplate(st["setting"])
        assert t.nparams == 0
        lay_d, nl_d = solver.build_stacks(t, torch.zeros((1, 0), dtype=torch.float32, device="cuda"), lmax=160)
        lay, n = lay_d.cpu().numpy()[:, 0], int(nl_d.cpu().numpy()[0])
        h = np.array(st["h"]); keep = h > 1e-3
        assert n == int(keep.sum())
        for row, key in ((1, "vs"), (0, "vp"), (2, "rho"), (3, "h")):
            np.testing.assert_allclose(lay[row, :n], np.array(st[key])[keep], rtol=2e-7, atol=1e-7)
        np.testing.assert_allclose(lay[4, :n], 1.0 / np.array(st["qs"])[keep], rtol=2e-7, atol=1e-12)


def test_overflow_and_bad_arguments(solver):
    import torch
    from pysurfinv_b200 import api
    t = S.StackTemplate(CONTINENTAL)
    params = torch.from_numpy(_random_params(t, 8, seed=1)).cuda()
    lay, nl = solver.build_stacks(t, params, lmax=40)      # needs 96 layers: reported as nlay = 0, no overrun
    assert int(nl.abs().sum()) == 0
    with pytest.raises(ValueError):
        solver.build_stacks(t, params[:, :3].contiguous())
    bad = t.to_c(); bad.ngroups = 0
    rc = solver.lib.surfdisp_build_stacks(bad, 8, params.data_ptr(), 96, lay.data_ptr(), nl.data_ptr(), None)
    assert rc == -1
    assert isinstance(api.SurfdispError("x"), RuntimeError)


def test_parameters_to_dispersion_chain(solver):
    """MC-style use: parameter vectors -> device-built stacks -> phase velocities, against the CPU chain
    (model-assembly oracle -> dispersion oracle)."""
    t = S.StackTemplate(CONTINENTAL)
    params = _random_params(t, 512, seed=9)
    lay_d, nl_d, lay, nl = _compare(solver, t, params, 142)
    per = np.array([8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
    g = solver.forward(lay_d, nl_d, per, kind=2)
    c0, u0, nf0, st0 = O.forward_batch(2, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
    ok = st0 != 3
    assert np.array_equal(g["nfound"].cpu().numpy()[ok], nf0[ok])
    assert np.abs(g["c"].cpu().numpy() - c0)[ok].max() <= 1e-4


THERMAL = {"OceanWater": {"H": 2.5}, "OceanSedimentCascadia": {"H": [1, "rel_pos", 100, 0.1]}, "OceanCrust": {"H": 7, "Vs": [3.25, 3.94]},
           "OceanMantleHybrid": {"BottomDepth": 200, "Conversion": "Ritzwoller", "ThermAge": [4, "rel_pos", 200, 0.4],
                                 "Vs": [[0, "abs", 0.4, 0.01], [0, "abs", 0.4, 0.01], [0, "abs", 0.4, 0.01], [0, "abs", 0.2, 0.01]]},
           "Info": {"modelType": "CascadiaOcean", "period": 10, "refLayer": True, "lithoAgeQ": True, "lithoAge": 0.6, "topo": -2.5}}


def test_thermal_mantle_group_matches_reference_classes(solver):
    """the thermal parameterisation (OceanMantleHybrid, layers.py:297-363: half-space cooling -> Vs by
    OceanSeisRitz, B-spline perturbation below the melting depth joined by a not-a-knot spline, Qs by OceanSeisRuan)
    assembled on the device: against stacks produced by the reference's own classes (thermal_reference.json), against
    the numpy restatement on random parameter vectors of the config-1 setting of point.py:374-391, and the
    CascadiaOcean prior verdicts on those thermal profiles."""
    import torch
    with open(os.path.join(os.path.dirname(__file__), "golden", "thermal_reference.json")) as f:
        gold = json.load(f)
    none = torch.zeros((1, 0), dtype=torch.float32, device="cuda")
    for c in gold["stacks"]:
        t = S.StackTemplate(c["setting"])
        lay, nl = solver.build_stacks(t, none, lmax=t.max_layers())
        n = int(nl[0])
        assert n == len(c["h"])
        got = lay[:, 0, :n].cpu().numpy()
        for row, key in ((0, "vp"), (1, "vs"), (2, "rho"), (3, "h")):
            np.testing.assert_allclose(got[row], np.array(c[key]), rtol=3e-7, atol=1e-7)
        np.testing.assert_allclose(1.0 / got[4], np.array(c["qs"]), rtol=1e-6)
        bad = int(solver.check_priors(t, none).cpu()[0]) & S.PRIOR_OCEAN
        assert (bad == 0) == c["isgood"]
    t = S.StackTemplate(THERMAL, prior_mask=S.PRIOR_OCEAN)
    assert t.nparams == 6 and t.max_layers() == 86
    params = _random_params(t, 400, 17)
    _compare(solver, t, params, t.max_layers())
    got = solver.check_priors(t, torch.from_numpy(params).cuda()).cpu().numpy() & S.PRIOR_OCEAN
    want = np.array([MB.priors_ocean(t, p.astype(np.float64)) for p in params]) & S.PRIOR_OCEAN
    assert (got != want).mean() < 0.01 and 0 < (got == 0).sum() < len(got)
    # the whole chain on the real config-1 parameterisation: parameters -> stacks -> dispersion, against the oracle
    per = np.array([10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
    good = params[got == 0][:64]
    lay, nl = solver.build_stacks(t, torch.from_numpy(good).cuda())
    g = solver.forward(lay, nl, per, kind=2)
    c0, u0, nf0, st0 = O.forward_batch(2, lay.cpu().numpy(), nl.cpu().numpy(), per, opts=O.make_opts(precision=0), nthreads=8)
    assert np.array_equal(g["nfound"].cpu().numpy(), nf0) and np.abs(g["c"].cpu().numpy() - c0).max() <= 1e-4
