"""GPU tier: prior checks, proposals, Metropolis rule and the ensemble driver."""
import json
import os

import numpy as np
import pytest

from oracle import model_builder as MB
from pysurfinv_b200 import stack as S

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "layers_reference.json")
SETTING = {"Sediment": {"H": [2.0, "abs_pos", 1.5, 0.1], "Vs": [[1.2, 0.8, 2.0, 0.05], [2.0, 1.2, 2.8, 0.05]]},
           "Crust": {"H": [30.0, "abs", 12.0, 1.0], "Vs": [[3.3, "rel", 10, 0.02], [3.5, "rel", 10, 0.02],
                                                         [3.7, "rel", 10, 0.02], [3.9, "rel", 10, 0.02]]},
           "Mantle": {"BottomDepth": 200.0, "Vs": [[4.4, "abs", 0.3, 0.02], [4.3, "abs", 0.3, 0.02],
                                                  [4.5, "abs", 0.3, 0.02], [4.4, "abs", 0.3, 0.02],
                                                  [4.6, "abs", 0.3, 0.02]]},
           "Info": {"refLayer": True}}


@pytest.fixture(scope="module")
def solver():
    import torch
    from pysurfinv_b200 import api
    assert torch.cuda.is_available()
    return api.DispersionSolver("cuda:0")


def test_priors_match_oracle_and_reference(solver):
    import torch
    t = S.StackTemplate(SETTING)
    lo, hi, _ = t.bounds()
    rng = np.random.default_rng(3)
    params = (lo + (hi - lo) * rng.random((2000, t.nparams))).astype(np.float32)
    got = solver.check_priors(t, torch.from_numpy(params).cuda()).cpu().numpy() & S.P_ALL   # (the ocean rules are evaluated too)
    want = np.array([MB.priors(t, p.astype(np.float64)) for p in params])
    assert np.array_equal(got, want)
    assert 0 < (got == 0).sum() < len(got)
    with open(GOLD) as f:
        gold = json.load(f)
    for case in gold["priors"]:          # the reference's own CascadiaPrism.isgood
        tt = S.StackTemplate(case["setting"])
        bad = int(solver.check_priors(tt, torch.zeros((1, 0), dtype=torch.float32, device="cuda")).cpu()[0]) & S.PRIOR_PRISM
        assert (bad == 0) == case["isgood"]


def test_proposals_respect_bounds_priors_and_are_reproducible(solver):
    import torch
    t = S.StackTemplate(SETTING, prior_mask=S.P_ALL)
    lo, hi, st = t.bounds()
    M = 20000
    cur = torch.from_numpy(np.tile(t.start_values(), (M, 1))).cuda().contiguous()
    status = torch.empty(M, dtype=torch.int32, device="cuda")
    a = solver.mc_propose(t, cur, seed=11, step_index=5, status=status)
    b = solver.mc_propose(t, cur, seed=11, step_index=5)
    c = solver.mc_propose(t, cur, seed=11, step_index=6)
    assert torch.equal(a, b) and not torch.equal(a, c)            # counter-based generator
    an = a.cpu().numpy()
    assert np.all(an > lo[None, :]) and np.all(an < hi[None, :])   # brownian.py:22 (strict)
    assert int(((solver.check_priors(t, a) & S.P_ALL) != 0).sum()) == 0        # every proposal is admissible
    assert int((status < 1).sum()) == 0
    # a parameter whose Gaussian step never hits a bound (wide interval) keeps its N(v, step) distribution unless
    # the prior rejection reshapes it: without priors the moments must match
    t0 = S.StackTemplate(SETTING, prior_mask=0)
    p0 = solver.mc_propose(t0, cur, seed=3, step_index=1).cpu().numpy()
    d = (p0 - t0.start_values()[None, :]) / st[None, :]
    j = int(np.argmax((hi - lo) / st))                             # widest interval in units of its step
    assert abs(d[:, j].mean()) < 0.03 and abs(d[:, j].std() - 1.0) < 0.03
    # chain restart: uniform over the box (brownian.py:17-19)
    rs = torch.ones(M, dtype=torch.uint8, device="cuda")
    r = solver.mc_propose(t0, cur, seed=3, step_index=2, reset_mask=rs).cpu().numpy()
    u = (r - lo[None, :]) / (hi - lo)[None, :]
    assert np.all((u > 0) & (u < 1)) and abs(u.mean() - 0.5) < 0.01 and abs(u.std() - 12 ** -0.5) < 0.01


def test_metropolis_rule(solver):
    import torch
    M, P = 200000, 3
    chi0 = torch.full((M,), 10.0, device="cuda")
    chi1 = torch.full((M,), 12.0, device="cuda"); chi1[: M // 2] = 9.0
    cur = torch.zeros((M, P), device="cuda"); prop = torch.ones((M, P), device="cuda")
    acc = solver.mc_accept(chi1, prop, chi0, cur, seed=5, step_index=0).cpu().numpy()
    assert acc[: M // 2].all()                                      # chi1 < chi0: always (point.py:35-36)
    rate = acc[M // 2:].mean()                                      # else with probability exp(-(chi1-chi0)/2)
    assert abs(rate - np.exp(-1.0)) < 0.005
    c = cur.cpu().numpy(); x = chi0.cpu().numpy()
    assert np.all(c[acc == 1] == 1) and np.all(c[acc == 0] == 0)
    assert np.all(x[acc == 1] == chi1.cpu().numpy()[acc == 1]) and np.all(x[acc == 0] == 10.0)
    force = torch.ones(M, dtype=torch.uint8, device="cuda")
    chi1b = torch.full((M,), 1e6, device="cuda")
    assert solver.mc_accept(chi1b, prop, chi0, cur, seed=5, step_index=1, force_mask=force).cpu().numpy().all()


def test_chain_ensemble_recovers_a_model_and_writes_reference_layout(solver, tmp_path):
    import torch
    from pysurfinv_b200 import mc
    t = S.StackTemplate(SETTING, prior_mask=S.P_ALL)
    per = np.array([8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
    # synthetic observation: the dispersion curve of an admissible model inside the box
    start = torch.from_numpy(t.start_values()[None, :]).cuda().contiguous()
    truth = solver.mc_propose(t, start, seed=99, step_index=0, reset_mask=torch.ones(1, dtype=torch.uint8, device="cuda"))
    lay, nl = solver.build_stacks(t, truth)
    obs = solver.forward(lay, nl, per, kind=2)["c"][0].cpu().numpy()
    assert obs.min() > 1.0
    ens = mc.ChainEnsemble(solver, t, per, obs, np.full(len(per), 0.01, np.float32), n_chains=256, seed=1, chain_length=150,
                           track_steps=150)
    ens.run(150)
    tr = ens.mc_track()
    assert tr.shape == (256, 150, 3 + t.nparams)
    misfit = tr[:, :, 0]
    accepted = tr[:, :, 2]
    assert np.all(accepted[:, 0] == 1)
    # chain 0 starts from the start model itself (point.py:47-50), the others from a uniform redraw (point.py:52)
    assert np.array_equal(tr[0, 0, 3:], t.start_values()) and not np.array_equal(tr[1, 0, 3:], t.start_values())
    # the walk goes downhill: the misfit of the chains' current states (last accepted sample) at the end is
    # well below the misfit of their first samples
    state = np.empty_like(misfit)
    for k in range(misfit.shape[1]):
        state[:, k] = np.where(accepted[:, k] == 1, misfit[:, k], state[:, k - 1] if k else misfit[:, 0])
    assert state[:, -1].mean() < 0.6 * state[:, 0].mean(), (state[:, 0].mean(), state[:, -1].mean())
    assert 0.02 < accepted[:, 1:].mean() < 0.98
    # every recorded proposal is admissible and inside its box
    lo, hi, _ = t.bounds()
    flat = tr[:, :, 3:].reshape(-1, t.nparams).astype(np.float32)
    assert np.all(flat >= lo[None, :]) and np.all(flat <= hi[None, :])
    assert int((solver.check_priors(t, torch.from_numpy(flat[::37].copy()).cuda()) & t.prior_mask).abs().sum()) == 0
    # the graph replay and the eager launches are the same chain: rerun without the graph
    ens2 = mc.ChainEnsemble(solver, t, per, obs, np.full(len(per), 0.01, np.float32), n_chains=256, seed=1, chain_length=150,
                            track_steps=150, use_graph=False)
    ens2.run(40)
    assert np.array_equal(ens2.mc_track(), tr[:, :40])
    # the misfit of the step kernel is the misfit kernel's (and the reference's, tests/test_point_reference.py)
    m3 = solver.misfit(ens.c_pred, ens.nfound, obs, np.full(len(per), 0.01, np.float32)).cpu().numpy()
    np.testing.assert_allclose(ens.misfit3.cpu().numpy(), m3, rtol=1e-6)
    paths = ens.save_npz(str(tmp_path), "pt", SETTING, chain_length=150)
    z = np.load(paths[3], allow_pickle=True)
    assert z["mcTrack"].shape == (150, 3 + t.nparams) and z["invMeta"].item()["chainL"] == 150
    assert set(z["obs"].item().keys()) == {"T", "c", "uncer"} and "Crust" in z["setting"].item()
    merged = np.load(ens.save_point_npz(str(tmp_path), "-120.0_45.0", SETTING, chain_length=150), allow_pickle=True)
    assert merged["mcTrack"].shape == (256 * 150, 3 + t.nparams) and np.array_equal(merged["mcTrack"][150:300], z["mcTrack"] * 0 + tr[1])


def test_restarts_and_points_side_by_side(solver):
    """Sub-chain restarts every chain_length steps (point.py:45-57) and several points in one ensemble, each with its
    own observed curve: a point's chains do not depend on which other points run beside them."""
    import torch
    from pysurfinv_b200 import mc
    t = S.StackTemplate(SETTING, prior_mask=S.PRIOR_CONTINENT)
    per = np.array([10, 14, 20, 28, 40, 60, 80], np.float32)
    start = torch.from_numpy(t.start_values()[None, :]).cuda().contiguous()
    obs = []
    for sd in (5, 6, 7):
        truth = solver.mc_propose(t, start, seed=sd, step_index=0, reset_mask=torch.ones(1, dtype=torch.uint8, device="cuda"))
        lay, nl = solver.build_stacks(t, truth)
        obs.append(solver.forward(lay, nl, per, kind=2)["c"][0].cpu().numpy())
    obs = np.array(obs)
    sig = np.full_like(obs, 0.01)
    ens = mc.ChainEnsemble(solver, t, per, obs, sig, n_chains=16, seed=3, n_points=3, chain_length=10, track_steps=30).run(30)
    tr = ens.mc_track().reshape(3, 16, 30, -1)
    assert np.all(tr[:, :, [0, 10, 20], 2] == 1)                     # first sample of every sub-chain is kept
    assert np.array_equal(tr[1, 0, 0, 3:], t.start_values())         # chain 0 of a point starts from the start model,
    assert not np.array_equal(tr[1, 1, 0, 3:], t.start_values())     # the others from a uniform redraw (point.py:101, 47-52)
    assert not np.array_equal(tr[1, 0, 10, 3:], t.start_values())    # later sub-chains: always a redraw (point.py:52)
    assert ens.best_misfit().shape == (3,)
    # a point's chains do not depend on the observations of the points beside it
    ens_b = mc.ChainEnsemble(solver, t, per, obs[[0, 2, 1]], sig, n_chains=16, seed=3, n_points=3, chain_length=10, track_steps=30).run(30)
    trb = ens_b.mc_track().reshape(3, 16, 30, -1)
    assert np.array_equal(trb[0], tr[0]) and not np.array_equal(trb[1], tr[1])


def test_prior_kernels_match_reference_verdicts(solver):
    """check_priors against the verdicts of the reference's own CascadiaContinent.isgood / CascadiaOcean.isgood
    (tests/golden/point_reference.json) and, bit by bit, against the numpy restatement on random ocean models."""
    import torch
    with open(os.path.join(os.path.dirname(__file__), "golden", "point_reference.json")) as f:
        gold = json.load(f)
    none = torch.zeros((1, 0), dtype=torch.float32, device="cuda")
    for case in gold["priors"]["continent"]:
        bad = int(solver.check_priors(S.StackTemplate(case["setting"]), none).cpu()[0])
        assert ((bad & S.PRIOR_CONTINENT) == 0) == case["isgood"]
    for case in gold["priors"]["ocean"]:
        tt = S.StackTemplate(case["setting"])
        bad = int(solver.check_priors(tt, none).cpu()[0])
        assert ((bad & S.PRIOR_OCEAN) == 0) == case["isgood"]
        assert (bad & S.PRIOR_OCEAN) == (MB.priors_ocean(tt, np.zeros(0)) & S.PRIOR_OCEAN)
    ocean = dict(gold["ocean_setting"])
    t = S.StackTemplate(ocean, prior_mask=S.PRIOR_OCEAN)
    lo, hi, _ = t.bounds()
    rng = np.random.default_rng(9)
    params = (lo + (hi - lo) * rng.random((3000, t.nparams))).astype(np.float32)
    got = solver.check_priors(t, torch.from_numpy(params).cuda()).cpu().numpy() & S.PRIOR_OCEAN
    want = np.array([MB.priors_ocean(t, p.astype(np.float64)) for p in params]) & S.PRIOR_OCEAN
    assert (got != want).mean() < 2e-3          # (threshold ties of the wavelet rule: summation order)
    assert 0 < (got == 0).sum() < len(got)
    # proposals under the ocean rules are admissible
    cur = torch.from_numpy(np.tile(t.start_values(), (4096, 1))).cuda().contiguous()
    status = torch.empty(4096, dtype=torch.int32, device="cuda")
    prop = solver.mc_propose(t, cur, seed=2, step_index=1, status=status)
    assert int((solver.check_priors(t, prop) & S.PRIOR_OCEAN != 0).sum()) == 0 and int((status < 1).sum()) == 0


def test_misfit_kernel_matches_reference_fixtures(solver):
    """surfdisp_misfit_batch against Point.misfit / PointCascadia.misfit of the reference itself (golden)."""
    import torch
    with open(os.path.join(os.path.dirname(__file__), "golden", "point_reference.json")) as f:
        gold = json.load(f)
    for c in gold["misfit"]:
        K = len(c["T"])
        pred = np.zeros((1, K), np.float32) if c["pred"] is None else np.array([c["pred"]], np.float32)
        nf = torch.tensor([0 if c["pred"] is None else K], dtype=torch.int32, device="cuda")
        mask = (~np.array(c["mask"])).astype(np.uint8)
        for mode, key in ((0, "point"), (1, "cascadia")):
            m = solver.misfit(torch.from_numpy(pred).cuda(), nf, c["obs"], c["uncer"], mask=mask, periods=c["T"], mode=mode).cpu().numpy()[0]
            # float32 inputs (obs, prediction, 1/sigma) against the reference's float64: 1e-4 relative on chi-square
            np.testing.assert_allclose(m, c[key], rtol=3e-4, atol=1e-30)
