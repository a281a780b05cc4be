"""CPU tier: the numpy restatement of the reference's model assembly against fixtures produced by running the
reference's own classes (tests/golden/make_golden_layers.py), and the template compiler."""
import json
import os

import numpy as np
import pytest

from oracle import model_builder as MB
from pysurfinv_b200 import stack as S

GOLD = os.path.join(os.path.dirname(__file__), "golden", "layers_reference.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


def test_bspline_bases_match_reference(gold):
    for b in gold["bspline"]:
        ref = np.array(b["basis"])
        mine = MB.bspl_basis(b["N"] + 1, b["nbasis"])
        assert mine.shape == ref.shape
        np.testing.assert_allclose(mine, ref, rtol=0, atol=2e-13)


def test_stacks_match_reference(gold):
    for st in gold["stacks"]:
        t = S.StackTemplate(st["setting"])
        assert t.nparams == 0
        h, vs, vp, rho, qs = MB.build_one(t, np.zeros(0))
        assert len(h) == len(st["h"])
        for mine, key in ((h, "h"), (vs, "vs"), (vp, "vp"), (rho, "rho"), (qs, "qs")):
            np.testing.assert_allclose(mine, np.array(st[key]), rtol=1e-11, atol=1e-11)


def test_template_parameter_order_and_bounds():
    setting = {"Sediment": {"H": [2.0, "abs_pos", 1.5, 0.1], "Vs": [1.5, 0.8, 2.6, 0.05]},
               "Crust": {"H": [30.0, "abs", 10.0, 1.0], "Vs": [[3.3, "rel", 10, 0.02], [3.5, "fixed"], [3.7, "rel", 10, 0.02], 3.9]},
               "Mantle": {"BottomDepth": 200.0, "Vs": [[4.4, "abs", 0.3, 0.02], [4.3, "abs", 0.3, 0.02], [4.5, "abs", 0.3, 5.0]]},
               "Info": {"refLayer": True}}
    t = S.StackTemplate(setting)
    where = [p.where for p in t.params]
    assert where == ["Sediment.H", "Sediment.Vs[0]", "Crust.H", "Crust.Vs[0]", "Crust.Vs[2]", "Mantle.Vs[0]", "Mantle.Vs[1]", "Mantle.Vs[2]"]
    lo, hi, step = t.bounds()
    assert np.allclose(lo[:3], [0.5, 0.8, 20.0]) and np.allclose(hi[:3], [3.5, 2.6, 40.0])
    assert np.isclose(lo[3], 3.3 * 0.9) and np.isclose(step[7], 0.3)      # step clamped to half the range (brownian.py:8)
    assert len(t.groups) == 4 and t.groups[3].kind == S.G_REFMANTLE and t.max_layers() == 1 + 15 + 60 + 20 and t.max_layers(hi=t.bounds()[1] + 30) == 1 + 30 + 60 + 20
    c = t.to_c()
    assert c.ngroups == 4 and c.nparams == 8 and c.groups[1].v_param[1] == -1 and abs(c.groups[1].v_fixed[3] - 3.9) < 1e-12
    lay, nl = MB.build_stacks(t, t.start_values()[None, :], 96)
    assert nl[0] == 1 + 15 + 60 + 20 and abs(lay[3, 0, :nl[0]].sum() - 500.0) < 1e-3


def test_priors_match_reference_isgood(gold):
    """The numpy restatement of the prior rules against the reference's own CascadiaPrism.isgood, run on 60
    perturbed models by make_golden_layers.py."""
    n_good = 0
    for case in gold["priors"]:
        t = S.StackTemplate(case["setting"])
        bad = MB.priors(t, np.zeros(0))
        assert (bad == 0) == case["isgood"], (case["setting"], bad)
        n_good += case["isgood"]
    assert 0 < n_good < len(gold["priors"])


def test_layer_bound_follows_the_parameter_boxes():
    """StackTemplate.max_layers(lo, hi): the fine-layer rules grow with the group thickness, so the bound is taken at the
    largest thickness inside the boxes -- per-point boxes ([n_points, P]) count with their union -- and every model
    drawn from the boxes stays within it."""
    setting = {"Sediment": {"H": [2.0, "abs_pos", 1.5, 0.1], "Vs": [1.5, 0.8, 2.6, 0.05]},
               "Crust": {"H": [30.0, "abs", 10.0, 1.0], "Vs": [[3.3, "rel", 10, 0.02], [3.5, "rel", 10, 0.02], [3.7, "rel", 10, 0.02]]},
               "Mantle": {"BottomDepth": 200.0, "Vs": [[4.4, "abs", 0.3, 0.02], [4.3, "abs", 0.3, 0.02], [4.5, "abs", 0.3, 0.02]]},
               "Info": {}}
    t = S.StackTemplate(setting)
    lo, hi, _ = t.bounds()
    assert t.max_layers() == t.max_layers(lo, hi) == 1 + 15 + 60
    wide = np.tile(hi, (3, 1)); wide[1, 2] = 70.0            # one point lets the crust grow beyond 60 km: 30 fine layers
    assert t.max_layers(np.tile(lo, (3, 1)), wide) == 1 + 30 + 60
    thin = hi.copy(); thin[2] = 19.0                         # crust never above 20 km: 10 fine layers
    assert t.max_layers(lo, thin) == 1 + 10 + 60
    rng = np.random.default_rng(5)
    params = (lo + (hi - lo) * rng.random((200, t.nparams))).astype(np.float32)
    lay, nl = MB.build_stacks(t, params, t.max_layers())
    assert nl.max() <= t.max_layers() and nl.min() >= 1 + 10 + 60
