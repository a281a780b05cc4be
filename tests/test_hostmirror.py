"""CPU tier: the CUDA kernels' per-lane arithmetic (surfdisp_core.cuh compiled with g++) replayed lane by
lane against the oracle.  This checks the kernel math and the round logic (cluster / window / interpolation
rounds, coarse scan with its kink guard, uniform-section and sequential polish)
without a GPU; the CUDA build itself is checked in test_gpu_parity.py."""
import os
import numpy as np
import pytest

from oracle import oracle as O
from pysurfinv_b200 import synth
from tests.hostmirror import mirror as HM


def _run(lay, nl, per, kind, G=4):
    c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
    dc, du = [], []
    for i in range(lay.shape[1]):
        n = int(nl[i])
        r = HM.forward(kind, lay[0, i, :n], lay[1, i, :n], lay[2, i, :n], lay[3, i, :n], lay[4, i, :n], per, G=G)
        if st0[i] == 3:
            continue
        assert r["nfound"] == nf0[i], "model %d: %d vs %d" % (i, r["nfound"], nf0[i])
        dc.append(np.abs(r["c"] - c0[i])); du.append(np.abs(r["u"] - u0[i]))
    return np.array(dc), np.array(du)


@pytest.mark.parametrize("kind", [2, 1])
def test_crustal(kind):
    lay, nl = synth.crustal_models(150, seed=3)
    dc, du = _run(lay, nl, synth.log_periods(), kind)
    assert dc.max() < 1e-4 and np.median(dc) < 2e-6
    assert (du > 1e-4).mean() < 2e-3 and np.median(du) < 5e-6


@pytest.mark.parametrize("kind", [2, 1])
def test_hand_models_unclamped_ndiv(kind):
    lay, nl = synth.hand_models(150, seed=4)
    dc, du = _run(lay, nl, synth.log_periods(16, 6.0, 60.0), kind)
    assert dc.max() < 1e-4 and du.max() < 1e-4


@pytest.mark.parametrize("G", [4, 8])   # the two widths the kernel can be instantiated with
def test_group_width_does_not_change_results(G):
    lay, nl = synth.crustal_models(40, seed=5)
    dc, du = _run(lay, nl, synth.log_periods(12), 2, G=G)
    assert dc.max() < 1e-4


def test_water_layer_and_ragged():
    lay, nl = synth.ragged_models(120, seed=6)
    per = np.array([10, 14, 20, 28, 40, 60, 80], np.float32)
    for kind in (2, 1):
        dc, du = _run(lay, nl, per, kind)
        assert dc.max() < 1e-4 and (du > 1e-4).mean() < 5e-3


def test_love_roots_just_below_half_space_velocity():
    """Long periods: the Love root sits ~1e-4 km/s below the flattened half-space velocity (a square-root cusp of
    the secular function).  The coarse scan has to fall back to the point-by-point one there (15 of these 400
    models lose their last root without that rule), and a bracket that contains the cusp has to be polished with
    the reference's own bisection/Neville sequence (model 104: the function turns back 9e-5 km/s above the root,
    a uniform section steps over both sign changes and ends on a third one above the half-space velocity)."""
    lay, nl = synth.crustal_models(400, seed=303)
    per = synth.log_periods(100, 5.0, 120.0)
    c0, u0, nf0, st0 = O.forward_batch(1, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
    assert (nf0 < len(per)).sum() >= 50       # the family does exercise the cut-off
    dc, du = _run(lay, nl, per, 1)
    assert dc.max() < 1e-4


def test_steep_branch_on_a_coarse_period_list_is_not_extrapolated():
    """Deep stacks, Love, periods 10..150 s in 10 s steps: the curve jumps by 1.7 km/s between the first two periods
    (thick slow sediments); a linear extrapolation to the third period lands on a higher mode with an even number of
    roots between c1 and the cluster (wrong by 0.7 km/s, same root count).  Extrapolations that move the root by more
    than 0.15 km/s are not trusted."""
    per = np.arange(10.0, 151.0, 10.0, dtype=np.float32)
    for seed in (3005, 3009):
        lay, nl = synth.crustal_models(100, seed=seed, n_crust=15, n_mantle=130, zmax=400.0)
        dc, du = _run(lay, nl, per, 1)
        assert dc.max() < 1e-4


def test_float32_group_velocity_state_against_float64_state(tmp_path):
    """REIGEN's ODE state in float32 as packed pairs with every sub-layer re-orthogonalised (the product's default) against
    the reference's float64 state (opts.group_f64 = 1), through the host build of the same per-lane code: thin-layer
    stacks, thick layers with ndiv = 5 sub-layers (the case that needs the re-orthogonalisation per SUB-layer), water
    layers.  The mirror reads its switches once per process, hence the two child processes."""
    import subprocess
    import sys
    script = (
        "import sys, numpy as np\n"
        "from pysurfinv_b200 import synth\n"
        "from tests.hostmirror import mirror as HM\n"
        "out = []\n"
        "for (lay, nl), per in ((synth.crustal_models(40, seed=811), synth.log_periods(20)), (synth.hand_models(60, seed=812), synth.log_periods(16, 6.0, 60.0)),\n"
        "                       (synth.ragged_models(40, seed=813), synth.log_periods(12))):\n"
        "    for m in range(lay.shape[1]):\n"
        "        n = nl[m]\n"
        "        r = HM.forward(2, lay[0, m, :n], lay[1, m, :n], lay[2, m, :n], lay[3, m, :n], lay[4, m, :n], per)\n"
        "        u = np.zeros(len(per), np.float32); u[:len(r['u'])] = r['u']; out.append(u)\n"
        "np.save(sys.argv[1], np.concatenate(out))\n")
    res = {}
    for mode, env in (("f32", {}), ("f64", {"HM_REIGEN_F64": "1"})):
        e = dict(os.environ); e.update(env)
        e["PYTHONPATH"] = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + os.pathsep + e.get("PYTHONPATH", "")
        f = str(tmp_path / ("u_%s.npy" % mode))
        subprocess.check_call([sys.executable, "-c", script, f], env=e)
        res[mode] = np.load(f)
    d = np.abs(res["f32"] - res["f64"])
    assert res["f64"].max() > 2.0 and np.count_nonzero(res["f64"]) > 2000
    assert d.max() <= 1.5e-5, d.max()
    assert np.median(d) <= 1.0e-6
