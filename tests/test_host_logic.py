"""CPU tier: synthetic workload generator, host mirrors of the reference consumers, sharding helpers."""
import numpy as np
import pytest

from pysurfinv_b200 import synth
from pysurfinv_b200.distributed import shard_range, padded_shard_size
from pysurfinv_b200.forward import misfit, accept


def test_generator_is_deterministic_and_shaped():
    a, na = synth.crustal_models(32, seed=1)
    b, nb = synth.crustal_models(32, seed=1)
    assert a.dtype == np.float32 and a.shape == (5, 32, 77) and np.array_equal(a, b) and np.all(na == 77)
    vp, vs, rho, h, qi = a
    assert np.all(vs[:, 1:16] >= 3.2 - 1e-6) and np.all(np.diff(vs[:, 1:16], axis=1) >= -1e-6)  # monotone crust
    assert np.allclose(h[:, :76].sum(1), 200.0, atol=1e-3)
    assert np.allclose(qi[:, 0], 1 / 80.0) and np.allclose(qi[:, 1:16], 1 / 600.0) and np.allclose(qi[:, 16:], 1 / 150.0)
    w, nw = synth.crustal_models(4, seed=1, water=True)
    assert w.shape[2] == 78 and np.all(w[1, :, 0] == 0)
    r, nr = synth.ragged_models(64, seed=2)
    assert set(np.unique(nr)) <= {4, 77, 78, 92} and r.shape == (5, 64, 96)
    for i in range(64):
        assert np.all(r[:, i, nr[i]:] == 0)
    p = synth.log_periods()
    assert len(p) == 40 and abs(p[0] - 8) < 1e-5 and abs(p[-1] - 80) < 1e-4 and np.all(np.diff(p) > 0)


def test_bspline_partition_of_unity():
    for n in (1, 2, 3, 4, 5, 7):
        B = synth.bspline_basis(np.linspace(0, 30, 41), n)
        assert B.shape == (n, 41) and np.allclose(B.sum(0), 1.0) and np.all(B >= -1e-12)


def test_misfit_matches_reference_formula():
    # observed Juan-de-Fuca curve quoted (commented) in reference point.py:400-404
    cO = np.array([3.572, 3.629, 3.672, 3.705, 3.730, 3.749, 3.764, 3.777, 3.787, 3.796, 3.803, 3.810, 3.823, 3.835,
                   3.860, 3.878, 3.892, 3.903])
    sig = np.full(18, 0.01)
    cP = cO + 0.004
    m, chi, L = misfit(cO, cP, sig)
    assert np.isclose(chi, 18 * 0.16) and np.isclose(m, 0.4) and np.isclose(L, np.exp(-0.5 * chi))
    m, chi, L = misfit(cO, cO + 0.05, sig)          # chi2 = 450 -> soft clip sqrt(50*450)
    assert np.isclose(chi, np.sqrt(50 * 450.0)) and np.isclose(m, 5.0)
    assert misfit(cO, None, sig) == (88888, 88888, 0)
    cm = np.ma.masked_array(cO, mask=[True] + [False] * 17)
    m, chi, L = misfit(cm, cP, sig)
    assert np.isclose(chi, 17 * 0.16)


def test_accept_rule():
    assert accept(10.0, 9.0, 0.999) is True
    assert bool(accept(10.0, 12.0, 0.99)) and not bool(accept(10.0, 12.0, 0.5))   # threshold 1-exp(-1) = 0.632


@pytest.mark.parametrize("M,W", [(10, 3), (8, 8), (5, 8), (1048576, 8), (0, 4)])
def test_shard_ranges_cover_exactly(M, W):
    spans = [shard_range(M, r, W) for r in range(W)]
    assert spans[0][0] == 0 and spans[-1][1] == M
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and b >= a
    assert max(b - a for a, b in spans) == (padded_shard_size(M, W) if M else 0)
