"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run on the B200 box.

Tolerance (BASELINE.json north_star): |dc|, |dU| <= 1e-4 km/s and identical root counts, measured
against the float32-faithful oracle (precision 0 = fast_surf semantics).  The oracle's own float32 noise
(float32 solver vs float64 solver on the same float32 model, precision 0 vs 1) is computed alongside:
where the reference itself is ill-conditioned (|dU_noise| large, short periods on slow sediments) the
group-velocity tolerance is widened by that noise, and the fraction of such points is bounded.
"""
import numpy as np
import pytest

from oracle import oracle as O
from pysurfinv_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def solver():
    import torch
    from pysurfinv_b200 import api
    assert torch.cuda.is_available()
    return api.DispersionSolver("cuda:0")


def _gpu(solver, lay, nl, per, kind):
    import torch
    out = solver.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per, kind=kind)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _check(g, lay, nl, per, kind):
    c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
    c1, u1, nf1, st1 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=1), nthreads=8)
    ok = st0 != 3  # LSTOP aborts of the reference are excluded and counted
    assert ok.mean() > 0.99
    assert np.array_equal(g["nfound"][ok], nf0[ok]), "root counts differ"
    dc = np.abs(g["c"] - c0)[ok]
    du = np.abs(g["u"] - u0)[ok]
    noise_u = np.abs(u0 - u1)[ok]
    assert dc.max() <= TOL, "phase velocity off by %g" % dc.max()
    assert np.median(dc) < 2e-6
    # group velocity: |dU| <= 1e-4 except where the reference's own U is ill-conditioned (its float32-vs-float64-solver
    # spread on the same float32 model reaches 1e-3..1e-2 at ~1e-4 of the points: short periods on slow sediments).
    # The bar is the reference's own noise, measured on the same models: at most 3x its noisy fraction (+1e-4 for the
    # sampling error of a rare event), the same for the 99.9 % quantile and the maximum.
    frac_bad = (du > TOL).mean()
    assert frac_bad <= 3.0 * (noise_u > TOL).mean() + 1e-4, (frac_bad, (noise_u > TOL).mean())
    assert np.quantile(du, 0.999) <= 3.0 * max(np.quantile(noise_u, 0.999), 1e-5)
    assert du.max() <= max(3.0 * noise_u.max(), TOL), (du.max(), noise_u.max())
    assert np.median(du) < 5e-6
    # beyond nfound everything is zero
    K = len(per)
    beyond = np.arange(K)[None, :] >= g["nfound"][:, None]
    assert np.all(g["c"][beyond] == 0) and np.all(g["u"][beyond] == 0)
    return dc, du


@pytest.mark.parametrize("kind", [2, 1])
def test_crustal_models(solver, kind):
    lay, nl = synth.crustal_models(1500, seed=11)
    per = synth.log_periods()
    g = _gpu(solver, lay, nl, per, kind)
    _check(g, lay, nl, per, kind)


@pytest.mark.parametrize("kind", [2, 1])
def test_hand_models_ndiv5(solver, kind):
    lay, nl = synth.hand_models(1000, seed=12)
    per = synth.log_periods(24, 6.0, 60.0)
    g = _gpu(solver, lay, nl, per, kind)
    _check(g, lay, nl, per, kind)


@pytest.mark.parametrize("kind", [2, 1])
def test_ragged_with_water(solver, kind):
    lay, nl = synth.ragged_models(800, seed=13)
    per = np.array([10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)  # point.py:400
    g = _gpu(solver, lay, nl, per, kind)
    _check(g, lay, nl, per, kind)


def test_golden_test1_model(solver, golden_test1):
    """The reference's own test model (68 layers to 1300 km): float32 fast_surf semantics must land
    within the float32 flattening noise (3e-4) of the real*8 golden curve, and on the oracle."""
    m = np.array(golden_test1["model"])
    h, vp, vs, rho, Q = m.T
    lay = np.stack([vp, vs, rho, h, 1 / Q]).astype(np.float32)[:, None, :]
    nl = np.array([len(h)], np.int32)
    per = np.array(golden_test1["periods"], np.float32)
    for kind, w in ((2, "R"), (1, "L")):
        g = _gpu(solver, np.ascontiguousarray(lay), nl, per, kind)
        assert g["nfound"][0] == 10
        assert np.abs(g["c"][0] - np.array(golden_test1[w]["c"][0])).max() < 3e-4
        _check(g, np.ascontiguousarray(lay), nl, per, kind)


def test_failure_and_edge_cases(solver):
    import torch
    per = synth.log_periods(8)
    # (a) empty batch
    out = solver.forward(torch.empty((5, 0, 8), device="cuda"), torch.empty(0, dtype=torch.int32, device="cuda"), per)
    assert out["c"].shape == (0, 8)
    # (b) a model with no root: half-space slower than everything above -> nfound = 0, zeros, flag set
    lay = np.zeros((5, 2, 4), np.float32)
    lay[:, 0, :3] = np.array([[6.0, 6.5, 5.0], [3.5, 3.8, 2.0], [2.7, 2.9, 2.5], [10, 20, 10], [1 / 600.] * 3], np.float32)
    good, gn = synth.hand_models(1, seed=5)
    lay[:, 1, :4] = good[:, 0]
    nl = np.array([3, 4], np.int32)
    g = _gpu(solver, lay, nl, per, 2)
    c0, u0, nf0, st0 = O.forward_batch(2, lay, nl, per, opts=O.make_opts(precision=0))
    assert np.array_equal(g["nfound"], nf0)
    assert g["nfound"][1] == 8
    if g["nfound"][0] == 0:
        assert g["flags"][0] & 1 and np.all(g["c"][0] == 0)
    # (c) invalid layer count is rejected per model, not fatal
    nl_bad = np.array([1, 4], np.int32)
    g = _gpu(solver, lay, nl_bad, per, 2)
    assert g["nfound"][0] == 0 and g["nfound"][1] == 8


def test_fast_surf_shim_and_cal_forward(solver):
    from pysurfinv_b200 import fast_surf as FS
    from pysurfinv_b200.forward import cal_forward
    lay, nl = synth.crustal_models(1, seed=21)
    vp, vs, rho, h, qsinv = (lay[i, 0].astype(np.float64) for i in range(5))
    periods = [8.0, 10, 20, 40, 60, 80]
    per = np.zeros(200); per[:6] = periods
    ur0, ul0, cr0, cl0 = FS.fast_surf(len(h), 2, vp, vs, rho, h, qsinv, per, 6)
    assert cr0.dtype == np.float32 and cr0.shape == (200,)
    ref = O.forward(2, vp, vs, rho, h, qsinv, periods)
    assert np.abs(cr0[:6] - ref["c"][0]).max() < TOL and np.abs(ur0[:6] - ref["u"][0]).max() < TOL
    assert np.all(cr0[6:] == 0) and np.all(cl0 == 0) and np.all(ul0 == 0)
    prof = np.stack([h, vs, vp, rho, 1 / qsinv, 2 / qsinv])
    cp = cal_forward(prof, "Ray", periods)
    assert np.allclose(cp, cr0[:6])
    ur0, ul0, cr0, cl0 = FS.fast_surf(len(h), 1, vp, vs, rho, h, qsinv, per, 6)
    refl = O.forward(1, vp, vs, rho, h, qsinv, periods)
    assert np.abs(cl0[:6] - refl["c"][0]).max() < TOL and np.all(cr0 == 0)


def test_host_batch_matches_device_path(solver):
    from pysurfinv_b200 import api
    lay, nl = synth.crustal_models(64, seed=31)
    per = synth.log_periods(12)
    g = _gpu(solver, lay, nl, per, 2)
    h = api.host_batch(lay, nl, per, 2)
    assert np.array_equal(h["c"], g["c"]) and np.array_equal(h["u"], g["u"]) and np.array_equal(h["nfound"], g["nfound"])
    p = solver.forward_host(lay, nl, per, 2)
    assert np.array_equal(p["c"], g["c"]) and np.array_equal(p["u"], g["u"])


def test_pipelined_host_batch_large(solver):
    """surfdisp_host_batch above 65536 models runs as an 8-chunk pipeline (copies under the kernels, first period per
    chunk, later periods in one launch, group velocities per chunk): identical to the device-resident call, Rayleigh
    and Love, pageable host memory through the pure C ABI and pinned memory through the solver."""
    from pysurfinv_b200 import api
    lay, nl = synth.ragged_models(70001, seed=33)
    per = np.array([10, 14, 20, 28, 40, 60], np.float32)
    for kind in (2, 1):
        g = _gpu(solver, lay, nl, per, kind)
        h = api.host_batch(lay, nl, per, kind)
        for key in ("c", "u", "nfound", "flags"):
            assert np.array_equal(h[key], g[key]), key
        p = solver.forward_host(lay, nl, per, kind)
        for key in ("c", "u", "nfound", "flags"):
            assert np.array_equal(p[key], g[key]), key
    # phase velocities only
    p = solver.forward_host(lay, nl, per, 2, group=False)
    assert np.array_equal(p["c"], _gpu(solver, lay, nl, per, 2)["c"]) and p["u"] is None


def test_misfit_kernel(solver):
    import torch
    from pysurfinv_b200.forward import misfit as host_misfit
    lay, nl = synth.crustal_models(256, seed=41)
    per = np.array([10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
    # observed curve of the reference's commented example (point.py:400-404 layout: c, sigma per period)
    rng = np.random.default_rng(0)
    out = solver.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per, kind=2)
    c = out["c"].cpu().numpy()
    obs = c[7] + rng.normal(0, 0.01, len(per)).astype(np.float32)
    sig = np.full(len(per), 0.01, np.float32)
    mask = np.ones(len(per), np.uint8); mask[3] = 0
    m = solver.misfit(out["c"], out["nfound"], obs, sig, mask=mask, mode=0).cpu().numpy()
    for i in (0, 7, 100, 255):
        mo = np.ma.masked_array(obs.astype(np.float64), mask=(mask == 0))
        e = host_misfit(mo, c[i].astype(np.float64), sig.astype(np.float64))
        assert np.allclose(m[i], np.array(e, dtype=np.float64), rtol=2e-5, atol=1e-30)
    # failure sentinel
    nf = out["nfound"].clone(); nf[3] = 5
    m2 = solver.misfit(out["c"], nf, obs, sig, mode=0).cpu().numpy()
    assert tuple(m2[3]) == (88888.0, 88888.0, 0.0)
    # Cascadia variant: mean of the two period bands
    m3 = solver.misfit(out["c"], out["nfound"], obs, sig, periods=per, mode=1).cpu().numpy()
    bias = (obs.astype(np.float64) - c[9]) / sig
    chi = ((bias[per <= 40] ** 2).mean() + (bias[per > 40] ** 2).mean()) / 2 * len(per)
    mis = np.sqrt(chi / len(per)); chi = chi if chi < 50 else np.sqrt(chi * 50)
    assert np.allclose(m3[9], [mis, chi, np.exp(-0.5 * chi)], rtol=2e-5, atol=1e-30)


def test_full_size_properties(solver):
    """Config-2 sized slice (131072 models x 40 periods): size-independent properties -- permutation
    invariance (a model's result does not depend on its batch position or neighbours), determinism,
    c <= half-space velocity, U < c for the normally dispersive long periods, all roots found."""
    import torch
    M = 131072
    lay, nl = synth.crustal_models(M, seed=51)
    per = synth.log_periods()
    dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
    a = solver.forward(dl, dn, per, kind=2)
    c_a, u_a, nf_a = a["c"].clone(), a["u"].clone(), a["nfound"].clone()
    perm = torch.randperm(M, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    b = solver.forward(dl[:, perm].contiguous(), dn[perm].contiguous(), per, kind=2)
    assert torch.equal(b["c"], c_a[perm]) and torch.equal(b["u"], u_a[perm]) and torch.equal(b["nfound"], nf_a[perm])
    full = nf_a == len(per)
    assert int(full.sum()) >= M - max(4, M // 10000)     # a handful of stacks lose the root at the longest periods
    vs_half = dl[1, :, 76]
    assert bool((c_a.max(dim=1).values <= vs_half * 1.05).all())
    assert bool((c_a[full] > 0.5).all()) and bool(torch.isfinite(u_a).all())
    r = u_a[full] / c_a[full]
    assert float(r.min()) > 0.05 and float(r.max()) < 1.5
    # oracle spot check on a strided subset at the full layout
    idx = np.unique(np.concatenate([np.arange(0, M, M // 256), np.nonzero(~full.cpu().numpy())[0]]))
    c0, u0, nf0, st0 = O.forward_batch(2, lay[:, idx], nl[idx], per, opts=O.make_opts(precision=0), nthreads=8)
    assert np.array_equal(nf_a.cpu().numpy()[idx], nf0)
    assert np.abs(c_a.cpu().numpy()[idx] - c0).max() <= TOL


def test_exact_scan_mode_gives_same_roots():
    """The coarse-to-fine scan + clustered polish (default) against the plain every-grid-point scan +
    uniform section (exact_scan=1): same root counts, same brackets -> roots equal to float32 noise."""
    import torch
    from pysurfinv_b200 import api
    lay, nl = synth.crustal_models(20000, seed=61)
    per = synth.log_periods()
    dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
    for kind in (2, 1):
        a = api.DispersionSolver("cuda:0").forward(dl, dn, per, kind=kind)
        b = api.DispersionSolver("cuda:0", opts=api.default_opts(exact_scan=1)).forward(dl, dn, per, kind=kind)
        assert torch.equal(a["nfound"], b["nfound"])
        assert float((a["c"] - b["c"]).abs().max()) < 5e-5
        du = (a["u"] - b["u"]).abs()
        assert float(du.median()) < 2e-6 and float((du > 1e-4).float().mean()) < 1e-3


def test_love_roots_just_below_half_space_velocity(solver):
    """Long periods: the Love root sits ~1e-4 km/s below the flattened half-space velocity, where the secular
    function has a square-root cusp (surfa.f:407-416 switches to the evanescent branch).  The coarse scan must
    fall back to point-by-point there; root counts against the oracle and against exact_scan=1."""
    import torch
    from pysurfinv_b200 import api
    lay, nl = synth.crustal_models(4096, seed=303)
    per = synth.log_periods(100, 5.0, 120.0)
    dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
    a = solver.forward(dl, dn, per, kind=1)
    b = api.DispersionSolver("cuda:0", opts=api.default_opts(exact_scan=1)).forward(dl, dn, per, kind=1)
    assert torch.equal(a["nfound"], b["nfound"])
    partial = np.nonzero((a["nfound"] < len(per)).cpu().numpy())[0]
    assert len(partial) > 100            # the family does exercise the cut-off
    idx = partial[:256]
    c0, u0, nf0, st0 = O.forward_batch(1, lay[:, idx], nl[idx], per, opts=O.make_opts(precision=0), nthreads=8)
    assert np.array_equal(a["nfound"].cpu().numpy()[idx], nf0)
    assert np.abs(a["c"].cpu().numpy()[idx] - c0).max() <= TOL
    # a bracket that contains the cusp: the function turns back 9e-5 km/s above the root of model 104 / period 58; only
    # the reference's own bisection/Neville sequence ends on the root the reference finds (tests/test_hostmirror.py)
    lay, nl = synth.crustal_models(400, seed=303)
    g = _gpu(solver, lay, nl, per, 1)
    c0, u0, nf0, st0 = O.forward_batch(1, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
    ok = st0 != 3
    assert np.array_equal(g["nfound"][ok], nf0[ok]) and nf0[104] == 59
    assert np.abs(g["c"] - c0)[ok].max() <= TOL


def test_velocity_inversion_scan_rounds_on_own_truncations(solver):
    """A stack whose half-space is slower than the mantle above it (crustal_models(400, seed=2089)[324], Rayleigh,
    T = 8.4 s): the scan passes the half-space velocity of the whole stack, where the truncations of the lower points
    of a round (calcul.f:155-159: own layer dropping per point, surfa.f:92-106) do not have the sign of the deepest
    one.  Round 1 reported 17 roots there, the reference 16.  Default and exact_scan modes, plus the LVZ family."""
    from pysurfinv_b200 import api
    lay, nl = synth.crustal_models(400, seed=2089)
    per = synth.log_periods(100, 5.0, 120.0)
    c0, u0, nf0, st0 = O.forward_batch(2, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
    assert nf0[324] == 16
    ok = st0 != 3
    for sv in (solver, api.DispersionSolver("cuda:0", opts=api.default_opts(exact_scan=1))):
        g = _gpu(sv, lay, nl, per, 2)
        assert g["nfound"][324] == 16
        assert np.array_equal(g["nfound"][ok], nf0[ok])
        assert np.abs(g["c"] - c0)[ok].max() <= TOL
    for kind in (2, 1):
        lay, nl = synth.crustal_models(1500, seed=2090, lvz=True)
        g = _gpu(solver, lay, nl, per, kind)
        c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
        ok = st0 != 3
        assert np.array_equal(g["nfound"][ok], nf0[ok])
        assert np.abs(g["c"] - c0)[ok].max() <= TOL


def test_config4_deep_stacks_ndiv_zero(solver):
    """BASELINE config 4 shape: ~150 fine layers, Rayleigh 10-150 s.  n >= 101 clamps ndiv to 99/(n-1) = 0
    (no sub-division, surfa.f:783-787)."""
    lay, nl = synth.crustal_models(300, seed=71, n_crust=15, n_mantle=130, zmax=400.0)
    assert lay.shape[2] == 147
    per = np.arange(10.0, 151.0, 10.0, dtype=np.float32)
    g = _gpu(solver, lay, nl, per, 2)
    _check(g, lay, nl, per, 2)


def test_config4_thermal_ocean_models(solver):
    """BASELINE config 4 on its real parameterisation: the ocean model of point.py:374-391 (water, sediment, crust,
    OceanMantleHybrid thermal mantle, reference mantle: 86 layers) with random admissible parameters, Rayleigh
    10-150 s.  The stacks are assembled by the numpy restatement of the reference's classes; the water layer makes
    the search start at 0.5 km/s (fast_surf.f:170) and the sweep pass a liquid layer."""
    import bench
    from oracle import model_builder as MB
    from pysurfinv_b200 import stack as S
    t = S.StackTemplate(bench.THERMAL_SETTING, prior_mask=S.PRIOR_OCEAN)
    lo, hi, _ = t.bounds()
    rng = np.random.default_rng(404)
    params = []
    while len(params) < 300:
        p = (lo + (hi - lo) * rng.random(t.nparams)).astype(np.float32)
        if MB.priors_ocean(t, p.astype(np.float64)) & S.PRIOR_OCEAN == 0:
            params.append(p)
    lay, nl = MB.build_stacks(t, np.array(params), t.max_layers())
    assert int(nl.max()) == 86 and float(lay[1, :, 0].max()) == 0.0
    g = _gpu(solver, lay, nl, bench.THERMAL_PERIODS, 2)
    _check(g, lay, nl, bench.THERMAL_PERIODS, 2)


def test_config4_love_steep_branch_coarse_periods(solver):
    """Config-4 stacks, Love, 10 s period steps: the curve jumps by 1.7 km/s between the first two periods; the third
    period must not be extrapolated onto a higher mode (tests/test_hostmirror.py, same models)."""
    per = np.arange(10.0, 151.0, 10.0, dtype=np.float32)
    for seed in (3005, 3009):
        lay, nl = synth.crustal_models(100, seed=seed, n_crust=15, n_mantle=130, zmax=400.0)
        g = _gpu(solver, lay, nl, per, 1)
        c0, u0, nf0, st0 = O.forward_batch(1, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
        ok = st0 != 3
        assert np.array_equal(g["nfound"][ok], nf0[ok])
        assert np.abs(g["c"] - c0)[ok].max() <= TOL


def test_very_deep_stack_and_many_periods(solver):
    """Shared-memory sizing path: 500-layer stacks (CTA shrinks) and the maximum of 200 periods."""
    lay, nl = synth.crustal_models(24, seed=72, n_crust=60, n_mantle=438, zmax=600.0)
    assert lay.shape[2] == 500
    per = np.linspace(8.0, 100.0, 200).astype(np.float32)
    for kind in (2, 1):
        g = _gpu(solver, lay, nl, per, kind)
        c0, u0, nf0, st0 = O.forward_batch(kind, lay, nl, per, opts=O.make_opts(precision=0), nthreads=8)
        ok = st0 != 3
        assert np.array_equal(g["nfound"][ok], nf0[ok])
        assert np.abs(g["c"] - c0)[ok].max() <= TOL
        assert np.median(np.abs(g["u"] - u0)[ok]) < 1e-5


def test_config3_joint_rayleigh_love_anisotropy(solver):
    """BASELINE config 3: Love on Vsh, Rayleigh on Vsv of the same stacks (radial anisotropy is outside the
    reference; parity is per wave type)."""
    lay, nl = synth.crustal_models(600, seed=73)
    rng = np.random.default_rng(73)
    xi = 1.0 + rng.uniform(-0.05, 0.05, (600, 1)).astype(np.float32)
    lay_sh = lay.copy()
    lay_sh[1, :, 16:] *= xi          # Vsh = xi * Vsv in the mantle
    per = synth.log_periods(24)
    gr = _gpu(solver, lay, nl, per, 2)
    gl = _gpu(solver, np.ascontiguousarray(lay_sh), nl, per, 1)
    _check(gr, lay, nl, per, 2)
    _check(gl, np.ascontiguousarray(lay_sh), nl, per, 1)


def test_chunked_host_path_matches_device_path(solver):
    """forward_host cuts large batches into chunks whose PCIe copies overlap the kernels: same results as one
    device-resident call, for a chunk count that does not divide the batch."""
    import torch
    lay, nl = synth.crustal_models(1001, seed=77)
    per = synth.log_periods(12)
    d = solver.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), per, kind=2)
    torch.cuda.synchronize()
    for chunks in (1, 3, 8):
        h = solver.forward_host(lay, nl, per, 2, chunks=chunks)
        assert np.array_equal(h["c"], d["c"].cpu().numpy()) and np.array_equal(h["u"], d["u"].cpu().numpy())
        assert np.array_equal(h["nfound"], d["nfound"].cpu().numpy()) and np.array_equal(h["flags"], d["flags"].cpu().numpy())


def test_neighbour_curve_hints_change_sweeps_not_results(solver):
    """surfdisp_batch_hinted: the curve of a nearby model (the chain's current model in a Monte-Carlo walk) centres the
    trial velocities of the later periods.  Good hints save sweeps; good, wrong and missing hints all give the root
    counts of the un-hinted search and the same roots to float32 noise -- and the oracle's."""
    import torch
    P18 = np.array([8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 50, 60, 70, 80], np.float32)
    lay, nl = synth.ragged_models(6000, seed=91)
    rng = np.random.default_rng(91)
    lay2 = lay.copy()
    lay2[1] *= (1.0 + rng.normal(0, 0.004, lay2[1].shape)).astype(np.float32)      # the neighbour: Vs moved by ~0.4 %
    dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
    for kind in (2, 1):
        base = solver.forward(dl, dn, P18, kind=kind)
        base = {k: v.clone() for k, v in base.items()}
        sw0 = solver.counters()[1]
        hint = solver.forward(torch.from_numpy(lay2).cuda(), dn, P18, kind=kind)["c"].clone()
        good = solver.forward(dl, dn, P18, kind=kind, hint=hint)
        sw1 = solver.counters()[1]
        assert torch.equal(good["nfound"], base["nfound"])
        assert float((good["c"] - base["c"]).abs().max()) < 3e-5
        assert sw1 < 0.9 * sw0, (sw0, sw1)                       # the hint pays
        junk = torch.from_numpy(rng.uniform(1.0, 5.0, tuple(hint.shape)).astype(np.float32)).cuda()
        junk[::3] = 0.0                                           # every third model: no hint
        bad = solver.forward(dl, dn, P18, kind=kind, hint=junk)
        assert torch.equal(bad["nfound"], base["nfound"])
        assert float((bad["c"] - base["c"]).abs().max()) < 3e-5
        idx = np.arange(0, 6000, 12)
        c0, u0, nf0, st0 = O.forward_batch(kind, lay[:, idx], nl[idx], P18, opts=O.make_opts(precision=0), nthreads=8)
        ok = st0 != 3
        assert np.array_equal(good["nfound"].cpu().numpy()[idx][ok], nf0[ok])
        assert np.abs(good["c"].cpu().numpy()[idx] - c0)[ok].max() <= TOL


def test_hand_over_between_root_search_launches_changes_nothing(solver):
    """Large batches run the later periods as two launches: a fast-path instantiation that contains no scan code and
    hands the models whose period needs the point-by-point scan (calcul.f:155-167) over to the general one, which
    resumes them where the scan would have started (stale layer records rebuilt from the recorded dropping depths)
    on a side stream beside the first group-velocity pass.  Forced here on small batches of the families that scan
    most (velocity inversions, water layers, 100-period lists, Love): bit-identical c, U, root counts and flags, on
    the device path, with hints, and through both host-buffer pipelines."""
    import torch
    from pysurfinv_b200 import stack as S
    lib = solver.lib
    P100 = synth.log_periods(100, 5.0, 120.0)
    cases = [(synth.crustal_models(3000, seed=501, lvz=True), P100, 2), (synth.crustal_models(3000, seed=502, lvz=True), P100, 1),
             (synth.ragged_models(3000, seed=503), synth.log_periods(24), 2), (synth.crustal_models(4000, seed=504), synth.log_periods(), 2),
             (synth.crustal_models(1500, seed=505, n_crust=15, n_mantle=130, zmax=400.0), np.arange(10.0, 151.0, 10.0, dtype=np.float32), 1)]
    handed = 0
    try:
        for (lay, nl), per, kind in cases:
            dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
            lib.surfdisp_set_split_min_models(0)
            ref = {k: v.cpu().numpy() for k, v in solver.forward(dl, dn, per, kind=kind).items()}
            hint = torch.from_numpy(ref["c"] * np.float32(1.002)).cuda().contiguous()
            ref_h = {k: v.cpu().numpy() for k, v in solver.forward(dl, dn, per, kind=kind, hint=hint).items()}
            lib.surfdisp_set_split_min_models(1)
            got = {k: v.cpu().numpy() for k, v in solver.forward(dl, dn, per, kind=kind).items()}
            got_h = {k: v.cpu().numpy() for k, v in solver.forward(dl, dn, per, kind=kind, hint=hint).items()}
            host = solver.forward_host(lay, nl, per, kind, chunks=3)
            for a, b in ((ref, got), (ref_h, got_h), (ref, host)):
                for key in ("c", "u", "nfound", "flags"):
                    assert np.array_equal(a[key], b[key]), key
            cnt = torch.zeros(64, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            # (the number of handed-over models of the last call sits in the workspace header)
            cnt.copy_(solver._ws[:256].view(torch.int32))
            handed += int(cnt[32])
        # parameters-fed pipeline (config-2 setting)
        t, _ = S.config2_template()
        params = S.config2_params(5000, seed=506)
        per = synth.log_periods()
        lib.surfdisp_set_split_min_models(0)
        a = solver.forward_params_pinned(t, torch.from_numpy(params).pin_memory(), per)
        a = {k: np.array(v) for k, v in a.items()}
        lib.surfdisp_set_split_min_models(1)
        b = solver.forward_params_pinned(t, torch.from_numpy(params).pin_memory(), per)
        for key in ("c", "u", "nfound", "flags"):
            assert np.array_equal(a[key], b[key]), key
    finally:
        lib.surfdisp_set_split_min_models(0)
    assert handed > 100      # the path was taken


def test_group_velocity_float32_state_against_float64_state(solver):
    """REIGEN's ODE state and energy sums in float32 (default: every sub-layer re-orthogonalised) against float64 like the
    reference (opts.group_f64 = 1): same roots, |dU| <= 3e-5 km/s on thin-layer stacks, thick layers with sub-division,
    water layers, velocity inversions, 147- and 497-layer stacks up to 200 s -- and both within the parity bars."""
    import torch
    from pysurfinv_b200 import api
    f64 = api.DispersionSolver("cuda:0", opts=api.default_opts(group_f64=1))
    cases = [(synth.crustal_models(3000, seed=601), synth.log_periods()), (synth.crustal_models(1500, seed=602, lvz=True), synth.log_periods(100, 5.0, 120.0)),
             (synth.hand_models(2000, seed=603), synth.log_periods(24, 6.0, 60.0)), (synth.ragged_models(2000, seed=604), synth.log_periods(24)),
             (synth.crustal_models(400, seed=605, n_crust=15, n_mantle=130, zmax=400.0), np.arange(10.0, 151.0, 10.0, dtype=np.float32)),
             (synth.crustal_models(60, seed=606, n_crust=40, n_mantle=455, zmax=600.0), synth.log_periods(60, 5.0, 200.0))]
    for (lay, nl), per in cases:
        a = _gpu(solver, lay, nl, per, 2)
        dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
        b = {k: v.cpu().numpy() for k, v in f64.forward(dl, dn, per, kind=2).items()}
        assert np.array_equal(a["c"], b["c"]) and np.array_equal(a["nfound"], b["nfound"])
        assert np.abs(a["u"] - b["u"]).max() <= 3e-5, np.abs(a["u"] - b["u"]).max()
    lay, nl = synth.crustal_models(1500, seed=607)
    per = synth.log_periods()
    dl, dn = torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda()
    _check({k: v.cpu().numpy() for k, v in f64.forward(dl, dn, per, kind=2).items()}, lay, nl, per, 2)
