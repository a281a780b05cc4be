"""Partial derivatives of the Rayleigh phase velocity (REIGEN's dcda / dcdb / dcdr, surfa.f:1130-1135,
1179-1185, 1202-1208).  The reference computes them and never hands them out, and it holds no golden vector for them
(the TEST1 kernel files come from PHV_SENS_KERNEL on a re-layered model); they are pinned by Rayleigh's principle:
the oracle's restatement of those lines against finite differences of the oracle's own phase velocities, and the
kernels' per-lane code (host build) and the CUDA path against the oracle."""
import numpy as np
import pytest

from oracle import oracle as O
from pysurfinv_b200 import synth, senskernel as SK
from tests.hostmirror import mirror as HM

PER = np.array([8, 12, 20, 30, 50, 80], np.float32)


def test_oracle_partials_satisfy_rayleighs_principle(golden_test1):
    m = np.array(golden_test1["model"])
    h, vp, vs, rho, Q = m.T
    per = [10.0, 20.0, 40.0, 80.0]
    o = O.make_opts(precision=2, atten=0, flat=0, ndiv_cap_r=999)       # plain layered model, un-clamped sub-division
    P = O.partials(vp, vs, rho, h, 1 / Q, per, opts=o)
    scale = np.abs(P["dcdb"]).max()
    for i in (0, 3, 10, 20, 40):
        for arr, name, which in ((vs, "dcdb", 1), (vp, "dcda", 0), (rho, "dcdr", 2)):
            e = 1e-4 * arr[i]
            hi, lo = arr.copy(), arr.copy()
            hi[i] += e; lo[i] -= e
            args = lambda x: [x if which == 0 else vp, x if which == 1 else vs, x if which == 2 else rho]
            fd = (O.forward(2, *args(hi), h, 1 / Q, per, opts=o)["c"][0] - O.forward(2, *args(lo), h, 1 / Q, per, opts=o)["c"][0]) / (2 * e)
            assert np.abs(P[name][:, i] - fd).max() < 2e-3 * scale, (i, name)


@pytest.mark.parametrize("family", ["crustal", "hand", "water"])
def test_kernel_code_partials_match_oracle(family):
    lay, nl = {"crustal": synth.crustal_models(5, seed=3), "hand": synth.hand_models(5, seed=4),
               "water": synth.crustal_models(4, seed=5, water=True)}[family]
    for i in range(lay.shape[1]):
        n = int(nl[i])
        args = [lay[j, i, :n] for j in range(5)]
        r = HM.forward(2, *args, PER)
        P = HM.partials(*args, PER, r["c"], r["ratio"])
        Pr = O.partials(*[a.astype(np.float64) for a in args], PER.astype(np.float64), opts=O.make_opts(precision=1))
        for key in ("dcda", "dcdb", "dcdr"):
            assert np.abs(P[key] - Pr[key]).max() < 1e-4 * max(1.0, np.abs(Pr[key]).max()), key


def test_input_kernels_against_finite_differences():
    """The chain back through attenuation and flattening (senskernel.input_kernels): dc/dVs of the INPUT layers against
    central differences of the float64 oracle -- what SensKernelPert (senskernel.py:146-158) measures."""
    lay, nl = synth.crustal_models(2, seed=8)
    n = int(nl[0])
    vp, vs, rho, h, q = [lay[j, 0, :n].astype(np.float64) for j in range(5)]
    per = np.array([10, 20, 40, 80], np.float32)
    r = HM.forward(2, vp, vs, rho, h, q, per)
    P = HM.partials(vp, vs, rho, h, q, per, r["c"], r["ratio"])
    dvs, dvp, drho = SK.input_kernels({k: P[k].astype(np.float64) for k in P}, vp, vs, rho, h, q, per.astype(np.float64))
    o = O.make_opts(precision=2)
    scale = np.abs(dvs).max()
    for L in (0, 2, 8, 16, 30, 50, 76):
        e = 1e-4 * vs[L]
        hi, lo = vs.copy(), vs.copy()
        hi[L] += e; lo[L] -= e
        fd = (O.forward(2, vp, hi, rho, h, q, per, opts=o)["c"][0] - O.forward(2, vp, lo, rho, h, q, per, opts=o)["c"][0]) / (2 * e)
        assert np.abs(dvs[:, L] - fd).max() < 0.01 * scale, L


@pytest.mark.gpu
def test_gpu_partials_and_senskernel():
    import torch
    from pysurfinv_b200 import api
    solver = api.DispersionSolver("cuda:0")
    lay, nl = synth.ragged_models(300, seed=44)
    out = solver.partials(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), PER)
    ref = solver.forward(torch.from_numpy(lay).cuda(), torch.from_numpy(nl).cuda(), PER, kind=2)
    assert torch.equal(out["c"], ref["c"]) and torch.equal(out["nfound"], ref["nfound"])
    c = out["c"].cpu().numpy()
    for i in range(0, 300, 7):
        n = int(nl[i])
        args = [lay[j, i, :n].astype(np.float64) for j in range(5)]
        Pr = O.partials(*args, PER.astype(np.float64), opts=O.make_opts(precision=1))
        if Pr["status"] != 0 or int(out["nfound"][i]) != len(PER):
            continue
        for key in ("dcda", "dcdb", "dcdr"):
            got = out[key][i, :, :n].cpu().numpy()
            # (1e-3 of the largest kernel value: the derivatives are evaluated at the float32 root, whose 1e-6 km/s noise moves
            # the eigenfunction of thick slow sediment layers at short periods by a few 1e-4 -- measured 5e-4 worst)
            assert np.abs(got - Pr[key]).max() < 1e-3 * max(1.0, np.abs(Pr[key]).max()), (i, key)
            assert np.all(out[key][i, :, n:].cpu().numpy() == 0)
    # the consumer: SensKernelPert's quantity for a hand model, against its own arithmetic on the float64 oracle
    hm, _ = synth.hand_models(1, seed=9)
    vp, vs, rho, h, q = [hm[j, 0].astype(np.float64) for j in range(5)]
    per = np.arange(20.0, 101.0, 10.0)                                   # SensKernelPert defaults (senskernel.py:131)
    sk = SK.SensKernel(solver, vp, vs, rho, h, q, per)
    o = O.make_opts(precision=2)
    fd = np.zeros((len(per), 4))
    for L in range(4):
        lo, hi = vs.copy(), vs.copy()
        lo[L] *= 0.999; hi[L] *= 1.001
        vL = O.forward(2, vp, lo, rho, h, q, per, opts=o)["c"][0]
        vH = O.forward(2, vp, hi, rho, h, q, per, opts=o)["c"][0]
        fd[:, L] = (vH - vL) / 0.2 / h[L]                                 # senskernel.py:150
    assert np.abs(sk.kernel["Vs"] - fd).max() < 0.02 * np.abs(fd).max()
