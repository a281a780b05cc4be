"""CPU restatement (numpy, float64) of the reference's model assembly -- TEST INFRASTRUCTURE ONLY.

Follows layers.py:4-45 (BsplBasis), :104-136 (SeisLayerVs.seisPropGrids), the per-class rules at
:139-295, and models.py:72-102 (Model1D.seisPropGrids / seisPropLayers).  Pinned against fixtures produced by
running the reference classes themselves (tests/golden/layers_reference.json, make_golden_layers.py).
"""
import numpy as np

from pysurfinv_b200 import stack as S   # template records only (shared description of the inputs)


def bspl_basis(npts, n, alpha=2.0):
    """Basis [n, npts] on an equally spaced grid of npts points (layers.py:4-45); independent of the span."""
    eps = np.finfo(float).eps
    z = np.linspace(0.0, 1.0, npts)
    if n == 1:
        return np.ones((1, npts))
    if n == 2:
        return np.stack([np.linspace(1, 0, npts), np.linspace(0, 1, npts)])
    deg = 3 + (n >= 4)
    x = np.zeros(n + deg)
    x[:deg - 1] = -eps
    x[deg - 1] = 0.0
    x[deg:n] = alpha ** np.arange(n - deg) * (alpha - 1) / (alpha ** (n - deg + 1) - 1)
    x[n] = 1.0
    x[n + 1:] = 1 + eps
    nc = len(x) - 1
    b0 = np.zeros((npts, nc))
    for i in range(nc):
        b0[(z >= x[i]) & (z < x[i + 1]), i] = 1
    b1 = b0.copy()
    for r in range(deg - 1):
        for i in range(nc - r - 1):
            col = np.zeros(npts)
            d1 = x[i + r + 1] - x[i]
            d2 = x[i + r + 2] - x[i + 1]
            if d1 != 0:
                col += b0[:, i] * (z - x[i]) / d1
            if d2 != 0:
                col += b0[:, i + 1] * (x[i + r + 2] - z) / d2
            b1[:, i] = col
        b0 = b1.copy()
    return b1[:, :n].T.copy()


def nfine(rule, fixed, H):
    if rule == S.N_FIXED:
        return fixed
    if rule == S.N_CRUST:   # layers.py:161-173
        return 60 if H >= 150 else 30 if H > 60 else 15 if H > 20 else 10 if H > 10 else 5
    return min(max(int(round(H / 2)), 2), 10)   # layers.py:226 (python round = half to even)


def rho_of(rule, const, vs, vp):
    if rule == S.R_QUARTIC:
        return 1.22679 + 1.53201 * vs - 0.83668 * vs * vs + 0.20673 * vs ** 3 - 0.01656 * vs ** 4
    if rule == S.R_OCEAN:
        return 0.541 + 0.3601 * vp
    if rule == S.R_MANTLE:
        return 3.4268 + (vs - 4.5) / 4.5
    return np.full_like(vs, const)


def build_one(tmpl, p):
    """Grid assembly of models.py:72-91 then layer averaging of :93-102 for one parameter vector."""
    z_all, vs_all, vp_all, rho_all, qs_all = [], [], [], [], []
    z0 = -max(tmpl.topo, 0.0)
    for g in tmpl.groups:
        hv = p[g.h_param] if g.h_param >= 0 else g.h_fixed
        top = z_all[-1] if z_all else 0.0   # layersAbove[0][-1]; empty list -> BottomDepth is taken from 0
        H = float(hv) if g.h_mode == 0 else float(hv) - top
        if g.h_mode == 1 and not z_all:
            H = float(hv)
        N = nfine(g.nfine_rule, g.nfine, H)
        z = np.linspace(0, H, N + 1)
        coef = np.array([p[g.v_param[i]] if g.v_param[i] >= 0 else g.v_fixed[i] for i in range(g.ncoef)], dtype=np.float64)
        if g.kind == S.G_WATER:
            vs = np.zeros(N + 1)
        elif g.kind == S.G_CONST:
            vs = np.full(N + 1, coef[0])
        elif g.kind == S.G_LINEAR:
            vs = np.linspace(coef[0], coef[1], N + 1)
        elif g.kind == S.G_BSPLINE:
            vs = coef @ bspl_basis(N + 1, g.ncoef)
        elif g.kind == S.G_CASCADIA:
            vs = np.full(N + 1, (0.02 * H ** 2 + 1.27 * H + 0.29 * 0.1) / (H + 0.29))
        else:   # reference mantle, layers.py:267-285
            vs0 = vs_all[-1]
            vs = np.linspace(vs0, vs0 + H * g.slope, N + 1)
        vp = g.vp_a * vs + g.vp_b
        rho = rho_of(g.rho_rule, g.rho_const, vs, vp)
        qs = np.full(N + 1, g.qs, dtype=np.float64)
        if g.kind == S.G_REFMANTLE:
            vp = vp_all[-1] + (vp - vp[0]); rho = rho_all[-1] + (rho - rho[0]); qs = qs_all[-1] + (qs - qs[0])
        if z[-1] - z[0] < 0.01:
            continue
        z_all += list(z + z0); vs_all += list(vs); vp_all += list(vp); rho_all += list(rho); qs_all += list(qs)
        z0 = z_all[-1]
    z = np.array(z_all); h = np.diff(z)
    mid = lambda a: (np.array(a)[1:] + np.array(a)[:-1]) / 2
    keep = h > 0.01
    return h[keep], mid(vs_all)[keep], mid(vp_all)[keep], mid(rho_all)[keep], mid(qs_all)[keep]


def build_stacks(tmpl, params, lmax):
    """params [M, P] -> layers float32 [5, M, lmax] (Vp, Vs, rho, h, 1/Qs), nlay int32 [M]."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float32))
    M = params.shape[0]
    out = np.zeros((5, M, lmax), np.float32)
    nl = np.zeros(M, np.int32)
    for m in range(M):
        h, vs, vp, rho, qs = build_one(tmpl, params[m].astype(np.float64))
        k = (h > 1e-3)              # models.py:20
        h, vs, vp, rho, qs = h[k], vs[k], vp[k], rho[k], qs[k]
        n = len(h)
        if n > lmax:
            raise ValueError("lmax too small")
        out[0, m, :n] = vp; out[1, m, :n] = vs; out[2, m, :n] = rho; out[3, m, :n] = h; out[4, m, :n] = 1.0 / qs
        nl[m] = n
    return out, nl


def grids_one(tmpl, p):
    """Fine grid (z, vs, group class) of models.py:72-91 without the reference mantle (the default of
    Model1D.seisPropGrids, which is what the prior checks look at)."""
    z_all, vs_all, vp_all, rho_all, qs_all, cls = [], [], [], [], [], []
    z0 = -max(tmpl.topo, 0.0)
    for g in tmpl.groups:
        if g.kind == S.G_REFMANTLE:
            continue
        hv = p[g.h_param] if g.h_param >= 0 else g.h_fixed
        top = z_all[-1] if z_all else 0.0
        H = float(hv) if (g.h_mode == 0 or not z_all) else float(hv) - top
        N = nfine(g.nfine_rule, g.nfine, H)
        z = np.linspace(0, H, N + 1)
        coef = np.array([p[g.v_param[i]] if g.v_param[i] >= 0 else g.v_fixed[i] for i in range(g.ncoef)], dtype=np.float64)
        if g.kind == S.G_WATER:
            vs = np.zeros(N + 1)
        elif g.kind == S.G_CONST:
            vs = np.full(N + 1, coef[0])
        elif g.kind == S.G_LINEAR:
            vs = np.linspace(coef[0], coef[1], N + 1)
        elif g.kind == S.G_BSPLINE:
            vs = coef @ bspl_basis(N + 1, g.ncoef)
        else:
            vs = np.full(N + 1, (0.02 * H ** 2 + 1.27 * H + 0.29 * 0.1) / (H + 0.29))
        if z[-1] - z[0] < 0.01:
            continue
        z_all += list(z + z0); vs_all += list(vs); cls += [g.gclass] * (N + 1)
        z0 = z_all[-1]
    return np.array(z_all), np.array(vs_all), np.array(cls)


def priors(tmpl, p):
    """Bits of the violated rules of CascadiaPrism.isgood (reference models.py:294-360)."""
    z, vs, cls = grids_one(tmpl, np.asarray(p, dtype=np.float64))
    eps = np.finfo(float).eps
    bad = 0
    for i in np.where(cls[1:] != cls[:-1])[0]:
        if vs[i + 1] < vs[i]:
            bad |= S.P_JUMP
    if np.any(vs > 4.9):
        bad |= S.P_VSMAX
    for c in (S.C_SEDIMENT, S.C_CRUST):
        v = vs[cls == c]
        if not np.all(np.diff(v) >= eps):
            bad |= S.P_MONO
    zm, vm = z[cls == S.C_MANTLE], vs[cls == S.C_MANTLE]
    if len(vm) >= 2 and (vm[-1] - vm[-2]) / (zm[-1] - zm[-2]) <= 0:
        bad |= S.P_BOTTOM
    return bad


def _argrel(v, greater):
    """scipy.signal.argrelmax / argrelmin (order 1, mode 'clip'): strict comparison with both neighbours, the
    neighbours of the end points being the end points themselves -- so the ends are never extrema."""
    v = np.asarray(v, dtype=np.float64)
    if len(v) < 3:
        return np.zeros(0, dtype=int)
    l, r, c = v[:-2], v[2:], v[1:-1]
    m = (c > l) & (c > r) if greater else (c < l) & (c < r)
    return np.nonzero(m)[0] + 1


def ricker(points, a):
    """scipy.signal.ricker (SciPy <= 1.14, _wavelets.py; removed in 1.15)."""
    A = 2 / (np.sqrt(3 * a) * (np.pi ** 0.25))
    vec = np.arange(0, points) - (points - 1.0) / 2
    return A * (1 - vec ** 2 / a ** 2) * np.exp(-vec ** 2 / (2 * a ** 2))


def cwt_ricker(data, width):
    """scipy.signal.cwt(data, ricker, [width])[0] (SciPy <= 1.14)."""
    n = min(10 * width, len(data))
    return np.convolve(data, ricker(n, width)[::-1], mode="same")


def priors_ocean(tmpl, p):
    """Bits of the violated rules of CascadiaOcean.isgood (reference models.py:571-677), as that code behaves:
    `grp` is a Python LIST there, so `grp[1:] != grp[:-1]` is one bool (the jump rule only ever compares the first
    two grid points, models.py:586-588) and `vs[grp == 'sediment']` is an empty selection (the monotonicity rules
    of models.py:591-594 never fire)."""
    z, vs, cls = grids_one(tmpl, np.asarray(p, dtype=np.float64))
    bad = 0
    vm, zm = vs[cls == S.C_MANTLE], z[cls == S.C_MANTLE]
    if np.any(vs[cls == S.C_SEDIMENT] < 0.2):
        bad |= S.P_SEDMIN
    if len(vs) >= 2 and vs[1] < vs[0]:
        bad |= S.P_FIRSTPAIR
    if (vs[-1] - vs[-2]) / (z[-1] - z[-2]) <= 0:
        bad |= S.P_BOTTOM
    ext = np.sort(np.append(_argrel(vm, True), _argrel(vm, False)))
    if len(ext) > 1 and np.any(np.abs(np.diff(vm[ext])) > 0.1 * vm.mean()):
        bad |= S.P_OSCI
    if len(_argrel(vm, True)) > 0:
        bad |= S.P_LOCALMAX
    slope = np.diff(vm) / np.diff(zm)
    if slope.min() < slope[0] * 1.5:
        bad |= S.P_SLOPE
    width = 30 // (zm[1] - zm[0])
    cw = cwt_ricker(vm - np.interp(zm, [zm[0], zm[-1]], [vm[0], vm[-1]]), width)
    e2 = np.sort(np.append(_argrel(cw, True), _argrel(cw, False)))
    if np.any(np.abs(np.diff(cw[e2])) > 0.3):
        bad |= S.P_CWT
    return bad
