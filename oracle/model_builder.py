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
    crust_h = 0.0
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
        elif g.kind == S.G_HYBRID:
            age = p[g.age_param] if g.age_param >= 0 else g.age_fixed
            vs, _vp, _rho, qs_h = hybrid_mantle(z, coef, float(age), crust_h, z0, Tp=g.tp, period=g.period,
                                               q_age=None if g.q_age < 0 else g.q_age)
        else:   # reference mantle, layers.py:267-285
            vs0 = vs_all[-1]
            vs = np.linspace(vs0, vs0 + H * g.slope, N + 1)
        vp = g.vp_a * vs + g.vp_b
        rho = rho_of(g.rho_rule, g.rho_const, vs, vp)
        qs = np.full(N + 1, g.qs, dtype=np.float64)
        if g.kind == S.G_HYBRID:
            qs = qs_h
        if g.gclass == S.C_CRUST and H / N > 0.01:
            crust_h += H
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
    crust_h = 0.0
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
        elif g.kind == S.G_HYBRID:
            age = p[g.age_param] if g.age_param >= 0 else g.age_fixed
            vs = hybrid_mantle(z, coef, float(age), crust_h, z0, Tp=g.tp, period=g.period, q_age=None if g.q_age < 0 else g.q_age)[0]
        else:
            vs = np.full(N + 1, (0.02 * H ** 2 + 1.27 * H + 0.29 * 0.1) / (H + 0.29))
        if g.gclass == S.C_CRUST and H / N > 0.01:
            crust_h += H
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


# ---------------------------------------------------------------------------------------------- thermal mantle (f-4)
# OceanMantleHybrid (reference layers.py:297-363): half-space-cooling temperature (ThermSeis.HSCM, ThermSeis.py:56-101)
# converted to Vs by the mineral-physics relations of OceanSeisRitz (ThermSeis.py:103-176), plus a B-spline
# perturbation below the depth where melting starts, joined by a not-a-knot cubic spline (scipy CubicSpline); Qs from
# the anelasticity model of OceanSeisRuan / OceanSeisYaTa (ThermSeis.py:320-448).
def hscm(age, zdeps, Tp=1325.0, kappa=1e-6, rho0=3.43e3):
    """ThermSeis.HSCM: returns (T [K], P [Pa], rho [kg/m^3]) on zdeps [km]."""
    from scipy.special import erf
    zdeps = np.asarray(zdeps, dtype=np.float64)
    P = 3.4e3 * 9.8 * zdeps * 1000                     # TherModel._calP
    T0, Da = 0.0, 0.4
    den = 2 * np.sqrt(age * 365 * 24 * 3600 * 1 * (kappa / 1e-6))

    def f(z):
        return erf(z * 1e3 / den)

    def g(z):
        dz = 0.001; fz = f(z); dfz = (f(z + dz) - fz) / dz + 1e-10
        return fz / dfz - z - (Tp - T0) / Da
    z0, z1 = 0.0, 400.0
    while z1 - z0 > 0.01:
        z2 = (z1 + z0) / 2
        if g(z2) < 0:
            z0 = z2
        else:
            z1 = z2
    Tm = (Da * z1 + Tp - T0) / f(z1) + T0
    T = (Tm - T0) * f(zdeps) + T0
    above = np.where(zdeps > z0)[0]
    if len(above):
        a = above[0]
        T_ad = Tp + zdeps * Da
        if a == 0:
            T = T_ad
        else:
            T[a:] = T_ad[a:]
    T = T + 273.15
    rho = rho0 * (1 - 4.4e-5 * (T - (500 + 273.15))) * (1 + 6.12e-12 * (P - 0.6e9))    # TherModel._calRho(rho0)
    return T, P, rho


_RITZ = [  # rho0, rho_X, K0, K_T, K_P, K_X, mu0, mu_T, mu_P, mu_X, alpha0..3   (ThermSeis.py:108-129)
    (3.222e3, 1.182e3, 129, -16e-3, 4.2, 0, 82, -14e-3, 1.4, -30, 0.2010e-4, 0.1390e-7, 0.1627e-2, -0.3380),
    (3.198e3, 0.804e3, 111, -12e-3, 6.0, -10, 81, -11e-3, 2.0, -29, 0.3871e-4, 0.0446e-7, 0.0343e-2, -1.7278),
    (3.280e3, 0.377e3, 105, -13e-3, 6.2, 13, 67, -10e-3, 1.7, -6, 0.3206e-4, 0.0811e-7, 0.1347e-2, -1.8167),
    (3.578e3, 0.702e3, 198, -28e-3, 5.7, 12, 108, -12e-3, 0.8, -24, 0.6969e-4, -0.0108e-7, -3.0799e-2, 5.0395),
    (3.565e3, 0.758e3, 173, -21e-3, 4.9, 7, 92, -10e-3, 1.4, -7, 0.0991e-4, 0.1165e-7, 1.0624e-2, -2.5000)]
_RITZ_W = [0.75, 0.21, 0.035, 0, 0.005]


def ritz_vs(T, P_pa, X=0.1):
    """OceanSeisRitz._pt2vs with RhoType 'raw' (ThermSeis.py:132-176): Vs [km/s]."""
    P = P_pa / 1e9
    T0, P0 = 273.15, 101.325e-6
    mus, Ks, rhos = [], [], []
    for (rho0, rho_X, K0, K_T, K_P, K_X, mu0, mu_T, mu_P, mu_X, a0, a1, a2, a3) in _RITZ:
        alpha = a0 + a1 * T + a2 * T ** (-1) + a3 * T ** (-2)
        rho0X = rho0 * rho_X / 1e3
        mu = mu0 + (T - T0) * mu_T + (P - P0) * mu_P + X * mu_X
        K = K0 + (T - T0) * K_T + (P - P0) * K_P + X * K_X
        mus.append(mu); Ks.append(K); rhos.append(rho0X * (1 - alpha * (T - T0) + (P - P0) / K))
    w = np.array(_RITZ_W)[:, None]
    mus, rhos = np.array(mus), np.array(rhos)
    rho = (w * rhos).sum(axis=0)
    mu = 0.5 * ((w * mus).sum(axis=0) + 1 / ((w / mus).sum(axis=0)))
    return np.sqrt(mu * 1e9 / rho) / 1000


def ruan_qs(T, P, period):
    """OceanSeisRuan: Qs = J1 / J2 of OceanSeisYaTa._anel with the Ruan2018 solidus (ThermSeis.py:320-448)."""
    from scipy.special import erf
    Pg = P / 1e9
    Tn = T / (-5.1 * Pg ** 2 + 92.5 * Pg + 1120.6 + 273.15)
    Aeta = np.where(Tn < 0.94, 1.0, np.where(Tn < 1, np.exp(-(Tn - 0.94) / (Tn - Tn * 0.94) * np.log(5)), 1 / 5))
    mu_U = (72.45 - 0.01094 * (T - 273.15) + 1.75 * P * 1e-9) * 1e9
    eta = 6.22e21 * np.exp(4.625e5 / 8.314 * (1 / T - 1 / (1200 + 273.15))) * np.exp(7.913e-6 / 8.314 * (P / T - 1.5e9 / (1200 + 273.15))) * Aeta
    tau_ns = period / (2 * np.pi * (eta / mu_U))
    A_P = np.where(Tn < 0.91, 0.01, np.where(Tn < 0.96, 0.01 + 0.4 * (Tn - 0.91), 0.03))
    sig = np.where(Tn < 0.92, 4.0, np.where(Tn < 1, 4 + 37.5 * (Tn - 0.92), 7.0))
    A_B, tau_np, alpha = 0.664, 6e-5, 0.38
    J1 = 1 + A_B * (tau_ns ** alpha) / alpha + np.sqrt(2 * np.pi) / 2 * A_P * sig * (1 - erf(np.log(tau_np / tau_ns) / (np.sqrt(2) * sig)))
    J2 = np.pi / 2 * A_B * (tau_ns ** alpha) + np.pi / 2 * (A_P * np.exp(-((np.log(tau_np / tau_ns) / (np.sqrt(2) * sig)) ** 2))) + tau_ns
    return J1 / J2


def cubic_spline_not_a_knot(xs, ys, x):
    """scipy.interpolate.CubicSpline(xs, ys)(x) with the default not-a-knot ends (scipy/interpolate/_cubic.py):
    tridiagonal system for the knot slopes, piecewise cubic evaluation, extrapolation with the end pieces."""
    xs, ys, x = np.asarray(xs, float), np.asarray(ys, float), np.asarray(x, float)
    n = len(xs)
    dx = np.diff(xs); slope = np.diff(ys) / dx
    if n == 2:
        s = np.array([slope[0], slope[0]])
    elif n == 3:
        # parabola through the three points
        A = np.array([[1.0, 1.0, 0.0], [dx[1], 2 * (dx[0] + dx[1]), dx[0]], [0.0, 1.0, 1.0]])
        b = np.array([2 * slope[0], 3 * (dx[0] * slope[1] + dx[1] * slope[0]), 2 * slope[1]])
        s = np.linalg.solve(A, b)
    else:
        A = np.zeros((n, n)); b = np.zeros(n)
        for i in range(1, n - 1):
            A[i, i - 1] = dx[i]; A[i, i] = 2 * (dx[i - 1] + dx[i]); A[i, i + 1] = dx[i - 1]
            b[i] = 3 * (dx[i] * slope[i - 1] + dx[i - 1] * slope[i])
        A[0, 0] = dx[1]; A[0, 1] = xs[2] - xs[0]
        d = xs[2] - xs[0]
        b[0] = ((dx[0] + 2 * d) * dx[1] * slope[0] + dx[0] ** 2 * slope[1]) / d
        A[-1, -1] = dx[-2]; A[-1, -2] = xs[-1] - xs[-3]
        d = xs[-1] - xs[-3]
        b[-1] = (dx[-1] ** 2 * slope[-2] + (2 * d + dx[-1]) * dx[-2] * slope[-1]) / d
        s = np.linalg.solve(A, b)
    t = (s[:-1] + s[1:] - 2 * slope) / dx
    c0, c1, c2, c3 = t / dx, (slope - s[:-1]) / dx - t, s[:-1], ys[:-1]
    i = np.clip(np.searchsorted(xs, x, side="right") - 1, 0, n - 2)
    h = x - xs[i]
    return ((c0[i] * h + c1[i]) * h + c2[i]) * h + c3[i]


def hybrid_mantle(z, coefs, therm_age, crust_h, z_top, Tp=1325.0, period=1.0, q_age=None):
    """OceanMantleHybrid._calVs / _calOthers (layers.py:302-363) on the group's grid z (from 0): (vs, vp, rho, qs)."""
    z = np.asarray(z, dtype=np.float64)
    age = max(1e-3, therm_age)
    T, P, _ = hscm(age, crust_h + z, Tp=Tp)
    vs_t = ritz_vs(T, P)
    # depth where melting starts (layers.py:312-319)
    Tm_, Pm_, _ = hscm(age, np.linspace(0, 200, 200))
    sol = -5.1 * (Pm_ / 1e9) ** 2 + 92.5 * (Pm_ / 1e9) + 1120.6 + 273.15
    idx = np.where(Tm_ > 0.92 * sol)[0]
    z_melt = (np.linspace(0, 200, 200)[idx[0]] if len(idx) else 200.0) - crust_h
    nb = len(coefs) + 1
    y2 = np.concatenate([[0.0], coefs]) @ bspl_basis(len(z), nb) + vs_t
    xL, xH = z_melt, (z_melt + crust_h) * 1.7 - crust_h
    keep1, keep2 = z < xL, z > xH
    vs = cubic_spline_not_a_knot(np.concatenate([z[keep1], z[keep2]]), np.concatenate([vs_t[keep1], y2[keep2]]), z)
    Tq, Pq, _ = hscm(max(1e-3, therm_age if q_age is None else q_age), z_top + z)
    qs = np.minimum(ruan_qs(Tq, Pq, period), 5000.0)
    return vs, 1.76 * vs, 3.4268 + (vs - 4.5) / 4.5, qs
