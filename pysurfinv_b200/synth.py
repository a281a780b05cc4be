"""Synthetic layered-model generators for the benchmark and parity workloads.

Host-side numpy only.  The shapes follow the reference's layer parameterisation rules so the stacks
look like what ``Model1D.seisPropLayers`` (reference models.py:93-102) hands to ``fast_surf``:

* sediment: 1 layer, Vp = 2 Vs, rho = quartic in Vs, Qs = 80        (layers.py:139-156)
* crust:    15 fine layers for 20 < H <= 60 km, cubic B-spline Vs, Vp = 1.8 Vs, Qs = 600 (layers.py:158-189)
* mantle:   60 fine layers down to 200 km, B-spline Vs, Vp = 1.76 Vs, rho = 3.4268+(Vs-4.5)/4.5, Qs = 150
* grid values are averaged to layer mid-points (models.py:93-102); a half-space with the deepest
  grid values closes the stack  ->  n = 1 + 15 + 60 + 1 = 77 layers.

Layout returned: ``layers`` float32 [5, M, Lmax] in the order (Vp, Vs, rho, h, 1/Qs) -- the argument
order of ``fast_surf.fast_surf`` (fast_surf.pyf:6-19) -- and ``nlay`` int32 [M].
"""
import numpy as np

DEFAULT_SEED = 20261018


def bspline_basis(z, ncoef, order=4, alpha=2.0):
    """Clamped B-spline basis [ncoef, len(z)] on z[0]..z[-1] with geometrically growing interior knots
    (same family as reference layers.py:4-45: cubic once ncoef >= 4, knot spacing ratio alpha)."""
    z = np.asarray(z, dtype=np.float64)
    if ncoef == 1:
        return np.ones((1, z.size))
    if ncoef == 2:
        u = (z - z[0]) / (z[-1] - z[0])
        return np.stack([1 - u, u])
    order = min(order, ncoef)
    nint = ncoef - order  # number of interior knots
    if nint > 0:
        w = alpha ** np.arange(nint + 1)
        inner = np.cumsum(w)[:-1] / w.sum()
    else:
        inner = np.zeros(0)
    knots = np.concatenate([np.zeros(order), inner, np.ones(order)])
    u = (z - z[0]) / (z[-1] - z[0])
    nb = len(knots) - 1
    B = np.zeros((nb, u.size))
    for i in range(nb):
        if knots[i + 1] > knots[i]:
            B[i] = (u >= knots[i]) & (u < knots[i + 1])
    last = np.max(np.nonzero(knots[1:] > knots[:-1])[0])
    B[last, u >= 1.0] = 1.0
    for k in range(2, order + 1):
        Bn = np.zeros((nb - k + 1, u.size))
        for i in range(nb - k + 1):
            d1 = knots[i + k - 1] - knots[i]
            d2 = knots[i + k] - knots[i + 1]
            if d1 > 0:
                Bn[i] += (u - knots[i]) / d1 * B[i]
            if d2 > 0:
                Bn[i] += (knots[i + k] - u) / d2 * B[i + 1]
        B = Bn
    return B[:ncoef]


def _rho_quartic(vs):
    return 1.22679 + 1.53201 * vs - 0.83668 * vs * vs + 0.20673 * vs ** 3 - 0.01656 * vs ** 4


def crustal_models(M, seed=DEFAULT_SEED, n_crust=15, n_mantle=60, zmax=200.0, water=False, lvz=False):
    """Config-2 workload: M random sediment + crust + mantle stacks (n = 77, or 78 with a water layer).
    lvz=True: the crustal coefficients are NOT sorted (low-velocity zones inside the crust, which the prior of
    the reference's model classes forbids but fast_surf itself accepts) and the mantle range is wider (stronger
    inversions, half-spaces slower than the layers above them)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Hsed = rng.uniform(0.5, 4.0, M)
    Vsed = rng.uniform(1.0, 2.5, M)
    Hcr = rng.uniform(20.0, 45.0, M)
    if lvz:
        ccr = rng.uniform(3.0, 4.1, (M, 4))
        cma = rng.uniform(3.9, 4.8, (M, 5))
    else:
        ccr = np.sort(rng.uniform(3.2, 4.0, (M, 4)), axis=1)          # monotone increasing crust
        cma = rng.uniform(4.1, 4.7, (M, 5))
    Hw = rng.uniform(0.5, 4.0, M) if water else None

    Bc = bspline_basis(np.linspace(0, 1, n_crust + 1), 4)          # [4, 16]
    Bm = bspline_basis(np.linspace(0, 1, n_mantle + 1), 5)         # [5, 61]
    vs_c = ccr @ Bc                                                # [M, 16] grid values
    vs_m = cma @ Bm                                                # [M, 61]
    mid = lambda g: 0.5 * (g[:, 1:] + g[:, :-1])

    h_c = np.repeat((Hcr / n_crust)[:, None], n_crust, 1)
    z_top_m = Hsed + Hcr + (Hw if water else 0.0)
    h_m = np.repeat(((zmax - z_top_m) / n_mantle)[:, None], n_mantle, 1)

    vs = np.concatenate([Vsed[:, None], mid(vs_c), mid(vs_m), vs_m[:, -1:]], 1)
    vp = np.concatenate([2.0 * Vsed[:, None], mid(1.8 * vs_c), mid(1.76 * vs_m), 1.76 * vs_m[:, -1:]], 1)
    rho_m = 3.4268 + (vs_m - 4.5) / 4.5
    rho = np.concatenate([_rho_quartic(Vsed)[:, None], mid(_rho_quartic(vs_c)), mid(rho_m), rho_m[:, -1:]], 1)
    h = np.concatenate([Hsed[:, None], h_c, h_m, np.full((M, 1), 10.0)], 1)
    qs = np.concatenate([np.full((M, 1), 80.0), np.full((M, n_crust), 600.0),
                         np.full((M, n_mantle + 1), 150.0)], 1)
    if water:
        vs = np.concatenate([np.zeros((M, 1)), vs], 1)
        vp = np.concatenate([np.full((M, 1), 1.475), vp], 1)
        rho = np.concatenate([np.full((M, 1), 1.027), rho], 1)
        h = np.concatenate([Hw[:, None], h], 1)
        qs = np.concatenate([np.full((M, 1), 10000.0), qs], 1)
    n = vs.shape[1]
    layers = np.stack([vp, vs, rho, h, 1.0 / qs]).astype(np.float32)
    return np.ascontiguousarray(layers), np.full(M, n, dtype=np.int32)


def hand_models(M, seed=DEFAULT_SEED + 1):
    """True 3-layer + half-space stacks (n = 4): exercises the un-clamped ndiv = 5 subdivision path."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h = np.stack([rng.uniform(1.0, 4.0, M), rng.uniform(10.0, 20.0, M), rng.uniform(10.0, 25.0, M),
                  np.full(M, 10.0)], 1)
    vs = np.stack([rng.uniform(1.5, 2.8, M), rng.uniform(3.2, 3.6, M), rng.uniform(3.7, 4.0, M),
                   rng.uniform(4.3, 4.7, M)], 1)
    vp = vs * np.array([2.0, 1.75, 1.75, 1.78])
    rho = np.concatenate([_rho_quartic(vs[:, :3]), 3.4268 + (vs[:, 3:] - 4.5) / 4.5], 1)
    qs = np.broadcast_to(np.array([80.0, 600.0, 600.0, 150.0]), (M, 4))
    layers = np.stack([vp, vs, rho, h, 1.0 / qs]).astype(np.float32)
    return np.ascontiguousarray(layers), np.full(M, 4, dtype=np.int32)


def ragged_models(M, seed=DEFAULT_SEED + 2, lmax=96):
    """Mixed stack depths (4 .. lmax layers, some with a water layer) padded to a common Lmax."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.zeros((5, M, lmax), dtype=np.float32)
    nl = np.zeros(M, dtype=np.int32)
    a, na = crustal_models(M, seed + 10)
    b, nb = hand_models(M, seed + 11)
    w, nw = crustal_models(M, seed + 12, water=True)
    f, nf = crustal_models(M, seed + 13, n_crust=30, n_mantle=60)
    pick = rng.integers(0, 4, M)
    for i in range(M):
        src, n = ((a, na), (b, nb), (w, nw), (f, nf))[pick[i]]
        n = int(n[i])
        out[:, i, :n] = src[:, i, :n]
        nl[i] = n
    return out, nl


def log_periods(K=40, tmin=8.0, tmax=80.0):
    """K log-spaced periods, ascending (config 2)."""
    return np.exp(np.linspace(np.log(tmin), np.log(tmax), K)).astype(np.float32)
