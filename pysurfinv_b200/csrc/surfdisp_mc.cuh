// surfdisp_mc.cuh -- the callers on either side of the solver, included by surfdisp_kernels.cu:
// model assembly from parameter vectors, the prior rules of the reference's model classes, bounded-Gaussian
// proposals, misfit + Metropolis rule + chain-track rows.
//
// One WARP per model / chain, lanes over the fine grid points of a layer group (the reference evaluates every
// group on linspace(0, H, N + 1), layers.py:111-116): coalesced stores of the five layer rows, the B-spline value
// of a grid point computed by its own lane, the prior rules as warp votes.  All arithmetic in double like the
// reference's numpy code; results are rounded to float32 when stored (that is what f2py does to the arrays handed
// to fast_surf, fast_surf.pyf:6-19).
#pragma once

namespace mcdev {

constexpr int kMaxMantleGrid = 128;      // grid points of the mantle-class groups kept for the CascadiaOcean rules
constexpr double kEps = 2.220446049250313e-16;

// value at z in [0, 1] of sum_i coef_i B_i(z) with the basis of reference layers.py:4-45: degree 3 (n = 3) or
// 4 (n >= 4) de Boor recursion on the knot vector [-eps x (deg-1), 0, geometric interior knots, 1, 1+eps ...].
// The reference runs the recursion over all n + deg - 1 columns; only the deg columns around the knot interval of z
// are non-zero, and a term whose lower-order factor is exactly zero adds nothing: the recursion below visits the
// non-zero terms only, in the reference's order (same operations on them, bit-identical values, a quarter of the
// divisions).  Knots are generated from their index (no array, no local memory).
struct Knots {
  int n, deg;
  double q[SURFDISP_MAX_COEF];     // interior knots 2^k / (2^(m+1) - 1), k < m = n - deg
  __device__ double at(int i) const {
    if (i < deg - 1) return -kEps;
    if (i == deg - 1) return 0.0;
    if (i < n) return q[i - deg];
    return (i == n) ? 1.0 : 1.0 + kEps;
  }
};

__device__ Knots make_knots(int n) {
  Knots k;
  k.n = n; k.deg = 3 + (n >= 4);
  const int m = n - k.deg;
  const double den = (double)((1 << (m + 1)) - 1);
#pragma unroll
  for (int kk = 0; kk < SURFDISP_MAX_COEF; ++kk) k.q[kk] = (kk < m) ? (double)(1 << kk) / den : 2.0;
  return k;
}

__device__ double bspl_profile(const double* coef, const Knots& kn, double z) {
  const int n = kn.n, deg = kn.deg, nc = n + deg - 1;
  // knot interval of z: x[k] <= z < x[k+1]
  int k = deg - 1;
#pragma unroll
  for (int kk = 0; kk < SURFDISP_MAX_COEF; ++kk) k += (kk < n - deg && z >= kn.q[kk]) ? 1 : 0;
  if (z >= 1.0) k = n;
  // w[j] = value of column k - r - 1 + j after round r (columns k-r-1 .. k can be non-zero)
  double w[6] = {0.0, 1.0, 0.0, 0.0, 0.0, 0.0};   // before round 0: column k is 1 (w[1]); w[0] = column k-1 = 0
  for (int r = 0; r < deg - 1; ++r) {
    double nw[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    // new columns i = k-r-1 .. k; old column i sits at w[i - (k-r)] + 1 shift: old window starts at column k-r (w[1])
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (j > r + 1) continue;
      const int i = k - r - 1 + j;
      if (i < 0 || i > nc - r - 2) continue;
      const double b_i = w[j], b_i1 = w[j + 1];      // old columns i and i+1
      double col = 0.0;
      if (b_i != 0.0) { const double xi = kn.at(i), d1 = kn.at(i + r + 1) - xi; if (d1 != 0.0) col += b_i * (z - xi) / d1; }
      if (b_i1 != 0.0) { const double xe = kn.at(i + r + 2), d2 = xe - kn.at(i + 1); if (d2 != 0.0) col += b_i1 * (xe - z) / d2; }
      nw[j + 1] = col;
    }
    // the window of the next round starts one column lower: new column k-r-1 (nw[1]) becomes w[1]
#pragma unroll
    for (int j = 0; j < 6; ++j) w[j] = nw[j];
  }
  // columns k-deg+1 .. k sit in w[1 .. deg]
  double v = 0.0;
#pragma unroll
  for (int j = 1; j <= 5; ++j) {
    const int i = k - (deg - 1) + (j - 1);
    if (j <= deg && i >= 0 && i < n) v += coef[i] * w[j];
  }
  return v;
}

__device__ __forceinline__ double stack_rho(int rule, double cst, double vs, double vp) {
  if (rule == SURFDISP_R_QUARTIC) return 1.22679 + 1.53201 * vs - 0.83668 * vs * vs + 0.20673 * vs * vs * vs - 0.01656 * vs * vs * vs * vs;
  if (rule == SURFDISP_R_OCEAN) return 0.541 + 0.3601 * vp;
  if (rule == SURFDISP_R_MANTLE) return 3.4268 + (vs - 4.5) / 4.5;
  return cst;
}

struct WarpScratch {
  double vm[kMaxMantleGrid];    // Vs and depth of the mantle-class grid points
  double zm[kMaxMantleGrid];
  double cw[kMaxMantleGrid];    // detrended profile, then its wavelet transform (thermal group: spline abscissae)
  double dt[kMaxMantleGrid];    //                                                (thermal group: spline ordinates)
  double hv[kMaxMantleGrid];    // thermal group: Vs, Qs and spline slopes on its grid
  double hq[kMaxMantleGrid];
  double hs[kMaxMantleGrid];
  float q[64];                  // the proposal being tested
};

// ------------------------------------------------------------------------------------ thermal mantle (OceanMantleHybrid)
// ThermSeis.HSCM (ThermSeis.py:56-101): half-space cooling with an adiabat below the depth where the conductive
// gradient falls to 0.4 K/km; the mantle temperature Tm follows from a bisection on that depth (the reference's
// finite-difference derivative and its 0.01 km stopping rule included).
struct Hscm { double den, Tm, z_ad, Tp; };

// Called by all lanes of a warp.  The reference bisects [0, 400] km until the interval is <= 0.01 km: always 16
// halvings, every midpoint a multiple of 400 / 2^16 (exact in binary).  Here five halvings at a time: the lanes
// evaluate the 31 interior points of a 32-section, then every lane walks the bisection's own path through those
// signs (so the result is the reference's even where the signs are not monotone); the 16th halving on its own.
__device__ Hscm hscm_setup(double age, double Tp) {
  Hscm h;
  h.Tp = Tp;
  h.den = 2.0 * sqrt(age * 365.0 * 24.0 * 3600.0 * 1.0 * (1e-6 / 1e-6));
  const double T0 = 0.0, Da = 0.4;
  const int lane = threadIdx.x & 31;
  auto below = [&](double z2) -> bool {
    const double fz = erf(z2 * 1e3 / h.den), dfz = (erf((z2 + 0.001) * 1e3 / h.den) - fz) / 0.001 + 1e-10;
    return fz / dfz - z2 - (Tp - T0) / Da < 0.0;
  };
  double z0 = 0.0, w = 400.0;
  for (int r = 0; r < 3; ++r) {
    const double sw = w / 32.0;
    const unsigned neg = __ballot_sync(0xffffffffu, lane > 0 && below(z0 + (double)lane * sw));
    int lo = 0, hi = 32;
#pragma unroll
    for (int t = 0; t < 5; ++t) { const int mid = (lo + hi) >> 1; if ((neg >> mid) & 1u) lo = mid; else hi = mid; }
    z0 = z0 + (double)lo * sw;
    w = sw;
  }
  double z1 = z0 + w;
  {
    const double z2 = (z1 + z0) / 2.0;
    if (below(z2)) z0 = z2; else z1 = z2;
  }
  h.Tm = (Da * z1 + Tp - T0) / erf(z1 * 1e3 / h.den) + T0;
  h.z_ad = z0;
  return h;
}
// temperature [K] at depth z [km]; zfirst = first grid depth (a grid that starts below z_ad is all adiabat)
__device__ __forceinline__ double hscm_T(const Hscm& h, double z) {
  const double T = (z > h.z_ad) ? h.Tp + z * 0.4 : h.Tm * erf(z * 1e3 / h.den);
  return T + 273.15;
}
__device__ __forceinline__ double hscm_P(double z) { return 3.4e3 * 9.8 * z * 1000.0; }

// OceanSeisRitz._pt2vs, RhoType 'raw' (ThermSeis.py:132-176): Voigt-Reuss-Hill shear modulus of five minerals
__constant__ double kRitz[5][14] = {
    {3.222e3, 1.182e3, 129, -16e-3, 4.2, 0, 82, -14e-3, 1.4, -30, 0.2010e-4, 0.1390e-7, 0.1627e-2, -0.3380},
    {3.198e3, 0.804e3, 111, -12e-3, 6.0, -10, 81, -11e-3, 2.0, -29, 0.3871e-4, 0.0446e-7, 0.0343e-2, -1.7278},
    {3.280e3, 0.377e3, 105, -13e-3, 6.2, 13, 67, -10e-3, 1.7, -6, 0.3206e-4, 0.0811e-7, 0.1347e-2, -1.8167},
    {3.578e3, 0.702e3, 198, -28e-3, 5.7, 12, 108, -12e-3, 0.8, -24, 0.6969e-4, -0.0108e-7, -3.0799e-2, 5.0395},
    {3.565e3, 0.758e3, 173, -21e-3, 4.9, 7, 92, -10e-3, 1.4, -7, 0.0991e-4, 0.1165e-7, 1.0624e-2, -2.5000}};
__constant__ double kRitzW[5] = {0.75, 0.21, 0.035, 0.0, 0.005};

__device__ double ritz_vs(double T, double P_pa) {
  const double P = P_pa / 1e9, T0 = 273.15, P0 = 101.325e-6, X = 0.1;
  double srho = 0.0, smu = 0.0, simu = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const double* d = kRitz[i];
    const double alpha = d[10] + d[11] * T + d[12] / T + d[13] / (T * T);
    const double rho0X = d[0] * d[1] / 1e3;
    const double mu = d[6] + (T - T0) * d[7] + (P - P0) * d[8] + X * d[9];
    const double K = d[2] + (T - T0) * d[3] + (P - P0) * d[4] + X * d[5];
    const double rho = rho0X * (1.0 - alpha * (T - T0) + (P - P0) / K);
    srho += kRitzW[i] * rho; smu += kRitzW[i] * mu; simu += kRitzW[i] / mu;
  }
  const double mu = 0.5 * (smu + 1.0 / simu);
  return sqrt(mu * 1e9 / srho) / 1000.0;
}

// OceanSeisRuan: Qs = J1 / J2 of OceanSeisYaTa._anel with the Ruan2018 solidus (ThermSeis.py:320-448)
__device__ double ruan_qs(double T, double P, double period) {
  const double Pg = P / 1e9;
  const double Tn = T / (-5.1 * Pg * Pg + 92.5 * Pg + 1120.6 + 273.15);
  const double Aeta = (Tn < 0.94) ? 1.0 : ((Tn < 1.0) ? exp(-(Tn - 0.94) / (Tn - Tn * 0.94) * log(5.0)) : 1.0 / 5.0);
  const double mu_U = (72.45 - 0.01094 * (T - 273.15) + 1.75 * P * 1e-9) * 1e9;
  const double eta = 6.22e21 * exp(4.625e5 / 8.314 * (1.0 / T - 1.0 / (1200.0 + 273.15))) *
                     exp(7.913e-6 / 8.314 * (P / T - 1.5e9 / (1200.0 + 273.15))) * Aeta;
  const double tau_ns = period / (2.0 * 3.141592653589793 * (eta / mu_U));
  const double A_P = (Tn < 0.91) ? 0.01 : ((Tn < 0.96) ? 0.01 + 0.4 * (Tn - 0.91) : 0.03);
  const double sig = (Tn < 0.92) ? 4.0 : ((Tn < 1.0) ? 4.0 + 37.5 * (Tn - 0.92) : 7.0);
  const double A_B = 0.664, tau_np = 6e-5, alpha = 0.38;
  const double ta = pow(tau_ns, alpha), lg = log(tau_np / tau_ns) / (sqrt(2.0) * sig);
  const double J1 = 1.0 + A_B * ta / alpha + sqrt(2.0 * 3.141592653589793) / 2.0 * A_P * sig * (1.0 - erf(lg));
  const double J2 = 3.141592653589793 / 2.0 * A_B * ta + 3.141592653589793 / 2.0 * (A_P * exp(-(lg * lg))) + tau_ns;
  return J1 / J2;
}

// OceanMantleHybrid._calVs / _calOthers (layers.py:302-363) for the N + 1 <= 64 grid points of the group, by one warp:
// ws.hv = Vs, ws.hq = Qs.  coef = the len(Vs) perturbation coefficients (the basis has one more, its first coefficient 0).
// WANT_QS = false (prior checks): ws.hq is not filled.
template <bool WANT_QS>
__device__ void hybrid_profile(const SurfdispStackGroup& g, const double* coef, double therm_age, double H, int N,
                               double crust_h, double z_top, WarpScratch& ws) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const double age = fmax(1e-3, therm_age);
  const Hscm hv = hscm_setup(age, g.tp);
  const Hscm hm = (g.tp == 1325.0) ? hv : hscm_setup(age, 1325.0);     // meltStart: HSCM(age) with the default Tp (layers.py:313)
  // depth where melting starts: first of linspace(0, 200, 200) with T > 0.92 x (damp solidus)
  int first = 200;
  for (int i = lane; i < 200; i += 32) {
    const double z = (i == 199) ? 200.0 : (double)i * (200.0 / 199.0);
    const double Pg = hscm_P(z) / 1e9;
    if (hscm_T(hm, z) > 0.92 * (-5.1 * Pg * Pg + 92.5 * Pg + 1120.6 + 273.15)) { first = i; break; }
  }
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(full, first, o));
  const double z_melt = ((first >= 199) ? 200.0 : (double)first * (200.0 / 199.0)) - crust_h;
  const double xL = z_melt, xH = (z_melt + crust_h) * 1.7 - crust_h;
  const double zstep = H / (double)N, ustep = 1.0 / (double)N;
  double pc[SURFDISP_MAX_COEF];
  pc[0] = 0.0;
  for (int i = 0; i < g.ncoef; ++i) pc[i + 1] = coef[i];
  const Knots kn = make_knots(g.ncoef + 1);
  // thermal Vs and perturbed Vs of this lane's grid points (two chunks), compaction of the spline's knots
  double zj[2], vt[2];
  int npts = 0;
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int j = 32 * c + lane;
    const bool act = j <= N;
    zj[c] = (j == N) ? H : (double)j * zstep;
    double y2 = 0.0;
    vt[c] = 0.0;
    if (act) {
      const double zz = crust_h + zj[c];
      vt[c] = ritz_vs(hscm_T(hv, zz), hscm_P(zz));
      y2 = bspl_profile(pc, kn, (j == N) ? 1.0 : (double)j * ustep) + vt[c];
    }
    const bool k1 = act && zj[c] < xL, k2 = act && zj[c] > xH;
    const unsigned km = __ballot_sync(full, k1 || k2);
    if (k1 || k2) { const int idx = npts + __popc(km & ((1u << lane) - 1u)); ws.cw[idx] = zj[c]; ws.dt[idx] = k1 ? vt[c] : y2; }
    npts += __popc(km);
  }
  __syncwarp();
  // knot slopes of scipy's CubicSpline (not-a-knot): tridiagonal system.  Interval widths, secant slopes, diagonal and
  // right-hand side of the interior rows by the lanes (kept in the free upper halves of the arrays: npts <= 64), the
  // two-term elimination by one lane
  const double* xs = ws.cw; const double* ys = ws.dt;
  double* sdx = ws.cw + 64; double* sm = ws.dt + 64;
  if (npts >= 4) {
    for (int i = lane; i < npts - 1; i += 32) { sdx[i] = xs[i + 1] - xs[i]; sm[i] = (ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]); }
    __syncwarp();
    for (int i = 1 + lane; i < npts - 1; i += 32) {
      ws.hv[i] = 2.0 * (sdx[i - 1] + sdx[i]);
      ws.hq[i] = 3.0 * (sdx[i] * sm[i - 1] + sdx[i - 1] * sm[i]);
    }
    __syncwarp();
  }
  if (lane == 0 && npts >= 2) {
    double* sl = ws.hs; double* bd = ws.hv; double* rh = ws.hq;      // slopes, eliminated diagonal, eliminated right-hand side
    if (npts == 2) { sl[0] = sl[1] = (ys[1] - ys[0]) / (xs[1] - xs[0]); }
    else if (npts == 3) {
      // parabola through three points: s0 + s1 = 2 m0, dx1 s0 + 2 (dx0 + dx1) s1 + dx0 s2 = 3 (dx0 m1 + dx1 m0), s1 + s2 = 2 m1
      const double d0 = xs[1] - xs[0], d1 = xs[2] - xs[1], m0 = (ys[1] - ys[0]) / d0, m1 = (ys[2] - ys[1]) / d1;
      const double s1 = (3.0 * (d0 * m1 + d1 * m0) - 2.0 * d1 * m0 - 2.0 * d0 * m1) / (d0 + d1);
      sl[0] = 2.0 * m0 - s1; sl[1] = s1; sl[2] = 2.0 * m1 - s1;
    } else {
      const int n = npts;
      // row 0: [dx1, x2 - x0]
      double d = xs[2] - xs[0];
      bd[0] = sdx[1];
      double cprev = d;
      rh[0] = ((sdx[0] + 2.0 * d) * sdx[1] * sm[0] + sdx[0] * sdx[0] * sm[1]) / d;
      double bprev = bd[0], rprev = rh[0];
      for (int i = 1; i < n - 1; ++i) {       // bd[i], rh[i] hold the row's diagonal and right-hand side
        const double w = sdx[i] / bprev;
        bprev = bd[i] - w * cprev; rprev = rh[i] - w * rprev;
        bd[i] = bprev; rh[i] = rprev;
        cprev = sdx[i - 1];
      }
      d = xs[n - 1] - xs[n - 3];
      {
        const double a = d, b = sdx[n - 2];
        const double r = (sdx[n - 2] * sdx[n - 2] * sm[n - 3] + (2.0 * d + sdx[n - 2]) * sdx[n - 3] * sm[n - 2]) / d;
        const double w = a / bprev;
        bd[n - 1] = b - w * cprev; rh[n - 1] = r - w * rprev;
      }
      double snext = rh[n - 1] / bd[n - 1];
      sl[n - 1] = snext;
      for (int i = n - 2; i >= 0; --i) {
        const double c = (i == 0) ? (xs[2] - xs[0]) : sdx[i - 1];
        snext = (rh[i] - c * snext) / bd[i];
        sl[i] = snext;
      }
    }
  }
  __syncwarp();
  // evaluation on the grid (extrapolation with the end pieces), Qs of the anelastic model
  const double age_q = fmax(1e-3, (g.q_age < 0.0) ? therm_age : g.q_age);
  Hscm hqm = hm;
  if (WANT_QS && age_q != age) hqm = hscm_setup(age_q, 1325.0);
  double vs_out[2], qs_out[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int j = 32 * c + lane;
    vs_out[c] = vt[c]; qs_out[c] = 0.0;
    if (j <= N) {
      if (npts >= 2) {
        int i = 0;
        while (i < npts - 2 && xs[i + 1] <= zj[c]) ++i;
        const double dx = xs[i + 1] - xs[i], m = (ys[i + 1] - ys[i]) / dx, t = (ws.hs[i] + ws.hs[i + 1] - 2.0 * m) / dx;
        const double h = zj[c] - xs[i];
        vs_out[c] = ((t / dx * h + ((m - ws.hs[i]) / dx - t)) * h + ws.hs[i]) * h + ys[i];
      }
      if (WANT_QS) {
        const double zq = z_top + zj[c];
        qs_out[c] = fmin(ruan_qs(hscm_T(hqm, zq), hscm_P(zq), g.period), 5000.0);
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 2; ++c) { const int j = 32 * c + lane; if (j <= N) { ws.hv[j] = vs_out[c]; ws.hq[j] = qs_out[c]; } }
  __syncwarp();
}

// Rules of CascadiaOcean.isgood that look at the mantle profile (reference models.py:603-635), on the nm grid points
// in ws.vm / ws.zm.  scipy.signal.argrelmax / argrelmin: strict comparison with both neighbours, end points never
// extrema; scipy.signal.cwt(data, ricker, [width]) of SciPy <= 1.14: convolution ('same') with the Mexican-hat
// wavelet of min(10 width, len) points.  All lanes return the same bits.
__device__ int ocean_mantle_rules(WarpScratch& ws, int nm) {
  const int lane = threadIdx.x & 31;
  int bad = 0;
  if (nm < 2) return 0;
  // mean, slope rule (models.py:621-623)
  double s = 0.0, smin = 1.0e300;
  for (int i = lane; i < nm; i += 32) s += ws.vm[i];
  for (int i = lane; i < nm - 1; i += 32) smin = fmin(smin, (ws.vm[i + 1] - ws.vm[i]) / (ws.zm[i + 1] - ws.zm[i]));
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); smin = fmin(smin, __shfl_xor_sync(0xffffffffu, smin, o)); }
  const double mean = s / (double)nm;
  const double slope0 = (ws.vm[1] - ws.vm[0]) / (ws.zm[1] - ws.zm[0]);
  if (smin < slope0 * 1.5) bad |= SURFDISP_P_SLOPE;
  // extrema of the profile: local maxima (models.py:616-619), oscillation limit (models.py:603-611)
  {
    // the lanes flag the extrema of their grid points; the (few) flagged ones are then visited in order
    bool anymax = false;
    double prev = 0.0; int have = 0, next = 0; bool osc = false;
    for (int base = 0; base < nm; base += 32) {
      const int i = base + lane;
      bool mx = false, mn = false;
      if (i >= 1 && i < nm - 1) { const double v = ws.vm[i], l = ws.vm[i - 1], r = ws.vm[i + 1]; mx = (v > l && v > r); mn = (v < l && v < r); }
      anymax |= __any_sync(0xffffffffu, mx);
      unsigned ext = __ballot_sync(0xffffffffu, mx || mn);
      while (ext) {
        const double v = ws.vm[base + __ffs(ext) - 1];
        ext &= ext - 1u;
        if (have && fabs(v - prev) > 0.1 * mean) osc = true;
        prev = v; have = 1; ++next;
      }
    }
    if (anymax) bad |= SURFDISP_P_LOCALMAX;
    if (next > 1 && osc) bad |= SURFDISP_P_OSCI;
  }
  // wavelet rule (models.py:626-635)
  {
    const double dz = ws.zm[1] - ws.zm[0];
    const double width = floor(30.0 / dz);
    if (width >= 1.0) {
      const int L = (int)fmin(10.0 * width, (double)nm);
      const double z0 = ws.zm[0], zN = ws.zm[nm - 1], v0 = ws.vm[0], vN = ws.vm[nm - 1];
      const double sl = (vN - v0) / (zN - z0);
      const double A = 2.0 / (sqrt(3.0 * width) * 1.3313353638003897);   // pi^(1/4)
      for (int i = lane; i < nm; i += 32) {
        ws.dt[i] = ws.vm[i] - ((i == nm - 1) ? vN : sl * (ws.zm[i] - z0) + v0);
        if (i < L) {
          const double vec = (double)i - ((double)L - 1.0) / 2.0, x2 = vec * vec, w2 = width * width;
          ws.cw[i] = A * (1.0 - x2 / w2) * exp(-x2 / (2.0 * w2));       // (symmetric: the reversal of cwt is a no-op)
        }
      }
      __syncwarp();
      const int sh = (L - 1) / 2;
      double out[(kMaxMantleGrid + 31) / 32];
      for (int r = 0, i = lane; i < nm; i += 32, ++r) {
        double acc = 0.0;
        for (int k = 0; k < L; ++k) { const int j = i + sh - k; if (j >= 0 && j < nm) acc += ws.dt[j] * ws.cw[k]; }
        out[r] = acc;
      }
      __syncwarp();
      for (int r = 0, i = lane; i < nm; i += 32, ++r) ws.cw[i] = out[r];
      __syncwarp();
      double prev = 0.0; int have = 0; bool big = false;
      for (int base = 0; base < nm; base += 32) {
        const int i = base + lane;
        bool ex = false;
        if (i >= 1 && i < nm - 1) { const double v = ws.cw[i], l = ws.cw[i - 1], r = ws.cw[i + 1]; ex = (v > l && v > r) || (v < l && v < r); }
        unsigned ext = __ballot_sync(0xffffffffu, ex);
        while (ext) {
          const double v = ws.cw[base + __ffs(ext) - 1];
          ext &= ext - 1u;
          if (have && fabs(v - prev) > 0.3) big = true;
          prev = v; have = 1;
        }
      }
      if (big) bad |= SURFDISP_P_CWT;
      __syncwarp();
    }
  }
  return bad;
}

// Assembles one model from its parameter vector, by one warp.  EMIT: write the layers.  Returns (same value in
// every lane) the SURFDISP_P_* bits of the violated prior rules -- CascadiaPrism / CascadiaContinent.isgood
// (models.py:294-360, 385-523) and, where `rules` asks for them, CascadiaOcean.isgood (models.py:571-677) --
// evaluated on the fine grid without the reference mantle like Model1D.seisPropGrids() does by default.
// HYB: the template has a thermal mantle group (compiled out otherwise: its erf / spline code costs registers)
template <bool EMIT, bool HYB>
__device__ int assemble_stack_warp(const SurfdispStackTemplate& t, const float* pm, int rules, int lmax, float* o_vp,
                                   float* o_vs, float* o_rho, float* o_h, float* o_qs, int* nl_out, WarpScratch& ws) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int nl = 0, bad = 0, nm = 0, ngrid = 0;
  bool overflow = false, any = false;
  double z0 = -fmax(t.topo, 0.0);        // depth of the top of the next group (models.py:76-77)
  double ztop_prev = 0.0;                 // bottom depth of the stack so far, for BottomDepth groups (0 if none)
  double last_vs = 0.0, last_vp = 0.0, last_rho = 0.0, last_qs = 0.0;   // deepest grid values so far
  double end_vs1 = 0.0, end_vs2 = 0.0, end_z1 = 0.0, end_z2 = 0.0;       // last two grid points of the model proper
  double first_vs0 = 0.0, first_vs1 = 0.0;
  double crust_h = 0.0;                   // thickness of the crust-class groups above (thermal mantle)
  int last_class = -1;
  double bot_grad = 1.0;                  // Vs gradient at the bottom of the deepest mantle group
  for (int gi = 0; gi < t.ngroups; ++gi) {
    const SurfdispStackGroup& g = t.groups[gi];
    const double hv = (g.h_param >= 0) ? (double)pm[g.h_param] : g.h_fixed;
    double H = hv;
    if (g.h_mode == 1 && any) H = hv - ztop_prev;
    int N = g.nfine;
    if (g.nfine_rule == SURFDISP_N_CRUST) N = (H >= 150.0) ? 60 : ((H > 60.0) ? 30 : ((H > 20.0) ? 15 : ((H > 10.0) ? 10 : 5)));
    else if (g.nfine_rule == SURFDISP_N_OCRUST) N = min(max((int)rint(H / 2.0), 2), 10);
    double coef[SURFDISP_MAX_COEF];
    for (int i = 0; i < g.ncoef; ++i) coef[i] = (g.v_param[i] >= 0) ? (double)pm[g.v_param[i]] : g.v_fixed[i];
    // z = linspace(0, H, N+1); a group thinner than 0.01 km is skipped altogether (models.py:82)
    if (H - 0.0 < 0.01) continue;
    const bool is_ref = (g.kind == SURFDISP_G_REFMANTLE);
    const bool is_hyb = HYB && (g.kind == SURFDISP_G_HYBRID);
    if (is_hyb) {
      if (N > 63) N = 63;
      hybrid_profile<EMIT>(g, coef, (g.age_param >= 0) ? (double)pm[g.age_param] : g.age_fixed, H, N, crust_h, z0, ws);
    }
    const Knots kn = make_knots((g.kind == SURFDISP_G_BSPLINE && g.ncoef >= 3) ? g.ncoef : 3);
    const bool mono_class = (g.gclass == SURFDISP_C_SEDIMENT || g.gclass == SURFDISP_C_CRUST);
    const double zstep = H / (double)N, ustep = 1.0 / (double)N;
    const double vs0_ref = last_vs;
    // Vp, rho, Qs of the first grid point of a reference-mantle group (layers.py:279-283)
    double vp_first = 0.0, rho_first = 0.0;
    if (is_ref) { vp_first = g.vp_a * vs0_ref + g.vp_b; rho_first = stack_rho(g.rho_rule, g.rho_const, vs0_ref, vp_first); }
    double c_z = 0.0, c_vs = 0.0, c_vp = 0.0, c_rho = 0.0, c_qs = 0.0;   // carry: grid point base-1 (lane 31 of the previous chunk)
    for (int base = 0; base <= N; base += 32) {
      const int j = base + lane;
      const bool act = j <= N;
      const double zz = (j == N) ? H : (double)j * zstep;
      const double u = (j == N) ? 1.0 : (double)j * ustep;
      double vs = 0.0;
      if (act) {
        if (g.kind == SURFDISP_G_WATER) vs = 0.0;
        else if (g.kind == SURFDISP_G_CONST) vs = coef[0];
        else if (g.kind == SURFDISP_G_LINEAR) vs = (j == N) ? coef[1] : coef[0] + (double)j * ((coef[1] - coef[0]) / (double)N);
        else if (g.kind == SURFDISP_G_BSPLINE) {
          if (g.ncoef == 1) vs = coef[0];
          else if (g.ncoef == 2) vs = coef[0] * ((j == N) ? 0.0 : 1.0 + (double)j * ((0.0 - 1.0) / (double)N)) + coef[1] * u;
          else vs = bspl_profile(coef, kn, u);
        } else if (g.kind == SURFDISP_G_CASCADIA) vs = (0.02 * H * H + 1.27 * H + 0.29 * 0.1) / (H + 0.29);
        else if (is_hyb) vs = ws.hv[j];
        else {  // reference mantle: linear continuation of the deepest Vs (layers.py:267-285)
          const double vend = vs0_ref + H * g.slope;
          vs = (j == N) ? vend : vs0_ref + (double)j * ((vend - vs0_ref) / (double)N);
        }
      }
      double vp = g.vp_a * vs + g.vp_b;
      double rho = stack_rho(g.rho_rule, g.rho_const, vs, vp);
      double qs = (is_hyb && act) ? ws.hq[j] : g.qs;
      if (is_ref) { vp = last_vp + (vp - vp_first); rho = last_rho + (rho - rho_first); qs = last_qs + (qs - g.qs); }
      // the grid point before this lane's
      double p_z = __shfl_up_sync(full, zz, 1), p_vs = __shfl_up_sync(full, vs, 1), p_vp = __shfl_up_sync(full, vp, 1);
      double p_rho = __shfl_up_sync(full, rho, 1), p_qs = __shfl_up_sync(full, qs, 1);
      if (lane == 0) { p_z = c_z; p_vs = c_vs; p_vp = c_vp; p_rho = c_rho; p_qs = c_qs; }
      if (act && !is_ref) {
        // ---- prior rules on the grid (models.py:301-356, 576-600)
        if (vs > 4.9) bad |= SURFDISP_P_VSMAX;
        if (j == 0 && any && g.gclass != last_class && vs < last_vs) bad |= SURFDISP_P_JUMP;
        if (j > 0 && mono_class && !(vs - p_vs >= kEps)) bad |= SURFDISP_P_MONO;
        if (j == 0 && any && g.gclass == last_class && mono_class && !(vs - last_vs >= kEps))
          bad |= SURFDISP_P_MONO;   // two groups of the same class form one array in the reference's test
        if (g.gclass == SURFDISP_C_SEDIMENT && vs < 0.2) bad |= SURFDISP_P_SEDMIN;
        if (g.gclass == SURFDISP_C_MANTLE && nm + j < kMaxMantleGrid) { ws.vm[nm + j] = vs; ws.zm[nm + j] = zz + z0; }
      }
      if (EMIT) {
        const double h = (zz + z0) - (p_z + z0);
        const bool keep = act && j > 0 && h > 0.01;   // models.py:102 (and models.py:20: h > 1e-3)
        const unsigned km = __ballot_sync(full, keep);
        const int idx = nl + __popc(km & ((1u << lane) - 1u));
        if (keep) {
          if (idx < lmax) {
            o_vp[idx] = (float)(0.5 * (vp + p_vp)); o_vs[idx] = (float)(0.5 * (vs + p_vs));
            o_rho[idx] = (float)(0.5 * (rho + p_rho)); o_h[idx] = (float)h;
            o_qs[idx] = (float)(1.0 / (0.5 * (qs + p_qs)));
          } else overflow = true;
        }
        nl += __popc(km);
      }
      c_z = __shfl_sync(full, zz, 31); c_vs = __shfl_sync(full, vs, 31); c_vp = __shfl_sync(full, vp, 31);
      c_rho = __shfl_sync(full, rho, 31); c_qs = __shfl_sync(full, qs, 31);
      // values of the group's last grid point (and the one before it), first two of the model
      if (base + 31 >= N) {
        const int ln = N - base, lp = ln - 1;           // lane of grid point N; N - 1 sits in lane lp (or in the carry)
        const double e_vs = __shfl_sync(full, vs, ln), e_vp = __shfl_sync(full, vp, ln), e_rho = __shfl_sync(full, rho, ln);
        const double e_qs = __shfl_sync(full, qs, ln), e_z = __shfl_sync(full, zz, ln);
        const double e_pvs = __shfl_sync(full, p_vs, ln), e_pz = __shfl_sync(full, p_z, ln);
        (void)lp;
        if (!is_ref) {
          if (g.gclass == SURFDISP_C_MANTLE) bot_grad = (e_vs - e_pvs) / (e_z - e_pz);
          end_vs1 = e_vs; end_vs2 = e_pvs; end_z1 = e_z + z0; end_z2 = e_pz + z0;
        }
        last_vs = e_vs; last_vp = e_vp; last_rho = e_rho; last_qs = e_qs;
      }
      if (base == 0 && !is_ref && ngrid == 0) { first_vs0 = __shfl_sync(full, vs, 0); first_vs1 = __shfl_sync(full, vs, 1); }
    }
    if (g.gclass == SURFDISP_C_CRUST && H / (double)N > 0.01) crust_h += H;    // OceanMantleHybrid.getCrustH (layers.py:304-311)
    if (!is_ref) {
      last_class = g.gclass;
      ngrid += N + 1;
      if (g.gclass == SURFDISP_C_MANTLE) nm = min(nm + N + 1, kMaxMantleGrid);
    }
    z0 = z0 + H;
    ztop_prev = z0;
    any = true;
  }
  bad = __reduce_or_sync(full, (unsigned)bad);
  if (!(bot_grad > 0.0)) bad |= SURFDISP_P_BOTTOM;
  if (rules & (SURFDISP_P_FIRSTPAIR | SURFDISP_P_OSCI | SURFDISP_P_LOCALMAX | SURFDISP_P_SLOPE | SURFDISP_P_CWT)) {
    // CascadiaOcean.isgood (models.py:571-677).  `grp` is a Python list there: the jump rule compares only the first
    // two grid points (models.py:586-588), the bottom rule looks at the last two points of the whole grid
    // (models.py:597-598, folded into SURFDISP_P_BOTTOM for these models).
    if (ngrid >= 2 && first_vs1 < first_vs0) bad |= SURFDISP_P_FIRSTPAIR;
    if (ngrid >= 2 && !((end_vs1 - end_vs2) / (end_z1 - end_z2) > 0.0)) bad |= SURFDISP_P_BOTTOM;
    __syncwarp();
    bad |= ocean_mantle_rules(ws, nm);
  }
  if (EMIT) {
    overflow = __any_sync(full, overflow);
    *nl_out = overflow ? -1 : nl;
  }
  return bad;
}

// ------------------------------------------------------------------------------------ kernels: builder, priors
constexpr int kMcThreads = 128;     // 4 warps = 4 models / chains per block

template <bool HYB>
__global__ void __launch_bounds__(kMcThreads, 4) build_stacks_kernel(const __grid_constant__ SurfdispStackTemplate t, int M,
                                                                  const float* __restrict__ params, int lmax,
                                                                  float* __restrict__ layers, int* __restrict__ nlay) {
  __shared__ WarpScratch scratch[kMcThreads / 32];
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  const float* pm = params + (size_t)m * t.nparams;
  const size_t pl = (size_t)M * lmax;
  float* o_vp = layers + 0 * pl + (size_t)m * lmax;
  float* o_vs = layers + 1 * pl + (size_t)m * lmax;
  float* o_rho = layers + 2 * pl + (size_t)m * lmax;
  float* o_h = layers + 3 * pl + (size_t)m * lmax;
  float* o_qs = layers + 4 * pl + (size_t)m * lmax;
  int nl = 0;
  assemble_stack_warp<true, HYB>(t, pm, 0, lmax, o_vp, o_vs, o_rho, o_h, o_qs, &nl, scratch[threadIdx.x >> 5]);
  const int nz = nl < 0 ? 0 : nl;
  for (int j = nz + lane; j < lmax; j += 32) { o_vp[j] = 0.f; o_vs[j] = 0.f; o_rho[j] = 0.f; o_h[j] = 0.f; o_qs[j] = 0.f; }
  if (lane == 0) nlay[m] = nz;
}

template <bool HYB>
__global__ void __launch_bounds__(kMcThreads, 4) check_priors_kernel(const __grid_constant__ SurfdispStackTemplate t, int M,
                                                                  const float* __restrict__ params, int* __restrict__ priors) {
  __shared__ WarpScratch scratch[kMcThreads / 32];
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  const int bad = assemble_stack_warp<false, HYB>(t, params + (size_t)m * t.nparams, 0x7fffffff, 0, nullptr, nullptr, nullptr, nullptr,
                                             nullptr, nullptr, scratch[threadIdx.x >> 5]);
  if ((threadIdx.x & 31) == 0) priors[m] = bad;
}

// ------------------------------------------------------------------------------------ Monte-Carlo step
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so chain m at step s draws the same numbers whatever
// the launch geometry.  counter = (chain, step, block, stream), key = seed; stream = (try << 8) | (parameter << 2) | 1
// for the proposal of one parameter in one try, 2 for the Metropolis draw.
struct Philox {
  unsigned int c0, c1, c2, c3, k0, k1;
  unsigned int buf[4];
  int have;
  __device__ Philox(unsigned long long seed, unsigned int chain, unsigned int step, unsigned int stream_id)
      : c0(chain), c1(step), c2(0u), c3(stream_id), k0((unsigned int)seed), k1((unsigned int)(seed >> 32)), have(0) {}
  __device__ void block() {
    unsigned int x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const unsigned int h0 = __umulhi(0xD2511F53u, x0), l0 = 0xD2511F53u * x0;
      const unsigned int h1 = __umulhi(0xCD9E8D57u, x2), l1 = 0xCD9E8D57u * x2;
      const unsigned int y0 = h1 ^ x1 ^ a, y1 = l1, y2 = h0 ^ x3 ^ b, y3 = l0;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    buf[0] = x0; buf[1] = x1; buf[2] = x2; buf[3] = x3;
    c2++;
    have = 4;
  }
  __device__ unsigned int next() { if (!have) block(); return buf[--have]; }
  __device__ float uniform() { return ((float)(next() >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)
  __device__ float gauss() {   // Box-Muller
    const float u1 = uniform(), u2 = uniform();
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
};

struct McBounds { float lo[64], hi[64], step[64]; };
constexpr int kMaxParams = 64;

// One admissible proposal for the chain of this warp, into ws.q (lanes over parameters; the prior rules by the whole
// warp).  MCinv.perturb (models.py:190-205): up to 1000 proposals, then MCinv.reset (models.py:206-219): up to 10000.
// Returns the number of tries, -1 if no admissible model was found (the reference raises there).
template <bool HYB>
__device__ int propose_warp(const SurfdispStackTemplate& t, const McBounds& bd, const float* cur, bool restart,
                            unsigned long long seed, unsigned int chain, unsigned int step_index, WarpScratch& ws) {
  const int lane = threadIdx.x & 31;
  const int P = t.nparams;
  int tries = 0;
  for (int phase = restart ? 1 : 0; phase < 2; ++phase) {
    const int limit = phase == 0 ? 1000 : 10000;
    for (int a = 0; a < limit; ++a) {
      ++tries;
      for (int i = lane; i < P; i += 32) {
        Philox rng(seed, chain, step_index, ((unsigned)tries << 8) | ((unsigned)i << 2) | 1u);
        const float lo = bd.lo[i], hi = bd.hi[i];
        float v = 0.f;
        bool ok = false;
        if (phase == 0) {
          // BrownianVar.move (brownian.py:20-27): Gaussian step, redrawn until strictly inside the bounds
          const float c = cur[i];
          for (int r = 0; r < 1000 && !ok; ++r) {
            v = c + bd.step[i] * rng.gauss();
            ok = (v < hi && v > lo);
          }
        }
        if (!ok) v = lo + (hi - lo) * rng.uniform();   // BrownianVar.reset (brownian.py:17-19)
        ws.q[i] = v;
      }
      __syncwarp();
      const int bad = t.prior_mask ? (assemble_stack_warp<false, HYB>(t, ws.q, t.prior_mask, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                  nullptr, ws) & t.prior_mask) : 0;
      __syncwarp();
      if (!bad) return tries;
    }
  }
  return -1;
}

template <bool HYB>
__global__ void __launch_bounds__(kMcThreads, 4) mc_propose_kernel(const __grid_constant__ SurfdispStackTemplate t,
                                                                const __grid_constant__ McBounds bd, int M,
                                                                const float* __restrict__ cur,
                                                                const unsigned char* __restrict__ reset_mask,
                                                                float* __restrict__ prop, int* __restrict__ status,
                                                                unsigned long long seed, unsigned int step_index) {
  __shared__ WarpScratch scratch[kMcThreads / 32];
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  WarpScratch& ws = scratch[threadIdx.x >> 5];
  const int P = t.nparams;
  const bool restart = reset_mask && reset_mask[m];
  const int result = propose_warp<HYB>(t, bd, cur + (size_t)m * P, restart, seed, (unsigned)m, step_index, ws);
  for (int i = lane; i < P; i += 32) prop[(size_t)m * P + i] = ws.q[i];
  if (status && lane == 0) status[m] = result;
}

// Fused first half of a Monte-Carlo step (point.py:45-59): proposal + model assembly of every chain.
//   step = *step_ptr (device counter, bumped by mc_bump_kernel at the end of the step: the same launches replay as a
//   CUDA graph).  step % chain_len == 0 starts a sub-chain: chains flagged in init_mask take the start model as it
//   is -- perturbed if it violates the priors -- (point.py:47-50), the others a uniform redraw (point.py:52).
struct McStepParams {
  int M, P, lmax, chain_len;
  const float* cur; float* prop; int* status;
  const unsigned char* init_mask;
  float* layers; int* nlay;
  const unsigned int* step_ptr;
  unsigned long long seed;
  const McBounds* bounds;       // device: one McBounds per point; chain m belongs to point m / chains_per_point
  int chains_per_point;
};

template <bool HYB>
__global__ void __launch_bounds__(kMcThreads, 4) mc_propose_build_kernel(const __grid_constant__ SurfdispStackTemplate t,
                                                                      const __grid_constant__ McStepParams p) {
  __shared__ WarpScratch scratch[kMcThreads / 32];
  __shared__ McBounds sbd[kMcThreads / 32];
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (m >= p.M) return;
  WarpScratch& ws = scratch[w];
  const int P = p.P;
  {
    const McBounds& src = p.bounds[m / p.chains_per_point];
    for (int i = lane; i < P; i += 32) { sbd[w].lo[i] = src.lo[i]; sbd[w].hi[i] = src.hi[i]; sbd[w].step[i] = src.step[i]; }
    __syncwarp();
  }
  const unsigned int step = *p.step_ptr;
  const float* cur = p.cur + (size_t)m * P;
  const bool start = (p.chain_len > 0) && (step % (unsigned)p.chain_len == 0u);
  const bool init = start && step == 0u && p.init_mask && p.init_mask[m];   // (point.py:47: `init` is consumed by the first sub-chain)
  int result = 0;
  bool take_cur = false;
  if (init) {
    for (int i = lane; i < P; i += 32) ws.q[i] = cur[i];
    __syncwarp();
    const int bad = t.prior_mask ? (assemble_stack_warp<false, HYB>(t, ws.q, t.prior_mask, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                nullptr, ws) & t.prior_mask) : 0;
    __syncwarp();
    take_cur = !bad;
  }
  if (!take_cur) {
    result = propose_warp<HYB>(t, sbd[w], cur, start && !init, p.seed, (unsigned)m, step, ws);
    if (result < 0) {   // no admissible model: the chain keeps its state, the row is flagged (status < 0)
      for (int i = lane; i < P; i += 32) ws.q[i] = cur[i];
      __syncwarp();
    }
  }
  for (int i = lane; i < P; i += 32) p.prop[(size_t)m * P + i] = ws.q[i];
  if (lane == 0) p.status[m] = result;
  // model assembly of the proposal
  const size_t pl = (size_t)p.M * p.lmax;
  float* o_vp = p.layers + 0 * pl + (size_t)m * p.lmax;
  float* o_vs = p.layers + 1 * pl + (size_t)m * p.lmax;
  float* o_rho = p.layers + 2 * pl + (size_t)m * p.lmax;
  float* o_h = p.layers + 3 * pl + (size_t)m * p.lmax;
  float* o_qs = p.layers + 4 * pl + (size_t)m * p.lmax;
  int nl = 0;
  assemble_stack_warp<true, HYB>(t, ws.q, 0, p.lmax, o_vp, o_vs, o_rho, o_h, o_qs, &nl, ws);
  const int nz = nl < 0 ? 0 : nl;
  for (int j = nz + lane; j < p.lmax; j += 32) { o_vp[j] = 0.f; o_vs[j] = 0.f; o_rho[j] = 0.f; o_h[j] = 0.f; o_qs[j] = 0.f; }
  if (lane == 0) p.nlay[m] = nz;
}

// Second half of the step: misfit of the proposal (Point.misfit point.py:15-31 / PointCascadia.misfit :337-366),
// Metropolis rule (point.py:34-37), state update and the chain-track row [misfit, L, accepted, parameters of the
// proposal] (Model1D._dump, models.py:243-245).  One thread per chain.
struct McFinishParams {
  int M, P, K, mode, chain_len, chains_per_point, track_steps;
  const float* c_pred; const int* nfound; const int* status;
  const float* obs; const float* isig; const unsigned char* use;   // device [n_points][K]
  float per[SURFDISP_MAX_PERIODS];
  const float* prop; float* cur; float* chi0;
  unsigned char* accepted; float* misfit_out;                       // [M][3] (misfit, chiSqr, L)
  float* c_cur;                                                     // [M][K] curve of the chain's current model (hint of the next step) or nullptr
  float* track;                                                     // [track_steps][M][3 + P] or nullptr
  const unsigned int* step_ptr;
  unsigned long long seed;
};

__device__ __forceinline__ void misfit_one(int mode, int K, const float* row, const float* obs, const float* isig,
                                           const unsigned char* use, const float* per, bool failed, double& misfit,
                                           double& chi, double& L) {
  if (failed) { misfit = 88888.0; chi = 88888.0; L = 0.0; return; }   // point.py:20-21
  double s1 = 0.0, s2 = 0.0;
  int n1 = 0, n2 = 0;
  for (int k = 0; k < K; ++k) {
    if (!use[k]) continue;
    const double bias = ((double)obs[k] - (double)row[k]) * (double)isig[k];
    if (mode == 1 && per[k] > 40.f) { s2 += bias * bias; n2++; }
    else { s1 += bias * bias; n1++; }
  }
  const int N = n1 + n2;
  if (mode == 1) {
    if (n1 > 0 && n2 > 0) chi = (s1 / n1 + s2 / n2) / 2.0 * N;
    else if (n2 > 0) chi = s2 / n2 * N;
    else chi = s1 / (n1 > 0 ? n1 : 1) * N;
  } else chi = s1;
  misfit = sqrt(chi / (N > 0 ? N : 1));
  if (!(chi < 50.0)) chi = sqrt(chi * 50.0);  // point.py:29
  L = exp(-0.5 * chi);
}

__global__ void __launch_bounds__(128) mc_finish_kernel(const __grid_constant__ McFinishParams p) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  const unsigned int step = *p.step_ptr;
  const int pt = m / p.chains_per_point;
  double misfit, chi, L;
  misfit_one(p.mode, p.K, p.c_pred + (size_t)m * p.K, p.obs + (size_t)pt * p.K, p.isig + (size_t)pt * p.K, p.use + (size_t)pt * p.K,
             p.per, p.nfound[m] < p.K, misfit, chi, L);
  const bool start = (p.chain_len > 0) && (step % (unsigned)p.chain_len == 0u);   // first sample of a sub-chain: taken (point.py:57)
  const bool stuck = p.status && p.status[m] < 0;
  const float x0 = p.chi0[m], x1 = (float)chi;
  bool acc = start;
  if (!acc && !stuck) {
    if (x1 < x0) acc = true;
    else {
      Philox rng(p.seed, (unsigned int)m, step, 2u);
      acc = (double)rng.uniform() > 1.0 - exp(-0.5 * ((double)x1 - (double)x0));
    }
  }
  if (stuck && !start) acc = false;
  if (acc) {
    p.chi0[m] = x1;
    for (int i = 0; i < p.P; ++i) p.cur[(size_t)m * p.P + i] = p.prop[(size_t)m * p.P + i];
    if (p.c_cur) {   // the accepted model's curve guides the root search of the chain's next proposals (zeros: no hint)
      const bool full = p.nfound[m] >= p.K;
      for (int k = 0; k < p.K; ++k) p.c_cur[(size_t)m * p.K + k] = full ? p.c_pred[(size_t)m * p.K + k] : 0.f;
    }
  }
  p.accepted[m] = acc ? 1 : 0;
  if (p.misfit_out) { float* o = p.misfit_out + (size_t)m * 3; o[0] = (float)misfit; o[1] = (float)chi; o[2] = (float)L; }
  if (p.track) {
    float* row = p.track + ((size_t)(step % (unsigned)p.track_steps) * p.M + m) * (3 + p.P);
    row[0] = (float)misfit; row[1] = (float)L; row[2] = acc ? 1.f : 0.f;
    for (int i = 0; i < p.P; ++i) row[3 + i] = p.prop[(size_t)m * p.P + i];
  }
}

__global__ void mc_bump_kernel(unsigned int* step_ptr) { *step_ptr += 1u; }

__global__ void __launch_bounds__(256) mc_accept_kernel(int M, int P, const float* __restrict__ chi1,
                                                        const float* __restrict__ prop, float* __restrict__ chi0,
                                                        float* __restrict__ cur, const unsigned char* __restrict__ force,
                                                        unsigned char* __restrict__ accepted, unsigned long long seed,
                                                        unsigned int step_index) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  Philox rng(seed, (unsigned int)m, step_index, 2u);
  const float x0 = chi0[m], x1 = chi1[m];
  bool acc = force && force[m];
  if (!acc) {
    // point.py:34-37: accept if chi1 < chi0, else if u > 1 - exp(-(chi1 - chi0) / 2)   (= (L0 - L1) / L0)
    if (x1 < x0) acc = true;
    else acc = (double)rng.uniform() > 1.0 - exp(-0.5 * ((double)x1 - (double)x0));
  }
  if (acc) {
    chi0[m] = x1;
    for (int i = 0; i < P; ++i) cur[(size_t)m * P + i] = prop[(size_t)m * P + i];
  }
  accepted[m] = acc ? 1 : 0;
}

}  // namespace mcdev
