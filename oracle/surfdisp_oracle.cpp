// oracle/surfdisp_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference dispersion forward path
// (001cat/pySurfInv fast_surf_src/*.f, and its real*8 sibling senskernel-1.0/src/SURF_PERTURB/*.f).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library; the product path (pysurfinv_b200/) never does.
//
// It is a from-scratch C++ restatement, templated on two scalar types:
//   RM  -- type of the "model preparation" arithmetic (attenuation correction + earth flattening)
//   R   -- type of everything the reference declares as default REAL (secular functions, root
//          search, the float parts of the eigen-integrals).  Variables the reference declares
//          DOUBLE PRECISION (REIGEN's ODE state, surfa.f:717-722) are always double.
// precision 0: RM=float,  R=float   -> float32-faithful fast_surf semantics (same op order, libm)
// precision 1: RM=float,  R=double  -> "what the float32 model's exact root is" (noise-free solver)
// precision 2: RM=double, R=double  -> full double; with sibling=1 it follows SURF_PERTURB
//                                      (the program that produced senskernel-1.0/TEST1/*)
//
// Parity status: pinned against senskernel-1.0/TEST1/test.{R,L}.{phv,grv} (real*8 sibling, golden
// vectors committed under tests/golden/).  The float32 last digits of fast_surf itself are
// UNPINNED: no Fortran compiler exists in the build container, so gfortran output could not be
// generated.
//
// Every routine cites the reference lines it follows (paths relative to /root/reference).
// Build: make -C oracle   (g++ -O2 -ffp-contract=off; no fast-math, so float32 ops round as in
// gfortran -O3 on x86-64 without FMA).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

extern "C" {

struct OracleOpts {
  int precision;    // 0,1,2 see above
  int sibling;      // 1: SURF_PERTURB variant (double constants, strict '>' in scan stop)
  int nmode;        // fast_surf: 1 (init.f:58); TEST1: 2
  int ndiv;         // initial sub-layer count, 5 (init.f:25)
  int ndiv_cap_r;   // REIGEN: ndiv<=cap/(n-1); fast_surf 99 (surfa.f:783), sibling 999
  int ndiv_cap_l;   // LEIGEN: 999 (surfa.f:414)
  int neville_cap;  // 50 (surfa.f:17), sibling 5000
  int stale_mmax;   // 1 = reference behaviour; 0 = refresh all layers every period
  int atten;        // KEY_ATTEN (init.f:43)
  int flat;         // flat1 applied (calcul.f:133)
  double dc;        // 0.01 (init.f:25)
  double fact;      // 4.0
  double t_base;    // 1.0 (fast_surf.f:77)
};

struct OracleCounters {
  long long sweeps_R, steps_R, sweeps_L, steps_L, sub_U, flat_layers, scan_evals, polish_evals;
};

}  // extern "C"

namespace {

// ---------------------------------------------------------------- math wrappers (libm, per type)
inline float  m_log(float x)  { return logf(x); }
inline double m_log(double x) { return log(x); }
inline float  m_exp(float x)  { return expf(x); }
inline double m_exp(double x) { return exp(x); }
inline float  m_sin(float x)  { return sinf(x); }
inline double m_sin(double x) { return sin(x); }
inline float  m_cos(float x)  { return cosf(x); }
inline double m_cos(double x) { return cos(x); }
inline float  m_sqrt(float x)  { return sqrtf(x); }
inline double m_sqrt(double x) { return sqrt(x); }
inline float  m_pow(float x, float y)   { return powf(x, y); }
inline double m_pow(double x, double y) { return pow(x, y); }
inline float  m_abs(float x)  { return fabsf(x); }
inline double m_abs(double x) { return fabs(x); }
// Fortran SIGN(1.,x): follows the sign bit (calcul.f:160, surfa.f:37,46)
template <typename T> inline int sgn(T x) { return std::signbit(x) ? -1 : 1; }

enum { NSIZE = 1000 };

template <typename RM, typename R>
struct Solver {
  OracleOpts o;
  OracleCounters cnt{};
  int kind = 2;
  // COMMON /c/
  int nmax = 0, mmax = 0, kmax = 0, idrop = 0, ndiv = 5, mode = 1;
  R fact = 4, dc = R(0.01);
  int lstop = 0;
  // COMMON /ref/ (1-based)
  std::vector<RM> a_ref, b_ref, rho_ref, d_ref, qs_ref;
  // COMMON /d/ working model (1-based)
  std::vector<R> a, b, rho, d;
  std::vector<RM> t;  // periods, 1-based
  // constants
  RM pi_att;
  R twopi_love, twopi_ray, twopi_leig, twopi_reig;
  RM t_base;
  // results (1-based k, 1-based iq)
  std::vector<std::vector<R>> c, ratio, ugr, cvar;
  // partial derivatives of the phase velocity of the last REIGEN call, per ORIGINAL layer of the period's model (the
  // reference keeps them per sub-layer in COMMON /derivd/, surfa.f:1130-1135, 1182-1185, 1202-1208; a layer's
  // derivative is the sum over its sub-layers); filled only when want_partials is set
  bool want_partials = false;
  std::vector<double> p_dcda, p_dcdb, p_dcdr;
  std::vector<std::vector<double>> all_dcda, all_dcdb, all_dcdr;   // [period][layer], fundamental mode
  std::vector<int> imax;

  void setup_consts() {
    if (o.sibling) {
      // SURF_PERTURB.f:117 (pi=datan(1)*4); surfa.f(sibling):190,263,591,892
      pi_att = RM(3.14159265358979323846);
      twopi_love = R(6.2831853);
      twopi_ray = R(6.28318531f);
      twopi_leig = R(6.2831853);
      twopi_reig = R(6.28318530717958647692);
    } else {
      // float32 literals: calcul.f:32, surfa.f:143,193,488,871
      pi_att = RM(3.1415927f);
      twopi_love = R(6.2831853f);
      twopi_ray = R(6.28318531f);
      twopi_leig = R(6.2831853f);
      twopi_reig = R(6.2831853072f);
    }
    t_base = RM(o.t_base);
    fact = R((float)o.fact);
    dc = R((float)o.dc);
  }

  // ------------------------------------------------------------------ flat1.f:2-73
  // in-place on RM arrays h,ro,vp,vs [1..n]
  void flat1(RM* h, RM* ro, RM* vp, RM* vs, int n) {
    std::vector<RM> hh(n + 2);
    const RM A = RM(6371.0f);
    for (int i = 1; i <= n; ++i) hh[i] = h[i];
    RM pwr = RM(2.2750f);
    if (kind == 1) pwr = RM(5.0f);
    int nm = n - 1;
    RM hs = 0;
    for (int i = 1; i <= n; ++i) {  // flat1.f:33-37
      RM ht = hs;
      hs = hs + hh[i];
      hh[i] = A - ht;
    }
    for (int i = 1; i <= nm; ++i) {  // flat1.f:41-56
      int ii = i + 1;
      RM fltd = m_log(hh[i] / hh[ii]);
      RM dif = (RM(1) / hh[ii] - RM(1) / hh[i]) * A / fltd;
      RM difr = m_pow(hh[i], pwr) - m_pow(hh[ii], pwr);
      RM qqq;
      if (o.sibling) {
        ro[i] = ro[i] * difr / (fltd * m_pow(A, pwr) * pwr);  // sibling flat1.f
      } else {
        qqq = difr / (fltd * m_pow(A, pwr) * pwr);
        ro[i] = ro[i] * qqq;
      }
      vp[i] = vp[i] * dif;
      vs[i] = vs[i] * dif;
    }
    RM fct = A / hh[n];  // flat1.f:58-62
    vp[n] = vp[n] * fct;
    vs[n] = vs[n] * fct;
    ro[n] = ro[n] * m_pow(RM(1) / fct, pwr);
    RM z0 = 0;
    for (int i = 2; i <= n; ++i) {  // flat1.f:65-68
      RM z1 = A * m_log(A / hh[i]);
      h[i - 1] = z1 - z0;
      z0 = z1;
    }
    h[n] = 0;
    cnt.flat_layers += n;
  }

  // ------------------------------------------ calcul.f:112-133 (also :239-251, :325-337)
  // rebuild working layers 1..m for period T: attenuation correction then flattening with layer m
  // treated as the half-space.
  void refresh(int m, RM T) {
    std::vector<RM> ta(m + 2), tb(m + 2), tr(m + 2), td(m + 2);
    for (int i = 1; i <= m; ++i) {
      tb[i] = b_ref[i];
      ta[i] = a_ref[i];
      tr[i] = rho_ref[i];
      td[i] = d_ref[i];
      if (o.atten) {
        if (o.sibling) {
          // calcul_deep.f:125-128
          RM lg = m_log(t_base / T);
          tb[i] = b_ref[i] * (RM(1) + qs_ref[i] * lg / pi_att);
          RM qp = qs_ref[i] * RM(4) / RM(3) * (b_ref[i] * b_ref[i]) / (a_ref[i] * a_ref[i]);
          ta[i] = a_ref[i] * (RM(1) + qp * lg / pi_att);
        } else {
          // calcul.f:122-126
          RM qsq = qs_ref[i] * m_log(t_base / T) / pi_att;
          RM qpq = qsq * RM(1.33333333f) * (b_ref[i] * b_ref[i]) / (a_ref[i] * a_ref[i]);
          tb[i] = b_ref[i] * (RM(1) + qsq);
          ta[i] = a_ref[i] * (RM(1) + qpq);
        }
      }
    }
    if (o.flat) flat1(td.data(), tr.data(), ta.data(), tb.data(), m);
    for (int i = 1; i <= m; ++i) {
      a[i] = R(ta[i]);
      b[i] = R(tb[i]);
      rho[i] = R(tr[i]);
      d[i] = R(td[i]);
    }
  }

  // ------------------------------------------------------------------ DLTAR1 surfa.f:135-183
  R dltar1(R cc, R tt_) {
    R wvno = twopi_love / (cc * tt_);
    R covb = cc / b[mmax];
    R h = rho[mmax] * b[mmax] * b[mmax];
    R rb = m_sqrt(m_abs(covb * covb - R(1)));
    R ut = 1;
    R tt = h * rb;
    R ett = tt;  // (reference leaves ett unset when mmax==1; never happens, mmax>=2)
    int mmm1 = mmax - 1;
    cnt.sweeps_L++;
    for (int k = 1; k <= mmm1; ++k) {
      int m = mmax - k;
      if (b[m] == R(0)) continue;
      cnt.steps_L++;
      covb = cc / b[m];
      rb = m_sqrt(m_abs(covb * covb - R(1)));
      h = rho[m] * b[m] * b[m];
      R q = -wvno * d[m] * rb;
      R y, z, cosq;
      if (rb < R(0.1e-20f) || cc == b[m]) {  // 1221
        y = -wvno * d[m];
        z = 0;
        cosq = 1;
      } else if (cc > b[m]) {  // 1231
        R sinq = m_sin(q);
        y = sinq / rb;
        z = rb * sinq;
        cosq = m_cos(q);
      } else {  // 1209
        R exqp = m_exp(q);
        R exqm = R(1) / exqp;
        y = (exqp - exqm) / (R(2) * rb);
        z = -rb * rb * y;
        cosq = (exqp + exqm) / R(2);
      }
      R eut = cosq * ut - y * tt / h;
      ett = h * z * ut + cosq * tt;
      ut = eut;
      tt = ett;
    }
    return -ett;
  }

  // ------------------------------------------------------------------ DLTAR4 surfa.f:185-372
  // mup=1 dispersion function, mup=2 ellipticity (two sweeps)
  R dltar4(R c_, R t_, int mup) {
    const R accur = R(1.e-8f), accurs = R(1.e-8f);
    R wvno = twopi_ray / (c_ * t_);
    R csq = c_ * c_;
    int jump = 1;
    if (mup > 1) jump = 2;
    R r12 = 0;
    for (;;) {
      R b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
      if (jump == 1) b1 = 1;
      else if (jump == 2) b2 = 1;
      else b3 = 1;
      R ra = 0, rb = 0, g = 0, g1 = 0;
      int m;
      bool reached_half = false;
      cnt.sweeps_R++;
      for (m = 1; m <= mmax; ++m) {
        R arga = R(1) - csq / (a[m] * a[m]);
        ra = m_sqrt(m_abs(arga));
        if (arga > R(0)) ra = -ra;
        R a11, a12, a13, a14, a15, a21, a22, a23, a24, a31, a32, a33, a41, a42, a51;
        if (!(m_abs(b[m]) > accurs)) {
          // liquid surface layer, surfa.f:219-251
          R pm = wvno * ra * d[m];
          if (mup > 1) continue;
          cnt.steps_R++;
          R rhoc = rho[m] * csq;
          R sinpr, cosp;
          if (m_abs(ra) < accur || ra == R(0)) {
            sinpr = wvno * d[m];
            cosp = 1;
          } else if (ra < R(0)) {
            sinpr = (m_exp(pm) - m_exp(-pm)) / (R(2) * ra);
            cosp = R(0.5f) * (m_exp(pm) + m_exp(-pm));
          } else {
            sinpr = m_sin(pm) / ra;
            cosp = m_cos(pm);
          }
          a11 = cosp;
          a21 = rhoc * sinpr;
          a31 = a41 = a51 = a12 = a22 = a32 = a42 = a13 = a23 = a33 = a14 = a24 = a15 = 0;
        } else {
          // surfa.f:253-320
          R argb = R(1) - csq / (b[m] * b[m]);
          rb = m_sqrt(m_abs(argb));
          if (argb > R(0)) rb = -rb;
          g = R(2) * (b[m] * b[m]) / csq;
          g1 = g - R(1);
          if (mmax == m) { reached_half = true; break; }
          cnt.steps_R++;
          R rhoc = rho[m] * csq;
          R pm = wvno * ra * d[m];
          R qm = wvno * rb * d[m];
          R rsinp, sinpr, cosp, rsinq, sinqr, cosq;
          if (ra < R(0)) {  // 213
            rsinp = -ra * R(0.5f) * (m_exp(pm) - m_exp(-pm));
            sinpr = -rsinp / (ra * ra);
            cosp = R(0.5f) * (m_exp(pm) + m_exp(-pm));
          } else if (ra == R(0)) {  // 212
            rsinp = 0;
            sinpr = wvno * d[m];
            cosp = 1;
          } else {  // 214
            rsinp = ra * m_sin(pm);
            sinpr = rsinp / (ra * ra);
            cosp = m_cos(pm);
          }
          if (m_abs(rb) < accur) {  // 218
            rsinq = 0;
            sinqr = wvno * d[m];
            cosq = 1;
          } else if (rb > R(0)) {  // 217
            rsinq = rb * m_sin(qm);
            sinqr = rsinq / (rb * rb);
            cosq = m_cos(qm);
          } else {
            rsinq = -rb * R(0.5f) * (m_exp(qm) - m_exp(-qm));
            sinqr = -rsinq / (rb * rb);
            cosq = R(0.5f) * (m_exp(qm) + m_exp(-qm));
          }
          R rr = rsinp * rsinq;
          R ss = sinpr * sinqr;
          R cc = cosp * cosq;
          R rs1 = rsinp * cosq;
          R rs2 = sinqr * cosp;
          R rs3 = sinpr * cosq;
          R rs4 = rsinq * cosp;
          R gm = R(2) * g - R(1);
          R gs = g * g;
          R g1s = g1 * g1;
          R ccm = R(1) - cc;
          R gg1 = g * g1;
          R rhocs = rhoc * rhoc;
          R suu = gs * rr + g1s * ss;
          a11 = R(2) * gs - gm;
          a11 = a11 * cc - suu - R(2) * gg1;
          a12 = -(rs1 + rs2) / rhoc;
          a13 = gm * ccm + g1 * ss + g * rr;
          a13 = -R(2) * a13 / rhoc;
          a14 = (rs3 + rs4) / rhoc;
          a15 = R(2) * ccm + rr + ss;
          a15 = a15 / rhocs;
          a21 = rhoc * (g1s * rs3 + gs * rs4);
          a22 = cc;
          a23 = R(2) * (g * rs4 + g1 * rs3);
          a24 = sinpr * rsinq;
          a31 = rhoc * (gg1 * gm * ccm + g1s * g1 * ss + gs * g * rr);
          a32 = g1 * rs2 + g * rs1;
          a33 = R(1) + R(2) * (R(2) * gg1 * ccm + suu);
          a41 = -rhoc * (g1s * rs2 + gs * rs1);
          a42 = rsinp * sinqr;
          a51 = rhocs * (R(2) * gs * g1s * ccm + gs * gs * rr + g1s * g1s * ss);
        }
        // surfa.f:326-335
        R bb1 = a11 * b1 + a12 * b2 + a13 * b3 + a14 * b4 + a15 * b5;
        R bb2 = a21 * b1 + a22 * b2 + a23 * b3 + a24 * b4 - a14 * b5;
        R bb3 = a31 * b1 + a32 * b2 + a33 * b3 - R(0.5f) * a23 * b4 + R(0.5f) * a13 * b5;
        R bb4 = a41 * b1 + a42 * b2 - R(2) * a32 * b3 + a22 * b4 - a12 * b5;
        R bb5 = a51 * b1 - a41 * b2 + R(2) * a31 * b3 - a21 * b4 + a11 * b5;
        b1 = bb1; b2 = bb2; b3 = bb3; b4 = bb4; b5 = bb5;
      }
      if (!reached_half) m = mmax;  // liquid half-space is outside the contract
      // half-space row, surfa.f:341-354
      R pp = a[m];
      R sss = b[m] * b[m];
      R ppp = pp * pp;
      R rhp = rho[m] * pp;
      R gra = g * ra;
      R g1s = g1 * g1;
      R rba = rb - R(1) / ra;
      R h11 = -R(2) * rb * sss / ppp + csq * g1s / ppp / gra;
      R h12 = rhp * pp;
      R h13 = -rb / h12 + g1 / h12 / gra;
      R h14 = rb / h12 / gra;
      R h15 = rba / rhp / rhp / csq / g;
      h12 = -R(1) / g / h12;
      R bb1 = h11 * b1 + h12 * b2 + R(2) * h13 * b3 + h14 * b4 + h15 * b5;
      if (mup == 1) return -bb1;
      if (jump == 2) r12 = bb1;
      jump = jump + 1;
      if (jump == 3) continue;
      return R(0.5f) * bb1 / r12;
    }
  }

  // ------------------------------------------------------------------ DLTAR surfa.f:85-133
  R dltar(R cc, R tt, int kk) {
    if (idrop <= 0) {
      R dmax = fact * cc * tt;
      mmax = nmax;
      R sum = 0;
      for (int ii = 1; ii <= nmax; ++ii) {
        if (cc - b[ii] < R(0)) {
          sum = sum + d[ii];
          if (sum - dmax > R(0)) { mmax = ii; break; }
        }
      }
      idrop = 1;
      if (mmax < 2) mmax = 2;
    }
    if (kk == 1) return dltar1(cc, tt);
    if (kk == 2) return dltar4(cc, tt, 1);
    return dltar4(cc, tt, 2);
  }

  // ------------------------------------------------------------------ NEVILL surfa.f:2-83
  // returns false on "too many cycles" (lstop)
  bool nevill(R tt, R c1, R c2, R del1, R del2, int ifunc, R& cc_out) {
    const R accur1 = R(0.1e-5f), accur2 = R(0.1e-7f);
    R x[21], y[21];
    int ic = 0;
    R c3 = (c1 + c2) / R(2);
    R del3 = dltar(c3, tt, ifunc); cnt.polish_evals++;
    int nev = 1;
    int m = 1;
    for (;;) {
      ic = ic + 1;
      if (!(ic < o.neville_cap)) { lstop += 10; return false; }
      bool inside = (c1 <= c3) ? (c2 > c3) : (c2 < c3);  // surfa.f:32-34
      bool bisect = !inside;
      if (inside) {
        R s13 = del1 - del3;
        R s32 = del3 - del2;
        if (sgn(del3) * sgn(del1) <= 0) { c2 = c3; del2 = del3; }
        else { c1 = c3; del1 = del3; }
        if (m_abs(c1 - c2) - accur1 <= R(0)) { cc_out = c3; return true; }
        if (sgn(s13) != sgn(s32)) nev = 0;
        R ss1 = m_abs(del1), s1 = R(0.1f) * ss1;
        R ss2 = m_abs(del2), s2 = R(0.1f) * ss2;
        if (s1 > ss2 || s2 > ss1) bisect = true;
        else if (nev == 0) bisect = true;
        else {
          if (nev == 2) { x[m + 1] = c3; y[m + 1] = del3; }  // 1350
          else { x[1] = c1; y[1] = del1; x[2] = c2; y[2] = del2; m = 1; }
          bool fail = false;
          for (int kk = 1; kk <= m; ++kk) {  // 1355
            int j = m - kk + 1;
            if (m_abs(y[m + 1] - y[j]) <= accur2) { fail = true; break; }
            x[j] = (-y[j] * x[j + 1] + y[m + 1] * x[j]) / (y[m + 1] - y[j]);
          }
          if (fail) bisect = true;
          else {
            c3 = x[1];
            del3 = dltar(c3, tt, ifunc); cnt.polish_evals++;
            nev = 2;
            m = m + 1;
            if (m > 10) m = 10;
            continue;
          }
        }
      }
      if (bisect) {  // 1344
        c3 = (c1 + c2) / R(2);
        del3 = dltar(c3, tt, ifunc); cnt.polish_evals++;
        nev = 1;
        m = 1;
      }
    }
  }

  // ------------------------------------------------------------------ LEIGEN surfa.f:374-631
  // working arrays hold the refreshed model (n layers); returns ugr, cvar
  bool leigen(R T, R cph, R& ugr_out, R& cvar_out) {
    int jm = mmax;  // COMMON mmax as seen through "jmax"
    int lm = jm, ln = jm;
    std::vector<R> ld(d.begin(), d.end()), lb(b.begin(), b.end()), lrho(rho.begin(), rho.end());
    int mm1 = lm - 1;
    int ivre = o.ndiv_cap_l / mm1;
    if (ndiv > ivre) ndiv = ivre;
    R div = R((float)ndiv);
    if (ndiv > 1) {  // surfa.f:418-445
      int jj = 1;
      if (lb[1] <= R(0.1e-10f)) jj = 2;
      int newm = (mm1 - jj + 1) * ndiv + jj;
      std::vector<R> nd(newm + 2), nb(newm + 2), nr(newm + 2);
      for (int j = 1; j < jj; ++j) { nd[j] = ld[j]; nb[j] = lb[j]; nr[j] = lrho[j]; }
      for (int j = jj; j <= mm1; ++j) {
        int ldiv = (j - jj) * ndiv;
        for (int i = 1; i <= ndiv; ++i) {
          nd[jj + ldiv + i - 1] = ld[j] / div;
          nb[jj + ldiv + i - 1] = lb[j];
          nr[jj + ldiv + i - 1] = lrho[j];
        }
      }
      nd[newm] = 0; nb[newm] = lb[ln]; nr[newm] = lrho[ln];
      ld.swap(nd); lb.swap(nb); lrho.swap(nr);
      lm = newm; ln = newm;
    }
    (void)ln;
    R t_ = T, c_ = cph;
    // layer dropping, surfa.f:475-487
    R any = fact;
    if (any <= R(0)) any = R(7.0f);
    R dmax = any * c_ * t_;
    {
      R sum = 0;
      int mx = lm;
      bool jumped = false;
      for (int ii = 1; ii <= lm; ++ii) {
        mx = ii;
        if (c_ - lb[ii] >= R(0)) continue;
        sum = sum + ld[ii];
        if (ii == lm) { jumped = true; break; }
        if (sum <= dmax) continue;
        R db = lb[ii + 1] - lb[ii];
        if (db < R(0)) { jumped = true; break; }
        if (db == R(0)) continue;
        mx = mx + 1; jumped = true; break;
      }
      if (!jumped) mx = std::min(mx + 1, lm);  // fall-through to 90009 (guarded against overrun)
      lm = mx;
    }
    R wvno = twopi_leig / (c_ * t_);
    R tw = twopi_leig / t_;
    R omegsq = tw * tw;
    const R const_lim = R(1.e10f), const_lim1 = R(1.e5f);
    R ut0 = 1;
    R ut, tq, sumi0, sumi1, sumi2;
    for (;;) {  // 7777
      ut = ut0;
      R covb = c_ / lb[lm];
      R h = lrho[lm] * lb[lm] * lb[lm];
      R rb = wvno * m_sqrt(m_abs(covb * covb - R(1)));
      tq = -h * rb * ut0;
      R dm, sm;
      if (rb == R(0)) { dm = R(1.0e25f); sm = 0; }
      else { dm = R(0.5f) / rb; sm = R(0.5f) * rb; }
      sumi0 = lrho[lm] * dm;
      sumi1 = h * dm;
      sumi2 = h * sm;
      bool restart = false;
      for (int k = 1; k <= lm - 1; ++k) {
        if (m_abs(ut) > const_lim) { ut0 = ut0 / const_lim1; restart = true; break; }
        int m = lm - k;
        if (lb[m] == R(0)) continue;
        cnt.sub_U++;
        covb = c_ / lb[m];
        rb = wvno * m_sqrt(m_abs(covb * covb - R(1)));
        h = lrho[m] * lb[m] * lb[m];
        R dz = ld[m] / R(4);
        R dmm[6], smm[6];
        dmm[1] = ut * ut;
        smm[1] = (tq / h) * (tq / h);
        R eut = ut, ett = tq;
        for (int kk = 2; kk <= 5; ++kk) {
          R xkk = R(kk - 1);
          R q = rb * dz * xkk;
          R y, z, cosq;
          if (c_ - lb[m] < R(0)) {  // 1207
            R exqp = m_exp(q);
            R exqm = R(1) / exqp;
            y = (exqp - exqm) / (R(2) * rb);
            z = rb * rb * y;
            cosq = (exqp + exqm) / R(2);
          } else if (c_ - lb[m] == R(0)) {  // 1221
            y = dz * xkk; z = 0; cosq = 1;
          } else {  // 1231
            R sinq = m_sin(q);
            y = sinq / rb;
            z = -rb * sinq;
            cosq = m_cos(q);
          }
          eut = cosq * ut - y * tq / h;
          ett = -h * z * ut + cosq * tq;
          dmm[kk] = eut * eut;
          smm[kk] = (ett * ett) / (h * h);
        }
        ut = eut;
        tq = ett;
        dm = (dz / R(22.5f)) * (R(7) * (dmm[1] + dmm[5]) + R(32) * (dmm[2] + dmm[4]) + R(12) * dmm[3]);
        sm = (dz / R(22.5f)) * (R(7) * (smm[1] + smm[5]) + R(32) * (smm[2] + smm[4]) + R(12) * smm[3]);
        sumi0 = sumi0 + lrho[m] * dm;
        sumi1 = sumi1 + h * dm;
        sumi2 = sumi2 + h * sm;
      }
      if (!restart) break;
    }
    sumi0 = sumi0 / (ut * ut);
    sumi1 = sumi1 / (ut * ut);
    sumi2 = sumi2 / (ut * ut);
    R wvar = (omegsq * sumi0 - sumi2) / sumi1;
    cvar_out = m_sqrt(omegsq / wvar);
    ugr_out = sumi1 / (c_ * sumi0);
    return true;
  }

  // ------------------------------------------------------------------ REIGEN surfa.f:714-1190
  bool reigen(R T, R cph, R ratio_, R& ugr_out, R& cvar_out) {
    int lm = mmax;
    std::vector<R> ld(d.begin(), d.end()), la(a.begin(), a.end()), lb(b.begin(), b.end()),
        lrho(rho.begin(), rho.end());
    int mm1 = lm - 1;
    int ivre = o.ndiv_cap_r / mm1;
    if (ndiv > ivre) ndiv = ivre;
    R div = R((float)ndiv);
    std::vector<int> orig(lm + 2);          // original layer of every (sub-)layer
    for (int j = 0; j <= lm + 1; ++j) orig[j] = j;
    const int lm_orig = lm;
    if (ndiv > 1) {  // surfa.f:787-820
      int jj = 1;
      if (lb[1] <= R(0.1e-10f)) jj = 2;
      int newm = (mm1 - jj + 1) * ndiv + jj;
      std::vector<R> nd(newm + 2), na(newm + 2), nb(newm + 2), nr(newm + 2);
      std::vector<int> no(newm + 2, 0);
      for (int j = 1; j < jj; ++j) { nd[j] = ld[j]; na[j] = la[j]; nb[j] = lb[j]; nr[j] = lrho[j]; no[j] = j; }
      for (int j = jj; j <= mm1; ++j) {
        int ldiv = (j - jj) * ndiv;
        for (int i = 1; i <= ndiv; ++i) {
          int q = jj + ldiv + i - 1;
          nd[q] = ld[j] / div; na[q] = la[j]; nb[q] = lb[j]; nr[q] = lrho[j]; no[q] = j;
        }
      }
      nd[newm] = 0; na[newm] = la[lm]; nb[newm] = lb[lm]; nr[newm] = lrho[lm]; no[newm] = lm;
      orig.swap(no);
      ld.swap(nd); la.swap(na); lb.swap(nb); lrho.swap(nr);
      lm = newm;
    }
    int ntot = lm;
    std::vector<R> xmu(ntot + 2), xlamb(ntot + 2);
    for (int i = 1; i <= ntot; ++i) {  // surfa.f:828-834
      xmu[i] = lrho[i] * lb[i] * lb[i];
      xlamb[i] = lrho[i] * (la[i] * la[i] - R(2) * lb[i] * lb[i]);
    }
    R t_ = T, c_ = cph;
    // layer dropping surfa.f:854-866
    R dmax = fact * t_ * c_;
    {
      R sum = 0;
      int mx = lm;
      bool jumped = false;
      for (int ii = 1; ii <= lm; ++ii) {
        mx = ii;
        if (c_ - lb[ii] >= R(0)) continue;
        sum = sum + ld[ii];
        if (ii == lm) { jumped = true; break; }
        if (sum <= dmax) continue;
        R da = la[ii + 1] - la[ii];
        if (da < R(0)) { jumped = true; break; }
        if (da > R(0)) { mx = mx + 1; jumped = true; break; }
        R db = lb[ii + 1] - lb[ii];
        if (db < R(0)) { jumped = true; break; }
        if (db == R(0)) continue;
        mx = mx + 1; jumped = true; break;
      }
      if (!jumped) mx = std::min(mx + 1, lm);
      lm = mx;
    }
    R sumi0 = 0, sumi1 = 0, sumi2 = 0, sumi3 = 0;
    R wvno = twopi_reig / (c_ * t_);
    R wvnosq = wvno * wvno;
    R omega = twopi_reig / t_;
    R omegsq = omega * omega;
    R tzz = 0;
    if (!(lb[1] > R(0))) {
      // water layer integrals, surfa.f:879-910 (complex sqrt written out by sign of c^2/a^2-1)
      R ra = c_ / la[1];
      R x = ra * ra - R(1);
      R mag = wvno * m_sqrt(m_abs(x));
      if (mag <= R(1.0e-35f)) {
        sumi0 = lrho[1] * ld[1]; sumi1 = sumi2 = sumi3 = 0; tzz = 0;
      } else {
        R sin2ra, cosra, rab1, sdr;
        if (x >= R(0)) {  // cra real
          sin2ra = m_sin(R(2) * mag * ld[1]) / (R(4) * mag);
          cosra = m_cos(mag * ld[1]);
          rab1 = mag * mag;
          sdr = m_sin(mag * ld[1]) / mag;
        } else {  // cra = i*mag
          R e2 = m_exp(R(2) * mag * ld[1]);
          sin2ra = (R(0.5f) * (e2 - R(1) / e2)) / (R(4) * mag);
          R e1 = m_exp(mag * ld[1]);
          cosra = R(0.5f) * (e1 + R(1) / e1);
          rab1 = -(mag * mag);
          sdr = (R(0.5f) * (e1 - R(1) / e1)) / mag;
        }
        R cos2rm = R(1) / (cosra * cosra);
        R fac1 = (R(0.5f) * ld[1] + sin2ra) * cos2rm;
        R fac3 = wvno * (R(0.5f) * ld[1] - sin2ra) * cos2rm;
        R fac2 = wvno * fac3 / rab1;
        R fac4 = rab1 * fac3 / wvno;
        sumi0 = lrho[1] * (fac1 + fac2);
        sumi1 = xlamb[1] * fac2;
        sumi2 = xlamb[1] * fac3;
        sumi3 = xlamb[1] * fac4;
        tzz = -lrho[1] * omegsq * sdr / cosra;
      }
    }
    // half-space, surfa.f:913-926
    R cova = c_ / la[lm];
    R covb = c_ / lb[lm];
    R gam = R(2) / (covb * covb);
    R gamm1 = gam - R(1);
    R ra = wvno * m_sqrt(m_abs(cova * cova - R(1)));
    R rb = wvno * m_sqrt(m_abs(covb * covb - R(1)));
    R det = wvnosq - ra * rb;
    R h = lrho[lm] * omegsq;
    R brkt = -gamm1 * wvno + gam * ra * rb / wvno;
    int iter = 0;
    const R wwt[5] = {0, 0, R(0.5f), R(0.5f), R(1.0f)};
    const R wt[5] = {0, R(1) / R(6), R(1) / R(3), R(1) / R(3), R(1) / R(6)};
    typedef double D;
    // yy/yz [m][k], k=1..5 ; fill values surfa.f:749-759
    std::vector<D> yy1((lm + 2) * 6, 1.0), yy2((lm + 2) * 6, 1.0), yy3((lm + 2) * 6, 1.0), yy4((lm + 2) * 6, 1.0);
    std::vector<D> yz1((lm + 2) * 6, 2.0), yz2((lm + 2) * 6, 1.0), yz3((lm + 2) * 6, 1.0), yz4((lm + 2) * 6, 1.0);
    auto IX = [](int m, int k) { return m * 6 + k; };
    auto integrate = [&](D ur, D uz, D tz, D tr, std::vector<D>& y1, std::vector<D>& y2,
                         std::vector<D>& y3, std::vector<D>& y4) {
      for (int mm = 1; mm <= lm - 1; ++mm) {  // surfa.f:928-979 / 999-1049
        int m = lm - mm;
        if (lb[m] <= R(0)) continue;
        cnt.sub_U++;
        R xdiv = 1;
        R ddz = -ld[m] / (R(4) * xdiv);
        R a12 = R(1) / (xlamb[m] + R(2) * xmu[m]);
        R a13 = wvno * xlamb[m] * a12;
        R a21 = -omegsq * lrho[m];
        R a24 = wvno;
        R a31 = -wvno;
        R a34 = R(1) / xmu[m];
        R a42 = -a13;
        R a43 = a21 + R(4) * wvnosq * xmu[m] * (xlamb[m] + xmu[m]) * a12;
        y3[IX(m, 5)] = ur; y1[IX(m, 5)] = uz; y2[IX(m, 5)] = tz; y4[IX(m, 5)] = tr;
        for (int kk = 2; kk <= 5; ++kk) {
          int k = 6 - kk;
          D eur = ur, euz = uz, etz = tz, etr = tr;
          D dur = 0, duz = 0, dtz = 0, dtr = 0;
          for (int ll = 1; ll <= 4; ++ll) {
            R w = wwt[ll] * ddz;
            D sur = ur + w * dur;
            D suz = uz + w * duz;
            D stz = tz + w * dtz;
            D str = tr + w * dtr;
            dur = a31 * suz + a34 * str;
            duz = a12 * stz + a13 * sur;
            dtz = a21 * suz + a24 * str;
            dtr = a42 * stz + a43 * sur;
            R v = wt[ll] * ddz;
            eur = eur + v * dur;
            euz = euz + v * duz;
            etz = etz + v * dtz;
            etr = etr + v * dtr;
          }
          ur = eur; uz = euz; tz = etz; tr = etr;
          y1[IX(m, k)] = uz; y2[IX(m, k)] = tz; y3[IX(m, k)] = ur; y4[IX(m, k)] = tr;
        }
      }
      if (!(lb[1] > R(0))) {  // surfa.f:980-984
        y1[IX(1, 1)] = y1[IX(2, 1)]; y2[IX(1, 1)] = y2[IX(2, 1)];
        y3[IX(1, 1)] = y3[IX(2, 1)]; y4[IX(1, 1)] = y4[IX(2, 1)];
      }
    };
    R s_atz1 = -h * brkt / det, s_atr1 = -h * ra / det;
    integrate(1.0, 0.0, D(s_atz1), D(s_atr1), yy1, yy2, yy3, yy4);
    D xnorm = 0, bb = 0;
    for (;;) {  // 4003
      D aur2 = 0.0, auz2 = 1.0;
      D atz2 = D(R(-h * rb / det));
      D atr2 = D(R(-h * brkt / det));
      if (iter != 0) {
        D aur1 = 1.0, auz1 = 0.0, atz1 = D(s_atz1), atr1 = D(s_atr1);
        aur2 = aur2 + xnorm * aur1;
        auz2 = auz2 + xnorm * auz1;
        atz2 = atz2 + xnorm * atz1;
        atr2 = atr2 + xnorm * atr1;
      }
      integrate(aur2, auz2, atz2, atr2, yz1, yz2, yz3, yz4);
      D aa = yz3[IX(1, 1)] - ratio_ * yz1[IX(1, 1)];
      bb = ratio_ * yy1[IX(1, 1)] - yy3[IX(1, 1)];
      if (fabs(bb) < 1.e-10) bb = copysign(1.e-10, bb);
      xnorm = aa / bb;
      bb = xnorm * yy1[IX(1, 1)] + yz1[IX(1, 1)];
      if (fabs(bb) < 1.e-10) bb = copysign(1.e-10, bb);
      R ampur_ns = R((xnorm * yy3[IX(1, 1)] + yz3[IX(1, 1)]) / bb);
      iter = iter + 1;
      if (iter > 1) break;
      R xtest = m_abs(ampur_ns / ratio_ - R(1));
      if (xtest >= R(0.00001f)) continue;
      break;
    }
    // integrals surfa.f:1087-1135
    R aur = 0, auz = 0, atz = 0, atr = 0;
    int m;
    bool early = false;
    std::vector<double> sub_dcda(ntot + 2, 0.0), sub_dcdb(ntot + 2, 0.0), sub_dcdr(ntot + 2, 0.0);
    for (m = 1; m <= lm; ++m) {
      if (lb[m] <= R(0)) continue;
      if (m >= lm) break;
      R dz = ld[m] / R(4);
      R dmr[6], dmz[6], smr[6], smz[6], dmrsmz[6], dmzsmr[6];
      for (int kk = 1; kk <= 5; ++kk) {
        aur = R((xnorm * yy3[IX(m, kk)] + yz3[IX(m, kk)]) / bb);
        auz = R((xnorm * yy1[IX(m, kk)] + yz1[IX(m, kk)]) / bb);
        atz = R((xnorm * yy2[IX(m, kk)] + yz2[IX(m, kk)]) / bb);
        atr = R((xnorm * yy4[IX(m, kk)] + yz4[IX(m, kk)]) / bb);
        R durdz = atr / xmu[m] - wvno * auz;
        R duzdz = (atz + wvno * xlamb[m] * aur) / (xlamb[m] + R(2) * xmu[m]);
        dmr[kk] = aur * aur;
        dmz[kk] = auz * auz;
        smr[kk] = durdz * durdz;
        smz[kk] = duzdz * duzdz;
        dmrsmz[kk] = aur * duzdz;
        dmzsmr[kk] = auz * durdz;
      }
      const R q = dz / R(22.5f);
      D dmmr = q * (R(7) * (dmr[1] + dmr[5]) + R(32) * (dmr[2] + dmr[4]) + R(12) * dmr[3]);
      D dmmz = q * (R(7) * (dmz[1] + dmz[5]) + R(32) * (dmz[2] + dmz[4]) + R(12) * dmz[3]);
      D smmz = q * (R(7) * (smz[1] + smz[5]) + R(32) * (smz[2] + smz[4]) + R(12) * smz[3]);
      D smmr = q * (R(7) * (smr[1] + smr[5]) + R(32) * (smr[2] + smr[4]) + R(12) * smr[3]);
      D drsz = q * (R(7) * (dmrsmz[1] + dmrsmz[5]) + R(32) * (dmrsmz[2] + dmrsmz[4]) + R(12) * dmrsmz[3]);
      D dzsr = q * (R(7) * (dmzsmr[1] + dmzsmr[5]) + R(32) * (dmzsmr[2] + dmzsmr[4]) + R(12) * dmzsmr[3]);
      sumi0 = R(sumi0 + lrho[m] * (dmmr + dmmz));
      sumi1 = R((xlamb[m] + R(2) * xmu[m]) * dmmr + xmu[m] * dmmz + sumi1);
      sumi2 = R(xmu[m] * dzsr - xlamb[m] * drsz + sumi2);
      sumi3 = R((xlamb[m] + R(2) * xmu[m]) * smmz + xmu[m] * smmr + sumi3);
      if (want_partials) {  // surfa.f:1130-1135
        D dldl = -wvnosq * dmmr + R(2) * wvno * drsz - smmz;
        D dldm = -wvnosq * (R(2) * dmmr + dmmz) - R(2) * wvno * dzsr - (R(2) * smmz + smmr);
        D dldr = omegsq * (dmmr + dmmz);
        sub_dcdb[m] = R(2) * lrho[m] * lb[m] * c_ * (dldm - R(2) * dldl) / wvno;
        sub_dcda[m] = R(2) * lrho[m] * la[m] * c_ * dldl / wvno;
        sub_dcdr[m] = (c_ / wvno) * (dldr + xlamb[m] * dldl / lrho[m] + xmu[m] * dldm / lrho[m]);
      }
      if (m_abs(auz) + m_abs(aur) - R(1.0e-15f) <= R(0)) { early = true; break; }  // -> 7002
    }
    if (m > lm) m = lm;
    if (!early) {
      // 77777, surfa.f:1140-1144
      if (!((lb[1] > R(0.1e-10f)) || m != 2)) { aur = ratio_; auz = 1; atr = 0; atz = tzz; }
    }
    // 7002: half-space tail surfa.f:1151-1178
    {
      R ap = -lrho[m] * (wvno * aur + rb * auz) / det;
      R bp = -lrho[m] * (-ra * aur / wvno - auz) / det;
      R a1 = -wvno * ap / lrho[m];
      R a2 = -wvno * rb * bp / lrho[m];
      R a3 = ra * ap / lrho[m];
      R a4 = wvnosq * bp / lrho[m];
      if (rb == R(0)) {  // 7006
        ugr_out = lb[m];
        cvar_out = 0;
        return true;
      }
      D dmmr = a1 * a1 / (R(2) * ra) + R(2) * a1 * a2 / (ra + rb) + a2 * a2 / (R(2) * rb);
      D dmmz = a3 * a3 / (R(2) * ra) + R(2) * a3 * a4 / (ra + rb) + a4 * a4 / (R(2) * rb);
      D smmz = ra * a3 * a3 / R(2) + R(2) * ra * rb * a3 * a4 / (ra + rb) + rb * a4 * a4 / R(2);
      D smmr = ra * a1 * a1 / R(2) + R(2) * ra * rb * a1 * a2 / (ra + rb) + rb * a2 * a2 / R(2);
      D drsz = -a1 * a3 / R(2) - (a1 * a4 * rb + a2 * a3 * ra) / (ra + rb) - a2 * a4 / R(2);
      D dzsr = -a1 * a3 / R(2) - (a1 * a4 * ra + a2 * a3 * rb) / (ra + rb) - a2 * a4 / R(2);
      sumi0 = R(sumi0 + lrho[m] * (dmmr + dmmz));
      sumi1 = R((xlamb[m] + R(2) * xmu[m]) * dmmr + xmu[m] * dmmz + sumi1);
      sumi2 = R(xmu[m] * dzsr - xlamb[m] * drsz + sumi2);
      sumi3 = R((xlamb[m] + R(2) * xmu[m]) * smmz + xmu[m] * smmr + sumi3);
      if (want_partials) {  // surfa.f:1179-1185 (half-space), 1202-1208 (normalisation by dL/dk)
        D dldr = omegsq * (dmmr + dmmz);
        D dldm = -wvnosq * (R(2) * dmmr + dmmz) - R(2) * wvno * dzsr - (R(2) * smmz + smmr);
        D dldl = -wvnosq * dmmr + R(2) * wvno * drsz - smmz;
        sub_dcda[m] = R(2) * lrho[m] * la[m] * c_ * dldl / wvno;
        sub_dcdb[m] = R(2) * lrho[m] * lb[m] * c_ * (dldm - R(2) * dldl) / wvno;
        sub_dcdr[m] = (c_ / wvno) * (dldr + xlamb[m] * dldl / lrho[m] + xmu[m] * dldm / lrho[m]);
        const D dldk = -R(2) * (wvno * sumi1 + sumi2);
        p_dcda.assign(lm_orig + 1, 0.0); p_dcdb.assign(lm_orig + 1, 0.0); p_dcdr.assign(lm_orig + 1, 0.0);
        const int first = (lb[1] <= R(0)) ? 2 : 1;
        for (int q = first; q <= m; ++q) {
          p_dcda[orig[q]] += sub_dcda[q] / dldk; p_dcdb[orig[q]] += sub_dcdb[q] / dldk; p_dcdr[orig[q]] += sub_dcdr[q] / dldk;
        }
      }
    }
    ugr_out = (wvno * sumi1 + sumi2) / (omega * sumi0);  // surfa.f:1186
    R wvar = (-sumi2 + m_sqrt(m_abs(sumi2 * sumi2 - sumi1 * (sumi3 - omegsq * sumi0)))) / sumi1;
    cvar_out = omega / wvar;
    return true;
  }

  // ---------------------------------------------- FAST_SURF fast_surf.f:2-211 + calcul calcul.f:2-408
  // returns status: 0 ok, 1 no root at first period, 2 stopped early (no root at k>1), 3 LSTOP abort
  int run(int kind_, int n, const double* ia, const double* ib, const double* irho, const double* id,
          const double* iqs, int nper, const double* per) {
    kind = kind_;
    setup_consts();
    int sz = std::max(n + 2, 8);
    a_ref.assign(sz, 0); b_ref.assign(sz, 0); rho_ref.assign(sz, 0); d_ref.assign(sz, 0); qs_ref.assign(sz, 0);
    a.assign(sz, 0); b.assign(sz, 0); rho.assign(sz, 0); d.assign(sz, 0);
    for (int i = 1; i <= n; ++i) {  // fast_surf.f:89-99 (f2py casts inputs to real*4)
      a_ref[i] = RM(ia[i - 1]); b_ref[i] = RM(ib[i - 1]); rho_ref[i] = RM(irho[i - 1]);
      d_ref[i] = RM(id[i - 1]); qs_ref[i] = RM(iqs[i - 1]);
    }
    // INIT init.f:58-73
    mode = o.nmode;
    kmax = nper;
    t.assign(kmax + 2, 0);
    for (int i = 1; i <= kmax; ++i) t[i] = RM(per[i - 1]);
    ndiv = o.ndiv;
    c.assign(mode + 1, std::vector<R>(kmax + 2, R(0)));
    ratio = c; ugr = c; cvar = c;
    imax.assign(mode + 1, 0);
    // fast_surf.f:113,150-171
    mmax = n; nmax = n; lstop = 0; idrop = 0;
    int ilay = 1;
    if (o.sibling ? (b_ref[1] == RM(0)) : (b_ref[1] < RM(0.1f))) ilay = 2;
    RM b_corr = 0;
    RM t1m = t[1];
    if (o.atten) b_corr = qs_ref[ilay] * m_log(t_base / t1m) / pi_att;
    RM qq = b_ref[ilay];
    if (kind == 2) qq = RM(o.sibling ? 0.9 : 0.9f) * qq;
    RM c1m = qq * (RM(1) + b_corr);
    if (!o.sibling && b_ref[1] < RM(0.1f)) c1m = RM(0.5f);  // fast_surf.f:171
    R c1 = R(c1m);
    int ifunc = kind;
    int kmode = mode;
    int status = 0;
    bool abort_all = false;
    // ---- phase 1: calcul.f:104-220
    for (int k = 1; k <= kmax && !abort_all; ++k) {
      RM Tm = t[k];
      R t1 = R(Tm);
      int mref = o.stale_mmax ? mmax : nmax;
      refresh(mref, Tm);
      bool stop_periods = false;
      for (int iq = 1; iq <= kmode; ++iq) {
        if (k > 1) {
          if (iq < 2) c1 = R(o.sibling ? 0.90 : 0.90f) * c[1][k - 1];
          else {
            R x = dc * (c[iq][k - 1] - c[iq - 1][k]);
            if (x > R(0)) c1 = c[iq][k - 1];
            else c1 = c[iq - 1][k] + R(o.sibling ? 0.01 : 0.01f) * dc;
          }
        }
        idrop = 0;
        R del1 = dltar(c1, t1, ifunc); cnt.scan_evals++;
        R c2 = c1, del2 = del1;
        bool found = false, fail = false;
        for (;;) {  // label 80
          c2 = c1 + dc;
          idrop = 0;
          del2 = dltar(c2, t1, ifunc); cnt.scan_evals++;
          if (sgn(del1) != sgn(del2)) { found = true; break; }
          c1 = c2;
          del1 = del2;
          if (c1 - (o.sibling ? R(0.8) : R(0.8f)) * b[1] < R(0)) { fail = true; break; }
          if (o.sibling) { if (c1 - (b[mmax] + R(0.3)) > R(0)) { fail = true; break; } }
          else { if (!(c1 - (b[mmax] + R(0.3f)) < R(0))) { fail = true; break; } }
          if (!(c1 == c1)) { fail = true; break; }  // NaN guard (reference would spin)
        }
        if (getenv("ORACLE_DEBUG")) fprintf(stderr, "k=%d iq=%d T=%g found=%d fail=%d c1=%.7f c2=%.7f del1=%g del2=%g mmax=%d b1=%.6f bmm=%.6f\n", k, iq, (double)t1, (int)found, (int)fail, (double)c1, (double)c2, (double)del1, (double)del2, mmax, (double)b[1], (double)b[mmax]);
        if (found) {
          R cn = 0;
          if (!nevill(t1, c1, c2, del1, del2, ifunc, cn)) {
            lstop = 0; status = 3; abort_all = true; break;  // calcul.f:173-189 -> 9999
          }
          c1 = cn;
          if (c1 - b[mmax] > R(0)) fail = true;  // calcul.f:191
          else {
            c[iq][k] = c1;
            if (ifunc == 2) ratio[iq][k] = dltar(c1, t1, 3);
            c1 = c1 + R(o.sibling ? 0.01 : 0.01f) * dc;
            imax[iq] = k;
            continue;
          }
        }
        if (fail) {  // label 250
          if (k * iq <= 1) { status = 1; abort_all = true; break; }
          kmode = iq - 1;
          if (kmode <= 0) { status = 2; stop_periods = true; }
          break;
        }
      }
      if (stop_periods) break;
    }
    if (abort_all) {
      // calcul.f:9999 -- phase 2 skipped; reference would copy stale data: report none
      for (int iq = 1; iq <= mode; ++iq) imax[iq] = 0;
      return status;
    }
    // ---- phase 2: calcul.f:224-404
    mmax = nmax;
    for (int iq = 1; iq <= mode; ++iq) {
      int j = imax[iq];
      for (int lip = 1; lip <= j; ++lip) {
        refresh(mmax, t[lip]);
        R u = 0, cv = 0;
        if (ifunc == 1) leigen(R(t[lip]), c[iq][lip], u, cv);
        else reigen(R(t[lip]), c[iq][lip], ratio[iq][lip], u, cv);
        ugr[iq][lip] = u;
        cvar[iq][lip] = cv;
        if (want_partials && ifunc == 2 && iq == 1) {
          if ((int)all_dcda.size() < kmax + 1) { all_dcda.resize(kmax + 1); all_dcdb.resize(kmax + 1); all_dcdr.resize(kmax + 1); }
          all_dcda[lip] = p_dcda; all_dcdb[lip] = p_dcdb; all_dcdr[lip] = p_dcdr;
        }
      }
    }
    return status;
  }
};

template <typename RM, typename R>
int run_one(const OracleOpts& o, int kind, int n, const double* a, const double* b, const double* rho,
            const double* d, const double* qs, int nper, const double* per, double* c_out,
            double* u_out, double* ratio_out, double* cvar_out, int* imax_out, OracleCounters* cnt) {
  Solver<RM, R> s;
  s.o = o;
  int st = s.run(kind, n, a, b, rho, d, qs, nper, per);
  for (int iq = 1; iq <= o.nmode; ++iq) {
    imax_out[iq - 1] = s.imax[iq];
    for (int k = 1; k <= nper; ++k) {
      bool ok = k <= s.imax[iq];
      size_t ix = (size_t)(iq - 1) * nper + (k - 1);
      c_out[ix] = ok ? (double)s.c[iq][k] : 0.0;
      u_out[ix] = ok ? (double)s.ugr[iq][k] : 0.0;
      if (ratio_out) ratio_out[ix] = ok ? (double)s.ratio[iq][k] : 0.0;
      if (cvar_out) cvar_out[ix] = ok ? (double)s.cvar[iq][k] : 0.0;
    }
  }
  if (cnt) {
    cnt->sweeps_R += s.cnt.sweeps_R; cnt->steps_R += s.cnt.steps_R;
    cnt->sweeps_L += s.cnt.sweeps_L; cnt->steps_L += s.cnt.steps_L;
    cnt->sub_U += s.cnt.sub_U; cnt->flat_layers += s.cnt.flat_layers;
    cnt->scan_evals += s.cnt.scan_evals; cnt->polish_evals += s.cnt.polish_evals;
  }
  return st;
}

int dispatch(const OracleOpts& o, int kind, int n, const double* a, const double* b, const double* rho,
             const double* d, const double* qs, int nper, const double* per, double* c_out,
             double* u_out, double* ratio_out, double* cvar_out, int* imax_out, OracleCounters* cnt) {
  if (o.precision == 0)
    return run_one<float, float>(o, kind, n, a, b, rho, d, qs, nper, per, c_out, u_out, ratio_out, cvar_out, imax_out, cnt);
  if (o.precision == 1)
    return run_one<float, double>(o, kind, n, a, b, rho, d, qs, nper, per, c_out, u_out, ratio_out, cvar_out, imax_out, cnt);
  return run_one<double, double>(o, kind, n, a, b, rho, d, qs, nper, per, c_out, u_out, ratio_out, cvar_out, imax_out, cnt);
}

}  // namespace

// Rayleigh phase-velocity partial derivatives of the fundamental mode (REIGEN, surfa.f:1130-1135, 1179-1185, 1202-1208):
// dcda, dcdb, dcdr [nper][n] with respect to Vp, Vs, density of every layer of the period's (attenuation-corrected,
// flattened) model; rows of periods without a root are zero.  Returns the oracle status.
template <typename RM, typename R>
static int partials_one(const OracleOpts& o, int n, const double* a, const double* b, const double* rho, const double* d,
                        const double* qs, int nper, const double* per, double* c_out, double* dcda, double* dcdb, double* dcdr) {
  Solver<RM, R> s;
  s.o = o;
  s.want_partials = true;
  const int st = s.run(2, n, a, b, rho, d, qs, nper, per);
  for (int k = 1; k <= nper; ++k) {
    const bool ok = k <= s.imax[1];
    if (c_out) c_out[k - 1] = ok ? (double)s.c[1][k] : 0.0;
    for (int i = 1; i <= n; ++i) {
      const bool have = ok && k < (int)s.all_dcda.size() && i < (int)s.all_dcda[k].size();
      dcda[(size_t)(k - 1) * n + i - 1] = have ? s.all_dcda[k][i] : 0.0;
      dcdb[(size_t)(k - 1) * n + i - 1] = have ? s.all_dcdb[k][i] : 0.0;
      dcdr[(size_t)(k - 1) * n + i - 1] = have ? s.all_dcdr[k][i] : 0.0;
    }
  }
  return st;
}

extern "C" {

void surfdisp_oracle_default_opts(OracleOpts* o) {
  o->precision = 0; o->sibling = 0; o->nmode = 1; o->ndiv = 5; o->ndiv_cap_r = 99; o->ndiv_cap_l = 999;
  o->neville_cap = 50; o->stale_mmax = 1; o->atten = 1; o->flat = 1; o->dc = 0.01; o->fact = 4.0; o->t_base = 1.0;
}

// One model.  Layer arrays are [n] (top to half-space): a=Vp, b=Vs, rho, d=thickness, qs=1/Qs.
// Outputs are [nmode][nper] (row-major), zero beyond imax[mode].
int surfdisp_oracle(const OracleOpts* o, int kind, int n, const double* a, const double* b,
                    const double* rho, const double* d, const double* qs, int nper, const double* per,
                    double* c_out, double* u_out, double* ratio_out, double* cvar_out, int* imax_out,
                    OracleCounters* cnt) {
  if (n < 2 || n > NSIZE - 1 || nper < 1 || (kind != 1 && kind != 2) || o->nmode < 1) return -1;
  return dispatch(*o, kind, n, a, b, rho, d, qs, nper, per, c_out, u_out, ratio_out, cvar_out, imax_out, cnt);
}

// Batch of M models sharing the period list.  layers is [5][M][lmax] float64 in the order
// (a, b, rho, d, qs); nlay[M].  Outputs [M][nper] (mode 1 only), status[M], nfound[M].
// nthreads>1 splits models over std::threads (the reference itself is one model per process).
int surfdisp_oracle_batch(const OracleOpts* o, int kind, int M, int lmax, const int* nlay,
                          const double* layers, int nper, const double* per, double* c_out,
                          double* u_out, int* nfound, int* status, OracleCounters* cnt, int nthreads) {
  if (o->nmode != 1) return -1;
  if (nthreads < 1) nthreads = 1;
  std::vector<OracleCounters> cs(nthreads);
  for (auto& x : cs) memset(&x, 0, sizeof(x));
  auto work = [&](int tid) {
    std::vector<double> rt(nper), cv(nper);
    for (int i = tid; i < M; i += nthreads) {
      const size_t off = (size_t)i * lmax;
      const size_t pl = (size_t)M * lmax;
      int im = 0;
      int st = surfdisp_oracle(o, kind, nlay[i], layers + 0 * pl + off, layers + 1 * pl + off,
                               layers + 2 * pl + off, layers + 3 * pl + off, layers + 4 * pl + off, nper,
                               per, c_out + (size_t)i * nper, u_out + (size_t)i * nper, rt.data(), cv.data(),
                               &im, &cs[tid]);
      nfound[i] = im;
      status[i] = st;
    }
  };
  if (nthreads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int tI = 0; tI < nthreads; ++tI) th.emplace_back(work, tI);
    for (auto& x : th) x.join();
  }
  if (cnt) {
    memset(cnt, 0, sizeof(*cnt));
    for (auto& x : cs) {
      cnt->sweeps_R += x.sweeps_R; cnt->steps_R += x.steps_R; cnt->sweeps_L += x.sweeps_L;
      cnt->steps_L += x.steps_L; cnt->sub_U += x.sub_U; cnt->flat_layers += x.flat_layers;
      cnt->scan_evals += x.scan_evals; cnt->polish_evals += x.polish_evals;
    }
  }
  return 0;
}


int surfdisp_oracle_partials(const OracleOpts* o, int n, const double* a, const double* b, const double* rho, const double* d,
                             const double* qs, int nper, const double* per, double* c_out, double* dcda, double* dcdb,
                             double* dcdr) {
  if (!o || n < 2 || nper < 1 || !dcda || !dcdb || !dcdr) return -1;
  if (o->precision == 0) return partials_one<float, float>(*o, n, a, b, rho, d, qs, nper, per, c_out, dcda, dcdb, dcdr);
  if (o->precision == 1) return partials_one<float, double>(*o, n, a, b, rho, d, qs, nper, per, c_out, dcda, dcdb, dcdr);
  return partials_one<double, double>(*o, n, a, b, rho, d, qs, nper, per, c_out, dcda, dcdb, dcdr);
}

}  // extern "C"
