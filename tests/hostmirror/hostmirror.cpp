// tests/hostmirror/hostmirror.cpp -- TEST ONLY.
// Compiles pysurfinv_b200/csrc/surfdisp_core.cuh (the per-lane arithmetic of the CUDA kernels) with
// g++ and replays the phase-1 group algorithm of surfdisp_kernels.cu lane by lane on the CPU, so the
// kernel math and the scan / G-section / secant design can be checked against the oracle in the
// "not gpu" test tier.  Never loaded by the product package.
#include <vector>
#include <cstring>
#include <cmath>
#include "../../pysurfinv_b200/csrc/surfdisp_core.cuh"

using namespace sd;

extern "C" {

// one model, shared periods; returns nfound
int hm_forward(int G, int kind, int n, const float* a, const float* b, const float* rho, const float* d,
               const float* qs, int K, const float* per, float dc, float fact, float t_base, int atten,
               int flatten, int stale, int ndiv0, int ndiv_cap, float* c_out, float* u_out, float* ratio_out,
               long long* sweeps) {
  const int ld = n;
  std::vector<float> cst((size_t)NCONST * ld);
  prep_model(n, kind, flatten, a, b, rho, d, qs, cst.data(), ld);
  std::vector<float4> q0(n + 1), q1(n + 1);
  std::vector<float> lt(K);
  for (int k = 0; k < K; ++k) { lt[k] = logf(t_base / per[k]); c_out[k] = 0; u_out[k] = 0; ratio_out[k] = 0; }
  float c1;
  {
    const float b0 = cst[C_BREF * ld + 0];
    const int ilay = (b0 < 0.1f) ? 1 : 0;
    float b_corr = 0.f;
    if (atten) b_corr = SD_DIV(SD_MUL(cst[C_QS * ld + ilay], lt[0]), SD_PI_ATT);
    float qq = cst[C_BREF * ld + ilay];
    if (kind == 2) qq = SD_MUL(0.9f, qq);
    c1 = SD_MUL(qq, SD_ADD(1.0f, b_corr));
    if (b0 < 0.1f) c1 = 0.5f;
  }
  int mm = n, nfound = 0;
  float c_prev = 0.f;
  long long nsw = 0;
  std::vector<float> cj(G), dj(G);
  std::vector<int> mj(G);
  auto sweep = [&](float c, float T, int m) {
    nsw++;
    return (kind == 2) ? rayleigh_sweep(c, T, m, q0.data(), q1.data(), 1) : love_sweep(c, T, m, q0.data(), q1.data());
  };
  for (int k = 0; k < K; ++k) {
    const float T = per[k];
    const int mref = stale ? mm : n;
    for (int i = 0; i < mref; ++i) {
      const bool hs = (i == mref - 1);
      float aa, bb;
      layer_ab(cst.data(), ld, i, lt[k], atten, hs, aa, bb);
      const float r = hs ? cst[C_RHOHS * ld + i] : cst[C_RHOFL * ld + i];
      const float dd = hs ? 0.f : cst[C_DFL * ld + i];
      LayerRec rec = make_rec(aa, bb, r, dd);
      q0[i] = rec.q0; q1[i] = rec.q1;
    }
    if (k > 0) c1 = SD_MUL(0.90f, c_prev);
    const float b_top = q1[0].y;
    float lo = 0, hi = 0, dlo = 0, dhi = 0;
    bool found = false, done = false;
    float cbase = c1, cl_prev = 0, dl_prev = 0;
    bool have_prev = false;
    int mnew = mm;
    for (int round = 0; round < 4096 && !done; ++round) {
      for (int g = 0; g < G; ++g) {
        float c = cbase;
        for (int t = 0; t < g; ++t) c = SD_ADD(c, dc);
        cj[g] = c; mj[g] = layer_drop(c, T, fact, n, q1.data()); dj[g] = sweep(c, T, mj[g]);
      }
      for (int g = 0; g < G && !done; ++g) {
        const bool hasp = (g > 0) || have_prev;
        const float dp = g ? dj[g - 1] : dl_prev, cp = g ? cj[g - 1] : cl_prev;
        const bool change = hasp && (std::signbit(dp) != std::signbit(dj[g]));
        const float b_hs = q1[mj[g] - 1].y;
        const bool stop = hasp && !change && ((cj[g] < 0.8f * b_top) || !(cj[g] < b_hs + 0.3f) || !(cj[g] == cj[g]));
        if (change || stop) { found = change; lo = cp; hi = cj[g]; dlo = dp; dhi = dj[g]; mnew = mj[g]; done = true; }
      }
      if (!done) { cl_prev = cj[G - 1]; dl_prev = dj[G - 1]; have_prev = true; cbase = SD_ADD(cl_prev, dc); }
    }
    mm = mnew;
    float croot = 0;
    if (found) {
      const float lo0 = lo, hi0 = hi, dlo0 = dlo, dhi0 = dhi;
      bool multi = false;
      for (int it = 0; it < 12 && (hi - lo) > 2.0e-5f; ++it) {
        const float step = (hi - lo) / (float)(G + 1);
        for (int g = 0; g < G; ++g) { cj[g] = lo + (float)(g + 1) * step; dj[g] = sweep(cj[g], T, mm); }
        int j = -1, nchg = 0;
        for (int g = 0; g < G; ++g) {
          const float dp = g ? dj[g - 1] : dlo;
          if (std::signbit(dp) != std::signbit(dj[g])) { if (j < 0) j = g; nchg++; }
        }
        if (std::signbit(dj[G - 1]) != std::signbit(dhi)) nchg++;
        if (it == 0 && nchg > 1) { multi = true; break; }
        if (j >= 0) {
          const float nlo = j ? cj[j - 1] : lo, ndlo = j ? dj[j - 1] : dlo;
          hi = cj[j]; dhi = dj[j]; lo = nlo; dlo = ndlo;
        } else { lo = cj[G - 1]; dlo = dj[G - 1]; }
      }
      if (!multi) {
        const float den = dhi - dlo;
        float cs = (den != 0.f) ? lo - dlo * (hi - lo) / den : 0.5f * (lo + hi);
        if (!(cs >= lo && cs <= hi)) cs = 0.5f * (lo + hi);
        croot = cs;
      } else {
        int ev_n = 0;
        auto f = [&](float cc) { return sweep(cc, T, mm); };
        if (!nevill_seq(f, lo0, hi0, dlo0, dhi0, croot, ev_n)) { nfound = 0; break; }
      }
      if (croot > q1[mm - 1].y) found = false;
    }
    if (!found) break;
    float ratio = 0;
    if (kind == 2) {
      const float r12 = rayleigh_sweep(croot, T, mm, q0.data(), q1.data(), 2);
      const float r3 = rayleigh_sweep(croot, T, mm, q0.data(), q1.data(), 3);
      nsw += 2;
      ratio = 0.5f * r3 / r12;
    }
    c_out[k] = croot; ratio_out[k] = ratio; c_prev = croot; nfound = k + 1;
  }
  // phase 2
  for (int k = 0; k < nfound; ++k) {
    ModelView mv;
    mv.cst = cst.data(); mv.ld = ld; mv.n = n; mv.atten = atten; mv.lt = lt[k];
    int ndiv = ndiv0;
    const int ivre = ndiv_cap / (n - 1);
    if (ndiv > ivre) ndiv = ivre;
    mv.ndiv = ndiv;
    mv.jj0 = (cst[C_BREF * ld + 0] <= 0.1e-10f) ? 1 : 0;
    unsigned long long ns = 0;
    u_out[k] = (kind == 2) ? reigen_thread(mv, per[k], c_out[k], ratio_out[k], fact, ns)
                           : leigen_thread(mv, per[k], c_out[k], fact, ns);
  }
  if (sweeps) *sweeps += nsw;
  return nfound;
}

}  // extern "C"
