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
               long long* sweeps, int algo, float delta0, float wfin, long long* rounds_out) {
  long long nrounds = 0, npolish = 0;
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
  float c_prev = 0.f, c_prev2 = 0.f;
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
    const int mm_in = mm;
    float croot = 0;
    bool found = false, lstop = false;
    for (int attempt = 0; attempt < 2; ++attempt) {
    mm = mm_in; found = false; lstop = false; croot = 0;
    const bool coarse = !(algo == 0 || k < 2 || attempt == 1);
    float lo = 0, hi = 0, dlo = 0, dhi = 0;
    float xn = 0, yn = 0; bool have_n = false;   // a third scan point next to the bracket
    bool done = false;
    int mnew = mm;
    auto stopc = [&](float c, int m) { return (c < 0.8f * b_top) || !(c < q1[m - 1].y + 0.3f) || !(c == c); };
    {
      // coarse-to-fine scan: coarse points every S grid steps, then the S-1 interior points of the first
      // coarse interval that shows an event.  S = 1 reproduces the plain scan.
      const float c_1 = SD_ADD(c1, dc);
      int S = (!coarse || c_1 < 0.8f * b_top) ? 1 : G;
      float cbase = c1;             // grid value of the first coarse point of this round
      float cP = 0, dP = 0; int have_prev = 0;
      for (int round = 0; round < 4096 && !done; ++round) {
        nrounds++;
        for (int g = 0; g < G; ++g) {
          float c = cbase;
          for (int t = 0; t < g * S; ++t) c = SD_ADD(c, dc);
          cj[g] = c; mj[g] = layer_drop(c, T, fact, n, q1.data()); dj[g] = sweep(c, T, mj[g]);
        }
        int jev = -1;
        for (int g = 0; g < G; ++g) {
          const bool hasp = (g > 0) || have_prev;
          const float dp = g ? dj[g - 1] : dP;
          const bool change = hasp && (std::signbit(dp) != std::signbit(dj[g]));
          const bool stop = hasp && !change && stopc(cj[g], mj[g]);
          // above the half-space velocity the secular function can change sign twice inside one coarse
          // interval (kink at c = b(mmax)): force such intervals to be resolved point by point
          const bool risky = hasp && S > 1 && !(cj[g] < q1[mj[g] - 1].y);
          if (change || stop || risky) { jev = g; break; }
        }
        if (jev < 0) { cP = cj[G - 1]; dP = dj[G - 1]; have_prev = 1; cbase = cP; for (int t = 0; t < S; ++t) cbase = SD_ADD(cbase, dc); continue; }
        // sequence of S+1 fine points from the previous coarse point to the event point
        std::vector<float> sc(S + 1), sd(S + 1); std::vector<int> sm(S + 1);
        sc[0] = jev ? cj[jev - 1] : cP; sd[0] = jev ? dj[jev - 1] : dP; sm[0] = 0;
        sc[S] = cj[jev]; sd[S] = dj[jev]; sm[S] = mj[jev];
        if (S > 1) {
          nrounds++;
          for (int q = 1; q < S; ++q) {
            float c = sc[0];
            for (int t = 0; t < q; ++t) c = SD_ADD(c, dc);
            sc[q] = c; sm[q] = layer_drop(c, T, fact, n, q1.data()); sd[q] = sweep(c, T, sm[q]);
          }
        }
        for (int q = 1; q <= S; ++q) {
          const bool change = std::signbit(sd[q - 1]) != std::signbit(sd[q]);
          const bool stop = !change && stopc(sc[q], sm[q]);
          if (change || stop) {
            found = change; lo = sc[q - 1]; hi = sc[q]; dlo = sd[q - 1]; dhi = sd[q]; mnew = sm[q]; done = true;
            // neighbour: prefer the side whose |value| is smaller (closer to the root)
            have_n = false;
            if (q >= 2) { xn = sc[q - 2]; yn = sd[q - 2]; have_n = true; }
            if (q < S && (!have_n || fabsf(sd[q]) < fabsf(sd[q - 1]))) { xn = sc[q + 1]; yn = sd[q + 1]; have_n = true; }
            break;
          }
        }
        if (!done) {  // only the 'risky' flag fired: go on from the end of this interval with the plain scan
          S = 1; cP = sc[sc.size() - 1]; dP = sd[sd.size() - 1]; have_prev = 1; cbase = SD_ADD(cP, dc);
        }
      }
    }
    mm = mnew;
    if (found) {
      const float lo0 = lo, hi0 = hi, dlo0 = dlo, dhi0 = dhi;
      bool multi = false;
      const float b_hs = q1[mm - 1].y;
      const bool careful = !coarse || (b_hs > lo - 0.011f && b_hs < hi + 0.011f);
      for (int it = 0; it < 16 && (hi - lo) > wfin; ++it) {
        nrounds++; npolish++;
        const float w = hi - lo;
        if (careful || it >= 4) {
          const float step = w / (float)(G + 1);
          for (int g = 0; g < G; ++g) cj[g] = lo + (float)(g + 1) * step;
        } else {
          // points clustered geometrically around the secant estimate
          const float den = dhi - dlo;
          float e = (den != 0.f) ? lo - dlo * w / den : 0.5f * (lo + hi);
          if (it == 0 && have_n && algo >= 2) {
            // inverse quadratic interpolation through (lo, hi, neighbour) evaluated at y = 0
            const float y0 = dlo, y1 = dhi, y2 = yn;
            const float d01 = y0 - y1, d02 = y0 - y2, d12 = y1 - y2;
            if (d01 != 0.f && d02 != 0.f && d12 != 0.f) {
              const float q = lo * (y1 * y2) / (d01 * d02) - hi * (y0 * y2) / (d01 * d12) + xn * (y0 * y1) / (d02 * d12);
              if (q > lo && q < hi) e = q;
            }
          }
          const float dl = fmaxf(delta0, w * (1.0f / 2048.f));
          for (int g = 0; g < G; ++g) {
            const int h = g - G / 2;                       // -4..3 for G = 8
            const float mag = (h >= 0) ? (float)(1 << (2 * h)) : -(float)(1 << (2 * (-h - 1)));
            float pnt = e + mag * dl;
            const float eps = w * 1.0e-3f;
            pnt = fminf(fmaxf(pnt, lo + eps), hi - eps);
            cj[g] = pnt;
          }
        }
        for (int g = 0; g < G; ++g) dj[g] = sweep(cj[g], T, mm);
        int j = -1, nchg = 0;
        for (int g = 0; g < G; ++g) {
          const float dp = g ? dj[g - 1] : dlo;
          if (std::signbit(dp) != std::signbit(dj[g])) { if (j < 0) j = g; nchg++; }
        }
        if (std::signbit(dj[G - 1]) != std::signbit(dhi)) nchg++;
        if (careful && it == 0 && nchg > 1) { multi = true; break; }
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
        if (!nevill_seq(f, lo0, hi0, dlo0, dhi0, croot, ev_n)) { found = false; lstop = true; }
      }
      if (found && croot > q1[mm - 1].y) found = false;
    }
    if (!coarse) break;
    bool suspicious = !found || lstop;
    if (!suspicious) {
      const float r = (lt[k - 1] - lt[k]) / (lt[k - 2] - lt[k - 1]);
      const float stepp = (c_prev - c_prev2) * r;
      suspicious = !(croot - (c_prev + stepp) <= fmaxf(0.1f, fabsf(stepp)));
    }
    if (!suspicious) break;
    }  // attempt
    if (lstop) { nfound = 0; break; }
    if (!found) break;
    float ratio = 0;
    if (kind == 2) {
      const float r12 = rayleigh_sweep(croot, T, mm, q0.data(), q1.data(), 2);
      const float r3 = rayleigh_sweep(croot, T, mm, q0.data(), q1.data(), 3);
      nsw += 2; nrounds++;
      ratio = 0.5f * r3 / r12;
    }
    c_out[k] = croot; ratio_out[k] = ratio; c_prev2 = c_prev; c_prev = croot; nfound = k + 1;
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
  if (rounds_out) { rounds_out[0] += nrounds; rounds_out[1] += npolish; }
  return nfound;
}

}  // extern "C"
