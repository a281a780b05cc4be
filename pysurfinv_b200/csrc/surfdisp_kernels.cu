// surfdisp_kernels.cu -- sm_100a kernels + C ABI of the batched dispersion forward solver.
//
// Data flow of one surfdisp_batch() call (all on one stream):
//   prep_kernel    warp per model, lanes over layers   layers[5][M][Lmax] -> consts[M][8][lpad]  (flat1.f, once per model)
//   phase1_kernel  4 lanes x 2 packed trial velocities per model, 8 models per warp; every group is a state
//                  machine and one loop iteration = one sweep of the secular function for every group.
//                  Launch 1 (P1_FIRST): first period (the reference's scan, calcul.f:155-167).  Later periods
//                  (cluster of trial velocities around the extrapolated root, inverse interpolation, scan as
//                  fall-back): one general launch (P1_GENERAL), or -- large batches -- a fast-path launch without
//                  scan code (P1_FAST) that hands the models needing a scan over to the general one
//                                                       -> c[M][K], ratio[M][K], nfound[M], flags[M]
//   phase2_kernel  thread per (model, period)          energy integrals -> U[M][K]  (calcul.f:224-404,
//                  REIGEN surfa.f:714-1190 with a float32 state as packed pairs (default) or the reference's float64
//                  state (opts.group_f64) / LEIGEN surfa.f:374-606 in float32)
//   partials_kernel, misfit_kernel and, in surfdisp_mc.cuh, build_stacks / check_priors / mc_propose[_build] /
//   mc_finish / mc_accept kernels: the callers on either side of the solver (model assembly, Monte-Carlo step),
//   warp per model.
// No tensor cores: the work is a serial chain of tiny structured propagator products with
// data-dependent branches.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "surfdisp_core.cuh"
#include "../../include/surfdisp_b200.h"

namespace {

using namespace sd;

constexpr int kMaxPer = SURFDISP_MAX_PERIODS;
constexpr int kHdrBytes = 256;  // workspace header: counters + work queues

struct PeriodTab {
  float per[kMaxPer];
  float lt[kMaxPer];  // ln(t_base / T), host libm (same logf the reference calls)
};

struct WsLayout {
  size_t consts_off, ratio_off, mm_off, order_off, dtot_off, mmh_off, defer_off, dstate_off, bucket_off, total;
  int lpad;
};
constexpr int kOrderBuckets = 1024;   // models are handed out in the order of their top shear velocity (bucket sort)

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

WsLayout ws_layout(int M, int lmax, int K) {
  WsLayout w;
  w.lpad = round_up(lmax, 4);
  w.consts_off = kHdrBytes;
  size_t consts = (size_t)M * NCONST * w.lpad * sizeof(float);
  w.ratio_off = w.consts_off + ((consts + 255) / 256) * 256;
  size_t ratio = (size_t)M * K * sizeof(float);
  w.mm_off = w.ratio_off + ((ratio + 255) / 256) * 256;
  w.order_off = w.mm_off + (((size_t)M * sizeof(int) + 255) / 256) * 256;
  w.dtot_off = w.order_off + (((size_t)M * sizeof(int) + 255) / 256) * 256;
  w.mmh_off = w.dtot_off + (((size_t)M * sizeof(float) + 255) / 256) * 256;
  w.defer_off = w.mmh_off + (((size_t)M * K * sizeof(short) + 255) / 256) * 256;
  w.dstate_off = w.defer_off + (((size_t)M * sizeof(int) + 255) / 256) * 256;
  w.bucket_off = w.dstate_off + (((size_t)M * sizeof(float2) + 255) / 256) * 256;
  w.total = w.bucket_off + 2 * kOrderBuckets * sizeof(int);
  return w;
}

thread_local char g_cuda_err[256] = "";
thread_local cudaEvent_t* g_prof_events = nullptr;  // set by surfdisp_batch_profiled: 4 events around the 3 kernels

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
  return SURFDISP_ECUDA;
}
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, #call); } while (0)

// ------------------------------------------------------------------------------------ prep kernel
// One warp per model, lanes over layers (coalesced loads of the five input rows, coalesced stores of the eight
// constant rows).  Same arithmetic, operation by operation, as prep_model() in surfdisp_core.cuh (the form the
// host mirror runs): the only sequential part of flat1.f:33-37 is the float32 running sum of the thicknesses,
// done here with a 32-step shuffle loop per chunk of 32 layers.
__global__ void __launch_bounds__(128) prep_kernel(int M, int lmax, int lpad, int kind, int flatten,
                                                   const int* __restrict__ nlay,
                                                   const float* __restrict__ layers, size_t comp_stride,
                                                   float* __restrict__ consts, float* __restrict__ dtot) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const int n = nlay[m];
  if (n < 2 || n > lmax) { if (lane == 0) dtot[m] = -1.f; return; }
  double dsum = 0.0;    // thickness of the flattened stack above the half-space (see below)
  const size_t pl = comp_stride;   // distance between the five input rows (the whole batch, not the launched range)
  const float* a = layers + 0 * pl + (size_t)m * lmax;
  const float* b = layers + 1 * pl + (size_t)m * lmax;
  const float* rho = layers + 2 * pl + (size_t)m * lmax;
  const float* d = layers + 3 * pl + (size_t)m * lmax;
  const float* qs = layers + 4 * pl + (size_t)m * lmax;
  float* out = consts + (size_t)m * NCONST * lpad;
  const int ld = lpad;
  const float A = SD_R0;
  const float pwr = (kind == 1) ? 5.0f : 2.2750f;
  const float apw = sd_powf_cr(A, pwr);
  float hs = 0.f;       // running sum of thicknesses (uniform across the warp)
  float zcarry = 0.f;   // z1 of the last layer of the previous chunk
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool act = i < n;
    const float di = act ? d[i] : 0.f;
    float my_ht = 0.f, my_hs = 0.f;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const float dj = __shfl_sync(0xffffffffu, di, j);
      const float ht = hs;
      hs = SD_ADD(hs, dj);
      if (lane == j) { my_ht = ht; my_hs = hs; }
    }
    float z1 = 0.f;
    float o_dif = 1.f, o_rhofl = 0.f, o_dfl = 0.f, o_hsf = 1.f, o_rhohs = 0.f;
    const float ai = act ? a[i] : 0.f, bi = act ? b[i] : 0.f, ri = act ? rho[i] : 0.f, qi = act ? qs[i] : 0.f;
    if (act && flatten) {
      const float r_i = SD_SUB(A, my_ht);
      const float r_n = SD_SUB(A, my_hs);  // radius of the top of layer i+1
      const float fltd = sd_logf_cr(SD_DIV(r_i, r_n));
      o_dif = SD_DIV(SD_MUL(SD_SUB(SD_DIV(1.0f, r_n), SD_DIV(1.0f, r_i)), A), fltd);
      const float difr = SD_SUB(sd_powf_cr(r_i, pwr), sd_powf_cr(r_n, pwr));
      const float qqq = SD_DIV(difr, SD_MUL(SD_MUL(fltd, apw), pwr));
      o_rhofl = SD_MUL(ri, qqq);
      z1 = SD_MUL(A, sd_logf_cr(SD_DIV(A, r_n)));
      const float fct = SD_DIV(A, r_i);
      o_hsf = fct;
      o_rhohs = SD_MUL(ri, sd_powf_cr(SD_DIV(1.0f, fct), pwr));
    }
    float z0 = __shfl_up_sync(0xffffffffu, z1, 1);
    if (lane == 0) z0 = zcarry;
    zcarry = __shfl_sync(0xffffffffu, z1, 31);
    if (act) {
      if (flatten) o_dfl = SD_SUB(z1, z0);
      else { o_rhofl = ri; o_dfl = di; o_rhohs = ri; }
      out[C_AREF * ld + i] = ai; out[C_BREF * ld + i] = bi; out[C_QS * ld + i] = qi;
      out[C_DIF * ld + i] = o_dif; out[C_RHOFL * ld + i] = o_rhofl; out[C_DFL * ld + i] = o_dfl;
      out[C_HSF * ld + i] = o_hsf; out[C_RHOHS * ld + i] = o_rhohs;
      if (i < n - 1) dsum += (double)fmaxf(o_dfl, 0.f);
    }
  }
  // Upper bound of every running thickness sum the layer-dropping walks can form (surfa.f:92-106, 854-866: float32
  // additions of a subset of these thicknesses in layer order; each addition rounds by at most 2^-24, and sub-layer
  // thicknesses RN(d / ndiv) add up to d (1 + 2^-24) at most).  A walk whose limit fact*c*T is not below it drops
  // nothing: both phases skip the walk then.  -1 = unknown (NaN thickness).
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  if (lane == 0) dtot[m] = (dsum == dsum) ? __double2float_ru(dsum * 1.0001) : -1.f;
}

// ------------------------------------------------------------------------------------ phase 1
struct P1Params {
  int kind, M, lpad, K, lmax;
  const int* nlay;
  const float* consts;
  float* c_out;
  float* ratio_out;
  int* nfound;
  int* flags;
  unsigned long long* counters;
  unsigned int* queue;
  float dc, fact;
  int atten, stale, exact_scan;
  int k_begin, k_end;   // periods [k_begin, k_end) are done by this launch (the first period runs as a launch of its own)
  int* mm_state;        // layer-dropping depth carried from launch to launch
  const float* dtot;    // per model: upper bound of the thickness sums of the layer-dropping walk (prep_kernel), -1 = unknown
  // ---- hand-over from the fast-path instantiation to the general one (see phase1_kernel)
  short* mmh;           // [M][K] dropping depth left by every finished period: what a resumed model replays its refreshes with
  int* defer_list;      // models the fast-path launch handed over; their count
  unsigned int* ndefer;
  float2* dstate;       // per handed-over model: (last prediction error, periods of good predictions still required)
  int resume;           // 1: this launch continues the models of `order` (ndefer of them) at the period they stopped at
  const float* hint;    // [M][K] neighbour curves (phase velocities of a nearby model on the same periods) or nullptr
  const int* order;     // order[i] - order_base = i-th model to hand out (nullptr: index order)
  int order_base;
  int mstride;  // float4 units between consecutive groups' shared-memory records
  PeriodTab tab;
};

// One copy of the secular-function loop for all call sites (cluster / window, refinement, scan, polish,
// ellipticity): inlining it at every site made the kernel ~95 KB of SASS and the warps stalled on instruction
// fetch.  Every lane carries a PAIR of trial velocities (packed FP32 arithmetic, see surfdisp_core.cuh);
// Rayleigh sweeps are done in the adjoint form, which yields the two ellipticity minors along with the
// dispersion function.
struct Sec2 { float2 d, e2, e3; };

__device__ __noinline__ Sec2 secular2(int kind, float2 c, float T, int mm, const float4* rec, int ell_only) {
  V2 d, e2 = v2(0.f, 0.f), e3 = e2;
  const V2 cp = v2(c.x, c.y);
  if (kind == 2) d = rayleigh_adjoint2(cp, T, mm, rec, ell_only != 0, e2, e3);
  else d = love_sweep2(cp, T, mm, rec, e2, e3);
  Sec2 r;
  r.d = make_float2(vx(d), vy(d)); r.e2 = make_float2(vx(e2), vy(e2)); r.e3 = make_float2(vx(e3), vy(e3));
  return r;
}

struct SecFn;
template <int G>
__device__ __forceinline__ float gshfl(unsigned mask, float v, int src) { return __shfl_sync(mask, v, src, G); }
template <int G>
__device__ __forceinline__ int gshfl(unsigned mask, int v, int src) { return __shfl_sync(mask, v, src, G); }

struct SamplePt { float c, d, e2, e3; };

constexpr float kInterpTol = 1.0e-5f;   // agreement of the 4- and 6-point root estimates that ends the refinement
constexpr float kClusterTol = 2.0e-6f;  // same, for the first round (cluster around the predicted root)
constexpr float kClusterWmax = 6.0e-3f; // widest bracket of a first-round cluster that may be accepted (the cluster's spacing never exceeds it;
                                        // window rounds on the 0.01 grid never are)
constexpr float kBracketTol = 2.0e-5f;  // bracket width below which a secant step is final
constexpr float kClusterH0 = 2.5e-4f;   // smallest innermost spacing of the first-round cluster
constexpr float kMaxPredStep = 0.15f;   // largest change of the root from one period to the next that the extrapolation is trusted with

// bit i of the result = bit i/2 of a (i even) or of b (i odd)
template <int G>
__device__ __forceinline__ unsigned interleave(unsigned a, unsigned b) {
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < G; ++i) r |= (((a >> i) & 1u) << (2 * i)) | (((b >> i) & 1u) << (2 * i + 1));
  return r;
}

// Layer dropping (surfa.f:92-106) for ONE trial velocity, computed by the G lanes of a group together.  Each lane
// owns a contiguous chunk of layers, cut into four sub-chunks whose thickness sums (layers with c < b only) are
// accumulated side by side -- four independent loads and additions per step instead of one dependent chain.  The
// chunk sums are prefixed across the lanes; the lane whose chunk crosses dmax = fact c T picks the sub-chunk from its
// partial sums and walks only that one, in the reference's order, to find the layer.  The additions inside a
// sub-chunk are in the reference's order; the total rounds differently from the reference's single running sum
// only in the last ulp, i.e. the result can differ by one layer on an exact tie (1e-5 per call).
template <int G>
__device__ __noinline__ int layer_drop_coop(float c, float T, float fact, int nmax, const float4* rec, unsigned gmask, int gl) {
  const float dmax = SD_MUL(SD_MUL(fact, c), T);
  const int Q = (nmax + 4 * G - 1) / (4 * G);      // layers per sub-chunk
  const int i0 = gl * 4 * Q;
  float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
  for (int t = 0; t < Q; ++t) {
    const int j0 = i0 + t, j1 = j0 + Q, j2 = j1 + Q, j3 = j2 + Q;
    const float4 e0 = rec[min(j0, nmax - 1)], e1 = rec[min(j1, nmax - 1)], e2 = rec[min(j2, nmax - 1)], e3 = rec[min(j3, nmax - 1)];
    if (j0 < nmax && c < e0.y) u0 = SD_ADD(u0, e0.w);
    if (j1 < nmax && c < e1.y) u1 = SD_ADD(u1, e1.w);
    if (j2 < nmax && c < e2.y) u2 = SD_ADD(u2, e2.w);
    if (j3 < nmax && c < e3.y) u3 = SD_ADD(u3, e3.w);
  }
  const float p1 = SD_ADD(u0, u1), p2 = SD_ADD(p1, u2), s = SD_ADD(p2, u3);
  // exclusive prefix over the lanes of the group
  float base = 0.f;
#pragma unroll
  for (int j = 0; j < G - 1; ++j) {
    const float sj = __shfl_sync(gmask, s, j, G);
    if (gl > j) base = SD_ADD(base, sj);
  }
  int found = nmax;
  if (!(base > dmax) && SD_ADD(base, s) > dmax) {
    // first sub-chunk whose end is beyond dmax
    int q = 3; float sum = SD_ADD(base, p2);
    if (SD_ADD(base, u0) > dmax) { q = 0; sum = base; }
    else if (SD_ADD(base, p1) > dmax) { q = 1; sum = SD_ADD(base, u0); }
    else if (SD_ADD(base, p2) > dmax) { q = 2; sum = SD_ADD(base, p1); }
    const int b0 = i0 + q * Q, b1 = min(b0 + Q, nmax);
    found = b1;      // (rounding of the split sums: the crossing is at the end of the sub-chunk at the latest)
    for (int i = b0; i < b1; ++i) {
      const float4 e = rec[i];
      if (c < e.y) { sum = SD_ADD(sum, e.w); if (sum > dmax) { found = i + 1; break; } }
    }
  }
  // the first lane that crosses wins
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(gmask, found, o, G));
  return found < 2 ? 2 : found;
}

struct SecFn {
  int kind; float T; int mm; const float4* rec;
  __device__ float operator()(float cc) const;
};

__device__ float SecFn::operator()(float cc) const { return secular2(kind, make_float2(cc, cc), T, mm, rec, 0).d.x; }

// the reference's sequential polish (rare path), kept out of line
__device__ __noinline__ bool nevill_out_of_line(const SecFn& f, float c1, float c2, float d1, float d2, float* cc, int* evals) {
  return nevill_seq(f, c1, c2, d1, d2, *cc, *evals);
}

#ifndef P1_G
#define P1_G 4
#endif
#ifndef P1_SOFT_SYNC
#define P1_SOFT_SYNC 4     // 0 = every group pulls its next model as soon as it is done
#endif
#ifndef P1_MINBLK
#define P1_MINBLK 4
#endif
#ifndef P1_THREADS
#define P1_THREADS 128
#endif

// Stages of a group's (= one model's) state machine.  Every iteration of the kernel's main loop performs ONE
// sweep of the secular function for every group of the warp, whatever stage each group is in: the groups
// advance through their periods and models independently, and the expensive code (the sweep) is always
// executed by all of them together (a structured loop nest would make groups that need an extra round, or
// that sit in the scan of their first period, serialise the others).
enum { ST_FETCH = 0, ST_PERIOD, ST_FAST, ST_REFINE, ST_SCAN, ST_POLISH, ST_ELL, ST_DONE, ST_DEFER };
enum { P1_GENERAL = 0, P1_FIRST = 1, P1_FAST = 2 };
// Which trial velocities the next sweep of a group needs (built in ONE place, right before the sweep)
enum { NB_NONE = 0, NB_FAST, NB_REFINE, NB_SCAN, NB_POLISH, NB_ELL };

// FIRST: the launch that does one period per model from scratch (k_begin = 0, k_end = 1): every model scans, the
// cluster / window path of the later periods is compiled out -- a smaller loop body for the launch whose short
// sweeps make it the most sensitive to instruction fetch.
template <int G, int MODE>
__global__ void __launch_bounds__(P1_THREADS, P1_MINBLK) phase1_kernel(const __grid_constant__ P1Params p) {
  static_assert(G == 4 || G == 8, "4 or 8 lanes x 2 trial velocities per model");
  constexpr int P = 2 * G;   // trial velocities per round; point i lives in lane i/2, component i%2
  constexpr bool FIRST = (MODE == P1_FIRST);
  constexpr bool NOSCAN = (MODE == P1_FAST);
  extern __shared__ float4 smem[];
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = ((1u << G) - 1u) << gbase;
  const unsigned gbits = (1u << G) - 1u;
  const int grp = threadIdx.x / G;
  float4* rec = smem + (size_t)grp * p.mstride;   // this model's layer records (a, b, rho, d)
  // ordered samples of the current round: slot 0 / P+1 = bracket ends of the previous round, 1..P = the points
  float4* slots = smem + (size_t)(blockDim.x / G) * p.mstride + (size_t)grp * (P + 2);
  const int K = p.K;
  const int ld = p.lpad;
  unsigned long long my_steps = 0, my_sweeps = 0;
  int my_models = 0;

  // ordered sign-change mask of the P points of a round: bit i set <=> sign(point i) != sign(point i-1)
  // (point -1 = `before`)
  auto change_mask = [&](float2 d, float before) -> unsigned {
    float dprev = __shfl_up_sync(gmask, d.y, 1, G);
    if (gl == 0) dprev = before;
    const unsigned ex = (__ballot_sync(gmask, signbit(dprev) != signbit(d.x)) >> gbase) & gbits;
    const unsigned ey = (__ballot_sync(gmask, signbit(d.x) != signbit(d.y)) >> gbase) & gbits;
    return interleave<G>(ex, ey);
  };
  auto pair_mask = [&](bool bx, bool by) -> unsigned {
    const unsigned ex = (__ballot_sync(gmask, bx) >> gbase) & gbits;
    const unsigned ey = (__ballot_sync(gmask, by) >> gbase) & gbits;
    return interleave<G>(ex, ey);
  };

  // ---- per-model state (uniform inside a group)
  int stage = ST_FETCH;
  int model = 0, n = 2, k = 0, mm = 2, nfound = 0, flag = 0;
  const float* cst = p.consts;
  float* crow = p.c_out;
  float* rrow = p.ratio_out;
  float c1 = 1.f, c_prev = 0.f, c_prev2 = 0.f, c_prev3 = 0.f, pred_err = 1.0e-3f, c_pred = 0.f, b_top = 0.f, T = 1.f;
  int hopped = 0;            // > 0: the root left the extrapolation of its branch; periods of good predictions still required
  int kfirst = 0;            // the first period of the current model in this launch
  bool force_scan = false;   // resumed model: its first period here goes to the scan at once
  bool mid_liquid = false;
  // ---- the sweep request of the current iteration (per lane) and its result
  float2 pc = make_float2(1.f, 1.f), pd = make_float2(0.f, 0.f), pe2 = pd, pe3 = pd;
  int meval = 2, ell_only = 0;
  // ---- fast path (cluster / window / refinement rounds)
  int j0 = 2, w0 = 2, dir = 0, fstage = 0, wtry = 0, mw = 2, jb = 1, it = 0, nvalid = 0;
  bool has_ends = false;
  SamplePt E0 = {0.f, 0.f, 0.f, 0.f}, E1 = E0;
  // ---- point-by-point path (scan, polish)
  float cbase = 0.f, cP = 0.f, dP = 0.f, tP = 0.f, cP2 = 0.f, dP2 = 0.f, tP2 = 0.f, lo = 0.f, hi = 0.f, dlo = 0.f, dhi = 0.f, lo0 = 0.f, hi0 = 0.f, dlo0 = 0.f, dhi0 = 0.f;
  int mjx = 2, mjy = 2, round = 0, pit = 0, stride = 1;
  bool have_prev = false, own_mj = false;
  bool own_eval = false;    // this scan round evaluates every point on its own truncation (see build_scan)
  bool own_half = false;    // ... and its lower points are done
  bool from_scan = false;   // the interpolation rounds refine a bracket found by the scan (fall-back: uniform-section polish)
  float bmin = 0.f;   // smallest b below the top layer (this period's records)
  // ---- result of the period
  float croot = 0.f, ratio = 0.f;
  // ---- what to build before the next sweep; refinement round parameters
  int need = NB_NONE;
  float rf_e = 0.f, rf_s0 = 0.f;
  bool rf_uni = false;

  const float cspan = (float)(1 << ((P - 2) / 2 - 1)) - 0.5f;   // outermost cluster offset in units of the spacing

  // layer dropping for one trial velocity by the lanes of the group; nothing is dropped -- and the walk skipped --
  // when even the whole stack is thinner than the limit fact*c*T (most periods beyond the first few)
  auto drop_coop = [&](float c) -> int {
    const float dt = rec[ld].x;
    if (dt >= 0.f && dt <= SD_MUL(SD_MUL(p.fact, c), T)) return n;
    return layer_drop_coop<G>(c, T, p.fact, n, rec, gmask, gl);
  };

  // ---- request builders
  auto build_fast = [&]() {
    // stages 0/1: point 0 = c1 itself, point 1 = half way to the window, points 2..P-1 = cluster / window; an odd
    // number of roots below shows as a sign difference between points 0, 1 and 2.  stage 2: P window points.
    const float hc = fminf(fmaxf(6.0f * pred_err / cspan, kClusterH0), 16.f * kClusterH0);  // cluster spans ~6x the last prediction error
    float cc[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pi = 2 * gl + h;
      if (pi >= w0 && fstage == 0) cc[h] = c_pred + geometric_offset(pi - 2, P - 2) * hc;
      else {
        // grid points in closed form: only the signs matter there (the reference's sequentially accumulated
        // grid differs by a few ulps, i.e. 1e-4 of a grid step)
        const int idx = (pi < w0) ? ((pi == 0) ? 0 : j0 / 2) : j0 + (pi - w0);
        cc[h] = c1 + (float)idx * p.dc;
      }
    }
    pc = make_float2(cc[0], cc[1]);
    mw = drop_coop(gshfl<G>(gmask, pc.y, G - 1));
    meval = mw; ell_only = 0;
  };
  auto build_scan = [&]() {
    // 2G consecutive grid points, accumulated like the reference does (calcul.f:157).  The reference gives each
    // point its own layer dropping (calcul.f:155-159, surfa.f:92-106).  As long as the round's top
    // point is below the half-space velocity of the deepest truncation, all truncations have the same sign there
    // and the whole round is evaluated on the deepest one.  When it is not -- the walk ran through the whole stack
    // and the true half-space is slower than the trial velocities (velocity inversion at depth) -- points whose own
    // truncation is shallower do not have the same sign: every point is then evaluated on its own truncation, like
    // the reference does (two sweeps for the lane; exact_scan: whenever the truncations of a round differ).
    // The own truncation of a point also matters for the stop test c >= b(mmax) + 0.3 (calcul.f:166), which
    // cannot fire while c < min b + 0.3.
    float cx = cbase;
    for (int t = 0; t < 2 * gl * stride; ++t) cx = SD_ADD(cx, p.dc);
    float cy = cx;
    for (int t = 0; t < stride; ++t) cy = SD_ADD(cy, p.dc);
    pc = make_float2(cx, cy);
    const float ctop = gshfl<G>(gmask, pc.y, G - 1);
    meval = drop_coop(ctop);
    own_mj = !(ctop < bmin + 0.3f);
    const bool above_hs = !(ctop < rec[meval - 1].y) || p.exact_scan;
    if (own_mj || above_hs) {
      mjx = layer_drop(pc.x, T, p.fact, n, rec); mjy = layer_drop(pc.y, T, p.fact, n, rec);
      own_mj = true;
      if (above_hs) {
        const int mtop = gshfl<G>(gmask, mjy, G - 1);
        own_eval = (__ballot_sync(gmask, mjx != mtop || mjy != mtop) & gmask) != 0u;
      }
    }
    ell_only = 0;
  };
  auto start_scan = [&]() {
    if constexpr (NOSCAN) { stage = ST_DEFER; need = NB_NONE; return; }
    cbase = c1; cP = 0.f; dP = 0.f; have_prev = false; round = 0; stride = 1;
    stage = ST_SCAN; need = NB_SCAN;
  };
  // the interpolation rounds gave up: the cluster / window rounds restart as a scan; a bracket that came from
  // the scan is polished by uniform section instead (mmax is pinned already)
  auto interp_failed = [&]() {
    if constexpr (NOSCAN) { stage = ST_DEFER; need = NB_NONE; return; }
    if (from_scan) { from_scan = false; lo0 = lo; hi0 = hi; dlo0 = dlo; dhi0 = dhi; pit = 0; stage = ST_POLISH; need = NB_POLISH; }
    else start_scan();
  };
  auto build_polish = [&]() {
    const float st = (hi - lo) / (float)(P + 1);
    pc = make_float2(lo + (float)(2 * gl + 1) * st, lo + (float)(2 * gl + 2) * st);
    meval = mm; ell_only = 0;
  };
  auto build_refine = [&]() {
    // P points around the root estimate, spaced by the disagreement of the two interpolation orders
    float cc[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pi = 2 * gl + h;
      cc[h] = E0.c + (rf_uni ? (float)(pi + 1) * ((E1.c - E0.c) / (float)(P + 1)) : rf_e + geometric_offset(pi, P) * rf_s0);
    }
    pc = make_float2(cc[0], cc[1]);
    meval = mw; ell_only = 0;
  };
  auto build_ell = [&]() {
    pc = make_float2(croot, croot);
    meval = mm; ell_only = 1;
  };

  // refresh layers 0..mref-1 for the period with log term lt, layer mref-1 flattened as the half-space
  // (calcul.f:112-133); returns the smallest b below the top layer among them
  auto refresh = [&](int mref, float lt) -> float {
    __syncwarp(gmask);
    float bm = 3.0e38f;
    for (int i0 = 4 * gl; i0 < mref; i0 += 4 * G) {
      // four layers per lane and iteration: six float4 loads of the per-model constants
      const float4 AR = *reinterpret_cast<const float4*>(cst + C_AREF * ld + i0);
      const float4 BR = *reinterpret_cast<const float4*>(cst + C_BREF * ld + i0);
      const float4 QS = *reinterpret_cast<const float4*>(cst + C_QS * ld + i0);
      const float4 DF = *reinterpret_cast<const float4*>(cst + C_DIF * ld + i0);
      const float4 RF = *reinterpret_cast<const float4*>(cst + C_RHOFL * ld + i0);
      const float4 DL = *reinterpret_cast<const float4*>(cst + C_DFL * ld + i0);
      const float ar[4] = {AR.x, AR.y, AR.z, AR.w}, br[4] = {BR.x, BR.y, BR.z, BR.w}, qs[4] = {QS.x, QS.y, QS.z, QS.w};
      const float df[4] = {DF.x, DF.y, DF.z, DF.w}, rf[4] = {RF.x, RF.y, RF.z, RF.w}, dl[4] = {DL.x, DL.y, DL.z, DL.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j;
        if (i < mref) {
          const bool hs = (i == mref - 1);
          float a, b;
          layer_ab_vals(ar[j], br[j], qs[j], hs ? cst[C_HSF * ld + i] : df[j], lt, p.atten, a, b);
          rec[i] = make_rec(a, b, hs ? cst[C_RHOHS * ld + i] : rf[j], hs ? 0.f : dl[j]);
          if (i >= 1) bm = fminf(bm, b);
        }
      }
    }
    __syncwarp(gmask);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) bm = fminf(bm, __shfl_xor_sync(gmask, bm, o, G));
    return bm;
  };

  for (;;) {
    // =========================================================== set-up: next model / next period
    // The groups of a warp start their models together: inside a model they drift apart by a few rounds only,
    // so that the sweeps issued together have similar depths (the truncation depth grows with the period; a
    // warp whose groups sit at very different periods would run every sweep to the depth of the deepest one).
#ifdef P1_SYNC_MODELS
    const bool warp_fetch = __all_sync(0xffffffffu, stage == ST_FETCH || stage == ST_DONE);
#elif P1_SOFT_SYNC > 0
    // soft re-alignment at model boundaries: a group that is done waits for the groups of its warp that are
    // within P1_SOFT_SYNC periods of finishing (not for stragglers, not for models that are being scanned every
    // period), so that the next models start together again
    const bool near_end = (stage != ST_FETCH && stage != ST_DONE) && !hopped && (p.k_end - p.k_begin > 8) &&
                          (p.k_end - k <= P1_SOFT_SYNC);
    // (a launch of one period per model -- the first-period launch -- hands out models to the whole warp at once:
    // the eight scans then pass through their stages together, 26 -> 23 ms)
    const bool first_wait = (p.k_end - p.k_begin == 1) && (stage != ST_FETCH && stage != ST_DONE);
    const bool warp_fetch = __ballot_sync(0xffffffffu, near_end || first_wait) == 0u;
#else
    const bool warp_fetch = true;
#endif
    if (stage == ST_FETCH && warp_fetch) {
      if (gl == 0) {
        model = (int)atomicAdd(p.queue, 1u);
        const int mtot = p.resume ? min((int)*p.ndefer, p.M) : p.M;
        if (model >= mtot) model = p.M;
        else if (p.order) model = p.order[model] - p.order_base;
      }
      model = gshfl<G>(gmask, model, 0);
      if (model >= p.M) stage = ST_DONE;
      else {
        n = p.nlay[model];
        crow = p.c_out + (size_t)model * K;
        rrow = p.ratio_out + (size_t)model * K;
        if (n < 2 || n > p.lmax) {
          for (int kk = gl; kk < K; kk += G) { crow[kk] = 0.f; rrow[kk] = 0.f; }
          if (gl == 0) { p.nfound[model] = 0; if (p.flags) p.flags[model] = SURFDISP_F_NO_ROOT_FIRST; }
          n = 2;  // stays in ST_FETCH: sits this iteration's sweep out and pulls another model next time
        } else {
          my_models += (gl == 0 && p.k_begin == 0);
          cst = p.consts + (size_t)model * NCONST * p.lpad;
          // (kept in the padding record of the group's layer records: see drop_coop)
          if (gl == 0) rec[ld] = make_float4(p.dtot[model], 0.f, 0.f, 0.f);
          // first start velocity, fast_surf.f:157-171
          {
            const float b0 = cst[C_BREF * ld + 0];
            const int ilay = (b0 < 0.1f) ? 1 : 0;
            float b_corr = 0.f;
            if (p.atten) b_corr = SD_DIV(SD_MUL(cst[C_QS * ld + ilay], p.tab.lt[0]), SD_PI_ATT);
            float qq = cst[C_BREF * ld + ilay];
            if (p.kind == 2) qq = SD_MUL(0.9f, qq);
            c1 = SD_MUL(qq, SD_ADD(1.0f, b_corr));
            if (b0 < 0.1f) c1 = 0.5f;
          }
          mm = n;  // reference COMMON mmax carried from period to period
          nfound = 0; flag = 0; k = p.resume ? p.nfound[model] : p.k_begin; hopped = 0;
          kfirst = k; force_scan = false;
          c_prev = c_prev2 = c_prev3 = 0.f; pred_err = 1.0e-3f;
          stage = ST_PERIOD;
          if (k > 0) {
            // continuing a model whose earlier periods were done by a previous launch
            nfound = p.nfound[model];
            if (nfound < k || k >= p.k_end) { stage = ST_FETCH; n = 2; }   // it ended there: nothing to do
            else {
              // the layers below the dropping depth keep the values of the last period that refreshed them (SURVEY
              // Q1): all of them the first period's, then every later period its own dropping depth
              bmin = refresh(n, p.tab.lt[0]) - 0.05f;
              if (p.stale) for (int kk = 1; kk < k; ++kk) refresh((int)p.mmh[(size_t)model * K + kk - 1], p.tab.lt[kk]);
              if (p.resume) {
                // handed over by the fast-path launch at the point where it would have started the scan of period k
                const float2 ds = p.dstate[model];
                pred_err = ds.x; hopped = (int)ds.y; force_scan = true;
              }
              mm = p.mm_state[model];
              flag = p.flags ? p.flags[model] : 0;
              c_prev = crow[k - 1];
              if (k >= 2) c_prev2 = crow[k - 2];
              if (k >= 3) c_prev3 = crow[k - 3];
            }
          }
        }
      }
    }
    if (stage == ST_PERIOD) {
      T = p.tab.per[k];
      const float lt = p.tab.lt[k];
      const float bm = refresh(p.stale ? mm : n, lt);
      if (k == kfirst) {
        // liquid layers below the top one: the in-sweep ellipticity is not valid for such stacks (all n records
        // are in shared memory at a model's first period here)
        bool l = false;
        for (int i = 1 + gl; i < n; i += G) l |= !(rec[i].y > 0.f);
        mid_liquid = (__ballot_sync(gmask, l) & gmask) != 0u;
      }
      // smallest b below the top layer: taken at the first period, where all layers are refreshed (mm = n); the
      // attenuation correction moves b by < 1 % over the period range, which the margin covers
      if (k == 0) bmin = bm - 0.05f;
      if (k > 0) c1 = SD_MUL(0.90f, c_prev);  // calcul.f:143
      b_top = rec[0].y;
      // ---- fast path (every period after the first, unless exact_scan).  The reference scans
      // c1, c1+dc, ... for the first sign change (calcul.f:155-167) and polishes inside that bracket
      // (NEVILL, surfa.f:2-83).  The root it ends on moves smoothly with the period, so the P = 2G trial
      // velocities of the first round are: c1 itself, a point half way, and P-2 points clustered geometrically
      // around the extrapolation (in ln T) of the previous roots.  The round is accepted when c1, the
      // half-way point and the lowest cluster point have the same sign (no odd number of roots was skipped;
      // a model whose root ever leaves the extrapolation by more than 0.1 km/s -- mode hopping -- is scanned
      // point by point from then on), there is exactly one sign change inside the cluster and the half-space
      // velocity (kink of the secular function) is not nearby.  If the cluster misses the root, a window of
      // P-2 grid points takes its place (moved up or down once more if needed); anything else goes to the
      // point-by-point path.  All points of a round are evaluated on the round's deepest truncation
      // (layer dropping, surfa.f:92-106) so that they sample ONE smooth function -- each truncation depth
      // scales the unnormalised secular function differently but has the same root to ~1e-10 -- and the root
      // is taken by inverse polynomial interpolation; if the 4- and 6-point estimates disagree, P more points
      // are clustered around the estimate and the test is repeated.
      c_pred = c_prev;
      // Neighbour curve (optional, P1Params::hint): the dispersion curve of a nearby model on the same periods -- the
      // chain's current model in a Monte-Carlo walk, whose proposal differs by one small step.  The DIFFERENCE to that
      // curve varies slowly with the period, so it is extrapolated instead of the curve itself (linear in ln T from the
      // third period on).  Only the centre of the cluster changes: the sign guards, the acceptance rules and the
      // fall-back to the scan are the same, so a wrong or missing hint costs sweeps, never the root.
      bool hinted = false;
      if (!FIRST && p.hint && k >= 1) {
        const float* hrow = p.hint + (size_t)model * K;
        const float hk = hrow[k], hk1 = hrow[k - 1];
        if (hk > 0.f && hk1 > 0.f) {
          hinted = true;
          const float d1 = c_prev - hk1;
          c_pred = hk + d1;
          if (k >= 2 && hrow[k - 2] > 0.f) {
            const float d2 = c_prev2 - hrow[k - 2];
            c_pred += (d1 - d2) * ((lt - p.tab.lt[k - 1]) / (p.tab.lt[k - 1] - p.tab.lt[k - 2]));
          }
        }
      }
      if (FIRST || hinted) { /* no prediction needed / c_pred is set */ }
      else if (k == 1) c_pred = c_prev + 0.02f;   // phase velocity grows with period: bias the first window upwards
      else if (k >= 2) {
        // extrapolation of the previous roots in ln T: linear, quadratic from the fourth period on
        const float x0 = p.tab.lt[k - 1], x1 = p.tab.lt[k - 2], x = lt;
        c_pred = c_prev + (c_prev - c_prev2) * ((x - x0) / (x0 - x1));
        if (k >= 3) {
          const float x2 = p.tab.lt[k - 3];
          const float d01 = (c_prev - c_prev2) / (x0 - x1), d12 = (c_prev2 - c_prev3) / (x1 - x2);
          c_pred += (d01 - d12) / (x0 - x2) * (x - x0) * (x - x1);
        }
      }
      j0 = (int)floorf((c_pred - c1) / p.dc) - (P - 4) / 2;
      if (j0 < 2) j0 = 2;
      // (an extrapolation that moves the root by more than 0.15 km/s is not trusted: where the branch is that steep --
      // thick slow sediments, coarse period lists -- the cluster can land on a higher mode with an even number of roots
      // between c1 and it, which the sign guard cannot see; scan from c1 like the reference)
      if (!FIRST && !p.exact_scan && k >= 1 && !hopped && !force_scan && !(SD_ADD(c1, p.dc) < 0.8f * b_top) && j0 < 1000 &&
          fabsf(c_pred - c_prev) <= kMaxPredStep) {
        // fstage 0 (from the third period on): cluster around the predicted root; 1: window of P-2 grid points
        // around it; 2: window of P grid points moved up or down
        fstage = ((k >= 2 || hinted) && j0 >= 4) ? 0 : 1;
        w0 = 2; dir = 0; wtry = 0; from_scan = false;
        stage = ST_FAST; need = NB_FAST;
      } else start_scan();
      force_scan = false;
    }
    if (__all_sync(0xffffffffu, stage == ST_DONE)) break;
    // ---- trial velocities of this iteration's sweep
    if (need != NB_NONE) { own_eval = false; own_half = false; }
    if (!FIRST && need == NB_FAST) build_fast();
    else if (need == NB_REFINE) build_refine();
    else if (!NOSCAN && need == NB_SCAN) build_scan();
    else if (!NOSCAN && need == NB_POLISH) build_polish();
    else if (need == NB_ELL) build_ell();
    need = NB_NONE;

    // =========================================================== the sweep (the only call site)
    const bool active = (stage >= ST_FAST && stage <= ST_ELL);
    bool process = active;
    if (active) {
      // (a scan round on own truncations takes two iterations: first every lane's lower point, then the upper one)
      float2 cc = pc;
      int me = meval;
      if (own_eval) { cc = own_half ? make_float2(pc.y, pc.y) : make_float2(pc.x, pc.x); me = own_half ? mjy : mjx; }
      const Sec2 sv = secular2(p.kind, cc, T, me, rec, ell_only);
      my_steps += 2u * (unsigned)(me - 1); my_sweeps += 2;
      if (!own_eval) { pd = sv.d; pe2 = sv.e2; pe3 = sv.e3; }
      else if (!own_half) { pd.x = sv.d.x; pe2.x = sv.e2.x; pe3.x = sv.e3.x; own_half = true; process = false; }
      else { pd.y = sv.d.x; pe2.y = sv.e2.x; pe3.y = sv.e3.x; own_half = false; }
    }

#ifdef P1_DEBUG
    if (active && model == P1_DEBUG && gl == 0 && k < 4)
      printf("k=%d stage=%d fstage=%d wtry=%d j0=%d w0=%d meval=%d pc=(%.5f,%.5f) pd=(%g,%g) c1=%.5f c_pred=%.5f lo=%.5f hi=%.5f cbase=%.5f round=%d\n",
             k, stage, fstage, wtry, j0, w0, meval, pc.x, pc.y, pd.x, pd.y, c1, c_pred, lo, hi, cbase, round);
#endif
    // =========================================================== what the results mean, per stage
    bool do_interp = false, period_done = false, model_done = false;
    if (process) {
      __syncwarp(gmask);
      slots[1 + 2 * gl] = make_float4(pc.x, pd.x, pe2.x, pe3.x);
      slots[2 + 2 * gl] = make_float4(pc.y, pd.y, pe2.y, pe3.y);
      __syncwarp(gmask);
    }
    auto sample = [&](int i) {
      // i-th entry of the ordered list: [point w0 .. point P-1] or [E0, point 0 .. point P-1, E1]
      const float4 v = slots[has_ends ? i : i + w0 + 1];
      SamplePt sp; sp.c = v.x; sp.d = v.y; sp.e2 = v.z; sp.e3 = v.w;
      return sp;
    };
    if (!FIRST && stage == ST_FAST) {
      const float d0 = gshfl<G>(gmask, pd.x, 0);
      const unsigned evc = change_mask(pd, d0) & ~((2u << w0) - 1u);     // changes between cluster / window points
      const unsigned evw = pair_mask(signbit(pd.x) != signbit(d0), signbit(pd.y) != signbit(d0));
      const bool below_ok = !(evw & ((2u << w0) - 1u));     // points 0..w0 have the sign of c1
      const int jev = __ffs(evc) - 1;
      bool retry = false, ok = false;
      if (below_ok && jev >= 1) ok = !(fstage == 0 && (evc & (evc - 1u)));   // not several sign changes inside the cluster
      else if (wtry < 3) {
        if (fstage == 0) { fstage = 1; retry = true; }           // the cluster does not bracket the root
        else if (below_ok && !evc && dir >= 0) { dir = 1; j0 += (w0 ? P - 3 : P - 1); w0 = 0; fstage = 2; retry = true; }   // root above the window
        else if (!below_ok && dir <= 0) {
          // root below the window: only if it is between the half-way point and the window
          const bool lower_ok = !(evw & ((1u << w0) - 1u));
          const int jmin = w0 ? j0 / 2 : 0;
          if (lower_ok && j0 > jmin) { dir = -1; j0 = max(j0 - (P - 1), jmin); w0 = 0; fstage = 2; retry = true; }
        }
      }
      if (ok) {
        // the half-space velocity of the round's truncation is a kink of the sampled function: not within a
        // grid step of the bracket, and the interpolation only uses points below it
        has_ends = false;
        const float bh2 = rec[mw - 1].y;
        jb = jev - w0;  // index (in the ordered sample list) of the upper end of the bracket
        const float br_lo = sample(jb - 1).c, br_hi = sample(jb).c;
        nvalid = __popc(pair_mask(pc.x < bh2, pc.y < bh2) & ~((1u << w0) - 1u));  // list points below the kink
        const bool kink = (bh2 > br_lo - 0.011f && bh2 < br_hi + 0.011f) || nvalid < 6 || jb > nvalid - 1;
        if (kink) start_scan(); else { it = 0; do_interp = true; }
      } else if (retry) { wtry++; need = NB_FAST; }
      else start_scan();
    } else if (stage == ST_REFINE) {
      const unsigned ev = change_mask(pd, E0.d);
      const float dlast = gshfl<G>(gmask, pd.y, G - 1);
      has_ends = true;
      if (ev) { jb = __ffs(ev); it++; do_interp = true; }                                   // point i is list entry i + 1
      else if (signbit(dlast) != signbit(E1.d)) { jb = P + 1; it++; do_interp = true; }
      else interp_failed();
    } else if (!NOSCAN && stage == ST_SCAN && process) {
      // ---- scan for the first sign change on the grid c1 + i dc (calcul.f:155-167).  The reference examines
      // every grid point.  Here only the first round does; after it every 4th grid point is evaluated (stride 4)
      // and the skipped ones are examined only where they can matter: around a sign change between two coarse
      // points, and around a coarse point where log|Delta| has a kink -- two roots hidden between two coarse
      // points multiply the smooth background by (c - r1)(c - r2), whose second difference in log2 at one of the
      // two neighbouring coarse points is >= 3, while the background's is ~0.01 (Delta is normalised by the other
      // minors of the same sweep, which removes the scale that changes with the truncation depth).  The fine rounds
      // follow the reference exactly, so the bracket found is the reference's.  exact_scan keeps stride 1.
      constexpr int S = 4;
      constexpr float kKinkThr = 1.0f;
      // left neighbours in sequence order of the two points of this lane
      float dpx = __shfl_up_sync(gmask, pd.y, 1, G), cpx = __shfl_up_sync(gmask, pc.y, 1, G);
      if (gl == 0) { dpx = dP; cpx = cP; }
      const bool haspx = (gl > 0) || have_prev;
      const bool nanx = !(pc.x == pc.x) || !(pd.x == pd.x), nany = !(pc.y == pc.y) || !(pd.y == pd.y);
      const bool stopx = (pc.x < 0.8f * b_top) || (own_mj && !(pc.x < rec[mjx - 1].y + 0.3f)) || nanx;
      const bool stopy = (pc.y < 0.8f * b_top) || (own_mj && !(pc.y < rec[mjy - 1].y + 0.3f)) || nany;
      const bool chx = haspx && (signbit(dpx) != signbit(pd.x));
      const bool chy = (signbit(pd.x) != signbit(pd.y));
      if (stride == 1) {
        const bool stx = haspx && !chx && stopx;
        const bool sty = !chy && stopy;
        const unsigned ev = pair_mask(chx || stx, chy || sty);
        if (!ev) {
          // nothing in this fine round: go on with coarse rounds, whose left neighbours (at coarse spacing) are the
          // last point of this round and the point S before it
          constexpr int i2 = P - 1 - S;
          cP2 = gshfl<G>(gmask, (i2 & 1) ? pc.y : pc.x, i2 >> 1);
          dP2 = gshfl<G>(gmask, (i2 & 1) ? pd.y : pd.x, i2 >> 1);
          tP2 = __log2f(fabsf(dP2) / (fabsf(gshfl<G>(gmask, (i2 & 1) ? pe2.y : pe2.x, i2 >> 1)) + fabsf(gshfl<G>(gmask, (i2 & 1) ? pe3.y : pe3.x, i2 >> 1))));
          cP = gshfl<G>(gmask, pc.y, G - 1);
          dP = gshfl<G>(gmask, pd.y, G - 1);
          tP = __log2f(fabsf(dP) / (fabsf(gshfl<G>(gmask, pe2.y, G - 1)) + fabsf(gshfl<G>(gmask, pe3.y, G - 1))));
          have_prev = true;
          // (the half-space velocity of the sampled truncation is a kink, not a smooth place: fine rounds there)
          const bool near_kink = !(cP + 2.f * (float)S * p.dc < rec[meval - 1].y);
          if (p.exact_scan || near_kink) cbase = SD_ADD(cP, p.dc);
          else {
            stride = S;
            cbase = cP;
            for (int t = 0; t < S; ++t) cbase = SD_ADD(cbase, p.dc);
          }
          if (++round >= 2048) { flag |= SURFDISP_F_SCAN_LIMIT; flag |= (k == 0) ? SURFDISP_F_NO_ROOT_FIRST : SURFDISP_F_NO_ROOT_AT_K; model_done = true; }
          else need = NB_SCAN;
        } else {
          const int j = __ffs(ev) - 1, sl = j >> 1;
          const bool jy = (j & 1) != 0;
          const bool found = gshfl<G>(gmask, (int)(jy ? chy : chx), sl) != 0;
          lo = gshfl<G>(gmask, jy ? pc.x : cpx, sl); hi = gshfl<G>(gmask, jy ? pc.y : pc.x, sl);
          dlo = gshfl<G>(gmask, jy ? pd.x : dpx, sl); dhi = gshfl<G>(gmask, jy ? pd.y : pd.x, sl);
          mm = drop_coop(hi);   // the last DLTAR with idrop=0 leaves COMMON mmax (surfa.f:94-105)
          if (found) {
            // ---- polish inside [lo,hi] (replaces NEVILL, surfa.f:2-83).  The round's points sample one smooth
            // function around the bracket (same truncation depth), so the interpolation rounds of the fast path
            // take over: one more round instead of three of uniform section plus the ellipticity sweep.  With
            // the half-space velocity (kink, possibly several roots in the bracket) nearby, or in exact_scan
            // mode: uniform (P+1)-section with mmax pinned until the bracket is <= 2e-5, then one
            // secant step.
            bool handed = false;
            if (!p.exact_scan && !own_eval && j >= 1) {
              const float bh2 = rec[meval - 1].y;
              nvalid = __popc(pair_mask(pc.x < bh2, pc.y < bh2));
              const bool kink = (bh2 > lo - 0.011f && bh2 < hi + 0.011f) || nvalid < 6 || j > nvalid - 1;
              if (!kink) { w0 = 0; has_ends = false; jb = j; it = 0; fstage = 1; mw = meval; from_scan = true; do_interp = true; handed = true; }
            }
            if (!handed) {
              lo0 = lo; hi0 = hi; dlo0 = dlo; dhi0 = dhi; pit = 0;
              stage = ST_POLISH; need = NB_POLISH;
            }
          } else {
            flag |= (k == 0) ? SURFDISP_F_NO_ROOT_FIRST : SURFDISP_F_NO_ROOT_AT_K;
            model_done = true;
          }
        }
      } else {
        // coarse round: t = log2 of the normalised |Delta| of every point, second differences along the sequence
        const float tx = __log2f(fabsf(pd.x) / (fabsf(pe2.x) + fabsf(pe3.x))), ty = __log2f(fabsf(pd.y) / (fabsf(pe2.y) + fabsf(pe3.y)));
        float tpy = __shfl_up_sync(gmask, ty, 1, G), tpx = __shfl_up_sync(gmask, tx, 1, G);   // t of points 2g-1 and 2g-2
        float cp2 = __shfl_up_sync(gmask, pc.x, 1, G), dp2 = __shfl_up_sync(gmask, pd.x, 1, G); // point 2g-2
        if (gl == 0) { tpy = tP; tpx = tP2; cp2 = cP2; dp2 = dP2; }
        const float qx = tpx - 2.f * tpy + tx;      // kink at point 2g-1 seen from point 2g
        const float qy = tpy - 2.f * tx + ty;       // kink at point 2g seen from point 2g+1
        const float bhs = rec[meval - 1].y;
        const bool evx = chx || !(fabsf(qx) <= kKinkThr) || stopx || !(pc.x + (float)S * p.dc < bhs);
        const bool evy = chy || !(fabsf(qy) <= kKinkThr) || stopy || !(pc.y + (float)S * p.dc < bhs);
        const unsigned ev = pair_mask(evx, evy);
        if (!ev) {
          cP2 = gshfl<G>(gmask, pc.x, G - 1); dP2 = gshfl<G>(gmask, pd.x, G - 1); tP2 = gshfl<G>(gmask, tx, G - 1);
          cP = gshfl<G>(gmask, pc.y, G - 1); dP = gshfl<G>(gmask, pd.y, G - 1); tP = gshfl<G>(gmask, ty, G - 1);
          cbase = cP;
          for (int t = 0; t < S; ++t) cbase = SD_ADD(cbase, p.dc);
          if (++round >= 2048) { flag |= SURFDISP_F_SCAN_LIMIT; flag |= (k == 0) ? SURFDISP_F_NO_ROOT_FIRST : SURFDISP_F_NO_ROOT_AT_K; model_done = true; }
          else need = NB_SCAN;
        } else {
          // examine the 2 S grid points of the two coarse intervals before the event point like the reference does:
          // the point two before the event becomes the "previous point" of a fine round
          const int j = __ffs(ev) - 1, sl = j >> 1;
          const bool jy = (j & 1) != 0;
          cP = gshfl<G>(gmask, jy ? cpx : cp2, sl);    // point j-2: for an odd j that is point 2g-1, else point 2g-2
          dP = gshfl<G>(gmask, jy ? dpx : dp2, sl);
          have_prev = true;
          stride = 1;
          cbase = SD_ADD(cP, p.dc);
          if (++round >= 2048) { flag |= SURFDISP_F_SCAN_LIMIT; flag |= (k == 0) ? SURFDISP_F_NO_ROOT_FIRST : SURFDISP_F_NO_ROOT_AT_K; model_done = true; }
          else need = NB_SCAN;
        }
      }
    } else if (!NOSCAN && stage == ST_POLISH) {
      const unsigned ev = change_mask(pd, dlo);
      const float dlast = gshfl<G>(gmask, pd.y, G - 1);
      bool have_root = false, lstop = false;
      const float bhk = rec[mm - 1].y;
      if (pit == 0 && (__popc(ev) + (int)(signbit(dlast) != signbit(dhi)) > 1 || (bhk > lo0 && bhk < hi0))) {
        // several roots inside the scan bracket -- or the half-space velocity of the truncation inside it: beyond
        // that kink the function can turn back within 1e-4 km/s of a root just below it, which a uniform section
        // steps over --: follow the reference's own sequential bisection/Neville sequence so that the same root is
        // picked (all lanes run it redundantly)
        int ev_n = 0;
        SecFn f; f.kind = p.kind; f.T = T; f.mm = mm; f.rec = rec;
        const bool okp = nevill_out_of_line(f, lo0, hi0, dlo0, dhi0, &croot, &ev_n);
        if (gl == 0) { my_steps += (unsigned long long)ev_n * (unsigned)(mm - 1); my_sweeps += ev_n; }
        if (okp) have_root = true; else lstop = true;
      } else {
        if (ev) {
          const int j = __ffs(ev) - 1, sl = j >> 1;
          const bool jy = (j & 1) != 0;
          float ppx = __shfl_up_sync(gmask, pc.y, 1, G), dpx = __shfl_up_sync(gmask, pd.y, 1, G);
          if (gl == 0) { ppx = lo; dpx = dlo; }
          const float nlo = gshfl<G>(gmask, jy ? pc.x : ppx, sl), ndlo = gshfl<G>(gmask, jy ? pd.x : dpx, sl);
          hi = gshfl<G>(gmask, jy ? pc.y : pc.x, sl); dhi = gshfl<G>(gmask, jy ? pd.y : pd.x, sl);
          lo = nlo; dlo = ndlo;
        } else {
          lo = gshfl<G>(gmask, pc.y, G - 1); dlo = dlast;
        }
        if ((hi - lo) > kBracketTol && ++pit < 16) need = NB_POLISH;
        else {
          const float den = dhi - dlo;
          float cs = (den != 0.f) ? lo - dlo * (hi - lo) / den : 0.5f * (lo + hi);
          if (!(cs >= lo && cs <= hi)) cs = 0.5f * (lo + hi);
          croot = cs;
          have_root = true;
        }
      }
      if (lstop) { flag |= SURFDISP_F_LSTOP; nfound = 0; model_done = true; }   // reference aborts the whole call (calcul.f:173-189)
      else if (have_root) {
        if (croot > rec[mm - 1].y) {   // calcul.f:191
          flag |= SURFDISP_F_ROOT_ABOVE_HS;
          flag |= (k == 0) ? SURFDISP_F_NO_ROOT_FIRST : SURFDISP_F_NO_ROOT_AT_K;
          model_done = true;
        } else if (p.kind == 2) { stage = ST_ELL; need = NB_ELL; }
        else { ratio = 0.f; period_done = true; }
      }
    } else if (stage == ST_ELL) {
      // ellipticity = 0.5 * bb1(e3) / bb1(e2) at the root (surfa.f:360-363)
      ratio = 0.5f * pe3.x / pe2.x;
      period_done = true;
    }

    if (do_interp) {
      const int np = has_ends ? P + 2 : nvalid;
      const int s6 = min(max(jb - 3, 0), np - 6), s4 = min(max(jb - 2, 0), np - 4);
      const SamplePt B0 = sample(jb - 1), B1 = sample(jb);
      float x[6], y[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) { const SamplePt sp = sample(s6 + i); x[i] = sp.c - B0.c; y[i] = sp.d; }
      float e4, e6;
      inv_interp6(x, y, s4 - s6, e4, e6);
      const float w = B1.c - B0.c;
      const bool inside = (e6 > 0.f && e6 < w);
      const float delta = fabsf(e6 - e4);
      float e = e6;
      if (!inside) { const float den = B1.d - B0.d; e = (den != 0.f) ? -B0.d * w / den : 0.5f * w; }
      // accepted when the two orders agree AND the bracket has two samples on either side (one-sided
      // estimates agree with each other without being right); never on the 0.01 km/s grid of a window
      const bool interior = (jb >= 2 && jb <= np - 2);
      const float tol = (it > 0) ? kInterpTol : ((fstage == 0 && w <= kClusterWmax) ? kClusterTol : -1.f);
      if ((inside && interior && delta <= tol) || w <= kBracketTol) {
        croot = B0.c + e;
        if (p.kind == 2) {
          float xs[4], f2[4], f3[4], wl[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { const SamplePt sp = sample(s4 + i); xs[i] = sp.c - B0.c; f2[i] = sp.e2; f3[i] = sp.e3; }
          lagrange4(xs, e, wl);
          const float se2 = wl[0] * f2[0] + wl[1] * f2[1] + wl[2] * f2[2] + wl[3] * f2[3];
          const float se3 = wl[0] * f3[0] + wl[1] * f3[1] + wl[2] * f3[2] + wl[3] * f3[3];
          ratio = 0.5f * se3 / se2;
        } else ratio = 0.f;
        // the reference's bracket is the grid interval around the root; its upper end fixes mmax
        float hg = c1 + (floorf((croot - c1) / p.dc) + 1.f) * p.dc;
        if (!(hg > croot)) hg += p.dc;
        const int mnew = drop_coop(hg);
        const float bh1 = rec[mnew - 1].y;
        if (croot > bh1 || (bh1 > croot - 0.021f && bh1 < croot + 0.021f)) interp_failed();   // calcul.f:191 / kink: point-by-point path
        else {
          mm = mnew;
          if (p.kind == 2 && mid_liquid) { stage = ST_ELL; need = NB_ELL; } else period_done = true;
        }
      } else if (it >= 3) interp_failed();
      else {
        // one more round: P points around the estimate, spaced by the disagreement of the two orders
        const float span = (float)(1 << (P / 2 - 1));   // outermost offset of the round, in units of s0
        const float s0 = fmaxf(inside ? 0.5f * delta : w / (2.f * span), 1.0e-5f);
        const bool uni = !(e - span * s0 > 0.f && e + span * s0 < w);
        E0 = B0; E1 = B1;
        __syncwarp(gmask);
        if (gl == 0) { slots[0] = make_float4(B0.c, B0.d, B0.e2, B0.e3); slots[P + 1] = make_float4(B1.c, B1.d, B1.e2, B1.e3); }
        rf_e = e; rf_s0 = s0; rf_uni = uni;   // relative to E0.c
        stage = ST_REFINE; need = NB_REFINE;
      }
    }

    if (NOSCAN && stage == ST_DEFER) {
      // this period needs the point-by-point path, which this instantiation does not contain: the model is handed
      // over, with what the general instantiation needs to go on exactly where the scan would have started
      if (gl == 0) {
        p.nfound[model] = nfound; p.mm_state[model] = mm; if (p.flags) p.flags[model] = flag;
        p.dstate[model] = make_float2(pred_err, (float)hopped);
        p.defer_list[atomicAdd(p.ndefer, 1u)] = model;
      }
      stage = ST_FETCH;
    }
    if (period_done) {
      if (gl == 0) { crow[k] = croot; rrow[k] = ratio; p.mmh[(size_t)model * K + k] = (short)mm; }
      if (k >= 2 || (k == 1 && p.hint)) {
        // A root more than 0.1 km/s off the extrapolation: mode hopping, or a branch that bends too fast to be
        // extrapolated (thick slow sediments at the short-period end).  Scan from c1 like the reference until
        // two periods in a row were predictable again.
        pred_err = fabsf(croot - c_pred);
        if (pred_err > 0.1f) hopped = 2;
        else if (hopped > 0) hopped = (pred_err < 0.01f) ? hopped - 1 : 2;
      }
      c_prev3 = c_prev2; c_prev2 = c_prev; c_prev = croot;
      nfound = ++k;
      if (k == p.k_end) model_done = true; else stage = ST_PERIOD;
    }
    if (model_done) {
      if (nfound < p.k_end || p.k_end == K)
        for (int kk = nfound + gl; kk < K; kk += G) { crow[kk] = 0.f; rrow[kk] = 0.f; }
      if (gl == 0) { p.nfound[model] = nfound; p.mm_state[model] = mm; if (p.flags) p.flags[model] = flag; }
      stage = ST_FETCH;
    }
  }
  // work counters (roofline numerator)
  for (int o = 16; o > 0; o >>= 1) {
    my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    my_sweeps += __shfl_xor_sync(0xffffffffu, my_sweeps, o);
    my_models += __shfl_xor_sync(0xffffffffu, my_models, o);
  }
  if (lane == 0) {
    atomicAdd(&p.counters[0], my_steps);
    atomicAdd(&p.counters[1], my_sweeps);
    atomicAdd(&p.counters[3], (unsigned long long)my_models);
  }
}

// ------------------------------------------------------------------------------------ phase 2
struct P2Params {
  int kind, M, lpad, K, mpb, lmax;
  const int* nlay;
  const float* consts;
  const float* c_in;
  const float* ratio_in;
  const int* nfound;
  const float* dtot;
  float* u_out;
  unsigned long long* counters;
  float fact;
  int atten, ndiv, ndiv_cap;
  PeriodTab tab;
};

#ifndef P2_THREADS
#define P2_THREADS 256     // upper bound of a block; the launch uses the multiple of 32 that the work items fill
#define P2_MINBLK 2        // = 128 registers per thread
#endif
#ifndef P2_MINBLK_LOVE
#define P2_MINBLK_LOVE 3   // 65536 / (3 x 256) = 85 registers per thread
#endif
#ifndef P2_MINBLK_F32
#define P2_MINBLK_F32 2   // float32 Rayleigh state, packed pairs: 119 registers, no spills, 46.4 ms (80 registers with 250 B of spills: 46.7; 64: 49.2)
#endif
// KIND 1: Love (float32 analytic propagation, latency bound: 85 registers and a third more resident warps); KIND 2:
// Rayleigh with the float64 ODE state of the reference (128 registers); KIND 3: Rayleigh with the float32 state (default).
template <int KIND>
__global__ void __launch_bounds__(P2_THREADS, KIND == 2 ? P2_MINBLK : (KIND == 3 ? P2_MINBLK_F32 : P2_MINBLK_LOVE)) phase2_kernel(const __grid_constant__ P2Params p) {
  extern __shared__ float4 smem[];
  float* sc = reinterpret_cast<float*>(smem);
  const int K = p.K;
  const int per_model = NCONST * p.lpad;
  const int model0 = blockIdx.x * p.mpb;
  const int nmod = min(p.mpb, p.M - model0);
  unsigned long long nsub = 0;
  // stage the constants of this block's models: coalesced float4 reads of the rows [8][lpad], written as records
  // [layer][8] -- a thread then reads all constants of a layer from one address (two 16-byte-aligned groups)
  {
    const float4* src = reinterpret_cast<const float4*>(p.consts + (size_t)model0 * per_model);
    const int n4 = nmod * per_model / 4, l4 = p.lpad / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = src[i];
      const int ml = i / (NCONST * l4), r = i - ml * (NCONST * l4), comp = r / l4, lay = (r - comp * l4) * 4;
      float* dst = sc + (size_t)ml * (per_model + NCONST) + (size_t)lay * NCONST + comp;   // (+ one record: the models' bank sets differ)
      dst[0] = v.x; dst[NCONST] = v.y; dst[2 * NCONST] = v.z; dst[3 * NCONST] = v.w;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nmod * K; t += blockDim.x) {
    // period-major inside the block: the 32 work items of a warp are neighbouring periods of the block's models,
    // i.e. integrations of similar depth (the depth grows with the period)
    const int k = t / nmod, ml = t - k * nmod;
    const int model = model0 + ml;
    float* urow = p.u_out + (size_t)model * K;
    const int n = p.nlay[model];
    if (k >= p.nfound[model] || n < 2 || n > p.lmax) {
      urow[k] = 0.f;
    } else {
      ModelView mv;
      mv.cst = sc + (size_t)ml * (per_model + NCONST);
      mv.sc = 1; mv.sl = NCONST; mv.n = n; mv.atten = p.atten; mv.lt = p.tab.lt[k]; mv.dtot = p.dtot[model];
      int ndiv = p.ndiv;
      const int ivre = p.ndiv_cap / (n - 1);
      if (ndiv > ivre) ndiv = ivre;
      mv.ndiv = ndiv;
      mv.jj0 = (mv.at(C_BREF, 0) <= 0.1e-10f) ? 1 : 0;
      const float T = p.tab.per[k];
      const float c = p.c_in[(size_t)model * K + k];
      float u;
      if (KIND == 2) u = reigen_thread2(mv, T, c, p.ratio_in[(size_t)model * K + k], p.fact, nsub);
      else if (KIND == 3) u = reigen_thread2_t<float, true>(mv, T, c, p.ratio_in[(size_t)model * K + k], p.fact, nsub, 1);
      else u = leigen_thread(mv, T, c, p.fact, nsub);
      urow[k] = u;
    }
  }
  // counters: one atomic per warp (all lanes arrive here)
  for (int o = 16; o > 0; o >>= 1) nsub += __shfl_xor_sync(0xffffffffu, nsub, o);
  if ((threadIdx.x & 31) == 0 && nsub) atomicAdd(&p.counters[2], nsub);
}

// ------------------------------------------------------------------------------------ partial derivatives
// Thread per (model, period) like phase 2: REIGEN's dcda / dcdb / dcdr (surfa.f:1130-1135, 1179-1185, 1202-1208).
struct PdParams {
  int M, lpad, K, mpb, lmax;
  const int* nlay;
  const float* consts;
  const float* c_in;
  const float* ratio_in;
  const int* nfound;
  const float* dtot;
  float *dcda, *dcdb, *dcdr;     // [M][K][lmax]
  float fact;
  int atten, ndiv, ndiv_cap;
  PeriodTab tab;
};

__global__ void __launch_bounds__(P2_THREADS, P2_MINBLK) partials_kernel(const __grid_constant__ PdParams p) {
  extern __shared__ float4 smem[];
  float* sc = reinterpret_cast<float*>(smem);
  const int K = p.K;
  const int model0 = blockIdx.x * p.mpb;
  const int nmod = min(p.mpb, p.M - model0);
  const int per_model = NCONST * p.lpad;
  {
    const float4* src = reinterpret_cast<const float4*>(p.consts + (size_t)model0 * per_model);
    const int n4 = nmod * per_model / 4, l4 = p.lpad / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = src[i];
      const int ml = i / (NCONST * l4), r = i - ml * (NCONST * l4), comp = r / l4, lay = (r - comp * l4) * 4;
      float* dst = sc + (size_t)ml * (per_model + NCONST) + (size_t)lay * NCONST + comp;
      dst[0] = v.x; dst[NCONST] = v.y; dst[2 * NCONST] = v.z; dst[3 * NCONST] = v.w;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < nmod * K; t += blockDim.x) {
    const int k = t / nmod, ml = t - k * nmod;
    const int model = model0 + ml;
    const size_t o = ((size_t)model * K + k) * p.lmax;
    const int n = p.nlay[model];
    bool done = false;
    if (k < p.nfound[model] && n >= 2 && n <= p.lmax) {
      ModelView mv;
      mv.cst = sc + (size_t)ml * (per_model + NCONST);
      mv.sc = 1; mv.sl = NCONST; mv.n = n; mv.atten = p.atten; mv.lt = p.tab.lt[k]; mv.dtot = p.dtot[model];
      int ndiv = p.ndiv;
      const int ivre = p.ndiv_cap / (n - 1);
      if (ndiv > ivre) ndiv = ivre;
      mv.ndiv = ndiv;
      mv.jj0 = (mv.at(C_BREF, 0) <= 0.1e-10f) ? 1 : 0;
      done = reigen_partials_thread(mv, p.tab.per[k], p.c_in[(size_t)model * K + k], p.ratio_in[(size_t)model * K + k], p.fact,
                                    p.dcda + o, p.dcdb + o, p.dcdr + o, 1);
      for (int j = n; j < p.lmax; ++j) { p.dcda[o + j] = 0.f; p.dcdb[o + j] = 0.f; p.dcdr[o + j] = 0.f; }
    }
    if (!done) for (int j = 0; j < p.lmax; ++j) { p.dcda[o + j] = 0.f; p.dcdb[o + j] = 0.f; p.dcdr[o + j] = 0.f; }
  }
}

// ------------------------------------------------------------------------------------ misfit
struct MisfitParams {
  int mode, M, K, ncount;
  const float* c_pred;
  const int* nfound;
  float* out;
  float obs[kMaxPer], isig[kMaxPer], per[kMaxPer];
  unsigned char use[kMaxPer];
};

// one warp per model, lanes over periods, shuffle reduction (point.py:15-31, 337-366)
__global__ void __launch_bounds__(256) misfit_kernel(const __grid_constant__ MisfitParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.M) return;
  const float* row = p.c_pred + (size_t)warp * p.K;
  double s1 = 0.0, s2 = 0.0;
  int n1 = 0, n2 = 0;
  for (int k = lane; k < p.K; k += 32) {
    if (!p.use[k]) continue;
    const double bias = ((double)p.obs[k] - (double)row[k]) * (double)p.isig[k];
    if (p.mode == 1 && p.per[k] > 40.f) { s2 += bias * bias; n2++; }
    else { s1 += bias * bias; n1++; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    n1 += __shfl_xor_sync(0xffffffffu, n1, o); n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  if (lane != 0) return;
  float* o = p.out + (size_t)warp * 3;
  if (p.nfound[warp] < p.K) { o[0] = 88888.f; o[1] = 88888.f; o[2] = 0.f; return; }  // point.py:20-21
  const int N = n1 + n2;
  double chi;
  if (p.mode == 1) {
    if (n1 > 0 && n2 > 0) chi = (s1 / n1 + s2 / n2) / 2.0 * N;
    else if (n2 > 0) chi = s2 / n2 * N;
    else chi = s1 / (n1 > 0 ? n1 : 1) * N;
  } else chi = s1;
  const double misfit = sqrt(chi / (N > 0 ? N : 1));
  if (!(chi < 50.0)) chi = sqrt(chi * 50.0);  // point.py:29
  o[0] = (float)misfit; o[1] = (float)chi; o[2] = (float)exp(-0.5 * chi);
}



// ------------------------------------------------------------------------------------ model builder, Monte-Carlo step
}  // namespace
#include "surfdisp_mc.cuh"
namespace {
using namespace mcdev;

// ------------------------------------------------------------------------------------ pipe peaks
// Register-resident FMA / MUFU chains: the denominators of the FP-pipe roofline (MEASURED_PEAKS.json
// only has HBM and bf16 tensor figures, neither of which bounds this path).
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters) {
  T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
  T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
  const T m = (T)0.999999, c = (T)1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
      a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(float* out, int iters) {
  float a0 = 1.0f + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = __expf(a0) * 0.1f; a1 = __expf(a1) * 0.1f; a2 = __expf(a2) * 0.1f; a3 = __expf(a3) * 0.1f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}

int fill_tab(PeriodTab& tab, int K, const float* periods, float t_base) {
  if (K < 1 || K > kMaxPer) return SURFDISP_EINVAL;
  memset(&tab, 0, sizeof(tab));
  for (int k = 0; k < K; ++k) {
    if (!(periods[k] > 0.f)) return SURFDISP_EINVAL;
    tab.per[k] = periods[k];
    tab.lt[k] = logf(t_base / periods[k]);  // host libm, as gfortran's alog (calcul.f:122)
  }
  return 0;
}

template <int G, int MODE>
int launch_phase1_t(const P1Params& p, cudaStream_t st) {
  P1Params q = p;
  q.mstride = p.lpad + 1;  // +1 float4: consecutive groups start 16 B apart mod 128 B (bank spread)
  // 128-thread CTAs (128/G models) for ordinary stacks; deep stacks (up to 1000 layers, 16 KB of layer records
  // per model) shrink the CTA until the records of its models fit in shared memory
  int threads = P1_THREADS;
  while (threads > 32 && (size_t)(threads / G) * (q.mstride + 2 * G + 2) * sizeof(float4) > 100 * 1024) threads /= 2;
  const int groups = threads / G;
  size_t smem = (size_t)groups * (q.mstride + 2 * G + 2) * sizeof(float4);   // layer records + sample slots
  if (smem > 220 * 1024) return SURFDISP_EINVAL;
  CK(cudaFuncSetAttribute(phase1_kernel<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148, occ = 1;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, phase1_kernel<G, MODE>, threads, smem));
  if (occ < 1) occ = 1;
  long long need = ((long long)p.M + groups - 1) / groups;
  long long grid = (long long)sms * occ;  // persistent: one resident wave, models pulled from a queue
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  phase1_kernel<G, MODE><<<(unsigned)grid, threads, smem, st>>>(q);
  CK(cudaGetLastError());
  return 0;
}

#ifndef P1_SPLIT
#define P1_SPLIT 1      // later periods: fast-path instantiation + hand-over to the general one (0: the general one alone)
#endif
// Three instantiations of the root search:
//   P1_FIRST    one period per model from scratch: every model scans, no cluster / window / refinement code
//   P1_FAST     later periods, interpolation rounds only: no scan / polish code and none of its state in registers;
//               a model whose period needs the point-by-point path (1 % of the config-2 models) is handed over
//   P1_GENERAL  everything; runs the handed-over models from the period they stopped at (and exact_scan launches)
// The hand-over pays for large batches only: the handed-over models -- among them the few that scan at every period,
// a serial chain of a few milliseconds -- start when the fast-path launch has ended.  Small batches (Monte-Carlo
// ensembles of a few thousand chains) keep the single general launch.  Measured: Rayleigh sweep of
// 2^20 models unchanged, Love sweep 229 -> 203 ms (a third of the Love models scan somewhere: in one launch their
// scans sit in warps beside models on the fast path), 131072-chain Monte-Carlo step 21.8 -> 21.2 ms, 256 chains
// 1.35 -> 1.96 ms (hence the threshold).
constexpr int kSplitMinDefault = 98304;
int g_split_min_models = kSplitMinDefault;   // (surfdisp_set_split_min_models: tests force the hand-over on small batches)
static bool p1_split(const P1Params& p) { return P1_SPLIT && p.k_begin >= 1 && !p.exact_scan && p.M >= g_split_min_models; }

template <int G>
int launch_phase1(const P1Params& p, cudaStream_t st) {
  if (p.k_begin == 0 && p.k_end == 1 && !p.exact_scan) return launch_phase1_t<G, P1_FIRST>(p, st);
  if (!p1_split(p)) return launch_phase1_t<G, P1_GENERAL>(p, st);
  CK(cudaMemsetAsync(p.ndefer, 0, sizeof(unsigned int), st));
  CK(cudaMemsetAsync(p.queue, 0, sizeof(unsigned int), st));
  const int rc = launch_phase1_t<G, P1_FAST>(p, st);
  if (rc) return rc;
  P1Params r = p;
  r.resume = 1; r.order = p.defer_list; r.order_base = 0;
  CK(cudaMemsetAsync(p.queue, 0, sizeof(unsigned int), st));
  return launch_phase1_t<G, P1_GENERAL>(r, st);
}

}  // namespace

// =============================================================================================== C ABI
extern "C" {

void surfdisp_default_opts(SurfdispOpts* o) {
  o->dc = 0.01f; o->fact = 4.0f; o->t_base = 1.0f; o->ndiv = 5; o->ndiv_cap_rayleigh = 99;
  o->ndiv_cap_love = 999; o->atten = 1; o->flatten = 1; o->stale_mmax = 1; o->compute_group = 1; o->exact_scan = 0;
  o->group_f64 = 0;
}

size_t surfdisp_workspace_bytes(int n_models, int n_layers_max, int n_periods) {
  if (n_models < 0 || n_layers_max < 2 || n_periods < 1) return 0;
  return ws_layout(n_models, n_layers_max, n_periods).total;
}

// ---- order in which the models are handed out: by the shear velocity of the top solid layer (1024 buckets).  The
// first-period scan starts at 0.9 of that velocity (fast_surf.f:157-171), so models with similar keys scan for
// similarly long, and a warp's eight models pass through their periods at a similar pace (first-period launch and
// later periods together: -5 %).  The order inside a bucket is whatever the atomics give; results do not depend on it.
__device__ __forceinline__ int order_bucket(const float* __restrict__ layers, size_t comp_stride, int lmax, const int* nlay, int m) {
  const int n = nlay[m];
  if (n < 2 || n > lmax) return 0;
  const float* vs = layers + comp_stride + (size_t)m * lmax;
  const float b = (vs[0] < 0.1f) ? vs[1] : vs[0];
  int k = (int)(b * 200.0f);
  return k < 0 ? 0 : (k >= kOrderBuckets ? kOrderBuckets - 1 : k);
}
__global__ void order_count_kernel(int M, int lmax, const int* __restrict__ nlay, const float* __restrict__ layers,
                                   size_t comp_stride, int* __restrict__ count) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) atomicAdd(&count[order_bucket(layers, comp_stride, lmax, nlay, m)], 1);
}
__global__ void order_scan_kernel(const int* __restrict__ count, int* __restrict__ offset) {
  // one block of kOrderBuckets threads: exclusive prefix sum
  __shared__ int sh[kOrderBuckets];
  const int t = threadIdx.x;
  sh[t] = count[t];
  __syncthreads();
  for (int o = 1; o < kOrderBuckets; o <<= 1) {
    const int v = (t >= o) ? sh[t - o] : 0;
    __syncthreads();
    sh[t] += v;
    __syncthreads();
  }
  offset[t] = sh[t] - count[t];
}
__global__ void order_scatter_kernel(int M, int lmax, const int* __restrict__ nlay, const float* __restrict__ layers,
                                     size_t comp_stride, int* __restrict__ offset, int* __restrict__ order, int base) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) order[atomicAdd(&offset[order_bucket(layers, comp_stride, lmax, nlay, m)], 1)] = base + m;
}

// ---- the batch as a plan plus stages over a range of models [a, b): surfdisp_batch runs every stage on the whole
// batch, the host path runs the first stages chunk by chunk under the host->device copies
struct Plan {
  SurfdispOpts o;
  int kind, M, lmax, K;
  WsLayout w;
  const int* nlay; const float* layers;
  float *c_out, *u_out, *consts, *ratio;
  int *nfound, *flags, *mm_state, *order, *buckets;
  float* dtot;     // per model: upper bound of the layer-dropping thickness sums (prep_kernel)
  short* mmh; int* defer_list; float2* dstate; unsigned int* ndefer;   // hand-over between the root-search launches
  unsigned long long* counters;
  unsigned int* queue;
  const float* hint;
  PeriodTab tab;
};

static int make_plan(Plan& pl, const SurfdispOpts* opts, int kind, int n_models, int n_layers_max, const int* n_layers,
                     const float* layers, int n_periods, const float* periods, float* c_out, float* u_out,
                     int* nfound, int* flags, void* workspace, size_t workspace_bytes) {
  if (opts) pl.o = *opts; else surfdisp_default_opts(&pl.o);
  if ((kind != 1 && kind != 2) || n_models < 0 || n_layers_max < 2 || n_layers_max > SURFDISP_MAX_LAYERS ||
      n_periods < 1 || n_periods > kMaxPer || !periods || !(pl.o.dc > 0.f))
    return SURFDISP_EINVAL;
  if (n_models == 0) return 0;
  if (!n_layers || !layers || !c_out || !nfound || !workspace) return SURFDISP_EINVAL;
  pl.w = ws_layout(n_models, n_layers_max, n_periods);
  if (workspace_bytes < pl.w.total) return SURFDISP_ENOMEM;
  char* ws = (char*)workspace;
  pl.kind = kind; pl.M = n_models; pl.lmax = n_layers_max; pl.K = n_periods;
  pl.nlay = n_layers; pl.layers = layers; pl.c_out = c_out; pl.u_out = u_out; pl.nfound = nfound; pl.flags = flags;
  pl.counters = (unsigned long long*)ws;
  pl.queue = (unsigned int*)(ws + 64);
  pl.consts = (float*)(ws + pl.w.consts_off);
  pl.ratio = (float*)(ws + pl.w.ratio_off);
  pl.mm_state = (int*)(ws + pl.w.mm_off);
  pl.order = (int*)(ws + pl.w.order_off);
  pl.dtot = (float*)(ws + pl.w.dtot_off);
  pl.mmh = (short*)(ws + pl.w.mmh_off); pl.defer_list = (int*)(ws + pl.w.defer_off); pl.dstate = (float2*)(ws + pl.w.dstate_off);
  pl.ndefer = (unsigned int*)(ws + 128);
  pl.buckets = (int*)(ws + pl.w.bucket_off);
  pl.hint = nullptr;
  return fill_tab(pl.tab, n_periods, periods, pl.o.t_base);
}

// the first launch of the root search covers the periods [0, k_split)
static int k_split(const Plan& pl) { return (pl.K > 1 && !pl.o.exact_scan) ? 1 : pl.K; }

static int stage_prep(const Plan& pl, int a, int b, cudaStream_t st) {
  const int m = b - a;
  if (m <= 0) return 0;
  prep_kernel<<<(unsigned)(((size_t)m * 32 + 127) / 128), 128, 0, st>>>(
      m, pl.lmax, pl.w.lpad, pl.kind, pl.o.flatten, pl.nlay + a, pl.layers + (size_t)a * pl.lmax, (size_t)pl.M * pl.lmax,
      pl.consts + (size_t)a * NCONST * pl.w.lpad, pl.dtot + a);
  CK(cudaGetLastError());
  // hand-out order of this range: order[a .. b) = the models a .. b-1 sorted by bucket
  CK(cudaMemsetAsync(pl.buckets, 0, 2 * kOrderBuckets * sizeof(int), st));
  const float* lay = pl.layers + (size_t)a * pl.lmax;
  const size_t cs = (size_t)pl.M * pl.lmax;
  order_count_kernel<<<(m + 255) / 256, 256, 0, st>>>(m, pl.lmax, pl.nlay + a, lay, cs, pl.buckets);
  order_scan_kernel<<<1, kOrderBuckets, 0, st>>>(pl.buckets, pl.buckets + kOrderBuckets);
  order_scatter_kernel<<<(m + 255) / 256, 256, 0, st>>>(m, pl.lmax, pl.nlay + a, lay, cs, pl.buckets + kOrderBuckets, pl.order + a, a);
  CK(cudaGetLastError());
  return 0;
}

// root search of the periods [k_begin, k_end) for the models [a, b)
static void fill_p1(const Plan& pl, int a, int b, int k_begin, int k_end, P1Params& p1);
static int stage_p1(const Plan& pl, int a, int b, int k_begin, int k_end, cudaStream_t st) {
  if (b <= a || k_end <= k_begin) return 0;
  P1Params p1;
  fill_p1(pl, a, b, k_begin, k_end, p1);
  if (!p1_split(p1)) CK(cudaMemsetAsync(pl.queue, 0, sizeof(unsigned int), st));   // (the split launches reset it themselves)
  return launch_phase1<P1_G>(p1, st);
}
static void fill_p1(const Plan& pl, int a, int b, int k_begin, int k_end, P1Params& p1) {
  memset(&p1, 0, sizeof(p1));
  p1.kind = pl.kind; p1.M = b - a; p1.lpad = pl.w.lpad; p1.lmax = pl.lmax; p1.K = pl.K; p1.nlay = pl.nlay + a;
  p1.consts = pl.consts + (size_t)a * NCONST * pl.w.lpad;
  p1.c_out = pl.c_out + (size_t)a * pl.K; p1.ratio_out = pl.ratio + (size_t)a * pl.K; p1.nfound = pl.nfound + a;
  p1.flags = pl.flags ? pl.flags + a : nullptr; p1.counters = pl.counters; p1.queue = pl.queue;
  p1.dc = pl.o.dc; p1.fact = pl.o.fact; p1.atten = pl.o.atten; p1.stale = pl.o.stale_mmax; p1.exact_scan = pl.o.exact_scan;
  p1.tab = pl.tab;
  p1.mm_state = pl.mm_state + a; p1.dtot = pl.dtot + a;
  p1.mmh = pl.mmh + (size_t)a * pl.K; p1.defer_list = pl.defer_list + a; p1.dstate = pl.dstate + a; p1.ndefer = pl.ndefer; p1.resume = 0;
  p1.order = pl.order + a; p1.order_base = a;
  p1.hint = pl.hint ? pl.hint + (size_t)a * pl.K : nullptr;
  p1.k_begin = k_begin; p1.k_end = k_end;
}
static int stage_p2(const Plan& pl, int a, int b, cudaStream_t st) {
  const int m = b - a;
  if (m <= 0 || !pl.u_out) return 0;
  if (!pl.o.compute_group) {
    CK(cudaMemsetAsync(pl.u_out + (size_t)a * pl.K, 0, (size_t)m * pl.K * sizeof(float), st));
    return 0;
  }
  P2Params p2;
  memset(&p2, 0, sizeof(p2));
  p2.kind = pl.kind; p2.M = m; p2.lpad = pl.w.lpad; p2.lmax = pl.lmax; p2.K = pl.K; p2.nlay = pl.nlay + a;
  p2.consts = pl.consts + (size_t)a * NCONST * pl.w.lpad;
  p2.c_in = pl.c_out + (size_t)a * pl.K; p2.ratio_in = pl.ratio + (size_t)a * pl.K; p2.nfound = pl.nfound + a;
  p2.u_out = pl.u_out + (size_t)a * pl.K; p2.counters = pl.counters; p2.dtot = pl.dtot + a;
  p2.fact = pl.o.fact; p2.atten = pl.o.atten; p2.ndiv = pl.o.ndiv;
  p2.ndiv_cap = (pl.kind == 2) ? pl.o.ndiv_cap_rayleigh : pl.o.ndiv_cap_love;
  p2.tab = pl.tab;
  const size_t per_model = (size_t)NCONST * (pl.w.lpad + 1) * sizeof(float);   // shared-memory record stride of a model
  // models per block: the one whose (models x periods) work items fill whole warps best (40 periods: 4 models =
  // 160 threads, no idle lane; 3 models in 128 threads leave 8 of 128 lanes idle in every FP64 instruction)
  int mpb = 1;
  {
    const int mmax_blk = P2_THREADS / pl.K < 1 ? 1 : P2_THREADS / pl.K;
    double best = -1.0;
    for (int m2 = 1; m2 <= mmax_blk; ++m2) {
      if (m2 > 1 && m2 * per_model > 96 * 1024) break;
      const int items = m2 * pl.K, thr = round_up(items, 32) > P2_THREADS ? P2_THREADS : round_up(items, 32);
      const double eff = (double)items / ((double)thr * ((items + thr - 1) / thr));
      if (eff > best + 1e-9) { best = eff; mpb = m2; }
    }
  }
  if (mpb * per_model > 200 * 1024) return SURFDISP_EINVAL;
  p2.mpb = mpb;
  int threads = round_up(mpb * pl.K, 32);
  if (threads > P2_THREADS) threads = P2_THREADS;
  const size_t smem = mpb * per_model;
  const int grid = (m + mpb - 1) / mpb;
  if (pl.kind == 2 && pl.o.group_f64) {
    CK(cudaFuncSetAttribute(phase2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    phase2_kernel<2><<<grid, threads, smem, st>>>(p2);
  } else if (pl.kind == 2) {
    CK(cudaFuncSetAttribute(phase2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    phase2_kernel<3><<<grid, threads, smem, st>>>(p2);
  } else {
    CK(cudaFuncSetAttribute(phase2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    phase2_kernel<1><<<grid, threads, smem, st>>>(p2);
  }
  CK(cudaGetLastError());
  return 0;
}

static thread_local const float* g_batch_hint = nullptr;   // set by surfdisp_batch_hinted around its call of surfdisp_batch

int surfdisp_batch_hinted(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max, const int* n_layers,
                          const float* layers, int n_periods, const float* periods, const float* c_hint, float* c_out,
                          float* u_out, int* nfound, int* flags, void* workspace, size_t workspace_bytes, void* stream) {
  g_batch_hint = c_hint;
  const int rc = surfdisp_batch(opts, kind, n_models, n_layers_max, n_layers, layers, n_periods, periods, c_out, u_out, nfound,
                                flags, workspace, workspace_bytes, stream);
  g_batch_hint = nullptr;
  return rc;
}

int surfdisp_batch(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max, const int* n_layers,
                   const float* layers, int n_periods, const float* periods, float* c_out, float* u_out,
                   int* nfound, int* flags, void* workspace, size_t workspace_bytes, void* stream) {
  Plan pl;
  int rc = make_plan(pl, opts, kind, n_models, n_layers_max, n_layers, layers, n_periods, periods, c_out, u_out, nfound,
                     flags, workspace, workspace_bytes);
  if (rc || n_models == 0) return rc;
  pl.hint = g_batch_hint;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(workspace, 0, kHdrBytes, st));
  if (g_prof_events) CK(cudaEventRecord(g_prof_events[0], st));
  if ((rc = stage_prep(pl, 0, n_models, st))) return rc;
  if (g_prof_events) CK(cudaEventRecord(g_prof_events[1], st));
  // The scan of the first period (calcul.f:155-167 from 0.9 b(1)) consists of many shallow sweeps, the later
  // periods of one or two deep ones each: run as separate launches, the groups of a warp never mix the two.
  const int ks = k_split(pl);
  if ((rc = stage_p1(pl, 0, n_models, 0, ks, st))) return rc;
  if ((rc = stage_p1(pl, 0, n_models, ks, n_periods, st))) return rc;
  if (g_prof_events) CK(cudaEventRecord(g_prof_events[2], st));
  if ((rc = stage_p2(pl, 0, n_models, st))) return rc;
  if (g_prof_events) CK(cudaEventRecord(g_prof_events[3], st));
  return 0;
}

static bool has_hybrid(const SurfdispStackTemplate* tmpl) {
  for (int g = 0; g < tmpl->ngroups; ++g) if (tmpl->groups[g].kind == SURFDISP_G_HYBRID) return true;
  return false;
}

static int check_template(const SurfdispStackTemplate* tmpl) {
  if (!tmpl) return SURFDISP_EINVAL;
  if (tmpl->ngroups < 1 || tmpl->ngroups > SURFDISP_MAX_GROUPS || tmpl->nparams < 0) return SURFDISP_EINVAL;
  for (int g = 0; g < tmpl->ngroups; ++g) {
    const SurfdispStackGroup& G = tmpl->groups[g];
    if (G.ncoef < 0 || G.ncoef > SURFDISP_MAX_COEF || G.h_param >= tmpl->nparams) return SURFDISP_EINVAL;
    if (G.kind < SURFDISP_G_WATER || G.kind > SURFDISP_G_HYBRID) return SURFDISP_EINVAL;
    if (G.kind == SURFDISP_G_HYBRID && (G.ncoef < 1 || G.ncoef + 1 > SURFDISP_MAX_COEF || G.age_param >= tmpl->nparams)) return SURFDISP_EINVAL;
    if (G.nfine_rule == SURFDISP_N_FIXED && G.nfine < 1) return SURFDISP_EINVAL;
    if ((G.kind == SURFDISP_G_LINEAR && G.ncoef < 2) || ((G.kind == SURFDISP_G_CONST || G.kind == SURFDISP_G_BSPLINE) && G.ncoef < 1)) return SURFDISP_EINVAL;
    for (int i = 0; i < G.ncoef; ++i) if (G.v_param[i] >= tmpl->nparams) return SURFDISP_EINVAL;
  }
  return 0;
}

int surfdisp_partials_batch(const SurfdispOpts* opts, int n_models, int n_layers_max, const int* n_layers, const float* layers,
                            int n_periods, const float* periods, float* c_out, float* dcda, float* dcdb, float* dcdr,
                            int* nfound, int* flags, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dcda || !dcdb || !dcdr) return SURFDISP_EINVAL;
  Plan pl;
  int rc = make_plan(pl, opts, SURFDISP_KIND_RAYLEIGH, n_models, n_layers_max, n_layers, layers, n_periods, periods, c_out, nullptr,
                     nfound, flags, workspace, workspace_bytes);
  if (rc || n_models == 0) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(workspace, 0, kHdrBytes, st));
  if ((rc = stage_prep(pl, 0, n_models, st))) return rc;
  const int ks = k_split(pl);
  if ((rc = stage_p1(pl, 0, n_models, 0, ks, st))) return rc;
  if ((rc = stage_p1(pl, 0, n_models, ks, n_periods, st))) return rc;
  PdParams pd;
  memset(&pd, 0, sizeof(pd));
  pd.M = n_models; pd.lpad = pl.w.lpad; pd.K = pl.K; pd.lmax = pl.lmax; pd.nlay = pl.nlay; pd.consts = pl.consts; pd.c_in = pl.c_out;
  pd.ratio_in = pl.ratio; pd.nfound = pl.nfound; pd.dtot = pl.dtot; pd.dcda = dcda; pd.dcdb = dcdb; pd.dcdr = dcdr; pd.fact = pl.o.fact;
  pd.atten = pl.o.atten; pd.ndiv = pl.o.ndiv; pd.ndiv_cap = pl.o.ndiv_cap_rayleigh; pd.tab = pl.tab;
  const size_t per_model = (size_t)NCONST * (pl.w.lpad + 1) * sizeof(float);
  if (per_model > 200 * 1024) return SURFDISP_EINVAL;
  int mpb = (int)((96 * 1024) / per_model);
  if (mpb < 1) mpb = 1;
  if (mpb * pl.K > P2_THREADS) mpb = P2_THREADS / pl.K < 1 ? 1 : P2_THREADS / pl.K;
  pd.mpb = mpb;
  int threads = round_up(mpb * pl.K, 32);
  if (threads > P2_THREADS) threads = P2_THREADS;
  const size_t smem = mpb * per_model;
  CK(cudaFuncSetAttribute(partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  partials_kernel<<<(n_models + mpb - 1) / mpb, threads, smem, st>>>(pd);
  CK(cudaGetLastError());
  return 0;
}

size_t surfdisp_pipelined_bytes(int n_models, int n_layers_max, int n_periods) {
  if (n_models < 0 || n_layers_max < 2 || n_periods < 1) return 0;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t nl = (size_t)5 * n_models * n_layers_max * sizeof(float), no = (size_t)n_models * n_periods * sizeof(float);
  return al(nl) + 3 * al((size_t)n_models * sizeof(int)) + 2 * al(no) + surfdisp_workspace_bytes(n_models, n_layers_max, n_periods);
}

size_t surfdisp_params_pipelined_bytes(int n_models, int n_params, int n_layers_max, int n_periods) {
  if (n_params < 0) return 0;
  const size_t base = surfdisp_pipelined_bytes(n_models, n_layers_max, n_periods);
  return base ? base + ((size_t)n_models * n_params * sizeof(float) + 255) / 256 * 256 : 0;
}

// The host-buffer pipeline.  Source of the stacks: host layers (copied chunk by chunk under the first stages) or host
// PARAMETER vectors (one small copy, the stacks are assembled on the device by build_stacks_kernel).
static int host_pipeline(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max, const int* n_layers,
                         const float* layers, const SurfdispStackTemplate* tmpl, const float* params, int n_periods,
                         const float* periods, float* c_out, float* u_out, int* nfound, int* flags, void* device_buffer,
                         size_t device_bytes, int n_chunks, void* compute_stream, void* copy_stream) {
  if (n_models < 0 || n_layers_max < 2 || n_periods < 1 || n_periods > kMaxPer) return SURFDISP_EINVAL;
  if (n_models == 0) return 0;
  const bool from_params = (tmpl != nullptr);
  if (!c_out || !nfound || !periods || !device_buffer) return SURFDISP_EINVAL;
  if (from_params ? (!params && tmpl->nparams > 0) : (!n_layers || !layers)) return SURFDISP_EINVAL;
  const int P = from_params ? tmpl->nparams : 0;
  if (device_bytes < (from_params ? surfdisp_params_pipelined_bytes(n_models, P, n_layers_max, n_periods)
                                  : surfdisp_pipelined_bytes(n_models, n_layers_max, n_periods)))
    return SURFDISP_ENOMEM;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > 64) n_chunks = 64;
  if (n_chunks > n_models) n_chunks = n_models;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t M = (size_t)n_models, L = (size_t)n_layers_max, K = (size_t)n_periods;
  const size_t nl = 5 * M * L * sizeof(float), no = M * K * sizeof(float), ni = M * sizeof(int), np_ = M * P * sizeof(float);
  char* dev = (char*)device_buffer;
  const size_t o_lay = 0, o_n = o_lay + al(nl), o_c = o_n + al(ni), o_u = o_c + al(no), o_nf = o_u + al(no),
               o_fl = o_nf + al(ni), o_par = o_fl + al(ni), o_ws = o_par + (from_params ? al(np_) : 0);
  float* d_lay = (float*)(dev + o_lay);
  int* d_n = (int*)(dev + o_n);
  float *d_c = (float*)(dev + o_c), *d_u = u_out ? (float*)(dev + o_u) : nullptr;
  int *d_nf = (int*)(dev + o_nf), *d_fl = (int*)(dev + o_fl);
  float* d_par = (float*)(dev + o_par);
  Plan pl;
  int rc = make_plan(pl, opts, kind, n_models, n_layers_max, d_n, d_lay, n_periods, periods, d_c, d_u, d_nf, d_fl,
                     dev + o_ws, device_bytes - o_ws);
  if (rc) return rc;
  cudaStream_t cs = (cudaStream_t)compute_stream, xs = (cudaStream_t)copy_stream;
  cudaEvent_t ev[2 * 64 + 2];
  int nev = 0;
  auto new_event = [&](cudaEvent_t& e) -> cudaError_t { cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming); if (r == cudaSuccess) ev[nev++] = e; return r; };
  cudaError_t e = cudaSuccess;
#define PCK(call, what) if ((e = (call)) != cudaSuccess) { rc = cuda_fail(e, what); break; }
  do {
    cudaEvent_t start, roots;
    PCK(new_event(start), "event");
    PCK(cudaMemsetAsync(dev + o_ws, 0, kHdrBytes, cs), "memset");
    PCK(cudaEventRecord(start, cs), "record");         // earlier work on the compute stream may still read the buffers
    PCK(cudaStreamWaitEvent(xs, start, 0), "wait");
    const int ks = k_split(pl);
    if (from_params) {
      // stage 1': one copy of the parameter vectors, model assembly on the device, preparation + first-period search
      if (P > 0) PCK(cudaMemcpyAsync(d_par, params, np_, cudaMemcpyHostToDevice, cs), "H2D params");
      if (has_hybrid(tmpl)) build_stacks_kernel<true><<<(unsigned)((M * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, cs>>>(*tmpl, n_models, d_par, n_layers_max, d_lay, d_n);
      else build_stacks_kernel<false><<<(unsigned)((M * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, cs>>>(*tmpl, n_models, d_par, n_layers_max, d_lay, d_n);
      PCK(cudaGetLastError(), "build_stacks");
      if ((rc = stage_prep(pl, 0, n_models, cs))) break;
      if ((rc = stage_p1(pl, 0, n_models, 0, ks, cs))) break;
    } else {
      // stage 1: host->device copy of chunk i under preparation + first-period root search of chunk i-1
      for (int i = 0; i < n_chunks && rc == 0; ++i) {
        const int a = (int)((long long)n_models * i / n_chunks), b = (int)((long long)n_models * (i + 1) / n_chunks);
        const size_t cnt = (size_t)(b - a);
        for (int comp = 0; comp < 5 && e == cudaSuccess; ++comp)
          e = cudaMemcpyAsync(d_lay + comp * M * L + (size_t)a * L, layers + comp * M * L + (size_t)a * L, cnt * L * sizeof(float),
                              cudaMemcpyHostToDevice, xs);
        if (e != cudaSuccess) { rc = cuda_fail(e, "H2D layers"); break; }
        PCK(cudaMemcpyAsync(d_n + a, n_layers + a, cnt * sizeof(int), cudaMemcpyHostToDevice, xs), "H2D nlay");
        cudaEvent_t up;
        PCK(new_event(up), "event");
        PCK(cudaEventRecord(up, xs), "record");
        PCK(cudaStreamWaitEvent(cs, up, 0), "wait");
        if ((rc = stage_prep(pl, a, b, cs))) break;
        if ((rc = stage_p1(pl, a, b, 0, ks, cs))) break;
      }
    }
    if (rc) break;
    // stage 2: the later periods on the whole batch (one persistent launch: its tail is paid once)
    if ((rc = stage_p1(pl, 0, n_models, ks, n_periods, cs))) break;
    PCK(new_event(roots), "event");
    PCK(cudaEventRecord(roots, cs), "record");
    PCK(cudaStreamWaitEvent(xs, roots, 0), "wait");
    PCK(cudaMemcpyAsync(nfound, d_nf, ni, cudaMemcpyDeviceToHost, xs), "D2H nfound");
    if (flags) PCK(cudaMemcpyAsync(flags, d_fl, ni, cudaMemcpyDeviceToHost, xs), "D2H flags");
    PCK(cudaMemcpyAsync(c_out, d_c, no, cudaMemcpyDeviceToHost, xs), "D2H c");
    // stage 3: group velocities chunk by chunk, each chunk copied out under the next one
    if (u_out) {
      for (int i = 0; i < n_chunks && rc == 0; ++i) {
        const int a = (int)((long long)n_models * i / n_chunks), b = (int)((long long)n_models * (i + 1) / n_chunks);
        if ((rc = stage_p2(pl, a, b, cs))) break;
        cudaEvent_t done;
        PCK(new_event(done), "event");
        PCK(cudaEventRecord(done, cs), "record");
        PCK(cudaStreamWaitEvent(xs, done, 0), "wait");
        PCK(cudaMemcpyAsync(u_out + (size_t)a * K, d_u + (size_t)a * K, (size_t)(b - a) * K * sizeof(float), cudaMemcpyDeviceToHost, xs),
            "D2H u");
      }
      if (rc) break;
    }
  } while (0);
#undef PCK
  // results are valid on return
  cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(xs);
  for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
  if (rc == 0 && e1 != cudaSuccess) rc = cuda_fail(e1, "sync compute");
  if (rc == 0 && e2 != cudaSuccess) rc = cuda_fail(e2, "sync copy");
  return rc;
}

int surfdisp_host_batch_pipelined(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max,
                                  const int* n_layers, const float* layers, int n_periods, const float* periods,
                                  float* c_out, float* u_out, int* nfound, int* flags, void* device_buffer,
                                  size_t device_bytes, int n_chunks, void* compute_stream, void* copy_stream) {
  return host_pipeline(opts, kind, n_models, n_layers_max, n_layers, layers, nullptr, nullptr, n_periods, periods, c_out, u_out,
                       nfound, flags, device_buffer, device_bytes, n_chunks, compute_stream, copy_stream);
}

int surfdisp_host_params_pipelined(const SurfdispOpts* opts, const SurfdispStackTemplate* tmpl, int kind, int n_models,
                                   int n_layers_max, const float* params, int n_periods, const float* periods,
                                   float* c_out, float* u_out, int* nfound, int* flags, void* device_buffer,
                                   size_t device_bytes, int n_chunks, void* compute_stream, void* copy_stream) {
  if (int rc = check_template(tmpl)) return rc;
  return host_pipeline(opts, kind, n_models, n_layers_max, nullptr, nullptr, tmpl, params, n_periods, periods, c_out, u_out,
                       nfound, flags, device_buffer, device_bytes, n_chunks, compute_stream, copy_stream);
}

int surfdisp_misfit_batch(int mode, int n_models, int n_periods, const float* c_pred, const int* nfound,
                          const float* obs, const float* sigma, const unsigned char* mask,
                          const float* periods, float* out, void* stream) {
  if ((mode != 0 && mode != 1) || n_models < 0 || n_periods < 1 || n_periods > kMaxPer || !obs || !sigma)
    return SURFDISP_EINVAL;
  if (mode == 1 && !periods) return SURFDISP_EINVAL;
  if (n_models == 0) return 0;
  if (!c_pred || !nfound || !out) return SURFDISP_EINVAL;
  MisfitParams p;
  memset(&p, 0, sizeof(p));
  p.mode = mode; p.M = n_models; p.K = n_periods; p.c_pred = c_pred; p.nfound = nfound; p.out = out;
  for (int k = 0; k < n_periods; ++k) {
    p.obs[k] = obs[k];
    p.isig[k] = 1.0f / sigma[k];
    p.per[k] = periods ? periods[k] : 0.f;
    p.use[k] = mask ? (mask[k] != 0) : 1;
    p.ncount += p.use[k];
  }
  if (p.ncount == 0) return SURFDISP_EINVAL;  // point.py:356 raises when everything is masked
  const int threads = 256;
  const long long warps = n_models;
  const unsigned grid = (unsigned)((warps * 32 + threads - 1) / threads);
  misfit_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(p);
  CK(cudaGetLastError());
  return 0;
}

// Per-thread context of the host entry points (surfdisp_host_batch, fast_surf_): the device block and the two streams
// are kept between calls -- an unmodified models.py:27 loop calls fast_surf 50 000 times per point, and a
// cudaMalloc + two cudaStreamCreate + cudaFree per call cost more than the solve.  One context per host thread and
// device (the entry points stay re-entrant: no state is shared between threads); the block only grows.
struct HostCtx {
  int device = -1;
  char* dev = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr, xs = nullptr;
  ~HostCtx() { release(); }
  void release() {
    if (device < 0) return;
    // (at thread exit the CUDA context may already be gone: errors are ignored)
    if (dev) cudaFree(dev);
    if (st) cudaStreamDestroy(st);
    if (xs) cudaStreamDestroy(xs);
    dev = nullptr; st = xs = nullptr; bytes = 0; device = -1;
    cudaGetLastError();
  }
  int acquire(int dev_id, size_t need) {
    if (device != dev_id) release();
    CK(cudaSetDevice(dev_id));
    if (device < 0) {
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithFlags(&xs, cudaStreamNonBlocking));
      device = dev_id;
    }
    if (bytes < need) {
      if (dev) { cudaFree(dev); dev = nullptr; bytes = 0; }
      const size_t want = need < (size_t)(1 << 20) ? (size_t)(1 << 20) : need;
      CK(cudaMalloc(&dev, want));
      bytes = want;
    }
    return 0;
  }
};
static thread_local HostCtx g_host_ctx;

int surfdisp_host_batch(const SurfdispOpts* opts, int device, int kind, int n_models, int n_layers_max,
                        const int* n_layers, const float* layers, int n_periods, const float* periods,
                        float* c_out, float* u_out, int* nfound, int* flags) {
  if (n_models < 0 || n_layers_max < 2 || n_periods < 1 || n_periods > kMaxPer) return SURFDISP_EINVAL;
  if (n_models == 0) return 0;
  if (!n_layers || !layers || !c_out || !nfound || !periods) return SURFDISP_EINVAL;
  const size_t total = surfdisp_pipelined_bytes(n_models, n_layers_max, n_periods);
  if (int rc = g_host_ctx.acquire(device, total)) return rc;
  const int chunks = n_models >= (1 << 16) ? 8 : 1;
  const int rc = surfdisp_host_batch_pipelined(opts, kind, n_models, n_layers_max, n_layers, layers, n_periods, periods, c_out,
                                               u_out, nfound, flags, g_host_ctx.dev, g_host_ctx.bytes, chunks, g_host_ctx.st,
                                               g_host_ctx.xs);
  // a batch that needed more than 1 GiB does not keep it
  if (g_host_ctx.bytes > ((size_t)1 << 30)) { cudaFree(g_host_ctx.dev); g_host_ctx.dev = nullptr; g_host_ctx.bytes = 0; }
  return rc;
}

void surfdisp_host_release(void) { g_host_ctx.release(); }

void surfdisp_set_split_min_models(int n_models) { g_split_min_models = (n_models > 0) ? n_models : kSplitMinDefault; }

void fast_surf_(const int* n_layer0, const int* kind0, const float* a_ref0, const float* b_ref0,
                const float* rho_ref0, const float* d_ref0, const float* qs_ref0, const float* cvper,
                const int* ncvper, float* uR0, float* uL0, float* cR0, float* cL0) {
  for (int i = 0; i < kMaxPer; ++i) { uR0[i] = 0.f; uL0[i] = 0.f; cR0[i] = 0.f; cL0[i] = 0.f; }
  const int n = *n_layer0, kind = *kind0;
  int K = *ncvper;
  if (K > kMaxPer) K = kMaxPer;  // init.f:63-66
  if (n < 2 || n > SURFDISP_MAX_LAYERS || K < 1 || (kind != 1 && kind != 2)) return;
  float* lay = (float*)malloc((size_t)5 * n * sizeof(float));
  if (!lay) return;
  memcpy(lay + 0 * n, a_ref0, n * sizeof(float));
  memcpy(lay + 1 * n, b_ref0, n * sizeof(float));
  memcpy(lay + 2 * n, rho_ref0, n * sizeof(float));
  memcpy(lay + 3 * n, d_ref0, n * sizeof(float));
  memcpy(lay + 4 * n, qs_ref0, n * sizeof(float));
  float c[kMaxPer], u[kMaxPer];
  int nf = 0, fl = 0, dev = 0;
  cudaGetDevice(&dev);
  int rc = surfdisp_host_batch(nullptr, dev, kind, 1, n, &n, lay, K, cvper, c, u, &nf, &fl);
  free(lay);
  if (rc != 0) return;
  for (int i = 0; i < nf && i < K; ++i) {  // fast_surf.f:197-208
    if (kind == 1) { cL0[i] = c[i]; uL0[i] = u[i]; }
    else { cR0[i] = c[i]; uR0[i] = u[i]; }
  }
}

int surfdisp_read_counters(const void* workspace, unsigned long long out[4], void* stream) {
  if (!workspace || !out) return SURFDISP_EINVAL;
  CK(cudaMemcpyAsync(out, workspace, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}


int surfdisp_batch_profiled(const SurfdispOpts* opts, int kind, int n_models, int n_layers_max, const int* n_layers,
                            const float* layers, int n_periods, const float* periods, float* c_out, float* u_out,
                            int* nfound, int* flags, void* workspace, size_t workspace_bytes, void* stream,
                            float kernel_ms[3]) {
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&ev[i]));
  g_prof_events = ev;
  int rc = surfdisp_batch(opts, kind, n_models, n_layers_max, n_layers, layers, n_periods, periods, c_out, u_out,
                          nfound, flags, workspace, workspace_bytes, stream);
  g_prof_events = nullptr;
  if (rc == 0) {
    cudaError_t e = cudaEventSynchronize(ev[3]);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaEventSynchronize");
    for (int i = 0; i < 3 && rc == 0; ++i) {
      kernel_ms[i] = 0.f;
      cudaEventElapsedTime(&kernel_ms[i], ev[i], ev[i + 1]);
    }
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

int surfdisp_measure_peaks(double out[3]) {
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  void* buf = nullptr;
  CK(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0.f;
  for (int which = 0; which < 3; ++which) {
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaEventRecord(e0, 0));
      if (which == 0) fma_peak_kernel<float><<<blocks, threads>>>((float*)buf, iters);
      else if (which == 1) fma_peak_kernel<double><<<blocks, threads>>>((double*)buf, iters / 4);
      else mufu_peak_kernel<<<blocks, threads>>>((float*)buf, iters);
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double n = (double)blocks * threads * (which == 1 ? iters / 4 : iters) * 8.0 * (which == 2 ? 4.0 : 8.0);
      const double rate = (which == 2 ? n : 2.0 * n) / (ms * 1e-3) * 1e-12;  // TFLOP/s, or T-ex2/s
      if (rep > 0 && rate > best) best = rate;
    }
    out[which] = best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  return 0;
}


int surfdisp_build_stacks(const SurfdispStackTemplate* tmpl, int n_models, const float* params, int n_layers_max,
                          float* layers, int* n_layers, void* stream) {
  if (n_models < 0 || n_layers_max < 2 || n_layers_max > SURFDISP_MAX_LAYERS) return SURFDISP_EINVAL;
  if (int rc = check_template(tmpl)) return rc;
  if (n_models == 0) return 0;
  if (!layers || !n_layers || (tmpl->nparams > 0 && !params)) return SURFDISP_EINVAL;
  if (has_hybrid(tmpl)) build_stacks_kernel<true><<<(unsigned)(((size_t)n_models * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, (cudaStream_t)stream>>>(*tmpl, n_models, params, n_layers_max, layers, n_layers);
  else build_stacks_kernel<false><<<(unsigned)(((size_t)n_models * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, (cudaStream_t)stream>>>(*tmpl, n_models, params, n_layers_max, layers, n_layers);
  CK(cudaGetLastError());
  return 0;
}

int surfdisp_check_priors(const SurfdispStackTemplate* tmpl, int n_models, const float* params, int* priors, void* stream) {
  if (n_models < 0) return SURFDISP_EINVAL;
  if (int rc = check_template(tmpl)) return rc;
  if (n_models == 0) return 0;
  if (!priors || (tmpl->nparams > 0 && !params)) return SURFDISP_EINVAL;
  if (has_hybrid(tmpl)) check_priors_kernel<true><<<(unsigned)(((size_t)n_models * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, (cudaStream_t)stream>>>(*tmpl, n_models, params, priors);
  else check_priors_kernel<false><<<(unsigned)(((size_t)n_models * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, (cudaStream_t)stream>>>(*tmpl, n_models, params, priors);
  CK(cudaGetLastError());
  return 0;
}

int surfdisp_mc_propose(const SurfdispStackTemplate* tmpl, int n_chains, const float* lo, const float* hi,
                        const float* step, const float* cur, const unsigned char* reset_mask, float* prop,
                        int* status, unsigned long long seed, unsigned int step_index, void* stream) {
  if (n_chains < 0) return SURFDISP_EINVAL;
  if (int rc = check_template(tmpl)) return rc;
  const int P = tmpl->nparams;
  if (P < 1 || P > kMaxParams || !lo || !hi || !step) return SURFDISP_EINVAL;
  if (n_chains == 0) return 0;
  if (!cur || !prop) return SURFDISP_EINVAL;
  McBounds bd;
  memset(&bd, 0, sizeof(bd));
  for (int i = 0; i < P; ++i) {
    if (!(hi[i] > lo[i]) || !(step[i] > 0.f)) return SURFDISP_EINVAL;
    bd.lo[i] = lo[i]; bd.hi[i] = hi[i]; bd.step[i] = step[i];
  }
  if (has_hybrid(tmpl))
    mc_propose_kernel<true><<<(unsigned)(((size_t)n_chains * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, (cudaStream_t)stream>>>(
        *tmpl, bd, n_chains, cur, reset_mask, prop, status, seed, step_index);
  else
    mc_propose_kernel<false><<<(unsigned)(((size_t)n_chains * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, (cudaStream_t)stream>>>(
        *tmpl, bd, n_chains, cur, reset_mask, prop, status, seed, step_index);
  CK(cudaGetLastError());
  return 0;
}

int surfdisp_mc_accept(int n_chains, int n_params, const float* chi1, const float* prop, float* chi0, float* cur,
                       const unsigned char* force_mask, unsigned char* accepted, unsigned long long seed,
                       unsigned int step_index, void* stream) {
  if (n_chains < 0 || n_params < 1) return SURFDISP_EINVAL;
  if (n_chains == 0) return 0;
  if (!chi1 || !prop || !chi0 || !cur || !accepted) return SURFDISP_EINVAL;
  mc_accept_kernel<<<(n_chains + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n_chains, n_params, chi1, prop, chi0, cur,
                                                                             force_mask, accepted, seed, step_index);
  CK(cudaGetLastError());
  return 0;
}

int surfdisp_mc_step(const SurfdispOpts* opts, const SurfdispStackTemplate* tmpl, const SurfdispMcState* s,
                     const float* periods, void* stream) {
  if (!s || !periods) return SURFDISP_EINVAL;
  if (int rc = check_template(tmpl)) return rc;
  const int M = s->n_chains, P = s->n_params, K = s->n_periods;
  if (M < 0 || P < 1 || P > kMaxParams || P != tmpl->nparams || K < 1 || K > kMaxPer || s->chains_per_point < 1 ||
      (s->misfit_mode != 0 && s->misfit_mode != 1) || s->track_steps < 0)
    return SURFDISP_EINVAL;
  if (M == 0) return 0;
  if (!s->cur || !s->prop || !s->chi0 || !s->status || !s->accepted || !s->step || !s->bounds || !s->obs || !s->isig ||
      !s->use || !s->layers || !s->n_layers || !s->c_pred || !s->nfound || !s->workspace || (s->track_steps > 0 && !s->track))
    return SURFDISP_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  // ---- proposal + model assembly
  McStepParams mp;
  memset(&mp, 0, sizeof(mp));
  mp.M = M; mp.P = P; mp.lmax = s->n_layers_max; mp.chain_len = s->chain_len; mp.cur = s->cur; mp.prop = s->prop;
  mp.status = s->status; mp.init_mask = s->init_mask; mp.layers = s->layers; mp.nlay = s->n_layers; mp.step_ptr = s->step;
  mp.seed = s->seed; mp.bounds = reinterpret_cast<const McBounds*>(s->bounds); mp.chains_per_point = s->chains_per_point;
  if (has_hybrid(tmpl)) mc_propose_build_kernel<true><<<(unsigned)(((size_t)M * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, st>>>(*tmpl, mp);
  else mc_propose_build_kernel<false><<<(unsigned)(((size_t)M * 32 + kMcThreads - 1) / kMcThreads), kMcThreads, 0, st>>>(*tmpl, mp);
  CK(cudaGetLastError());
  // ---- phase velocities of the proposals
  SurfdispOpts o;
  if (opts) o = *opts; else surfdisp_default_opts(&o);
  o.compute_group = 0;
  int rc = surfdisp_batch_hinted(&o, s->kind, M, s->n_layers_max, s->n_layers, s->layers, K, periods, s->c_cur, s->c_pred, nullptr,
                                 s->nfound, s->flags, s->workspace, s->workspace_bytes, stream);
  if (rc) return rc;
  // ---- misfit, Metropolis rule, state update, track row
  McFinishParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.M = M; fp.P = P; fp.K = K; fp.mode = s->misfit_mode; fp.chain_len = s->chain_len; fp.chains_per_point = s->chains_per_point;
  fp.track_steps = s->track_steps > 0 ? s->track_steps : 1; fp.c_pred = s->c_pred; fp.nfound = s->nfound; fp.status = s->status;
  fp.obs = s->obs; fp.isig = s->isig; fp.use = s->use; fp.prop = s->prop; fp.cur = s->cur; fp.chi0 = s->chi0;
  fp.accepted = s->accepted; fp.misfit_out = s->misfit; fp.c_cur = s->c_cur; fp.track = s->track_steps > 0 ? s->track : nullptr; fp.step_ptr = s->step;
  fp.seed = s->seed;
  for (int k = 0; k < K; ++k) fp.per[k] = periods[k];
  mc_finish_kernel<<<(M + 127) / 128, 128, 0, st>>>(fp);
  mc_bump_kernel<<<1, 1, 0, st>>>(s->step);
  CK(cudaGetLastError());
  return 0;
}

const char* surfdisp_version(void) { return "surfdisp_b200 0.2 (sm_100a)"; }
const char* surfdisp_last_cuda_error(void) { return g_cuda_err; }

}  // extern "C"
