// surfdisp_core.cuh -- per-lane arithmetic of the batched dispersion solver.
//
// Everything here is written from the physics/algorithm of the reference hot path
// (fast_surf_src/{flat1,calcul,surfa}.f of 001cat/pySurfInv; line numbers cited per function), laid
// out for one GPU lane evaluating a pair of trial phase velocities (or one period) against a layer stack
// staged in shared memory.  The functions are __host__ __device__ so that tests/hostmirror can compile
// the same source with g++ and check the math against the CPU oracle without a GPU; the product only
// ever runs the __device__ instantiation (see surfdisp_kernels.cu).
#pragma once
#include <math.h>
#include <stdint.h>
#include "sd_libm.cuh"

#if defined(__CUDACC__)
#define SD_HD __host__ __device__ __forceinline__
#else
#define SD_HD inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
#endif

// Exactly-rounded float32 basic ops without FMA contraction.  The model-preparation stage (attenuation
// + earth flattening) suffers catastrophic cancellation in float32 (relative 1e-4 per layer), so the
// reference's results are only reproducible if these roundings happen in the reference's order.
#if defined(__CUDA_ARCH__)
#define SD_ADD(a, b) __fadd_rn((a), (b))
#define SD_SUB(a, b) __fsub_rn((a), (b))
#define SD_MUL(a, b) __fmul_rn((a), (b))
#define SD_DIV(a, b) __fdiv_rn((a), (b))
#else
static inline float sd_add(float a, float b) { volatile float r = a + b; return r; }
static inline float sd_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float sd_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float sd_div(float a, float b) { volatile float r = a / b; return r; }
#define SD_ADD(a, b) sd_add((a), (b))
#define SD_SUB(a, b) sd_sub((a), (b))
#define SD_MUL(a, b) sd_mul((a), (b))
#define SD_DIV(a, b) sd_div((a), (b))
#endif

// Division where the last ulp does not matter (interpolation weights, the small attenuation terms): one
// MUFU.RCP + multiply on the device.
#if defined(__CUDA_ARCH__)
#define SD_FDIV(a, b) __fdividef((a), (b))
#else
#define SD_FDIV(a, b) ((a) / (b))
#endif

namespace sd {

// float32 literals of the reference
#define SD_PI_ATT 3.1415927f      /* calcul.f:32 */
#define SD_TWOPI 6.2831855f       /* 6.28318531 / 6.2831853 / 6.2831853072 all round to this float */
#define SD_R0 6371.0f             /* flat1.f:21 */

// Per-model, per-layer constants written once by the prep kernel (period independent part of
// calcul.f:112-133 + flat1.f).  One row of NCONST arrays per model.
enum { C_AREF = 0, C_BREF = 1, C_QS = 2, C_DIF = 3, C_RHOFL = 4, C_DFL = 5, C_HSF = 6, C_RHOHS = 7, NCONST = 8 };

// Working layer record in shared memory: one float4 (a, b, rho, d) per layer for the current period
// (attenuation-corrected, flattened).  b == 0 flags a liquid layer.  The reciprocals the layer step needs
// (1/a^2, 1/b^2, 1/rho) are MUFU.RCP results formed on the fly: 16 bytes per layer instead of 32 lets twice
// as many models stay resident per SM.

// ----------------------------------------------------------------------------------------------
// Model preparation.  flat1.f:33-69 evaluated once per model: radii by sequential float32 prefix sum,
// velocity factor dif(i), density factor, flattened thickness, plus the factors layer i would get if it
// were the (effective) half-space.  logs and powers follow glibc's logf / powf bit by bit (sd_libm.cuh):
// r_i**p - r_(i+1)**p cancels 3-4 digits, so even a 1-ulp difference in a power shows up as 1e-5 km/s in c.
// Arrays are indexed [0..n-1]; out rows have stride ld.
SD_HD float sd_logf_cr(float x) { return sdm::logf_glibc(x); }
SD_HD float sd_powf_cr(float x, float p) { return sdm::powf_glibc(x, p); }

SD_HD void prep_model(int n, int kind, int flatten, const float* a, const float* b, const float* rho,
                      const float* d, const float* qs, float* out, int ld) {
  const float A = SD_R0;
  const float pwr = (kind == 1) ? 5.0f : 2.2750f;
  const float apw = sd_powf_cr(A, pwr);
  float hs = 0.f, z0 = 0.f;
  float r_i = A;  // radius of the top of layer i
  for (int i = 0; i < n; ++i) {
    float ht = hs;
    hs = SD_ADD(hs, d[i]);
    r_i = SD_SUB(A, ht);
    float r_n = SD_SUB(A, hs);  // radius of the top of layer i+1
    out[C_AREF * ld + i] = a[i];
    out[C_BREF * ld + i] = b[i];
    out[C_QS * ld + i] = qs[i];
    if (!flatten) {
      out[C_DIF * ld + i] = 1.f; out[C_RHOFL * ld + i] = rho[i]; out[C_DFL * ld + i] = d[i];
      out[C_HSF * ld + i] = 1.f; out[C_RHOHS * ld + i] = rho[i];
      continue;
    }
    // as a regular layer (flat1.f:41-56), needs the next boundary
    float fltd = sd_logf_cr(SD_DIV(r_i, r_n));
    float dif = SD_DIV(SD_MUL(SD_SUB(SD_DIV(1.0f, r_n), SD_DIV(1.0f, r_i)), A), fltd);
    float difr = SD_SUB(sd_powf_cr(r_i, pwr), sd_powf_cr(r_n, pwr));
    float qqq = SD_DIV(difr, SD_MUL(SD_MUL(fltd, apw), pwr));
    out[C_DIF * ld + i] = dif;
    out[C_RHOFL * ld + i] = SD_MUL(rho[i], qqq);
    // flattened thickness of layer i: z(i+1) - z(i)  (flat1.f:65-68)
    float z1 = SD_MUL(A, sd_logf_cr(SD_DIV(A, r_n)));
    out[C_DFL * ld + i] = SD_SUB(z1, z0);
    z0 = z1;
    // as the half-space (flat1.f:58-62)
    float fct = SD_DIV(A, r_i);
    out[C_HSF * ld + i] = fct;
    out[C_RHOHS * ld + i] = SD_MUL(rho[i], sd_powf_cr(SD_DIV(1.0f, fct), pwr));
  }
}

// Upper bound of every running thickness sum the layer-dropping walks can form (surfa.f:92-106, 854-866: float32
// additions, in layer order, of the flattened thicknesses of a subset of the layers above the half-space; every
// addition rounds by at most 2^-24 and sub-layer thicknesses RN(d / ndiv) add up to d (1 + 2^-24) at most, so the
// real sum with a margin of 1e-4 covers 1600 additions).  A walk whose limit fact*c*T is not below the bound drops
// nothing; both phases skip it then.  -1 = unknown.  (prep_kernel forms the same sum with a warp reduction.)
SD_HD float stack_depth_bound(const float* cst, int ld, int n) {
  double s = 0.0;
  for (int i = 0; i < n - 1; ++i) { const float d = cst[C_DFL * ld + i]; s += (double)(d > 0.f ? d : 0.f); }
  if (!(s == s)) return -1.f;
  float f = (float)(s * 1.0001);
  return ((double)f < s * 1.00005) ? -1.f : f;   // (never: the conversion rounds by 6e-8)
}

// Attenuation-corrected, flattened (a, b) of layer i for the period whose log term is lt = ln(t_base/T)
// (calcul.f:121-127 then flat1 scaling).  as_half selects the half-space flattening factor.
SD_HD void layer_ab_vals(float ar, float br, float qs, float f, float lt, int atten, float& a, float& b) {
  a = ar; b = br;
  if (atten) {
    // qsq, qpq ~ 1e-2: an ulp of them is 1e-9 of the velocity, far below float32 resolution
    const float qsq = SD_FDIV(SD_MUL(qs, lt), SD_PI_ATT);
    const float qpq = SD_FDIV(SD_MUL(SD_MUL(qsq, 1.33333333f), SD_MUL(br, br)), SD_MUL(ar, ar));
    b = SD_MUL(br, SD_ADD(1.0f, qsq));
    a = SD_MUL(ar, SD_ADD(1.0f, qpq));
  }
  a = SD_MUL(a, f);
  b = SD_MUL(b, f);
}

SD_HD void layer_ab(const float* cst, int ld, int i, float lt, int atten, bool as_half, float& a, float& b) {
  const float f = as_half ? cst[C_HSF * ld + i] : cst[C_DIF * ld + i];
  layer_ab_vals(cst[C_AREF * ld + i], cst[C_BREF * ld + i], cst[C_QS * ld + i], f, lt, atten, a, b);
}

SD_HD float4 make_rec(float a, float b, float rho, float d) { return make_float4(a, b, rho, d); }

// ----------------------------------------------------------------------------------------------
// Layer dropping, surfa.f:92-106.  Returns mmax (1-based count of layers kept, >= 2).
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
int layer_drop(float c, float T, float fact, int nmax, const float4* rec) {
  const float dmax = SD_MUL(SD_MUL(fact, c), T);
  float sum = 0.f;
  int mmax = nmax;
  // four layers per shared-memory round trip (the loop is latency bound: load -> compare -> add -> compare);
  // the additions stay in the reference's order
  int ii = 0;
  for (; ii + 4 <= nmax; ii += 4) {
    const float4 e0 = rec[ii], e1 = rec[ii + 1], e2 = rec[ii + 2], e3 = rec[ii + 3];
    if (c < e0.y) { sum = SD_ADD(sum, e0.w); if (sum > dmax) { mmax = ii + 1; goto done; } }
    if (c < e1.y) { sum = SD_ADD(sum, e1.w); if (sum > dmax) { mmax = ii + 2; goto done; } }
    if (c < e2.y) { sum = SD_ADD(sum, e2.w); if (sum > dmax) { mmax = ii + 3; goto done; } }
    if (c < e3.y) { sum = SD_ADD(sum, e3.w); if (sum > dmax) { mmax = ii + 4; goto done; } }
  }
  for (; ii < nmax; ++ii) {
    const float4 e = rec[ii];
    if (c < e.y) {
      sum = SD_ADD(sum, e.w);
      if (sum > dmax) { mmax = ii + 1; break; }
    }
  }
done:
  return mmax < 2 ? 2 : mmax;
}

// ----------------------------------------------------------------------------------------------
// Fast scalar primitives of the secular-function inner loop.  On the device they map to single MUFU
// instructions (1-2 ulp): the secular function only has to be good to the float32 noise of the
// reference's own evaluation (root noise 1e-6 km/s, tolerance 1e-4).  On the host (tests/hostmirror)
// they are the exact libm functions.
#if defined(__CUDA_ARCH__)
SD_HD float sd_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SD_HD float sd_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SD_HD float sd_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
SD_HD void sd_sincos(float x, float& s, float& c) {
  // two-constant Cody-Waite reduction to [-pi, pi], then MUFU.SIN / MUFU.COS (abs err 2^-21.4 there)
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(-n, 6.28318548202514648f, x);
  r = fmaf(-n, -1.74845553146951715e-7f, r);
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(r));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(r));
}
#else
SD_HD float sd_rsqrt(float x) { return 1.0f / sqrtf(x); }
SD_HD float sd_rcp(float x) { return 1.0f / x; }
SD_HD float sd_ex2(float x) { return exp2f(x); }
SD_HD void sd_sincos(float x, float& s, float& c) { sincosf(x, &s, &c); }
#endif
#define SD_LOG2E 1.4426950408889634f
// block-placement hints: the common path of the sweep loop (solid layer, thin-tier series) should fall through
#define SD_LIKELY(x) __builtin_expect(!!(x), 1)
#define SD_UNLIKELY(x) __builtin_expect(!!(x), 0)

// One "half" of a layer propagator: for r = sqrt(|arg|) and x = kd*r returns
//   (r sin x, sin x / r, cos x)  when arg < 0 (oscillatory, c above the layer velocity)
//   (-r sinh x, sinh x / r, cosh x) when arg > 0 (evanescent)          surfa.f:262-288
//   (0, kd, 1) in the limit arg -> 0 (surfa.f:262-265, 275, 285-287)
// With u = kd^2 * arg (signed) all three are entire functions of u:
//   sin x / r = kd S(u),  r sin x = -kd arg S(u),  cos x = C(u),   S(u) = sum u^n/(2n+1)!,  C(u) = sum u^n/(2n)!
// For |u| < 0.5 (thin layers / long periods: the vast majority of layer steps) the series are used: no
// square root, no MUFU, no branch on the sign of arg, and none of the cancellation that
// (exp(x)-exp(-x))/2 suffers for small x (the reference's float32 form loses ~x^-1 ulps there).
// Thin tier: degree-3 / degree-4 minimax polynomials on |u| <= 0.5 (constant term exact); their float32 evaluation
// error (9e-8) is the rounding floor of the Horner scheme itself -- the Taylor polynomials need one more term each.
#define SD_S1 1.66666666e-1f
#define SD_S2 8.33389324e-3f
#define SD_S3 1.98420544e-4f
#define SD_C2 4.16666670e-2f
#define SD_C3 1.38898493e-3f
#define SD_C4 2.48008682e-5f
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
void half_terms(float arg, float kd, float kd2, float& rsin, float& sinr, float& cs) {
  const float u = kd2 * arg;
  if (fabsf(u) < 0.5f) {
    const float S = 1.f + u * (SD_S1 + u * (SD_S2 + u * SD_S3));
    cs = 1.f + u * (0.5f + u * (SD_C2 + u * (SD_C3 + u * SD_C4)));
    sinr = kd * S;
    rsin = -arg * sinr;
    return;
  }
  const float t = fabsf(arg);
  const float ir = sd_rsqrt(t);
  const float r = t * ir;
  const float x = kd * r;
  if (arg > 0.f) {
    const float xl = x * SD_LOG2E;
    const float ex = sd_ex2(xl), em = sd_ex2(-xl);
    const float sh = 0.5f * (ex - em);
    cs = 0.5f * (ex + em);
    rsin = -r * sh;
    sinr = sh * ir;
  } else {
    float sn, c_;
    sd_sincos(x, sn, c_);
    cs = c_;
    rsin = r * sn;
    sinr = sn * ir;
  }
}

// ----------------------------------------------------------------------------------------------
// Rayleigh secular function (surfa.f:193-357), adjoint form.  The reference propagates the Dunkin
// compound-matrix COLUMN vector top-down through layers 1..mmax-1 and contracts it with the half-space row h
// (surfa.f:341-354): the dispersion function and the two ellipticity sweeps are h^T P e1, h^T P e2 and
// h^T P e3 with the same layer product P = A(mmax-1) ... A(1).  Propagating the ROW vector upwards instead
// (r <- r A(m), m = mmax-1 .. 1) yields all three from one sweep: Delta = -r1, ellipticity = 0.5 r3 / r2
// (surfa.f:360-363), with (e2, e3) = (r2, r3) returned separately (the caller interpolates them to the root
// before dividing).  Same 25 multiply-adds per layer as the column form, but every trial velocity of the root
// search carries its ellipticity, so the reference's two extra sweeps per period disappear.
// The 15 matrix entries of surfa.f:289-320 are formed from shared sub-expressions
// (w = 2 g g1 (1-cc) + g^2 rr + g1^2 ss gives a11 = cc - w and a33 = 1 + 2w, etc.).
// A liquid top layer (surfa.f:219-251) closes the sweep: the ellipticity is taken below it (the reference
// skips liquid layers there, surfa.f:220), the dispersion function includes it.  Liquid layers deeper in the
// stack get dispersion-function semantics; for such stacks (e2, e3) must come from a second sweep with
// ell_only = true, which skips every liquid layer like the reference's ellipticity sweeps do.

// Layer records live in shared memory in the kernel: a 32-bit shared-window address and ld.shared spare the
// generic-address descriptor set-up that a plain pointer dereference costs on every layer step.
#if defined(__CUDA_ARCH__) && !defined(SD_REC_GENERIC)
struct RecLoader {
  unsigned base;
  __device__ __forceinline__ explicit RecLoader(const float4* rec) {
    unsigned b = (unsigned)__cvta_generic_to_shared(rec);
    // Passed through an identity shuffle: where the sweep has a single call site in a kernel, ptxas specialises it for
    // that kernel and re-derives the address from the thread index and the kernel parameters in EVERY layer step (17
    // instructions of 200: cheaper than a register in its cost model).  A shuffle result cannot be recomputed.
    unsigned lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    base = __shfl_sync(__activemask(), b, lane);
  }
  __device__ __forceinline__ float4 operator()(int m) const {
    float4 r;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(base + 16u * (unsigned)m));
    return r;
  }
};
#else
struct RecLoader {
  const float4* rec;
  SD_HD explicit RecLoader(const float4* r) : rec(r) {}
  SD_HD float4 operator()(int m) const { return rec[m]; }
};
#endif

// ----------------------------------------------------------------------------------------------
// Two trial velocities per lane.  sm_100 has packed FP32 arithmetic (fma/mul/add.rn.f32x2 -> FFMA2 / FMUL2 /
// FADD2): one issue slot for two lanes' worth of work, and the packed multiply / add also run at twice the
// scalar rate (tools/microbench/ffma2_bench.cu).  The secular-function loop is issue bound, and everything
// in it is elementwise in the trial velocity (the layer data are shared), so each lane carries a PAIR of
// velocities through the sweep.
// The pair travels as ONE 64-bit value (not a float2): the compiler front end splits a float2 into two scalars at
// every control-flow join and re-packs it before the next packed instruction, which cost ~40 register moves per
// layer step (a quarter of the sweep's instructions); a 64-bit carrier keeps the pair in an aligned register pair.
#if defined(__CUDACC__)
struct V2 { unsigned long long v; };
#if defined(__CUDA_ARCH__)
SD_HD V2 v2(float a, float b) { V2 r; r.v = ((unsigned long long)__float_as_uint(b) << 32) | (unsigned long long)__float_as_uint(a); return r; }
SD_HD float vx(V2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
SD_HD float vy(V2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
SD_HD V2 vmul(V2 a, V2 b) { V2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SD_HD V2 vadd(V2 a, V2 b) { V2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SD_HD V2 vfma(V2 a, V2 b, V2 c) { V2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else   // host pass of nvcc: never executed, only has to compile
SD_HD V2 v2(float a, float b) { V2 r; unsigned ua, ub; memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); r.v = ((unsigned long long)ub << 32) | ua; return r; }
SD_HD float vx(V2 a) { const unsigned u = (unsigned)(a.v & 0xffffffffull); float f; memcpy(&f, &u, 4); return f; }
SD_HD float vy(V2 a) { const unsigned u = (unsigned)(a.v >> 32); float f; memcpy(&f, &u, 4); return f; }
SD_HD V2 vmul(V2 a, V2 b) { return v2(vx(a) * vx(b), vy(a) * vy(b)); }
SD_HD V2 vadd(V2 a, V2 b) { return v2(vx(a) + vx(b), vy(a) + vy(b)); }
SD_HD V2 vfma(V2 a, V2 b, V2 c) { return v2(fmaf(vx(a), vx(b), vx(c)), fmaf(vy(a), vy(b), vy(c))); }
#endif
#else
struct V2 { float x, y; };   // the host mirror's plain pair
SD_HD V2 v2(float a, float b) { V2 r; r.x = a; r.y = b; return r; }
SD_HD float vx(V2 a) { return a.x; }
SD_HD float vy(V2 a) { return a.y; }
SD_HD V2 vmul(V2 a, V2 b) { return v2(a.x * b.x, a.y * b.y); }
SD_HD V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
SD_HD V2 vfma(V2 a, V2 b, V2 c) { return v2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
SD_HD V2 vs(float s) { return v2(s, s); }
SD_HD V2 vneg(V2 a) { return v2(-vx(a), -vy(a)); }
SD_HD V2 vsub(V2 a, V2 b) { return vfma(b, vs(-1.f), a); }

// half_terms for a pair, negation-free form (packed FP32 has no operand negation in PTX; an explicit sign flip of a
// pair costs two scalar instructions).  Inputs either (a, k2s) = (arg, (k d)^2) [NEG = false] or
// (-arg, -(k d)^2) [NEG = true]: u = k2s * a either way.  Outputs sinr = sin x / r, cs = cos x and p = a * sinr,
// i.e. r sin x for NEG = true and -r sin x for NEG = false (oscillatory and evanescent alike).  Two series tiers, taken only if both velocities qualify: |u| < 0.5 (thin
// layers, the vast majority) and |u| < 3 (thick layers at short periods / low trial velocities, e.g. the whole scan
// of the first period), where the entire functions S and C need 9 / 10 terms for float32 accuracy; beyond that
// each velocity goes through the MUFU-based scalar form.
template <bool NEG>
SD_HD void half_terms2_scalar(V2 a, V2 kd, V2 k2s, V2& p, V2& sinr, V2& cs) {
  // (temporaries: the out-of-line scalar form takes addresses, which must not pin the pair registers to memory)
  float r0, s0, c0, r1, s1, c1;
  half_terms(NEG ? -vx(a) : vx(a), vx(kd), NEG ? -vx(k2s) : vx(k2s), r0, s0, c0);
  half_terms(NEG ? -vy(a) : vy(a), vy(kd), NEG ? -vy(k2s) : vy(k2s), r1, s1, c1);
  p = NEG ? v2(r0, r1) : v2(-r0, -r1); sinr = v2(s0, s1); cs = v2(c0, c1);
}

template <bool NEG>
SD_HD void half_terms2(V2 narg, V2 kd, V2 nkd2, V2& rsin, V2& sinr, V2& cs) {
  const V2 u = vmul(nkd2, narg);
  const float um = fmaxf(fabsf(vx(u)), fabsf(vy(u)));
  if (SD_LIKELY(um < 0.5f)) {
    V2 S = vfma(u, vs(SD_S3), vs(SD_S2));
    S = vfma(u, S, vs(SD_S1));
    S = vfma(u, S, vs(1.f));
    V2 C = vfma(u, vs(SD_C4), vs(SD_C3));
    C = vfma(u, C, vs(SD_C2));
    C = vfma(u, C, vs(0.5f));
    cs = vfma(u, C, vs(1.f));
    sinr = vmul(kd, S);
    rsin = vmul(narg, sinr);
    return;
  }
  if (um < 3.0f) {
    // S(u) = sum u^n / (2n+1)!,  C(u) = sum u^n / (2n)!;  |u|^10/21! < 2e-15, |u|^10/20! < 3e-14 relative
    V2 S = vfma(u, vs(8.2206352e-18f), vs(2.8114573e-15f));   // 1/19!, 1/17!
    S = vfma(u, S, vs(7.6471637e-13f));                         // 1/15!
    S = vfma(u, S, vs(1.6059044e-10f));                         // 1/13!
    S = vfma(u, S, vs(2.5052108e-8f));                          // 1/11!
    S = vfma(u, S, vs(2.7557319e-6f));                          // 1/9!
    S = vfma(u, S, vs(1.9841270e-4f));                          // 1/7!
    S = vfma(u, S, vs(8.3333333e-3f));                          // 1/5!
    S = vfma(u, S, vs(1.6666667e-1f));                          // 1/3!
    S = vfma(u, S, vs(1.f));
    V2 C = vfma(u, vs(1.5619207e-16f), vs(4.7794773e-14f));   // 1/18!, 1/16!
    C = vfma(u, C, vs(1.1470746e-11f));                         // 1/14!
    C = vfma(u, C, vs(2.0876757e-9f));                          // 1/12!
    C = vfma(u, C, vs(2.7557319e-7f));                          // 1/10!
    C = vfma(u, C, vs(2.4801587e-5f));                          // 1/8!
    C = vfma(u, C, vs(1.3888889e-3f));                          // 1/6!
    C = vfma(u, C, vs(4.1666667e-2f));                          // 1/4!
    C = vfma(u, C, vs(0.5f));
    cs = vfma(u, C, vs(1.f));
    sinr = vmul(kd, S);
    rsin = vmul(narg, sinr);
    return;
  }
  half_terms2_scalar<NEG>(narg, kd, nkd2, rsin, sinr, cs);
}

// Half-space row of the Rayleigh secular function (surfa.f:341-354) for one velocity; R = (a, b, rho, d)
SD_HD void rayleigh_hs_row(float csq, float icsq, const float4 R, float& r1, float& r2, float& r3, float& r4, float& r5) {
  const float pp = R.x, b2 = 2.0f * R.y * R.y;
  const float ippp = SD_FDIV(1.0f, pp * pp);
  const float arga = 1.0f - csq * ippp;
  const float argb = 1.0f - SD_FDIV(csq, R.y * R.y);
  float ra = sqrtf(fabsf(arga)); if (arga > 0.f) ra = -ra;
  float rb = sqrtf(fabsf(argb)); if (argb > 0.f) rb = -rb;
  const float g = b2 * icsq;
  const float g1 = g - 1.0f;
  const float sss = 0.5f * b2;
  const float rhp = R.z * pp;
  const float igra = SD_FDIV(1.0f, g * ra);
  const float g1s = g1 * g1;
  const float rba = rb - SD_FDIV(1.0f, ra);
  const float ih12 = SD_FDIV(1.0f, rhp * pp);
  r1 = -2.f * rb * sss * ippp + csq * g1s * ippp * igra;
  r3 = 2.f * (-rb * ih12 + g1 * ih12 * igra);
  r4 = rb * ih12 * igra;
  r5 = SD_FDIV(rba, rhp * rhp * csq * g);
  r2 = -SD_FDIV(ih12, g);
}

// P and SV series of one Rayleigh layer step together (u = nkd2 * narg per term): the four Horner chains are
// interleaved in one basic block -- no branch between them and four independent dependency chains for the
// scheduler.  thin: all four |u| < 0.5;  thick: all four |u| < 3.
SD_HD void series_thin2x2(V2 up, V2 uq, V2 kd, V2& Sp_, V2& Cp_, V2& Sq_, V2& Cq_) {
  V2 Sp = vfma(up, vs(SD_S3), vs(SD_S2)), Sq = vfma(uq, vs(SD_S3), vs(SD_S2));
  V2 Cp = vfma(up, vs(SD_C4), vs(SD_C3)), Cq = vfma(uq, vs(SD_C4), vs(SD_C3));
  Sp = vfma(up, Sp, vs(SD_S1)); Sq = vfma(uq, Sq, vs(SD_S1));
  Cp = vfma(up, Cp, vs(SD_C2)); Cq = vfma(uq, Cq, vs(SD_C2));
  Sp = vfma(up, Sp, vs(1.f)); Sq = vfma(uq, Sq, vs(1.f));
  Cp = vfma(up, Cp, vs(0.5f)); Cq = vfma(uq, Cq, vs(0.5f));
  Cp_ = vfma(up, Cp, vs(1.f)); Cq_ = vfma(uq, Cq, vs(1.f));
  Sp_ = vmul(kd, Sp); Sq_ = vmul(kd, Sq);
}

SD_HD void series_thick2x2(V2 up, V2 uq, V2 kd, V2& Sp_, V2& Cp_, V2& Sq_, V2& Cq_) {
  // S(u) = sum u^n / (2n+1)!,  C(u) = sum u^n / (2n)!;  |u|^10/21! < 2e-15, |u|^10/20! < 3e-14 relative
  V2 Sp = vfma(up, vs(8.2206352e-18f), vs(2.8114573e-15f)), Sq = vfma(uq, vs(8.2206352e-18f), vs(2.8114573e-15f));
  V2 Cp = vfma(up, vs(1.5619207e-16f), vs(4.7794773e-14f)), Cq = vfma(uq, vs(1.5619207e-16f), vs(4.7794773e-14f));
  Sp = vfma(up, Sp, vs(7.6471637e-13f)); Sq = vfma(uq, Sq, vs(7.6471637e-13f));
  Cp = vfma(up, Cp, vs(1.1470746e-11f)); Cq = vfma(uq, Cq, vs(1.1470746e-11f));
  Sp = vfma(up, Sp, vs(1.6059044e-10f)); Sq = vfma(uq, Sq, vs(1.6059044e-10f));
  Cp = vfma(up, Cp, vs(2.0876757e-9f)); Cq = vfma(uq, Cq, vs(2.0876757e-9f));
  Sp = vfma(up, Sp, vs(2.5052108e-8f)); Sq = vfma(uq, Sq, vs(2.5052108e-8f));
  Cp = vfma(up, Cp, vs(2.7557319e-7f)); Cq = vfma(uq, Cq, vs(2.7557319e-7f));
  Sp = vfma(up, Sp, vs(2.7557319e-6f)); Sq = vfma(uq, Sq, vs(2.7557319e-6f));
  Cp = vfma(up, Cp, vs(2.4801587e-5f)); Cq = vfma(uq, Cq, vs(2.4801587e-5f));
  Sp = vfma(up, Sp, vs(1.9841270e-4f)); Sq = vfma(uq, Sq, vs(1.9841270e-4f));
  Cp = vfma(up, Cp, vs(1.3888889e-3f)); Cq = vfma(uq, Cq, vs(1.3888889e-3f));
  Sp = vfma(up, Sp, vs(8.3333333e-3f)); Sq = vfma(uq, Sq, vs(8.3333333e-3f));
  Cp = vfma(up, Cp, vs(4.1666667e-2f)); Cq = vfma(uq, Cq, vs(4.1666667e-2f));
  Sp = vfma(up, Sp, vs(1.6666667e-1f)); Sq = vfma(uq, Sq, vs(1.6666667e-1f));
  Cp = vfma(up, Cp, vs(0.5f)); Cq = vfma(uq, Cq, vs(0.5f));
  Sp = vfma(up, Sp, vs(1.f)); Sq = vfma(uq, Sq, vs(1.f));
  Cp_ = vfma(up, Cp, vs(1.f)); Cq_ = vfma(uq, Cq, vs(1.f));
  Sp_ = vmul(kd, Sp); Sq_ = vmul(kd, Sq);
}

#ifndef SD_RAY_UNROLL
#define SD_RAY_UNROLL 1
#endif
constexpr int kRayUnroll = SD_RAY_UNROLL;
// Rayleigh sweep for a pair of trial velocities (same truncation depth mmax, same period).
// The row vector is carried as s = (r1, r2, r3, -r4, r5): with that sign convention every entry of the layer matrix
// enters both of its positions with one sign, and with tp = narg_p sin(x_p)/r_p = r sin x (same for q) no operand
// ever has to be negated inside the loop.
SD_HD V2 rayleigh_adjoint2(V2 c, float T, int mmax, const float4* rec, bool ell_only, V2& e2, V2& e3) {
  const V2 csq = vmul(c, c);
  const V2 icsq = v2(1.0f / vx(csq), 1.0f / vy(csq));
  const V2 nicsq = v2(-vx(icsq), -vy(icsq));
  const V2 wvno = v2(SD_TWOPI / (vx(c) * T), SD_TWOPI / (vy(c) * T));
  const V2 nwvno = v2(-vx(wvno), -vy(wvno));
  const int last = mmax - 1;
  const RecLoader ld(rec);
  V2 r1, r2, r3, s4, r5;
  {
    float x1, x2, x3, x4, x5, y1, y2, y3, y4, y5;
    const float4 Rh = ld(last);
    rayleigh_hs_row(vx(csq), vx(icsq), Rh, x1, x2, x3, x4, x5);
    rayleigh_hs_row(vy(csq), vy(icsq), Rh, y1, y2, y3, y4, y5);
    r1 = v2(x1, y1); r2 = v2(x2, y2); r3 = v2(x3, y3); s4 = v2(-x4, -y4); r5 = v2(x5, y5);
  }
  const V2 one = vs(1.f), two = vs(2.f), mone = vs(-1.f);
  // The loop body is laid out by hand (labels): the common path -- solid layer, thin-tier series -- runs straight
  // through; the thick-tier / MUFU terms and the liquid layer sit behind it.  (Taken branches cost instruction-fetch
  // bubbles: with them on the common path the kernel stalled on "no instruction" as often as on arithmetic.)
  for (int m = last - 1; m >= 0; --m) {
    float4 R;
    float ia2, ib2, b2, irho, um;
    V2 kd, nkd2, nargp, nargq, up, uq, sinpr, cosp, sinqr, cosq, tp, tq;   // tp = r sin x (P), tq = r sin x (SV)
    R = ld(m);
    { const float ia = sd_rcp(R.x); ia2 = ia * ia; }
    kd = vmul(wvno, vs(R.w));
    nkd2 = vmul(kd, vmul(nwvno, vs(R.w)));
    nargp = vfma(csq, vs(ia2), mone);
    if (R.y == 0.f) goto liquid_layer;
    { const float ib = sd_rcp(R.y); ib2 = ib * ib; b2 = 2.0f * R.y * R.y; irho = sd_rcp(R.z); }
    nargq = vfma(csq, vs(ib2), mone);
    up = vmul(nkd2, nargp); uq = vmul(nkd2, nargq);
    um = fmaxf(fmaxf(fabsf(vx(up)), fabsf(vy(up))), fmaxf(fabsf(vx(uq)), fabsf(vy(uq))));
    if (!(um < 0.5f)) goto slow_terms;
    series_thin2x2(up, uq, kd, sinpr, cosp, sinqr, cosq);
    tp = vmul(nargp, sinpr); tq = vmul(nargq, sinqr);
  have_terms:
    {
      const V2 g = vmul(vs(b2), icsq);
      const V2 g1 = vadd(g, mone);
      const V2 rhoc = vmul(vs(R.z), csq);
      const V2 nirhoc = vmul(vs(irho), nicsq);
      const V2 rr = vmul(tp, tq), ss = vmul(sinpr, sinqr), cc = vmul(cosp, cosq);
      const V2 rs1 = vmul(tp, cosq), rs2 = vmul(sinqr, cosp), rs3 = vmul(sinpr, cosq), rs4 = vmul(tq, cosp);
      const V2 nss = vmul(ss, mone);
      const V2 m24 = vmul(nargq, nss), m42 = vmul(nargp, nss);   // -a24 = -sinpr tq,  -a42 = -tp sinqr
      const V2 gs = vmul(g, g), g1s = vmul(g1, g1);
      const V2 ccm = vfma(cc, mone, one);
      // The five entries built from (rr, ss, 1 - cc) are  P_n = g^n rr + g1^n ss + (cross term) (1 - cc), n = 0..4.
      // With A = g rr + g1 ccm, B = g1 ss + g ccm and g - g1 = 1:  P1 = A + B,  P2 = g A + g1 B,
      // P3 = g^2 A + g1^2 B,  P4 = g^3 A + g1^3 B - g g1 ccm.
      const V2 A = vfma(g, rr, vmul(g1, ccm)), B = vfma(g1, ss, vmul(g, ccm));
      const V2 w = vfma(g, A, vmul(g1, B));                                              // P2
      const V2 tA = vmul(gs, A), tB = vmul(g1s, B);
      const V2 a11 = vfma(w, mone, cc);
      const V2 a33 = vfma(two, w, one);
      const V2 a12 = vmul(vadd(rs1, rs2), nirhoc);
      const V2 m14 = vmul(vadd(rs3, rs4), nirhoc);                                       // = -a14
      const V2 a13h = vmul(vadd(A, B), nirhoc);                                          // = 0.5 * a13   (P1)
      const V2 a15 = vmul(vfma(two, ccm, vadd(rr, ss)), vmul(nirhoc, nirhoc));           // P0
      const V2 a21 = vmul(rhoc, vfma(g1s, rs3, vmul(gs, rs4)));
      const V2 m41 = vmul(rhoc, vfma(g1s, rs2, vmul(gs, rs1)));                          // = -a41
      const V2 a23h = vfma(g, rs4, vmul(g1, rs3));                                       // = 0.5 * a23
      const V2 a32 = vfma(g1, rs2, vmul(g, rs1));
      const V2 a31 = vmul(rhoc, vadd(tA, tB));                                           // P3
      const V2 a51 = vmul(vmul(rhoc, rhoc), vfma(vmul(vmul(g, g1), ccm), mone, vfma(g, tA, vmul(g1, tB))));   // P4
      // s <- s A' (rows of A as in surfa.f:326-330, signs for s4 = -r4)
      const V2 n1 = vfma(r1, a11, vfma(r2, a21, vfma(r3, a31, vfma(s4, m41, vmul(r5, a51)))));
      const V2 n2 = vfma(r1, a12, vfma(r2, cc, vfma(r3, a32, vfma(s4, m42, vmul(r5, m41)))));
      const V2 n3 = vfma(two, vfma(r1, a13h, vfma(r2, a23h, vfma(s4, a32, vmul(r5, a31)))), vmul(r3, a33));
      const V2 n4 = vfma(r1, m14, vfma(r2, m24, vfma(r3, a23h, vfma(s4, cc, vmul(r5, a21)))));
      const V2 n5 = vfma(r1, a15, vfma(r2, m14, vfma(r3, a13h, vfma(s4, a12, vmul(r5, a11)))));
      r1 = n1; r2 = n2; r3 = n3; s4 = n4; r5 = n5;
    }
    continue;
  slow_terms:
    if (um < 3.0f) {
      series_thick2x2(up, uq, kd, sinpr, cosp, sinqr, cosq);
      tp = vmul(nargp, sinpr); tq = vmul(nargq, sinqr);
    } else {
      half_terms2<true>(nargp, kd, nkd2, tp, sinpr, cosp);
      half_terms2<true>(nargq, kd, nkd2, tq, sinqr, cosq);
    }
    goto have_terms;
  liquid_layer:
    {
      V2 rsinp;
      half_terms2<true>(nargp, kd, nkd2, rsinp, sinpr, cosp);
      if (ell_only) continue;
      const V2 a21 = vmul(vmul(vs(R.z), csq), sinpr);
      const V2 n1 = vfma(r1, cosp, vmul(r2, a21));
      if (m == 0) {
        e2 = r2; e3 = r3;
        return vneg(n1);
      }
      const V2 n4 = vmul(r5, a21), n5 = vmul(r5, cosp);   // s4 = -r4 = +r5 a21
      r1 = n1; r2 = vs(0.f); r3 = vs(0.f); s4 = n4; r5 = n5;
    }
  }
  e2 = r2; e3 = r3;
  return vneg(r1);
}

// Love secular function for a pair of trial velocities: (displacement, stress) propagated from the half-space
// up (surfa.f:143-182).  With q = -k d rb (surfa.f:156): y = sin(q)/rb = -sinr, z = rb sin(q) = -rsin (both
// branches and the rb -> 0 limit of surfa.f:164-166), cos(q) = cs.
SD_HD V2 love_sweep2(V2 c, float T, int mmax, const float4* rec, V2& n1_out, V2& n2_out) {
  const V2 csq = vmul(c, c);
  const V2 wvno = v2(SD_TWOPI / (vx(c) * T), SD_TWOPI / (vy(c) * T));
  const V2 ncsq = v2(-vx(csq), -vy(csq));
  const int last = mmax - 1;
  const RecLoader ld(rec);
  V2 ut, tt;
  float htop;
  {
    const float4 R = ld(last);
    const float h = R.z * R.y * R.y, ib2 = 1.0f / (R.y * R.y);
    htop = h;
    ut = vs(1.f);
    tt = v2(h * sqrtf(fabsf(vx(csq) * ib2 - 1.0f)), h * sqrtf(fabsf(vy(csq) * ib2 - 1.0f)));
  }
  for (int m = last - 1; m >= 0; --m) {
    const float4 R = ld(m);
    if (R.y == 0.f) continue;  // liquid layer skipped (surfa.f:152)
    const float ib = sd_rcp(R.y), ib2 = ib * ib;
    const V2 kd = vmul(wvno, vs(R.w));
    V2 nrsin, sinr, cs;   // nrsin = arg sin x / r = -r sin x
    half_terms2<false>(vfma(ncsq, vs(ib2), vs(1.f)), kd, vmul(kd, kd), nrsin, sinr, cs);
    const float h = R.z * R.y * R.y;
    const float ih = sd_rcp(R.z) * ib2;
    const V2 eut = vfma(cs, ut, vmul(vmul(sinr, tt), vs(ih)));
    const V2 ett = vfma(cs, tt, vmul(vmul(vs(h), nrsin), ut));
    ut = eut;
    tt = ett;
    htop = h;
  }
  // (displacement x rigidity x wavenumber, stress) at the surface: together they carry the scale of the sweep and
  // never vanish together -- the scan normalises -tt by |n1| + |n2|
  n1_out = vmul(vmul(ut, vs(htop)), wvno);
  n2_out = tt;
  return vneg(tt);
}

// ----------------------------------------------------------------------------------------------
// Root of a sampled function by inverse polynomial interpolation: Neville's tableau for x(y) at y = 0 through
// 6 points (x relative to some nearby origin, y the secular-function samples, ordered by x).  e6 uses all six,
// e4 the four contiguous points starting at i4 (0..2), which is a level-3 entry of the same tableau.
SD_HD void inv_interp6(const float* xin, const float* y, int i4, float& e4, float& e6) {
  float x[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 6; ++i) x[i] = xin[i];
  float l3[3] = {0.f, 0.f, 0.f};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int lev = 1; lev < 6; ++lev) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6 - lev; ++i) x[i] = SD_FDIV(y[i + lev] * x[i] - y[i] * x[i + 1], y[i + lev] - y[i]);
    if (lev == 3) { l3[0] = x[0]; l3[1] = x[1]; l3[2] = x[2]; }
  }
  e6 = x[0];
  e4 = (i4 == 0) ? l3[0] : ((i4 == 1) ? l3[1] : l3[2]);
}

// 4-point Lagrange weights at x (abscissae xs[0..3])
SD_HD void lagrange4(const float* xs, float x, float* w) {
  const float d0 = x - xs[0], d1 = x - xs[1], d2 = x - xs[2], d3 = x - xs[3];
  w[0] = SD_FDIV(d1 * d2 * d3, (xs[0] - xs[1]) * (xs[0] - xs[2]) * (xs[0] - xs[3]));
  w[1] = SD_FDIV(d0 * d2 * d3, (xs[1] - xs[0]) * (xs[1] - xs[2]) * (xs[1] - xs[3]));
  w[2] = SD_FDIV(d0 * d1 * d3, (xs[2] - xs[0]) * (xs[2] - xs[1]) * (xs[2] - xs[3]));
  w[3] = SD_FDIV(d0 * d1 * d2, (xs[3] - xs[0]) * (xs[3] - xs[1]) * (xs[3] - xs[2]));
}

// Offsets of the points of a clustered round around a root estimate, in units of the innermost half spacing:
// +-(0.5, 1.5, 3.5, 7.5, 15.5, ...) for i = 0 .. n-1 in ascending order (n even): every interval is twice as
// wide as its inner neighbour, so a bracket found at distance e from the estimate is about e/2 wide.
SD_HD float geometric_offset(int i, int n) {
  const int half = n / 2;
  const int h = (i >= half) ? i - half : half - 1 - i;
  const float m = (float)(1 << h) - 0.5f;
  return (i >= half) ? m : -m;
}

// ----------------------------------------------------------------------------------------------
// Sequential root polish exactly as the reference does it (NEVILL, surfa.f:2-83): interval halving that
// switches to Neville inverse interpolation when the bracket values are monotone and within 10x.
// Only used when the G-section samples show more than one sign change inside the scan bracket (a kink at
// the half-space velocity can put three roots in one 0.01 km/s bracket): which of them the reference
// lands on -- and therefore whether it accepts the period -- depends on this exact evaluation order.
// F: float(float c) evaluates the secular function with mmax pinned.  Returns false on "too many cycles".
template <typename F>
SD_HD bool nevill_seq(F f, float c1, float c2, float del1, float del2, float& cc, int& evals) {
  const float accur1 = 0.1e-5f, accur2 = 0.1e-7f;
  float x[12], y[12];
  int ic = 0, nev = 1, m = 1;
  float c3 = (c1 + c2) / 2.f;
  float del3 = f(c3); evals++;
  for (;;) {
    ic++;
    if (!(ic < 50)) return false;
    const bool inside = (c1 <= c3) ? (c2 > c3) : (c2 < c3);
    bool bisect = !inside;
    if (inside) {
      const float s13 = del1 - del3, s32 = del3 - del2;
      if (signbit(del3) != signbit(del1)) { c2 = c3; del2 = del3; } else { c1 = c3; del1 = del3; }
      if (fabsf(c1 - c2) <= accur1) { cc = c3; return true; }
      if (signbit(s13) != signbit(s32)) nev = 0;
      const float ss1 = fabsf(del1), ss2 = fabsf(del2);
      if (0.1f * ss1 > ss2 || 0.1f * ss2 > ss1 || nev == 0) bisect = true;
      else {
        if (nev == 2) { x[m + 1] = c3; y[m + 1] = del3; }
        else { x[1] = c1; y[1] = del1; x[2] = c2; y[2] = del2; m = 1; }
        bool fail = false;
        for (int kk = 1; kk <= m; ++kk) {
          const int j = m - kk + 1;
          const float den = SD_SUB(y[m + 1], y[j]);
          if (fabsf(den) <= accur2) { fail = true; break; }
          x[j] = SD_DIV(SD_ADD(SD_MUL(-y[j], x[j + 1]), SD_MUL(y[m + 1], x[j])), den);
        }
        if (fail) bisect = true;
        else {
          c3 = x[1];
          del3 = f(c3); evals++;
          nev = 2;
          m = m + 1;
          if (m > 10) m = 10;
          continue;
        }
      }
    }
    if (bisect) {
      c3 = (c1 + c2) / 2.f;
      del3 = f(c3); evals++;
      nev = 1;
      m = 1;
    }
  }
}

// ==============================================================================================
// Group velocity (phase 2 of calcul.f:224-404)
// On-the-fly view of the period-T model of calcul.f:325-337 (all n layers refreshed, flat1 with n).
struct ModelView {
  const float* cst;
  int sc, sl;            // constant (component, layer) sits at cst[component * sc + layer * sl]: rows [8][ld] (sc = ld, sl = 1)
                         // as the prep kernel writes them, or records [layer][8] (sc = 1, sl = 8) as phase 2 stages them
  int n, atten, ndiv, jj0;
  float lt;
  float dtot = -1.f;     // upper bound of the thickness sums of eigen_drop() (prep_kernel), -1 = unknown
  SD_HD float at(int comp, int j) const { return cst[comp * sc + j * sl]; }
  SD_HD void ab(int j, float& a, float& b) const {
    layer_ab_vals(at(C_AREF, j), at(C_BREF, j), at(C_QS, j), (j == n - 1) ? at(C_HSF, j) : at(C_DIF, j), lt, atten, a, b);
  }
  // b only (the layer-dropping walk needs nothing else): same roundings as ab()
  SD_HD float b_only(int j) const {
    const float br = at(C_BREF, j);
    float b = br;
    if (atten) b = SD_MUL(br, SD_ADD(1.0f, SD_FDIV(SD_MUL(at(C_QS, j), lt), SD_PI_ATT)));
    return SD_MUL(b, (j == n - 1) ? at(C_HSF, j) : at(C_DIF, j));
  }
  SD_HD float rho(int j) const { return (j == n - 1) ? at(C_RHOHS, j) : at(C_RHOFL, j); }
  SD_HD int nsub(int j) const { return (ndiv > 1 && j >= jj0 && j < n - 1) ? ndiv : 1; }
  SD_HD float dsub(int j) const {
    if (j == n - 1) return 0.f;
    const float d = at(C_DFL, j);
    return (ndiv > 1 && j >= jj0) ? SD_DIV(d, (float)ndiv) : d;
  }
};

// REIGEN / LEIGEN layer dropping on the subdivided stack (surfa.f:854-866, 475-487).  Returns the layer
// whose properties serve as half-space (jh) and how many sub-layers of layer jl = jh or jh-1 are
// integrated (the walk can end on the last sub-layer of a layer).
struct DropResult { int jh; int jlast; int nlast; };

SD_HD DropResult eigen_drop(const ModelView& mv, float c, float T, float fact, bool use_a) {
  const float dmax = SD_MUL(SD_MUL(fact, T), c);
  float sum = 0.f;
  const int n = mv.n;
  DropResult r; r.jh = n - 1; r.jlast = n - 2; r.nlast = mv.nsub(n - 2 < 0 ? 0 : n - 2);
  // the whole stack is thinner than the limit: the walk below would run to the end without a decision
  if (mv.dtot >= 0.f && mv.dtot <= dmax) return r;
  for (int j = 0; j < n; ++j) {
    const float b = mv.b_only(j);
    if (!(c - b < 0.f)) continue;
    const int ns = mv.nsub(j);
    const float ds = mv.dsub(j);
    for (int s = 0; s < ns; ++s) sum = SD_ADD(sum, ds);  // interior sub-layers never trigger (equal neighbours)
    if (j == n - 1) break;  // ii == mmax
    if (sum <= dmax) continue;
    // sum may have crossed dmax on an interior sub-layer: the reference then keeps walking because the
    // next sub-layer has identical a,b (surfa.f:862-863 fall to 900), so the decision is taken here
    float a2, b2;
    mv.ab(j + 1, a2, b2);
    float a = 0.f, b_chk;
    if (use_a) mv.ab(j, a, b_chk);
    int dec;  // -1: half-space = this sub-layer, +1: next one, 0: keep going
    if (use_a) {
      const float da = a2 - a;
      if (da < 0.f) dec = -1; else if (da > 0.f) dec = 1;
      else { const float db = b2 - b; dec = (db < 0.f) ? -1 : ((db > 0.f) ? 1 : 0); }
    } else {
      const float db = b2 - b; dec = (db < 0.f) ? -1 : ((db > 0.f) ? 1 : 0);
    }
    if (dec == 0) continue;
    if (dec < 0) {
      r.jh = j; r.jlast = j; r.nlast = ns - 1;
      if (r.nlast == 0) { r.jlast = j - 1; r.nlast = (j > 0) ? mv.nsub(j - 1) : 0; }
    } else {
      r.jh = j + 1; r.jlast = j; r.nlast = ns;
    }
    return r;
  }
  return r;
}

// (test instrumentation of the host mirror: how often the reference's second REIGEN iteration is taken)
#if !defined(__CUDACC__) && defined(SD_HOST_COUNTERS)
static long sd_second_pass_count = 0;
#define SD_COUNT_SECOND_PASS() (++sd_second_pass_count)
#else
#define SD_COUNT_SECOND_PASS() ((void)0)
#endif

// The reference integrates the 4x4 stress-displacement system  v' = A v,  v = (ur, uz, tz, tr), with
// classical RK4, 4 sub-steps per sub-layer (surfa.f:945-978).  A is constant inside a layer, so one RK4
// step is exactly the matrix polynomial  P = I + hA + (hA)^2/2 + (hA)^3/6 + (hA)^4/24  applied to v.
// A couples p = (ur, tz) with q = (uz, tr) only (p' = B q, q' = C p), which makes P cheap to build from
// 2x2 blocks once per layer; each step is then 16 FMAs per solution instead of 64.
struct RkCoef { double a12, a13, a21, a24, a31, a34, a42, a43; double h; };
// B and C are trace-free (a24 = -a31, a42 = -a13), so adj(B) = -B, adj(C) = -C and CB = adj(BC): with X = BC,
// t = tr X, d = det X (X^2 = t X - d I) the blocks are
//   pp = (1 - h^4/24 d) I + (h^2/2 + h^4/24 t) X,   qq = adj(pp),
//   pq = h F B,  qp = h adj(F) C   with  F = I + h^2/6 X
// -- 12 stored entries and about half the arithmetic of forming the four blocks separately.
struct StepMat { double pp00, pp01, pp10, pp11, pq00, pq01, pq10, pq11, qp00, qp01, qp10, qp11; };

SD_HD StepMat make_stepmat(const RkCoef& k) {
  // B = [[a31, a34], [a21, a24]] : (uz, tr) -> (ur', tz');  C = [[a13, a12], [a43, a42]] : (ur, tz) -> (uz', tr')
  const double b00 = k.a31, b01 = k.a34, b10 = k.a21;     // b11 = -b00
  const double c00 = k.a13, c01 = k.a12, c10 = k.a43;     // c11 = -c00
  const double h = k.h, h2 = h * h;
  const double x00 = b00 * c00 + b01 * c10, x01 = b00 * c01 - b01 * c00, x10 = b10 * c00 - b00 * c10, x11 = b10 * c01 + b00 * c00;
  const double e2 = h2 * 0.5, e4 = h2 * h2 * (1.0 / 24.0), e3 = h2 * (1.0 / 6.0);
  const double t = x00 + x11, d = x00 * x11 - x01 * x10;
  const double al = 1.0 - e4 * d, be = e2 + e4 * t;
  StepMat m;
  m.pp00 = al + be * x00; m.pp01 = be * x01; m.pp10 = be * x10; m.pp11 = al + be * x11;
  const double f00 = 1.0 + e3 * x00, f01 = e3 * x01, f10 = e3 * x10, f11 = 1.0 + e3 * x11;
  const double hb00 = h * b00, hb01 = h * b01, hb10 = h * b10;    // h B, (1,1) entry = -hb00
  m.pq00 = f00 * hb00 + f01 * hb10; m.pq01 = f00 * hb01 - f01 * hb00;
  m.pq10 = f10 * hb00 + f11 * hb10; m.pq11 = f10 * hb01 - f11 * hb00;
  const double hc00 = h * c00, hc01 = h * c01, hc10 = h * c10;    // h C, (1,1) entry = -hc00
  // adj(F) = [[f11, -f01], [-f10, f00]]
  m.qp00 = f11 * hc00 - f01 * hc10; m.qp01 = f11 * hc01 + f01 * hc00;
  m.qp10 = f00 * hc10 - f10 * hc00; m.qp11 = -(f10 * hc01 + f00 * hc00);
  return m;
}

SD_HD void rk4_step(const StepMat& m, double& ur, double& uz, double& tz, double& tr) {
  const double nur = m.pp00 * ur + m.pp01 * tz + m.pq00 * uz + m.pq01 * tr;
  const double ntz = m.pp10 * ur + m.pp11 * tz + m.pq10 * uz + m.pq11 * tr;
  const double nuz = m.qp00 * ur + m.qp01 * tz + m.pp11 * uz - m.pp01 * tr;   // qq = adj(pp)
  const double ntr = m.qp10 * ur + m.qp11 * tz - m.pp10 * uz + m.pp00 * tr;
  ur = nur; uz = nuz; tz = ntz; tr = ntr;
}

struct Quad9 { double i0yy, i0yz, i0zz, i1yy, i1yz, i1zz, i2yy, i2yz, i2zz; };

// Rayleigh group velocity by energy integrals (REIGEN, surfa.f:714-1190), streaming form: both
// half-space solutions are integrated together and the Boole-rule integrals are accumulated as
// quadratic forms in (y, z), combined with xnorm at the end -- no per-knot storage.
SD_HD float reigen_thread(const ModelView& mv, float T, float c, float ratio, float fact,
                               unsigned long long& nsubsteps) {
  const int n = mv.n;
  const DropResult dr = eigen_drop(mv, c, T, fact, true);
  const float wvno = SD_DIV(SD_TWOPI, SD_MUL(c, T));
  const float wvnosq = SD_MUL(wvno, wvno);
  const float omega = SD_DIV(SD_TWOPI, T);
  const float omegsq = SD_MUL(omega, omega);
  const bool water = !(mv.at(C_BREF, 0) > 0.f);
  // water layer integrals (surfa.f:879-910)
  float w0 = 0.f, w1 = 0.f, w2 = 0.f;
  if (water) {
    float a1, b1;
    mv.ab(0, a1, b1);
    const float rho1 = mv.rho(0), d1 = mv.dsub(0);
    const float xl1 = rho1 * (a1 * a1 - 2.f * b1 * b1);
    const float ra = c / a1;
    const float x = ra * ra - 1.f;
    const float mag = wvno * sqrtf(fabsf(x));
    if (mag <= 1.0e-35f) { w0 = rho1 * d1; }
    else {
      float sin2ra, cosra, rab1;
      if (x >= 0.f) {
        sin2ra = sinf(2.f * mag * d1) / (4.f * mag); cosra = cosf(mag * d1); rab1 = mag * mag;
      } else {
        const float e2 = expf(2.f * mag * d1);
        sin2ra = (0.5f * (e2 - 1.f / e2)) / (4.f * mag);
        const float e1 = expf(mag * d1);
        cosra = 0.5f * (e1 + 1.f / e1); rab1 = -(mag * mag);
      }
      const float cos2rm = 1.f / (cosra * cosra);
      const float fac1 = (0.5f * d1 + sin2ra) * cos2rm;
      const float fac3 = wvno * (0.5f * d1 - sin2ra) * cos2rm;
      const float fac2 = wvno * fac3 / rab1;
      w0 = rho1 * (fac1 + fac2); w1 = xl1 * fac2; w2 = xl1 * fac3;
    }
  }
  // half-space (surfa.f:913-926), float32 like the reference
  float ah, bh;
  mv.ab(dr.jh, ah, bh);
  const float rhoh = mv.rho(dr.jh);
  const float cova = SD_DIV(c, ah), covb = SD_DIV(c, bh);
  const float gam = SD_DIV(2.f, SD_MUL(covb, covb));
  const float gamm1 = SD_SUB(gam, 1.f);
  const float ra = SD_MUL(wvno, sqrtf(fabsf(SD_SUB(SD_MUL(cova, cova), 1.f))));
  const float rb = SD_MUL(wvno, sqrtf(fabsf(SD_SUB(SD_MUL(covb, covb), 1.f))));
  const float det = SD_SUB(wvnosq, SD_MUL(ra, rb));
  const float h = SD_MUL(rhoh, omegsq);
  const float brkt = SD_ADD(SD_MUL(-gamm1, wvno), SD_DIV(SD_MUL(SD_MUL(gam, ra), rb), wvno));
  const double y0tz = (double)SD_DIV(SD_MUL(-h, brkt), det), y0tr = (double)SD_DIV(SD_MUL(-h, ra), det);
  const double z0tz = (double)SD_DIV(SD_MUL(-h, rb), det), z0tr = y0tz;
  if (rb == 0.f) return bh;  // surfa.f:1165

  double xnorm = 0.0, bb = 1.0;
  double zs_ur = 0.0, zs_uz = 1.0, zs_tz = z0tz, zs_tr = z0tr;  // start vector of solution 2
  Quad9 Q;
  const int jfirst = water ? 1 : 0;
  for (int pass = 0; pass < 2; ++pass) {
    double yur = 1.0, yuz = 0.0, ytz = y0tz, ytr = y0tr;
    double zur = zs_ur, zuz = zs_uz, ztz = zs_tz, ztr = zs_tr;
    Q.i0yy = Q.i0yz = Q.i0zz = Q.i1yy = Q.i1yz = Q.i1zz = Q.i2yy = Q.i2yz = Q.i2zz = 0.0;
    for (int j = dr.jlast; j >= jfirst; --j) {
      float a, b;
      mv.ab(j, a, b);
      if (!(b > 0.f)) continue;
      const float rho = mv.rho(j);
      const float xmu = SD_MUL(SD_MUL(rho, b), b);
      const float xlamb = SD_MUL(rho, SD_SUB(SD_MUL(a, a), SD_MUL(SD_MUL(2.f, b), b)));
      const float ds = mv.dsub(j);
      const float ddz = SD_DIV(-ds, 4.f);
      const float f12 = SD_DIV(1.f, SD_ADD(xlamb, SD_MUL(2.f, xmu)));
      const float f13 = SD_MUL(SD_MUL(wvno, xlamb), f12);
      const float f21 = SD_MUL(-omegsq, rho);
      const float f34 = SD_DIV(1.f, xmu);
      const float f43 = SD_ADD(f21, SD_MUL(SD_MUL(SD_MUL(SD_MUL(4.f, wvnosq), xmu), SD_ADD(xlamb, xmu)), f12));
      RkCoef kc;
      kc.a12 = f12; kc.a13 = f13; kc.a21 = f21; kc.a24 = wvno; kc.a31 = -wvno; kc.a34 = f34; kc.a42 = -f13; kc.a43 = f43;
      kc.h = (double)ddz;
      const StepMat sm = make_stepmat(kc);
      const int ns = (j == dr.jlast) ? dr.nlast : mv.nsub(j);
      const double dk = (double)wvno, dlam = (double)xlamb, dmu = (double)xmu;
      const double dkl = dk * dlam;
      // Boole-weighted raw sums over the knots of this layer's sub-layers (weights 7,32,12,32,7; the
      // common factor dz/22.5 and the material constants are applied once per layer)
      // Raw Boole-weighted sums of products of the state components; the derivative terms of the integrands
      // (ur' = tr/mu - k uz, uz' = (tz + k lambda ur)/(lambda + 2 mu), surfa.f:1100-1110) are linear in the state
      // with layer-constant coefficients, so they are applied once per layer instead of once per knot.
      double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, q6 = 0, q7 = 0, q8 = 0, q9 = 0, q10 = 0, q11 = 0;
      for (int s = 0; s < ns; ++s) {
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) {
          const double w = (kk == 0 || kk == 4) ? 7.0 : ((kk == 2) ? 12.0 : 32.0);
          const double ay = w * yur, by = w * yuz, az = w * zur, bz = w * zuz;
          r0 += ay * yur; r1 += ay * zur; r2 += az * zur;
          r3 += by * yuz; r4 += by * zuz; r5 += bz * zuz;
          q6 += by * ytr; q7 += by * ztr + bz * ytr; q8 += bz * ztr;
          q9 += ay * ytz; q10 += ay * ztz + az * ytz; q11 += az * ztz;
          if (kk < 4) { rk4_step(sm, yur, yuz, ytz, ytr); rk4_step(sm, zur, zuz, ztz, ztr); }
        }
      }
      const double r6 = kc.a34 * q6 - dk * r3, r7 = kc.a34 * q7 - 2.0 * dk * r4, r8 = kc.a34 * q8 - dk * r5;
      const double r9 = kc.a12 * (q9 + dkl * r0), r10 = kc.a12 * (q10 + 2.0 * dkl * r1), r11 = kc.a12 * (q11 + dkl * r2);
      {
        const double qw = (double)SD_DIV(SD_DIV(ds, 4.f), 22.5f);
        const double qr = qw * (double)rho, l2m = (double)SD_ADD(xlamb, SD_MUL(2.f, xmu));
        Q.i0yy += qr * (r0 + r3); Q.i0yz += 2.0 * qr * (r1 + r4); Q.i0zz += qr * (r2 + r5);
        Q.i1yy += qw * (l2m * r0 + dmu * r3); Q.i1yz += 2.0 * qw * (l2m * r1 + dmu * r4); Q.i1zz += qw * (l2m * r2 + dmu * r5);
        Q.i2yy += qw * (dmu * r6 - dlam * r9); Q.i2yz += qw * (dmu * r7 - dlam * r10); Q.i2zz += qw * (dmu * r8 - dlam * r11);
      }
      nsubsteps += (unsigned)ns;
    }
    // combine with the surface ellipticity (surfa.f:1056-1065)
    const double aa = zur - (double)ratio * zuz;
    double b_ = (double)ratio * yuz - yur;
    if (fabs(b_) < 1.e-10) b_ = copysign(1.e-10, b_);
    xnorm = aa / b_;
    bb = xnorm * yuz + zuz;
    if (fabs(bb) < 1.e-10) bb = copysign(1.e-10, bb);
    if (pass == 0) {
      const float ampur = (float)((xnorm * yur + zur) / bb);
      const float xtest = fabsf(ampur / ratio - 1.f);
      // The reference forms the eigenfunction knot by knot (xnorm y + z, surfa.f:1076-1086) and squares it; the
      // quadratic forms used here square the dynamic range of the cancellation instead: where y and z have grown
      // nearly parallel (short periods, c above the crustal velocities: the P part grows by e^18 over the depth
      // range) x^2 Qyy + x Qyz + Qzz loses every digit although the reference's first iteration is still fine.
      // The conditioning of the two positive integrals is checked; below 1e-7 the second iteration is taken too
      // (solution 2 restarted from the eigenfunction itself: no cancellation left in the sums).
      const double a0 = xnorm * xnorm * Q.i0yy, b0 = xnorm * Q.i0yz, a1 = xnorm * xnorm * Q.i1yy, b1 = xnorm * Q.i1yz;
      const bool lost = fabs(a0 + b0 + Q.i0zz) < 1.0e-7 * (fabs(a0) + fabs(b0) + fabs(Q.i0zz)) ||
                        fabs(a1 + b1 + Q.i1zz) < 1.0e-7 * (fabs(a1) + fabs(b1) + fabs(Q.i1zz));
      if (!(xtest >= 0.00001f) && !lost) break;  // the reference keeps the first iteration (surfa.f:1068-1069)
      // second iteration of the reference (surfa.f:986-998): solution 2 restarted from z + xnorm*y
      SD_COUNT_SECOND_PASS();
      zs_ur = 0.0 + xnorm * 1.0; zs_uz = 1.0 + xnorm * 0.0; zs_tz = z0tz + xnorm * y0tz; zs_tr = z0tr + xnorm * y0tr;
    }
  }
  const double ib2 = 1.0 / (bb * bb);
  double s0 = (double)w0 + (xnorm * xnorm * Q.i0yy + xnorm * Q.i0yz + Q.i0zz) * ib2;
  double s1 = (double)w1 + (xnorm * xnorm * Q.i1yy + xnorm * Q.i1yz + Q.i1zz) * ib2;
  double s2 = (double)w2 + (xnorm * xnorm * Q.i2yy + xnorm * Q.i2yz + Q.i2zz) * ib2;
  // half-space tail (surfa.f:1151-1178)
  {
    double aur = (xnorm * 1.0 + zs_ur) / bb, auz = (xnorm * 0.0 + zs_uz) / bb;
    if (water && dr.jh == 1) { aur = ratio; auz = 1.0; }
    const double dra = ra, drb = rb, ddet = det, dk = wvno, drho = rhoh;
    const double xmu = (double)SD_MUL(SD_MUL(rhoh, bh), bh);
    const double xlamb = (double)SD_MUL(rhoh, SD_SUB(SD_MUL(ah, ah), SD_MUL(SD_MUL(2.f, bh), bh)));
    const double ap = -drho * (dk * aur + drb * auz) / ddet;
    const double bp = -drho * (-dra * aur / dk - auz) / ddet;
    const double a1 = -dk * ap / drho, a2 = -dk * drb * bp / drho, a3 = dra * ap / drho, a4 = (double)wvnosq * bp / drho;
    const double dmmr = a1 * a1 / (2. * dra) + 2. * a1 * a2 / (dra + drb) + a2 * a2 / (2. * drb);
    const double dmmz = a3 * a3 / (2. * dra) + 2. * a3 * a4 / (dra + drb) + a4 * a4 / (2. * drb);
    const double drsz = -a1 * a3 / 2. - (a1 * a4 * drb + a2 * a3 * dra) / (dra + drb) - a2 * a4 / 2.;
    const double dzsr = -a1 * a3 / 2. - (a1 * a4 * dra + a2 * a3 * drb) / (dra + drb) - a2 * a4 / 2.;
    s0 += drho * (dmmr + dmmz);
    s1 += (xlamb + 2. * xmu) * dmmr + xmu * dmmz;
    s2 += xmu * dzsr - xlamb * drsz;
  }
  (void)n;
  return (float)(((double)wvno * s1 + s2) / ((double)omega * s0));  // surfa.f:1186
}

// ----------------------------------------------------------------------------------------------
// REIGEN, second formulation (the one the kernel runs; reigen_thread above is kept as the plain statement of the
// same algorithm and is what tests/hostmirror compares it with).  Same integration -- RK4 step matrix per layer,
// Boole sums as quadratic forms in the two half-space solutions -- with the per-layer arithmetic trimmed:
//  * the float32 coefficient set-up uses MUFU reciprocals (1 ulp) instead of IEEE divisions with their slow-path
//    calls: the coefficients are layer constants of an ODE, an ulp of them is 1e-7 of U;
//  * the step matrix is formed from hB and hC (X' = (hB)(hC) = h^2 BC makes the RK4 coefficients constants):
//    37 FP64 operations instead of 48;
//  * Boole weights divided by 32: the two interior knots of weight 32 need no multiply; every sum is a chain of
//    FMAs (the cross sums take two FMAs per knot instead of multiply + FMA + add); the first knot initialises
//    the sums (no zeroing): 210 FP64 operations per sub-layer instead of 228;
//  * closing of a layer: the integrands' layer constants are folded once (mu uz ur' - lambda ur uz' =
//    mu f34 uz tr - mu k uz^2 - lambda f12 ur tz - lambda^2 f12 k ur^2): 35 operations instead of 42;
//  * the float32 set-up of the NEXT layer is issued before the FP64 work of the current one, so that its latency
//    (a dependent chain of ~25 float operations and 6 conversions) hides under the FP64 stream.
// +0 tied to a value of the FP64 stream (see reigen_thread2)
#if defined(__CUDA_ARCH__)
SD_HD int sd_anchor(double x) { return __double2int_rz(x * 0.0); }
SD_HD int sd_anchor(float x) { return __float2int_rz(x * 0.0f); }
#else
SD_HD int sd_anchor(double) { return 0; }
SD_HD int sd_anchor(float) { return 0; }
#endif

struct LayerF {
  float hb00, hb01, hb10, hc00, hc01, hc10;   // h B, h C (float32 products; B = [[-k, f34], [f21, k]], C = [[f13, f12], [f43, -f13]])
  float qw, rho, mu, lam, f12, f34;           // closing constants
  int ns;                                     // sub-layers to integrate (0: liquid layer, skipped)
};

SD_HD LayerF reigen_layer_setup(const ModelView& mv, int j, int ns, float wvno, float wvnosq, float omegsq) {
  LayerF L;
  float a, b;
  mv.ab(j, a, b);
  const float rho = mv.rho(j);
  const float b2 = b * b;
  const float xmu = rho * b2;
  const float xlamb = rho * (a * a - 2.f * b2);
  const float ds = mv.at(C_DFL, j) * ((mv.ndiv > 1 && j >= mv.jj0) ? sd_rcp((float)mv.ndiv) : 1.0f);   // surfa.f:800-815
  const float h = -0.25f * ds;
  const float l2m = xlamb + 2.f * xmu;
  const float f12 = sd_rcp(l2m);
  const float f13 = wvno * xlamb * f12;
  const float f21 = -omegsq * rho;
  const float f34 = sd_rcp(xmu);
  const float f43 = f21 + 4.f * wvnosq * xmu * (xlamb + xmu) * f12;
  L.hb00 = -h * wvno; L.hb01 = h * f34; L.hb10 = h * f21;
  L.hc00 = h * f13; L.hc01 = h * f12; L.hc10 = h * f43;
  L.qw = ds * (0.25f * 32.f / 22.5f);     // Boole: dz / 22.5 with the weights carried as w / 32
  L.rho = rho; L.mu = xmu; L.lam = xlamb; L.f12 = f12; L.f34 = f34;
  L.ns = (b > 0.f) ? ns : 0;
  return L;
}

template <typename F> struct StepMat2T { F pp00, pp01, pp10, pp11, pq00, pq01, pq10, pq11, qp00, qp01, qp10, qp11; };
using StepMat2 = StepMat2T<double>;

template <typename F> SD_HD StepMat2T<F> make_stepmat2_t(const LayerF& L) {
  const F hb00 = L.hb00, hb01 = L.hb01, hb10 = L.hb10, hc00 = L.hc00, hc01 = L.hc01, hc10 = L.hc10;
  // X' = (hB)(hC), B = [[b00, b01], [b10, -b00]], C = [[c00, c01], [c10, -c00]]
  const F x00 = hb00 * hc00 + hb01 * hc10, x01 = hb00 * hc01 - hb01 * hc00;
  const F x10 = hb10 * hc00 - hb00 * hc10, x11 = hb10 * hc01 + hb00 * hc00;
  const F t = x00 + x11, d = x00 * x11 - x01 * x10;
  const F al = F(1.0) - d * (F(1.0) / F(24.0)), be = F(0.5) + t * (F(1.0) / F(24.0));
  StepMat2T<F> m;
  m.pp00 = al + be * x00; m.pp01 = be * x01; m.pp10 = be * x10; m.pp11 = al + be * x11;
  const F f00 = F(1.0) + x00 * (F(1.0) / F(6.0)), f01 = x01 * (F(1.0) / F(6.0)), f10 = x10 * (F(1.0) / F(6.0)), f11 = F(1.0) + x11 * (F(1.0) / F(6.0));
  m.pq00 = f00 * hb00 + f01 * hb10; m.pq01 = f00 * hb01 - f01 * hb00;
  m.pq10 = f10 * hb00 + f11 * hb10; m.pq11 = f10 * hb01 - f11 * hb00;
  m.qp00 = f11 * hc00 - f01 * hc10; m.qp01 = f11 * hc01 + f01 * hc00;
  m.qp10 = f00 * hc10 - f10 * hc00; m.qp11 = -f10 * hc01 - f00 * hc00;
  return m;
}

SD_HD StepMat2 make_stepmat2(const LayerF& L) { return make_stepmat2_t<double>(L); }

template <typename F> SD_HD void rk4_step2(const StepMat2T<F>& m, F& ur, F& uz, F& tz, F& tr) {
  const F nur = m.pp00 * ur + m.pp01 * tz + m.pq00 * uz + m.pq01 * tr;
  const F ntz = m.pp10 * ur + m.pp11 * tz + m.pq10 * uz + m.pq11 * tr;
  const F nuz = m.qp00 * ur + m.qp01 * tz + m.pp11 * uz - m.pp01 * tr;   // qq = adj(pp)
  const F ntr = m.qp10 * ur + m.qp11 * tz - m.pp10 * uz + m.pp00 * tr;
  ur = nur; uz = nuz; tz = ntz; tr = ntr;
}

// the twelve Boole sums of a layer: (ur^2, uz^2, uz tr, ur tz) x (yy, yz, zz)
template <typename F> struct Raw12T { F r0, r1, r2, r3, r4, r5, q6, q7, q8, q9, q10, q11; };
template <typename F> struct Quad9T { F i0yy, i0yz, i0zz, i1yy, i1yz, i1zz, i2yy, i2yz, i2zz; };

template <bool FIRST, typename F>
SD_HD void boole_knot(Raw12T<F>& R, F w, bool unit, F yur, F yuz, F ytz, F ytr, F zur, F zuz, F ztz, F ztr) {
  const F ay = unit ? yur : w * yur, by = unit ? yuz : w * yuz, az = unit ? zur : w * zur, bz = unit ? zuz : w * zuz;
  if (FIRST) {
    R.r0 = ay * yur; R.r1 = ay * zur; R.r2 = az * zur;
    R.r3 = by * yuz; R.r4 = by * zuz; R.r5 = bz * zuz;
    R.q6 = by * ytr; R.q7 = by * ztr; R.q8 = bz * ztr;
    R.q9 = ay * ytz; R.q10 = ay * ztz; R.q11 = az * ztz;
  } else {
    R.r0 = fma(ay, yur, R.r0); R.r1 = fma(ay, zur, R.r1); R.r2 = fma(az, zur, R.r2);
    R.r3 = fma(by, yuz, R.r3); R.r4 = fma(by, zuz, R.r4); R.r5 = fma(bz, zuz, R.r5);
    R.q6 = fma(by, ytr, R.q6); R.q7 = fma(by, ztr, R.q7); R.q8 = fma(bz, ztr, R.q8);
    R.q9 = fma(ay, ytz, R.q9); R.q10 = fma(ay, ztz, R.q10); R.q11 = fma(az, ztz, R.q11);
  }
  R.q7 = fma(bz, ytr, R.q7); R.q10 = fma(az, ytz, R.q10);
}

// F = double: the reference's precision of the ODE state (REIGEN is implicit double precision), re-orthogonalised every
// `orth_every` layers.  F = float (the product's default, orth_every = 1): every sub-layer is closed into the sums and
// re-orthogonalised; measured against the double state on eight model families (4 .. 497 layers, water layers,
// velocity inversions, thermal ocean models, periods 5 .. 200 s): |dU| <= 6e-6 km/s (2e-5 at 497 layers), and the same
// error statistics against the oracle -- the noise of U comes from the float32 root c, not from the integration.
// PK (float state only): the two solutions travel as packed pairs (y, z) per component -- the RK4 step applies ONE
// matrix to both, and of the twelve Boole sums the (yy, zz) halves of every triple are a packed operation: 16 instead of
// 32 instructions per step, 12 instead of 18 per knot.
template <typename F, bool PK = false>
SD_HD float reigen_thread2_t(const ModelView& mv, float T, float c, float ratio, float fact,
                             unsigned long long& nsubsteps, int orth_every) {
  static_assert(!PK || sizeof(F) == 4, "packed pairs are float32");
  const DropResult dr = eigen_drop(mv, c, T, fact, true);
  const float wvno = SD_DIV(SD_TWOPI, SD_MUL(c, T));
  const float wvnosq = SD_MUL(wvno, wvno);
  const float omega = SD_DIV(SD_TWOPI, T);
  const float omegsq = SD_MUL(omega, omega);
  const bool water = !(mv.at(C_BREF, 0) > 0.f);
  // water layer integrals (surfa.f:879-910)
  float w0 = 0.f, w1 = 0.f, w2 = 0.f;
  if (water) {
    float a1, b1;
    mv.ab(0, a1, b1);
    const float rho1 = mv.rho(0), d1 = mv.dsub(0);
    const float xl1 = rho1 * (a1 * a1 - 2.f * b1 * b1);
    const float ra = c / a1;
    const float x = ra * ra - 1.f;
    const float mag = wvno * sqrtf(fabsf(x));
    if (mag <= 1.0e-35f) { w0 = rho1 * d1; }
    else {
      float sin2ra, cosra, rab1;
      if (x >= 0.f) {
        sin2ra = sinf(2.f * mag * d1) / (4.f * mag); cosra = cosf(mag * d1); rab1 = mag * mag;
      } else {
        const float e2 = expf(2.f * mag * d1);
        sin2ra = (0.5f * (e2 - 1.f / e2)) / (4.f * mag);
        const float e1 = expf(mag * d1);
        cosra = 0.5f * (e1 + 1.f / e1); rab1 = -(mag * mag);
      }
      const float cos2rm = 1.f / (cosra * cosra);
      const float fac1 = (0.5f * d1 + sin2ra) * cos2rm;
      const float fac3 = wvno * (0.5f * d1 - sin2ra) * cos2rm;
      const float fac2 = wvno * fac3 / rab1;
      w0 = rho1 * (fac1 + fac2); w1 = xl1 * fac2; w2 = xl1 * fac3;
    }
  }
  // half-space (surfa.f:913-926), float32 like the reference
  float ah, bh;
  mv.ab(dr.jh, ah, bh);
  const float rhoh = mv.rho(dr.jh);
  const float cova = SD_DIV(c, ah), covb = SD_DIV(c, bh);
  const float gam = SD_DIV(2.f, SD_MUL(covb, covb));
  const float gamm1 = SD_SUB(gam, 1.f);
  const float ra = SD_MUL(wvno, sqrtf(fabsf(SD_SUB(SD_MUL(cova, cova), 1.f))));
  const float rb = SD_MUL(wvno, sqrtf(fabsf(SD_SUB(SD_MUL(covb, covb), 1.f))));
  const float det = SD_SUB(wvnosq, SD_MUL(ra, rb));
  const float h = SD_MUL(rhoh, omegsq);
  const float brkt = SD_ADD(SD_MUL(-gamm1, wvno), SD_DIV(SD_MUL(SD_MUL(gam, ra), rb), wvno));
  const F y0tz = (F)SD_DIV(SD_MUL(-h, brkt), det), y0tr = (F)SD_DIV(SD_MUL(-h, ra), det);
  const F z0tz = (F)SD_DIV(SD_MUL(-h, rb), det), z0tr = y0tz;
  if (rb == 0.f) return bh;  // surfa.f:1165

  F xnorm = F(0.0), bb = F(1.0), alpha_sum = F(0.0);
  F zs_ur = F(0.0), zs_uz = F(1.0), zs_tz = z0tz, zs_tr = z0tr;  // start vector of solution 2
  Quad9T<F> Q;
  const int jfirst = water ? 1 : 0;
  const F dk = (F)wvno;
  constexpr F W7 = F(7.0) / F(32.0), W12 = F(12.0) / F(32.0);
  for (int pass = 0; pass < 2; ++pass) {
    F yur = F(1.0), yuz = F(0.0), ytz = y0tz, ytr = y0tr;
    F zur = zs_ur, zuz = zs_uz, ztz = zs_tz, ztr = zs_tr;
    Q.i0yy = Q.i0yz = Q.i0zz = Q.i1yy = Q.i1yz = Q.i1zz = Q.i2yy = Q.i2yz = Q.i2zz = F(0.0);
    int since = 0;
    alpha_sum = F(0.0);
    LayerF cur;
    cur.ns = 0;
    if (dr.jlast >= jfirst) cur = reigen_layer_setup(mv, dr.jlast, dr.nlast, wvno, wvnosq, omegsq);
    for (int j = dr.jlast; j >= jfirst; --j) {
      // float32 set-up of the next layer: it does not depend on the FP64 work of this one, and sits in the same
      // basic block (computed unconditionally, for a clamped index), so that the compiler interleaves the two
      const int jn = (j > jfirst) ? j - 1 : jfirst;
      LayerF nxt;
      if (sizeof(F) == 4 && cur.ns > 0) {
        // float32 state: every SUB-layer is closed into the sums and followed by a re-orthogonalisation.  The pair is
        // then never more parallel than one sub-layer makes it (growth ratio of the P and S parts over <= a few km), so
        // the quadratic forms cancel a digit or two, not eight: on every model family of the hunt the group velocities
        // agree with the float64 state's to ~3e-7 km/s and have the same error statistics against the oracle.
        const StepMat2T<F> sm = make_stepmat2_t<F>(cur);
        nxt = reigen_layer_setup(mv, jn, mv.nsub(jn), wvno, wvnosq, omegsq);
        const F qw = (F)cur.qw, mu = (F)cur.mu, lam = (F)cur.lam, f12 = (F)cur.f12;
        const F c0 = qw * (F)cur.rho;
        const F c1b = qw * mu, c1a = fma(F(2.0), c1b, qw * lam);
        const F cA = c1b * (F)cur.f34, cB = -c1b * dk;
        const F cC = -(qw * lam) * f12, cD = cC * (dk * lam);
        // (PK: the step matrix broadcast into pairs once per layer; qq = adj(pp) needs two entries negated)
        V2 Ppp00, Ppp01, Ppp10, Ppp11, Ppq00, Ppq01, Ppq10, Ppq11, Pqp00, Pqp01, Pqp10, Pqp11, Pnpp01, Pnpp10;
        if (PK) {
          Ppp00 = vs((float)sm.pp00); Ppp01 = vs((float)sm.pp01); Ppp10 = vs((float)sm.pp10); Ppp11 = vs((float)sm.pp11);
          Ppq00 = vs((float)sm.pq00); Ppq01 = vs((float)sm.pq01); Ppq10 = vs((float)sm.pq10); Ppq11 = vs((float)sm.pq11);
          Pqp00 = vs((float)sm.qp00); Pqp01 = vs((float)sm.qp01); Pqp10 = vs((float)sm.qp10); Pqp11 = vs((float)sm.qp11);
          Pnpp01 = vs(-(float)sm.pp01); Pnpp10 = vs(-(float)sm.pp10);
        }
        for (int s = 0; s < cur.ns; ++s) {
          Raw12T<F> R;
          if (PK) {
            V2 pur = v2((float)yur, (float)zur), puz = v2((float)yuz, (float)zuz), ptz = v2((float)ytz, (float)ztz), ptr = v2((float)ytr, (float)ztr);
            V2 R02, R35, Q68, Q911;
            float r1, r4, q7, q10;
            const V2 w7 = vs(7.0f / 32.0f), w12 = vs(12.0f / 32.0f);
            // knot 0 (weight 7/32) initialises the sums
            {
              const V2 A = vmul(w7, pur), B = vmul(w7, puz);
              R02 = vmul(A, pur); r1 = vx(A) * vy(pur); R35 = vmul(B, puz); r4 = vx(B) * vy(puz);
              Q68 = vmul(B, ptr); q7 = vx(B) * vy(ptr); Q911 = vmul(A, ptz); q10 = vx(A) * vy(ptz);
              q7 = fmaf(vy(B), vx(ptr), q7); q10 = fmaf(vy(A), vx(ptz), q10);
            }
#pragma unroll
            for (int kn = 1; kn <= 4; ++kn) {
              {
                const V2 nur = vfma(Ppp00, pur, vfma(Ppp01, ptz, vfma(Ppq00, puz, vmul(Ppq01, ptr))));
                const V2 ntz = vfma(Ppp10, pur, vfma(Ppp11, ptz, vfma(Ppq10, puz, vmul(Ppq11, ptr))));
                const V2 nuz = vfma(Pqp00, pur, vfma(Pqp01, ptz, vfma(Ppp11, puz, vmul(Pnpp01, ptr))));
                const V2 ntr = vfma(Pqp10, pur, vfma(Pqp11, ptz, vfma(Pnpp10, puz, vmul(Ppp00, ptr))));
                pur = nur; puz = nuz; ptz = ntz; ptr = ntr;
              }
              // weights 32/32 (knots 1, 3), 12/32 (knot 2), 7/32 (knot 4)
              const V2 A = (kn == 2) ? vmul(w12, pur) : ((kn == 4) ? vmul(w7, pur) : pur);
              const V2 B = (kn == 2) ? vmul(w12, puz) : ((kn == 4) ? vmul(w7, puz) : puz);
              R02 = vfma(A, pur, R02); r1 = fmaf(vx(A), vy(pur), r1); R35 = vfma(B, puz, R35); r4 = fmaf(vx(B), vy(puz), r4);
              Q68 = vfma(B, ptr, Q68); q7 = fmaf(vx(B), vy(ptr), q7); Q911 = vfma(A, ptz, Q911); q10 = fmaf(vx(A), vy(ptz), q10);
              q7 = fmaf(vy(B), vx(ptr), q7); q10 = fmaf(vy(A), vx(ptz), q10);
            }
            yur = vx(pur); zur = vy(pur); yuz = vx(puz); zuz = vy(puz); ytz = vx(ptz); ztz = vy(ptz); ytr = vx(ptr); ztr = vy(ptr);
            R.r0 = vx(R02); R.r2 = vy(R02); R.r1 = r1; R.r3 = vx(R35); R.r5 = vy(R35); R.r4 = r4;
            R.q6 = vx(Q68); R.q8 = vy(Q68); R.q7 = q7; R.q9 = vx(Q911); R.q11 = vy(Q911); R.q10 = q10;
          } else {
          boole_knot<true>(R, W7, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, F(1.0), true, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, W12, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, F(1.0), true, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, W7, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          }
          const F R1 = R.r1 + R.r1, R4 = R.r4 + R.r4;
          Q.i0yy = fma(c0, R.r3, fma(c0, R.r0, Q.i0yy)); Q.i0yz = fma(c0, R4, fma(c0, R1, Q.i0yz)); Q.i0zz = fma(c0, R.r5, fma(c0, R.r2, Q.i0zz));
          Q.i1yy = fma(c1b, R.r3, fma(c1a, R.r0, Q.i1yy)); Q.i1yz = fma(c1b, R4, fma(c1a, R1, Q.i1yz)); Q.i1zz = fma(c1b, R.r5, fma(c1a, R.r2, Q.i1zz));
          Q.i2yy = fma(cD, R.r0, fma(cC, R.q9, fma(cB, R.r3, fma(cA, R.q6, Q.i2yy))));
          Q.i2yz = fma(cD, R1, fma(cC, R.q10, fma(cB, R4, fma(cA, R.q7, Q.i2yz))));
          Q.i2zz = fma(cD, R.r2, fma(cC, R.q11, fma(cB, R.r5, fma(cA, R.q8, Q.i2zz))));
          if ((j > jfirst || s + 1 < cur.ns) && ++since >= orth_every) {
            since = 0;
            const F syy = yur * yur + yuz * yuz + ytz * ytz + ytr * ytr;
            const F syz = yur * zur + yuz * zuz + ytz * ztz + ytr * ztr;
            const F al = (F)SD_FDIV((float)syz, (float)syy);      // (any coefficient is carried through the sums exactly: no IEEE division needed)
            zur = fma(-al, yur, zur); zuz = fma(-al, yuz, zuz); ztz = fma(-al, ytz, ztz); ztr = fma(-al, ytr, ztr);
            alpha_sum += al;
            { const F n = fma(F(-2.0) * al, Q.i0yy, Q.i0yz); Q.i0zz = fma(F(-0.5) * al, Q.i0yz + n, Q.i0zz); Q.i0yz = n; }
            { const F n = fma(F(-2.0) * al, Q.i1yy, Q.i1yz); Q.i1zz = fma(F(-0.5) * al, Q.i1yz + n, Q.i1zz); Q.i1yz = n; }
            { const F n = fma(F(-2.0) * al, Q.i2yy, Q.i2yz); Q.i2zz = fma(F(-0.5) * al, Q.i2yz + n, Q.i2zz); Q.i2yz = n; }
          }
        }
        nsubsteps += (unsigned)cur.ns;
      } else if (cur.ns > 0) {
        const StepMat2T<F> sm = make_stepmat2_t<F>(cur);
        Raw12T<F> R;
        // first sub-layer peeled: its first knot initialises the sums, and the straight-line code lets the compiler
        // interleave the next layer's float32 set-up with this FP64 stream (stacks of >= 21 layers have ns = 1)
        boole_knot<true>(R, W7, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
        rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
        // The compiler's list scheduler issues the (independent) set-up chain of the next layer first and the FP64
        // stream after it; a warp then spends ~200 cycles in dependent float32 instructions per layer while the FP64
        // pipe waits for the other warps.  The layer index is tied to a value of this point of the stream (+0, not
        // foldable for IEEE doubles), which places the chain in the middle of the FP64 work.
        // (measured: one anchor F(74.2) -> F(72.7) ms; the chain cut in three anchored stages F(73.8) ms)
        nxt = reigen_layer_setup(mv, jn + sd_anchor(yur), mv.nsub(jn), wvno, wvnosq, omegsq);
        boole_knot<false>(R, F(1.0), true, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
        rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
        boole_knot<false>(R, W12, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
        rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
        boole_knot<false>(R, F(1.0), true, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
        rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
        boole_knot<false>(R, W7, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
        for (int s = 1; s < cur.ns; ++s) {
          boole_knot<false>(R, W7, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, F(1.0), true, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, W12, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, F(1.0), true, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
          rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr);
          boole_knot<false>(R, W7, false, yur, yuz, ytz, ytr, zur, zuz, ztz, ztr);
        }
        // closing: layer constants applied once (the yz sums of the squares enter doubled)
        {
          const F qw = (F)cur.qw, mu = (F)cur.mu, lam = (F)cur.lam, f12 = (F)cur.f12;
          const F c0 = qw * (F)cur.rho;
          const F c1b = qw * mu, c1a = fma(F(2.0), c1b, qw * lam);      // qw (lambda + 2 mu), qw mu
          const F cA = c1b * (F)cur.f34, cB = -c1b * dk;
          const F cC = -(qw * lam) * f12, cD = cC * (dk * lam);
          const F R1 = R.r1 + R.r1, R4 = R.r4 + R.r4;
          Q.i0yy = fma(c0, R.r3, fma(c0, R.r0, Q.i0yy)); Q.i0yz = fma(c0, R4, fma(c0, R1, Q.i0yz)); Q.i0zz = fma(c0, R.r5, fma(c0, R.r2, Q.i0zz));
          Q.i1yy = fma(c1b, R.r3, fma(c1a, R.r0, Q.i1yy)); Q.i1yz = fma(c1b, R4, fma(c1a, R1, Q.i1yz)); Q.i1zz = fma(c1b, R.r5, fma(c1a, R.r2, Q.i1zz));
          Q.i2yy = fma(cD, R.r0, fma(cC, R.q9, fma(cB, R.r3, fma(cA, R.q6, Q.i2yy))));
          Q.i2yz = fma(cD, R1, fma(cC, R.q10, fma(cB, R4, fma(cA, R.q7, Q.i2yz))));
          Q.i2zz = fma(cD, R.r2, fma(cC, R.q11, fma(cB, R.r5, fma(cA, R.q8, Q.i2zz))));
        }
        nsubsteps += (unsigned)cur.ns;
        // Re-orthogonalisation: both solutions grow upwards and turn parallel (the P part dominates), and the
        // eigenfunction is their small difference.  The reference notices that at the surface (xtest, surfa.f:1066)
        // and integrates a second time from z + xnorm y -- 36 % of the evaluations of the benchmark set.  Here, every
        // kOrthEvery layers the component of z along y is removed (z <- z - al y): the sums are bilinear forms in
        // (y, z), so they are carried into the new basis exactly (Syz' = Syz - al Syy, Sz'z' = Szz - al (Syz + Syz')),
        // the pair never becomes parallel and the first integration is already the accurate one.
        if (++since >= orth_every && j > jfirst) {
          since = 0;
          const F syy = yur * yur + yuz * yuz + ytz * ytz + ytr * ytr;
          const F syz = yur * zur + yuz * zuz + ytz * ztz + ytr * ztr;
          const F al = syz / syy;
          zur = fma(-al, yur, zur); zuz = fma(-al, yuz, zuz); ztz = fma(-al, ytz, ztz); ztr = fma(-al, ytr, ztr);
          alpha_sum += al;
          // (the yz sums are stored doubled: S = 2 Syz)
          { const F n = fma(-F(2.0) * al, Q.i0yy, Q.i0yz); Q.i0zz = fma(-F(0.5) * al, Q.i0yz + n, Q.i0zz); Q.i0yz = n; }
          { const F n = fma(-F(2.0) * al, Q.i1yy, Q.i1yz); Q.i1zz = fma(-F(0.5) * al, Q.i1yz + n, Q.i1zz); Q.i1yz = n; }
          { const F n = fma(-F(2.0) * al, Q.i2yy, Q.i2yz); Q.i2zz = fma(-F(0.5) * al, Q.i2yz + n, Q.i2zz); Q.i2yz = n; }
        }
      } else nxt = reigen_layer_setup(mv, jn, mv.nsub(jn), wvno, wvnosq, omegsq);
      cur = nxt;
    }
    // combine with the surface ellipticity (surfa.f:1056-1065)
    const F aa = zur - (F)ratio * zuz;
    F b_ = (F)ratio * yuz - yur;
    if (fabs(b_) < F(1.e-10)) b_ = copysign(F(1.e-10), b_);
    xnorm = aa / b_;
    bb = xnorm * yuz + zuz;
    if (fabs(bb) < F(1.e-10)) bb = copysign(F(1.e-10), bb);
    if (pass == 0) {
      const float ampur = (float)((xnorm * yur + zur) / bb);
      const float xtest = fabsf(ampur / ratio - 1.f);
      // The reference forms the eigenfunction knot by knot (xnorm y + z, surfa.f:1076-1086) and squares it; the
      // quadratic forms used here square the dynamic range of the cancellation instead: where y and z have grown
      // nearly parallel (short periods, c above the crustal velocities: the P part grows by e^18 over the depth
      // range) x^2 Qyy + x Qyz + Qzz loses every digit although the reference's first iteration is still fine.
      // The conditioning of the two positive integrals is checked; below F(1e-7) the second iteration is taken too
      // (solution 2 restarted from the eigenfunction itself: no cancellation left in the sums).
      const F a0 = xnorm * xnorm * Q.i0yy, b0 = xnorm * Q.i0yz, a1 = xnorm * xnorm * Q.i1yy, b1 = xnorm * Q.i1yz;
      const bool lost = fabs(a0 + b0 + Q.i0zz) < F(1.0e-7) * (fabs(a0) + fabs(b0) + fabs(Q.i0zz)) ||
                        fabs(a1 + b1 + Q.i1zz) < F(1.0e-7) * (fabs(a1) + fabs(b1) + fabs(Q.i1zz));
      if (!(xtest >= 0.00001f) && !lost) break;  // the reference keeps the first iteration (surfa.f:1068-1069)
      // second iteration of the reference (surfa.f:986-998): solution 2 restarted from z + xnorm*y
      SD_COUNT_SECOND_PASS();
      // (in terms of this pass's start vector the eigenfunction is (xnorm - alpha_sum) y + z)
      const F xo = xnorm - alpha_sum;
      zs_ur = zs_ur + xo * F(1.0); zs_uz = zs_uz + xo * F(0.0); zs_tz = zs_tz + xo * y0tz; zs_tr = zs_tr + xo * y0tr;
    }
  }
  const F ib2 = F(1.0) / (bb * bb);
  F s0 = (F)w0 + (xnorm * xnorm * Q.i0yy + xnorm * Q.i0yz + Q.i0zz) * ib2;
  F s1 = (F)w1 + (xnorm * xnorm * Q.i1yy + xnorm * Q.i1yz + Q.i1zz) * ib2;
  F s2 = (F)w2 + (xnorm * xnorm * Q.i2yy + xnorm * Q.i2yz + Q.i2zz) * ib2;
  // half-space tail (surfa.f:1151-1178)
  {
    const F xo = xnorm - alpha_sum;   // coefficient of y relative to the start vectors of the last pass
    F aur = (xo * F(1.0) + zs_ur) / bb, auz = (xo * F(0.0) + zs_uz) / bb;
    if (water && dr.jh == 1) { aur = ratio; auz = F(1.0); }
    const F dra = ra, drb = rb, ddet = det, drho = rhoh;
    const F xmu = (F)SD_MUL(SD_MUL(rhoh, bh), bh);
    const F xlamb = (F)SD_MUL(rhoh, SD_SUB(SD_MUL(ah, ah), SD_MUL(SD_MUL(2.f, bh), bh)));
    const F ap = -drho * (dk * aur + drb * auz) / ddet;
    const F bp = -drho * (-dra * aur / dk - auz) / ddet;
    const F a1 = -dk * ap / drho, a2 = -dk * drb * bp / drho, a3 = dra * ap / drho, a4 = (F)wvnosq * bp / drho;
    const F dmmr = a1 * a1 / (F(2.) * dra) + F(2.) * a1 * a2 / (dra + drb) + a2 * a2 / (F(2.) * drb);
    const F dmmz = a3 * a3 / (F(2.) * dra) + F(2.) * a3 * a4 / (dra + drb) + a4 * a4 / (F(2.) * drb);
    const F drsz = -a1 * a3 / F(2.) - (a1 * a4 * drb + a2 * a3 * dra) / (dra + drb) - a2 * a4 / F(2.);
    const F dzsr = -a1 * a3 / F(2.) - (a1 * a4 * dra + a2 * a3 * drb) / (dra + drb) - a2 * a4 / F(2.);
    s0 += drho * (dmmr + dmmz);
    s1 += (xlamb + F(2.) * xmu) * dmmr + xmu * dmmz;
    s2 += xmu * dzsr - xlamb * drsz;
  }
  return (float)(((F)wvno * s1 + s2) / ((F)omega * s0));  // surfa.f:1186
}

// The reference's precision (surfa.f: implicit double precision in REIGEN)
SD_HD float reigen_thread2(const ModelView& mv, float T, float c, float ratio, float fact, unsigned long long& nsubsteps) {
  return reigen_thread2_t<double>(mv, T, c, ratio, fact, nsubsteps, 8);
}

// ----------------------------------------------------------------------------------------------
// Partial derivatives of the Rayleigh phase velocity with respect to Vp, Vs and density of every layer of the
// period's model (REIGEN's dcda, dcdb, dcdr: surfa.f:1130-1135 per (sub-)layer, 1179-1185 for the half-space,
// 1202-1208 normalisation by dL/dk) -- what the reference computes along with the group velocity and never hands
// out; they replace the 2n + 1 forward solves of the finite-difference SensKernelPert (senskernel.py:130-158).
// Two integrations with the re-orthogonalisation schedule of reigen_thread2: the first one yields the surface
// combination (xnorm, bb) and the total of the re-orthogonalisation coefficients; in the second one the
// eigenfunction is known in every segment, (x_seg y + z) / bb with x_seg = xnorm - (alpha_total - alpha_so_far), so
// the Boole integrals of the layer are formed from the eigenfunction itself.
// out_da / out_db / out_dr: [n] entries with stride `os` (zero where the layer is below the truncation or liquid).
// Returns false when the half-space is degenerate (rb = 0, surfa.f:1165).
SD_HD bool reigen_partials_thread(const ModelView& mv, float T, float c, float ratio, float fact, float* out_da,
                                  float* out_db, float* out_dr, int os) {
  const int n = mv.n;
  for (int j = 0; j < n; ++j) { out_da[j * os] = 0.f; out_db[j * os] = 0.f; out_dr[j * os] = 0.f; }
  const DropResult dr = eigen_drop(mv, c, T, fact, true);
  const float wvno = SD_DIV(SD_TWOPI, SD_MUL(c, T));
  const float wvnosq = SD_MUL(wvno, wvno);
  const float omega = SD_DIV(SD_TWOPI, T);
  const float omegsq = SD_MUL(omega, omega);
  const bool water = !(mv.at(C_BREF, 0) > 0.f);
  double w1 = 0.0, w2 = 0.0;     // water-layer parts of I1, I2 (surfa.f:879-910)
  if (water) {
    float a1, b1;
    mv.ab(0, a1, b1);
    const float rho1 = mv.rho(0), d1 = mv.dsub(0);
    const float xl1 = rho1 * (a1 * a1 - 2.f * b1 * b1);
    const float ra = c / a1, x = ra * ra - 1.f, mag = wvno * sqrtf(fabsf(x));
    if (mag > 1.0e-35f) {
      float sin2ra, cosra, rab1;
      if (x >= 0.f) { sin2ra = sinf(2.f * mag * d1) / (4.f * mag); cosra = cosf(mag * d1); rab1 = mag * mag; }
      else {
        const float e2 = expf(2.f * mag * d1), e1 = expf(mag * d1);
        sin2ra = (0.5f * (e2 - 1.f / e2)) / (4.f * mag); cosra = 0.5f * (e1 + 1.f / e1); rab1 = -(mag * mag);
      }
      const float cos2rm = 1.f / (cosra * cosra);
      const float fac3 = wvno * (0.5f * d1 - sin2ra) * cos2rm, fac2 = wvno * fac3 / rab1;
      w1 = xl1 * fac2; w2 = xl1 * fac3;
    }
  }
  float ah, bh;
  mv.ab(dr.jh, ah, bh);
  const float rhoh = mv.rho(dr.jh);
  const float cova = SD_DIV(c, ah), covb = SD_DIV(c, bh);
  const float gam = SD_DIV(2.f, SD_MUL(covb, covb)), gamm1 = SD_SUB(gam, 1.f);
  const float ra = SD_MUL(wvno, sqrtf(fabsf(SD_SUB(SD_MUL(cova, cova), 1.f))));
  const float rb = SD_MUL(wvno, sqrtf(fabsf(SD_SUB(SD_MUL(covb, covb), 1.f))));
  const float det = SD_SUB(wvnosq, SD_MUL(ra, rb));
  const float hh = SD_MUL(rhoh, omegsq);
  const float brkt = SD_ADD(SD_MUL(-gamm1, wvno), SD_DIV(SD_MUL(SD_MUL(gam, ra), rb), wvno));
  const double y0tz = (double)SD_DIV(SD_MUL(-hh, brkt), det), y0tr = (double)SD_DIV(SD_MUL(-hh, ra), det);
  const double z0tz = (double)SD_DIV(SD_MUL(-hh, rb), det), z0tr = y0tz;
  if (rb == 0.f) return false;
  const int jfirst = water ? 1 : 0;
  const double dk = (double)wvno, dk2 = (double)wvnosq, dom2 = (double)omegsq;
  constexpr int kOrthEvery = 8;
  constexpr double W7 = 7.0 / 32.0, W12 = 12.0 / 32.0;
  double xnorm = 0.0, bb = 1.0, alpha_total = 0.0, sumi1 = w1, sumi2 = w2;
  for (int pass = 0; pass < 2; ++pass) {
    double yur = 1.0, yuz = 0.0, ytz = y0tz, ytr = y0tr;
    double zur = 0.0, zuz = 1.0, ztz = z0tz, ztr = z0tr;
    double alpha_sofar = 0.0;
    int since = 0;
    for (int j = dr.jlast; j >= jfirst; --j) {
      const int ns = (j == dr.jlast) ? dr.nlast : mv.nsub(j);
      const LayerF L = reigen_layer_setup(mv, j, ns, wvno, wvnosq, omegsq);
      if (L.ns <= 0) continue;
      const StepMat2 sm = make_stepmat2(L);
      const double xs = xnorm - (alpha_total - alpha_sofar), ib = 1.0 / bb;
      double Err = 0, Ezz = 0, Eztr = 0, Ertz = 0, Etrtr = 0, Etztz = 0;   // Boole sums (weights / 32) of the eigenfunction products
      for (int sub = 0; sub < L.ns; ++sub) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int kk = 0; kk < 5; ++kk) {
          if (pass == 1) {
            const double w = (kk == 0 || kk == 4) ? W7 : ((kk == 2) ? W12 : 1.0);
            const double eur = (xs * yur + zur) * ib, euz = (xs * yuz + zuz) * ib, etz = (xs * ytz + ztz) * ib, etr = (xs * ytr + ztr) * ib;
            Err += w * eur * eur; Ezz += w * euz * euz; Eztr += w * euz * etr; Ertz += w * eur * etz;
            Etrtr += w * etr * etr; Etztz += w * etz * etz;
          }
          if (kk < 4) { rk4_step2(sm, yur, yuz, ytz, ytr); rk4_step2(sm, zur, zuz, ztz, ztr); }
        }
      }
      if (pass == 1) {
        // integrals of the layer (surfa.f:1100-1129) from the raw sums; ur' = tr / mu - k uz, uz' = (tz + k lambda ur) / (lambda + 2 mu)
        const double qw = (double)L.qw, mu = (double)L.mu, lam = (double)L.lam, f12 = (double)L.f12, f34 = (double)L.f34;
        const double dmmr = qw * Err, dmmz = qw * Ezz;
        const double smmr = qw * (f34 * f34 * Etrtr - 2.0 * f34 * dk * Eztr + dk2 * Ezz);
        const double smmz = qw * f12 * f12 * (Etztz + 2.0 * dk * lam * Ertz + dk2 * lam * lam * Err);
        const double drsz = qw * f12 * (Ertz + dk * lam * Err);
        const double dzsr = qw * (f34 * Eztr - dk * Ezz);
        sumi1 += (lam + 2.0 * mu) * dmmr + mu * dmmz;
        sumi2 += mu * dzsr - lam * drsz;
        const double dldl = -dk2 * dmmr + 2.0 * dk * drsz - smmz;
        const double dldm = -dk2 * (2.0 * dmmr + dmmz) - 2.0 * dk * dzsr - (2.0 * smmz + smmr);
        const double dldr = dom2 * (dmmr + dmmz);
        float a, b;
        mv.ab(j, a, b);
        const double rho = (double)L.rho;
        out_db[j * os] = (float)(2.0 * rho * (double)b * (double)c * (dldm - 2.0 * dldl) / dk);      // surfa.f:1133-1135 (un-normalised)
        out_da[j * os] = (float)(2.0 * rho * (double)a * (double)c * dldl / dk);
        out_dr[j * os] = (float)(((double)c / dk) * (dldr + lam * dldl / rho + mu * dldm / rho));
      }
      if (++since >= kOrthEvery && j > jfirst) {
        since = 0;
        const double syy = yur * yur + yuz * yuz + ytz * ytz + ytr * ytr;
        const double syz = yur * zur + yuz * zuz + ytz * ztz + ytr * ztr;
        const double al = syz / syy;
        zur = fma(-al, yur, zur); zuz = fma(-al, yuz, zuz); ztz = fma(-al, ytz, ztz); ztr = fma(-al, ytr, ztr);
        alpha_sofar += al;
      }
    }
    if (pass == 0) {
      alpha_total = alpha_sofar;
      const double aa = zur - (double)ratio * zuz;
      double b_ = (double)ratio * yuz - yur;
      if (fabs(b_) < 1.e-10) b_ = copysign(1.e-10, b_);
      xnorm = aa / b_;
      bb = xnorm * yuz + zuz;
      if (fabs(bb) < 1.e-10) bb = copysign(1.e-10, bb);
    }
  }
  // half-space (surfa.f:1145-1185): amplitudes at its top from the start vectors
  {
    const double xo = xnorm - alpha_total;
    double aur = (xo * 1.0 + 0.0) / bb, auz = (xo * 0.0 + 1.0) / bb;
    if (water && dr.jh == 1) { aur = ratio; auz = 1.0; }
    const double dra = ra, drb = rb, ddet = det, drho = rhoh;
    const double xmu = (double)SD_MUL(SD_MUL(rhoh, bh), bh);
    const double xlamb = (double)SD_MUL(rhoh, SD_SUB(SD_MUL(ah, ah), SD_MUL(SD_MUL(2.f, bh), bh)));
    const double ap = -drho * (dk * aur + drb * auz) / ddet;
    const double bp = -drho * (-dra * aur / dk - auz) / ddet;
    const double a1 = -dk * ap / drho, a2 = -dk * drb * bp / drho, a3 = dra * ap / drho, a4 = dk2 * bp / drho;
    const double dmmr = a1 * a1 / (2. * dra) + 2. * a1 * a2 / (dra + drb) + a2 * a2 / (2. * drb);
    const double dmmz = a3 * a3 / (2. * dra) + 2. * a3 * a4 / (dra + drb) + a4 * a4 / (2. * drb);
    const double smmz = dra * a3 * a3 / 2. + 2. * dra * drb * a3 * a4 / (dra + drb) + drb * a4 * a4 / 2.;
    const double smmr = dra * a1 * a1 / 2. + 2. * dra * drb * a1 * a2 / (dra + drb) + drb * a2 * a2 / 2.;
    const double drsz = -a1 * a3 / 2. - (a1 * a4 * drb + a2 * a3 * dra) / (dra + drb) - a2 * a4 / 2.;
    const double dzsr = -a1 * a3 / 2. - (a1 * a4 * dra + a2 * a3 * drb) / (dra + drb) - a2 * a4 / 2.;
    sumi1 += (xlamb + 2. * xmu) * dmmr + xmu * dmmz;
    sumi2 += xmu * dzsr - xlamb * drsz;
    const double dldr = dom2 * (dmmr + dmmz);
    const double dldm = -dk2 * (2. * dmmr + dmmz) - 2. * dk * dzsr - (2. * smmz + smmr);
    const double dldl = -dk2 * dmmr + 2. * dk * drsz - smmz;
    out_da[dr.jh * os] = (float)(2. * drho * (double)ah * (double)c * dldl / dk);
    out_db[dr.jh * os] = (float)(2. * drho * (double)bh * (double)c * (dldm - 2. * dldl) / dk);
    out_dr[dr.jh * os] = (float)(((double)c / dk) * (dldr + xlamb * dldl / drho + xmu * dldm / drho));
  }
  const double idldk = 1.0 / (-2.0 * (dk * sumi1 + sumi2));      // surfa.f:1203
  for (int j = jfirst; j <= dr.jh; ++j) {
    out_da[j * os] = (float)((double)out_da[j * os] * idldk);
    out_db[j * os] = (float)((double)out_db[j * os] * idldk);
    out_dr[j * os] = (float)((double)out_dr[j * os] * idldk);
  }
  return true;
}

// Love group velocity (LEIGEN, surfa.f:374-606), float32 like the reference.
SD_HD float leigen_thread(const ModelView& mv, float T, float c, float fact, unsigned long long& nsubsteps) {
  const DropResult dr = eigen_drop(mv, c, T, (fact <= 0.f) ? 7.0f : fact, false);
  const float wvno = SD_TWOPI / (c * T);
  float ah, bh;
  mv.ab(dr.jh, ah, bh);
  const float rhoh = mv.rho(dr.jh);
  float ut0 = 1.f, ut = 1.f, sumi0 = 0.f, sumi1 = 0.f;
  for (int attempt = 0; attempt < 16; ++attempt) {
    ut = ut0;
    const float covb = c / bh;
    const float hh = rhoh * bh * bh;
    const float rbh = wvno * sqrtf(fabsf(covb * covb - 1.f));
    float tq = -hh * rbh * ut0;
    float dm = (rbh == 0.f) ? 1.0e25f : 0.5f / rbh;
    sumi0 = rhoh * dm;
    sumi1 = hh * dm;
    bool restart = false;
    for (int j = dr.jlast; j >= 0 && !restart; --j) {
      float a, b;
      mv.ab(j, a, b);
      if (b == 0.f) continue;
      const float rho = mv.rho(j);
      const float cb = c / b;
      const float rb = wvno * sqrtf(fabsf(cb * cb - 1.f));
      const float h = rho * b * b;
      const float dz = mv.dsub(j) / 4.f;
      const int ns = (j == dr.jlast) ? dr.nlast : mv.nsub(j);
      // the propagators to the four interior knots depend on the layer only: formed once per layer, not once
      // per sub-layer like surfa.f:524-552 does (same operations, same results)
      float yk[5], zk[5], ck[5];
#pragma unroll
      for (int kk = 1; kk < 5; ++kk) {
        const float q = rb * dz * (float)kk;
        if (c < b) {
          const float exqp = expf(q), exqm = 1.f / exqp;
          yk[kk] = (exqp - exqm) / (2.f * rb); zk[kk] = rb * rb * yk[kk]; ck[kk] = (exqp + exqm) / 2.f;
        } else if (c == b) {
          yk[kk] = dz * (float)kk; zk[kk] = 0.f; ck[kk] = 1.f;
        } else {
          float sn, cs;
          sincosf(q, &sn, &cs);
          yk[kk] = sn / rb; zk[kk] = -rb * sn; ck[kk] = cs;
        }
      }
      for (int s = 0; s < ns; ++s) {
        if (fabsf(ut) > 1.e10f) { ut0 = ut0 / 1.e5f; restart = true; break; }  // surfa.f:519-522
        float dmm[5];
        dmm[0] = ut * ut;
        float eut = ut, ett = tq;
#pragma unroll
        for (int kk = 1; kk < 5; ++kk) {
          eut = ck[kk] * ut - yk[kk] * tq / h;
          ett = -h * zk[kk] * ut + ck[kk] * tq;
          dmm[kk] = eut * eut;
        }
        ut = eut; tq = ett;
        dm = (dz / 22.5f) * (7.f * (dmm[0] + dmm[4]) + 32.f * (dmm[1] + dmm[3]) + 12.f * dmm[2]);
        sumi0 = sumi0 + rho * dm;
        sumi1 = sumi1 + h * dm;
      }
      nsubsteps += (unsigned)ns;
    }
    if (!restart) break;
  }
  return sumi1 / (c * sumi0);  // surfa.f:606 (the 1/ut^2 normalisation cancels)
}


}  // namespace sd
