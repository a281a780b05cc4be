// sd_libm.cuh -- bit-exact restatement of the float32 libm functions the reference reaches through gfortran
// (flat1.f:44-68 -> alog / ** -> glibc logf / powf), for device code.
//
// Earth flattening in float32 amplifies a 1-ulp difference in r**p into ~3e-5 of a thin layer's density
// (r_i**p - r_(i+1)**p cancels 3-4 digits), i.e. up to 5e-5 km/s in c.  glibc's logf/powf are NOT correctly
// rounded (0.82 / 0.52 ulp), so "compute in double and round once" differs from them for a few per cent of the
// arguments; the only way to reproduce the reference's model preparation is to follow glibc's algorithm
// (glibc 2.27+ sysdeps/ieee754/flt-32/e_logf.c, e_powf.c, by Szabolcs Nagy: table-driven log2 / exp2 with
// double-precision polynomials; the same code is published as ARM optimized-routines).  The constants below
// are the ones of the libm.so.6 of this image (GLIBC 2.39); tests/test_libm_port.py checks the port against
// the host libm bit by bit over the whole argument range the path uses.  x86-64 glibc dispatches to the
// FMA build of these functions, hence the explicit fma() calls.
// Domain: finite positive normal x; |y log2 x| < 126.  (No special-case handling: the path never needs it.)
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SDM_HD __host__ __device__ __forceinline__
#define SDM_TAB __device__ __constant__
#else
#define SDM_HD inline
#define SDM_TAB
#endif

namespace sdm {
// {1/c, ln c} for the 16 sub-intervals of [0x3f330000, 2*0x3f330000)
#define SDM_KLOGFTAB_VALUES \
0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, \
  0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2, \
  0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, \
  0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3, \
  0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, \
  0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3, \
  0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, \
  0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4, \
  0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, \
  0x1.0000000000000p+0, 0x0.0p+0, \
  0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5, \
  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4, \
  0x1.b2036576afce6p-1, 0x1.526e57720db08p-3, \
  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3, \
  0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2, \
  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2,
// {1/c, log2 c}
#define SDM_KPOWLOG2TAB_VALUES \
0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, \
  0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2, \
  0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2, \
  0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2, \
  0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, \
  0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3, \
  0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, \
  0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4, \
  0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, \
  0x1.0000000000000p+0, 0x0.0p+0, \
  0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4, \
  0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3, \
  0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3, \
  0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2, \
  0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2, \
  0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2,
#define SDM_KEXP2TAB_VALUES \
0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
  0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
  0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
  0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
  0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
  0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
  0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
  0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL,
// logf: ln2, poly[3];  powf log2: poly[5];  exp2f: shift_scaled, poly[3]
#define SDM_LN2 0x1.62e42fefa39efp-1
#define SDM_LOGF_A0 -0x1.00ea348b88334p-2
#define SDM_LOGF_A1 0x1.5575b0be00b6ap-2
#define SDM_LOGF_A2 -0x1.ffffef20a4123p-2
#define SDM_PLOG_A0 0x1.27616c9496e0bp-2
#define SDM_PLOG_A1 -0x1.71969a075c67ap-2
#define SDM_PLOG_A2 0x1.ec70a6ca7baddp-2
#define SDM_PLOG_A3 -0x1.7154748bef6c8p-1
#define SDM_PLOG_A4 0x1.71547652ab82bp+0
#define SDM_EXP2_SHIFT 0x1.8000000000000p+47
#define SDM_EXP2_C0 0x1.c6af84b912394p-5
#define SDM_EXP2_C1 0x1.ebfce50fac4f3p-3
#define SDM_EXP2_C2 0x1.62e42ff0c52d6p-1

static const double kLogfTab_h[32] = {SDM_KLOGFTAB_VALUES};
static const double kPowLog2Tab_h[32] = {SDM_KPOWLOG2TAB_VALUES};
static const uint64_t kExp2Tab_h[32] = {SDM_KEXP2TAB_VALUES};
#if defined(__CUDACC__)
static __device__ __constant__ double kLogfTab_d[32] = {SDM_KLOGFTAB_VALUES};
static __device__ __constant__ double kPowLog2Tab_d[32] = {SDM_KPOWLOG2TAB_VALUES};
static __device__ __constant__ uint64_t kExp2Tab_d[32] = {SDM_KEXP2TAB_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define SDM_T(name) name##_d
#else
#define SDM_T(name) name##_h
#endif

SDM_HD uint32_t as_u32(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
SDM_HD float as_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
SDM_HD uint64_t as_u64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
SDM_HD double as_f64(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

#define SDM_OFF 0x3f330000u

// glibc logf (e_logf.c): x = 2^k z, z in [OFF, 2 OFF); log x = log1p(z/c - 1) + log c + k ln2
SDM_HD float logf_glibc(float x) {
  const uint32_t ix = as_u32(x);
  if (ix == 0x3f800000u) return 0.f;
  const uint32_t tmp = ix - SDM_OFF;
  const int i = (int)((tmp >> (23 - 4)) % 16u);
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & (0x1ffu << 23));
  const double invc = SDM_T(kLogfTab)[2 * i], logc = SDM_T(kLogfTab)[2 * i + 1];
  const double z = (double)as_f32(iz);
  const double r = fma(z, invc, -1.0);
  const double y0 = fma((double)k, SDM_LN2, logc);
  const double r2 = r * r;
  double y = fma(SDM_LOGF_A1, r, SDM_LOGF_A2);
  y = fma(SDM_LOGF_A0, r2, y);
  y = fma(y, r2, y0 + r);
  return (float)y;
}

// glibc powf (e_powf.c), x > 0 normal: 2^(y log2 x) with log2 by table + degree-5 polynomial and exp2 by
// table + cubic, all in double
SDM_HD float powf_glibc(float x, float yf) {
  const uint32_t ix = as_u32(x);
  const uint32_t tmp = ix - SDM_OFF;
  const int i = (int)((tmp >> (23 - 4)) % 16u);
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)top >> 23;
  const double invc = SDM_T(kPowLog2Tab)[2 * i], logc = SDM_T(kPowLog2Tab)[2 * i + 1];
  const double z = (double)as_f32(iz);
  const double r = fma(z, invc, -1.0);
  const double y0 = logc + (double)k;
  const double r2 = r * r;
  double y = fma(SDM_PLOG_A0, r, SDM_PLOG_A1);
  const double p = fma(SDM_PLOG_A2, r, SDM_PLOG_A3);
  const double r4 = r2 * r2;
  double q = fma(SDM_PLOG_A4, r, y0);
  q = fma(p, r2, q);
  y = fma(y, r4, q);
  const double xd = (double)yf * y;   // y log2 x
  // exp2: x = k/N + r, |r| <= 1/(2N), N = 32
  double kd = xd + SDM_EXP2_SHIFT;
  const uint64_t ki = as_u64(kd);
  kd -= SDM_EXP2_SHIFT;
  const double rr = xd - kd;
  uint64_t t = SDM_T(kExp2Tab)[ki % 32u];
  t += ki << (52 - 5);
  const double s = as_f64(t);
  const double zz = fma(SDM_EXP2_C0, rr, SDM_EXP2_C1);
  const double rr2 = rr * rr;
  double yy = fma(SDM_EXP2_C2, rr, 1.0);
  yy = fma(zz, rr2, yy);
  yy = yy * s;
  return (float)yy;
}

}  // namespace sdm
