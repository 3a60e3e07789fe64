// mm_libm.cuh -- bit-exact device/host restatement of the float math the reference reaches
// through Eigen 3.3.7 (SSE packet sin/cos) and through glibc 2.39 libm (std::sin/cos/tan/
// atan/atan2/acos/asin on float).  Every routine is written with one IEEE operation per
// statement (no contraction) so that the B200 kernels land on exactly the same 1/16-pel
// position as the reference CPU build.  SURVEY.md section 8(c) risks R1/R6.
//
// What is restated, and how it was pinned:
//   * mm_sinf/mm_cosf     glibc 2.39 sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h.
//                         x86_64 selects the FMA build (__sinf_fma/__cosf_fma) when the
//                         host CPU has FMA+AVX2; `fma_build` reproduces that contraction
//                         pattern, read off this box's libm.so.6 (oracle/tools/libm_consts.py).
//   * mm_atanf, mm_atan2f, mm_acosf, mm_asinf, mm_tanf
//                         glibc 2.39 flt-32 s_atanf.c, e_atan2f.c, e_acosf.c, e_asinf.c,
//                         s_tanf.c + k_tanf.c (fdlibm float code, no FMA).
//   * mm_psin/mm_pcos     Eigen 3.3.7 Eigen/src/Core/arch/SSE/MathFunctions.h psin/pcos<Packet4f>
//                         (Cephes), one lane.  Eigen is not on this box: parity UNPINNED.
//   tests/native/libm_pin.cpp compares the glibc routines against the real libm of the box,
//   exhaustively over all 2^32 floats when asked.
//
// The same header is compiled by nvcc for sm_100a (device) and by g++ (host, for the pin test
// and for per-sequence setup code).  Constants are the bit patterns found in libm.so.6.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MM_HD __host__ __device__ __forceinline__
// the big libm routines: one copy each by default (k_map shrinks from 12k to 2.8k SASS instructions, which is what the
// instruction cache needs); -DMM_INLINE_LIBM inlines them at every call site
#ifndef MM_INLINE_LIBM
#define MM_HDL static __host__ __device__ __noinline__
#else
#define MM_HDL __host__ __device__ __forceinline__
#endif
#else
#define MM_HD static inline
#define MM_HDL static inline
#endif

#if defined(__CUDA_ARCH__)
#define MM_FMUL(a, b) __fmul_rn((a), (b))
#define MM_FADD(a, b) __fadd_rn((a), (b))
#define MM_FSUB(a, b) __fsub_rn((a), (b))
#define MM_FDIV(a, b) __fdiv_rn((a), (b))
#define MM_FSQRT(a) __fsqrt_rn((a))
#define MM_DMUL(a, b) __dmul_rn((a), (b))
#define MM_DADD(a, b) __dadd_rn((a), (b))
#define MM_DSUB(a, b) __dsub_rn((a), (b))
#define MM_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define MM_F2U(f) __float_as_uint((f))
#define MM_U2F(u) __uint_as_float((u))
#define MM_D2F(d) __double2float_rn((d))
#else
#include <math.h>
#include <string.h>
// Host build: compile with -ffp-contract=off (the Makefiles do) so these stay single ops.
#define MM_FMUL(a, b) ((float)(a) * (float)(b))
#define MM_FADD(a, b) ((float)(a) + (float)(b))
#define MM_FSUB(a, b) ((float)(a) - (float)(b))
#define MM_FDIV(a, b) ((float)(a) / (float)(b))
#define MM_FSQRT(a) __builtin_sqrtf((a))
#define MM_DMUL(a, b) ((double)(a) * (double)(b))
#define MM_DADD(a, b) ((double)(a) + (double)(b))
#define MM_DSUB(a, b) ((double)(a) - (double)(b))
#define MM_DFMA(a, b, c) __builtin_fma((a), (b), (c))
static inline uint32_t mm_f2u_host(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float mm_u2f_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define MM_F2U(f) mm_f2u_host((f))
#define MM_U2F(u) mm_u2f_host((u))
#define MM_D2F(d) ((float)(d))
#endif

// ------------------------------------------------------------------------------------------
// Small helpers
// ------------------------------------------------------------------------------------------
MM_HD float mm_fabsf(float x) { return MM_U2F(MM_F2U(x) & 0x7fffffffu); }
MM_HD float mm_negf(float x) { return MM_U2F(MM_F2U(x) ^ 0x80000000u); }
MM_HD bool mm_isnanf(float x) { return (MM_F2U(x) & 0x7fffffffu) > 0x7f800000u; }

// x86 cvttss2si / cvttps2dq: truncate, and return the "integer indefinite" 0x80000000 for NaN
// and for values that do not fit (CUDA's conversion saturates instead).
MM_HD int32_t mm_cvtt_f32_i32(float x)
{
  const uint32_t ax = MM_F2U(x) & 0x7fffffffu;
  if (ax >= 0x4f000000u) return (int32_t)0x80000000u;  // |x| >= 2^31, inf, NaN
  return (int32_t)x;                                    // in range: C truncation == cvtt
}
MM_HD int32_t mm_cvtt_f64_i32(double x)
{
  if (!(x > -2147483649.0 && x < 2147483648.0)) return (int32_t)0x80000000u;
  return (int32_t)x;
}

// ------------------------------------------------------------------------------------------
// glibc 2.39 sinf / cosf (sincosf.h, s_sinf.c, s_cosf.c)
// ------------------------------------------------------------------------------------------
// __sincosf_table[0] / [1]: c0..c4 (cosine poly), s1..s3 (sine poly).  Table 1 negates c*.
#define MM_SC_HPI_INV 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define MM_SC_HPI 0x1.921FB54442D18p0
#define MM_SC_C0 0x1p0
#define MM_SC_C1 -0x1.ffffffd0c621cp-2
#define MM_SC_C2 0x1.55553e1068f19p-5
#define MM_SC_C3 -0x1.6c087e89a359dp-10
#define MM_SC_C4 0x1.99343027bf8c3p-16
#define MM_SC_S1 -0x1.555545995a603p-3
#define MM_SC_S2 0x1.1107605230bc4p-7
#define MM_SC_S3 -0x1.994eb3774cf24p-13
#define MM_SC_PI63 0x1.921FB54442D18p-62

// sinf_poly(): n even -> sine polynomial, n odd -> cosine polynomial.  `neg` selects
// __sincosf_table[1] (cosine coefficients negated).  With fma_build every a+b*c below is a
// fused multiply-add exactly where __sinf_fma/__cosf_fma have vfmadd (libm.so.6 0x7e800/0x7e330).
MM_HD float mm_sincosf_poly(double x, double x2, bool neg, int n, bool fma_build)
{
  if ((n & 1) == 0)
  {
    const double x3 = MM_DMUL(x, x2);
    double s1, s, r;
    const double x7 = MM_DMUL(x3, x2);
    if (fma_build)
    {
      s1 = MM_DFMA(x2, MM_SC_S3, MM_SC_S2);
      s = MM_DFMA(x3, MM_SC_S1, x);
      r = MM_DFMA(x7, s1, s);
    }
    else
    {
      s1 = MM_DADD(MM_SC_S2, MM_DMUL(x2, MM_SC_S3));
      s = MM_DADD(x, MM_DMUL(x3, MM_SC_S1));
      r = MM_DADD(s, MM_DMUL(x7, s1));
    }
    return MM_D2F(r);
  }
  else
  {
    const double sg = neg ? -1.0 : 1.0;  // table[1] = -table[0] for c0..c4 (exact)
    const double c0 = sg * MM_SC_C0, c1c = sg * MM_SC_C1, c2c = sg * MM_SC_C2, c3c = sg * MM_SC_C3,
                 c4c = sg * MM_SC_C4;
    const double x4 = MM_DMUL(x2, x2);
    const double x6 = MM_DMUL(x4, x2);
    double c2, c1, c, r;
    if (fma_build)
    {
      c2 = MM_DFMA(x2, c4c, c3c);
      c1 = MM_DFMA(x2, c1c, c0);
      c = MM_DFMA(x4, c2c, c1);
      r = MM_DFMA(x6, c2, c);
    }
    else
    {
      c2 = MM_DADD(c3c, MM_DMUL(x2, c4c));
      c1 = MM_DADD(c0, MM_DMUL(x2, c1c));
      c = MM_DADD(c1, MM_DMUL(x4, c2c));
      r = MM_DADD(c, MM_DMUL(x6, c2));
    }
    return MM_D2F(r);
  }
}

// __inv_pio4[24] (sincosf_data.c), as found at libm.so.6 0xb80c0.
MM_HD uint32_t mm_inv_pio4(int i)
{
  switch (i)
  {
    case 0: return 0xa2u;
    case 1: return 0xa2f9u;
    case 2: return 0xa2f983u;
    case 3: return 0xa2f9836eu;
    case 4: return 0xf9836e4eu;
    case 5: return 0x836e4e44u;
    case 6: return 0x6e4e4415u;
    case 7: return 0x4e441529u;
    case 8: return 0x441529fcu;
    case 9: return 0x1529fc27u;
    case 10: return 0x29fc2757u;
    case 11: return 0xfc2757d1u;
    case 12: return 0x2757d1f5u;
    case 13: return 0x57d1f534u;
    case 14: return 0xd1f534ddu;
    case 15: return 0xf534ddc0u;
    case 16: return 0x34ddc0dbu;
    case 17: return 0xddc0db62u;
    case 18: return 0xc0db6295u;
    case 19: return 0xdb629599u;
    case 20: return 0x6295993cu;
    case 21: return 0x95993c43u;
    case 22: return 0x993c4390u;
    default: return 0x3c439041u;
  }
}

// reduce_fast(): |x| < 120.  n = round(x * 2/pi), returns x - n*pi/2 (double).
MM_HD double mm_sc_reduce_fast(double x, int* np, bool fma_build)
{
  const double r = MM_DMUL(x, MM_SC_HPI_INV);
  const int n = (int)(((uint32_t)mm_cvtt_f64_i32(r) + 0x800000u)) >> 24;
  *np = n;
  if (fma_build) return MM_DFMA(-(double)n, MM_SC_HPI, x);
  return MM_DSUB(x, MM_DMUL((double)n, MM_SC_HPI));
}

// reduce_large(): 120 <= |x| < inf, on the bit pattern of x.
MM_HD double mm_sc_reduce_large(uint32_t xi, int* np)
{
  const int idx = (int)((xi >> 26) & 15u);
  const int shift = (int)((xi >> 23) & 7u);
  uint64_t n, res0, res1, res2;
  xi = (xi & 0xffffffu) | 0x800000u;
  xi <<= shift;
  res0 = (uint32_t)(xi * mm_inv_pio4(idx));
  res1 = (uint64_t)xi * mm_inv_pio4(idx + 4);
  res2 = (uint64_t)xi * mm_inv_pio4(idx + 8);
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  const double x = (double)(int64_t)res0;
  *np = (int)n;
  return MM_DMUL(x, MM_SC_PI63);
}

MM_HD uint32_t mm_abstop12(float x) { return (MM_F2U(x) >> 20) & 0x7ffu; }

MM_HDL float mm_sinf(float y, bool fma_build)
{
  double x = (double)y;
  const uint32_t top = mm_abstop12(y);
  int n;
  if (top < 0x3f4u)  // |y| < pi/4
  {
    const double s = MM_DMUL(x, x);
    if (top < 0x398u) return y;  // |y| < 2^-12
    return mm_sincosf_poly(x, s, false, 0, fma_build);
  }
  else if (top < 0x42fu)  // |y| < 120
  {
    x = mm_sc_reduce_fast(x, &n, fma_build);
    const int q = n & 3;
    const double s = (q == 1 || q == 2) ? -1.0 : 1.0;  // sign[4] = {1,-1,-1,1}
    return mm_sincosf_poly(MM_DMUL(x, s), MM_DMUL(x, x), (n & 2) != 0, n, fma_build);
  }
  else if (top < 0x7f8u)
  {
    const uint32_t xi = MM_F2U(y);
    const int sign = (int)(xi >> 31);
    x = mm_sc_reduce_large(xi, &n);
    const int q = (n + sign) & 3;
    const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
    return mm_sincosf_poly(MM_DMUL(x, s), MM_DMUL(x, x), ((n + sign) & 2) != 0, n, fma_build);
  }
  return MM_FDIV(MM_FSUB(y, y), MM_FSUB(y, y));  // __math_invalidf: NaN
}

MM_HDL float mm_cosf(float y, bool fma_build)
{
  double x = (double)y;
  const uint32_t top = mm_abstop12(y);
  int n;
  if (top < 0x3f4u)
  {
    const double x2 = MM_DMUL(x, x);
    if (top < 0x398u) return 1.0f;
    return mm_sincosf_poly(x, x2, false, 1, fma_build);
  }
  else if (top < 0x42fu)
  {
    x = mm_sc_reduce_fast(x, &n, fma_build);
    const int q = n & 3;
    const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
    return mm_sincosf_poly(MM_DMUL(x, s), MM_DMUL(x, x), (n & 2) != 0, n ^ 1, fma_build);
  }
  else if (top < 0x7f8u)
  {
    const uint32_t xi = MM_F2U(y);
    const int sign = (int)(xi >> 31);
    x = mm_sc_reduce_large(xi, &n);
    const int q = (n + sign) & 3;
    const double s = (q == 1 || q == 2) ? -1.0 : 1.0;
    return mm_sincosf_poly(MM_DMUL(x, s), MM_DMUL(x, x), ((n + sign) & 2) != 0, n ^ 1, fma_build);
  }
  return MM_FDIV(MM_FSUB(y, y), MM_FSUB(y, y));
}

// ------------------------------------------------------------------------------------------
// glibc 2.39 flt-32/s_atanf.c
// ------------------------------------------------------------------------------------------
MM_HDL float mm_atanf(float x)
{
  const float atanhi0 = MM_U2F(0x3eed6338u), atanhi1 = MM_U2F(0x3f490fdau), atanhi2 = MM_U2F(0x3f7b985eu),
              atanhi3 = MM_U2F(0x3fc90fdau);
  const float atanlo0 = MM_U2F(0x31ac3769u), atanlo1 = MM_U2F(0x33222168u), atanlo2 = MM_U2F(0x33140fb4u),
              atanlo3 = MM_U2F(0x33a22168u);
  const float aT0 = MM_U2F(0x3eaaaaabu), aT1 = MM_U2F(0xbe4ccccdu), aT2 = MM_U2F(0x3e124925u),
              aT3 = MM_U2F(0xbde38e38u), aT4 = MM_U2F(0x3dba2e6eu), aT5 = MM_U2F(0xbd9d8795u),
              aT6 = MM_U2F(0x3d886b35u), aT7 = MM_U2F(0xbd6ef16bu), aT8 = MM_U2F(0x3d4bda59u),
              aT9 = MM_U2F(0xbd15a221u), aT10 = MM_U2F(0x3c8569d7u);
  const int32_t hx = (int32_t)MM_F2U(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  float hi = 0.f, lo = 0.f;
  if (ix >= 0x4c000000)  // |x| >= 2^25
  {
    if (ix > 0x7f800000) return MM_FADD(x, x);  // NaN
    if (hx > 0) return MM_FADD(atanhi3, atanlo3);
    return MM_FSUB(mm_negf(atanhi3), atanlo3);
  }
  if (ix < 0x3ee00000)  // |x| < 0.4375
  {
    if (ix < 0x31000000) return x;  // |x| < 2^-29 (huge+x>one always holds)
    id = -1;
  }
  else
  {
    x = mm_fabsf(x);
    if (ix < 0x3f980000)  // |x| < 1.1875
    {
      if (ix < 0x3f300000)
      {
        id = 0; hi = atanhi0; lo = atanlo0;
        x = MM_FDIV(MM_FSUB(MM_FADD(x, x), 1.0f), MM_FADD(2.0f, x));
      }
      else
      {
        id = 1; hi = atanhi1; lo = atanlo1;
        x = MM_FDIV(MM_FSUB(x, 1.0f), MM_FADD(x, 1.0f));
      }
    }
    else
    {
      if (ix < 0x401c0000)  // |x| < 2.4375
      {
        id = 2; hi = atanhi2; lo = atanlo2;
        x = MM_FDIV(MM_FSUB(x, 1.5f), MM_FADD(1.0f, MM_FMUL(1.5f, x)));
      }
      else
      {
        id = 3; hi = atanhi3; lo = atanlo3;
        x = MM_FDIV(-1.0f, x);
      }
    }
  }
  const float z = MM_FMUL(x, x);
  const float w = MM_FMUL(z, z);
  float s1 = MM_FADD(aT8, MM_FMUL(w, aT10));
  s1 = MM_FADD(aT6, MM_FMUL(w, s1));
  s1 = MM_FADD(aT4, MM_FMUL(w, s1));
  s1 = MM_FADD(aT2, MM_FMUL(w, s1));
  s1 = MM_FADD(aT0, MM_FMUL(w, s1));
  s1 = MM_FMUL(z, s1);
  float s2 = MM_FADD(aT7, MM_FMUL(w, aT9));
  s2 = MM_FADD(aT5, MM_FMUL(w, s2));
  s2 = MM_FADD(aT3, MM_FMUL(w, s2));
  s2 = MM_FADD(aT1, MM_FMUL(w, s2));
  s2 = MM_FMUL(w, s2);
  const float xs = MM_FMUL(x, MM_FADD(s1, s2));
  if (id < 0) return MM_FSUB(x, xs);
  const float r = MM_FSUB(hi, MM_FSUB(MM_FSUB(xs, lo), x));
  return (hx < 0) ? mm_negf(r) : r;
}

// ------------------------------------------------------------------------------------------
// glibc 2.39 flt-32/e_atan2f.c (the w_atan2f wrapper only touches errno)
// ------------------------------------------------------------------------------------------
MM_HDL float mm_atan2f_full(float y, float x)
{
  const float tiny = MM_U2F(0x0da24260u);
  const float pi_o_4 = MM_U2F(0x3f490fdbu), pi_o_2 = MM_U2F(0x3fc90fdbu), pi = MM_U2F(0x40490fdbu),
              pi_lo = MM_U2F(0xb3bbbd2eu);
  const int32_t hx = (int32_t)MM_F2U(x), hy = (int32_t)MM_F2U(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return MM_FADD(x, y);  // NaN
  if (hx == 0x3f800000) return mm_atanf(y);                       // x == 1.0
  const int m = (int)(((uint32_t)hy >> 31) & 1u) | (int)(((uint32_t)hx >> 30) & 2u);
  if (iy == 0)
  {
    switch (m)
    {
      case 0:
      case 1: return y;
      case 2: return MM_FADD(pi, tiny);
      default: return MM_FSUB(mm_negf(pi), tiny);
    }
  }
  if (ix == 0) return (hy < 0) ? MM_FSUB(mm_negf(pi_o_2), tiny) : MM_FADD(pi_o_2, tiny);
  if (ix == 0x7f800000)
  {
    if (iy == 0x7f800000)
    {
      switch (m)
      {
        case 0: return MM_FADD(pi_o_4, tiny);
        case 1: return MM_FSUB(mm_negf(pi_o_4), tiny);
        case 2: return MM_FADD(MM_FMUL(3.0f, pi_o_4), tiny);
        default: return MM_FSUB(MM_FMUL(-3.0f, pi_o_4), tiny);
      }
    }
    else
    {
      switch (m)
      {
        case 0: return 0.0f;
        case 1: return MM_U2F(0x80000000u);
        case 2: return MM_FADD(pi, tiny);
        default: return MM_FSUB(mm_negf(pi), tiny);
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? MM_FSUB(mm_negf(pi_o_2), tiny) : MM_FADD(pi_o_2, tiny);
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = MM_FADD(pi_o_2, MM_FMUL(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = mm_atanf(mm_fabsf(MM_FDIV(y, x)));
  switch (m)
  {
    case 0: return z;
    case 1: return mm_negf(z);
    case 2: return MM_FSUB(pi, MM_FSUB(z, pi_lo));
    default: return MM_FSUB(MM_FSUB(z, pi_lo), pi);
  }
}

// Common case first: x and y finite, non-zero, x != 1 and |exponent difference| <= 60 take exactly the statements
// `z = atanf(|y/x|)` + the quadrant fix-up of the code above (same operations, same order); everything else goes through
// the complete case analysis.  One test instead of the cascade: the cascade was ~40 of the ~50 instructions per call.
MM_HD float mm_atan2f(float y, float x)
{
  const int32_t hx = (int32_t)MM_F2U(x), hy = (int32_t)MM_F2U(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  const int32_t k = (iy - ix) >> 23;
  const bool common = (uint32_t)(ix - 1) < 0x7f7fffffu && (uint32_t)(iy - 1) < 0x7f7fffffu && hx != 0x3f800000 &&
                      (uint32_t)(k + 60) <= 120u;
  if (!common) return mm_atan2f_full(y, x);
  const float pi = MM_U2F(0x40490fdbu), pi_lo = MM_U2F(0xb3bbbd2eu);
  const float z = mm_atanf(mm_fabsf(MM_FDIV(y, x)));
  if (hx >= 0) return (hy < 0) ? mm_negf(z) : z;
  const float t = MM_FSUB(z, pi_lo);
  return (hy < 0) ? MM_FSUB(t, pi) : MM_FSUB(pi, t);
}

// ------------------------------------------------------------------------------------------
// glibc 2.39 flt-32/e_acosf.c
// ------------------------------------------------------------------------------------------
MM_HD float mm_acos_pq(float z)
{
  const float pS0 = MM_U2F(0x3e2aaaabu), pS1 = MM_U2F(0xbea6b090u), pS2 = MM_U2F(0x3e4e0aa8u),
              pS3 = MM_U2F(0xbd241146u), pS4 = MM_U2F(0x3a4f7f04u), pS5 = MM_U2F(0x3811ef08u);
  const float qS1 = MM_U2F(0xc019d139u), qS2 = MM_U2F(0x4001572du), qS3 = MM_U2F(0xbf303361u),
              qS4 = MM_U2F(0x3d9dc62eu);
  float p = MM_FADD(pS4, MM_FMUL(z, pS5));
  p = MM_FADD(pS3, MM_FMUL(z, p));
  p = MM_FADD(pS2, MM_FMUL(z, p));
  p = MM_FADD(pS1, MM_FMUL(z, p));
  p = MM_FADD(pS0, MM_FMUL(z, p));
  p = MM_FMUL(z, p);
  float q = MM_FADD(qS3, MM_FMUL(z, qS4));
  q = MM_FADD(qS2, MM_FMUL(z, q));
  q = MM_FADD(qS1, MM_FMUL(z, q));
  q = MM_FADD(1.0f, MM_FMUL(z, q));
  return MM_FDIV(p, q);
}

MM_HDL float mm_acosf(float x)
{
  const float pio2_hi = MM_U2F(0x3fc90fdau), pio2_lo = MM_U2F(0x33a22168u), pi = MM_U2F(0x40490fdau);
  const int32_t hx = (int32_t)MM_F2U(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix == 0x3f800000)
  {
    if (hx > 0) return 0.0f;
    return MM_FADD(pi, MM_FMUL(2.0f, pio2_lo));
  }
  else if (ix > 0x3f800000)
  {
    return MM_FDIV(MM_FSUB(x, x), MM_FSUB(x, x));  // NaN
  }
  if (ix < 0x3f000000)  // |x| < 0.5
  {
    if (ix <= 0x32800000) return MM_FADD(pio2_hi, pio2_lo);
    const float z = MM_FMUL(x, x);
    const float r = mm_acos_pq(z);
    return MM_FSUB(pio2_hi, MM_FSUB(x, MM_FSUB(pio2_lo, MM_FMUL(r, x))));
  }
  else if (hx < 0)  // x < -0.5
  {
    const float z = MM_FMUL(MM_FADD(1.0f, x), 0.5f);
    const float r = mm_acos_pq(z);
    const float s = MM_FSQRT(z);
    const float w = MM_FSUB(MM_FMUL(r, s), pio2_lo);
    return MM_FSUB(pi, MM_FMUL(2.0f, MM_FADD(s, w)));
  }
  else  // x > 0.5
  {
    const float z = MM_FMUL(MM_FSUB(1.0f, x), 0.5f);
    const float s = MM_FSQRT(z);
    const float df = MM_U2F(MM_F2U(s) & 0xfffff000u);
    const float c = MM_FDIV(MM_FSUB(z, MM_FMUL(df, df)), MM_FADD(s, df));
    const float r = mm_acos_pq(z);
    const float w = MM_FADD(MM_FMUL(r, s), c);
    return MM_FMUL(2.0f, MM_FADD(df, w));
  }
}

// ------------------------------------------------------------------------------------------
// glibc 2.39 flt-32/e_asinf.c (glibc's own polynomial, not fdlibm's rational)
// ------------------------------------------------------------------------------------------
MM_HD float mm_asin_poly(float t)
{
  const float p0 = MM_U2F(0x3e2aaae4u), p1 = MM_U2F(0x3d9980f2u), p2 = MM_U2F(0x3d3a3f25u),
              p3 = MM_U2F(0x3cc6141eu), p4 = MM_U2F(0x3d2cb694u);
  float w = MM_FADD(p3, MM_FMUL(t, p4));
  w = MM_FADD(p2, MM_FMUL(t, w));
  w = MM_FADD(p1, MM_FMUL(t, w));
  w = MM_FADD(p0, MM_FMUL(t, w));
  return MM_FMUL(t, w);
}

MM_HDL float mm_asinf(float x)
{
  const float pio2_hi = MM_U2F(0x3fc90fdbu), pio2_lo = MM_U2F(0xb33bbd2eu), pio4_hi = MM_U2F(0x3f490fdbu);
  const int32_t hx = (int32_t)MM_F2U(x);
  const int32_t ix = hx & 0x7fffffff;
  float t;
  if (ix == 0x3f800000) return MM_FADD(MM_FMUL(x, pio2_hi), MM_FMUL(x, pio2_lo));
  if (ix > 0x3f800000) return MM_FDIV(MM_FSUB(x, x), MM_FSUB(x, x));
  if (ix < 0x3f000000)
  {
    if (ix < 0x32000000) return x;
    t = MM_FMUL(x, x);
    const float w = mm_asin_poly(t);
    return MM_FADD(x, MM_FMUL(x, w));
  }
  float w = MM_FSUB(1.0f, mm_fabsf(x));
  t = MM_FMUL(w, 0.5f);
  float p = mm_asin_poly(t);
  const float s = MM_FSQRT(t);
  if (ix >= 0x3f79999a)
  {
    t = MM_FSUB(pio2_hi, MM_FSUB(MM_FMUL(2.0f, MM_FADD(s, MM_FMUL(s, p))), pio2_lo));
  }
  else
  {
    w = MM_U2F(MM_F2U(s) & 0xfffff000u);
    const float c = MM_FDIV(MM_FSUB(t, MM_FMUL(w, w)), MM_FADD(s, w));
    const float r = p;
    p = MM_FSUB(MM_FMUL(MM_FMUL(2.0f, s), r), MM_FSUB(pio2_lo, MM_FMUL(2.0f, c)));
    const float q = MM_FSUB(pio4_hi, MM_FMUL(2.0f, w));
    t = MM_FSUB(pio4_hi, MM_FSUB(p, q));
  }
  return (hx > 0) ? t : mm_negf(t);
}

// ------------------------------------------------------------------------------------------
// glibc 2.39 flt-32/s_tanf.c + k_tanf.c (reduction shared with sincosf, generic build: no FMA)
// ------------------------------------------------------------------------------------------
MM_HD float mm_kernel_tanf(float x, float y, int iy)
{
  const float pio4 = MM_U2F(0x3f490fdau), pio4lo = MM_U2F(0x33222168u);
  const float T0 = MM_U2F(0x3eaaaaabu), T1 = MM_U2F(0x3e088889u), T2 = MM_U2F(0x3d5d0dd1u),
              T3 = MM_U2F(0x3cb327a4u), T4 = MM_U2F(0x3c11371fu), T5 = MM_U2F(0x3b6b6916u),
              T6 = MM_U2F(0x3abede48u), T7 = MM_U2F(0x3a1a26c8u), T8 = MM_U2F(0x398137b9u),
              T9 = MM_U2F(0x38a3f445u), T10 = MM_U2F(0x3895c07au), T11 = MM_U2F(0xb79bae5fu),
              T12 = MM_U2F(0x37d95384u);
  const int32_t hx = (int32_t)MM_F2U(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix < 0x39000000)  // |x| < 2^-13
  {
    if (mm_cvtt_f32_i32(x) == 0)
    {
      if ((ix | (iy + 1)) == 0) return MM_FDIV(1.0f, mm_fabsf(x));
      else if (iy == 1) return x;
      else return MM_FDIV(-1.0f, x);
    }
  }
  if (ix >= 0x3f2ca140)  // |x| >= 0.6744
  {
    if (hx < 0) { x = mm_negf(x); y = mm_negf(y); }
    const float zz = MM_FSUB(pio4, x);
    const float ww = MM_FSUB(pio4lo, y);
    x = MM_FADD(zz, ww);
    y = 0.0f;
    if (mm_fabsf(x) < MM_U2F(0x39000000u))
    {
      const int sg = (1 - ((hx >> 30) & 2)) * iy;
      return MM_FMUL((float)sg, MM_FSUB(1.0f, MM_FMUL((float)(2 * iy), x)));
    }
  }
  const float z = MM_FMUL(x, x);
  float w = MM_FMUL(z, z);
  float r = MM_FADD(T9, MM_FMUL(w, T11));
  r = MM_FADD(T7, MM_FMUL(w, r));
  r = MM_FADD(T5, MM_FMUL(w, r));
  r = MM_FADD(T3, MM_FMUL(w, r));
  r = MM_FADD(T1, MM_FMUL(w, r));
  float v = MM_FADD(T10, MM_FMUL(w, T12));
  v = MM_FADD(T8, MM_FMUL(w, v));
  v = MM_FADD(T6, MM_FMUL(w, v));
  v = MM_FADD(T4, MM_FMUL(w, v));
  v = MM_FADD(T2, MM_FMUL(w, v));
  v = MM_FMUL(z, v);
  float s = MM_FMUL(z, x);
  r = MM_FADD(y, MM_FMUL(z, MM_FADD(MM_FMUL(s, MM_FADD(r, v)), y)));
  r = MM_FADD(r, MM_FMUL(T0, s));
  w = MM_FADD(x, r);
  if (ix >= 0x3f2ca140)
  {
    v = (float)iy;
    const float sgn = (float)(1 - ((hx >> 30) & 2));
    const float inner = MM_FSUB(x, MM_FSUB(MM_FDIV(MM_FMUL(w, w), MM_FADD(w, v)), r));
    return MM_FMUL(sgn, MM_FSUB(v, MM_FMUL(2.0f, inner)));
  }
  if (iy == 1) return w;
  // -1/(x+r), accurately
  const float zt = MM_U2F(MM_F2U(w) & 0xfffff000u);
  v = MM_FSUB(r, MM_FSUB(zt, x));
  const float a = MM_FDIV(-1.0f, w);
  const float t = MM_U2F(MM_F2U(a) & 0xfffff000u);
  s = MM_FADD(1.0f, MM_FMUL(t, zt));
  return MM_FADD(t, MM_FMUL(a, MM_FADD(s, MM_FMUL(t, v))));
}

MM_HDL float mm_tanf(float x)
{
  const int32_t ix = (int32_t)(MM_F2U(x) & 0x7fffffffu);
  if (ix <= 0x3f490fda) return mm_kernel_tanf(x, 0.0f, 1);
  if (ix >= 0x7f800000) return MM_FSUB(x, x);  // NaN
  // rem_pio2f(): double reduction shared with sincosf (generic build: mul then sub)
  double dx = (double)x;
  int n;
  if (mm_abstop12(x) < 0x42fu)
  {
    dx = mm_sc_reduce_fast(dx, &n, false);
  }
  else
  {
    const uint32_t xi = MM_F2U(x);
    dx = mm_sc_reduce_large(xi, &n);
    if (xi >> 31) dx = -dx;
  }
  const float y0 = MM_D2F(dx);
  const float y1 = MM_D2F(MM_DSUB(dx, (double)y0));
  return mm_kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}

// ------------------------------------------------------------------------------------------
// Eigen 3.3.7 SSE psin/pcos<Packet4f>, one lane (Cephes sinf/cosf; mul and add are separate
// roundings because the reference is built with -msse4.1 only, pmadd = pmul then padd).
// Used by Eigen for Array<float>::sin()/cos() on the vectorised part of an expression.
// ------------------------------------------------------------------------------------------
MM_HD void mm_ps_core(float ax, int32_t* j_out, float* ys_out, float* yc_out)
{
  const float FOPI = 1.27323954473516f;
  const float DP1 = -0.78515625f, DP2 = -2.4187564849853515625e-4f, DP3 = -3.77489497744594108e-8f;
  const float sincof_p0 = -1.9515295891E-4f, sincof_p1 = 8.3321608736E-3f, sincof_p2 = -1.6666654611E-1f;
  const float coscof_p0 = 2.443315711809948E-005f, coscof_p1 = -1.388731625493765E-003f,
              coscof_p2 = 4.166664568298827E-002f;
  float y = MM_FMUL(ax, FOPI);
  int32_t j = mm_cvtt_f32_i32(y);
  j = (int32_t)(((uint32_t)j + 1u) & ~1u);
  y = (float)j;
  float x = ax;
  const float x1 = MM_FMUL(y, DP1), x2 = MM_FMUL(y, DP2), x3 = MM_FMUL(y, DP3);
  x = MM_FADD(x, x1);
  x = MM_FADD(x, x2);
  x = MM_FADD(x, x3);
  const float z = MM_FMUL(x, x);
  float yc = coscof_p0;
  yc = MM_FADD(MM_FMUL(yc, z), coscof_p1);
  yc = MM_FADD(MM_FMUL(yc, z), coscof_p2);
  yc = MM_FMUL(yc, z);
  yc = MM_FMUL(yc, z);
  const float tmp = MM_FMUL(z, 0.5f);
  yc = MM_FSUB(yc, tmp);
  yc = MM_FADD(yc, 1.0f);
  float ys = sincof_p0;
  ys = MM_FADD(MM_FMUL(ys, z), sincof_p1);
  ys = MM_FADD(MM_FMUL(ys, z), sincof_p2);
  ys = MM_FMUL(ys, z);
  ys = MM_FMUL(ys, x);
  ys = MM_FADD(ys, x);
  *j_out = j;
  *ys_out = ys;
  *yc_out = yc;
}

MM_HD float mm_psin(float x)
{
  const uint32_t sign = MM_F2U(x) & 0x80000000u;
  int32_t j;
  float ys, yc;
  mm_ps_core(mm_fabsf(x), &j, &ys, &yc);
  const uint32_t swap = ((uint32_t)j & 4u) << 29;
  const bool poly_sin = (((uint32_t)j & 2u) == 0u);
  const float r = poly_sin ? ys : yc;
  return MM_U2F(MM_F2U(r) ^ (sign ^ swap));
}

MM_HD float mm_pcos(float x)
{
  int32_t j;
  float ys, yc;
  mm_ps_core(mm_fabsf(x), &j, &ys, &yc);
  const uint32_t j2 = (uint32_t)j - 2u;
  const uint32_t sign = ((~j2) & 4u) << 29;
  const bool poly_sin = ((j2 & 2u) == 0u);
  const float r = poly_sin ? ys : yc;
  return MM_U2F(MM_F2U(r) ^ sign);
}

// std::round(float): half away from zero (glibc roundf), used by Eigen's scalar round().
// t = trunc(x) and d = x - t are exact (|d| < 1, or 0 once |x| >= 2^23), so the result is t, or t +- 1 when |d| >= 0.5;
// NaN and infinities fall through the comparison and return trunc(x) = x.  Pinned exhaustively against glibc's roundf
// like the other functions (tests/native/libm_pin.cpp).
MM_HD float mm_roundf(float x)
{
#if defined(__CUDA_ARCH__)
  const float t = truncf(x);
#else
  const float t = __builtin_truncf(x);
#endif
  const float d = MM_FSUB(x, t);
  const float one = MM_U2F((MM_F2U(x) & 0x80000000u) | 0x3f800000u);
  return (mm_fabsf(d) >= 0.5f) ? MM_FADD(t, one) : t;
}

// ------------------------------------------------------------------------------------------
// float(sin((double)x)), float(cos((double)x)) for |x| <= pi/2 -- the two block constants of
// TangentialMotionModel.cpp:27-28,38,40, where the unqualified sin/cos bind to the DOUBLE libm
// overloads.  Host: glibc.  Device: CUDA's double sin/cos (<= 2 ulp) agree with glibc's
// (<= 1 ulp) after narrowing unless the double result lies within a few ulp of a float rounding
// breakpoint; exactly those inputs -- found by sweeping all 2^31 floats of the domain on the host
// (oracle/tools/tan_exceptions.cpp, guard 8 ulp) -- are answered from a table.  The GPU test
// test_tan_double_sincos_exhaustive re-checks the whole domain against the sweep's checksum.
// ------------------------------------------------------------------------------------------
#include "mm_tan_exceptions.h"
#if defined(__CUDA_ARCH__)
struct MmTanExc { uint32_t x, r; };
__device__ __forceinline__ bool mm_d2f_risky(double d)
{
  // would a value 4 ulp(double) below / above narrow to a different float?
  const long long b = __double_as_longlong(d);
  return __double2float_rn(__longlong_as_double(b - 4)) != __double2float_rn(__longlong_as_double(b + 4));
}
static __device__ __noinline__ float mm_tan_exc_lookup(const MmTanExc* t, int n, float x, float fallback)
{
  const uint32_t xb = __float_as_uint(x);
  for (int i = 0; i < n; ++i)
    if (t[i].x == xb) return __uint_as_float(t[i].r);
  return fallback;
}
__device__ __forceinline__ float mm_sin_d2f(float x)
{
  const MmTanExc exc[MM_TAN_EXC_SIN_N] = MM_TAN_EXC_SIN_INIT;
  const double d = sin((double)x);
  const float f = __double2float_rn(d);
  if (d != 0.0 && mm_d2f_risky(d)) return mm_tan_exc_lookup(exc, MM_TAN_EXC_SIN_N, x, f);
  return f;
}
__device__ __forceinline__ float mm_cos_d2f(float x)
{
  const MmTanExc exc[MM_TAN_EXC_COS_N] = MM_TAN_EXC_COS_INIT;
  const double d = cos((double)x);
  const float f = __double2float_rn(d);
  if (d != 0.0 && mm_d2f_risky(d)) return mm_tan_exc_lookup(exc, MM_TAN_EXC_COS_N, x, f);
  return f;
}
#else
static inline float mm_sin_d2f(float x) { return (float)sin((double)x); }
static inline float mm_cos_d2f(float x) { return (float)cos((double)x); }
#endif
