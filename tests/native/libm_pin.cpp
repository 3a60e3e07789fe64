// libm_pin.cpp -- pins vvc-extension-mm_b200/csrc/mm_libm.cuh (the restated glibc 2.39 float
// routines the reference reaches through std::sin/cos/tan/atan/atan2/acos/asin) against the
// REAL libm of the box it runs on, bit for bit.
//
//   libm_pin <stride> [threads]     stride 1 = all 2^32 float bit patterns per unary function
// atan2f has two arguments: it is checked on a seeded sample plus the structured corner grid.
// Exit code 0 and "mismatches 0" on every line means the restatement is pinned.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <thread>
#include <vector>
#include <atomic>
#include "../../vvc-extension-mm_b200/csrc/mm_libm.cuh"

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline bool same(float a, float b)
{
  if (std::isnan(a) && std::isnan(b)) return true;
  return f2u(a) == f2u(b);
}

static bool g_fma;

typedef float (*fn1)(float);
static float w_sinf(float x) { return mm_sinf(x, g_fma); }
static float w_cosf(float x) { return mm_cosf(x, g_fma); }
static float w_atanf(float x) { return mm_atanf(x); }
static float w_acosf(float x) { return mm_acosf(x); }
static float w_asinf(float x) { return mm_asinf(x); }
static float w_tanf(float x) { return mm_tanf(x); }
static float w_roundf(float x) { return mm_roundf(x); }
// volatile function pointers keep gcc from folding or substituting the libm calls
static float (*volatile l_sinf)(float) = sinf;
static float (*volatile l_cosf)(float) = cosf;
static float (*volatile l_atanf)(float) = atanf;
static float (*volatile l_acosf)(float) = acosf;
static float (*volatile l_asinf)(float) = asinf;
static float (*volatile l_tanf)(float) = tanf;
static float (*volatile l_roundf)(float) = roundf;
static float (*volatile l_atan2f)(float, float) = atan2f;
// gcc -O3 (the reference's release build) merges sin(x)/cos(x) pairs into sincosf: its two results must equal sinf/cosf
static void (*volatile l_sincosf)(float, float*, float*) = sincosf;
static float w_sincos_s(float x) { float s, c; l_sincosf(x, &s, &c); return s; }
static float w_sincos_c(float x) { float s, c; l_sincosf(x, &s, &c); return c; }
static float (*volatile l_sincos_s)(float) = w_sincos_s;
static float (*volatile l_sincos_c)(float) = w_sincos_c;

static uint64_t run1(const char* name, fn1 mine, float (*volatile* ref)(float), uint64_t stride, int nthreads)
{
  std::atomic<uint64_t> bad(0), cnt(0);
  std::atomic<uint32_t> first(0);
  std::vector<std::thread> th;
  fn1 reff = *ref;
  for (int t = 0; t < nthreads; ++t)
  {
    th.emplace_back([&, t]() {
      uint64_t lb = 0, lc = 0;
      for (uint64_t u = (uint64_t)t * stride; u < (1ULL << 32); u += stride * nthreads)
      {
        const float x = u2f((uint32_t)u);
        const float a = mine(x), b = reff(x);
        ++lc;
        if (!same(a, b))
        {
          if (lb == 0 && first.load() == 0) first.store((uint32_t)u);
          ++lb;
        }
      }
      bad += lb;
      cnt += lc;
    });
  }
  for (auto& x : th) x.join();
  printf("%-8s checked %llu mismatches %llu", name, (unsigned long long)cnt.load(), (unsigned long long)bad.load());
  if (bad.load())
  {
    const float x = u2f(first.load());
    printf("  first x=%a (0x%08x) mine=%a libm=%a", x, first.load(), mine(x), reff(x));
  }
  printf("\n");
  return bad.load();
}

static uint64_t run_atan2(uint64_t samples, int nthreads)
{
  std::atomic<uint64_t> bad(0), cnt(0);
  std::vector<std::thread> th;
  // corner grid: signed zeros, denormals, ones, huge, inf, nan and exponent ladders
  std::vector<float> grid;
  const uint32_t specials[] = {0x00000000u, 0x80000000u, 0x00000001u, 0x80000001u, 0x007fffffu, 0x00800000u,
                               0x3f800000u, 0xbf800000u, 0x3f7fffffu, 0x3f800001u, 0x7f7fffffu, 0xff7fffffu,
                               0x7f800000u, 0xff800000u, 0x7fc00000u, 0x40490fdbu, 0x3fc90fdbu};
  for (uint32_t s : specials) grid.push_back(u2f(s));
  for (int e = 1; e < 255; e += 3)
    for (uint32_t m : {0u, 0x400000u, 0x7fffffu})
    {
      grid.push_back(u2f(((uint32_t)e << 23) | m));
      grid.push_back(u2f(0x80000000u | ((uint32_t)e << 23) | m));
    }
  uint64_t gb = 0;
  for (float y : grid)
    for (float x : grid)
    {
      if (!same(mm_atan2f(y, x), l_atan2f(y, x)))
      {
        if (gb == 0) printf("  atan2f grid mismatch y=%a x=%a mine=%a libm=%a\n", y, x, mm_atan2f(y, x), l_atan2f(y, x));
        ++gb;
      }
    }
  bad += gb;
  cnt += grid.size() * grid.size();
  for (int t = 0; t < nthreads; ++t)
  {
    th.emplace_back([&, t]() {
      uint64_t s = 0x9E3779B97F4A7C15ULL * (t + 1), lb = 0;
      auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
      for (uint64_t i = 0; i < samples / nthreads; ++i)
      {
        const uint64_t r = next();
        float y, x;
        if (i & 1)
        {  // unit-sphere-like magnitudes: what the path feeds atan2
          y = (float)((double)(int32_t)(r >> 32) / 2147483648.0);
          x = (float)((double)(int32_t)(r & 0xffffffffu) / 2147483648.0);
          if (i & 2) x *= 0x1p-12f;
          if (i & 4) y *= 0x1p-9f;
        }
        else
        {  // raw bit patterns
          y = u2f((uint32_t)(r >> 32));
          x = u2f((uint32_t)r);
        }
        if (!same(mm_atan2f(y, x), l_atan2f(y, x)))
        {
          if (lb == 0) printf("  atan2f mismatch y=%a x=%a mine=%a libm=%a\n", y, x, mm_atan2f(y, x), l_atan2f(y, x));
          ++lb;
        }
      }
      bad += lb;
      cnt += samples / nthreads;
    });
  }
  for (auto& x : th) x.join();
  printf("%-8s checked %llu mismatches %llu\n", "atan2f", (unsigned long long)cnt.load(), (unsigned long long)bad.load());
  return bad.load();
}

int main(int argc, char** argv)
{
  const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 4099;
  const int nthreads = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
  __builtin_cpu_init();
  g_fma = __builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2");
  printf("libm_pin: stride %llu threads %d host_fma_build %d\n", (unsigned long long)stride, nthreads, (int)g_fma);
  uint64_t bad = 0;
  bad += run1("sinf", w_sinf, &l_sinf, stride, nthreads);
  bad += run1("cosf", w_cosf, &l_cosf, stride, nthreads);
  bad += run1("sincosf.s", w_sinf, &l_sincos_s, stride, nthreads);
  bad += run1("sincosf.c", w_cosf, &l_sincos_c, stride, nthreads);
  bad += run1("atanf", w_atanf, &l_atanf, stride, nthreads);
  bad += run1("acosf", w_acosf, &l_acosf, stride, nthreads);
  bad += run1("asinf", w_asinf, &l_asinf, stride, nthreads);
  bad += run1("tanf", w_tanf, &l_tanf, stride, nthreads);
  bad += run1("roundf", w_roundf, &l_roundf, stride, nthreads);
  bad += run_atan2(stride == 1 ? 4000000000ULL : 40000000ULL, nthreads);
  printf("total mismatches %llu\n", (unsigned long long)bad);
  return bad ? 1 : 0;
}
