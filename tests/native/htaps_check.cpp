// htaps_check.cpp -- host check of the PRODUCT's coefficient-table builder (csrc/mm_htaps.h): for every alignment o,
// phase and output j, a scalar model of k_fir's six dp2a steps over the table words equals the plain 8-tap sum
// sum_t c[phase][t] * s[o + j + t] (InterpolationFilter::filter<8>, InterpolationFilter.cpp:540-645) on random rows, and
// entry o = 0 carries the vertical pass' (a0,a1 | b0,b1,b2) words.  Exit code 0 = all equal.
#include <stdio.h>
#include <stdlib.h>
#include "../../vvc-extension-mm_b200/csrc/mm_htaps.h"

static const int8_t kLuma[16][8] = {   // InterpolationFilter.cpp:82-100
  {0, 0, 0, 64, 0, 0, 0, 0},       {0, 1, -3, 63, 4, -2, 1, 0},      {-1, 2, -5, 62, 8, -3, 1, 0},
  {-1, 3, -8, 60, 13, -4, 1, 0},   {-1, 4, -10, 58, 17, -5, 1, 0},   {-1, 4, -11, 52, 26, -8, 3, -1},
  {-1, 3, -9, 47, 31, -10, 4, -1}, {-1, 4, -11, 45, 34, -10, 4, -1}, {-1, 4, -11, 40, 40, -11, 4, -1},
  {-1, 4, -10, 34, 45, -11, 4, -1}, {-1, 4, -10, 31, 47, -9, 3, -1}, {-1, 3, -8, 26, 52, -11, 4, -1},
  {0, 1, -5, 17, 58, -10, 4, -1},  {0, 1, -4, 13, 60, -8, 3, -1},    {0, 1, -3, 8, 62, -5, 2, -1},
  {0, 1, -2, 4, 63, -3, 1, 0}};

// dp2a: two signed 16-bit halves of `a` times the low (hi = 0) or high (hi = 1) pair of signed bytes of `b`
static int dp2a(uint32_t a, uint32_t b, int hi, int c)
{
  const int a0 = (int16_t)(a & 0xffff), a1 = (int16_t)(a >> 16);
  const int b0 = (int8_t)((b >> (hi ? 16 : 0)) & 0xff), b1 = (int8_t)((b >> (hi ? 24 : 8)) & 0xff);
  return c + a0 * b0 + a1 * b1;
}

int main()
{
  uint32_t t[MM_HTAPS_WORDS];
  mm_build_htaps(t, kLuma);
  long bad = 0, n = 0;
  srand(7);
  for (int it = 0; it < 2000; ++it)
  {
    int16_t s[16];
    for (int i = 0; i < 16; ++i) s[i] = (int16_t)((it & 1) ? (rand() % 1024) : (rand() % 65536 - 32768));
    uint32_t v[8];
    for (int k = 0; k < 8; ++k) v[k] = (uint16_t)s[2 * k] | ((uint32_t)(uint16_t)s[2 * k + 1] << 16);
    for (int o = 0; o < 4; ++o)
      for (int ph = 0; ph < 16; ++ph)
        for (int j = 0; j < 4; ++j)
        {
          const uint32_t* c = t + ((o * 16 + ph) * 4 + j) * 3;
          int acc = 123;
          for (int k = 0; k < 6; ++k) acc = dp2a(v[(j >> 1) + k], c[k >> 1], k & 1, acc);
          int ref = 123;
          for (int tap = 0; tap < 8; ++tap) ref += kLuma[ph][tap] * s[o + j + tap];
          bad += acc != ref;
          ++n;
        }
  }
  for (int ph = 0; ph < 16; ++ph)   // vertical-pass words inside entry o = 0
  {
    const uint32_t* e = t + ph * 12;
    const int8_t* c = kLuma[ph];
    auto pk = [](int a, int b, int cc, int d) { return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(cc & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24); };
    bad += e[0] != pk(c[0], c[1], c[2], c[3]);
    bad += e[1] != pk(c[4], c[5], c[6], c[7]);
    bad += e[2] != 0;
    bad += e[3] != pk(0, c[0], c[1], c[2]);
    bad += e[4] != pk(c[3], c[4], c[5], c[6]);
    bad += e[5] != pk(c[7], 0, 0, 0);
  }
  for (int o = 0; o < 4; ++o)      // the shifted base words k_fir derives at run time == the table words
    for (int ph = 0; ph < 16; ++ph)
      for (int j = 0; j < 4; ++j)
      {
        uint32_t w[3];
        mm_htaps_shift(t[ph * 12 + 0], t[ph * 12 + 1], o + (j & 1), w);
        const uint32_t* c = t + ((o * 16 + ph) * 4 + j) * 3;
        bad += (w[0] != c[0]) + (w[1] != c[1]) + (w[2] != c[2]);
        ++n;
      }
  printf("htaps: %ld cases, %ld mismatches\n", n, bad);
  return bad ? 1 : 0;
}
