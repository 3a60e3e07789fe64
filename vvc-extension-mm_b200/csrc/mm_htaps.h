/* mm_htaps.h -- host-side builder of k_fir's alignment-folded luma coefficient table (plain C++, no CUDA: also compiled
 * by g++ for tests/native/htaps_check.cpp).  See mm_fir.cuh, "Wide-load variant". */
#pragma once
#include <stdint.h>

#define MM_HTAPS_WORDS (4 * 16 * 12)
#define MM_HTAPS_BASE_WORDS 32   /* compact table behind it: (C0, C1) of the 16 phases, one 128-byte line */
static inline void mm_build_htaps(uint32_t* t, const int8_t (*lf)[8])
{
  for (int o = 0; o < 4; ++o)
    for (int ph = 0; ph < 16; ++ph)
      for (int j = 0; j < 4; ++j)
        for (int m = 0; m < 3; ++m)
        {
          uint32_t w = 0;
          for (int b = 0; b < 4; ++b)
          {
            const int k = 2 * m + (b >> 1);                 // pair index inside the output's 6 pairs
            const int sample = 2 * ((j >> 1) + k) + (b & 1);  // sample index inside the 16-sample row
            const int tap = sample - (o + j);
            const int c = (tap >= 0 && tap < 8) ? lf[ph][tap] : 0;
            w |= (uint32_t)(c & 0xff) << (8 * b);
          }
          t[((o * 16 + ph) * 4 + j) * 3 + m] = w;
        }
}


/* The words of alignment o and output j are the phase's two base words C0 = (c0,c1,c2,c3), C1 = (c4,c5,c6,c7) -- a 12-byte
 * string c0..c7,0,0,0,0 -- moved up by sh = o + (j & 1) bytes (0..4): k_fir keeps only C0, C1 per phase (128 bytes of
 * shared memory) and derives the six words it needs (sh = o for the even outputs, o + 1 for the odd ones) with three
 * funnel shifts each.  Host model of __funnelshift_lc(lo, hi, s) = high word of (hi:lo) << min(s, 32): */
static inline uint32_t mm_fshl_clamp(uint32_t lo, uint32_t hi, unsigned s)
{
  if (s > 32) s = 32;
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return s == 32 ? lo : (uint32_t)((v << s) >> 32);
}
static inline void mm_htaps_shift(uint32_t c0, uint32_t c1, int sh, uint32_t w[3])
{
  w[0] = mm_fshl_clamp(0u, c0, 8u * (unsigned)sh);
  w[1] = mm_fshl_clamp(c0, c1, 8u * (unsigned)sh);
  w[2] = mm_fshl_clamp(c1, 0u, 8u * (unsigned)sh);
}
