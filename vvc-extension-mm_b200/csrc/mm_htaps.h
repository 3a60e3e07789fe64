/* mm_htaps.h -- host-side builder of k_fir's alignment-folded luma coefficient table (plain C++, no CUDA: also compiled
 * by g++ for tests/native/htaps_check.cpp).  See mm_fir.cuh, "Wide-load variant". */
#pragma once
#include <stdint.h>

#define MM_HTAPS_WORDS (4 * 16 * 12)
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

