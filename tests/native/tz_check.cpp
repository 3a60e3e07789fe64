// tz_check.cpp -- prints the point lists of the PRODUCT's search enumerators (host/mm_gpu_adapter.h) for the cases given on
// stdin, one per line:  D left right top bottom startX startY dist corners | P left right top bottom bestX bestY pointNr |
// R left right top bottom step.  Output: "x y nr dist" quadruples per line.  tests/test_native_cpu.py compares them with a
// Python restatement of InterSearch.cpp:456-481, 535-737, 5611-5647.
#include <cstdio>
#include <vector>
#include "../../vvc-extension-mm_b200/host/mm_gpu_adapter.h"
int main()
{
  char kind;
  while (scanf(" %c", &kind) == 1)
  {
    mmgpu_host::SearchRange sr;
    std::vector<mmgpu_host::TzPoint> pts;
    if (scanf("%d %d %d %d", &sr.left, &sr.right, &sr.top, &sr.bottom) != 4) return 1;
    if (kind == 'D') { int x, y, d, c; if (scanf("%d %d %d %d", &x, &y, &d, &c) != 4) return 1; mmgpu_host::tzDiamondRound(sr, x, y, d, c != 0, pts); }
    else if (kind == 'P') { int x, y, nr; if (scanf("%d %d %d", &x, &y, &nr) != 3) return 1; mmgpu_host::tz2Point(sr, x, y, (uint8_t)nr, pts); }
    else { int step; if (scanf("%d", &step) != 1) return 1; mmgpu_host::tzRaster(sr, step, pts); }
    for (auto& p : pts) printf("%d %d %d %u ", p.x, p.y, (int)p.pointNr, p.distance);
    printf("\n");
  }
  return 0;
}
