// adapter_check.cpp -- compiles the C++ host adapter against libmmgpu.so and exercises it.
// Without a GPU: init() must throw (no CPU fallback) -> prints "nogpu-throws".  With a GPU: runs MC + a 9-point
// refinement and prints checksums that the Python test compares with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../vvc-extension-mm_b200/host/mm_gpu_adapter.h"
using namespace mmgpu_host;

int main(int argc, char** argv)
{
  const int W = 512, H = 256, M = 288;
  GpuInterPrediction ip;
  try { ip.init(W, H, 1, 10, 128, M, 1.0f, 1, 2, 0); }
  catch (const std::exception& e) { printf("nogpu-throws: %s\n", e.what()); return argc > 1 ? 1 : 0; }
  // deterministic synthetic planes with replicated margins
  std::vector<int16_t> planes[3];
  CPelBuf reco[3];
  for (int c = 0; c < 3; ++c)
  {
    const int w = c ? W / 2 : W, h = c ? H / 2 : H, m = c ? M / 2 : M, st = w + 2 * m;
    planes[c].assign((size_t)st * (h + 2 * m), 0);
    for (int y = -m; y < h + m; ++y)
      for (int x = -m; x < w + m; ++x)
      {
        const int xx = x < 0 ? 0 : (x >= w ? w - 1 : x), yy = y < 0 ? 0 : (y >= h ? h - 1 : y);
        planes[c][(size_t)(y + m) * st + (x + m)] = (int16_t)((xx * 7 + yy * 13 + (xx * yy) % 31 + c * 101) & 1023);
      }
    reco[c].buf = planes[c].data() + (size_t)m * st + m; reco[c].stride = st; reco[c].width = w; reco[c].height = h;
  }
  ip.uploadReference(0, 0, reco);
  const int epi[3] = {1 << 24, 0, 0};
  ip.setEpipoleFixed(0, epi);
  std::vector<int16_t> dst(64 * 64);
  PelBuf d = {dst.data(), 64, 64, 64};
  ClpRng clp = {0, 1023, 10};
  ip.xPredInterBlkMM(0, Position{128, 64}, Size{64, 64}, 0, Mv{37, -21}, d, MMGPU_GEODESIC_CAMPOSE, false, clp);
  unsigned long long sum = 0;
  for (size_t i = 0; i < dst.size(); ++i) sum = sum * 1000003ULL + (unsigned short)dst[i];
  printf("mc %llu\n", sum);
  std::vector<int16_t> org(32 * 32);
  for (int i = 0; i < 32 * 32; ++i) org[i] = (int16_t)((i * 37) & 1023);
  CPelBuf key = {org.data(), 32, 32, 32};
  const Mv refine[9] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
  Mv best;
  Distortion c = ip.xPatternRefinementProjected(Position{64, 96}, Size{32, 32}, 0, MMGPU_MPA_LEFT_RIGHT, key, MV_PRECISION_HALF,
                                                Mv{6, -4}, refine, true, clp, [](const Mv& m) { return (Distortion)(abs(m.hor) + abs(m.ver)); }, best);
  printf("refine %llu %d %d\n", (unsigned long long)c, best.hor, best.ver);
  return 0;
}
