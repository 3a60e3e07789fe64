// map_host.cpp -- host build of the PRODUCT mapping code (vvc-extension-mm_b200/csrc/mm_map.cuh,
// mm_libm.cuh) behind a tiny C interface, so the CPU test-suite can compare it with the oracle
// without a GPU.  The sub-block enumeration mirrors mm_subblock_pos() in mm_sbpos.cuh.
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../vvc-extension-mm_b200/csrc/mm_map.cuh"
#include "../../vvc-extension-mm_b200/csrc/mm_mvp.cuh"

struct HostCtx
{
  MmGeom g;
  std::vector<float> px[3], py[3];
  std::vector<uint8_t> vip[3];
};

extern "C" void* mmhost_create(int picW, int picH, float off, int gedModulated, int fmaBuild)
{
  HostCtx* c = new HostCtx();
  MmGeom& g = c->g;
  g.picW = (float)picW; g.picH = (float)picH;
  g.focal = (float)(1. / tan(M_PI / picH));
  g.angRes = (float)(M_PI / picH);
  g.off = off; g.rows4 = picH / 4; g.cols4 = picW / 4; g.gedModulated = gedModulated; g.fmaBuild = fmaBuild;
  const long n = (long)g.rows4 * g.cols4;
  for (int m = 0; m < 3; ++m)
  {
    c->px[m].resize(n); c->py[m].resize(n); c->vip[m].resize(n);
    for (int row = 0; row < g.rows4; ++row)
      for (int col = 0; col < g.cols4; ++col)
      {
        const bool pk = ((long)col * g.rows4 + row) < (n & ~3L);
        float PX, PY; bool v;
        MmTrig none; none.valid = false; none.st = none.ct = none.sp = none.cp = 0.f;
        mm_mpa_to_perspective(g, m + 1, (float)(4 * col) + off, (float)(4 * row) + off, pk, none, &PX, &PY, &v);
        const long i = (long)row * g.cols4 + col;
        c->px[m][i] = PX; c->py[m][i] = PY; c->vip[m][i] = v;
      }
  }
  return c;
}
extern "C" void mmhost_destroy(void* h) { delete (HostCtx*)h; }

extern "C" void mmhost_reproject(void* h, int comp, int csx, int csy, int x, int y, int w, int hgt, int mvHor, int mvVer,
                                 int model, const float* R, int32_t* outX, int32_t* outY)
{
  HostCtx* c = (HostCtx*)h;
  const MmGeom& g = c->g;
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = hgt / sbh, cols = w / sbw, n = rows * cols;
  const bool isLuma = comp == 0;
  MmCand cand;
  mm_cand_setup(g, model, isLuma, x, y, w, hgt, mvHor, mvVer, R, &cand);
  for (int col = 0; col < cols; ++col)
    for (int row = 0; row < rows; ++row)
    {
      const int i = col * rows + row;
      const bool pk = i < (n & ~3);
      float X, Y;
      if (isLuma) { X = (float)(4 * ((x >> 2) + col)) + g.off; Y = (float)(4 * ((y >> 2) + row)) + g.off; }
      else { X = ((float)x * (float)(1 << csx) + g.off) + (float)(4 * col); Y = ((float)y * (float)(1 << csy) + g.off) + (float)(4 * row); }
      float cPX = 0, cPY = 0; bool cV = false;
      if (isLuma && model >= 1 && model <= 3)
      {
        const long ci = (long)((y >> 2) + row) * g.cols4 + ((x >> 2) + col);
        cPX = c->px[model - 1][ci]; cPY = c->py[model - 1][ci]; cV = c->vip[model - 1][ci] != 0;
      }
      // ERP toSphere table values of the grid cell (same evaluation the device tables hoist); every other
      // call goes through the direct path so both are exercised
      MmTrig tr; tr.valid = false; tr.st = tr.ct = tr.sp = tr.cp = 0.f;
      if (((row + col) & 1) == 0)
      {
        float th, ph;
        mm_erp_angles(g, X, Y, &th, &ph);
        mm_sincos_pk(th, pk, g.fmaBuild != 0, &tr.st, &tr.ct);
        mm_sincos_pk(ph, pk, g.fmaBuild != 0, &tr.sp, &tr.cp);
        tr.valid = true;
      }
      float MX, MY;
      mm_model_motion(g, model, isLuma, cand, R, X, Y, pk, cPX, cPY, cV, tr, &MX, &MY);
      int fx, fy;
      mm_to_fixed(g, X, Y, MX, MY, isLuma ? 0 : csx, isLuma ? 0 : csy, &fx, &fy);
      outX[i] = fx; outY[i] = fy;
    }
}

// MM-MVP conversion (mm_mvp.cuh) on the host
extern "C" void mmhost_mvp(void* h, int posX, int posY, int mvHor, int mvVer, int modelOrig, int modelDes, int shiftHor, int shiftVer,
                           const float* Rorig, const float* Rdes, int sameRot, int candX, int candY, int candW, int candH, int curX,
                           int curY, int curW, int curH, int32_t* out)
{
  HostCtx* c = (HostCtx*)h;
  mm_mvp_convert(c->g, posX, posY, mvHor, mvVer, modelOrig, modelDes, shiftHor, shiftVer, Rorig, Rdes, sameRot != 0, candX, candY, candW,
                 candH, curX, curY, curW, curH, &out[0], &out[1]);
}
