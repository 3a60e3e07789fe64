// mm_sbpos.cuh -- position of one sub-block: picture-grid lookup (MPA cache, ERP trig tables) + motion model + fixed point.
#pragma once
#include "mm_common.cuh"

// Candidate-independent state of one sub-block: its coordinates, packet flag and the model's prepared values.
struct MmSbState
{
  float X, Y;
  bool pk;
  MmInv inv;
};

__device__ __forceinline__ void mm_subblock_prepare(const MmGeom& g, const float2* const* mpa, const uint8_t* const* mpaVip,
                                                    const float2* const* thTab, const float2* const* phTab, int model,
                                                    int comp, int csx, int csy, int posX, int posY, int sizeW, int sizeH,
                                                    const MmCand& blk, const float* R, int row, int col, MmSbState* st)
{
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = sizeH / sbh, cols = sizeW / sbw;
  const int n = rows * cols;
  const int i = col * rows + row;  // Eigen column-major linear index
  const bool pk = i < (n & ~3);
  const bool isLuma = comp == 0;
  float X, Y;
  int gc = -1, gr = -1;  // picture grid cell (4x4 luma units) of the sub-block, when it sits on the grid
  if (isLuma)
  {
    gc = (posX >> 2) + col;
    gr = (posY >> 2) + row;
    X = MM_FADD((float)(4 * gc), g.off);
    Y = MM_FADD((float)(4 * gr), g.off);
  }
  else
  {
    // LinSpaced(n, start, end) with step exactly 4 (MVReprojection.cpp:113-120): all values exact
    X = MM_FADD(MM_FADD(MM_FMUL((float)posX, (float)(1 << csx)), g.off), (float)(4 * col));
    Y = MM_FADD(MM_FADD(MM_FMUL((float)posY, (float)(1 << csy)), g.off), (float)(4 * row));
    const int lx = posX << csx, ly = posY << csy;
    if (((lx | ly) & 3) == 0) { gc = (lx >> 2) + col; gr = (ly >> 2) + row; }
  }
  float cPX = 0.f, cPY = 0.f;
  bool cVip = false;
  const bool mpaModel = model >= MM_MPA_FRONT_BACK && model <= MM_MPA_TOP_BOTTOM;
  MmTrig tr;
  tr.valid = false;
  tr.st = tr.ct = tr.sp = tr.cp = 0.f;
  if (isLuma && mpaModel)
  {
    const int ci = gr * g.cols4 + gc;
    const float2 p = __ldg(&mpa[model - 1][ci]);
    cPX = p.x;
    cPY = p.y;
    cVip = __ldg(&mpaVip[model - 1][ci]) != 0;
  }
  else if (gc >= 0 && gc < g.cols4 && gr < g.rows4)
  {
    const float2 t = __ldg(&thTab[pk ? 1 : 0][gr]), q = __ldg(&phTab[pk ? 1 : 0][gc]);
    tr.st = t.x; tr.ct = t.y; tr.sp = q.x; tr.cp = q.y;
    tr.valid = true;
  }
  st->X = X;
  st->Y = Y;
  st->pk = pk;
  mm_model_prepare(g, model, isLuma, blk, R, X, Y, pk, cPX, cPY, cVip, tr, &st->inv);
}

__device__ __forceinline__ void mm_subblock_apply(const MmGeom& g, int model, int comp, int csx, int csy, const MmCand& c,
                                                  const float* R, const MmSbState& st, int* fx, int* fy)
{
  const bool isLuma = comp == 0;
  float MX, MY;
  mm_model_apply(g, model, isLuma, c, R, st.X, st.Y, st.pk, st.inv, &MX, &MY);
  mm_to_fixed(g, st.X, st.Y, MX, MY, isLuma ? 0 : csx, isLuma ? 0 : csy, fx, fy);
}

// Position of one sub-block in 1/16 (luma) or 1/32 (chroma 4:2:0) pel: mapping + fixed-point conversion.
__device__ __forceinline__ void mm_subblock_pos(const MmGeom& g, const float2* const* mpa, const uint8_t* const* mpaVip,
                                                const float2* const* thTab, const float2* const* phTab, int model,
                                                int comp, int csx, int csy, int posX, int posY, int sizeW, int sizeH,
                                                const MmCand& c, const float* R, int row, int col, int* fx, int* fy)
{
  MmSbState st;
  mm_subblock_prepare(g, mpa, mpaVip, thTab, phTab, model, comp, csx, csy, posX, posY, sizeW, sizeH, c, R, row, col, &st);
  mm_subblock_apply(g, model, comp, csx, csy, c, R, st, fx, fy);
}
