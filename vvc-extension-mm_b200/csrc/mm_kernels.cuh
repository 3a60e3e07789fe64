// mm_kernels.cuh -- sm_100a kernels of the multi-model MC + distortion path.
//
//   k_mpa_cache        MotionPlaneAdaptiveMotionModel::fillCache   (MotionPlaneAdaptiveMotionModel.cpp:97-104)
//   k_cand_setup       per-(block,mv) scalars of the motion models (mm_map.cuh: mm_cand_setup)
//   k_reproject        MVReprojection::reprojectMotionVectorSubblocks (MVReprojection.cpp:80-166)
//   k_mc_generic       xPredInterBlkMM / xPredInterBlk, any component, uni/bi (InterPrediction.cpp:683-856, 928-1096)
//   k_mc_jobs          whole-picture reconstruction shape: jobs of any component, uni or bi with fused addAvg
//   k_me<MODE>         fused map -> 8-tap separable FIR -> SAD/HAD for all candidates of all blocks
//                      (InterSearch.cpp:6277-6385 + RdCost.cpp:482-.., 2943-3067), or prediction write-out
//   k_extend_border    Picture::extendPicBorder (Picture.cpp:988-1048)
//   k_add_avg          PelBuf::addAvg (Buffer.cpp:551-587)
//
// HBM layout: reference planes int16, pitch a multiple of 64 samples, margins resident; positions never
// leave registers in k_me; one warp evaluates 32 sub-blocks (4x4 luma) of one block: either 32 consecutive
// sub-blocks of one candidate (blocks with >= 32 sub-blocks) or 32/n candidates of a small block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mmgpu.h"
#include "mm_map.cuh"

struct MmRef
{
  const int16_t* plane00[3];
  int32_t stride[3];
  int32_t poc;
};

struct MmItem  // one warp of work
{
  int32_t block;  // index into blocks
  int32_t cand;   // first candidate (global index) handled by this warp
  int32_t sb;     // first sub-block (tile-major index) handled by this warp
  int32_t ncand;  // number of consecutive candidates this warp evaluates
};

struct MmFir  // InterpolationFilter::filter<> shifts/offsets for the context's bit depth
{
  int32_t shift1, off1;    // first pass, not last:  (sum + off1) >> shift1
  int32_t shift2, off2;    // second pass, last:     (sum + off2) >> shift2, clip
  int32_t shift2bi;        // second pass, not last: sum >> 6
  int32_t shift1l, off1l;  // single pass first&last: (sum + 32) >> 6, clip
  int32_t cpShift;         // filterCopy first-not-last: (p << cpShift) - 8192
};

struct MmParams
{
  MmGeom g;
  MmFir fir;
  const mmgpu_block* blocks;
  const MmItem* items;
  int32_t nItems;
  const mmgpu_mv* mvs;
  int32_t mvShift;
  const MmCand* cands;
  const MmRef* refs;
  const float* rot;          // [slot][9]; slots 0..2 = GEODESIC_X/Y/Z, 3.. = user slots
  const float2* mpa[3];      // row-major [rows4][cols4] motion-plane coordinates of models 1..3
  const uint8_t* mpaVip[3];
  const float2* thTab[2];    // [scalar|packet][rows4]: (sin, cos) of theta(4*row + off)   (ERP toSphere hoisted)
  const float2* phTab[2];    // [scalar|packet][cols4]: (sin, cos) of phi(4*col + off)
  const int16_t* org;
  int32_t orgStride;
  int32_t distKind, subShift, hadMode;
  int32_t clpMin, clpMax;
  int32_t picW, picH, maxCU;
  int32_t oobMode;
  unsigned long long* dist;
  int16_t* dst;
  int32_t dstStride;
};

#define MM_ROT_USER0 3

__device__ __forceinline__ const float* mm_rot_ptr(const float* rot, int model, int rotSlot)
{
  const int s = (model == MM_GEODESIC_CAMPOSE) ? (MM_ROT_USER0 + rotSlot) : (model - MM_GEODESIC_X);
  return rot + 9 * (s < 0 ? 0 : s);
}

// -------------------------------------------------------------------------------------------------
// Filter tables (InterpolationFilter.cpp:82-100, 187-221)
// -------------------------------------------------------------------------------------------------
__constant__ int8_t c_lumaFilter[16][8] = {
  {0, 0, 0, 64, 0, 0, 0, 0},       {0, 1, -3, 63, 4, -2, 1, 0},      {-1, 2, -5, 62, 8, -3, 1, 0},
  {-1, 3, -8, 60, 13, -4, 1, 0},   {-1, 4, -10, 58, 17, -5, 1, 0},   {-1, 4, -11, 52, 26, -8, 3, -1},
  {-1, 3, -9, 47, 31, -10, 4, -1}, {-1, 4, -11, 45, 34, -10, 4, -1}, {-1, 4, -11, 40, 40, -11, 4, -1},
  {-1, 4, -10, 34, 45, -11, 4, -1}, {-1, 4, -10, 31, 47, -9, 3, -1}, {-1, 3, -8, 26, 52, -11, 4, -1},
  {0, 1, -5, 17, 58, -10, 4, -1},  {0, 1, -4, 13, 60, -8, 3, -1},    {0, 1, -3, 8, 62, -5, 2, -1},
  {0, 1, -2, 4, 63, -3, 1, 0}};
__constant__ int8_t c_chromaFilter[32][4] = {
  {0, 64, 0, 0},    {-1, 63, 2, 0},   {-2, 62, 4, 0},   {-2, 60, 7, -1},  {-2, 58, 10, -2}, {-3, 57, 12, -2},
  {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
  {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
  {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
  {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
  {0, 4, 62, -2},   {0, 2, 63, -1}};

// -------------------------------------------------------------------------------------------------
// Geometry helpers shared by the kernels
// -------------------------------------------------------------------------------------------------
struct MmSub  // where a sub-block sits inside its block
{
  int row, col;   // sub-block row/col inside the block
  int rows, cols; // sub-block grid of the block
  int T;          // sub-blocks per Hadamard tile (1,2,4,8), q = index inside the tile
  int q;
  int tileKind;   // 0: 4x4, 1: 8x8, 2: 16x8 or 8x16, 3: 8x4 or 4x8
};

// Tile-major enumeration: sub-block s -> (row, col).  Tiles follow RdCost::xGetHADs' tiling rule
// (RdCost.cpp:2962-3060) so that the sub-blocks of one Hadamard tile sit in adjacent lanes.
__device__ __forceinline__ MmSub mm_sub_from_index(int w, int h, int s)
{
  MmSub r;
  r.cols = w >> 2;
  r.rows = h >> 2;
  int ltw, lth;  // log2 of the tile size in sub-blocks (block sizes are powers of two)
  if (w > h && (h & 7) == 0 && (w & 15) == 0) { ltw = 2; lth = 1; r.tileKind = 2; }
  else if (w < h && (w & 7) == 0 && (h & 15) == 0) { ltw = 1; lth = 2; r.tileKind = 2; }
  else if (w > h && (h & 3) == 0 && (w & 7) == 0) { ltw = 1; lth = 0; r.tileKind = 3; }
  else if (w < h && (w & 3) == 0 && (h & 7) == 0) { ltw = 0; lth = 1; r.tileKind = 3; }
  else if ((h & 7) == 0 && (w & 7) == 0) { ltw = 1; lth = 1; r.tileKind = 1; }
  else { ltw = 0; lth = 0; r.tileKind = 0; }
  r.T = 1 << (ltw + lth);
  const int t = s >> (ltw + lth);
  r.q = s & (r.T - 1);
  const int ltpr = (31 - __clz(r.cols)) - ltw;  // log2(tiles per row)
  const int ty = t >> ltpr, tx = t & ((1 << ltpr) - 1);
  const int qy = r.q >> ltw, qx = r.q & ((1 << ltw) - 1);
  r.col = (tx << ltw) + qx;
  r.row = (ty << lth) + qy;
  return r;
}

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

// -------------------------------------------------------------------------------------------------
// k_mpa_cache: picture-wide motion-plane cache (always the vectorised Eigen path: n = W/4*H/4)
// -------------------------------------------------------------------------------------------------
__global__ void k_mpa_cache(MmGeom g, int model, float2* out, uint8_t* vip)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.rows4 * g.cols4) return;
  const int row = idx / g.cols4, col = idx - row * g.cols4;
  const float X = MM_FADD((float)(4 * col), g.off), Y = MM_FADD((float)(4 * row), g.off);
  const long n = (long)g.rows4 * g.cols4;
  const bool pk = ((long)col * g.rows4 + row) < (n & ~3L);
  float PX, PY;
  bool v;
  MmTrig none; none.valid = false; none.st = none.ct = none.sp = none.cp = 0.f;
  mm_mpa_to_perspective(g, model, X, Y, pk, none, &PX, &PY, &v);
  out[idx] = make_float2(PX, PY);
  vip[idx] = v ? 1 : 0;
}

// ERP toSphere angles of the picture grid, scalar-libm and Eigen-packet variants (Projection.cpp:213-221,
// Coordinate.cpp:53-60): tab[0] = scalar, tab[1] = packet; th over rows, ph over columns.
__global__ void k_trig_tables(MmGeom g, float2* th0, float2* th1, float2* ph0, float2* ph1)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool fma = g.fmaBuild != 0;
  if (i < g.rows4)
  {
    float th, ph, s, c;
    mm_erp_angles(g, 0.0f, MM_FADD((float)(4 * i), g.off), &th, &ph);
    mm_sincos_pk(th, false, fma, &s, &c); th0[i] = make_float2(s, c);
    mm_sincos_pk(th, true, fma, &s, &c);  th1[i] = make_float2(s, c);
  }
  if (i < g.cols4)
  {
    float th, ph, s, c;
    mm_erp_angles(g, MM_FADD((float)(4 * i), g.off), 0.0f, &th, &ph);
    mm_sincos_pk(ph, false, fma, &s, &c); ph0[i] = make_float2(s, c);
    mm_sincos_pk(ph, true, fma, &s, &c);  ph1[i] = make_float2(s, c);
  }
}

// -------------------------------------------------------------------------------------------------
// k_build_items: the work list of a batch, one warp per block.  itemBase[b] = first work item of block b
// (host-computed, grouped by motion model); also fills the candidate -> block map used by k_cand_setup.
// -------------------------------------------------------------------------------------------------
__global__ void k_build_items(const mmgpu_block* __restrict__ blocks, const int32_t* __restrict__ itemBase, int nBlocks,
                              int candChunk, MmItem* __restrict__ items, int32_t* __restrict__ candBlock)
{
  const int bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bi >= nBlocks) return;
  const mmgpu_block b = blocks[bi];
  const int n = (b.w >> 2) * (b.h >> 2);
  const int base = itemBase[bi];
  for (int j = lane; j < b.cand_count; j += 32) candBlock[b.cand_first + j] = bi;
  // a work item = up to candChunk candidates (n >= 32) or candChunk steps of 32/n candidates (n < 32) of one
  // 32-sub-block chunk: the warp prepares the candidate-independent state once and loops over the candidates
  if (n >= 32)
  {
    const int chunks = n >> 5, groups = (b.cand_count + candChunk - 1) / candChunk, total = groups * chunks;
    for (int t = lane; t < total; t += 32)
    {
      const int gidx = t / chunks, ch = t - gidx * chunks;
      const int c0 = gidx * candChunk;
      items[base + t] = MmItem{bi, b.cand_first + c0, ch << 5, min(candChunk, b.cand_count - c0)};
    }
  }
  else
  {
    const int per = (32 / n) * candChunk, total = (b.cand_count + per - 1) / per;
    for (int t = lane; t < total; t += 32)
      items[base + t] = MmItem{bi, b.cand_first + t * per, 0, min(per, b.cand_count - t * per)};
  }
}

// -------------------------------------------------------------------------------------------------
// k_cand_setup: one thread per candidate MV
// -------------------------------------------------------------------------------------------------
__global__ void k_cand_setup(MmGeom g, const mmgpu_block* __restrict__ blocks, const int32_t* __restrict__ candBlock,
                             const mmgpu_mv* __restrict__ mvs, int mvShift, const float* __restrict__ rot, int nCands,
                             MmCand* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nCands) return;
  const mmgpu_block b = blocks[candBlock[i]];
  const mmgpu_mv mv = mvs[i];
  MmCand c;
  // Mv::changePrecision(mvPrec, MV_PRECISION_INTERNAL) is a left shift (Mv.h)
  const int mh = mv.hor * (1 << mvShift), mvv = mv.ver * (1 << mvShift);
  if (b.model == MM_CLASSIC)
  {
    c.mvX = 0.f; c.mvY = 0.f; c.mvZero = 0;
    for (int k = 0; k < 9; ++k) c.p[k] = 0.f;
    c.p[0] = __int_as_float(mh);
    c.p[1] = __int_as_float(mvv);
  }
  else
  {
    mm_cand_setup(g, b.model, true, b.x, b.y, b.w, b.h, mh, mvv, mm_rot_ptr(rot, b.model, b.rot_slot), &c);
  }
  out[i] = c;
}

// -------------------------------------------------------------------------------------------------
// k_reproject: positions of one block (parity entry point a1)
// -------------------------------------------------------------------------------------------------
__global__ void k_reproject(MmParams P, int comp, int csx, int csy, int x, int y, int w, int h, mmgpu_mv mv, int model,
                            int rotSlot, int32_t* outX, int32_t* outY)
{
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = h / sbh, cols = w / sbw;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int col = i / rows, row = i - col * rows;
  const float* R = mm_rot_ptr(P.rot, model, rotSlot);
  MmCand c;
  mm_cand_setup(P.g, model, comp == 0, x, y, w, h, mv.hor, mv.ver, R, &c);
  int fx, fy;
  mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, comp, csx, csy, x, y, w, h, c, R, row, col, &fx, &fy);
  outX[i] = fx;
  outY[i] = fy;
}

// -------------------------------------------------------------------------------------------------
// Generic FIR of one sub-block (any component / tap count / uni or bi): the literal two-pass
// structure of xPredInterBlkMM (InterPrediction.cpp:786-828) with filter<>'s arithmetic.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mm_clip(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <int NTAPS>
__device__ void mm_fir_generic(const MmFir& f, const int16_t* __restrict__ ref, int stride, int xPos, int yPos, int xFrac,
                               int yFrac, int sbw, int sbh, bool rnd, int lo, int hi, int16_t* dst, int dstStride)
{
  const int8_t* ch = (NTAPS == 8) ? c_lumaFilter[xFrac] : c_chromaFilter[xFrac];
  const int8_t* cv = (NTAPS == 8) ? c_lumaFilter[yFrac] : c_chromaFilter[yFrac];
  const int16_t* p = ref + (long)yPos * stride + xPos;
  const int half = NTAPS / 2 - 1;
  if (yFrac == 0 || xFrac == 0)
  {
    for (int r = 0; r < sbh; ++r)
      for (int x = 0; x < sbw; ++x)
      {
        int val;
        if (xFrac == 0 && yFrac == 0)
        {
          // filterCopy<true,isLast> (:392-470)
          const int s = p[r * stride + x];
          val = rnd ? s : (int)(int16_t)((int16_t)(s << f.cpShift) - (int16_t)8192);
        }
        else
        {
          int sum = 0;
          if (yFrac == 0)
            for (int k = 0; k < NTAPS; ++k) sum += p[r * stride + x + k - half] * ch[k];
          else
            for (int k = 0; k < NTAPS; ++k) sum += p[(r + k - half) * stride + x] * cv[k];
          if (rnd) val = mm_clip((int)(int16_t)((sum + f.off1l) >> f.shift1l), lo, hi);
          else val = (int)(int16_t)((sum + f.off1) >> f.shift1);
        }
        dst[r * dstStride + x] = (int16_t)val;
      }
    return;
  }
  int16_t tmp[(4 + 7) * 4];
  for (int r = 0; r < sbh + NTAPS - 1; ++r)
    for (int x = 0; x < sbw; ++x)
    {
      int sum = 0;
      for (int k = 0; k < NTAPS; ++k) sum += p[(r - half) * stride + x + k - half] * ch[k];
      tmp[r * 4 + x] = (int16_t)((sum + f.off1) >> f.shift1);
    }
  for (int r = 0; r < sbh; ++r)
    for (int x = 0; x < sbw; ++x)
    {
      int sum = 0;
      for (int k = 0; k < NTAPS; ++k) sum += tmp[(r + k) * 4 + x] * cv[k];
      int val;
      if (rnd) val = mm_clip((int)(int16_t)((sum + f.off2) >> f.shift2), lo, hi);
      else val = (int)(int16_t)(sum >> f.shift2bi);
      dst[r * dstStride + x] = (int16_t)val;
    }
}

// One thread per sub-block of ONE block (mmgpu_mc).  model == CLASSIC: one MV for the block (xPredInterBlk).
__global__ void k_mc_generic(MmParams P, int refSlot, int comp, int csx, int csy, int x, int y, int w, int h, mmgpu_mv mv,
                             int model, int rotSlot, int bi, int16_t* dst, int dstStride)
{
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = h / sbh, cols = w / sbw;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int col = i / rows, row = i - col * rows;
  const MmRef ref = P.refs[refSlot];
  const int shH = 4 + csx, shV = 4 + csy;
  int fx, fy;
  int16_t* d = dst + (long)row * sbh * dstStride + col * sbw;
  if (model == MM_CLASSIC)
  {
    // xPredInterBlk (InterPrediction.cpp:928-1096, no wrap / RPR): every sample uses the block MV
    fx = ((x + col * sbw) << shH) + mv.hor;
    fy = ((y + row * sbh) << shV) + mv.ver;
  }
  else
  {
    const float* R = mm_rot_ptr(P.rot, model, rotSlot);
    MmCand c;
    mm_cand_setup(P.g, model, comp == 0, x, y, w, h, mv.hor, mv.ver, R, &c);
    mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, comp, csx, csy, x, y, w, h, c, R, row, col, &fx, &fy);
    const int xP = fx >> shH, yP = fy >> shV;
    const int picw = P.picW >> csx, pich = P.picH >> csy;
    const int mcw = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csx);
    const int mch = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csy);
    if (xP < -mcw || yP < -mch || xP >= picw + mcw - sbw || yP >= pich + mch - sbh)
    {
      for (int r = 0; r < sbh; ++r)
        for (int q = 0; q < sbw; ++q) d[r * dstStride + q] = 0;
      return;
    }
  }
  const int xP = fx >> shH, yP = fy >> shV, xF = fx & ((1 << shH) - 1), yF = fy & ((1 << shV) - 1);
  if (comp == 0)
    mm_fir_generic<8>(P.fir, ref.plane00[0], ref.stride[0], xP, yP, xF, yF, sbw, sbh, !bi, P.clpMin, P.clpMax, d, dstStride);
  else
    mm_fir_generic<4>(P.fir, ref.plane00[comp], ref.stride[comp], xP, yP, xF, yF, sbw, sbh, !bi, P.clpMin, P.clpMax, d,
                      dstStride);
}

// One sub-block of one component into `out` (row stride 4): mapping (or classic MV) + OOB rule + FIR.
__device__ __forceinline__ void mm_mc_subblock(const MmParams& P, int refSlot, int comp, int csx, int csy, int x, int y, int w,
                                               int h, mmgpu_mv mv, int model, int rotSlot, bool bi, int row, int col,
                                               int16_t* out)
{
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const MmRef ref = P.refs[refSlot];
  const int shH = 4 + csx, shV = 4 + csy;
  int fx, fy;
  if (model == MM_CLASSIC)
  {
    fx = ((x + col * sbw) << shH) + mv.hor;
    fy = ((y + row * sbh) << shV) + mv.ver;
  }
  else
  {
    const float* R = mm_rot_ptr(P.rot, model, rotSlot);
    MmCand c;
    mm_cand_setup(P.g, model, comp == 0, x, y, w, h, mv.hor, mv.ver, R, &c);
    mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, comp, csx, csy, x, y, w, h, c, R, row, col, &fx, &fy);
    const int xP = fx >> shH, yP = fy >> shV;
    const int picw = P.picW >> csx, pich = P.picH >> csy;
    const int mcw = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csx);
    const int mch = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csy);
    if (xP < -mcw || yP < -mch || xP >= picw + mcw - sbw || yP >= pich + mch - sbh)
    {
      for (int i = 0; i < 16; ++i) out[i] = 0;
      return;
    }
  }
  const int xP = fx >> shH, yP = fy >> shV, xF = fx & ((1 << shH) - 1), yF = fy & ((1 << shV) - 1);
  if (comp == 0)
    mm_fir_generic<8>(P.fir, ref.plane00[0], ref.stride[0], xP, yP, xF, yF, sbw, sbh, !bi, P.clpMin, P.clpMax, out, 4);
  else
    mm_fir_generic<4>(P.fir, ref.plane00[comp], ref.stride[comp], xP, yP, xF, yF, sbw, sbh, !bi, P.clpMin, P.clpMax, out, 4);
}

// Whole-picture reconstruction shape (DecCu::xReconInter -> motionCompensation -> xPredInterUni / xPredInterBi,
// InterPrediction.cpp:455-681): one CTA per job (one component of one PU); uni-prediction writes the clipped
// prediction, bi-prediction interpolates both lists at 14 bit and applies PelBuf::addAvg (Buffer.cpp:551-587) in
// registers before the store.  dst planes are un-padded device pictures.
__global__ void k_mc_jobs(MmParams P, const mmgpu_mc_job* __restrict__ jobs, int16_t* dstY, int16_t* dstCb, int16_t* dstCr,
                          int strideY, int strideC, int csxC, int csyC, int avgShift, int avgOffset)
{
  const mmgpu_mc_job j = jobs[blockIdx.x];
  const int csx = j.comp ? csxC : 0, csy = j.comp ? csyC : 0;
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = j.h / sbh, cols = j.w / sbw;
  int16_t* plane = j.comp == 0 ? dstY : (j.comp == 1 ? dstCb : dstCr);
  const int stride = j.comp == 0 ? strideY : strideC;
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x)
  {
    const int col = i / rows, row = i - col * rows;
    int16_t a[16], b[16];
    mm_mc_subblock(P, j.ref_slot0, j.comp, csx, csy, j.x, j.y, j.w, j.h, j.mv0, j.model, j.rot_slot0, j.bi != 0, row, col, a);
    if (j.bi)
    {
      mm_mc_subblock(P, j.ref_slot1, j.comp, csx, csy, j.x, j.y, j.w, j.h, j.mv1, j.model, j.rot_slot1, true, row, col, b);
      for (int r = 0; r < sbh; ++r)
        for (int q = 0; q < sbw; ++q)
          a[r * 4 + q] = (int16_t)mm_clip(((int)a[r * 4 + q] + (int)b[r * 4 + q] + avgOffset) >> avgShift, P.clpMin, P.clpMax);
    }
    int16_t* d = plane + (long)(j.y + row * sbh) * stride + (j.x + col * sbw);
    for (int r = 0; r < sbh; ++r)
      for (int q = 0; q < sbw; ++q) d[(long)r * stride + q] = a[r * 4 + q];
  }
}

// -------------------------------------------------------------------------------------------------
// Fast luma path: 4x4 sub-block, 8-tap separable FIR on packed 16-bit pairs with dp2a.
// All (xFrac, yFrac) cases, including 0, run the two-pass form: with the phase-0 kernel {0,0,0,64,0,0,0,0}
// it equals the reference's single-pass / copy results exactly for rndRes = true ("FIR identity");
// the copy case skips the clip like filterCopy<true,true>.
// -------------------------------------------------------------------------------------------------
struct MmTaps  // packed coefficient words of one phase, built once per CTA in shared memory
{
  int32_t a0, a1;      // (c0,c1,c2,c3) (c4,c5,c6,c7): output aligned with the sample pair
  int32_t b0, b1, b2;  // (0,c0,c1,c2) (c3,c4,c5,c6) (c7,0,0,0): output starts on the odd sample
};

__device__ __forceinline__ int mm_pack4(int c0, int c1, int c2, int c3)
{
  return (c0 & 0xff) | ((c1 & 0xff) << 8) | ((c2 & 0xff) << 16) | ((c3 & 0xff) << 24);
}
__device__ __forceinline__ void mm_build_taps(MmTaps* t /* [16] in shared memory */)
{
  for (int ph = threadIdx.x; ph < 16; ph += blockDim.x)
  {
    const int8_t* c = c_lumaFilter[ph];
    t[ph].a0 = mm_pack4(c[0], c[1], c[2], c[3]);
    t[ph].a1 = mm_pack4(c[4], c[5], c[6], c[7]);
    t[ph].b0 = mm_pack4(0, c[0], c[1], c[2]);
    t[ph].b1 = mm_pack4(c[3], c[4], c[5], c[6]);
    t[ph].b2 = mm_pack4(c[7], 0, 0, 0);
  }
}
// four outputs of an 8-tap filter over 12 packed samples w0..w5 (output j starts at sample j)
__device__ __forceinline__ void mm_fir4(const int w0, const int w1, const int w2, const int w3, const int w4,
                                        const int w5, const MmTaps& t, int init, int& o0, int& o1, int& o2, int& o3)
{
  o0 = __dp2a_lo(w0, t.a0, __dp2a_hi(w1, t.a0, __dp2a_lo(w2, t.a1, __dp2a_hi(w3, t.a1, init))));
  o1 = __dp2a_lo(w0, t.b0, __dp2a_hi(w1, t.b0, __dp2a_lo(w2, t.b1, __dp2a_hi(w3, t.b1, __dp2a_lo(w4, t.b2, init)))));
  o2 = __dp2a_lo(w1, t.a0, __dp2a_hi(w2, t.a0, __dp2a_lo(w3, t.a1, __dp2a_hi(w4, t.a1, init))));
  o3 = __dp2a_lo(w1, t.b0, __dp2a_hi(w2, t.b0, __dp2a_lo(w3, t.b1, __dp2a_hi(w4, t.b1, __dp2a_lo(w5, t.b2, init)))));
}

// pred[r*4+c]; BI = false: clipped prediction, BI = true: 14-bit intermediate
template <bool BI>
__device__ __forceinline__ void mm_fir_luma4x4(const MmFir& f, const MmTaps* taps, const int16_t* __restrict__ ref,
                                               int stride, int xPos, int yPos, int xFrac, int yFrac, int lo, int hi,
                                               int* pred)
{
  const MmTaps th = taps[xFrac], tv = taps[yFrac];
  const int a = xPos - 3;
  const int sh = (a & 1) << 4;
  const uint32_t* base = reinterpret_cast<const uint32_t*>(ref + (long)(yPos - 3) * stride + (a & ~1));
  const int wstride = stride >> 1;
  int P[6][4];  // vertical pairs of first-pass rows, per column
  int hprev[4];
#pragma unroll
  for (int r = 0; r < 11; ++r)
  {
    const uint32_t* rowp = base + (long)r * wstride;
    uint32_t w0 = __ldg(rowp + 0), w1 = __ldg(rowp + 1), w2 = __ldg(rowp + 2), w3 = __ldg(rowp + 3), w4 = __ldg(rowp + 4),
             w5 = __ldg(rowp + 5);
    w0 = __funnelshift_r(w0, w1, sh);
    w1 = __funnelshift_r(w1, w2, sh);
    w2 = __funnelshift_r(w2, w3, sh);
    w3 = __funnelshift_r(w3, w4, sh);
    w4 = __funnelshift_r(w4, w5, sh);
    w5 = __funnelshift_r(w5, w5, sh);
    int h0, h1, h2, h3;
    mm_fir4((int)w0, (int)w1, (int)w2, (int)w3, (int)w4, (int)w5, th, f.off1, h0, h1, h2, h3);
    h0 >>= f.shift1; h1 >>= f.shift1; h2 >>= f.shift1; h3 >>= f.shift1;
    if (r & 1)
    {
      P[r >> 1][0] = __byte_perm(hprev[0], h0, 0x5410);
      P[r >> 1][1] = __byte_perm(hprev[1], h1, 0x5410);
      P[r >> 1][2] = __byte_perm(hprev[2], h2, 0x5410);
      P[r >> 1][3] = __byte_perm(hprev[3], h3, 0x5410);
    }
    else if (r == 10)
    {
      P[5][0] = h0 & 0xffff; P[5][1] = h1 & 0xffff; P[5][2] = h2 & 0xffff; P[5][3] = h3 & 0xffff;
    }
    else
    {
      hprev[0] = h0; hprev[1] = h1; hprev[2] = h2; hprev[3] = h3;
    }
  }
  const bool copy = (xFrac | yFrac) == 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
  {
    int o0, o1, o2, o3;
    mm_fir4(P[0][c], P[1][c], P[2][c], P[3][c], P[4][c], P[5][c], tv, BI ? 0 : f.off2, o0, o1, o2, o3);
    if (BI)
    {
      pred[0 * 4 + c] = (int)(int16_t)(o0 >> f.shift2bi);
      pred[1 * 4 + c] = (int)(int16_t)(o1 >> f.shift2bi);
      pred[2 * 4 + c] = (int)(int16_t)(o2 >> f.shift2bi);
      pred[3 * 4 + c] = (int)(int16_t)(o3 >> f.shift2bi);
    }
    else
    {
      o0 = (int)(int16_t)(o0 >> f.shift2); o1 = (int)(int16_t)(o1 >> f.shift2);
      o2 = (int)(int16_t)(o2 >> f.shift2); o3 = (int)(int16_t)(o3 >> f.shift2);
      pred[0 * 4 + c] = copy ? o0 : mm_clip(o0, lo, hi);
      pred[1 * 4 + c] = copy ? o1 : mm_clip(o1, lo, hi);
      pred[2 * 4 + c] = copy ? o2 : mm_clip(o2, lo, hi);
      pred[3 * 4 + c] = copy ? o3 : mm_clip(o3, lo, hi);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Distortion pieces
// -------------------------------------------------------------------------------------------------
// 4x4 2-D Hadamard in place (d[r*4+c]); coefficient order/sign is irrelevant for the sum of magnitudes,
// d[0] ends up as the DC term (sum of all differences).
__device__ __forceinline__ void mm_had4x4(int* d)
{
#pragma unroll
  for (int r = 0; r < 4; ++r)
  {
    const int a = d[r * 4 + 0] + d[r * 4 + 1], b = d[r * 4 + 0] - d[r * 4 + 1];
    const int c = d[r * 4 + 2] + d[r * 4 + 3], e = d[r * 4 + 2] - d[r * 4 + 3];
    d[r * 4 + 0] = a + c; d[r * 4 + 1] = b + e; d[r * 4 + 2] = a - c; d[r * 4 + 3] = b - e;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
  {
    const int a = d[0 * 4 + c] + d[1 * 4 + c], b = d[0 * 4 + c] - d[1 * 4 + c];
    const int g = d[2 * 4 + c] + d[3 * 4 + c], e = d[2 * 4 + c] - d[3 * 4 + c];
    d[0 * 4 + c] = a + g; d[1 * 4 + c] = b + e; d[2 * 4 + c] = a - g; d[3 * 4 + c] = b - e;
  }
}
// one cross-lane butterfly stage of the tile transform: lanes `lane` and `lane^bit` hold the 16
// coefficients of neighbouring sub-blocks; afterwards the lower lane holds sums/differences of
// elements 0..7, the upper lane those of elements 8..15 (8 shuffles instead of 16).
__device__ __forceinline__ void mm_had_xlane(int* d, int bit, bool upper)
{
#pragma unroll
  for (int i = 0; i < 8; ++i)
  {
    const int send = upper ? d[i] : d[i + 8];
    const int keep = upper ? d[i + 8] : d[i];
    const int recv = __shfl_xor_sync(0xffffffffu, send, bit);
    d[i] = keep + recv;
    d[i + 8] = keep - recv;
  }
}

enum { MM_MODE_DIST = 0, MM_MODE_PRED = 1 };

// -------------------------------------------------------------------------------------------------
// k_me: one warp per work item.
// -------------------------------------------------------------------------------------------------
#ifndef MM_ME_MINBLOCKS
#define MM_ME_MINBLOCKS 8
#endif
// MODEL >= 0: kernel specialised for one motion model (work items are grouped by model, the host launches one
// kernel per model present in the batch: smaller code per kernel, no model dispatch); MODEL < 0: generic.
template <int MODE, int MODEL>
__global__ void __launch_bounds__(128, MM_ME_MINBLOCKS) k_me(const MmParams P)
{
  __shared__ MmTaps s_taps[16];
  mm_build_taps(s_taps);
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= P.nItems) return;
  const MmItem it = P.items[warp];
  const mmgpu_block b = P.blocks[it.block];
  const int n = (b.w >> 2) * (b.h >> 2);
  int cand, sb;
  if (n >= 32) { cand = it.cand; sb = it.sb + lane; }
  else { cand = it.cand + lane / n; sb = lane % n; }
  const bool valid = cand < b.cand_first + b.cand_count;
  const int candc = valid ? cand : it.cand;  // invalid lanes redo a valid candidate, result discarded
  const MmSub s = mm_sub_from_index(b.w, b.h, sb);
  const MmCand c = P.cands[candc];
  const MmRef ref = P.refs[b.ref_slot];

  int fx, fy;
  bool oob = false;
  const int model = (MODEL >= 0) ? MODEL : b.model;
  if (model == MM_CLASSIC)
  {
    fx = ((b.x + s.col * 4) << 4) + __float_as_int(c.p[0]);
    fy = ((b.y + s.row * 4) << 4) + __float_as_int(c.p[1]);
  }
  else
  {
    const float* R = mm_rot_ptr(P.rot, model, b.rot_slot);
    mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, 0, 0, 0, b.x, b.y, b.w, b.h, c, R, s.row, s.col, &fx, &fy);
    const int mc = (P.oobMode == MMGPU_OOB_ME) ? 0 : P.maxCU;
    const int xP = fx >> 4, yP = fy >> 4;
    oob = (xP < -mc || yP < -mc || xP >= P.picW + mc - 4 || yP >= P.picH + mc - 4);
  }
  int pred[16];
  if (oob)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i) pred[i] = 0;
  }
  else
  {
    mm_fir_luma4x4<false>(P.fir, s_taps, ref.plane00[0], ref.stride[0], fx >> 4, fy >> 4, fx & 15, fy & 15, P.clpMin,
                          P.clpMax, pred);
  }

  if (MODE == MM_MODE_PRED)
  {
    if (!valid) return;
    int16_t* d = P.dst + (long)(b.y + s.row * 4) * P.dstStride + (b.x + s.col * 4);
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
      const uint32_t lo = (uint32_t)(pred[r * 4 + 0] & 0xffff) | ((uint32_t)pred[r * 4 + 1] << 16);
      const uint32_t hi = (uint32_t)(pred[r * 4 + 2] & 0xffff) | ((uint32_t)pred[r * 4 + 3] << 16);
      *reinterpret_cast<uint2*>(d + (long)r * P.dstStride) = make_uint2(lo, hi);
    }
    return;
  }

  // differences org - pred
  const int16_t* o = P.org + (long)(b.org_y + s.row * 4) * P.orgStride + (b.org_x + s.col * 4);
  int d[16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
  {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(o + (long)r * P.orgStride));
    d[r * 4 + 0] = (int)(int16_t)(v.x & 0xffff) - pred[r * 4 + 0];
    d[r * 4 + 1] = (int)(int16_t)(v.x >> 16) - pred[r * 4 + 1];
    d[r * 4 + 2] = (int)(int16_t)(v.y & 0xffff) - pred[r * 4 + 2];
    d[r * 4 + 3] = (int)(int16_t)(v.y >> 16) - pred[r * 4 + 3];
  }
  unsigned int part;
  if (P.distKind == MMGPU_DIST_SAD)
  {
    // RdCost::xGetSAD*: rows stepped by 1 << subShift, sum << subShift (block rows are 4-aligned)
    unsigned int sad = 0;
    const int step = 1 << P.subShift;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if ((r & (step - 1)) == 0)
        sad += abs(d[r * 4 + 0]) + abs(d[r * 4 + 1]) + abs(d[r * 4 + 2]) + abs(d[r * 4 + 3]);
    part = sad << P.subShift;
  }
  else
  {
    mm_had4x4(d);
    // tile transform across the T adjacent lanes (q = lane & (T-1)); uniform per warp (same block)
    if (s.T >= 2) mm_had_xlane(d, 1, (lane & 1) != 0);
    if (s.T >= 4) mm_had_xlane(d, 2, (lane & 2) != 0);
    if (s.T >= 8) mm_had_xlane(d, 4, (lane & 4) != 0);
    unsigned int sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) sum += (unsigned int)abs(d[i]);
    if (s.q == 0)
    {
      const unsigned int dc = (unsigned int)abs(d[0]);  // JVET_R0164_MEAN_SCALED_SATD
      sum = sum - dc + (dc >> 2);
    }
    if (s.T >= 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (s.T >= 4) sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (s.T >= 8) sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    unsigned int t;
    if (s.tileKind == 0) t = (sum + 1) >> 1;
    else if (s.tileKind == 1) t = (sum + 2) >> 2;
    else if (P.hadMode == MMGPU_HAD_SIMD)
    {
      t = (unsigned int)(((unsigned long long)sum * 0xb504f334ULL) >> 32);
      t = (s.tileKind == 2) ? (t >> 2) : (t >> 1);
    }
    else
    {
      const double div = (s.tileKind == 2) ? 0x1.6a09e667f3bcdp+3 /* sqrt(16.0*8) */ : 0x1.6a09e667f3bcdp+2 /* sqrt(4.0*8) */;
      t = (unsigned int)(int)__dmul_rn(__ddiv_rn((double)sum, div), 2.0);
    }
    part = (s.q == 0) ? t : 0u;
  }
  // candidate reduction
  if (n >= 32)
  {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) atomicAdd(&P.dist[cand], (unsigned long long)part);
  }
  else
  {
    for (int off = n >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (valid && sb == 0) P.dist[cand] = (unsigned long long)part;
  }
}

// -------------------------------------------------------------------------------------------------
// k_dmvr: xProcessDMVRProjected's cost evaluations (InterPrediction.cpp:2510-2567).  One group of 2n lanes per
// (sub-PU, search offset): n lanes interpolate the L0 prediction of the sub-PU's n sub-blocks at mv0 + offset,
// n lanes the L1 prediction at mv1 - offset (bi = true: 14-bit intermediates, +-maxCU OOB rule, the sub-PU is
// "the block": quirk Q6); xDMVRCost = SAD over even rows (subShift 1), << 1, >> 1.
// -------------------------------------------------------------------------------------------------
struct MmDmvrItem
{
  int32_t subpu;   // index into the sub-PU array
  int32_t off;     // first m_pSearchOffset index handled by this warp
  int32_t pad0, pad1;
};

__global__ void __launch_bounds__(128) k_dmvr(const MmParams P, const mmgpu_dmvr_subpu* __restrict__ subpus,
                                              const MmDmvrItem* __restrict__ items, int nItems, int simdSad,
                                              unsigned long long* __restrict__ costs)
{
  __shared__ MmTaps s_taps[16];
  mm_build_taps(s_taps);
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nItems) return;
  const MmDmvrItem it = items[warp];
  const mmgpu_dmvr_subpu sp = subpus[it.subpu];
  const int cols = sp.w >> 2, rows = sp.h >> 2;
  const int n = cols * rows;                    // 4, 8 or 16 sub-blocks
  const int g = lane / (2 * n), within = lane - g * 2 * n;
  const int list = within / n, sb = within - list * n;
  const int off = it.off + g;
  const bool valid = off < 25;
  const int offc = valid ? off : it.off;
  const int oh = (offc % 5 - 2) * 16, ov = (offc / 5 - 2) * 16;   // m_pSearchOffset << MV_FRACTIONAL_BITS_INTERNAL
  const int mh = list ? (sp.mv1.hor - oh) : (sp.mv0.hor + oh);
  const int mv = list ? (sp.mv1.ver - ov) : (sp.mv0.ver + ov);
  const int row = sb / cols, col = sb - row * cols;
  const MmRef ref = P.refs[list ? sp.ref_slot1 : sp.ref_slot0];
  const float* R = mm_rot_ptr(P.rot, sp.model, list ? sp.rot_slot1 : sp.rot_slot0);
  MmCand c;
  mm_cand_setup(P.g, sp.model, true, sp.x, sp.y, sp.w, sp.h, mh, mv, R, &c);
  int fx, fy;
  mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, sp.model, 0, 0, 0, sp.x, sp.y, sp.w, sp.h, c, R, row, col, &fx, &fy);
  const int xP = fx >> 4, yP = fy >> 4;
  const bool oob = (xP < -P.maxCU || yP < -P.maxCU || xP >= P.picW + P.maxCU - 4 || yP >= P.picH + P.maxCU - 4);
  int pred[16];
  if (oob)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i) pred[i] = 0;
  }
  else
  {
    mm_fir_luma4x4<true>(P.fir, s_taps, ref.plane00[0], ref.stride[0], xP, yP, fx & 15, fy & 15, P.clpMin, P.clpMax, pred);
  }
  // rows 0 and 2 of the sub-block (even rows of the sub-PU): L0 lanes receive the L1 samples of lane ^ n
  unsigned int sad = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
  {
    const int mine = pred[(i >> 2) * 8 + (i & 3)];
    const int other = __shfl_xor_sync(0xffffffffu, mine, n);
    int d = mine - other;
    if (simdSad) { d = (int)(int16_t)d; d = abs(d) & 0xffff; }   // _mm_sub_epi16 + _mm_abs_epi16, zero-extended
    else d = abs(d);
    sad += (unsigned int)d;
  }
  if (list) sad = 0;
  for (int o = n >> 1; o > 0; o >>= 1) sad += __shfl_xor_sync(0xffffffffu, sad, o);
  if (valid && within == 0) costs[(long)it.subpu * 25 + off] = ((unsigned long long)sad << 1) >> 1;
}

// -------------------------------------------------------------------------------------------------
// Picture::extendPicBorder on the device (Picture.cpp:1008-1046): one thread per margin sample
// -------------------------------------------------------------------------------------------------
__global__ void k_extend_border(int16_t* plane00, int stride, int w, int h, int mx, int my)
{
  const int X = blockIdx.x * blockDim.x + threadIdx.x - mx;
  const int Y = blockIdx.y * blockDim.y + threadIdx.y - my;
  if (X >= w + mx || Y >= h + my) return;
  if (X >= 0 && X < w && Y >= 0 && Y < h) return;
  const int sx = min(max(X, 0), w - 1), sy = min(max(Y, 0), h - 1);
  plane00[(long)Y * stride + X] = plane00[(long)sy * stride + sx];
}

__global__ void k_add_avg(const int16_t* s0, int st0, const int16_t* s1, int st1, int16_t* dst, int ds, int w, int h,
                          int shift, int offset, int lo, int hi)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(long)y * ds + x] = (int16_t)mm_clip(((int)s0[(long)y * st0 + x] + (int)s1[(long)y * st1 + x] + offset) >> shift, lo, hi);
}
