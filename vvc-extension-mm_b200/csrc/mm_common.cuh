// mm_common.cuh -- types, tables and index helpers shared by the sm_100a kernels of the multi-model MC +
// distortion path, and the host-side launch wrappers of the three kernel translation units:
//
//   mm_kernels_map.cu   k_mpa_cache     MotionPlaneAdaptiveMotionModel::fillCache (MotionPlaneAdaptiveMotionModel.cpp:97-104)
//                       k_trig_tables   ERP toSphere angles of the picture grid (Projection.cpp:213-221)
//                       k_build_items   work list of a batch (one warp of 32 sub-block-candidates per item)
//                       k_cand_setup    per-(block,mv) scalars of the motion models (mm_map.cuh: mm_cand_setup)
//                       k_reproject     MVReprojection::reprojectMotionVectorSubblocks (MVReprojection.cpp:80-166)
//                       k_map           positions of every sub-block of every candidate of a batch (a1-a3)
//   mm_kernels_fir.cu   k_fir<MODE>     8-tap separable FIR at those positions -> SAD/HAD per candidate
//                                       (InterSearch.cpp:6277-6385 + RdCost.cpp:482-.., 2943-3067) or prediction write-out
//                       k_extend_border Picture::extendPicBorder (Picture.cpp:988-1048)
//                       k_add_avg       PelBuf::addAvg (Buffer.cpp:551-587)
//   mm_kernels_mc.cu    k_mc_generic    xPredInterBlkMM / xPredInterBlk, any component, uni/bi (InterPrediction.cpp:683-856, 928-1096)
//                       k_mc_jobs       whole-picture reconstruction: jobs of any component, uni or bi with fused addAvg
//                       k_dmvr          xProcessDMVRProjected's cost evaluations (InterPrediction.cpp:2510-2567)
//
// HBM layout: reference planes int16, pitch a multiple of 64 samples, margins resident.  A batch is enumerated as
// block -> candidate -> sub-block (tile-major); one warp owns 32 consecutive entries of one block: either 32
// consecutive sub-blocks of one candidate (blocks with >= 32 sub-blocks) or 32/n candidates of a small block.
// k_map writes the 1/16-pel position of every entry (8 B), k_fir reads it back coalesced: the float mapping and
// the integer stencil are separate kernels so that each fits the instruction cache and its own register budget.
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
  int32_t tma;     // 1: the slot's luma plane has TMA descriptors (MmParams::tmaps[slot * MM_TMA_TYPES ..])
  int32_t pad;
};

// Shared-memory boxes of k_fir, filled by TMA (cp.async.bulk.tensor.2d) from the padded luma plane.  One box shape per
// arrangement of a warp's 32 sub-blocks; the box pitch (words) is = 4 mod 8 so that lanes 4 picture rows apart land 16
// banks apart.  Box bytes = w * h * 2 <= MM_TMA_BOX_BYTES.
#define MM_TMA_TYPES 6
#define MM_TMA_BOX_BYTES 5120
// 16x2, 8x4 and 4x8 sub-block arrangements: tight boxes first (less L2 -> shared-memory traffic), then taller ones for
// warps whose motion shears the arrangement vertically
#define MM_TMA_BOX_DIMS {{104, 18}, {56, 32}, {40, 46}, {104, 24}, {56, 44}, {40, 62}}

struct __align__(16) MmItem  // one warp of work: everything k_map / k_fir need to know about their block (two 16-byte loads)
{
  int32_t x, y;      // luma position of the block
  uint32_t wh;       // w | h << 16
  uint32_t rmr;      // ref_slot | model << 16 | rot_slot << 24
  uint32_t orgxy;    // org_x | org_y << 16
  int32_t cand;      // first candidate (global index) handled by this warp
  int32_t sb;        // first sub-block (tile-major index) handled by this warp
  int32_t candEnd;   // cand_first + cand_count of the block
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
  const uint32_t* hTaps;     // mm_build_htaps table (MM_HTAPS_WORDS words) in global memory
  const void* tmaps;         // CUtensorMap[max_refs][MM_TMA_TYPES] in global memory (128 bytes each)
  int32_t tmaOffX, tmaOffY;  // luma margins: tensor coordinates = picture coordinates + offset
  const int16_t* org;
  int32_t orgStride;
  int32_t distKind, subShift, hadMode;
  int32_t clpMin, clpMax;
  int32_t picW, picH, maxCU;
  int32_t oobMode;
  int32_t refOverride;       // >= 0: every item reads this reference slot (multi-reference batches), else the item's own
  int32_t rotOverride;       // >= 0: every GEODESIC_CAMPOSE block uses this rotation slot
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
#define MM_LUMA_FILTER_INIT \
  {0, 0, 0, 64, 0, 0, 0, 0},       {0, 1, -3, 63, 4, -2, 1, 0},      {-1, 2, -5, 62, 8, -3, 1, 0}, \
  {-1, 3, -8, 60, 13, -4, 1, 0},   {-1, 4, -10, 58, 17, -5, 1, 0},   {-1, 4, -11, 52, 26, -8, 3, -1}, \
  {-1, 3, -9, 47, 31, -10, 4, -1}, {-1, 4, -11, 45, 34, -10, 4, -1}, {-1, 4, -11, 40, 40, -11, 4, -1}, \
  {-1, 4, -10, 34, 45, -11, 4, -1}, {-1, 4, -10, 31, 47, -9, 3, -1}, {-1, 3, -8, 26, 52, -11, 4, -1}, \
  {0, 1, -5, 17, 58, -10, 4, -1},  {0, 1, -4, 13, 60, -8, 3, -1},    {0, 1, -3, 8, 62, -5, 2, -1}, \
  {0, 1, -2, 4, 63, -3, 1, 0}
static __constant__ int8_t c_lumaFilter[16][8] = {
  MM_LUMA_FILTER_INIT};
static const int8_t h_lumaFilter[16][8] = {
  MM_LUMA_FILTER_INIT};
static __constant__ int8_t c_chromaFilter[32][4] = {
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

// Zero-fill rule of xPredInterBlkMM (InterPrediction.cpp:780) and xMVReprojectionInterpolation (InterSearch.cpp:6332):
//   xPos < -maxCU || yPos < -maxCU || xPos >= refBuf.width + maxCU - sbW || yPos >= refBuf.height + maxCU - sbH
// refBuf.width/height and the sub-block size are unsigned (SizeType), so the two upper-bound comparisons are unsigned: a
// negative position converts to a huge value and is "beyond" the bound -- the slack exists on the right/bottom only.
__device__ __forceinline__ bool mm_oob(int xP, int yP, int picW, int picH, int mcw, int mch, int sbw, int sbh)
{
  return xP < -mcw || yP < -mch || (uint32_t)xP >= (uint32_t)(picW + mcw - sbw) || (uint32_t)yP >= (uint32_t)(picH + mch - sbh);
}

enum { MM_MODE_DIST = 0, MM_MODE_PRED = 1 };
#define MM_POS_OOB ((int32_t)0x80000000)  // k_map -> k_fir: sub-block outside the picture, prediction := 0

struct MmDmvrItem
{
  int32_t subpu;   // index into the sub-PU array
  int32_t off;     // first m_pSearchOffset index handled by this warp
  int32_t pad0, pad1;
};

// ---- launch wrappers (defined next to their kernels) ---------------------------------------------------------
void mmk_mpa_cache(cudaStream_t s, const MmGeom& g, int model, float2* out, uint8_t* vip);
void mmk_trig_tables(cudaStream_t s, const MmGeom& g, float2* th0, float2* th1, float2* ph0, float2* ph1);
void mmk_build_items(cudaStream_t s, const mmgpu_block* blocks, const int32_t* itemBase, int nBlocks, MmItem* items,
                     int32_t* candBlock);
// dist != nullptr: also zeroes the per-candidate distortion accumulators
void mmk_cand_setup(cudaStream_t s, const MmGeom& g, const mmgpu_block* blocks, const int32_t* candBlock, const mmgpu_mv* mvs,
                    int mvShift, const float* rot, int rotOverride, int nCands, MmCand* out, unsigned long long* dist);
void mmk_reproject(cudaStream_t s, const MmParams& P, int comp, int csx, int csy, int x, int y, int w, int h, mmgpu_mv mv,
                   int model, int rotSlot, int32_t* outX, int32_t* outY);
// positions of items [first, first+count) of P.items; model >= 0: all of them use that motion model
void mmk_map(cudaStream_t s, const MmParams& P, int first, int count, int model, int2* pos);
// FIR + distortion (or prediction write-out) of items [first, first+count)
void mmk_fir(cudaStream_t s, const MmParams& P, int first, int count, int mode, const int2* pos);
void mmk_tan_selftest(cudaStream_t s, uint32_t stride, unsigned long long* out);
#define MM_PAD_MAX 24   // planes per k_pad_planes launch (8 pictures x 3 components)
struct MmPadPlane
{
  const int16_t* src;   // un-padded plane
  int16_t* dst00;       // sample (0,0) of the padded plane
  int32_t srcStride, dstStride, w, h, mx, my, xoff, pad;
};
struct MmPadJobs { MmPadPlane p[MM_PAD_MAX]; };
void mmk_pad_planes(cudaStream_t s, const MmPadJobs& J, int nPlanes, int maxGroups, int maxRows);
void mmk_pull_pad_planes(cudaStream_t s, const MmPadJobs& J, int nPlanes, int maxCtas);
void mmk_multicast_copy(cudaStream_t s, const void* src, void* dstMulticast, size_t bytes, int maxCtas);
void mmk_extend_border(cudaStream_t s, int16_t* plane00, int stride, int w, int h, int mx, int my, int xoff);
void mmk_add_avg(cudaStream_t s, const int16_t* s0, int st0, const int16_t* s1, int st1, int16_t* dst, int ds, int w, int h,
                 int shift, int offset, int lo, int hi);
void mmk_mc_generic(cudaStream_t s, const MmParams& P, int refSlot, int comp, int csx, int csy, int x, int y, int w, int h,
                    mmgpu_mv mv, int model, int rotSlot, int bi, int16_t* dst, int dstStride);
void mmk_bdof_ring(cudaStream_t s, const int16_t* ref, int stride, int x0, int y0, int xOff, int yOff, int w, int h, int shift, int16_t* tmp);
#define MM_JOB_CHUNK 64   // sub-blocks per k_mc_jobs CTA
void mmk_mc_jobs(cudaStream_t s, const MmParams& P, int nItems, const mmgpu_mc_job* jobs, const int2* items, int16_t* dstY, int16_t* dstCb,
                 int16_t* dstCr, int strideY, int strideC, int csxC, int csyC, int avgShift, int avgOffset);
void mmk_dmvr_classic(cudaStream_t s, const MmRef* refs, const mmgpu_dmvr_classic_subpu* subpus, int n, int bd, unsigned long long* costs);
void mmk_dmvr_final_mc(cudaStream_t s, const MmRef* refs, const MmFir& f, const mmgpu_dmvr_mc_job* jobs, int n, int rnd, int lo, int hi,
                       int16_t* dst);
void mmk_dmvr(cudaStream_t s, const MmParams& P, const mmgpu_dmvr_subpu* subpus, const MmDmvrItem* items, int nItems,
              int simdSad, unsigned long long* costs);
