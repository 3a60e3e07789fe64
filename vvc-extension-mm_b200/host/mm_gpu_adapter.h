// mm_gpu_adapter.h -- C++11 host side of the drop-in boundary: the reference's own argument lists
// (Position/Size/Mv/ClpRng/CPelBuf shaped structs, MotionModelID, ComponentID) over the C ABI in
// include/mmgpu.h.  A maintainer calls these from the bodies of
//   InterPrediction::xPredInterBlkMM            (CommonLib/InterPrediction.cpp:683-856)
//   InterSearch::xMVReprojectionInterpolation   (EncoderLib/InterSearch.cpp:6277-6385) + distFunc
//   InterSearch::xPatternRefinementProjected    (EncoderLib/InterSearch.cpp:6222-6249)
//   MVReprojection::init / Slice::constructRefPicList hooks
// (see INTEGRATION.md).  Error behaviour follows the reference's CHECK/THROW: a failing call throws
// (std::runtime_error) and, as in the reference, nothing catches it before main().
#pragma once
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/mmgpu.h"

namespace mmgpu_host
{
struct Position { int x, y; };
struct Size { unsigned width, height; };
struct Mv { int hor, ver; };                       // CommonLib/Mv.h, MV_PRECISION_INTERNAL unless stated
struct ClpRng { int min, max, bd; };               // CommonLib/Common.h
struct CPelBuf { const int16_t* buf; int stride; unsigned width, height; };   // AreaBuf<const Pel>
struct PelBuf { int16_t* buf; int stride; unsigned width, height; };
typedef uint64_t Distortion;                       // TypeDef.h:282
enum MvPrecision { MV_PRECISION_INTERNAL = 0, MV_PRECISION_QUARTER = 2, MV_PRECISION_HALF = 3, MV_PRECISION_INT = 4 };  // value = left shift to internal

// ---- a14: candidate enumerators of the integer motion search (data only) + the compare chain of xTZSearchHelp ------------
// Each enumerator lists, in the reference's visiting order, the points one call of the reference function evaluates; the
// whole list goes to the device as ONE batch (evalCandidates) and tzReplay() then walks the reference's compare chain over
// the returned distortions, so the search takes the same decisions with the same uiBestSad.
struct SearchRange { int left, right, top, bottom; };           // EncoderLib/InterSearch.h SearchRange
struct TzPoint { int x, y; uint8_t pointNr; unsigned distance; };
struct TzBest                                                    // the IntTZSearchStruct fields xTZSearchHelp updates
{
  Distortion uiBestSad; int iBestX, iBestY; unsigned uiBestDistance, uiBestRound; uint8_t ucPointNr;
};

// xTZ8PointDiamondSearch (EncoderLib/InterSearch.cpp:535-737)
inline void tzDiamondRound(const SearchRange& sr, int startX, int startY, int dist, bool cornersAtDist1, std::vector<TzPoint>& out)
{
  const int top = startY - dist, bottom = startY + dist, left = startX - dist, right = startX + dist;
  auto add = [&](int x, int y, int nr, int d) { out.push_back(TzPoint{x, y, (uint8_t)nr, (unsigned)d}); };
  if (dist == 1)
  {
    if (top >= sr.top)
    {
      if (cornersAtDist1)
      {
        if (left >= sr.left) add(left, top, 1, dist);
        add(startX, top, 2, dist);
        if (right <= sr.right) add(right, top, 3, dist);
      }
      else add(startX, top, 2, dist);
    }
    if (left >= sr.left) add(left, startY, 4, dist);
    if (right <= sr.right) add(right, startY, 5, dist);
    if (bottom <= sr.bottom)
    {
      if (cornersAtDist1)
      {
        if (left >= sr.left) add(left, bottom, 6, dist);
        add(startX, bottom, 7, dist);
        if (right <= sr.right) add(right, bottom, 8, dist);
      }
      else add(startX, bottom, 7, dist);
    }
    return;
  }
  const bool inside = top >= sr.top && left >= sr.left && right <= sr.right && bottom <= sr.bottom;
  if (dist <= 8)
  {
    const int h = dist >> 1, top2 = startY - h, bottom2 = startY + h, left2 = startX - h, right2 = startX + h;
    if (inside)
    {
      add(startX, top, 2, dist); add(left2, top2, 1, h); add(right2, top2, 3, h); add(left, startY, 4, dist);
      add(right, startY, 5, dist); add(left2, bottom2, 6, h); add(right2, bottom2, 8, h); add(startX, bottom, 7, dist);
      return;
    }
    if (top >= sr.top) add(startX, top, 2, dist);
    if (top2 >= sr.top)
    {
      if (left2 >= sr.left) add(left2, top2, 1, h);
      if (right2 <= sr.right) add(right2, top2, 3, h);
    }
    if (left >= sr.left) add(left, startY, 4, dist);
    if (right <= sr.right) add(right, startY, 5, dist);
    if (bottom2 <= sr.bottom)
    {
      if (left2 >= sr.left) add(left2, bottom2, 6, h);
      if (right2 <= sr.right) add(right2, bottom2, 8, h);
    }
    if (bottom <= sr.bottom) add(startX, bottom, 7, dist);
    return;
  }
  // dist > 8: the four tips, then three rings of four points along the edges of the diamond
  if (inside)
  {
    add(startX, top, 0, dist); add(left, startY, 0, dist); add(right, startY, 0, dist); add(startX, bottom, 0, dist);
    for (int i = 1; i < 4; ++i)
    {
      const int q = (dist >> 2) * i, yT = top + q, yB = bottom - q, xL = startX - q, xR = startX + q;
      add(xL, yT, 0, dist); add(xR, yT, 0, dist); add(xL, yB, 0, dist); add(xR, yB, 0, dist);
    }
    return;
  }
  if (top >= sr.top) add(startX, top, 0, dist);
  if (left >= sr.left) add(left, startY, 0, dist);
  if (right <= sr.right) add(right, startY, 0, dist);
  if (bottom <= sr.bottom) add(startX, bottom, 0, dist);
  for (int i = 1; i < 4; ++i)
  {
    const int q = (dist >> 2) * i, yT = top + q, yB = bottom - q, xL = startX - q, xR = startX + q;
    if (yT >= sr.top)
    {
      if (xL >= sr.left) add(xL, yT, 0, dist);
      if (xR <= sr.right) add(xR, yT, 0, dist);
    }
    if (yB <= sr.bottom)
    {
      if (xL >= sr.left) add(xL, yB, 0, dist);
      if (xR <= sr.right) add(xR, yB, 0, dist);
    }
  }
}

// xTZ2PointSearch (InterSearch.cpp:456-481): the two untested neighbours of the best point of a distance-1 round
inline void tz2Point(const SearchRange& sr, int bestX, int bestY, uint8_t pointNr, std::vector<TzPoint>& out)
{
  static const int xo[2][9] = {{0, -1, -1, 0, -1, +1, -1, -1, +1}, {0, 0, +1, +1, -1, +1, 0, +1, 0}};
  static const int yo[2][9] = {{0, 0, -1, -1, +1, -1, 0, +1, 0}, {0, -1, -1, 0, -1, +1, +1, +1, +1}};
  for (int k = 0; k < 2; ++k)
  {
    const int x = bestX + xo[k][pointNr], y = bestY + yo[k][pointNr];
    if (x >= sr.left && x <= sr.right && y >= sr.top && y <= sr.bottom) out.push_back(TzPoint{x, y, 0, 2});
  }
}

// raster scan of xTZSearch (InterSearch.cpp:5611-5647) and, with step 1, the full grid of xPatternSearch (:5287-5322)
inline void tzRaster(const SearchRange& sr, int step, std::vector<TzPoint>& out)
{
  for (int y = sr.top; y <= sr.bottom; y += step)
    for (int x = sr.left; x <= sr.right; x += step) out.push_back(TzPoint{x, y, 0, (unsigned)step});
}

// xTZSearchHelp (InterSearch.cpp:363-453, the branch taken by every motion model but CLASSIC with subShiftMode 1):
// motion cost only added when the raw distortion already beats the best; strict "<" twice.
template <typename MvCost>
inline void tzReplay(const std::vector<TzPoint>& pts, const Distortion* dist, MvCost mvCost, TzBest& b)
{
  for (size_t i = 0; i < pts.size(); ++i)
  {
    Distortion sad = dist[i];
    if (sad < b.uiBestSad)
    {
      sad += mvCost(pts[i].x, pts[i].y);
      if (sad < b.uiBestSad)
      {
        b.uiBestSad = sad; b.iBestX = pts[i].x; b.iBestY = pts[i].y; b.uiBestDistance = pts[i].distance; b.uiBestRound = 0;
        b.ucPointNr = pts[i].pointNr;
      }
    }
  }
}
// xPatternSearch's chain (InterSearch.cpp:5299-5317): cost always added, strict "<"
template <typename MvCost>
inline void fullSearchReplay(const std::vector<TzPoint>& pts, const Distortion* dist, MvCost mvCost, Distortion& best, int& bestX, int& bestY)
{
  for (size_t i = 0; i < pts.size(); ++i)
  {
    const Distortion sad = dist[i] + mvCost(pts[i].x, pts[i].y);
    if (sad < best) { best = sad; bestX = pts[i].x; bestY = pts[i].y; }
  }
}

class GpuInterPrediction
{
public:
  GpuInterPrediction() : m_ctx(nullptr), m_chFmt(1) {}
  ~GpuInterPrediction() { mmgpu_destroy(m_ctx); }
  GpuInterPrediction(const GpuInterPrediction&) = delete;
  GpuInterPrediction& operator=(const GpuInterPrediction&) = delete;

  // MVReprojection::init(projection, resolution, sps, epipoleList)  (MVReprojection.cpp:7-65)
  void init(int picW, int picH, int chromaFormat, int internalBitDepth, int maxCUWidth, int margin, float mmOffset4x4,
            int gedFlavor, int maxRefs, int device = 0)
  {
    mmgpu_config c;
    c.pic_w = picW; c.pic_h = picH; c.chroma_fmt = chromaFormat; c.bit_depth = internalBitDepth; c.max_cu = maxCUWidth;
    c.margin = margin; c.offset4x4 = mmOffset4x4; c.ged_flavor = gedFlavor; c.max_refs = maxRefs; c.device = device;
    c.libm_fma = -1; c.had_mode = -1;
    mmgpu_destroy(m_ctx);
    m_ctx = nullptr;
    if (mmgpu_create(&c, &m_ctx) != MMGPU_OK) throw std::runtime_error(std::string("mmgpu_create: ") + mmgpu_last_error(nullptr));
    m_chFmt = chromaFormat;
  }
  bool isInitialized() const { return m_ctx != nullptr; }

  // after Picture::extendPicBorder: refPic->getRecoBuf(compID, wrapRef) for Y, Cb, Cr
  void uploadReference(int slot, int poc, const CPelBuf reco[3])
  {
    const int16_t* p[3] = {reco[0].buf, reco[1].buf, reco[2].buf};
    const int32_t s[3] = {reco[0].stride, reco[1].stride, reco[2].stride};
    chk(mmgpu_set_ref(m_ctx, slot, poc, p, s));
  }
  // EpipoleList::findEpipole (Q24 fixed) -> GeodesicMotionModel::setEpipole
  void setEpipoleFixed(int rotSlot, const int epipoleQ24[3])
  {
    const float e[3] = {mmgpu_fixed_to_float(epipoleQ24[0], 24), mmgpu_fixed_to_float(epipoleQ24[1], 24),
                        mmgpu_fixed_to_float(epipoleQ24[2], 24)};
    float R[9];
    mmgpu_rodrigues(e, R);
    chk(mmgpu_set_rotation(m_ctx, rotSlot, R));
  }

  // InterPrediction::xPredInterBlkMM(compID, pu, refPic, mv, dstPic, motionModel, bi, clpRng, ...)
  void xPredInterBlkMM(int compID, const Position& blockPos, const Size& blockSize, int refSlot, const Mv& mv, PelBuf& dstBuf,
                       int motionModel, bool bi, const ClpRng& clpRng, int rotSlot = 0)
  {
    const mmgpu_mv m = {mv.hor, mv.ver};
    chk(mmgpu_mc(m_ctx, refSlot, compID, m_chFmt, blockPos.x, blockPos.y, (int)blockSize.width, (int)blockSize.height, m,
                 motionModel, rotSlot, bi ? 1 : 0, clpRng.min, clpRng.max, MMGPU_OOB_PRED, dstBuf.buf, dstBuf.stride));
  }
  // InterSearch::xMVReprojectionInterpolation(cuPosition, cuSize, refBuf, mv, mvPrec, dstBuf, motionModel, clpRng, ...)
  void xMVReprojectionInterpolation(const Position& cuPosition, const Size& cuSize, int refSlot, const Mv& mv, MvPrecision mvPrec,
                                    PelBuf& dstBuf, int motionModel, const ClpRng& clpRng, int rotSlot = 0)
  {
    const mmgpu_mv m = {mv.hor * (1 << (int)mvPrec), mv.ver * (1 << (int)mvPrec)};   // Mv::changePrecision
    chk(mmgpu_mc(m_ctx, refSlot, 0, 0, cuPosition.x, cuPosition.y, (int)cuSize.width, (int)cuSize.height, m, motionModel,
                 rotSlot, 0, clpRng.min, clpRng.max, MMGPU_OOB_ME, dstBuf.buf, dstBuf.stride));
  }
  // PelBuf::addAvg via xWeightedAverage
  void addAvg(const CPelBuf& s0, const CPelBuf& s1, PelBuf& dst, const ClpRng& clpRng)
  {
    chk(mmgpu_add_avg(m_ctx, s0.buf, s0.stride, s1.buf, s1.stride, dst.buf, dst.stride, (int)dst.width, (int)dst.height,
                      clpRng.min, clpRng.max));
  }

  // xApplyMvVA + m_cDistParam.distFunc for every MV of a candidate set the search would evaluate anyway
  // (diamond round, raster line, 9-point refinement, 81-point bi-pred full search): one call, then the
  // caller replays its compare chain in reference order over `dist`.
  void evalCandidates(const Position& blkPos, const Size& blkSize, int refSlot, int motionModel, const CPelBuf& pcPatternKey,
                      const std::vector<Mv>& mvs, MvPrecision mvPrec, bool useHadamard, int subShift, const ClpRng& clpRng,
                      std::vector<Distortion>& dist, int rotSlot = 0)
  {
    chk(mmgpu_set_org(m_ctx, pcPatternKey.buf, pcPatternKey.stride, (int)pcPatternKey.width, (int)pcPatternKey.height));
    mmgpu_block b;
    b.x = blkPos.x; b.y = blkPos.y; b.w = (int)blkSize.width; b.h = (int)blkSize.height; b.model = motionModel;
    b.ref_slot = refSlot; b.rot_slot = rotSlot; b.cand_first = 0; b.cand_count = (int)mvs.size(); b.org_x = 0; b.org_y = 0;
    b.reserved = 0;
    std::vector<mmgpu_mv> m(mvs.size());
    for (size_t i = 0; i < mvs.size(); ++i) { m[i].hor = mvs[i].hor; m[i].ver = mvs[i].ver; }
    dist.assign(mvs.size(), 0);
    chk(mmgpu_me_batch(m_ctx, 1, &b, (int)m.size(), m.data(), (int)mvPrec, useHadamard ? MMGPU_DIST_HAD : MMGPU_DIST_SAD,
                       subShift, clpRng.min, clpRng.max, dist.data()));
  }

  // InterSearch::xPatternRefinementProjected (InterSearch.cpp:6222-6249): 9 candidates around rcMvBase, best-first-wins
  // compare in the reference's order; mvCost(i) = m_pcRdCost->getCostOfVectorWithPredictor(mv.hor, mv.ver, 0).
  template <typename MvCost>
  Distortion xPatternRefinementProjected(const Position& blkPos, const Size& blkSize, int refSlot, int motionModel,
                                         const CPelBuf& pcPatternKey, MvPrecision refinementPrecision, const Mv& rcMvBase,
                                         const Mv pcMvRefine[9], bool useHadamard, const ClpRng& clpRng, MvCost mvCost,
                                         Mv& rcMvRefined, int rotSlot = 0)
  {
    std::vector<Mv> cand(9);
    for (int i = 0; i < 9; ++i) { cand[i].hor = rcMvBase.hor + pcMvRefine[i].hor; cand[i].ver = rcMvBase.ver + pcMvRefine[i].ver; }
    std::vector<Distortion> d;
    evalCandidates(blkPos, blkSize, refSlot, motionModel, pcPatternKey, cand, refinementPrecision, useHadamard, 0, clpRng, d, rotSlot);
    Distortion best = std::numeric_limits<Distortion>::max();
    for (int i = 0; i < 9; ++i)
    {
      const Distortion u = d[i] + mvCost(cand[i]);
      if (u < best) { best = u; rcMvRefined = cand[i]; }
    }
    return best;
  }

  // One call of xTZ8PointDiamondSearch / raster / 2-point on the device: the enumerated points as one batch, then the
  // reference's compare chain (uiBestRound is incremented by the diamond round itself, InterSearch.cpp:549).
  template <typename MvCost>
  void tzEvaluate(const std::vector<TzPoint>& pts, const Position& blkPos, const Size& blkSize, int refSlot, int motionModel,
                  const CPelBuf& pcPatternKey, int subShift, const ClpRng& clpRng, MvCost mvCost, TzBest& best, int rotSlot = 0)
  {
    if (pts.empty()) return;
    std::vector<Mv> mvs(pts.size());
    for (size_t i = 0; i < pts.size(); ++i) { mvs[i].hor = pts[i].x; mvs[i].ver = pts[i].y; }
    std::vector<Distortion> d;
    evalCandidates(blkPos, blkSize, refSlot, motionModel, pcPatternKey, mvs, MV_PRECISION_INT, false, subShift, clpRng, d, rotSlot);
    tzReplay(pts, d.data(), mvCost, best);
  }

  // InterPrediction::xPredInterBlk (classic translational MC, InterPrediction.cpp:928-1096): plain, with the alternative half-pel
  // filter (pu.cu->imv == IMV_HPEL, luma) or with bioApplied (luma: prediction + ring into the (w + 4)-stride BDOF work buffer)
  void xPredInterBlk(int compID, const Position& blockPos, const Size& blockSize, int refSlot, const Mv& mv, PelBuf& dstBuf, bool bi,
                     const ClpRng& clpRng, bool useAltHpelIf = false)
  {
    const mmgpu_mv m = {mv.hor, mv.ver};
    if (useAltHpelIf && compID == 0)
      chk(mmgpu_mc_alt_hpel(m_ctx, refSlot, blockPos.x, blockPos.y, (int)blockSize.width, (int)blockSize.height, m, bi ? 1 : 0, clpRng.min,
                            clpRng.max, dstBuf.buf, dstBuf.stride));
    else
      chk(mmgpu_mc(m_ctx, refSlot, compID, m_chFmt, blockPos.x, blockPos.y, (int)blockSize.width, (int)blockSize.height, m, MMGPU_CLASSIC, 0,
                   bi ? 1 : 0, clpRng.min, clpRng.max, MMGPU_OOB_PRED, dstBuf.buf, dstBuf.stride));
  }
  void xPredInterBlkBdof(const Position& blockPos, const Size& blockSize, int refSlot, const Mv& mv, const ClpRng& clpRng, int16_t* bdofBuf)
  {
    const mmgpu_mv m = {mv.hor, mv.ver};
    chk(mmgpu_mc_bdof(m_ctx, refSlot, blockPos.x, blockPos.y, (int)blockSize.width, (int)blockSize.height, m, clpRng.min, clpRng.max, bdofBuf));
  }
  // MVReprojection::motionVectorInDesiredMotionModel (MVReprojection.cpp:168-217)
  Mv motionVectorInDesiredMotionModel(const Position& position, const Mv& mvOrig, int modelOrig, int modelDesired, int shiftHor, int shiftVer,
                                      int rotSlotOrig, int rotSlotDesired, const Position& candPos, const Size& candSize,
                                      const Position& curPos, const Size& curSize)
  {
    const mmgpu_mv m = {mvOrig.hor, mvOrig.ver};
    mmgpu_mv out = {0, 0};
    chk(mmgpu_mvp_convert(m_ctx, position.x, position.y, m, modelOrig, modelDesired, shiftHor, shiftVer, rotSlotOrig, rotSlotDesired, candPos.x,
                          candPos.y, (int)candSize.width, (int)candSize.height, curPos.x, curPos.y, (int)curSize.width, (int)curSize.height, &out));
    return Mv{out.hor, out.ver};
  }
  // InterPrediction::xProcessDMVR (classic, InterPrediction.cpp:2239-2439): the integer search of a PU's sub-PUs in one call
  // (costs[i * 25 + k], then mmgpu_dmvr_decide per sub-PU) and the final MC of the (list, component) jobs of a sub-PU
  void dmvrClassicCosts(const std::vector<mmgpu_dmvr_classic_subpu>& subPus, std::vector<uint64_t>& costs)
  {
    costs.assign(subPus.size() * 25, 0);
    if (!subPus.empty()) chk(mmgpu_dmvr_classic_costs(m_ctx, (int)subPus.size(), subPus.data(), costs.data()));
  }
  void dmvrFinalMc(const std::vector<mmgpu_dmvr_mc_job>& jobs, std::vector<int16_t>& out)
  {
    if (!jobs.empty()) chk(mmgpu_dmvr_final_mc(m_ctx, (int)jobs.size(), jobs.data(), out.data(), (int)out.size()));
  }

  mmgpu_ctx* ctx() { return m_ctx; }

private:
  void chk(int rc)
  {
    if (rc != MMGPU_OK) throw std::runtime_error(std::string("mmgpu: ") + mmgpu_last_error(m_ctx));
  }
  mmgpu_ctx* m_ctx;
  int m_chFmt;
};
}  // namespace mmgpu_host
