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
