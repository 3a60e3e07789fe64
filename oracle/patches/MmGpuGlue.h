// MmGpuGlue.h -- reference-side binding of libmmgpu.so (include/mmgpu.h), added to the reference by oracle/patches/mmgpu.patch.
// Included only by the patched translation units (InterPrediction.cpp, InterSearch.cpp, MVReprojection.cpp, Picture.cpp):
// no reference header changes, so every other object file of the stock build is reused as is.
//
// MMGPU=0 in the environment switches the patched apps back to the reference's CPU code at run time.
#pragma once
#include "CommonDef.h"
#include "Unit.h"
#include "Mv.h"
#include "Buffer.h"
#include "Picture.h"
#include "RdCost.h"
#include "EpipoleList.h"

namespace mmglue
{
bool enabled();
bool mvpOnLibrary();    // MMGPU_MVP=0 keeps the MM-MVP conversion on the reference's code
bool classicOnGpu();   // MMGPU_CLASSIC=0 keeps the classic translational MC on the reference's code

// MVReprojection::init (MVReprojection.cpp:7-65): creates the device context (MPA caches, ERP tables on the device)
void init(const Size& resolution, const SPS* sps, const EpipoleList* epipoleList);
// Picture::extendPicBorder (Picture.cpp:988-1048): the reconstruction of `pic` changed -> its device copy is stale
void invalidate(const Picture* pic);
// make sure the (pic, wrap) variant is resident; returns its reference slot (lazy upload of the padded planes)
int ensureRef(const Picture* pic, bool wrap);

// body of InterPrediction::xPredInterBlkMM (InterPrediction.cpp:683-856)
void predInterBlkMM(ComponentID compID, ChromaFormat chFmt, const Position& blockPos, const Size& blockSize, const Picture* refPic,
                    bool wrapRef, const Mv& mv, MotionModelID motionModel, bool bi, const ClpRng& clpRng, int curPOC, PelBuf& dstBuf);

// plain case of InterPrediction::xPredInterBlk (InterPrediction.cpp:928-1096: classic translational MC of one component; no
// RPR / BDOF border / DMVR padding / IBC -- those have their own entry points below or stay on the reference's code)
void predInterBlkClassic(ComponentID compID, ChromaFormat chFmt, const Position& blockPos, int width, int height, const Picture* refPic,
                         bool wrapRef, const Mv& mv, bool bi, const ClpRng& clpRng, PelBuf& dstBuf);

// MVReprojection::motionVectorInDesiredMotionModel (MVReprojection.cpp:168-217): MM-MVP candidate conversion
Mv mvpConvert(const Position& position, const Mv& mvOrig, MotionModelID modelOrig, MotionModelID modelDesired, int shiftHor, int shiftVer,
              int curPOCOrig, int refPOCOrig, int curPOCDesired, int refPOCDesired, const Position& candPos, const Size& candSize,
              const Position& curPos, const Size& curSize);

// xPredInterBlk with bioApplied, luma: prediction + ring of integer samples into the BDOF work buffer (stride width + 4)
void predInterBlkClassicBdof(const Position& blockPos, int width, int height, const Picture* refPic, bool wrapRef, const Mv& mv,
                             const ClpRng& clpRng, Pel* bdofBuf);

// xPredInterBlk of a classic luma block with the alternative half-pel filter (pu.cu->imv == IMV_HPEL), no BDOF
void predInterBlkClassicAltHpel(const Position& blockPos, int width, int height, const Picture* refPic, bool wrapRef, const Mv& mv, bool bi,
                                const ClpRng& clpRng, PelBuf& dstBuf);

// ---- MMGPU_VERIFY=1: every hooked call also runs the reference's own code in the same process (the hook re-enters the
//      reference function with inVerify() set, so the second pass takes the CPU path) and the results are compared value by
//      value; the app stops at the first difference and the stats line reports how many values were compared.
bool verifying();
bool inVerify();
void setInVerify(bool on);
void verifySave(const Pel* buf, int stride, int w, int h);
void verifyCompare(const Pel* buf, int stride, int w, int h, const char* what);
void verifyMv(const Mv& viaLibrary, const Mv& viaReference);
void meVerify(const DistParam& dp);   // distortion of the last meEval() against the reference's distFunc on its own prediction

// ---- classic DMVR, InterPrediction::xProcessDMVR (InterPrediction.cpp:2239-2439).  MMGPU_DMVR=0 keeps it on the reference's code;
//      MMGPU_VERIFY=1 additionally runs the reference's own code next to the device and stops at the first difference.
bool dmvrOnGpu();
struct DmvrList
{
  const Picture* refPic;
  bool wrapRef;
  int orgX, orgY;     // xPrefetch's Rec_offset (luma): first prefetched sample
  int fracX, fracY;   // fraction xinitMC's xPredInterBlk filters with
};
// integer search of every sub-PU of a PU in one batch: Begin, Add per sub-PU (in the loop's order), Run; Select(num) makes
// sub-PU num's 25 costs current, Cost(k) = xDMVRCost at m_pSearchOffset[k], End forgets them (xBIPMVRefine asks Active()).
void dmvrBegin(int dx, int dy);
void dmvrAdd(const DmvrList& l0, const DmvrList& l1);
void dmvrRun();
void dmvrSelect(int num);
bool dmvrActive();
uint64_t dmvrCost(int k);
void dmvrVerifyCost(int k, uint64_t cpuCost);
void dmvrEnd();
// final MC of one sub-PU (xFinalPaddedMCForDMVR): Add per (list, component) that reads the padded window, Run at the end
struct DmvrFinal
{
  ComponentID compID;
  const Picture* refPic;
  bool wrapRef;
  int orgX, orgY;       // xPrefetch's Rec_offset of the component
  int width, height;    // component block
  int pad;              // xPad's padsize (0: sub-PU did not move, window not padded)
  int dIntX, dIntY;     // deltaIntMv
  int fracX, fracY;
  bool altHpel, bio;
  Pel* dst;             // pcYUVTemp.bufs[compID].buf, or the BDOF work buffer (bio)
  int dstStride;
};
void dmvrFinalAdd(const DmvrFinal& j);
void dmvrFinalRun();

// ---- motion estimation: xApplyMvVA / xMVReprojectionInterpolation + m_cDistParam.distFunc (InterSearch.cpp:363-453,
//      5287-5322, 6168-6385).  The search control flow is untouched: candidate sets the reference is about to visit are
//      evaluated ahead in ONE mmgpu_me_batch, and the reference's own compare chain then reads the distortions back one by
//      one through m_cDistParam.distFunc (identical decisions by construction).
struct MeBlock
{
  Position blkPos;
  Size blkSize;
  const Pel* refLuma00;   // cStruct.pcRefBuf->buf: sample (0,0) of the reference luma plane (identifies the resident slot)
  MotionModelID model;
  int curPOC, refPOC;
  ClpRng clpRng;
};
// after RdCost::setDistParam in the MM branches: remembers SAD/HAD, forgets the look-ahead of the previous search and
// points dp.distFunc at the function that returns the distortion of the last meEval()
void meSetup(DistParam& dp, bool useHadamard);
// evaluate one candidate (look-ahead hit, or a one-candidate batch); its distortion is what dp.distFunc returns next
void meEval(const MeBlock& b, const DistParam& dp, const Mv& mv, MvPrecision prec);
// classic fractional refinement (xPatternSearchFracDIF -> xPatternRefinement, InterSearch.cpp:739-853, 6093-6159: the
// reference up-samples the block once and compares 9 + 9 shifted views of it; the same predictions are the classic MC at the
// half / quarter-pel MVs): the caller announces the block and the integer base MV, xPatternRefinement asks per candidate.
void classicFracBegin(const MeBlock& b, const Mv& baseInt, bool usable);
bool classicFracActive();
void classicFracEnd();
const MeBlock& classicFracBlock();
Mv classicFracMv(const Mv& relQuarter);   // absolute quarter-pel MV of a candidate given relative to the integer base
// look-ahead: the listed MVs will (probably) be visited next
void mePrefetch(const MeBlock& b, const DistParam& dp, const Mv* mvs, int n, MvPrecision prec);
// dry run of one of the reference's own enumerators: while collecting, xTZSearchHelp only records the point
void beginCollect();
bool collecting();
void collect(int x, int y);
void endCollect(const MeBlock& b, const DistParam& dp);   // = mePrefetch of the collected integer MVs

void printStats();

// ---- debugging aid: MMGPU_DUMP=<file> (with MMGPU=0) records every xPredInterBlkMM call of the CPU path -- parameters,
//      the reference picture (un-padded planes, once per picture) and the predicted samples -- for tests/replay_mc_dump.py,
//      which replays the calls through the oracle (and, on a GPU box, through libmmgpu.so).
bool dumping();
// ME side of the dump (every MMGPU_DUMP_ME-th search; 0/unset = none): the original block once per search ('O'), then
// every candidate the CPU path evaluates with the distortion the reference's distFunc returned ('E').
void dumpRef(const Picture* refPic, bool wrapRef);
void meDumpSetup(DistParam& dp, bool useHadamard);
void meDumpCandidate(const MeBlock& b, const Mv& mv, MvPrecision prec);
void dumpMc(ComponentID compID, ChromaFormat chFmt, const Position& blockPos, const Size& blockSize, const Picture* refPic,
            bool wrapRef, const Mv& mv, MotionModelID motionModel, bool bi, const ClpRng& clpRng, int curPOC, const PelBuf& dstBuf);
}  // namespace mmglue
