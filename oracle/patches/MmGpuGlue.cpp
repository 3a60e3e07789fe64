// MmGpuGlue.cpp -- see MmGpuGlue.h.  Everything here is host bookkeeping around the C ABI of libmmgpu.so:
// which reference picture lives in which device slot, which epipole in which rotation slot, and the look-ahead
// table of the motion search.  No sample or distortion is ever computed here.
#include "MmGpuGlue.h"
#include "Slice.h"
#include "mmgpu.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

namespace mmglue
{
namespace
{
struct RefEntry
{
  const Picture* pic;
  bool wrap;
  int poc;
  bool valid;
  uint64_t lastUse;
  const Pel* luma00;
};
struct RotEntry
{
  float e[3];
  bool used;
  uint64_t lastUse;
};
struct State
{
  mmgpu_ctx* ctx = nullptr;
  const EpipoleList* epipoles = nullptr;
  ChromaFormat chFmt = CHROMA_420;
  std::vector<RefEntry> refs;
  RotEntry rots[64];
  uint64_t tick = 0;
  // ME
  bool useHadamard = false;
  Distortion pending = 0;
  Distortion (*cpuDistFunc)(const DistParam&) = nullptr;
  bool pendingValid = false;
  std::unordered_map<uint64_t, Distortion> ahead;
  // signature of the search the look-ahead table belongs to
  const Pel* sigOrg = nullptr;
  const Pel* sigRef = nullptr;
  int sigX = -1, sigY = -1, sigW = 0, sigH = 0, sigModel = -1, sigRefPOC = 0, sigSubShift = 0;
  bool collecting = false;
  std::vector<Mv> collected;
  std::vector<mmgpu_mv> mvScratch;
  std::vector<uint64_t> distScratch;
  // statistics
  uint64_t nMcClassic = 0, nMvp = 0;
  uint64_t nMc = 0, nMeBatches = 0, nMeCands = 0, nMeHits = 0, nMeSingles = 0, nUploads = 0;
};
State g;

int envFlag()
{
  const char* e = getenv("MMGPU");
  return (e && e[0] == '0') ? 0 : 1;
}
const bool g_enabled = envFlag() != 0;

void chk(int rc, const char* what)
{
  if (rc != MMGPU_OK)
  {
    THROW(std::string("mmgpu: ") + what + ": " + mmgpu_last_error(g.ctx));
  }
}
uint64_t mvKey(int hor, int ver, int prec) { return ((uint64_t)(uint32_t)hor << 32) ^ ((uint64_t)(uint32_t)ver << 4) ^ (uint64_t)(prec & 15); }

int rotSlot(int curPOC, int refPOC)
{
  // GeodesicMotionModel::setEpipole(m_epipoleList->findEpipole(curPOC, refPOC)) (MVReprojection.cpp:127-129)
  const Array3TCoord e = g.epipoles->findEpipole(curPOC, refPOC);
  const float ef[3] = {(float)e.x(), (float)e.y(), (float)e.z()};
  int freeSlot = -1, oldest = 0;
  for (int s = 0; s < 64; ++s)
  {
    RotEntry& r = g.rots[s];
    if (r.used && memcmp(r.e, ef, sizeof(ef)) == 0) { r.lastUse = ++g.tick; return s; }
    if (!r.used && freeSlot < 0) freeSlot = s;
    if (r.lastUse < g.rots[oldest].lastUse) oldest = s;
  }
  const int s = freeSlot >= 0 ? freeSlot : oldest;
  float R[9];
  mmgpu_rodrigues(ef, R);
  chk(mmgpu_set_rotation(g.ctx, s, R), "set_rotation");
  memcpy(g.rots[s].e, ef, sizeof(ef));
  g.rots[s].used = true;
  g.rots[s].lastUse = ++g.tick;
  return s;
}

int slotOfLuma(const Pel* luma00)
{
  for (size_t s = 0; s < g.refs.size(); ++s)
    if (g.refs[s].valid && g.refs[s].luma00 == luma00) { g.refs[s].lastUse = ++g.tick; return (int)s; }
  THROW("mmglue: motion search on a reference picture that was not registered (ensureRef)");
}

struct AtExit { ~AtExit() { if (g.ctx) { printStats(); mmgpu_destroy(g.ctx); g.ctx = nullptr; } } } g_atExit;
}  // namespace

bool enabled() { return g_enabled; }
bool mvpOnLibrary()
{
  static const bool on = !(getenv("MMGPU_MVP") && getenv("MMGPU_MVP")[0] == '0');
  return on;
}
bool classicOnGpu()
{
  static const bool on = !(getenv("MMGPU_CLASSIC") && getenv("MMGPU_CLASSIC")[0] == '0');
  return on;
}

void init(const Size& resolution, const SPS* sps, const EpipoleList* epipoleList)
{
  g.epipoles = epipoleList;
  if (!g_enabled) return;
  if (g.ctx) { mmgpu_destroy(g.ctx); g.ctx = nullptr; }
  mmgpu_config c;
  memset(&c, 0, sizeof(c));
  c.pic_w = (int)resolution.width;
  c.pic_h = (int)resolution.height;
  c.chroma_fmt = (int)sps->getChromaFormatIdc();
  c.bit_depth = sps->getBitDepth(CHANNEL_TYPE_LUMA);
  c.max_cu = (int)sps->getMaxCUWidth();
  c.margin = MAX_SCALING_RATIO * ((int)sps->getMaxCUWidth() + PIC_MARGIN);   // Picture::create (Picture.cpp:91), EncLib.cpp:986
  c.offset4x4 = sps->getMMOffset4x4() == 4 ? 1.5f : (float)sps->getMMOffset4x4();
  c.ged_flavor = (int)sps->getGEDFlavor();
  c.max_refs = 2 * (MAX_NUM_REF_PICS + 1);
  c.device = getenv("MMGPU_DEVICE") ? atoi(getenv("MMGPU_DEVICE")) : 0;
  c.libm_fma = -1;
  c.had_mode = -1;
  if (mmgpu_create(&c, &g.ctx) != MMGPU_OK) { THROW(std::string("mmgpu_create: ") + mmgpu_last_error(nullptr)); }
  g.epipoles = epipoleList;
  g.chFmt = sps->getChromaFormatIdc();
  g.refs.assign((size_t)c.max_refs, RefEntry{nullptr, false, 0, false, 0, nullptr});
  memset(g.rots, 0, sizeof(g.rots));
  fprintf(stderr, "[mmgpu] context %dx%d fmt %d depth %d maxCU %d margin %d\n", c.pic_w, c.pic_h, c.chroma_fmt, c.bit_depth, c.max_cu, c.margin);
}

void invalidate(const Picture* pic)
{
  if (!g_enabled) return;
  for (auto& r : g.refs)
    if (r.pic == pic) r.valid = false;
}

int ensureRef(const Picture* pic, bool wrap)
{
  CHECK(!g.ctx, "mmglue: MVReprojection::init has not run");
  int freeSlot = -1, oldest = 0;
  for (size_t s = 0; s < g.refs.size(); ++s)
  {
    RefEntry& r = g.refs[s];
    if (r.pic == pic && r.wrap == wrap)
    {
      if (r.valid && r.poc == pic->getPOC()) { r.lastUse = ++g.tick; return (int)s; }
      freeSlot = (int)s;   // stale copy of this very picture: refresh in place
      break;
    }
    if (!r.valid && freeSlot < 0) freeSlot = (int)s;
    if (r.lastUse < g.refs[oldest].lastUse) oldest = (int)s;
  }
  const int s = freeSlot >= 0 ? freeSlot : oldest;
  CHECK((int)pic->margin != MAX_SCALING_RATIO * ((int)pic->cs->sps->getMaxCUWidth() + PIC_MARGIN), "mmglue: unexpected picture margin");
  const int16_t* planes[3] = {nullptr, nullptr, nullptr};
  int32_t strides[3] = {0, 0, 0};
  const int nc = getNumberValidComponents(pic->chromaFormat);
  for (int c = 0; c < nc; ++c)
  {
    const CPelBuf b = pic->getRecoBuf(ComponentID(c), wrap);
    planes[c] = b.buf;
    strides[c] = (int32_t)b.stride;
  }
  chk(mmgpu_set_ref(g.ctx, s, pic->getPOC(), planes, strides), "set_ref");
  g.refs[s] = RefEntry{pic, wrap, pic->getPOC(), true, ++g.tick, planes[0]};
  g.nUploads++;
  return s;
}

void predInterBlkMM(ComponentID compID, ChromaFormat chFmt, const Position& blockPos, const Size& blockSize, const Picture* refPic,
                    bool wrapRef, const Mv& mv, MotionModelID motionModel, bool bi, const ClpRng& clpRng, int curPOC, PelBuf& dstBuf)
{
  const int slot = ensureRef(refPic, wrapRef);
  const int rot = (motionModel == GEODESIC_CAMPOSE) ? rotSlot(curPOC, refPic->getPOC()) : 0;
  const mmgpu_mv m = {mv.hor, mv.ver};
  chk(mmgpu_mc(g.ctx, slot, (int)compID, (int)chFmt, blockPos.x, blockPos.y, (int)blockSize.width, (int)blockSize.height, m,
               (int)motionModel, rot, bi ? 1 : 0, clpRng.min, clpRng.max, MMGPU_OOB_PRED, dstBuf.buf, (int)dstBuf.stride), "mc");
  g.nMc++;
}

void predInterBlkClassic(ComponentID compID, ChromaFormat chFmt, const Position& blockPos, int width, int height, const Picture* refPic,
                         bool wrapRef, const Mv& mv, bool bi, const ClpRng& clpRng, PelBuf& dstBuf)
{
  const int slot = ensureRef(refPic, wrapRef);
  const mmgpu_mv m = {mv.hor, mv.ver};
  chk(mmgpu_mc(g.ctx, slot, (int)compID, (int)chFmt, blockPos.x, blockPos.y, width, height, m, MMGPU_CLASSIC, 0, bi ? 1 : 0, clpRng.min,
               clpRng.max, MMGPU_OOB_PRED, dstBuf.buf, (int)dstBuf.stride), "mc (classic)");
  g.nMcClassic++;
}

Mv mvpConvert(const Position& position, const Mv& mvOrig, MotionModelID modelOrig, MotionModelID modelDesired, int shiftHor, int shiftVer,
              int curPOCOrig, int refPOCOrig, int curPOCDesired, int refPOCDesired, const Position& candPos, const Size& candSize,
              const Position& curPos, const Size& curSize)
{
  const int ro = (modelOrig == GEODESIC_CAMPOSE) ? rotSlot(curPOCOrig, refPOCOrig) : 0;
  const int rd = (modelDesired == GEODESIC_CAMPOSE) ? rotSlot(curPOCDesired, refPOCDesired) : 0;
  const mmgpu_mv in = {mvOrig.hor, mvOrig.ver};
  mmgpu_mv out = {0, 0};
  chk(mmgpu_mvp_convert(g.ctx, position.x, position.y, in, (int)modelOrig, (int)modelDesired, shiftHor, shiftVer, ro, rd, candPos.x, candPos.y,
                        (int)candSize.width, (int)candSize.height, curPos.x, curPos.y, (int)curSize.width, (int)curSize.height, &out),
      "mvp_convert");
  g.nMvp++;
  return Mv(out.hor, out.ver);
}

void predInterBlkClassicBdof(const Position& blockPos, int width, int height, const Picture* refPic, bool wrapRef, const Mv& mv,
                             const ClpRng& clpRng, Pel* bdofBuf)
{
  const int slot = ensureRef(refPic, wrapRef);
  const mmgpu_mv m = {mv.hor, mv.ver};
  chk(mmgpu_mc_bdof(g.ctx, slot, blockPos.x, blockPos.y, width, height, m, clpRng.min, clpRng.max, bdofBuf), "mc_bdof");
  g.nMcClassic++;
}

void predInterBlkClassicAltHpel(const Position& blockPos, int width, int height, const Picture* refPic, bool wrapRef, const Mv& mv, bool bi,
                                const ClpRng& clpRng, PelBuf& dstBuf)
{
  const int slot = ensureRef(refPic, wrapRef);
  const mmgpu_mv m = {mv.hor, mv.ver};
  chk(mmgpu_mc_alt_hpel(g.ctx, slot, blockPos.x, blockPos.y, width, height, m, bi ? 1 : 0, clpRng.min, clpRng.max, dstBuf.buf, (int)dstBuf.stride),
      "mc_alt_hpel");
  g.nMcClassic++;
}

// ---- classic DMVR -----------------------------------------------------------------------------------------------------
namespace
{
std::vector<mmgpu_dmvr_classic_subpu> g_dmvrSubs;
std::vector<uint64_t> g_dmvrCosts;
int g_dmvrDx = 0, g_dmvrDy = 0, g_dmvrCur = -1;
std::vector<mmgpu_dmvr_mc_job> g_dmvrJobs;
std::vector<DmvrFinal> g_dmvrFinals;
std::vector<int16_t> g_dmvrOut;
uint64_t g_nDmvrSubPus = 0, g_nDmvrFinal = 0, g_nDmvrVerified = 0;
}
bool dmvrOnGpu()
{
  static const bool on = !(getenv("MMGPU_DMVR") && getenv("MMGPU_DMVR")[0] == '0');
  return on;
}
bool verifying()
{
  static const bool on = getenv("MMGPU_VERIFY") && getenv("MMGPU_VERIFY")[0] == '1';
  return on;
}
namespace
{
bool g_inVerify = false;
std::vector<Pel> g_verifySaved;
uint64_t g_nVerMcCalls = 0, g_nVerMcSamples = 0, g_nVerMv = 0, g_nVerDist = 0;
}
bool inVerify() { return g_inVerify; }
void setInVerify(bool on) { g_inVerify = on; }
void verifySave(const Pel* buf, int stride, int w, int h)
{
  g_verifySaved.resize((size_t)w * h);
  for (int y = 0; y < h; ++y) memcpy(&g_verifySaved[(size_t)y * w], buf + (size_t)y * stride, sizeof(Pel) * w);
}
void verifyCompare(const Pel* buf, int stride, int w, int h, const char* what)
{
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if (buf[(size_t)y * stride + x] != g_verifySaved[(size_t)y * w + x])
      {
        fprintf(stderr, "[mmgpu] VERIFY: %s %dx%d at (%d,%d): device %d, reference %d\n", what, w, h, x, y, (int)g_verifySaved[(size_t)y * w + x],
                (int)buf[(size_t)y * stride + x]);
        THROW("mmglue: MMGPU_VERIFY found a prediction sample that differs from the reference's own");
      }
  g_nVerMcCalls++;
  g_nVerMcSamples += (uint64_t)w * h;
}
void verifyMv(const Mv& a, const Mv& b)
{
  if (a.hor != b.hor || a.ver != b.ver)
  {
    fprintf(stderr, "[mmgpu] VERIFY: MM-MVP conversion: library (%d,%d), reference (%d,%d)\n", a.hor, a.ver, b.hor, b.ver);
    THROW("mmglue: MMGPU_VERIFY found an MM-MVP conversion that differs from the reference's own");
  }
  g_nVerMv++;
}
void meVerify(const DistParam& dp)
{
  CHECK(!g.pendingValid || !g.cpuDistFunc, "mmglue: meVerify without a pending distortion");
  DistParam c = dp;
  c.distFunc = g.cpuDistFunc;
  c.maximumDistortionForEarlyExit = std::numeric_limits<Distortion>::max();   // the device always returns the complete sum
  const Distortion ref = g.cpuDistFunc(c);
  if (ref != g.pending)
  {
    fprintf(stderr, "[mmgpu] VERIFY: candidate distortion (%s, %dx%d): device %llu, reference %llu\n", g.useHadamard ? "HAD" : "SAD", (int)dp.org.width,
            (int)dp.org.height, (unsigned long long)g.pending, (unsigned long long)ref);
    THROW("mmglue: MMGPU_VERIFY found a candidate distortion that differs from the reference's own");
  }
  g_nVerDist++;
}

void dmvrBegin(int dx, int dy)
{
  g_dmvrSubs.clear();
  g_dmvrDx = dx; g_dmvrDy = dy; g_dmvrCur = -1;
}
void dmvrAdd(const DmvrList& l0, const DmvrList& l1)
{
  mmgpu_dmvr_classic_subpu s;
  s.w = g_dmvrDx; s.h = g_dmvrDy;
  s.ref_slot0 = ensureRef(l0.refPic, l0.wrapRef);
  s.ref_slot1 = ensureRef(l1.refPic, l1.wrapRef);
  s.org_x0 = l0.orgX; s.org_y0 = l0.orgY; s.frac_x0 = l0.fracX; s.frac_y0 = l0.fracY;
  s.org_x1 = l1.orgX; s.org_y1 = l1.orgY; s.frac_x1 = l1.fracX; s.frac_y1 = l1.fracY;
  g_dmvrSubs.push_back(s);
}
void dmvrRun()
{
  g_dmvrCosts.resize(g_dmvrSubs.size() * 25);
  chk(mmgpu_dmvr_classic_costs(g.ctx, (int)g_dmvrSubs.size(), g_dmvrSubs.data(), g_dmvrCosts.data()), "dmvr_classic_costs");
  g_nDmvrSubPus += g_dmvrSubs.size();
}
void dmvrSelect(int num)
{
  CHECK(num < 0 || num >= (int)g_dmvrSubs.size(), "mmglue: DMVR sub-PU index outside the batch");
  g_dmvrCur = num;
}
bool dmvrActive() { return g_dmvrCur >= 0; }
uint64_t dmvrCost(int k) { return g_dmvrCosts[(size_t)g_dmvrCur * 25 + k]; }
void dmvrVerifyCost(int k, uint64_t cpuCost)
{
  if (dmvrCost(k) != cpuCost)
  {
    fprintf(stderr, "[mmgpu] VERIFY: DMVR cost of sub-PU %d, offset %d: device %llu, reference %llu\n", g_dmvrCur, k,
            (unsigned long long)dmvrCost(k), (unsigned long long)cpuCost);
    THROW("mmglue: MMGPU_VERIFY found a classic-DMVR cost that differs from the reference's own");
  }
  g_nDmvrVerified++;
}
void dmvrEnd() { g_dmvrCur = -1; }

void dmvrFinalAdd(const DmvrFinal& f)
{
  mmgpu_dmvr_mc_job j;
  j.comp = (int)f.compID;
  j.ref_slot = ensureRef(f.refPic, f.wrapRef);
  j.org_x = f.orgX; j.org_y = f.orgY; j.w = f.width; j.h = f.height; j.pad = f.pad; j.dint_x = f.dIntX; j.dint_y = f.dIntY;
  j.frac_x = f.fracX; j.frac_y = f.fracY; j.alt_hpel = f.altHpel ? 1 : 0; j.bio = f.bio ? 1 : 0;
  int off = 0;
  if (!g_dmvrJobs.empty())
  {
    const mmgpu_dmvr_mc_job& p = g_dmvrJobs.back();
    off = p.dst_offset + (p.bio ? (p.w + 4) * (p.h + 4) : p.w * p.h);
  }
  j.dst_offset = off;
  j.dst_stride = f.width;
  g_dmvrJobs.push_back(j);
  g_dmvrFinals.push_back(f);
}
void dmvrFinalRun()
{
  if (g_dmvrJobs.empty()) return;
  const mmgpu_dmvr_mc_job& l = g_dmvrJobs.back();
  const int total = l.dst_offset + (l.bio ? (l.w + 4) * (l.h + 4) : l.w * l.h);
  g_dmvrOut.resize(total);
  chk(mmgpu_dmvr_final_mc(g.ctx, (int)g_dmvrJobs.size(), g_dmvrJobs.data(), g_dmvrOut.data(), total), "dmvr_final_mc");
  for (size_t i = 0; i < g_dmvrJobs.size(); ++i)
  {
    const mmgpu_dmvr_mc_job& j = g_dmvrJobs[i];
    const DmvrFinal& f = g_dmvrFinals[i];
    const int16_t* src = g_dmvrOut.data() + j.dst_offset;
    // bio: only the ring box (rows 1 .. h+2, columns 1 .. w+2) of the work buffer is defined by the reference
    const int sStride = j.bio ? j.w + 4 : j.w, rows = j.bio ? j.h + 2 : j.h, cols = j.bio ? j.w + 2 : j.w, first = j.bio ? sStride + 1 : 0;
    const int dStride = j.bio ? j.w + 4 : f.dstStride;
    for (int y = 0; y < rows; ++y)
    {
      Pel* d = f.dst + first + y * dStride;
      const int16_t* s = src + first + y * sStride;
      if (verifying())
      {
        for (int x = 0; x < cols; ++x)
          if (d[x] != s[x])
          {
            fprintf(stderr, "[mmgpu] VERIFY: DMVR final MC comp %d %dx%d pad %d dint (%d,%d) frac (%d,%d) alt %d bio %d at (%d,%d): device %d, reference %d\n",
                    j.comp, j.w, j.h, j.pad, j.dint_x, j.dint_y, j.frac_x, j.frac_y, j.alt_hpel, j.bio, x, y, (int)s[x], (int)d[x]);
            THROW("mmglue: MMGPU_VERIFY found a classic-DMVR final prediction that differs from the reference's own");
          }
        g_nDmvrVerified++;
      }
      else memcpy(d, s, sizeof(Pel) * cols);
    }
  }
  g_nDmvrFinal += g_dmvrJobs.size();
  g_dmvrJobs.clear();
  g_dmvrFinals.clear();
}

// ---- ME -----------------------------------------------------------------------------------------------------------------
static Distortion pendingDist(const DistParam&)
{
  CHECK(!g.pendingValid, "mmglue: distFunc called without a preceding candidate evaluation");
  g.pendingValid = false;
  return g.pending;
}

void meSetup(DistParam& dp, bool useHadamard)
{
  CHECK(dp.applyWeight || dp.useMR, "mmglue: weighted / mean-removed distortion is not on the GPU path");
  g.useHadamard = useHadamard;
  g.ahead.clear();
  g.sigOrg = nullptr;
  g.pendingValid = false;
  if (dp.distFunc != pendingDist) g.cpuDistFunc = dp.distFunc;   // what RdCost::setDistParam selected (MMGPU_VERIFY)
  dp.distFunc = pendingDist;
}

static void checkSignature(const MeBlock& b, const DistParam& dp)
{
  if (g.sigOrg == dp.org.buf && g.sigRef == b.refLuma00 && g.sigX == b.blkPos.x && g.sigY == b.blkPos.y && g.sigW == (int)b.blkSize.width &&
      g.sigH == (int)b.blkSize.height && g.sigModel == (int)b.model && g.sigRefPOC == b.refPOC && g.sigSubShift == dp.subShift)
    return;
  g.ahead.clear();
  g.sigOrg = dp.org.buf; g.sigRef = b.refLuma00; g.sigX = b.blkPos.x; g.sigY = b.blkPos.y; g.sigW = (int)b.blkSize.width;
  g.sigH = (int)b.blkSize.height; g.sigModel = (int)b.model; g.sigRefPOC = b.refPOC; g.sigSubShift = dp.subShift;
}

static void evalBatch(const MeBlock& b, const DistParam& dp, const Mv* mvs, int n, MvPrecision prec)
{
  CHECK((int)dp.org.width != (int)b.blkSize.width || (int)dp.org.height != (int)b.blkSize.height, "mmglue: pattern size != block size");
  mmgpu_block blk;
  memset(&blk, 0, sizeof(blk));
  blk.x = b.blkPos.x; blk.y = b.blkPos.y; blk.w = (int)b.blkSize.width; blk.h = (int)b.blkSize.height;
  blk.model = (int)b.model;
  blk.ref_slot = slotOfLuma(b.refLuma00);
  blk.rot_slot = (b.model == GEODESIC_CAMPOSE) ? rotSlot(b.curPOC, b.refPOC) : 0;
  blk.cand_first = 0;
  blk.cand_count = n;
  g.mvScratch.resize((size_t)n);
  g.distScratch.resize((size_t)n);
  for (int i = 0; i < n; ++i) { g.mvScratch[i].hor = mvs[i].hor; g.mvScratch[i].ver = mvs[i].ver; }
  chk(mmgpu_set_org(g.ctx, dp.org.buf, (int)dp.org.stride, (int)dp.org.width, (int)dp.org.height), "set_org");
  chk(mmgpu_me_batch(g.ctx, 1, &blk, n, g.mvScratch.data(), (int)MV_PRECISION_INTERNAL - (int)prec,
                     g.useHadamard ? MMGPU_DIST_HAD : MMGPU_DIST_SAD, dp.subShift, b.clpRng.min, b.clpRng.max, g.distScratch.data()),
      "me_batch");
  g.nMeBatches++;
  g.nMeCands += (uint64_t)n;
}

void mePrefetch(const MeBlock& b, const DistParam& dp, const Mv* mvs, int n, MvPrecision prec)
{
  checkSignature(b, dp);
  std::vector<Mv> todo;
  todo.reserve((size_t)n);
  for (int i = 0; i < n; ++i)
  {
    const uint64_t k = mvKey(mvs[i].hor, mvs[i].ver, (int)prec);
    if (g.ahead.find(k) == g.ahead.end()) { g.ahead[k] = std::numeric_limits<Distortion>::max(); todo.push_back(mvs[i]); }
  }
  if (todo.empty()) return;
  evalBatch(b, dp, todo.data(), (int)todo.size(), prec);
  for (size_t i = 0; i < todo.size(); ++i) g.ahead[mvKey(todo[i].hor, todo[i].ver, (int)prec)] = g.distScratch[i];
}

void meEval(const MeBlock& b, const DistParam& dp, const Mv& mv, MvPrecision prec)
{
  checkSignature(b, dp);
  const uint64_t k = mvKey(mv.hor, mv.ver, (int)prec);
  auto it = g.ahead.find(k);
  if (it != g.ahead.end())
  {
    g.pending = it->second;
    g.nMeHits++;
  }
  else
  {
    evalBatch(b, dp, &mv, 1, prec);
    g.pending = g.distScratch[0];
    g.ahead[k] = g.pending;
    g.nMeSingles++;
  }
  g.pendingValid = true;
}

namespace
{
MeBlock g_cfBlk;
Mv g_cfBase;
bool g_cfActive = false;
}
void classicFracBegin(const MeBlock& b, const Mv& baseInt, bool usable)
{
  g_cfBlk = b;
  g_cfBase = baseInt;
  g_cfActive = usable;
}
bool classicFracActive() { return g_cfActive; }
void classicFracEnd() { g_cfActive = false; }
const MeBlock& classicFracBlock() { return g_cfBlk; }
Mv classicFracMv(const Mv& relQuarter) { return Mv(g_cfBase.hor * 4 + relQuarter.hor, g_cfBase.ver * 4 + relQuarter.ver); }

void beginCollect()
{
  CHECK(g.collecting, "mmglue: nested dry run");
  g.collecting = true;
  g.collected.clear();
}
bool collecting() { return g.collecting; }
void collect(int x, int y) { g.collected.push_back(Mv(x, y)); }
void endCollect(const MeBlock& b, const DistParam& dp)
{
  g.collecting = false;
  if (!g.collected.empty()) mePrefetch(b, dp, g.collected.data(), (int)g.collected.size(), MV_PRECISION_INT);
}

// ---- dump of the CPU path's MC calls ---------------------------------------------------------------------------
namespace
{
FILE* g_dump = nullptr;
bool g_dumpInit = false;
std::vector<std::pair<const Picture*, int>> g_dumped;   // (picture, POC | wrap << 30) already written
std::vector<std::pair<const Pel*, int>> g_dumpedLuma;   // luma (0,0) pointer -> POC | wrap << 30
void w32(int32_t v) { fwrite(&v, 4, 1, g_dump); }
// ME dump state
int g_meEvery = -1;
long g_meSearches = 0;
bool g_meActive = false, g_meHeaderDone = false, g_meHad = false;
FpDistFunc g_meOrigDist = nullptr;
MeBlock g_meBlk;
Mv g_meMv;
int g_mePrec = 0;
bool g_meCandValid = false;
}
bool dumping()
{
  if (!g_dumpInit)
  {
    g_dumpInit = true;
    const char* f = getenv("MMGPU_DUMP");
    if (f && f[0]) g_dump = fopen(f, "wb");
  }
  return g_dump != nullptr;
}
void dumpRef(const Picture* refPic, bool wrapRef)
{
  const int key = refPic->getPOC() | (wrapRef ? (1 << 30) : 0);
  bool seen = false;
  for (auto& d : g_dumped) seen = seen || (d.first == refPic && d.second == key);
  if (!seen)
  {
    // a picture object is recycled for later frames: forget its older entries
    for (auto& d : g_dumped) if (d.first == refPic) d.second = -1;
    g_dumped.push_back({refPic, key});
    const Pel* luma00 = refPic->getRecoBuf(COMPONENT_Y, wrapRef).buf;
    for (auto& d : g_dumpedLuma) if (d.first == luma00) d.second = -1;
    g_dumpedLuma.push_back({luma00, key});
    w32('R'); w32(refPic->getPOC()); w32(wrapRef ? 1 : 0);
    const int nc = getNumberValidComponents(refPic->chromaFormat);
    w32(nc);
    for (int c = 0; c < nc; ++c)
    {
      const CPelBuf b = refPic->getRecoBuf(ComponentID(c), wrapRef);
      w32((int)b.width); w32((int)b.height);
      for (int y = 0; y < (int)b.height; ++y) fwrite(b.buf + (ptrdiff_t)y * b.stride, sizeof(Pel), b.width, g_dump);
    }
  }
}

static Distortion dumpingDist(const DistParam& dp)
{
  const Distortion d = g_meOrigDist(dp);
  if (g_meActive && g_meCandValid)
  {
    if (!g_meHeaderDone)
    {
      int key = -1;
      for (auto& r : g_dumpedLuma) if (r.first == g_meBlk.refLuma00 && r.second >= 0) key = r.second;
      CHECK(key < 0, "mmglue dump: search on an unregistered reference picture");
      float e[3] = {0, 0, 0};
      if (g_meBlk.model == GEODESIC_CAMPOSE && g.epipoles)
      {
        const Array3TCoord ep = g.epipoles->findEpipole(g_meBlk.curPOC, g_meBlk.refPOC);
        e[0] = (float)ep.x(); e[1] = (float)ep.y(); e[2] = (float)ep.z();
      }
      w32('O'); w32(g_meBlk.blkPos.x); w32(g_meBlk.blkPos.y); w32((int)g_meBlk.blkSize.width); w32((int)g_meBlk.blkSize.height);
      w32(key & 0x3fffffff); w32(key >> 30); w32((int)g_meBlk.model); w32(g_meBlk.curPOC); fwrite(e, 4, 3, g_dump);
      w32(g_meHad ? 1 : 0); w32(dp.subShift); w32(g_meBlk.clpRng.min); w32(g_meBlk.clpRng.max);
      for (int y = 0; y < (int)dp.org.height; ++y) fwrite(dp.org.buf + (ptrdiff_t)y * dp.org.stride, sizeof(Pel), dp.org.width, g_dump);
      g_meHeaderDone = true;
    }
    w32('E'); w32(g_meMv.hor); w32(g_meMv.ver); w32((int)MV_PRECISION_INTERNAL - g_mePrec);
    const uint64_t dd = d;
    fwrite(&dd, 8, 1, g_dump);
  }
  g_meCandValid = false;
  return d;
}
void meDumpSetup(DistParam& dp, bool useHadamard)
{
  if (g_meEvery < 0) { const char* e = getenv("MMGPU_DUMP_ME"); g_meEvery = e ? atoi(e) : 0; }
  if (g_meEvery <= 0) return;
  g_meActive = (g_meSearches++ % g_meEvery) == 0;
  g_meHeaderDone = false;
  g_meCandValid = false;
  g_meHad = useHadamard;
  g_meOrigDist = dp.distFunc;
  dp.distFunc = dumpingDist;
}
void meDumpCandidate(const MeBlock& b, const Mv& mv, MvPrecision prec)
{
  g_meBlk = b; g_meMv = mv; g_mePrec = (int)prec; g_meCandValid = true;
}

void dumpMc(ComponentID compID, ChromaFormat chFmt, const Position& blockPos, const Size& blockSize, const Picture* refPic,
            bool wrapRef, const Mv& mv, MotionModelID motionModel, bool bi, const ClpRng& clpRng, int curPOC, const PelBuf& dstBuf)
{
  dumpRef(refPic, wrapRef);
  float e[3] = {0, 0, 0};
  if (motionModel == GEODESIC_CAMPOSE && g.epipoles)
  {
    const Array3TCoord ep = g.epipoles->findEpipole(curPOC, refPic->getPOC());
    e[0] = (float)ep.x(); e[1] = (float)ep.y(); e[2] = (float)ep.z();
  }
  w32('M'); w32((int)compID); w32((int)chFmt); w32(blockPos.x); w32(blockPos.y); w32((int)blockSize.width); w32((int)blockSize.height);
  w32(refPic->getPOC()); w32(wrapRef ? 1 : 0); w32(mv.hor); w32(mv.ver); w32((int)motionModel); w32(bi ? 1 : 0); w32(clpRng.min);
  w32(clpRng.max); w32(curPOC);
  fwrite(e, 4, 3, g_dump);
  for (int y = 0; y < (int)blockSize.height; ++y) fwrite(dstBuf.buf + (ptrdiff_t)y * dstBuf.stride, sizeof(Pel), blockSize.width, g_dump);
  fflush(g_dump);
}

void printStats()
{
  fprintf(stderr, "[mmgpu] classic MC calls %llu | MM-MVP conversions %llu\n", (unsigned long long)g.nMcClassic, (unsigned long long)g.nMvp);
  fprintf(stderr, "[mmgpu] classic DMVR: %llu sub-PU searches, %llu final-MC jobs%s\n", (unsigned long long)g_nDmvrSubPus, (unsigned long long)g_nDmvrFinal,
          verifying() ? (" | VERIFY: " + std::to_string(g_nDmvrVerified) + " comparisons against the reference's own code, all equal").c_str() : "");
  if (verifying())
    fprintf(stderr, "[mmgpu] VERIFY (device vs the reference's own code in this process, all equal): %llu MC calls / %llu samples, %llu candidate distortions, "
                    "%llu MM-MVP conversions\n", (unsigned long long)g_nVerMcCalls, (unsigned long long)g_nVerMcSamples, (unsigned long long)g_nVerDist,
            (unsigned long long)g_nVerMv);
  fprintf(stderr, "[mmgpu] MC calls %llu | ME batches %llu, candidates %llu (%.1f per batch), look-ahead hits %llu, single evaluations %llu | "
                  "reference uploads %llu | kernels launched %llu\n",
          (unsigned long long)g.nMc, (unsigned long long)g.nMeBatches, (unsigned long long)g.nMeCands,
          g.nMeBatches ? (double)g.nMeCands / (double)g.nMeBatches : 0.0, (unsigned long long)g.nMeHits, (unsigned long long)g.nMeSingles,
          (unsigned long long)g.nUploads, (unsigned long long)(g.ctx ? mmgpu_launch_count(g.ctx) : 0));
}
}  // namespace mmglue
