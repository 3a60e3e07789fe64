// ref_wrap.cpp -- C entry points over the REFERENCE's own object code (TEST INFRASTRUCTURE ONLY).
//
// oracle/Makefile compiles, from where they lie under /root/reference/source/Lib/CommonLib, the untouched
//   InterpolationFilter.cpp, RdCost.cpp, x86/sse41/{InterpolationFilter,RdCost}_sse41.cpp,
//   x86/avx2/{InterpolationFilter,RdCost}_avx2.cpp
// against the parse-only Eigen stand-in (eigen_parse_shim/) and links them with this file into
// oracle/_ref/libvtmref.so.  tests/ use it to pin the oracle's integer steps (FIR, SAD, HAD, DMVR cost)
// against what the reference really executes, including the SSE4.1/AVX2 code paths.
// No reference source is copied: only declarations are used through the reference's headers.
#include <cstdlib>
#include <cstring>
#include "CommonLib/CommonDef.h"
#include "CommonLib/InterpolationFilter.h"
#include "CommonLib/RdCost.h"
#include "CommonLib/RdCostWeightPrediction.h"
#include "CommonLib/Rom.h"

// ---- symbols of translation units that are NOT part of the path (never reached from below) ---------------
int16_t** g_GeoParams = nullptr;
int16_t* g_globalGeoWeights[GEO_NUM_PRESTORED_MASK];
int16_t g_weightOffset[GEO_NUM_PARTITION_MODE][GEO_NUM_CU_SIZE][GEO_NUM_CU_SIZE][2];
int8_t g_angle2mask[GEO_NUM_ANGLES];
int8_t g_angle2mirror[GEO_NUM_ANGLES];
Distortion RdCostWeightPrediction::xGetSADw(const DistParam&) { abort(); }
Distortion RdCostWeightPrediction::xGetHADsw(const DistParam&) { abort(); }
Distortion RdCostWeightPrediction::xGetSSEw(const DistParam&) { abort(); }
template <> void AreaBuf<Pel>::subtract(const Pel) { abort(); }

// x86/InitX86.cpp is not compiled (it also wires ALF/affine/PelBuf SIMD); same dispatch for the two classes
// of the path: AVX2 when the host has it, else SSE4.1 (read_x86_extension_flags, x86/InitX86.cpp:57-123).
static int g_level = 0;  // 0 scalar, 1 sse41, 2 avx2
#if ENABLE_SIMD_OPT_MCIF
void InterpolationFilter::initInterpolationFilterX86()
{
  if (g_level == 2) _initInterpolationFilterX86<AVX2>();
  else if (g_level == 1) _initInterpolationFilterX86<SSE41>();
}
#endif
#if ENABLE_SIMD_OPT_DIST
void RdCost::initRdCostX86()
{
  if (g_level == 2) _initRdCostX86<AVX2>();
  else if (g_level == 1) _initRdCostX86<SSE41>();
}
#endif

static InterpolationFilter* g_if = nullptr;
static RdCost* g_rd = nullptr;

extern "C" int ref_host_level()
{
  __builtin_cpu_init();
  return __builtin_cpu_supports("avx2") ? 2 : (__builtin_cpu_supports("sse4.1") ? 1 : 0);
}

// level: 0 = scalar C++ (--SIMD=SCALAR), 1 = SSE4.1, 2 = AVX2, -1 = what the host would select
extern "C" void ref_init(int level)
{
  if (level < 0) level = ref_host_level();
  g_level = level;
  delete g_if;
  delete g_rd;
  g_if = new InterpolationFilter();      // constructor installs the scalar function pointers
  g_rd = new RdCost();                   // constructor -> init(): scalar m_afpDistortFunc
#if ENABLE_SIMD_OPT_MCIF
  g_if->initInterpolationFilterX86();
#endif
#if ENABLE_SIMD_OPT_DIST
  g_rd->initRdCostX86();
#endif
}

static ClpRng mk_clp(int bd, int lo, int hi)
{
  ClpRng c;
  c.min = lo; c.max = hi; c.bd = bd; c.n = 0;
  return c;
}

extern "C" void ref_filter_hor(int compID, const Pel* src, int srcStride, Pel* dst, int dstStride, int w, int h, int frac,
                               int isLast, int bd, int lo, int hi)
{
  g_if->filterHor(ComponentID(compID), src, srcStride, dst, dstStride, w, h, frac, isLast != 0, mk_clp(bd, lo, hi),
                  InterpolationFilter::FILTER_DEFAULT, false);
}
extern "C" void ref_filter_ver(int compID, const Pel* src, int srcStride, Pel* dst, int dstStride, int w, int h, int frac,
                               int isFirst, int isLast, int bd, int lo, int hi)
{
  g_if->filterVer(ComponentID(compID), src, srcStride, dst, dstStride, w, h, frac, isFirst != 0, isLast != 0,
                  mk_clp(bd, lo, hi), InterpolationFilter::FILTER_DEFAULT, false);
}

// the same two with the filter index / alternative half-pel flag exposed: dmvr != 0 -> FILTER_DMVR (bilinear, xinitMC),
// altHpel != 0 -> useAltHpelIf (pu.cu->imv == IMV_HPEL, InterPrediction.cpp:948)
extern "C" void ref_filter_hor_x(const Pel* src, int srcStride, Pel* dst, int dstStride, int w, int h, int frac, int isLast, int bd,
                                 int lo, int hi, int dmvr, int altHpel)
{
  g_if->filterHor(COMPONENT_Y, src, srcStride, dst, dstStride, w, h, frac, isLast != 0, mk_clp(bd, lo, hi),
                  dmvr ? InterpolationFilter::FILTER_DMVR : InterpolationFilter::FILTER_DEFAULT, altHpel != 0);
}
extern "C" void ref_filter_ver_x(const Pel* src, int srcStride, Pel* dst, int dstStride, int w, int h, int frac, int isFirst,
                                 int isLast, int bd, int lo, int hi, int dmvr, int altHpel)
{
  g_if->filterVer(COMPONENT_Y, src, srcStride, dst, dstStride, w, h, frac, isFirst != 0, isLast != 0, mk_clp(bd, lo, hi),
                  dmvr ? InterpolationFilter::FILTER_DMVR : InterpolationFilter::FILTER_DEFAULT, altHpel != 0);
}

// RdCost::setDistParam(rcDP, org, piRefY, iRefStride, bitDepth, compID, subShiftMode, step, useHadamard) + distFunc:
// what InterSearch does per ME candidate (InterSearch.cpp:6171, 6189, 6235; RdCost.cpp:225-311)
extern "C" unsigned long long ref_dist(const Pel* org, int orgStride, const Pel* cur, int curStride, int w, int h, int bd,
                                       int subShiftMode, int useHadamard)
{
  DistParam dp;
  CPelBuf o(org, orgStride, w, h);
  g_rd->setDistParam(dp, o, cur, curStride, bd, COMPONENT_Y, subShiftMode, 1, useHadamard != 0);
  return dp.distFunc(dp);
}
extern "C" int ref_dist_subshift(int w, int h, int subShiftMode)
{
  static Pel z[128 * 128];
  DistParam dp;
  CPelBuf o(z, 128, w, h);
  g_rd->setDistParam(dp, o, z, 128, 10, COMPONENT_Y, subShiftMode, 1, false);
  return dp.subShift;
}
// InterPrediction::xDMVRCost (InterPrediction.cpp:2147-2155)
extern "C" unsigned long long ref_dmvr_cost(const Pel* a, int as, const Pel* b, int bs, int w, int h, int bd)
{
  DistParam dp;
  dp.applyWeight = false;
  dp.useMR = false;
  g_rd->setDistParam(dp, a, b, as, bs, bd, COMPONENT_Y, w, h, 1);
  return dp.distFunc(dp) >> 1;
}
