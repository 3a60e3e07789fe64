// ref_wrap_dmvr.cpp -- C entry point over the REFERENCE's own DMVR error-surface code (TEST INFRASTRUCTURE ONLY).
// oracle/vtm.mk links it with the objects of the whole-reference build (CommonLib, Utilities, libmd5 -- compiled unchanged from
// /root/reference) into oracle/_ref/vtm/libvtmdmvr.so.  xDMVRSubPixelErrorSurface / xSubPelErrorSrfc / div_for_maxq7
// (InterPrediction.cpp:1958-2048, 2157-2175) are free functions of InterPrediction.cpp.
#include <cstdint>
void xDMVRSubPixelErrorSurface(bool notZeroCost, int16_t* totalDeltaMV, int16_t* deltaMV, uint64_t* pSADsArray);

// sads25: the 5x5 cost array (row-major, MAX_UINT64 = not evaluated), centre: index the search ended on; total: in/out (1/16 pel)
extern "C" void ref_dmvr_subpel(int notZeroCost, int16_t* total, int16_t* delta, uint64_t* sads25, int centre)
{
  xDMVRSubPixelErrorSurface(notZeroCost != 0, total, delta, sads25 + centre);
}

// ---- Buffer.cpp: PelBuf::addAvg (xWeightedAverage's default case, :551-587) and paddingCore (xPad, :372-396) ------------------
#include "CommonLib/Unit.h"   // brings Buffer.h in the order the reference needs
// simd != 0: the x86 function-pointer table the apps install (initPelBufOpsX86); 0: the scalar defaults of PelBufferOps()
extern "C" void ref_pelbuf_simd(int simd)
{
  g_pelBufOP = PelBufferOps();
#if ENABLE_SIMD_OPT_BUFFER && defined(TARGET_SIMD_X86)
  if (simd) g_pelBufOP.initPelBufOpsX86();
#endif
}
extern "C" void ref_add_avg(const Pel* s0, int st0, const Pel* s1, int st1, Pel* dst, int ds, int w, int h, int bd, int lo, int hi)
{
  ClpRng c;
  c.min = lo; c.max = hi; c.bd = bd; c.n = 0;
  PelBuf d(dst, ds, w, h);
  d.addAvg(CPelBuf(s0, st0, w, h), CPelBuf(s1, st1, w, h), c);
}
extern "C" void ref_padding(Pel* ptr, int stride, int w, int h, int padSize) { paddingCore(ptr, stride, w, h, padSize); }
