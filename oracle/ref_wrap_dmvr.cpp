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
