/* mm_oracle.h -- CPU oracle for the multi-model 360 MC + distortion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or executed by the
 * product (vvc-extension-mm_b200/, include/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the checker/baseline.
 *
 * It is a plain-C restatement of the reference's algorithm (FAU-LMS/vvc-extension-mm,
 * VTM-17.2 + MM); every function cites the reference file:line it follows.
 *
 * PARITY STATUS (details: DESIGN.md section 2)
 *   integer steps (FIR, clip, OOB zero fill, SAD, HAD, DMVR cost, weighted average):
 *       PINNED against the reference's own InterpolationFilter.cpp / RdCost.cpp + SSE4.1/AVX2 units compiled
 *       from /root/reference into oracle/_ref/libvtmref.so (tests/test_oracle_vs_reference_objects.py).
 *   float mapping, operation order / model algebra / caching / quirks:
 *       PINNED against the reference's own MVReprojection.cpp, MotionModels/<model>.cpp, Projection.cpp, Coordinate.cpp
 *       compiled unchanged into oracle/_ref/libvtmmap.so (tests/test_oracle_vs_reference_mapping.py).
 *   libm: the oracle calls the REAL glibc of the box (sinf cosf tanf atanf atan2f acosf asinf roundf sin cos).
 *   Eigen 3.3.7 itself is not on the box (fetched from the network by the reference's cmake,
 *       source/3rdparty/External-Eigen3.cmake:4-6) and the reference ships no tests or golden vectors:
 *       libvtmmap.so is built against a functional stand-in (oracle/eigen_func_shim).  What that stand-in and
 *       this file assert about Eigen -- SSE psin/pcos packets, packet-body/scalar-tail traversal, 3-term dot
 *       association, scalar min/max, round -- is "PARITY UNPINNED": restated from Eigen 3.3.7 as recalled.
 */
#ifndef MM_ORACLE_H
#define MM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* MotionModelID, source/Lib/CommonLib/TypeDef.h:865-880 */
enum {
  ORC_CLASSIC = 0, ORC_MPA_FRONT_BACK = 1, ORC_MPA_LEFT_RIGHT = 2, ORC_MPA_TOP_BOTTOM = 3,
  ORC_TANGENTIAL = 4, ORC_THREE_D_TRANSLATIONAL = 5, ORC_ROTATIONAL = 6,
  ORC_GEODESIC_X = 7, ORC_GEODESIC_Y = 8, ORC_GEODESIC_Z = 9, ORC_GEODESIC_CAMPOSE = 10, ORC_NUM_MODELS = 11
};
enum { ORC_CHROMA_400 = 0, ORC_CHROMA_420 = 1, ORC_CHROMA_422 = 2, ORC_CHROMA_444 = 3 };
enum { ORC_GED_ORIGINAL = 0, ORC_GED_MODULATED = 1 };
enum { ORC_OOB_PRED = 0, ORC_OOB_ME = 1 };       /* InterPrediction.cpp:776-780 vs InterSearch.cpp:6329-6332 */
enum { ORC_DIST_SAD = 0, ORC_DIST_HAD = 1 };
enum { ORC_HAD_SCALAR = 0, ORC_HAD_SIMD = 1 };   /* RdCost.cpp:2637 vs x86/RdCostX86.h:2661 */

typedef struct OrcCtx OrcCtx;

typedef struct {
  int pic_w, pic_h;      /* luma picture size */
  int chroma_fmt;        /* ORC_CHROMA_* of the sequence */
  int max_cu;            /* sps->getMaxCUWidth(), 128 in the VTM cfgs */
  float offset4x4;       /* MVReprojection::m_offset4x4 (1.0: EncApp.cpp:758) */
  int ged_flavor;        /* ORC_GED_MODULATED (EncApp.cpp:755) */
  int bit_depth;         /* internal bit depth (10) */
} OrcCfg;

OrcCtx* orc_create(const OrcCfg* cfg);
void orc_destroy(OrcCtx* c);

/* GeodesicMotionModel::setEpipole, GeodesicMotionModel.cpp:14-46: float epipole -> row-major R[9]. */
void orc_rodrigues(const float epipole[3], float R[9]);
/* FloatingFixedConversion::fixedToFloating, Coordinate.cpp:75-77 (Q24 epipole -> float). */
float orc_fixed_to_float(int fixed, int precision);
/* Fixed GEDA axes / camera pose rotation for the following reproject calls (per model id 7..10). */
void orc_set_rotation(OrcCtx* c, int model, const float R[9]);

/* MVReprojection::reprojectMotionVectorSubblocks, MVReprojection.cpp:80-166.
 * pos/size are in units of component `comp` (0=Y,1=Cb,2=Cr); mv in 1/16 luma pel.
 * out_x/out_y: (size_h/sbH) x (size_w/sbW) ints, COLUMN-major like Eigen (idx = col*rows+row). */
void orc_reproject(OrcCtx* c, int pos_x, int pos_y, int size_w, int size_h, int mv_hor, int mv_ver,
                   int model, int comp, int chroma_fmt, int32_t* out_x, int32_t* out_y);

/* MPA perspective cache (MotionPlaneAdaptiveMotionModel::fillCache, :97-104) of plane model 1..3,
 * column-major (H/4 rows x W/4 cols); exposed so the device-built cache can be checked. */
const float* orc_mpa_cache_x(OrcCtx* c, int model);
const float* orc_mpa_cache_y(OrcCtx* c, int model);
const uint8_t* orc_mpa_cache_vip(OrcCtx* c, int model);

/* InterpolationFilter::filterHor / filterVer (InterpolationFilter.cpp:675-809 -> filter<>/filterCopy<>).
 * is_luma selects the 8-tap (16 phases) or 4-tap (32 phases) table. */
void orc_filter_hor(int is_luma, const int16_t* src, int src_stride, int16_t* dst, int dst_stride, int w, int h,
                    int frac, int is_last, int bit_depth, int clp_min, int clp_max);
void orc_filter_ver(int is_luma, const int16_t* src, int src_stride, int16_t* dst, int dst_stride, int w, int h,
                    int frac, int is_first, int is_last, int bit_depth, int clp_min, int clp_max);

/* InterPrediction::xPredInterBlkMM (InterPrediction.cpp:683-856, oob_mode ORC_OOB_PRED) and
 * InterSearch::xMVReprojectionInterpolation (InterSearch.cpp:6277-6385, ORC_OOB_ME, luma, bi=0).
 * ref points at sample (0,0) of the padded component plane; pic_w/pic_h are that component's size. */
void orc_mc_block(OrcCtx* c, const int16_t* ref, int ref_stride, int comp, int chroma_fmt, int pos_x, int pos_y,
                  int size_w, int size_h, int mv_hor, int mv_ver, int model, int bi, int clp_min, int clp_max,
                  int oob_mode, int16_t* dst, int dst_stride);

/* InterPrediction::xPredInterBlk, classic translational path (InterPrediction.cpp:928-1096, no RPR/BDOF/
 * DMVR-bilinear): one MV for the block, mv in 1/16 (luma) units, component shift applied inside. */
void orc_mc_block_classic(const int16_t* ref, int ref_stride, int comp, int chroma_fmt, int pos_x, int pos_y,
                          int size_w, int size_h, int mv_hor, int mv_ver, int bi, int bit_depth, int clp_min,
                          int clp_max, int16_t* dst, int dst_stride);

/* RdCost::xGetSAD* (RdCost.cpp:482-..., x86/RdCostX86.h:378-495): rows stepped by 1<<sub_shift. */
uint64_t orc_sad(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride, int w, int h, int sub_shift);
/* RdCost::xGetHADs (RdCost.cpp:2943-3067) / xGetHADs_SIMD (x86/RdCostX86.h:4413-4557). */
uint64_t orc_had(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride, int w, int h, int had_mode);

/* ME candidate evaluation = xMVReprojectionInterpolation + distFunc for n MVs of one block
 * (InterSearch.cpp:363-453, 6222-6249, 6263-6275).  mv_prec_shift: 4 = integer MVs, 3 = half,
 * 2 = quarter, 0 = internal (Mv::changePrecision to MV_PRECISION_INTERNAL). */
/* xPredInterBlk with bioApplied (InterPrediction.cpp:1007-1015, 1054-1093): tmp = (w + 4) x (h + 4) BDOF work buffer */
void orc_mc_bdof(const int16_t* ref, int rs, int pos_x, int pos_y, int size_w, int size_h, int mv_hor, int mv_ver, int bd, int lo,
                 int hi, int16_t* tmp);
void orc_me_block(OrcCtx* c, const int16_t* ref, int ref_stride, const int16_t* org, int org_stride, int pos_x,
                  int pos_y, int size_w, int size_h, int model, int n, const int32_t* mv_hor, const int32_t* mv_ver,
                  int mv_prec_shift, int dist_kind, int sub_shift, int had_mode, int clp_min, int clp_max,
                  uint64_t* dist);

/* PelBuf::addAvg via InterPrediction::xWeightedAverage (Buffer.cpp:76-..., default BCW weights). */
void orc_add_avg(const int16_t* src0, int s0_stride, const int16_t* src1, int s1_stride, int16_t* dst, int dst_stride,
                 int w, int h, int bit_depth, int clp_min, int clp_max);

/* InterPrediction::xProcessDMVRProjected integer search of ONE sub-PU (InterPrediction.cpp:2485-2583) with
 * xDMVRCost (:2147-2155: sized SAD, subShift 1, >>1).  ref0/ref1: luma planes of the two lists, R0/R1: geodesic
 * rotations of the two (curPOC,refPOC) pairs.  costs25[k] = xDMVRCost at m_pSearchOffset[k] (InterPrediction.h:107-111;
 * k = 12 is the starting point).  simd_sad != 0 reproduces the 16-bit wrap of xGetSAD_NxN_SIMD's _mm_sub_epi16 on
 * 14-bit intermediates (x86/RdCostX86.h:378-495). */
void orc_dmvr_costs(OrcCtx* c, const int16_t* ref0, int rs0, const int16_t* ref1, int rs1, int x, int y, int dx, int dy,
                    int mv0_hor, int mv0_ver, int mv1_hor, int mv1_ver, int model, const float* R0, const float* R1,
                    int clp_min, int clp_max, int simd_sad, uint64_t costs25[25]);
/* Decision logic of the same loop + xDMVRSubPixelErrorSurface/xSubPelErrorSrfc/div_for_maxq7 (:1958-2048, 2157-2175):
 * from the 25 costs to pu.mvdL0SubPu (1/16 pel) and minCost; returns notZeroCost. */
int orc_dmvr_decide(const uint64_t costs25[25], int dx, int dy, int32_t delta_mv[2], uint64_t* min_cost);

uint64_t orc_dmvr_cost(const int16_t* a, int as, const int16_t* b, int bs, int w, int h, int simd_sad);

/* Classic DMVR (InterPrediction::xProcessDMVR, InterPrediction.cpp:2239-2439), one sub-PU.  ref0/ref1: luma sample (0,0) of the
 * padded planes of list 0 / list 1.  (ox, oy): first prefetched luma sample (xPrefetch's Rec_offset), (fx, fy): fraction of the
 * clipped merge MV.  costs25 as orc_dmvr_costs (index 12 = centre, raw); the decision is orc_dmvr_decide. */
void orc_dmvr_classic_costs(const int16_t* ref0, int rs0, const int16_t* ref1, int rs1, int ox0, int oy0, int fx0, int fy0,
                            int ox1, int oy1, int fx1, int fy1, int dx, int dy, int bd, uint64_t costs25[25]);
/* xFinalPaddedMCForDMVR (:2075-2145) of one component of one list from the prefetched + padded window. */
void orc_dmvr_final_mc(const int16_t* ref, int rs, int is_luma, int ox, int oy, int w, int h, int padsize, int dix, int diy,
                       int xFrac, int yFrac, int alt_hpel, int bio, int bd, int lo, int hi, int16_t* dst, int ds);
/* classic luma MC with the alternative half-pel filter (pu.cu->imv == IMV_HPEL, InterPrediction.cpp:948) */
void orc_mc_block_classic_alt(const int16_t* ref, int rs, int pos_x, int pos_y, int size_w, int size_h, int mv_hor, int mv_ver, int bi,
                              int bd, int lo, int hi, int16_t* dst, int ds);

/* Picture::extendPicBorder (Picture.cpp:988-1048): replicate a w x h plane into its margins in place. */
void orc_extend_border(int16_t* plane00, int stride, int w, int h, int margin_x, int margin_y);
/* Picture::extendWrapBorder (Picture.cpp:1049-1092): margins continue the picture from the opposite side for the first
 * xoffset columns (component units), replicate beyond; in place. */
void orc_extend_wrap_border(int16_t* plane00, int stride, int w, int h, int margin_x, int margin_y, int xoffset);

#ifdef __cplusplus
}
#endif
#endif
