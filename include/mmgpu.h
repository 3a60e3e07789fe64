/* mmgpu.h -- C ABI of the B200 (sm_100a) multi-model 360 MC + block-matching distortion path.
 *
 * Drop-in boundary for FAU-LMS/vvc-extension-mm (VTM-17.2 + MM).  Each entry point names the
 * reference interface whose body it replaces; INTEGRATION.md shows the C++ stubs a maintainer
 * adds inside the reference.  Plain pointers and sizes only, `int` return (0 = ok, <0 = error,
 * text via mmgpu_last_error); no exceptions cross the boundary.  One context per
 * InterPrediction/InterSearch instance; a context is thread-compatible, not thread-safe
 * (the reference is single threaded).  There is no CPU fallback: every compute entry point
 * fails with MMGPU_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MMGPU_H
#define MMGPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MMGPU_OK 0
#define MMGPU_ERR_ARG (-1)
#define MMGPU_ERR_CUDA (-2)
#define MMGPU_ERR_STATE (-3)

/* MotionModelID, source/Lib/CommonLib/TypeDef.h:865-880 */
enum mmgpu_model {
  MMGPU_CLASSIC = 0, MMGPU_MPA_FRONT_BACK = 1, MMGPU_MPA_LEFT_RIGHT = 2, MMGPU_MPA_TOP_BOTTOM = 3,
  MMGPU_TANGENTIAL = 4, MMGPU_THREE_D_TRANSLATIONAL = 5, MMGPU_ROTATIONAL = 6, MMGPU_GEODESIC_X = 7,
  MMGPU_GEODESIC_Y = 8, MMGPU_GEODESIC_Z = 9, MMGPU_GEODESIC_CAMPOSE = 10
};
enum mmgpu_oob { MMGPU_OOB_PRED = 0 /* InterPrediction.cpp:776-780 */, MMGPU_OOB_ME = 1 /* InterSearch.cpp:6329-6332 */ };
enum mmgpu_dist { MMGPU_DIST_SAD = 0 /* RdCost::xGetSAD*, DF_SAD.. */, MMGPU_DIST_HAD = 1 /* RdCost::xGetHADs, DF_HAD.. */ };
enum mmgpu_had_mode { MMGPU_HAD_SCALAR = 0 /* RdCost.cpp:2637 */, MMGPU_HAD_SIMD = 1 /* x86/RdCostX86.h:2661 (AVX2/SSE hosts) */ };

typedef struct mmgpu_ctx mmgpu_ctx;

/* What EncLib::init (EncLib.cpp:327-357) / DecLib::activateParameterSets (DecLib.cpp:2014-2063) know
 * when they call MVReprojection::init (MVReprojection.cpp:7-65). */
typedef struct {
  int32_t pic_w, pic_h;  /* luma picture size (ERP) */
  int32_t chroma_fmt;    /* ChromaFormat: 0=400 1=420 2=422 3=444 */
  int32_t bit_depth;     /* internal bit depth (<= 10: the SIMD distortion functions' domain) */
  int32_t max_cu;        /* sps->getMaxCUWidth() (128) */
  int32_t margin;        /* luma margin of reference planes, Picture.cpp:91 (288) */
  float offset4x4;       /* MMOffset4x4 as float (EncApp.cpp:758 -> 1.0) */
  int32_t ged_flavor;    /* 0 = VISHWANATH_ORIGINAL, 1 = VISHWANATH_MODULATED (EncApp.cpp:755) */
  int32_t max_refs;      /* reference-picture slots kept resident in HBM */
  int32_t device;        /* CUDA device ordinal */
  int32_t libm_fma;      /* which glibc sinf/cosf build the host CPU selects: 1 = __sinf_fma, 0 = sse2, -1 = detect */
  int32_t had_mode;      /* mmgpu_had_mode; -1 = MMGPU_HAD_SIMD */
} mmgpu_config;

typedef struct { int32_t hor, ver; } mmgpu_mv; /* Mv, source/Lib/CommonLib/Mv.h */

/* One block of a batch: pu.blocks[COMPONENT_Y] plus what IntTZSearchStruct carries
 * (InterSearch.h: blkPos, blkSize, motionModel, curPOC/refPOC -> rot_slot, pcPatternKey -> org_x/y). */
typedef struct {
  int32_t x, y, w, h;     /* luma position / size */
  int32_t model;          /* mmgpu_model (CLASSIC allowed: translational candidates, a8) */
  int32_t ref_slot;       /* reference picture slot set by mmgpu_set_ref* */
  int32_t rot_slot;       /* rotation slot for GEODESIC_CAMPOSE (mmgpu_set_rotation); ignored otherwise */
  int32_t cand_first;     /* first candidate MV of this block in the mv array */
  int32_t cand_count;     /* number of candidate MVs */
  int32_t org_x, org_y;   /* top-left of the original block inside the org plane (mmgpu_set_org*) */
  int32_t reserved;
} mmgpu_block;

/* ---- lifetime ---------------------------------------------------------------------------------- */
int mmgpu_create(const mmgpu_config* cfg, mmgpu_ctx** out);  /* MVReprojection::init + fillCache on device */
void mmgpu_destroy(mmgpu_ctx* ctx);
const char* mmgpu_last_error(const mmgpu_ctx* ctx);          /* ctx may be NULL: error of the last create */
int mmgpu_sync(mmgpu_ctx* ctx);
/* cudaStream_t the context launches on (for event timing by the caller) */
void* mmgpu_stream(mmgpu_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches claim) */
uint64_t mmgpu_launch_count(const mmgpu_ctx* ctx);
/* CUDA-event timing of the two batch kernels on the context's stream (k_map: positions, k_fir: interpolation +
 * distortion): enable, run, then read the accumulated device times and the number of batches (synchronises the
 * stream; resets the accumulators). */
int mmgpu_kernel_timing_enable(mmgpu_ctx* ctx, int on);
int mmgpu_kernel_timing_read(mmgpu_ctx* ctx, double* map_ms, double* fir_ms, uint64_t* batches);

/* Self-test of the one place where the reference narrows a DOUBLE libm result to float (TangentialMotionModel.cpp:27-28,
 * 38,40): checksum of (x, float(sin((double)x)), float(cos((double)x))) over every stride-th float x with |x| <= pi/2,
 * computed by the device code the kernels use.  Must equal the glibc sweep of oracle/tools/tan_exceptions.cpp. */
int mmgpu_selftest_tan_sincos(mmgpu_ctx* ctx, uint32_t stride, uint64_t* checksum);

/* ---- reference pictures: Picture::getRecoBuf(compID, wrapRef) after extendPicBorder (Picture.cpp:988-1092) */
/* plane00[c] points at sample (0,0) of component c in HOST memory; the margins around it are copied too. */
int mmgpu_set_ref(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane00[3], const int32_t stride[3]);
/* same, but plane00[c] are DEVICE pointers owned by the caller (kept resident, zero copy; 8-byte aligned, stride a
 * multiple of 4 samples): used when the
 * picture arrives by ncclBroadcast or was reconstructed on the device. */
int mmgpu_set_ref_device(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane00[3], const int32_t stride[3]);
/* un-padded planes in host memory: upload + Picture::extendPicBorder on the device. */
int mmgpu_set_ref_unpadded(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane[3], const int32_t stride[3]);
/* same with the horizontally wrapped border of Picture::extendWrapBorder (Picture.cpp:1049-1092), i.e. the
 * PIC_RECON_WRAP buffer getRecoBuf(compID, wrapRef = true) returns when the PPS enables wrap-around (360 ERP):
 * wrap_offset = pps->getWrapAroundOffset() in luma samples (the ERP picture width for a full wrap). */
int mmgpu_set_ref_unpadded_wrap(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane[3], const int32_t stride[3],
                                int wrap_offset);

/* un-padded planes that already live in DEVICE memory (e.g. the landing buffer of an NCCL all-gather: only the picture
 * itself crosses NVLink, the margins are rebuilt on arrival): device-to-device copy into the slot's own padded planes +
 * border extension, enqueued on `stream` (a cudaStream_t; NULL = the context's stream).  wrap_offset as above (0 = replicate).
 * The caller orders `stream` against the batches that read the slot (events), as with any double-buffered reference. */
int mmgpu_set_ref_unpadded_device(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const d_plane[3], const int32_t stride[3],
                                  int wrap_offset, void* stream);

/* the same for n pictures at once (n * components <= 24): ONE kernel launch copies and extends all planes -- the per-step
 * form of the multi-GPU picture exchange.  d_plane[i*3 + c], stride[i*3 + c]. */
int mmgpu_set_refs_unpadded_device(mmgpu_ctx* ctx, int n, const int32_t* slots, const int32_t* pocs, const int16_t* const* d_plane,
                                   const int32_t* stride, int wrap_offset, void* stream);

/* the multi-GPU form of it: d_plane are the OWNERS' un-padded planes in peer memory (NVLink-mapped device addresses, e.g.
 * torch symmetric memory; the caller's own picture is a local address).  A persistent kernel of max_ctas CTAs pulls every
 * row once over NVLink and writes it with its left / right margin into the slot's padded plane; a second one replicates the
 * top / bottom rows locally (Picture::extendPicBorder, Picture.cpp:988-1048; replicate only).  The exchange and the border
 * extension are one pass over the data; max_ctas is the SM budget it takes from the kernels of the running step.
 * The caller brackets the call with its own cross-rank barriers (pictures published / pulls finished). */
int mmgpu_pull_refs_unpadded(mmgpu_ctx* ctx, int n, const int32_t* slots, const int32_t* pocs, const int16_t* const* d_plane,
                             const int32_t* stride, int max_ctas, void* stream);

/* PUSH form of the same exchange: d_dst_multicast is a multicast address of symmetric memory (NVLS: the NVSwitch replicates one
 * store into the same offset of every GPU's buffer, e.g. torch symmetric memory's multicast_ptr + offset of a landing reference
 * slot registered with mmgpu_set_ref_device).  The owner reads its padded picture once and writes it once (`multimem.st`);
 * per GPU and step that is ONE picture of SM work and NVLink egress.  Asynchronous on `stream`; the caller brackets it with its
 * cross-rank barriers (slot free everywhere / pushes landed). */
int mmgpu_multicast_copy(mmgpu_ctx* ctx, const void* d_src, void* d_dst_multicast, size_t bytes, int max_ctas, void* stream);

/* ---- geodesic epipole: GeodesicMotionModel::setEpipole (GeodesicMotionModel.cpp:14-46) ------------ */
void mmgpu_rodrigues(const float epipole[3], float R[9]);          /* host helper, row-major R */
float mmgpu_fixed_to_float(int32_t fixed, int32_t precision);     /* Coordinate.cpp:75-77 (EPIPOLE_PRECISION_FIXED=24) */
int mmgpu_set_rotation(mmgpu_ctx* ctx, int rot_slot, const float R[9]);

/* ---- a1: MVReprojection::reprojectMotionVectorSubblocks (MVReprojection.cpp:80-166) --------------- */
/* out_x/out_y: (h/sbH)*(w/sbW) ints, column-major like the reference's ArrayXXi (idx = col*rows + row). */
int mmgpu_reproject(mmgpu_ctx* ctx, int comp, int chroma_fmt, int x, int y, int w, int h, mmgpu_mv mv, int model,
                    int rot_slot, int32_t* out_x, int32_t* out_y);

/* ---- 8f-1: MVReprojection::motionVectorInDesiredMotionModel (MVReprojection.cpp:168-217; callers UnitTools.cpp:2272,2302,
 *      2932,...): the MV that, in motion model `model_desired`, moves sample `pos` like `mv_orig` does in `model_orig` (MM-MVP
 *      candidate conversion).  A scalar per-candidate computation inside merge / AMVP list construction: evaluated on the HOST
 *      with the same restated float code the kernels use (no launch, no copy).  shift_hor / shift_ver: MV fraction bits
 *      (MV_FRACTIONAL_BITS_INTERNAL [+ chroma scale]); rot_slot_*: rotation slots of the (cur, ref) POC pairs for
 *      GEODESIC_CAMPOSE (equal slots == equal epipoles); cand / cur: position and size of the candidate's and the current block. */
int mmgpu_mvp_convert(mmgpu_ctx* ctx, int pos_x, int pos_y, mmgpu_mv mv_orig, int model_orig, int model_desired, int shift_hor,
                      int shift_ver, int rot_slot_orig, int rot_slot_desired, int cand_x, int cand_y, int cand_w, int cand_h,
                      int cur_x, int cur_y, int cur_w, int cur_h, mmgpu_mv* out);

/* ---- a5/a6/a8: InterPrediction::xPredInterBlkMM (InterPrediction.cpp:683-856), xPredInterBlk (:928-1096),
 *      InterSearch::xMVReprojectionInterpolation (InterSearch.cpp:6277-6385) --------------------------- */
/* x,y,w,h in units of component `comp`; mv in 1/16 luma pel (MV_PRECISION_INTERNAL); bi!=0 keeps the
 * 14-bit intermediate (rndRes = !bi); dst is HOST memory (dst_stride in samples). */
int mmgpu_mc(mmgpu_ctx* ctx, int ref_slot, int comp, int chroma_fmt, int x, int y, int w, int h, mmgpu_mv mv, int model,
             int rot_slot, int bi, int clp_min, int clp_max, int oob_mode, int16_t* dst, int dst_stride);

/* xPredInterBlk with bioApplied (InterPrediction.cpp:1007-1015, 1054-1093): classic bi-prediction MC of a luma block into the
 * BDOF work buffer -- the (w x h) 14-bit prediction at (2,2) of a buffer with stride w + 4 and, around it, the one-sample ring
 * of integer-position reference samples ((ref << IF_INTERNAL_FRAC_BITS) - IF_INTERNAL_OFFS, taken at the nearest integer
 * position: offset 1 where the fraction is below half a sample).  dst: HOST pointer to the start of that buffer
 * (m_filteredBlockTmp[2 + refList][COMPONENT_Y]), (w + 4) x (h + 4) samples. */
int mmgpu_mc_bdof(mmgpu_ctx* ctx, int ref_slot, int x, int y, int w, int h, mmgpu_mv mv, int clp_min, int clp_max, int16_t* dst);

/* Whole-picture reconstruction shape (DecCu::xReconInter -> InterPrediction::motionCompensation -> xPredInterUni /
 * xPredInterBi, InterPrediction.cpp:455-681): many PUs x components in one call. */
typedef struct {
  int32_t comp;                  /* ComponentID 0/1/2 */
  int32_t x, y, w, h;            /* pu.blocks[comp] position / size (component units) */
  int32_t model;                 /* mmgpu_model, CLASSIC included */
  int32_t bi;                    /* 0: uni (list 0 fields), 1: bi-prediction: both lists at 14 bit + PelBuf::addAvg */
  int32_t ref_slot0, rot_slot0;
  mmgpu_mv mv0;                  /* MV_PRECISION_INTERNAL */
  int32_t ref_slot1, rot_slot1;
  mmgpu_mv mv1;
} mmgpu_mc_job;
/* dst[c]: HOST planes of the (un-padded) picture being reconstructed; every job writes its block at (x,y). */
/* (dst planes in page-locked host memory -- cudaHostRegister or a pinned allocation -- are read back by one direct copy per plane:
 * 0.43 instead of 0.76-1.04 ms per 2048x1024 4:2:0 picture) */
int mmgpu_mc_jobs(mmgpu_ctx* ctx, int n_jobs, const mmgpu_mc_job* jobs, int clp_min, int clp_max, int16_t* const dst[3],
                  const int32_t dst_stride[3]);

/* same with the destination planes in DEVICE memory (a reconstruction that stays on the GPU: no picture crosses PCIe);
 * asynchronous on mmgpu_stream().  d_dst[c] 2-byte aligned device planes of the un-padded picture. */
int mmgpu_mc_jobs_device(mmgpu_ctx* ctx, int n_jobs, const mmgpu_mc_job* jobs, int clp_min, int clp_max, int16_t* const d_dst[3],
                         const int32_t dst_stride[3]);

/* ---- a13: InterPrediction::xWeightedAverage -> PelBuf::addAvg (Buffer.cpp:551-587) ------------------ */
int mmgpu_add_avg(mmgpu_ctx* ctx, const int16_t* src0, int s0_stride, const int16_t* src1, int s1_stride, int16_t* dst,
                  int dst_stride, int w, int h, int clp_min, int clp_max);

/* ---- a12: InterPrediction::xProcessDMVRProjected + xDMVRCost (InterPrediction.cpp:2442-2635, 2147-2155) -- */
/* One DMVR sub-PU (min(16,w) x min(16,h) luma, quirk Q6: the sub-PU is the block for the mapping). */
typedef struct {
  int32_t x, y, w, h;            /* luma position / size of the sub-PU (8x16, 16x8, 16x16; 8x8 accepted) */
  int32_t model;                 /* non-classic motion model of the PU */
  int32_t ref_slot0, ref_slot1;  /* reference pictures of list 0 / list 1 */
  int32_t rot_slot0, rot_slot1;  /* geodesic rotations of (curPOC, refPOC_L0) / (curPOC, refPOC_L1) */
  mmgpu_mv mv0, mv1;             /* merge MVs of the PU (MV_PRECISION_INTERNAL) */
} mmgpu_dmvr_subpu;
/* costs[i*25 + k] = xDMVRCost(MC(L0, mv0 + o_k), MC(L1, mv1 - o_k)) for the 25 offsets o_k of m_pSearchOffset
 * (InterPrediction.h:107-111), both MCs at bi = true.  HOST pointers. */
int mmgpu_dmvr_costs(mmgpu_ctx* ctx, int n_subpus, const mmgpu_dmvr_subpu* subpus, int clp_min, int clp_max, uint64_t* costs);
/* Host-side decision of the same loop incl. xDMVRSubPixelErrorSurface (InterPrediction.cpp:1958-2048, 2157-2175,
 * 2502-2583): 25 costs -> pu.mvdL0SubPu (1/16 pel) and minCost; returns notZeroCost (0/1). */
int mmgpu_dmvr_decide(const uint64_t costs25[25], int w, int h, int32_t delta_mv[2], uint64_t* min_cost);

/* ---- f4 (second half): classic DMVR, InterPrediction::xProcessDMVR (InterPrediction.cpp:2239-2439) ------------------ */
/* One sub-PU (min(16,w) x min(16,h) luma) of the integer search.  The reference prefetches a (w + 7) x (h + 7) window per
 * list (xPrefetch :1891-1936) and filters bilinear predictions from it (xinitMC :2177-2223, FILTER_DMVR): org = position of
 * the first prefetched luma sample = pu.lumaPos() + (clipMv(mergeMv - (3 << 4)) >> 4), frac = clipMv(mergeMv) & 15. */
typedef struct {
  int32_t w, h;                              /* 8 or 16 each */
  int32_t ref_slot0, ref_slot1;
  int32_t org_x0, org_y0, org_x1, org_y1;    /* luma samples, picture coordinates (may lie in the margin) */
  int32_t frac_x0, frac_y0, frac_x1, frac_y1;
} mmgpu_dmvr_classic_subpu;
/* costs[i*25 + k] = xDMVRCost (:2147-2155) of the bilinear predictions, L0 moved by +o_k and L1 by -o_k, o_k = m_pSearchOffset[k]
 * (index 12 = centre, raw: the caller's mmgpu_dmvr_decide applies minCost -= minCost >> 2 and the rest of :2364-2395). */
int mmgpu_dmvr_classic_costs(mmgpu_ctx* ctx, int n_subpus, const mmgpu_dmvr_classic_subpu* subpus, uint64_t* costs);
/* One component of one list of xFinalPaddedMCForDMVR (:2075-2145): xPredInterBlk(bi = true) reading from the prefetched
 * window padded by replication (xPad :1938-1956) instead of the picture. */
typedef struct {
  int32_t comp, ref_slot;
  int32_t org_x, org_y;        /* first prefetched sample of the component (xPrefetch's Rec_offset) */
  int32_t w, h;                /* component block (<= 16 x 16) */
  int32_t pad;                 /* xPad's padsize: 2 >> scaleY (0: the window is not padded, e.g. a sub-PU that did not move) */
  int32_t dint_x, dint_y;      /* deltaIntMv (:2126-2127), |d| <= pad */
  int32_t frac_x, frac_y;      /* of cMvClipped: luma & 15, chroma phase of 32 (:965-971) */
  int32_t alt_hpel;            /* pu.cu->imv == IMV_HPEL (:948) */
  int32_t bio;                 /* luma with bioApplied: (w + 4) x (h + 4) BDOF buffer, prediction at (2,2), ring of integer samples */
  int32_t dst_offset, dst_stride; /* where the job's output starts in dst (samples) and its stride (bio: stride is w + 4) */
} mmgpu_dmvr_mc_job;
/* dst: HOST buffer of dst_samples int16 (14-bit intermediates as xPredInterBlk with bi = true writes them). */
int mmgpu_dmvr_final_mc(mmgpu_ctx* ctx, int n_jobs, const mmgpu_dmvr_mc_job* jobs, int16_t* dst, int dst_samples);
/* xPredInterBlk (:928-1096) of a classic luma block with the alternative half-pel filter (pu.cu->imv == IMV_HPEL). */
int mmgpu_mc_alt_hpel(mmgpu_ctx* ctx, int ref_slot, int x, int y, int w, int h, mmgpu_mv mv, int bi, int clp_min, int clp_max,
                      int16_t* dst, int dst_stride);

/* ---- a9-a11 + a6: candidate evaluation, xTZSearchHelp/xPatternRefinementProjected's
 *      xMVReprojectionInterpolation + m_cDistParam.distFunc pair (InterSearch.cpp:363-453, 6222-6249) --- */
/* original samples (pcPatternKey or the whole current picture), HOST memory */
int mmgpu_set_org(mmgpu_ctx* ctx, const int16_t* org, int stride, int w, int h);
int mmgpu_set_org_device(mmgpu_ctx* ctx, const int16_t* org, int stride, int w, int h);
/* Evaluate every candidate MV of every block: dist[i] = distFunc(org block, MC(block, mvs[i])).
 * mv_prec_shift: mvs are given in MV_PRECISION_INT (4), HALF (3), QUARTER (2) or INTERNAL (0) units
 * (Mv::changePrecision, InterSearch.cpp:6294).  dist_kind/sub_shift as set by RdCost::setDistParam
 * (RdCost.cpp:225-311).  All pointers HOST memory; the call returns when dist[] is filled. */
int mmgpu_me_batch(mmgpu_ctx* ctx, int n_blocks, const mmgpu_block* blocks, int n_cands, const mmgpu_mv* mvs,
                   int mv_prec_shift, int dist_kind, int sub_shift, int clp_min, int clp_max, uint64_t* dist);

/* Out-of-picture rule of the candidate evaluation (mmgpu_me_batch*, _submit, _device, _multiref): MMGPU_OOB_ME (default,
 * xMVReprojectionInterpolation, InterSearch.cpp:6329-6332) for the motion search; MMGPU_OOB_PRED (xPredInterBlkMM,
 * InterPrediction.cpp:776-780) for distortions of decided MVs -- merge-candidate SATD pre-selection (EncCu.cpp:2375-2600) and
 * xGetInterPredictionError (InterSearch.cpp:855-868), whose predictions come from motionCompensation. */
int mmgpu_set_dist_oob_mode(mmgpu_ctx* ctx, int oob_mode);

/* Pipelined form of the same call for callers that have the next batch ready while this one runs (frame-level
 * look-ahead ME): submit copies org/blocks/mvs on a copy stream into buffers owned by `slot` (0..3) and queues the
 * kernels and the read-back of dist behind them; wait returns when dist[] of that slot is filled.  With two slots the
 * host->device copies of batch i+1 overlap the kernels of batch i.  All host buffers must stay valid (and should be
 * pinned) until the wait; a slot must be waited for before it is submitted again. */
int mmgpu_me_batch_submit(mmgpu_ctx* ctx, int slot, const int16_t* org, int org_stride, int org_w, int org_h, int n_blocks,
                          const mmgpu_block* blocks, int n_cands, const mmgpu_mv* mvs, int mv_prec_shift, int dist_kind,
                          int sub_shift, int clp_min, int clp_max, uint64_t* dist);
int mmgpu_me_batch_wait(mmgpu_ctx* ctx, int slot);

/* Frame-wide microbench entry (SURVEY 8b mmgpu_frame_batch): the same evaluation with blocks, mvs and dist
 * already resident in HBM; asynchronous on mmgpu_stream().  `plan` comes from mmgpu_plan_create (host-built
 * work list for these blocks, uploaded once) and can be reused while the block list is unchanged. */
typedef struct mmgpu_plan mmgpu_plan;
int mmgpu_plan_create(mmgpu_ctx* ctx, int n_blocks, const mmgpu_block* blocks, int n_cands, mmgpu_plan** out);
void mmgpu_plan_destroy(mmgpu_plan* plan);
int mmgpu_me_batch_device(mmgpu_ctx* ctx, const mmgpu_plan* plan, const mmgpu_mv* d_mvs, int mv_prec_shift, int dist_kind,
                          int sub_shift, int clp_min, int clp_max, uint64_t* d_dist);
/* number of luma sample-candidates (sum over blocks of w*h*cand_count) a plan evaluates */
uint64_t mmgpu_plan_samples(const mmgpu_plan* plan);

/* The same candidate set against SEVERAL reference pictures (the reference loops iRefIdxTemp over the same block and search
 * pattern, InterSearch.cpp:2695-...; the positions of MPA / ROT / TAN / 3DT do not depend on the reference picture, those of
 * GEODESIC_CAMPOSE only through the epipole of the (current, reference) POC pair): the float mapping (k_map) runs once per
 * distinct rotation slot, the interpolation + distortion (k_fir) once per reference.  ref_slots[r] replaces every block's
 * ref_slot, rot_slots[r] every block's rot_slot (NULL: keep the blocks' own); d_dist holds n_refs * n_cands distortions,
 * reference-major.  Device pointers for d_mvs / d_dist, host pointers for the slot lists; asynchronous like
 * mmgpu_me_batch_device. */
int mmgpu_me_batch_device_multiref(mmgpu_ctx* ctx, const mmgpu_plan* plan, const mmgpu_mv* d_mvs, int mv_prec_shift, int dist_kind,
                                   int sub_shift, int clp_min, int clp_max, int n_refs, const int32_t* ref_slots,
                                   const int32_t* rot_slots, uint64_t* d_dist);

/* MC of a whole batch (one MV per block: the first candidate), predictions written into a device picture
 * buffer at (x,y): final-MC / merge-candidate shape of the path (EncCu.cpp:2375-2600). */
int mmgpu_mc_batch_device(mmgpu_ctx* ctx, const mmgpu_plan* plan, const mmgpu_mv* d_mvs, int mv_prec_shift, int clp_min,
                          int clp_max, int16_t* d_dst, int dst_stride);

#ifdef __cplusplus
}
#endif
#endif
