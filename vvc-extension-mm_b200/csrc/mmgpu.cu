// mmgpu.cu -- C ABI (include/mmgpu.h) over the sm_100a kernels of mm_kernels_{map,fir,mc}.cu (launch wrappers: mm_common.cuh).
// Host side is plain C++ + CUDA runtime; no torch types, no CPU fallback of the compute path.
#include <cuda.h>          // CUtensorMap types; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "mm_common.cuh"
#include "mm_fir.cuh"
#include "mm_mvp.cuh"

#define MM_MAX_ROT 64

struct mmgpu_plan
{
  mmgpu_ctx* ctx;
  int nBlocks, nCands, nItems;
  uint64_t samples;
  mmgpu_block* dBlocks;
  MmItem* dItems;
  int32_t* dCandBlock;
  MmCand* dCands;
  int32_t* dItemBase;
  int2* dPos;  // 1/16-pel position of every (candidate, sub-block) entry: k_map -> k_fir
  int modelFirst[MM_NUM_MODELS], modelCount[MM_NUM_MODELS];  // work items of each motion model (contiguous)
  int orgMaxX, orgMaxY;   // max(org_x + w), max(org_y + h) over the blocks: checked against the org plane at launch
  int maxCandCount;       // largest cand_count of a block (mmgpu_mc_batch_device wants exactly one MV per block)
  size_t capBlocks, capItems, capCands, capCandParams, capItemBase, capPos;
};

struct MmStage  // pinned host staging for the small per-call tables of a plan (reused; guarded by an event)
{
  void* h;
  size_t cap;
  cudaEvent_t freeEv;
  bool busy;
};

#define MM_MAX_INFLIGHT 4
struct MmSlot  // one in-flight mmgpu_me_batch_submit: its own device buffers, so copies of the next batch overlap kernels
{
  mmgpu_plan plan;
  MmStage stage;
  mmgpu_mv* dMvs;
  size_t capMvs;
  unsigned long long* dDist;
  size_t capDist;
  int16_t* dOrg;
  size_t capOrg;
  cudaEvent_t copied, prepared, computed, done;
  bool init, pending;
};

struct mmgpu_ctx
{
  mmgpu_config cfg;
  MmGeom g;
  MmFir fir;
  cudaStream_t stream;
  std::string err;
  uint64_t launches;
  // reference slots
  MmRef* dRefs;
  std::vector<MmRef> hRefs;
  std::vector<int16_t*> ownedPlanes;  // [slot*3+c] base allocations owned by the context
  // rotations
  float* dRot;
  std::vector<float> hRot;  // host mirror (mmgpu_mvp_convert runs on the host)
  // MPA caches
  float2* dMpa[3];
  uint8_t* dMpaVip[3];
  float2* dThTab[2];
  float2* dPhTab[2];
  uint32_t* dHTaps;  // mm_build_htaps: alignment-folded luma coefficient words for k_fir
  CUtensorMap* dTmaps;  // [max_refs][MM_TMA_TYPES] TMA descriptors of the padded luma planes (k_fir's shared-memory boxes)
  void* encodeTiled;    // cuTensorMapEncodeTiled
  // original picture
  int16_t* dOrgOwned;
  size_t orgCap;
  const int16_t* dOrg;
  int orgStride, orgW, orgH;
  // scratch
  mmgpu_plan scratchPlan;
  mmgpu_mv* dMvs;
  size_t capMvs;
  unsigned long long* dDist;
  size_t capDist;
  int16_t* dTmp;  // small device buffers for mmgpu_mc / reproject / add_avg
  size_t capTmp;
  mmgpu_dmvr_subpu* dSubpus;
  size_t capSubpus;
  mmgpu_mc_job* dJobs;
  size_t capJobs;
  int2* dJobItems;
  size_t capJobItems;
  int16_t* dPic[3];   // reconstruction picture planes for mmgpu_mc_jobs
  size_t capPic[3];
  MmDmvrItem* dDmvrItems;
  size_t capDmvrItems;
  mmgpu_dmvr_classic_subpu* dSubpusC;   // classic DMVR: search sub-PUs / final-MC jobs
  size_t capSubpusC;
  mmgpu_dmvr_mc_job* dDmvrJobs;
  size_t capDmvrJobs;
  bool mapPerModel;
  int distOob;              // OOB rule of the DIST batches (mmgpu_set_dist_oob_mode)
  MmStage stage;            // staging of the synchronous entry points
  cudaStream_t copyStream;  // host->device copies of submitted batches
  cudaStream_t prepStream;  // work list + per-candidate scalars of a submitted batch (overlaps the previous batch's k_fir)
  cudaStream_t readStream;  // device->host read-back of their distortions (so it does not delay the next batch's kernels)
  MmSlot slots[MM_MAX_INFLIGHT];
  // optional CUDA-event timing of the batch kernels: events t0 | k_map | t1 | k_fir | t2 per launchMe
  bool timing;
  std::vector<cudaEvent_t> evPool;   // triples
  size_t evUsed;
  double evAccumMs[2];               // [0] = k_map, [1] = k_fir
  uint64_t evAccumN;
};

static thread_local std::string g_createErr;   // error text of the last failed mmgpu_create of this thread

static int fail(mmgpu_ctx* c, int code, const std::string& msg)
{
  if (c) c->err = msg;
  else g_createErr = msg;
  return code;
}
#define CU(call)                                                                                           \
  do                                                                                                       \
  {                                                                                                        \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
      return fail(ctx, MMGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
  } while (0)

static int scaleX(int comp, int fmt) { return (comp == 0) ? 0 : ((fmt == 1 || fmt == 2) ? 1 : 0); }
static int scaleY(int comp, int fmt) { return (comp == 0) ? 0 : ((fmt == 1) ? 1 : 0); }
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static MmParams baseParams(mmgpu_ctx* ctx)
{
  MmParams P;
  memset(&P, 0, sizeof(P));
  P.g = ctx->g;
  P.fir = ctx->fir;
  P.refs = ctx->dRefs;
  P.rot = ctx->dRot;
  for (int i = 0; i < 3; ++i) { P.mpa[i] = ctx->dMpa[i]; P.mpaVip[i] = ctx->dMpaVip[i]; }
  for (int i = 0; i < 2; ++i) { P.thTab[i] = ctx->dThTab[i]; P.phTab[i] = ctx->dPhTab[i]; }
  P.hTaps = ctx->dHTaps;
  P.tmaps = ctx->dTmaps;
  P.tmaOffX = ctx->cfg.margin;
  P.tmaOffY = ctx->cfg.margin;
  P.org = ctx->dOrg;
  P.orgStride = ctx->orgStride;
  P.hadMode = ctx->cfg.had_mode;
  P.picW = ctx->cfg.pic_w;
  P.picH = ctx->cfg.pic_h;
  P.maxCU = ctx->cfg.max_cu;
  P.clpMin = 0;
  P.clpMax = (1 << ctx->cfg.bit_depth) - 1;
  P.refOverride = -1;
  P.rotOverride = -1;
  return P;
}

extern "C" void mmgpu_destroy(mmgpu_ctx* ctx);
extern "C" const char* mmgpu_last_error(const mmgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createErr.c_str(); }

extern "C" void mmgpu_rodrigues(const float e[3], float R[9])
{
  // GeodesicMotionModel::setEpipole (GeodesicMotionModel.cpp:25-45); Eigen fixed 3x3 lazy products
  // evaluate each coefficient as a0*b0 + (a1*b1 + a2*b2).
  const float norm = sqrtf((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]));
  const float px = e[0] / norm, py = e[1] / norm, pz = e[2] / norm;
  const float cx = -py, cy = px, cz = 0.0f;
  const float s = sqrtf(cx * cx + cy * cy + cz * cz);
  float M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (s == 0)
  {
    if (pz < 0) M[8] = -1;
    memcpy(R, M, sizeof(M));
    return;
  }
  const float one = 1.0f, mone = -1.0f;
  const float mn = (pz < one) ? pz : one;        // std::min(TCoord(1), z)
  const float cc = (mone < mn) ? mn : mone;      // std::max(TCoord(-1), .)
  float K[9] = {0, -cz, cy, cz, 0, -cx, -cy, cx, 0};
  float KK[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      const float p0 = K[i * 3 + 0] * K[0 * 3 + j], p1 = K[i * 3 + 1] * K[1 * 3 + j], p2 = K[i * 3 + 2] * K[2 * 3 + j];
      const float s12 = p1 + p2;
      KK[i * 3 + j] = p0 + s12;
    }
  const float f = (1 - cc) / (s * s);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      const float id = (i == j) ? 1.0f : 0.0f;
      const float t = id + K[i * 3 + j];
      const float u = KK[i * 3 + j] * f;
      M[i * 3 + j] = t + u;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = M[j * 3 + i];  // transposeInPlace
}

extern "C" float mmgpu_fixed_to_float(int32_t fixed, int32_t precision)
{
  return (float)(fixed >> precision) + (float)(fixed & ((1 << precision) - 1)) / (float)(1 << precision);
}

// inside mmgpu_create a failing CUDA call must not leak the half-built context, and its message must be readable through
// mmgpu_last_error(NULL)
#define CUC(call)                                                                                          \
  do                                                                                                       \
  {                                                                                                        \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess)                                                                                 \
    {                                                                                                      \
      const std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
      cudaGetLastError();                                                                                  \
      mmgpu_destroy(ctx);                                                                                  \
      return fail(nullptr, MMGPU_ERR_CUDA, m_);                                                            \
    }                                                                                                      \
  } while (0)
extern "C" int mmgpu_create(const mmgpu_config* cfg, mmgpu_ctx** out)
{
  mmgpu_ctx* ctx = nullptr;
  if (!cfg || !out) return fail(nullptr, MMGPU_ERR_ARG, "null argument");
  if (cfg->pic_w <= 0 || cfg->pic_h <= 0 || (cfg->pic_w & 7) || (cfg->pic_h & 7))
    return fail(nullptr, MMGPU_ERR_ARG, "picture size must be a positive multiple of 8");
  if (cfg->bit_depth < 8 || cfg->bit_depth > 10) return fail(nullptr, MMGPU_ERR_ARG, "bit_depth must be 8..10");
  if (cfg->max_refs <= 0 || cfg->max_refs > 1024) return fail(nullptr, MMGPU_ERR_ARG, "max_refs out of range");
  if (cfg->chroma_fmt < 0 || cfg->chroma_fmt > 3) return fail(nullptr, MMGPU_ERR_ARG, "chroma_fmt must be 0..3 (400, 420, 422, 444)");
  if (!pow2(cfg->max_cu) || cfg->max_cu < 16 || cfg->max_cu > 128) return fail(nullptr, MMGPU_ERR_ARG, "max_cu must be 16, 32, 64 or 128");
  // memory safety of the kernels: the PRED zero-fill rule lets a position reach picW + max_cu, the 8-tap window adds 4+4 and
  // the wide row loads up to 12 more samples; classic MVs are held inside margin - 16.  8-byte row loads and the TMA
  // descriptors need plane (0,0) = base + margin on a 4-sample boundary for luma and chroma alike.
  if (cfg->margin < cfg->max_cu + 16 || (cfg->margin & 7))
    return fail(nullptr, MMGPU_ERR_ARG, "margin must be a multiple of 8 and >= max_cu + 16 (the reference uses 2 * (max_cu + 16))");
  if (cfg->ged_flavor != 0 && cfg->ged_flavor != 1) return fail(nullptr, MMGPU_ERR_ARG, "ged_flavor must be 0 or 1");
  if (cfg->had_mode > MMGPU_HAD_SIMD) return fail(nullptr, MMGPU_ERR_ARG, "bad had_mode");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, MMGPU_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MMGPU_ERR_ARG, "bad device ordinal");
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) return fail(nullptr, MMGPU_ERR_CUDA, cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10)
    return fail(nullptr, MMGPU_ERR_CUDA, "device is not sm_100 (this library carries sm_100a code only)");

  ctx = new mmgpu_ctx();
  ctx->cfg = *cfg;
  if (ctx->cfg.had_mode < 0) ctx->cfg.had_mode = MMGPU_HAD_SIMD;
  if (ctx->cfg.libm_fma < 0)
  {
    __builtin_cpu_init();
    ctx->cfg.libm_fma = (__builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2")) ? 1 : 0;
  }
  ctx->launches = 0;
  ctx->distOob = MMGPU_OOB_ME;
  ctx->timing = false;
  memset(&ctx->stage, 0, sizeof(ctx->stage));
  memset(ctx->slots, 0, sizeof(ctx->slots));
  ctx->mapPerModel = getenv("MMGPU_MAP_PER_MODEL") != nullptr;
  ctx->evUsed = 0;
  ctx->evAccumMs[0] = ctx->evAccumMs[1] = 0;
  ctx->evAccumN = 0;
  MmGeom& g = ctx->g;
  g.picW = (float)cfg->pic_w;
  g.picH = (float)cfg->pic_h;
  g.focal = (float)(1. / tan(M_PI / cfg->pic_h));   // Projection.h:128, evaluated by the host libm like the reference
  g.angRes = (float)(M_PI / cfg->pic_h);            // MVReprojection.cpp:27
  g.off = cfg->offset4x4;
  g.rows4 = cfg->pic_h / 4;
  g.cols4 = cfg->pic_w / 4;
  g.gedModulated = cfg->ged_flavor == 1;
  g.fmaBuild = ctx->cfg.libm_fma;
  const int hr = (14 - cfg->bit_depth) < 2 ? 2 : (14 - cfg->bit_depth);
  MmFir& f = ctx->fir;
  f.shift1 = 6 - hr;
  f.off1 = -8192 * (1 << f.shift1);
  f.shift2 = 6 + hr;
  f.off2 = (1 << (f.shift2 - 1)) + (8192 << 6);
  f.shift2bi = 6;
  f.shift1l = 6;
  f.off1l = 32;
  f.cpShift = hr;

  CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&ctx->readStream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithFlags(&ctx->prepStream, cudaStreamNonBlocking));
  CUC(cudaEventCreateWithFlags(&ctx->stage.freeEv, cudaEventDisableTiming));
  ctx->hRefs.assign(cfg->max_refs, MmRef());
  ctx->ownedPlanes.assign((size_t)cfg->max_refs * 3, nullptr);
  CUC(cudaMalloc(&ctx->dRefs, sizeof(MmRef) * cfg->max_refs));
  CUC(cudaMemset(ctx->dRefs, 0, sizeof(MmRef) * cfg->max_refs));
  CUC(cudaMalloc(&ctx->dRot, sizeof(float) * 9 * (MM_ROT_USER0 + MM_MAX_ROT)));
  {
    std::vector<float> rot(9 * (MM_ROT_USER0 + MM_MAX_ROT));
    const float ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1}, e0[3] = {0, 0, 0};
    mmgpu_rodrigues(ex, &rot[0]);   // MVReprojection.cpp:42-52
    mmgpu_rodrigues(ey, &rot[9]);
    mmgpu_rodrigues(ez, &rot[18]);
    for (int s = 0; s < MM_MAX_ROT; ++s) mmgpu_rodrigues(e0, &rot[9 * (MM_ROT_USER0 + s)]);  // EpipoleList.h:15-17
    CUC(cudaMemcpy(ctx->dRot, rot.data(), rot.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->hRot = rot;
  }
  const int n4 = g.rows4 * g.cols4;
  for (int m = 0; m < 3; ++m)
  {
    CUC(cudaMalloc(&ctx->dMpa[m], sizeof(float2) * n4));
    CUC(cudaMalloc(&ctx->dMpaVip[m], n4));
    mmk_mpa_cache(ctx->stream, g, m + 1, ctx->dMpa[m], ctx->dMpaVip[m]);
    ctx->launches++;
  }
  for (int v = 0; v < 2; ++v)
  {
    CUC(cudaMalloc(&ctx->dThTab[v], sizeof(float2) * g.rows4));
    CUC(cudaMalloc(&ctx->dPhTab[v], sizeof(float2) * g.cols4));
  }
  {
    cudaDriverEntryPointQueryResult qres;
    ctx->encodeTiled = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encodeTiled, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ctx->encodeTiled = nullptr;   // no descriptors: k_fir gathers from global memory
    cudaGetLastError();
    CUC(cudaMalloc(&ctx->dTmaps, sizeof(CUtensorMap) * (size_t)cfg->max_refs * MM_TMA_TYPES));
    CUC(cudaMemset(ctx->dTmaps, 0, sizeof(CUtensorMap) * (size_t)cfg->max_refs * MM_TMA_TYPES));
  }
  {
    std::vector<uint32_t> ht(MM_HTAPS_WORDS + MM_HTAPS_BASE_WORDS);
    mm_build_htaps(ht.data(), h_lumaFilter);
    for (int ph = 0; ph < 16; ++ph)   // entry o = 0, j = 0 holds the un-shifted words (C0, C1, 0)
    {
      ht[MM_HTAPS_WORDS + 2 * ph] = ht[ph * 12];
      ht[MM_HTAPS_WORDS + 2 * ph + 1] = ht[ph * 12 + 1];
    }
    CUC(cudaMalloc(&ctx->dHTaps, sizeof(uint32_t) * ht.size()));
    CUC(cudaMemcpy(ctx->dHTaps, ht.data(), sizeof(uint32_t) * ht.size(), cudaMemcpyHostToDevice));
  }
  mmk_trig_tables(ctx->stream, g, ctx->dThTab[0], ctx->dThTab[1], ctx->dPhTab[0], ctx->dPhTab[1]);
  ctx->launches++;
  CUC(cudaGetLastError());
  CUC(cudaStreamSynchronize(ctx->stream));
  *out = ctx;
  return MMGPU_OK;
}

static void freePlan(mmgpu_plan* p)
{
  cudaFree(p->dBlocks);
  cudaFree(p->dItems);
  cudaFree(p->dCandBlock);
  cudaFree(p->dCands);
  cudaFree(p->dItemBase);
  cudaFree(p->dPos);
  p->dItemBase = nullptr;
  p->dPos = nullptr;
  p->capItemBase = p->capPos = 0;
  p->dBlocks = nullptr; p->dItems = nullptr; p->dCandBlock = nullptr; p->dCands = nullptr;
  p->capBlocks = p->capItems = p->capCands = p->capCandParams = 0;
}

extern "C" void mmgpu_destroy(mmgpu_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto p : ctx->ownedPlanes) cudaFree(p);
  cudaFree(ctx->dRefs);
  cudaFree(ctx->dRot);
  for (int m = 0; m < 3; ++m) { cudaFree(ctx->dMpa[m]); cudaFree(ctx->dMpaVip[m]); }
  for (int v = 0; v < 2; ++v) { cudaFree(ctx->dThTab[v]); cudaFree(ctx->dPhTab[v]); }
  cudaFree(ctx->dHTaps);
  cudaFree(ctx->dTmaps);
  cudaFree(ctx->dOrgOwned);
  freePlan(&ctx->scratchPlan);
  cudaFree(ctx->dMvs);
  cudaFree(ctx->dDist);
  cudaFree(ctx->dTmp);
  if (ctx->stage.h) cudaFreeHost(ctx->stage.h);
  if (ctx->stage.freeEv) cudaEventDestroy(ctx->stage.freeEv);
  if (ctx->copyStream) cudaStreamSynchronize(ctx->copyStream);
  if (ctx->readStream) cudaStreamSynchronize(ctx->readStream);
  if (ctx->prepStream) cudaStreamSynchronize(ctx->prepStream);
  for (int i = 0; i < MM_MAX_INFLIGHT; ++i)
  {
    MmSlot& sl = ctx->slots[i];
    if (!sl.init) continue;
    freePlan(&sl.plan);
    if (sl.stage.h) cudaFreeHost(sl.stage.h);
    cudaEventDestroy(sl.stage.freeEv);
    cudaEventDestroy(sl.copied);
    cudaEventDestroy(sl.prepared);
    cudaEventDestroy(sl.computed);
    cudaEventDestroy(sl.done);
    cudaFree(sl.dMvs);
    cudaFree(sl.dDist);
    cudaFree(sl.dOrg);
  }
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  if (ctx->readStream) cudaStreamDestroy(ctx->readStream);
  if (ctx->prepStream) cudaStreamDestroy(ctx->prepStream);
  cudaFree(ctx->dSubpus);
  cudaFree(ctx->dDmvrItems);
  cudaFree(ctx->dSubpusC);
  cudaFree(ctx->dDmvrJobs);
  cudaFree(ctx->dJobs);
  cudaFree(ctx->dJobItems);
  for (int c = 0; c < 3; ++c) cudaFree(ctx->dPic[c]);
  for (auto e : ctx->evPool) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
}

extern "C" int mmgpu_sync(mmgpu_ctx* ctx)
{
  if (!ctx) return MMGPU_ERR_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}
static int drainTiming(mmgpu_ctx* ctx)
{
  CU(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i + 2 < ctx->evUsed; i += 3)
  {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ctx->evPool[i], ctx->evPool[i + 1]));
    ctx->evAccumMs[0] += ms;
    CU(cudaEventElapsedTime(&ms, ctx->evPool[i + 1], ctx->evPool[i + 2]));
    ctx->evAccumMs[1] += ms;
    ctx->evAccumN++;
  }
  ctx->evUsed = 0;
  return MMGPU_OK;
}
extern "C" int mmgpu_kernel_timing_enable(mmgpu_ctx* ctx, int on)
{
  if (!ctx) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  int rc = drainTiming(ctx);
  if (rc) return rc;
  ctx->timing = on != 0;
  ctx->evAccumMs[0] = ctx->evAccumMs[1] = 0;
  ctx->evAccumN = 0;
  return MMGPU_OK;
}
extern "C" int mmgpu_kernel_timing_read(mmgpu_ctx* ctx, double* map_ms, double* fir_ms, uint64_t* launches)
{
  if (!ctx || !map_ms || !fir_ms || !launches) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  int rc = drainTiming(ctx);
  if (rc) return rc;
  *map_ms = ctx->evAccumMs[0];
  *fir_ms = ctx->evAccumMs[1];
  *launches = ctx->evAccumN;
  ctx->evAccumMs[0] = ctx->evAccumMs[1] = 0;
  ctx->evAccumN = 0;
  return MMGPU_OK;
}
static int timingEvent(mmgpu_ctx* ctx)
{
  if (!ctx->timing) return MMGPU_OK;
  if (ctx->evUsed >= 3 * 1024 && ctx->evUsed % 3 == 0) { int rc = drainTiming(ctx); if (rc) return rc; }
  if (ctx->evUsed >= ctx->evPool.size())
  {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    ctx->evPool.push_back(e);
  }
  CU(cudaEventRecord(ctx->evPool[ctx->evUsed++], ctx->stream));
  return MMGPU_OK;
}

extern "C" void* mmgpu_stream(mmgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t mmgpu_launch_count(const mmgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- reference pictures --------------------------------------------------------------------------
static int compDims(const mmgpu_ctx* ctx, int c, int* w, int* h, int* mx, int* my)
{
  const int sx = scaleX(c, ctx->cfg.chroma_fmt), sy = scaleY(c, ctx->cfg.chroma_fmt);
  *w = ctx->cfg.pic_w >> sx;
  *h = ctx->cfg.pic_h >> sy;
  *mx = ctx->cfg.margin >> sx;
  *my = ctx->cfg.margin >> sy;
  return 0;
}
static int numComps(const mmgpu_ctx* ctx) { return ctx->cfg.chroma_fmt == 0 ? 1 : 3; }

static int ensureOwnedPlane(mmgpu_ctx* ctx, int slot, int c, int16_t** plane00, int* pitch)
{
  int w, h, mx, my;
  compDims(ctx, c, &w, &h, &mx, &my);
  const int p = (w + 2 * mx + 63) / 64 * 64;  // pitch in samples: 128-byte multiple
  int16_t*& base = ctx->ownedPlanes[(size_t)slot * 3 + c];
  if (!base) CU(cudaMalloc(&base, sizeof(int16_t) * (size_t)p * (h + 2 * my)));
  *plane00 = base + (size_t)my * p + mx;
  *pitch = p;
  return MMGPU_OK;
}
typedef CUresult (*MmEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// TMA descriptors of a slot's padded luma plane (one per box shape).  Needs a 16-byte aligned allocation and pitch;
// otherwise the slot simply has none (r->tma = 0) and k_fir gathers from global memory.
static int makeTmaps(mmgpu_ctx* ctx, int slot, MmRef* r)
{
  r->tma = 0;
  if (!ctx->encodeTiled || getenv("MMGPU_NO_TMA")) return MMGPU_OK;
  int w, h, mx, my;
  compDims(ctx, 0, &w, &h, &mx, &my);
  const int16_t* base = r->plane00[0] - (ptrdiff_t)my * r->stride[0] - mx;
  if (((uintptr_t)base & 15) || ((size_t)r->stride[0] * 2) % 16) return MMGPU_OK;
  const uint32_t dims[MM_TMA_TYPES][2] = MM_TMA_BOX_DIMS;
  CUtensorMap maps[MM_TMA_TYPES];
  for (int t = 0; t < MM_TMA_TYPES; ++t)
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)(w + 2 * mx), (cuuint64_t)(h + 2 * my)};
    const cuuint64_t gstr[1] = {(cuuint64_t)r->stride[0] * 2};
    const cuuint32_t box[2] = {dims[t][0], dims[t][1]};
    const cuuint32_t estr[2] = {1, 1};
    if (box[0] * box[1] * 2 > MM_TMA_BOX_BYTES) return fail(ctx, MMGPU_ERR_STATE, "TMA box larger than its shared-memory buffer");
    CUresult cr = ((MmEncodeTiled)ctx->encodeTiled)(&maps[t], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)base, gdim, gstr, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return MMGPU_OK;   // not encodable: gather path
  }
  // ordered behind the kernels already queued on the context's stream (a batch in flight keeps the descriptors it was
  // launched with); `maps` is pageable memory: the call returns once it has been staged
  CU(cudaMemcpyAsync(ctx->dTmaps + (size_t)slot * MM_TMA_TYPES, maps, sizeof(maps), cudaMemcpyHostToDevice, ctx->stream));
  r->tma = 1;
  return MMGPU_OK;
}

static int publishRef(mmgpu_ctx* ctx, int slot)
{
  int rc = makeTmaps(ctx, slot, &ctx->hRefs[slot]);
  if (rc) return rc;
  // stream-ordered like the descriptors; hRefs[slot] is pageable memory (staged before the call returns), so the
  // registration does not stall the host behind the kernels of the previous frame
  CU(cudaMemcpyAsync(ctx->dRefs + slot, &ctx->hRefs[slot], sizeof(MmRef), cudaMemcpyHostToDevice, ctx->stream));
  return MMGPU_OK;
}

extern "C" int mmgpu_set_ref(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane00[3], const int32_t stride[3])
{
  if (!ctx || !plane00 || !stride) return MMGPU_ERR_ARG;
  if (slot < 0 || slot >= ctx->cfg.max_refs) return fail(ctx, MMGPU_ERR_ARG, "ref slot out of range");
  cudaSetDevice(ctx->cfg.device);
  MmRef r;
  memset(&r, 0, sizeof(r));
  r.poc = poc;
  for (int c = 0; c < numComps(ctx); ++c)
  {
    if (!plane00[c]) return fail(ctx, MMGPU_ERR_ARG, "null plane");
    int w, h, mx, my, pitch;
    int16_t* d00;
    compDims(ctx, c, &w, &h, &mx, &my);
    int rc = ensureOwnedPlane(ctx, slot, c, &d00, &pitch);
    if (rc) return rc;
    const int16_t* src = plane00[c] - (ptrdiff_t)my * stride[c] - mx;
    CU(cudaMemcpy2DAsync(d00 - (size_t)my * pitch - mx, sizeof(int16_t) * pitch, src, sizeof(int16_t) * stride[c],
                         sizeof(int16_t) * (w + 2 * mx), h + 2 * my, cudaMemcpyHostToDevice, ctx->stream));
    r.plane00[c] = d00;
    r.stride[c] = pitch;
  }
  ctx->hRefs[slot] = r;
  return publishRef(ctx, slot);
}

extern "C" int mmgpu_set_ref_device(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane00[3],
                                    const int32_t stride[3])
{
  if (!ctx || !plane00 || !stride) return MMGPU_ERR_ARG;
  if (slot < 0 || slot >= ctx->cfg.max_refs) return fail(ctx, MMGPU_ERR_ARG, "ref slot out of range");
  cudaSetDevice(ctx->cfg.device);
  MmRef r;
  memset(&r, 0, sizeof(r));
  r.poc = poc;
  for (int c = 0; c < numComps(ctx); ++c)
  {
    if (!plane00[c]) { if (c == 0) return fail(ctx, MMGPU_ERR_ARG, "null luma plane"); continue; }
    if (((uintptr_t)plane00[c] & 7) || (stride[c] & 3))
      return fail(ctx, MMGPU_ERR_ARG, "device plane must be 8-byte aligned with a stride that is a multiple of 4 samples");
    r.plane00[c] = plane00[c];
    r.stride[c] = stride[c];
  }
  ctx->hRefs[slot] = r;
  return publishRef(ctx, slot);
}

static int setRefUnpadded(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane[3], const int32_t stride[3],
                          int wrap_offset)
{
  if (!ctx || !plane || !stride) return MMGPU_ERR_ARG;
  if (slot < 0 || slot >= ctx->cfg.max_refs) return fail(ctx, MMGPU_ERR_ARG, "ref slot out of range");
  if (wrap_offset < 0 || wrap_offset > ctx->cfg.pic_w) return fail(ctx, MMGPU_ERR_ARG, "wrap_offset must be 0..pic_w");
  cudaSetDevice(ctx->cfg.device);
  MmRef r;
  memset(&r, 0, sizeof(r));
  r.poc = poc;
  for (int c = 0; c < numComps(ctx); ++c)
  {
    if (!plane[c]) return fail(ctx, MMGPU_ERR_ARG, "null plane");
    int w, h, mx, my, pitch;
    int16_t* d00;
    compDims(ctx, c, &w, &h, &mx, &my);
    int rc = ensureOwnedPlane(ctx, slot, c, &d00, &pitch);
    if (rc) return rc;
    CU(cudaMemcpy2DAsync(d00, sizeof(int16_t) * pitch, plane[c], sizeof(int16_t) * stride[c], sizeof(int16_t) * w, h,
                         cudaMemcpyHostToDevice, ctx->stream));
    mmk_extend_border(ctx->stream, d00, pitch, w, h, mx, my, wrap_offset >> scaleX(c, ctx->cfg.chroma_fmt));
    ctx->launches++;
    r.plane00[c] = d00;
    r.stride[c] = pitch;
  }
  CU(cudaGetLastError());
  ctx->hRefs[slot] = r;
  return publishRef(ctx, slot);
}

extern "C" int mmgpu_set_ref_unpadded(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane[3],
                                      const int32_t stride[3])
{
  return setRefUnpadded(ctx, slot, poc, plane, stride, 0);
}

extern "C" int mmgpu_set_ref_unpadded_wrap(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const plane[3],
                                           const int32_t stride[3], int wrap_offset)
{
  return setRefUnpadded(ctx, slot, poc, plane, stride, wrap_offset);
}

extern "C" int mmgpu_set_ref_unpadded_device(mmgpu_ctx* ctx, int slot, int poc, const int16_t* const d_plane[3],
                                             const int32_t stride[3], int wrap_offset, void* stream)
{
  if (!ctx || !d_plane || !stride) return MMGPU_ERR_ARG;
  if (slot < 0 || slot >= ctx->cfg.max_refs) return fail(ctx, MMGPU_ERR_ARG, "ref slot out of range");
  if (wrap_offset < 0 || wrap_offset > ctx->cfg.pic_w) return fail(ctx, MMGPU_ERR_ARG, "wrap_offset must be 0..pic_w");
  cudaSetDevice(ctx->cfg.device);
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  MmRef r;
  memset(&r, 0, sizeof(r));
  r.poc = poc;
  bool fresh = false;   // first use of the slot: its planes are allocated and the device tables have to learn them
  for (int c = 0; c < numComps(ctx); ++c)
  {
    if (!d_plane[c]) return fail(ctx, MMGPU_ERR_ARG, "null plane");
    int w, h, mx, my, pitch;
    int16_t* d00;
    compDims(ctx, c, &w, &h, &mx, &my);
    if (!ctx->ownedPlanes[(size_t)slot * 3 + c]) fresh = true;
    int rc = ensureOwnedPlane(ctx, slot, c, &d00, &pitch);
    if (rc) return rc;
    CU(cudaMemcpy2DAsync(d00, sizeof(int16_t) * pitch, d_plane[c], sizeof(int16_t) * stride[c], sizeof(int16_t) * w, h,
                         cudaMemcpyDeviceToDevice, st));
    mmk_extend_border(st, d00, pitch, w, h, mx, my, wrap_offset >> scaleX(c, ctx->cfg.chroma_fmt));
    ctx->launches++;
    r.plane00[c] = d00;
    r.stride[c] = pitch;
  }
  CU(cudaGetLastError());
  if (fresh || ctx->hRefs[slot].plane00[0] != r.plane00[0])
  {
    ctx->hRefs[slot] = r;
    return publishRef(ctx, slot);   // descriptor / table upload on the context's stream (addresses never change afterwards)
  }
  ctx->hRefs[slot].poc = poc;
  return MMGPU_OK;
}

static int setRefsUnpadded(mmgpu_ctx* ctx, int n, const int32_t* slots, const int32_t* pocs, const int16_t* const* d_plane,
                           const int32_t* stride, int wrap_offset, void* stream, int pullCtas)
{
  if (!ctx || !slots || !pocs || !d_plane || !stride || n <= 0) return MMGPU_ERR_ARG;
  const int nc = numComps(ctx);
  if (n * nc > MM_PAD_MAX) return fail(ctx, MMGPU_ERR_ARG, "too many pictures for one call");
  if (wrap_offset < 0 || wrap_offset > ctx->cfg.pic_w) return fail(ctx, MMGPU_ERR_ARG, "wrap_offset must be 0..pic_w");
  if ((ctx->cfg.pic_w & 15) || (ctx->cfg.margin & 15)) return fail(ctx, MMGPU_ERR_ARG, "batched padding needs picture width and margin multiples of 16");
  cudaSetDevice(ctx->cfg.device);
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  MmPadJobs J;
  memset(&J, 0, sizeof(J));
  int maxGroups = 0, maxRows = 0, k = 0;
  for (int i = 0; i < n; ++i)
  {
    const int slot = slots[i];
    if (slot < 0 || slot >= ctx->cfg.max_refs) return fail(ctx, MMGPU_ERR_ARG, "ref slot out of range");
    MmRef r;
    memset(&r, 0, sizeof(r));
    r.poc = pocs[i];
    bool fresh = false;
    for (int c = 0; c < nc; ++c, ++k)
    {
      const int16_t* src = d_plane[i * 3 + c];
      if (!src || ((uintptr_t)src & 15) || (stride[i * 3 + c] & 7)) return fail(ctx, MMGPU_ERR_ARG, "planes must be 16-byte aligned, stride a multiple of 8");
      int w, h, mx, my, pitch;
      int16_t* d00;
      compDims(ctx, c, &w, &h, &mx, &my);
      if (!ctx->ownedPlanes[(size_t)slot * 3 + c]) fresh = true;
      int rc = ensureOwnedPlane(ctx, slot, c, &d00, &pitch);
      if (rc) return rc;
      MmPadPlane& p = J.p[k];
      p.src = src; p.dst00 = d00; p.srcStride = stride[i * 3 + c]; p.dstStride = pitch; p.w = w; p.h = h; p.mx = mx; p.my = my;
      p.xoff = wrap_offset >> scaleX(c, ctx->cfg.chroma_fmt);
      maxGroups = std::max(maxGroups, (w + 2 * mx) >> 3);
      maxRows = std::max(maxRows, h + 2 * my);
      r.plane00[c] = d00;
      r.stride[c] = pitch;
    }
    if (fresh || ctx->hRefs[slot].plane00[0] != r.plane00[0])
    {
      ctx->hRefs[slot] = r;
      int rc = publishRef(ctx, slot);
      if (rc) return rc;
    }
    else ctx->hRefs[slot].poc = pocs[i];
  }
  if (pullCtas > 0)
  {
    mmk_pull_pad_planes(st, J, k, pullCtas);
    ctx->launches += 3;
  }
  else
  {
    mmk_pad_planes(st, J, k, maxGroups, maxRows);
    ctx->launches++;
  }
  CU(cudaGetLastError());
  return MMGPU_OK;
}

extern "C" int mmgpu_set_refs_unpadded_device(mmgpu_ctx* ctx, int n, const int32_t* slots, const int32_t* pocs,
                                              const int16_t* const* d_plane, const int32_t* stride, int wrap_offset, void* stream)
{
  return setRefsUnpadded(ctx, n, slots, pocs, d_plane, stride, wrap_offset, stream, 0);
}

extern "C" int mmgpu_pull_refs_unpadded(mmgpu_ctx* ctx, int n, const int32_t* slots, const int32_t* pocs, const int16_t* const* d_plane,
                                        const int32_t* stride, int max_ctas, void* stream)
{
  if (max_ctas < 1 || max_ctas > 1024) return ctx ? fail(ctx, MMGPU_ERR_ARG, "max_ctas must be 1..1024") : MMGPU_ERR_ARG;
  return setRefsUnpadded(ctx, n, slots, pocs, d_plane, stride, 0, stream, max_ctas);
}

extern "C" int mmgpu_multicast_copy(mmgpu_ctx* ctx, const void* d_src, void* d_dst_multicast, size_t bytes, int max_ctas, void* stream)
{
  if (!ctx || !d_src || !d_dst_multicast || bytes == 0) return MMGPU_ERR_ARG;
  if (((uintptr_t)d_src & 15) || ((uintptr_t)d_dst_multicast & 15) || (bytes & 15)) return fail(ctx, MMGPU_ERR_ARG, "multicast_copy: 16-byte alignment");
  if (max_ctas < 1 || max_ctas > 1024) return fail(ctx, MMGPU_ERR_ARG, "max_ctas must be 1..1024");
  cudaSetDevice(ctx->cfg.device);
  mmk_multicast_copy(stream ? (cudaStream_t)stream : ctx->stream, d_src, d_dst_multicast, bytes, max_ctas);
  ctx->launches++;
  CU(cudaGetLastError());
  return MMGPU_OK;
}

extern "C" int mmgpu_set_rotation(mmgpu_ctx* ctx, int rot_slot, const float R[9])
{
  if (!ctx || !R) return MMGPU_ERR_ARG;
  if (rot_slot < 0 || rot_slot >= MM_MAX_ROT) return fail(ctx, MMGPU_ERR_ARG, "rotation slot out of range");
  cudaSetDevice(ctx->cfg.device);
  memcpy(&ctx->hRot[9 * (MM_ROT_USER0 + rot_slot)], R, 9 * sizeof(float));
  CU(cudaMemcpyAsync(ctx->dRot + 9 * (MM_ROT_USER0 + rot_slot), R, 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

// ---- scratch -------------------------------------------------------------------------------------
static int ensurePinned(mmgpu_ctx* ctx, MmStage* st, size_t bytes)
{
  if (st->busy) { CU(cudaEventSynchronize(st->freeEv)); st->busy = false; }
  if (st->cap >= bytes) return MMGPU_OK;
  if (st->h) cudaFreeHost(st->h);
  st->h = nullptr;
  st->cap = 0;
  const size_t n = bytes + bytes / 2 + 4096;
  CU(cudaMallocHost(&st->h, n));
  st->cap = n;
  return MMGPU_OK;
}

template <typename T>
static int ensureCap(mmgpu_ctx* ctx, T** p, size_t* cap, size_t need)
{
  if (*cap >= need) return MMGPU_OK;
  size_t n = need + need / 2 + 1024;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CU(cudaMalloc(p, sizeof(T) * n));
  *cap = n;
  return MMGPU_OK;
}

// anyMultiple: MC entry points take every size that is a multiple of the sub-block (the encoder's RD search also predicts
// 12/24/48-wide blocks); the batch / Hadamard work lists index with power-of-two arithmetic
static int badBlock(mmgpu_ctx* ctx, const char* what, int comp, int fmt, int x, int y, int w, int h)
{
  char dims[96];
  snprintf(dims, sizeof(dims), " (comp %d fmt %d: %dx%d at %d,%d)", comp, fmt, w, h, x, y);
  return fail(ctx, MMGPU_ERR_ARG, std::string(what) + dims);
}
static int checkBlockShape(mmgpu_ctx* ctx, int comp, int fmt, int x, int y, int w, int h, bool anyMultiple = false)
{
  // runs per block of every batch: nothing but comparisons on the good path
  const int sx = scaleX(comp, fmt), sy = scaleY(comp, fmt);
  const int sbw = 4 >> sx, sbh = 4 >> sy;
  if (w < sbw || h < sbh || (w % sbw) || (h % sbh) || (x % sbw) || (y % sbh) || x < 0 || y < 0)
    return badBlock(ctx, "block must be aligned to the sub-block grid", comp, fmt, x, y, w, h);
  if ((!anyMultiple && (!pow2(w) || !pow2(h))) || (w << sx) > 128 || (h << sy) > 128)
    return badBlock(ctx, "block size must be a power of two <= 128", comp, fmt, x, y, w, h);
  if (((x + w) << sx) > ctx->cfg.pic_w || ((y + h) << sy) > ctx->cfg.pic_h) return badBlock(ctx, "block outside the picture", comp, fmt, x, y, w, h);
  return MMGPU_OK;
}

extern "C" int mmgpu_selftest_tan_sincos(mmgpu_ctx* ctx, uint32_t stride, uint64_t* checksum)
{
  if (!ctx || !checksum || stride == 0) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  int rc = ensureCap(ctx, &ctx->dDist, &ctx->capDist, 1);
  if (rc) return rc;
  CU(cudaMemsetAsync(ctx->dDist, 0, sizeof(unsigned long long), ctx->stream));
  mmk_tan_selftest(ctx->stream, stride, ctx->dDist);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(checksum, ctx->dDist, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

// ---- a1 ------------------------------------------------------------------------------------------
extern "C" int mmgpu_reproject(mmgpu_ctx* ctx, int comp, int chroma_fmt, int x, int y, int w, int h, mmgpu_mv mv, int model,
                               int rot_slot, int32_t* out_x, int32_t* out_y)
{
  if (!ctx || !out_x || !out_y) return MMGPU_ERR_ARG;
  if (model <= MMGPU_CLASSIC || model > MMGPU_GEODESIC_CAMPOSE) return fail(ctx, MMGPU_ERR_ARG, "reproject: model must be a non-classic motion model");
  if (comp < 0 || comp > 2 || rot_slot < 0 || rot_slot >= MM_MAX_ROT) return fail(ctx, MMGPU_ERR_ARG, "bad comp/rot_slot");
  if (comp != 0 && chroma_fmt != ctx->cfg.chroma_fmt) return fail(ctx, MMGPU_ERR_ARG, "chroma_fmt differs from the context's");
  int rc = checkBlockShape(ctx, comp, chroma_fmt, x, y, w, h);
  if (rc) return rc;
  cudaSetDevice(ctx->cfg.device);
  const int csx = scaleX(comp, chroma_fmt), csy = scaleY(comp, chroma_fmt);
  const int n = (w / (4 >> csx)) * (h / (4 >> csy));
  rc = ensureCap(ctx, &ctx->dTmp, &ctx->capTmp, (size_t)n * 4 + 64);
  if (rc) return rc;
  int32_t* dx = reinterpret_cast<int32_t*>(ctx->dTmp);
  int32_t* dy = dx + n;
  MmParams P = baseParams(ctx);
  mmk_reproject(ctx->stream, P, comp, csx, csy, x, y, w, h, mv, model, rot_slot, dx, dy);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out_x, dx, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(out_y, dy, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

// ---- 8f-1: MM-MVP conversion (host) ---------------------------------------------------------------------------
extern "C" int mmgpu_mvp_convert(mmgpu_ctx* ctx, int pos_x, int pos_y, mmgpu_mv mv_orig, int model_orig, int model_desired,
                                 int shift_hor, int shift_ver, int rot_slot_orig, int rot_slot_desired, int cand_x, int cand_y,
                                 int cand_w, int cand_h, int cur_x, int cur_y, int cur_w, int cur_h, mmgpu_mv* out)
{
  if (!ctx || !out) return MMGPU_ERR_ARG;
  // the reference's two early returns come first: callers pass INVALID (-1) / NUM_MODELS for unavailable candidates with a
  // zero MV or equal models (MVReprojection.cpp:174-184)
  if (mv_orig.hor == 0 && mv_orig.ver == 0) { out->hor = 0; out->ver = 0; return MMGPU_OK; }
  if (model_desired == model_orig && (model_desired != MMGPU_GEODESIC_CAMPOSE || rot_slot_orig == rot_slot_desired)) { *out = mv_orig; return MMGPU_OK; }
  if (model_orig < MMGPU_CLASSIC || model_orig > MMGPU_GEODESIC_CAMPOSE || model_desired < MMGPU_CLASSIC || model_desired > MMGPU_GEODESIC_CAMPOSE)
    return fail(ctx, MMGPU_ERR_ARG, "bad model");
  if (shift_hor < 0 || shift_hor > 8 || shift_ver < 0 || shift_ver > 8) return fail(ctx, MMGPU_ERR_ARG, "bad shift");
  if (rot_slot_orig < 0 || rot_slot_orig >= MM_MAX_ROT || rot_slot_desired < 0 || rot_slot_desired >= MM_MAX_ROT) return fail(ctx, MMGPU_ERR_ARG, "bad rot_slot");
  auto rotOf = [&](int model, int slot) -> const float* {
    const int s = (model == MMGPU_GEODESIC_CAMPOSE) ? (MM_ROT_USER0 + slot) : (model - MMGPU_GEODESIC_X);
    return &ctx->hRot[9 * (s < 0 ? 0 : s)];
  };
  int h = 0, v = 0;
  mm_mvp_convert(ctx->g, pos_x, pos_y, mv_orig.hor, mv_orig.ver, model_orig, model_desired, shift_hor, shift_ver, rotOf(model_orig, rot_slot_orig),
                 rotOf(model_desired, rot_slot_desired), rot_slot_orig == rot_slot_desired, cand_x, cand_y, cand_w, cand_h, cur_x, cur_y, cur_w,
                 cur_h, &h, &v);
  out->hor = h;
  out->ver = v;
  return MMGPU_OK;
}

// ---- a5/a6/a8 ------------------------------------------------------------------------------------
extern "C" int mmgpu_mc(mmgpu_ctx* ctx, int ref_slot, int comp, int chroma_fmt, int x, int y, int w, int h, mmgpu_mv mv,
                        int model, int rot_slot, int bi, int clp_min, int clp_max, int oob_mode, int16_t* dst, int dst_stride)
{
  if (!ctx || !dst) return MMGPU_ERR_ARG;
  if (model < MMGPU_CLASSIC || model > MMGPU_GEODESIC_CAMPOSE) return fail(ctx, MMGPU_ERR_ARG, "bad model");
  if (comp < 0 || comp > 2 || rot_slot < 0 || rot_slot >= MM_MAX_ROT) return fail(ctx, MMGPU_ERR_ARG, "bad comp/rot_slot");
  if (comp != 0 && chroma_fmt != ctx->cfg.chroma_fmt) return fail(ctx, MMGPU_ERR_ARG, "chroma_fmt differs from the context's");
  if (ref_slot < 0 || ref_slot >= ctx->cfg.max_refs || !ctx->hRefs[ref_slot].plane00[comp])
    return fail(ctx, MMGPU_ERR_STATE, "reference slot not set");
  int rc = checkBlockShape(ctx, comp, chroma_fmt, x, y, w, h, true);
  if (rc) return rc;
  cudaSetDevice(ctx->cfg.device);
  const int csx = scaleX(comp, chroma_fmt), csy = scaleY(comp, chroma_fmt);
  rc = ensureCap(ctx, &ctx->dTmp, &ctx->capTmp, (size_t)w * h + 64);
  if (rc) return rc;
  MmParams P = baseParams(ctx);
  P.clpMin = clp_min;
  P.clpMax = clp_max;
  P.oobMode = oob_mode;
  mmk_mc_generic(ctx->stream, P, ref_slot, comp, csx, csy, x, y, w, h, mv, model, rot_slot, bi, ctx->dTmp, w);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy2DAsync(dst, sizeof(int16_t) * dst_stride, ctx->dTmp, sizeof(int16_t) * w, sizeof(int16_t) * w, h,
                       cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

extern "C" int mmgpu_mc_bdof(mmgpu_ctx* ctx, int ref_slot, int x, int y, int w, int h, mmgpu_mv mv, int clp_min, int clp_max, int16_t* dst)
{
  if (!ctx || !dst) return MMGPU_ERR_ARG;
  if (ref_slot < 0 || ref_slot >= ctx->cfg.max_refs || !ctx->hRefs[ref_slot].plane00[0]) return fail(ctx, MMGPU_ERR_STATE, "reference slot not set");
  int rc = checkBlockShape(ctx, 0, 0, x, y, w, h);
  if (rc) return rc;
  // the ring reads one sample beyond the block at integer MV offsets: same clipMv contract as xPredInterBlk
  const int mg = ctx->cfg.margin - 16;
  const int x0 = x + (mv.hor >> 4), y0 = y + (mv.ver >> 4);
  if (x0 < -mg || y0 < -mg || x0 + w > ctx->cfg.pic_w + mg || y0 + h > ctx->cfg.pic_h + mg)
    return fail(ctx, MMGPU_ERR_ARG, "mc_bdof: MV not clipped to the padded picture (clipMv)");
  cudaSetDevice(ctx->cfg.device);
  const int stride = w + 4, n = stride * (h + 4);
  if ((rc = ensureCap(ctx, &ctx->dTmp, &ctx->capTmp, (size_t)n + 64))) return rc;
  MmParams P = baseParams(ctx);
  P.clpMin = clp_min;
  P.clpMax = clp_max;
  P.oobMode = MMGPU_OOB_PRED;
  CU(cudaMemsetAsync(ctx->dTmp, 0, sizeof(int16_t) * n, ctx->stream));
  mmk_mc_generic(ctx->stream, P, ref_slot, 0, 0, 0, x, y, w, h, mv, MMGPU_CLASSIC, 0, 1, ctx->dTmp + 2 * stride + 2, stride);
  const MmRef& r = ctx->hRefs[ref_slot];
  mmk_bdof_ring(ctx->stream, r.plane00[0], r.stride[0], x0, y0, ((mv.hor & 15) < 8) ? 1 : 0, ((mv.ver & 15) < 8) ? 1 : 0, w, h, ctx->fir.cpShift,
                ctx->dTmp);
  ctx->launches += 2;
  CU(cudaGetLastError());
  // only the ring box (rows 1 .. h+2, columns 1 .. w+2) is defined by the reference: copy exactly that
  CU(cudaMemcpy2DAsync(dst + stride + 1, sizeof(int16_t) * stride, ctx->dTmp + stride + 1, sizeof(int16_t) * stride, sizeof(int16_t) * (w + 2), h + 2,
                       cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

extern "C" int mmgpu_add_avg(mmgpu_ctx* ctx, const int16_t* src0, int s0_stride, const int16_t* src1, int s1_stride,
                             int16_t* dst, int dst_stride, int w, int h, int clp_min, int clp_max)
{
  if (!ctx || !src0 || !src1 || !dst || w <= 0 || h <= 0) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  int rc = ensureCap(ctx, &ctx->dTmp, &ctx->capTmp, (size_t)w * h * 3 + 64);
  if (rc) return rc;
  int16_t *d0 = ctx->dTmp, *d1 = d0 + (size_t)w * h, *dd = d1 + (size_t)w * h;
  CU(cudaMemcpy2DAsync(d0, 2 * w, src0, 2 * s0_stride, 2 * w, h, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpy2DAsync(d1, 2 * w, src1, 2 * s1_stride, 2 * w, h, cudaMemcpyHostToDevice, ctx->stream));
  const int hr = ctx->fir.cpShift;
  const int shift = hr + 1, offset = (1 << (shift - 1)) + 2 * 8192;
  mmk_add_avg(ctx->stream, d0, w, d1, w, dd, w, w, h, shift, offset, clp_min, clp_max);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy2DAsync(dst, 2 * dst_stride, dd, 2 * w, 2 * w, h, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

// ---- org -----------------------------------------------------------------------------------------
extern "C" int mmgpu_set_org(mmgpu_ctx* ctx, const int16_t* org, int stride, int w, int h)
{
  if (!ctx || !org || w <= 0 || h <= 0 || stride < w) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  const int pitch = (w + 63) / 64 * 64;
  int rc = ensureCap(ctx, &ctx->dOrgOwned, &ctx->orgCap, (size_t)pitch * h);
  if (rc) return rc;
  CU(cudaMemcpy2DAsync(ctx->dOrgOwned, 2 * pitch, org, 2 * (size_t)stride, 2 * (size_t)w, h, cudaMemcpyHostToDevice, ctx->stream));
  ctx->dOrg = ctx->dOrgOwned;
  ctx->orgStride = pitch;
  ctx->orgW = w;
  ctx->orgH = h;
  return MMGPU_OK;
}
extern "C" int mmgpu_set_org_device(mmgpu_ctx* ctx, const int16_t* org, int stride, int w, int h)
{
  if (!ctx || !org || w <= 0 || h <= 0 || stride < w) return MMGPU_ERR_ARG;
  if (((uintptr_t)org & 7) || (stride & 3)) return fail(ctx, MMGPU_ERR_ARG, "device org must be 8-byte aligned, stride a multiple of 4");
  ctx->dOrg = org;
  ctx->orgStride = stride;
  ctx->orgW = w;
  ctx->orgH = h;
  return MMGPU_OK;
}

// ---- plans ---------------------------------------------------------------------------------------
// Host part of a plan: validate the blocks, count the warps of work of each and lay them out grouped by motion model
// (so the warps resident on an SM run the same model's code).  The work list itself and the candidate->block map are
// written by k_build_items on the device; nothing here is O(candidates).
// Copies go to `copyStream`; the device-side work list is built on `execStream` (which waits for the copies through
// `copied` when the two differ).  orgW/orgH >= 0: also check the org rectangles against that plane.
static int buildPlan(mmgpu_ctx* ctx, mmgpu_plan* p, int n_blocks, const mmgpu_block* blocks, int n_cands, int orgW, int orgH,
                     MmStage* st, cudaStream_t copyStream, cudaStream_t execStream, cudaEvent_t copied)
{
  int rc;
  const bool needOrg = orgW >= 0;
  if ((rc = ensurePinned(ctx, st, (size_t)n_blocks * sizeof(int32_t) * 2))) return rc;
  int32_t* itemBase = reinterpret_cast<int32_t*>(st->h);
  int32_t* order = itemBase + n_blocks;
  long modelItems[MM_NUM_MODELS] = {0};
  uint64_t samples = 0;
  long candSum = 0;
  int orgMaxX = 0, orgMaxY = 0, maxCand = 0;
  for (int bi = 0; bi < n_blocks; ++bi)
  {
    const mmgpu_block& b = blocks[bi];
    if ((rc = checkBlockShape(ctx, 0, 0, b.x, b.y, b.w, b.h))) return rc;
    if (b.model < MMGPU_CLASSIC || b.model > MMGPU_GEODESIC_CAMPOSE) return fail(ctx, MMGPU_ERR_ARG, "bad model in block");
    if (b.rot_slot < 0 || b.rot_slot >= MM_MAX_ROT) return fail(ctx, MMGPU_ERR_ARG, "bad rot_slot in block");
    if (b.ref_slot < 0 || b.ref_slot >= ctx->cfg.max_refs || !ctx->hRefs[b.ref_slot].plane00[0])
      return fail(ctx, MMGPU_ERR_STATE, "block references an unset reference slot");
    if (b.cand_count < 0 || b.cand_first < 0 || (long)b.cand_first + b.cand_count > n_cands)
      return fail(ctx, MMGPU_ERR_ARG, "candidate range outside the mv array");
    if (b.org_x < 0 || b.org_y < 0 || b.org_x > 0xffff || b.org_y > 0xffff) return fail(ctx, MMGPU_ERR_ARG, "org_x / org_y must be in 0..65535");
    if (b.org_x & 3) return fail(ctx, MMGPU_ERR_ARG, "org_x must be a multiple of 4 (8-byte loads of the original block)");
    if (needOrg && (b.org_x + b.w > orgW || b.org_y + b.h > orgH))
      return fail(ctx, MMGPU_ERR_ARG, "org block outside the org plane");
    orgMaxX = std::max(orgMaxX, b.org_x + b.w);
    orgMaxY = std::max(orgMaxY, b.org_y + b.h);
    maxCand = std::max(maxCand, b.cand_count);
    const int n = (b.w >> 2) * (b.h >> 2);
    const int nItems = (n >= 32) ? b.cand_count * (n >> 5) : (b.cand_count + (32 / n) - 1) / (32 / n);
    itemBase[bi] = nItems;
    order[bi] = bi;
    modelItems[b.model] += nItems;
    candSum += b.cand_count;
    samples += (uint64_t)b.w * b.h * b.cand_count;
  }
  // candidate ranges must tile [0, n_cands): equal total + no overlap between neighbours in cand_first order
  if (candSum != n_cands) return fail(ctx, MMGPU_ERR_ARG, "candidate ranges of the blocks do not cover the mv array exactly once");
  bool ordered = true;   // the usual case: blocks listed in candidate order -> no sort on the per-call path
  for (int i = 1; i < n_blocks && ordered; ++i) ordered = blocks[i - 1].cand_first <= blocks[i].cand_first;
  if (!ordered)
    std::sort(order, order + n_blocks, [blocks](int32_t a, int32_t b) { return blocks[a].cand_first < blocks[b].cand_first; });
  long expect = 0;
  for (int i = 0; i < n_blocks; ++i)
  {
    if (blocks[order[i]].cand_first != expect) return fail(ctx, MMGPU_ERR_ARG, "candidate ranges of blocks overlap or leave gaps");
    expect += blocks[order[i]].cand_count;
  }
  long modelBase[MM_NUM_MODELS], total = 0;
  for (int m = 0; m < MM_NUM_MODELS; ++m)
  {
    modelBase[m] = total;
    p->modelFirst[m] = (int)total;
    p->modelCount[m] = (int)modelItems[m];
    total += modelItems[m];
  }
  if (total > 0x7fffffffL) return fail(ctx, MMGPU_ERR_ARG, "batch too large");
  for (int bi = 0; bi < n_blocks; ++bi)
  {
    const int cnt = itemBase[bi];
    itemBase[bi] = (int32_t)modelBase[blocks[bi].model];
    modelBase[blocks[bi].model] += cnt;
  }
  if ((rc = ensureCap(ctx, &p->dBlocks, &p->capBlocks, (size_t)n_blocks))) return rc;
  if ((rc = ensureCap(ctx, &p->dItemBase, &p->capItemBase, (size_t)n_blocks))) return rc;
  if ((rc = ensureCap(ctx, &p->dItems, &p->capItems, (size_t)total))) return rc;
  if ((rc = ensureCap(ctx, &p->dCandBlock, &p->capCands, (size_t)n_cands))) return rc;
  if ((rc = ensureCap(ctx, &p->dCands, &p->capCandParams, (size_t)n_cands))) return rc;
  if ((rc = ensureCap(ctx, &p->dPos, &p->capPos, (size_t)total * 32))) return rc;
  CU(cudaMemcpyAsync(p->dBlocks, blocks, sizeof(mmgpu_block) * n_blocks, cudaMemcpyHostToDevice, copyStream));
  CU(cudaMemcpyAsync(p->dItemBase, itemBase, sizeof(int32_t) * n_blocks, cudaMemcpyHostToDevice, copyStream));
  CU(cudaEventRecord(st->freeEv, copyStream));  // the pinned staging buffer may be rewritten after this point
  st->busy = true;
  if (copyStream != execStream)
  {
    CU(cudaEventRecord(copied, copyStream));
    CU(cudaStreamWaitEvent(execStream, copied, 0));
  }
  mmk_build_items(execStream, p->dBlocks, p->dItemBase, n_blocks, p->dItems, p->dCandBlock);
  ctx->launches++;
  CU(cudaGetLastError());
  p->ctx = ctx;
  p->nBlocks = n_blocks;
  p->nCands = n_cands;
  p->nItems = (int)total;
  p->samples = samples;
  p->orgMaxX = orgMaxX;
  p->orgMaxY = orgMaxY;
  p->maxCandCount = maxCand;
  return MMGPU_OK;
}

extern "C" int mmgpu_plan_create(mmgpu_ctx* ctx, int n_blocks, const mmgpu_block* blocks, int n_cands, mmgpu_plan** out)
{
  if (!ctx || !blocks || !out || n_blocks <= 0 || n_cands <= 0) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  mmgpu_plan* p = new mmgpu_plan();
  memset(p, 0, sizeof(*p));
  int rc = buildPlan(ctx, p, n_blocks, blocks, n_cands, -1, -1, &ctx->stage, ctx->stream, ctx->stream, nullptr);
  if (rc) { freePlan(p); delete p; return rc; }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { freePlan(p); delete p; return fail(ctx, MMGPU_ERR_CUDA, "plan upload failed"); }
  *out = p;
  return MMGPU_OK;
}
extern "C" void mmgpu_plan_destroy(mmgpu_plan* p)
{
  if (!p) return;
  if (p->ctx) cudaSetDevice(p->ctx->cfg.device);
  freePlan(p);
  delete p;
}
extern "C" uint64_t mmgpu_plan_samples(const mmgpu_plan* p) { return p ? p->samples : 0; }

static int launchMe(mmgpu_ctx* ctx, const mmgpu_plan* plan, const mmgpu_mv* d_mvs, int mv_prec_shift, int mode, int dist_kind,
                    int sub_shift, int clp_min, int clp_max, unsigned long long* d_dist, int16_t* d_dst, int dst_stride,
                    const int16_t* d_org = nullptr, int org_stride = 0, cudaStream_t setupStream = nullptr,
                    cudaEvent_t setupDone = nullptr)
{
  // setupStream != nullptr (pipelined submit): k_cand_setup runs there, behind the slot's k_build_items, so that both
  // overlap the previous batch's k_fir; the main stream joins through setupDone before k_map.
  if (mv_prec_shift < 0 || mv_prec_shift > 4) return fail(ctx, MMGPU_ERR_ARG, "mv_prec_shift must be 0..4");
  MmParams P = baseParams(ctx);
  if (d_org) { P.org = d_org; P.orgStride = org_stride; }
  P.blocks = plan->dBlocks;
  P.items = plan->dItems;
  P.nItems = plan->nItems;
  P.mvs = d_mvs;
  P.mvShift = mv_prec_shift;
  P.cands = plan->dCands;
  P.clpMin = clp_min;
  P.clpMax = clp_max;
  P.distKind = dist_kind;
  P.subShift = sub_shift;
  P.dist = d_dist;
  P.dst = d_dst;
  P.dstStride = dst_stride;
  mmk_cand_setup(setupStream ? setupStream : ctx->stream, ctx->g, plan->dBlocks, plan->dCandBlock, d_mvs, mv_prec_shift, ctx->dRot, -1,
                 plan->nCands, plan->dCands, (mode == MM_MODE_DIST) ? d_dist : nullptr);
  ctx->launches++;
  if (setupStream)
  {
    CU(cudaEventRecord(setupDone, setupStream));
    CU(cudaStreamWaitEvent(ctx->stream, setupDone, 0));
  }
  int rc;
  P.oobMode = (mode == MM_MODE_DIST) ? ctx->distOob : MMGPU_OOB_PRED;
  if ((rc = timingEvent(ctx))) return rc;
  // k_map: positions of all entries.  Work items are grouped by motion model; ctx->mapPerModel launches one
  // specialised kernel per model present, otherwise one generic launch covers all items.
  if (ctx->mapPerModel)
  {
    for (int m = 0; m < MM_NUM_MODELS; ++m)
      if (plan->modelCount[m])
      {
        mmk_map(ctx->stream, P, plan->modelFirst[m], plan->modelCount[m], m, plan->dPos);
        ctx->launches++;
      }
  }
  else
  {
    mmk_map(ctx->stream, P, 0, plan->nItems, -1, plan->dPos);
    ctx->launches++;
  }
  if ((rc = timingEvent(ctx))) return rc;
  mmk_fir(ctx->stream, P, 0, plan->nItems, mode, plan->dPos);
  ctx->launches++;
  if ((rc = timingEvent(ctx))) return rc;
  CU(cudaGetLastError());
  return MMGPU_OK;
}

extern "C" int mmgpu_set_dist_oob_mode(mmgpu_ctx* ctx, int oob_mode)
{
  if (!ctx) return MMGPU_ERR_ARG;
  if (oob_mode != MMGPU_OOB_ME && oob_mode != MMGPU_OOB_PRED) return fail(ctx, MMGPU_ERR_ARG, "bad oob_mode");
  ctx->distOob = oob_mode;
  return MMGPU_OK;
}

extern "C" int mmgpu_me_batch_device(mmgpu_ctx* ctx, const mmgpu_plan* plan, const mmgpu_mv* d_mvs, int mv_prec_shift,
                                     int dist_kind, int sub_shift, int clp_min, int clp_max, uint64_t* d_dist)
{
  if (!ctx || !plan || !d_mvs || !d_dist) return MMGPU_ERR_ARG;
  if (!ctx->dOrg) return fail(ctx, MMGPU_ERR_STATE, "mmgpu_set_org* has not been called");
  if (dist_kind != MMGPU_DIST_SAD && dist_kind != MMGPU_DIST_HAD) return fail(ctx, MMGPU_ERR_ARG, "bad dist_kind");
  if (sub_shift < 0 || sub_shift > 2 || (dist_kind == MMGPU_DIST_HAD && sub_shift != 0)) return fail(ctx, MMGPU_ERR_ARG, "bad sub_shift");
  if (plan->orgMaxX > ctx->orgW || plan->orgMaxY > ctx->orgH)
    return fail(ctx, MMGPU_ERR_ARG, "a block of the plan lies outside the current org plane (mmgpu_set_org*)");
  cudaSetDevice(ctx->cfg.device);
  return launchMe(ctx, plan, d_mvs, mv_prec_shift, MM_MODE_DIST, dist_kind, sub_shift, clp_min, clp_max,
                  reinterpret_cast<unsigned long long*>(d_dist), nullptr, 0);
}

extern "C" int mmgpu_me_batch_device_multiref(mmgpu_ctx* ctx, const mmgpu_plan* plan, const mmgpu_mv* d_mvs, int mv_prec_shift,
                                              int dist_kind, int sub_shift, int clp_min, int clp_max, int n_refs,
                                              const int32_t* ref_slots, const int32_t* rot_slots, uint64_t* d_dist)
{
  if (!ctx || !plan || !d_mvs || !d_dist || !ref_slots || n_refs <= 0) return MMGPU_ERR_ARG;
  if (!ctx->dOrg) return fail(ctx, MMGPU_ERR_STATE, "mmgpu_set_org* has not been called");
  if (dist_kind != MMGPU_DIST_SAD && dist_kind != MMGPU_DIST_HAD) return fail(ctx, MMGPU_ERR_ARG, "bad dist_kind");
  if (sub_shift < 0 || sub_shift > 2 || (dist_kind == MMGPU_DIST_HAD && sub_shift != 0)) return fail(ctx, MMGPU_ERR_ARG, "bad sub_shift");
  if (mv_prec_shift < 0 || mv_prec_shift > 4) return fail(ctx, MMGPU_ERR_ARG, "mv_prec_shift must be 0..4");
  if (plan->orgMaxX > ctx->orgW || plan->orgMaxY > ctx->orgH)
    return fail(ctx, MMGPU_ERR_ARG, "a block of the plan lies outside the current org plane (mmgpu_set_org*)");
  for (int r = 0; r < n_refs; ++r)
  {
    if (ref_slots[r] < 0 || ref_slots[r] >= ctx->cfg.max_refs || !ctx->hRefs[ref_slots[r]].plane00[0])
      return fail(ctx, MMGPU_ERR_STATE, "multiref: unset reference slot");
    if (rot_slots && (rot_slots[r] < 0 || rot_slots[r] >= MM_MAX_ROT)) return fail(ctx, MMGPU_ERR_ARG, "multiref: bad rot_slot");
  }
  cudaSetDevice(ctx->cfg.device);
  unsigned long long* dist = reinterpret_cast<unsigned long long*>(d_dist);
  CU(cudaMemsetAsync(dist, 0, sizeof(unsigned long long) * (size_t)n_refs * plan->nCands, ctx->stream));
  MmParams P = baseParams(ctx);
  P.blocks = plan->dBlocks; P.items = plan->dItems; P.nItems = plan->nItems; P.mvs = d_mvs; P.mvShift = mv_prec_shift;
  P.cands = plan->dCands; P.clpMin = clp_min; P.clpMax = clp_max; P.distKind = dist_kind; P.subShift = sub_shift;
  P.oobMode = ctx->distOob;
  // the mapping depends on the reference only through the rotation slot of GEODESIC_CAMPOSE blocks: one k_cand_setup + k_map
  // per distinct rotation slot, one k_fir per reference
  const bool campose = plan->modelCount[MMGPU_GEODESIC_CAMPOSE] > 0 && rot_slots != nullptr;
  std::vector<char> done((size_t)n_refs, 0);
  int rc;
  for (int r = 0; r < n_refs; ++r)
  {
    if (done[r]) continue;
    const int rot = rot_slots ? rot_slots[r] : -1;
    mmk_cand_setup(ctx->stream, ctx->g, plan->dBlocks, plan->dCandBlock, d_mvs, mv_prec_shift, ctx->dRot, rot, plan->nCands, plan->dCands, nullptr);
    ctx->launches++;
    P.rotOverride = rot;
    if ((rc = timingEvent(ctx))) return rc;
    mmk_map(ctx->stream, P, 0, plan->nItems, -1, plan->dPos);
    ctx->launches++;
    if ((rc = timingEvent(ctx))) return rc;
    for (int q = r; q < n_refs; ++q)
    {
      if (done[q] || (campose && rot_slots[q] != rot)) continue;
      P.refOverride = ref_slots[q];
      P.dist = dist + (size_t)q * plan->nCands;
      mmk_fir(ctx->stream, P, 0, plan->nItems, MM_MODE_DIST, plan->dPos);
      ctx->launches++;
      done[q] = 1;
    }
    if ((rc = timingEvent(ctx))) return rc;
  }
  CU(cudaGetLastError());
  return MMGPU_OK;
}

extern "C" int mmgpu_mc_batch_device(mmgpu_ctx* ctx, const mmgpu_plan* plan, const mmgpu_mv* d_mvs, int mv_prec_shift,
                                     int clp_min, int clp_max, int16_t* d_dst, int dst_stride)
{
  if (!ctx || !plan || !d_mvs || !d_dst) return MMGPU_ERR_ARG;
  if (((uintptr_t)d_dst & 7) || (dst_stride & 3)) return fail(ctx, MMGPU_ERR_ARG, "dst must be 8-byte aligned, stride a multiple of 4");
  if (plan->maxCandCount != 1)
    return fail(ctx, MMGPU_ERR_ARG, "mmgpu_mc_batch_device wants exactly one MV per block (every candidate writes the block's destination)");
  cudaSetDevice(ctx->cfg.device);
  return launchMe(ctx, plan, d_mvs, mv_prec_shift, MM_MODE_PRED, 0, 0, clp_min, clp_max, nullptr, d_dst, dst_stride);
}

extern "C" int mmgpu_me_batch(mmgpu_ctx* ctx, int n_blocks, const mmgpu_block* blocks, int n_cands, const mmgpu_mv* mvs,
                              int mv_prec_shift, int dist_kind, int sub_shift, int clp_min, int clp_max, uint64_t* dist)
{
  if (!ctx || !blocks || !mvs || !dist || n_blocks <= 0 || n_cands <= 0) return MMGPU_ERR_ARG;
  if (!ctx->dOrg) return fail(ctx, MMGPU_ERR_STATE, "mmgpu_set_org* has not been called");
  cudaSetDevice(ctx->cfg.device);
  int rc = buildPlan(ctx, &ctx->scratchPlan, n_blocks, blocks, n_cands, ctx->orgW, ctx->orgH, &ctx->stage, ctx->stream, ctx->stream,
                     nullptr);
  if (rc) return rc;
  if ((rc = ensureCap(ctx, &ctx->dMvs, &ctx->capMvs, (size_t)n_cands))) return rc;
  if ((rc = ensureCap(ctx, &ctx->dDist, &ctx->capDist, (size_t)n_cands))) return rc;
  CU(cudaMemcpyAsync(ctx->dMvs, mvs, sizeof(mmgpu_mv) * n_cands, cudaMemcpyHostToDevice, ctx->stream));
  rc = mmgpu_me_batch_device(ctx, &ctx->scratchPlan, ctx->dMvs, mv_prec_shift, dist_kind, sub_shift, clp_min, clp_max,
                             reinterpret_cast<uint64_t*>(ctx->dDist));
  if (rc) return rc;
  CU(cudaMemcpyAsync(dist, ctx->dDist, sizeof(uint64_t) * n_cands, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

// ---- pipelined variant: the copies of batch i+1 overlap the kernels of batch i ----------------------------------
static int slotInit(mmgpu_ctx* ctx, MmSlot* sl)
{
  if (sl->init) return MMGPU_OK;
  CU(cudaEventCreateWithFlags(&sl->stage.freeEv, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&sl->copied, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&sl->prepared, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&sl->computed, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&sl->done, cudaEventDisableTiming));
  sl->init = true;
  return MMGPU_OK;
}

extern "C" int mmgpu_me_batch_submit(mmgpu_ctx* ctx, int slot, const int16_t* org, int org_stride, int org_w, int org_h,
                                     int n_blocks, const mmgpu_block* blocks, int n_cands, const mmgpu_mv* mvs,
                                     int mv_prec_shift, int dist_kind, int sub_shift, int clp_min, int clp_max, uint64_t* dist)
{
  if (!ctx || !org || !blocks || !mvs || !dist || n_blocks <= 0 || n_cands <= 0 || org_w <= 0 || org_h <= 0 || org_stride < org_w)
    return MMGPU_ERR_ARG;
  if (slot < 0 || slot >= MM_MAX_INFLIGHT) return fail(ctx, MMGPU_ERR_ARG, "slot out of range");
  if (dist_kind != MMGPU_DIST_SAD && dist_kind != MMGPU_DIST_HAD) return fail(ctx, MMGPU_ERR_ARG, "bad dist_kind");
  if (sub_shift < 0 || sub_shift > 2 || (dist_kind == MMGPU_DIST_HAD && sub_shift != 0)) return fail(ctx, MMGPU_ERR_ARG, "bad sub_shift");
  cudaSetDevice(ctx->cfg.device);
  MmSlot* sl = &ctx->slots[slot];
  int rc;
  if ((rc = slotInit(ctx, sl))) return rc;
  if (sl->pending) return fail(ctx, MMGPU_ERR_STATE, "slot still in flight: call mmgpu_me_batch_wait first");
  const int pitch = (org_w + 63) / 64 * 64;
  if ((rc = ensureCap(ctx, &sl->dOrg, &sl->capOrg, (size_t)pitch * org_h))) return rc;
  if ((rc = ensureCap(ctx, &sl->dMvs, &sl->capMvs, (size_t)n_cands))) return rc;
  if ((rc = ensureCap(ctx, &sl->dDist, &sl->capDist, (size_t)n_cands))) return rc;
  CU(cudaMemcpy2DAsync(sl->dOrg, 2 * (size_t)pitch, org, 2 * (size_t)org_stride, 2 * (size_t)org_w, org_h, cudaMemcpyHostToDevice,
                       ctx->copyStream));
  CU(cudaMemcpyAsync(sl->dMvs, mvs, sizeof(mmgpu_mv) * n_cands, cudaMemcpyHostToDevice, ctx->copyStream));
  rc = buildPlan(ctx, &sl->plan, n_blocks, blocks, n_cands, org_w, org_h, &sl->stage, ctx->copyStream, ctx->prepStream, sl->copied);
  if (!rc)
    rc = launchMe(ctx, &sl->plan, sl->dMvs, mv_prec_shift, MM_MODE_DIST, dist_kind, sub_shift, clp_min, clp_max, sl->dDist, nullptr, 0,
                  sl->dOrg, pitch, ctx->prepStream, sl->prepared);
  if (rc)
  {
    // copies of this batch may already be queued: drain them so that the caller's buffers are free again
    const std::string msg = ctx->err;
    cudaStreamSynchronize(ctx->copyStream);
    cudaStreamSynchronize(ctx->prepStream);
    cudaStreamSynchronize(ctx->stream);
    cudaGetLastError();
    ctx->err = msg;
    return rc;
  }
  CU(cudaEventRecord(sl->computed, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->readStream, sl->computed, 0));
  CU(cudaMemcpyAsync(dist, sl->dDist, sizeof(uint64_t) * n_cands, cudaMemcpyDeviceToHost, ctx->readStream));
  CU(cudaEventRecord(sl->done, ctx->readStream));
  sl->pending = true;
  return MMGPU_OK;
}

extern "C" int mmgpu_me_batch_wait(mmgpu_ctx* ctx, int slot)
{
  if (!ctx) return MMGPU_ERR_ARG;
  if (slot < 0 || slot >= MM_MAX_INFLIGHT) return fail(ctx, MMGPU_ERR_ARG, "slot out of range");
  MmSlot* sl = &ctx->slots[slot];
  if (!sl->init || !sl->pending) return fail(ctx, MMGPU_ERR_STATE, "nothing submitted on this slot");
  cudaSetDevice(ctx->cfg.device);
  CU(cudaEventSynchronize(sl->done));
  sl->pending = false;
  return MMGPU_OK;
}

// ---- a12: DMVR (projected) ------------------------------------------------------------------------------
extern "C" int mmgpu_dmvr_costs(mmgpu_ctx* ctx, int n_subpus, const mmgpu_dmvr_subpu* subpus, int clp_min, int clp_max,
                                uint64_t* costs)
{
  if (!ctx || !subpus || !costs || n_subpus <= 0) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  std::vector<MmDmvrItem> items;
  for (int i = 0; i < n_subpus; ++i)
  {
    const mmgpu_dmvr_subpu& s = subpus[i];
    int rc = checkBlockShape(ctx, 0, 0, s.x, s.y, s.w, s.h);
    if (rc) return rc;
    if (s.w > 16 || s.h > 16 || s.w < 8 || s.h < 8) return fail(ctx, MMGPU_ERR_ARG, "DMVR sub-PU must be 8..16 wide and high");
    if (s.model <= MMGPU_CLASSIC || s.model > MMGPU_GEODESIC_CAMPOSE) return fail(ctx, MMGPU_ERR_ARG, "DMVR projected needs a non-classic model");
    if (s.ref_slot0 < 0 || s.ref_slot0 >= ctx->cfg.max_refs || !ctx->hRefs[s.ref_slot0].plane00[0] || s.ref_slot1 < 0 ||
        s.ref_slot1 >= ctx->cfg.max_refs || !ctx->hRefs[s.ref_slot1].plane00[0])
      return fail(ctx, MMGPU_ERR_STATE, "DMVR sub-PU references an unset reference slot");
    if (s.rot_slot0 < 0 || s.rot_slot0 >= MM_MAX_ROT || s.rot_slot1 < 0 || s.rot_slot1 >= MM_MAX_ROT)
      return fail(ctx, MMGPU_ERR_ARG, "bad rot_slot");
    const int n = (s.w >> 2) * (s.h >> 2);
    const int per = 32 / (2 * n);
    for (int o = 0; o < 25; o += per) items.push_back(MmDmvrItem{i, o, 0, 0});
  }
  int rc;
  if ((rc = ensureCap(ctx, &ctx->dSubpus, &ctx->capSubpus, (size_t)n_subpus))) return rc;
  if ((rc = ensureCap(ctx, &ctx->dDmvrItems, &ctx->capDmvrItems, items.size()))) return rc;
  if ((rc = ensureCap(ctx, &ctx->dDist, &ctx->capDist, (size_t)n_subpus * 25))) return rc;
  CU(cudaMemcpyAsync(ctx->dSubpus, subpus, sizeof(mmgpu_dmvr_subpu) * n_subpus, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->dDmvrItems, items.data(), sizeof(MmDmvrItem) * items.size(), cudaMemcpyHostToDevice, ctx->stream));
  MmParams P = baseParams(ctx);
  P.clpMin = clp_min;
  P.clpMax = clp_max;
  P.oobMode = MMGPU_OOB_PRED;
  const int nItems = (int)items.size();
  mmk_dmvr(ctx->stream, P, ctx->dSubpus, ctx->dDmvrItems, nItems, ctx->cfg.had_mode == MMGPU_HAD_SIMD, ctx->dDist);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(costs, ctx->dDist, sizeof(uint64_t) * (size_t)n_subpus * 25, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

static int32_t divForMaxq7(int64_t N, int64_t D)   // InterPrediction.cpp:1958-1994
{
  int32_t sign = 0, q = 0;
  if (N < 0) { sign = 1; N = -N; }
  D = (D << 3);
  if (N >= D) { N -= D; q++; }
  q = (q << 1);
  D = (D >> 1);
  if (N >= D) { N -= D; q++; }
  q = (q << 1);
  if (N >= (D >> 1)) q++;
  return sign ? -q : q;
}

// ---- f4: classic DMVR (xProcessDMVR) ----------------------------------------------------------------------
// window [x, x + w) x [y, y + h) of component comp of a slot inside its padded plane?
static bool windowInside(mmgpu_ctx* ctx, int comp, int x, int y, int w, int h)
{
  const int sx = scaleX(comp, ctx->cfg.chroma_fmt), sy = scaleY(comp, ctx->cfg.chroma_fmt);
  const int mx = ctx->cfg.margin >> sx, my = ctx->cfg.margin >> sy, pw = ctx->cfg.pic_w >> sx, ph = ctx->cfg.pic_h >> sy;
  return x >= -mx && y >= -my && x + w <= pw + mx && y + h <= ph + my;
}

extern "C" int mmgpu_dmvr_classic_costs(mmgpu_ctx* ctx, int n_subpus, const mmgpu_dmvr_classic_subpu* subpus, uint64_t* costs)
{
  if (!ctx || !subpus || !costs || n_subpus <= 0) return MMGPU_ERR_ARG;
  for (int i = 0; i < n_subpus; ++i)
  {
    const mmgpu_dmvr_classic_subpu& s = subpus[i];
    if ((s.w != 8 && s.w != 16) || (s.h != 8 && s.h != 16)) return fail(ctx, MMGPU_ERR_ARG, "classic DMVR sub-PU must be 8 or 16 wide and high");
    if (s.ref_slot0 < 0 || s.ref_slot0 >= ctx->cfg.max_refs || !ctx->hRefs[s.ref_slot0].plane00[0] || s.ref_slot1 < 0 ||
        s.ref_slot1 >= ctx->cfg.max_refs || !ctx->hRefs[s.ref_slot1].plane00[0])
      return fail(ctx, MMGPU_ERR_STATE, "classic DMVR sub-PU references an unset reference slot");
    if ((unsigned)s.frac_x0 > 15u || (unsigned)s.frac_y0 > 15u || (unsigned)s.frac_x1 > 15u || (unsigned)s.frac_y1 > 15u)
      return fail(ctx, MMGPU_ERR_ARG, "classic DMVR: fraction out of range");
    if (!windowInside(ctx, 0, s.org_x0, s.org_y0, s.w + 7, s.h + 7) || !windowInside(ctx, 0, s.org_x1, s.org_y1, s.w + 7, s.h + 7))
      return fail(ctx, MMGPU_ERR_ARG, "classic DMVR: prefetch window outside the padded picture (clipMv)");
  }
  cudaSetDevice(ctx->cfg.device);
  int rc;
  if ((rc = ensureCap(ctx, &ctx->dSubpusC, &ctx->capSubpusC, (size_t)n_subpus))) return rc;
  if ((rc = ensureCap(ctx, &ctx->dDist, &ctx->capDist, (size_t)n_subpus * 25))) return rc;
  CU(cudaMemcpyAsync(ctx->dSubpusC, subpus, sizeof(mmgpu_dmvr_classic_subpu) * n_subpus, cudaMemcpyHostToDevice, ctx->stream));
  mmk_dmvr_classic(ctx->stream, ctx->dRefs, ctx->dSubpusC, n_subpus, ctx->cfg.bit_depth, ctx->dDist);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(costs, ctx->dDist, sizeof(uint64_t) * 25 * n_subpus, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

static int checkDmvrJob(mmgpu_ctx* ctx, const mmgpu_dmvr_mc_job& j, int dst_samples)
{
  const int ncomp = numComps(ctx);
  if (j.comp < 0 || j.comp >= ncomp) return fail(ctx, MMGPU_ERR_ARG, "dmvr_final_mc: bad component");
  if (j.ref_slot < 0 || j.ref_slot >= ctx->cfg.max_refs || !ctx->hRefs[j.ref_slot].plane00[j.comp])
    return fail(ctx, MMGPU_ERR_STATE, "dmvr_final_mc: unset reference slot");
  const int taps = j.comp == 0 ? 8 : 4;
  if (j.w < 1 || j.h < 1 || j.w > 16 || j.h > 16) return fail(ctx, MMGPU_ERR_ARG, "dmvr_final_mc: block must be 1..16 wide and high");
  if (j.pad < 0 || j.pad > 2 || abs(j.dint_x) > j.pad || abs(j.dint_y) > j.pad) return fail(ctx, MMGPU_ERR_ARG, "dmvr_final_mc: bad pad / deltaIntMv");
  const int nph = j.comp == 0 ? 16 : 32;
  if (j.frac_x < 0 || j.frac_x >= nph || j.frac_y < 0 || j.frac_y >= nph) return fail(ctx, MMGPU_ERR_ARG, "dmvr_final_mc: fraction out of range");
  if (j.bio && j.comp != 0) return fail(ctx, MMGPU_ERR_ARG, "dmvr_final_mc: bio is a luma flag");
  if (!windowInside(ctx, j.comp, j.org_x, j.org_y, j.w + taps - 1, j.h + taps - 1))
    return fail(ctx, MMGPU_ERR_ARG, "dmvr_final_mc: prefetch window outside the padded picture (clipMv)");
  const long last = j.bio ? (long)j.dst_offset + (long)(j.w + 4) * (j.h + 4) : (long)j.dst_offset + (long)(j.h - 1) * j.dst_stride + j.w;
  if (j.dst_offset < 0 || (!j.bio && j.dst_stride < j.w) || last > dst_samples) return fail(ctx, MMGPU_ERR_ARG, "dmvr_final_mc: output outside dst");
  return MMGPU_OK;
}

extern "C" int mmgpu_dmvr_final_mc(mmgpu_ctx* ctx, int n_jobs, const mmgpu_dmvr_mc_job* jobs, int16_t* dst, int dst_samples)
{
  if (!ctx || !jobs || !dst || n_jobs <= 0 || dst_samples <= 0) return MMGPU_ERR_ARG;
  int rc;
  for (int i = 0; i < n_jobs; ++i)
    if ((rc = checkDmvrJob(ctx, jobs[i], dst_samples))) return rc;
  cudaSetDevice(ctx->cfg.device);
  if ((rc = ensureCap(ctx, &ctx->dDmvrJobs, &ctx->capDmvrJobs, (size_t)n_jobs))) return rc;
  if ((rc = ensureCap(ctx, &ctx->dTmp, &ctx->capTmp, (size_t)dst_samples + 64))) return rc;
  CU(cudaMemcpyAsync(ctx->dDmvrJobs, jobs, sizeof(mmgpu_dmvr_mc_job) * n_jobs, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->dTmp, 0, sizeof(int16_t) * dst_samples, ctx->stream));
  mmk_dmvr_final_mc(ctx->stream, ctx->dRefs, ctx->fir, ctx->dDmvrJobs, n_jobs, 0, 0, 0, ctx->dTmp);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(dst, ctx->dTmp, sizeof(int16_t) * dst_samples, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}

extern "C" int mmgpu_mc_alt_hpel(mmgpu_ctx* ctx, int ref_slot, int x, int y, int w, int h, mmgpu_mv mv, int bi, int clp_min, int clp_max,
                                 int16_t* dst, int dst_stride)
{
  if (!ctx || !dst || dst_stride < w) return MMGPU_ERR_ARG;
  if (ref_slot < 0 || ref_slot >= ctx->cfg.max_refs || !ctx->hRefs[ref_slot].plane00[0]) return fail(ctx, MMGPU_ERR_STATE, "reference slot not set");
  int rc = checkBlockShape(ctx, 0, 0, x, y, w, h, true);
  if (rc) return rc;
  // 16x16 tiles, each its own un-padded window: identical to reading the picture
  std::vector<mmgpu_dmvr_mc_job> jobs;
  for (int ty = 0; ty < h; ty += 16)
    for (int tx = 0; tx < w; tx += 16)
    {
      mmgpu_dmvr_mc_job j = {};
      j.comp = 0; j.ref_slot = ref_slot;
      j.org_x = x + tx + (mv.hor >> 4) - 3; j.org_y = y + ty + (mv.ver >> 4) - 3;
      j.w = std::min(16, w - tx); j.h = std::min(16, h - ty);
      j.frac_x = mv.hor & 15; j.frac_y = mv.ver & 15;
      j.alt_hpel = 1;
      j.dst_offset = ty * w + tx; j.dst_stride = w;
      if ((rc = checkDmvrJob(ctx, j, w * h))) return rc;
      jobs.push_back(j);
    }
  cudaSetDevice(ctx->cfg.device);
  if ((rc = ensureCap(ctx, &ctx->dDmvrJobs, &ctx->capDmvrJobs, jobs.size()))) return rc;
  if ((rc = ensureCap(ctx, &ctx->dTmp, &ctx->capTmp, (size_t)w * h + 64))) return rc;
  CU(cudaMemcpyAsync(ctx->dDmvrJobs, jobs.data(), sizeof(mmgpu_dmvr_mc_job) * jobs.size(), cudaMemcpyHostToDevice, ctx->stream));
  mmk_dmvr_final_mc(ctx->stream, ctx->dRefs, ctx->fir, ctx->dDmvrJobs, (int)jobs.size(), bi ? 0 : 1, clp_min, clp_max, ctx->dTmp);
  ctx->launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpy2DAsync(dst, sizeof(int16_t) * dst_stride, ctx->dTmp, sizeof(int16_t) * w, sizeof(int16_t) * w, h, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));   // jobs (host vector) must outlive the upload
  return MMGPU_OK;
}

extern "C" int mmgpu_dmvr_decide(const uint64_t costs25[25], int w, int h, int32_t delta_mv[2], uint64_t* min_cost)
{
  // xProcessDMVRProjected's integer search bookkeeping (iterationCount == 1) + xDMVRSubPixelErrorSurface
  static const int offs[25][2] = {{-2, -2}, {-1, -2}, {0, -2}, {1, -2}, {2, -2}, {-2, -1}, {-1, -1}, {0, -1}, {1, -1}, {2, -1},
                                  {-2, 0},  {-1, 0},  {0, 0},  {1, 0},  {2, 0},  {-2, 1},  {-1, 1},  {0, 1},  {1, 1},  {2, 1},
                                  {-2, 2},  {-1, 2},  {0, 2},  {1, 2},  {2, 2}};
  uint64_t sads[25];
  for (int i = 0; i < 25; ++i) sads[i] = UINT64_MAX;
  uint64_t* p = &sads[12];
  int notZero = 1;
  int16_t total[2] = {0, 0}, delta[2] = {0, 0};
  uint64_t minCost = costs25[12];
  minCost -= (minCost >> 2);
  if (minCost < (uint64_t)(w * h)) notZero = 0;
  else
  {
    p[0] = minCost;
    for (int n = 0; n < 25; ++n)
    {
      const int o = offs[n][1] * 5 + offs[n][0];
      if (p[o] == UINT64_MAX) p[o] = costs25[n];
      if (p[o] < minCost) { minCost = p[o]; delta[0] = (int16_t)offs[n][0]; delta[1] = (int16_t)offs[n][1]; }
    }
    if (!(delta[0] == 0 && delta[1] == 0))
    {
      total[0] = (int16_t)(total[0] + delta[0]);
      total[1] = (int16_t)(total[1] + delta[1]);
      p += delta[1] * 5 + delta[0];
    }
  }
  total[0] = (int16_t)(total[0] << 4);
  total[1] = (int16_t)(total[1] << 4);
  if (notZero && (abs(total[0]) != (2 << 4)) && (abs(total[1]) != (2 << 4)))
  {
    const uint64_t sb[5] = {p[0], p[-1], p[-5], p[1], p[5]};
    int32_t t[2] = {0, 0};
    int64_t num = (int64_t)((sb[1] - sb[3]) << 4), den = (int64_t)((sb[1] + sb[3] - (sb[0] << 1)));
    if (den != 0) t[0] = ((sb[1] != sb[0]) && (sb[3] != sb[0])) ? divForMaxq7(num, den) : ((sb[1] == sb[0]) ? -8 : 8);
    num = (int64_t)((sb[2] - sb[4]) << 4);
    den = (int64_t)((sb[2] + sb[4] - (sb[0] << 1)));
    if (den != 0) t[1] = ((sb[2] != sb[0]) && (sb[4] != sb[0])) ? divForMaxq7(num, den) : ((sb[2] == sb[0]) ? -8 : 8);
    total[0] = (int16_t)(total[0] + t[0]);
    total[1] = (int16_t)(total[1] + t[1]);
  }
  delta_mv[0] = total[0];
  delta_mv[1] = total[1];
  if (min_cost) *min_cost = minCost;
  return notZero;
}

// ---- whole-picture reconstruction shape -------------------------------------------------------------------
static int checkJobs(mmgpu_ctx* ctx, int n_jobs, const mmgpu_mc_job* jobs, bool used[3], long area[3])
{
  const int fmt = ctx->cfg.chroma_fmt;
  for (int i = 0; i < n_jobs; ++i)
  {
    const mmgpu_mc_job& j = jobs[i];
    if (j.comp < 0 || j.comp >= numComps(ctx)) return fail(ctx, MMGPU_ERR_ARG, "bad component in job");
    int rc = checkBlockShape(ctx, j.comp, fmt, j.x, j.y, j.w, j.h, true);
    if (rc) return rc;
    if (j.model < MMGPU_CLASSIC || j.model > MMGPU_GEODESIC_CAMPOSE) return fail(ctx, MMGPU_ERR_ARG, "bad model in job");
    const int nl = j.bi ? 2 : 1;
    for (int l = 0; l < nl; ++l)
    {
      const int rs = l ? j.ref_slot1 : j.ref_slot0, ro = l ? j.rot_slot1 : j.rot_slot0;
      if (rs < 0 || rs >= ctx->cfg.max_refs || !ctx->hRefs[rs].plane00[j.comp]) return fail(ctx, MMGPU_ERR_STATE, "job references an unset reference slot");
      if (ro < 0 || ro >= MM_MAX_ROT) return fail(ctx, MMGPU_ERR_ARG, "bad rot_slot in job");
    }
    used[j.comp] = true;
    area[j.comp] += (long)j.w * j.h;
  }
  return MMGPU_OK;
}

static int launchJobs(mmgpu_ctx* ctx, int n_jobs, const mmgpu_mc_job* jobs, int clp_min, int clp_max, int16_t* const d[3], const int pitch[3])
{
  int rc;
  // work list: chunks of MM_JOB_CHUNK sub-blocks
  const int fmtc = ctx->cfg.chroma_fmt;
  std::vector<int2> items;
  items.reserve((size_t)n_jobs * 2);
  for (int i = 0; i < n_jobs; ++i)
  {
    const int sbw = 4 >> scaleX(jobs[i].comp, fmtc), sbh = 4 >> scaleY(jobs[i].comp, fmtc);
    const int n = (jobs[i].w / sbw) * (jobs[i].h / sbh);
    for (int f = 0; f < n; f += MM_JOB_CHUNK) items.push_back(make_int2(i, f));
  }
  if ((rc = ensureCap(ctx, &ctx->dJobs, &ctx->capJobs, (size_t)n_jobs))) return rc;
  if ((rc = ensureCap(ctx, &ctx->dJobItems, &ctx->capJobItems, items.size()))) return rc;
  CU(cudaMemcpyAsync(ctx->dJobs, jobs, sizeof(mmgpu_mc_job) * n_jobs, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->dJobItems, items.data(), sizeof(int2) * items.size(), cudaMemcpyHostToDevice, ctx->stream));
  MmParams P = baseParams(ctx);
  P.clpMin = clp_min;
  P.clpMax = clp_max;
  P.oobMode = MMGPU_OOB_PRED;
  const int fmt = ctx->cfg.chroma_fmt;
  const int shift = ctx->fir.cpShift + 1, offset = (1 << (shift - 1)) + 2 * 8192;
  mmk_mc_jobs(ctx->stream, P, (int)items.size(), ctx->dJobs, ctx->dJobItems, d[0], d[1], d[2], pitch[0], pitch[1], scaleX(1, fmt), scaleY(1, fmt),
              shift, offset);
  ctx->launches++;
  CU(cudaGetLastError());
  return MMGPU_OK;
}

extern "C" int mmgpu_mc_jobs_device(mmgpu_ctx* ctx, int n_jobs, const mmgpu_mc_job* jobs, int clp_min, int clp_max,
                                    int16_t* const d_dst[3], const int32_t dst_stride[3])
{
  if (!ctx || !jobs || !d_dst || !dst_stride || n_jobs <= 0) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  bool used[3] = {false, false, false};
  long area[3] = {0, 0, 0};
  int rc = checkJobs(ctx, n_jobs, jobs, used, area);
  if (rc) return rc;
  if (numComps(ctx) == 3 && used[1] && used[2] && dst_stride[1] != dst_stride[2]) return fail(ctx, MMGPU_ERR_ARG, "Cb and Cr planes must share one stride");
  for (int c = 0; c < numComps(ctx); ++c)
    if (used[c] && !d_dst[c]) return fail(ctx, MMGPU_ERR_ARG, "null destination plane");
  const int pitch[3] = {dst_stride[0], dst_stride[1], dst_stride[2]};
  return launchJobs(ctx, n_jobs, jobs, clp_min, clp_max, d_dst, pitch);
}

extern "C" int mmgpu_mc_jobs(mmgpu_ctx* ctx, int n_jobs, const mmgpu_mc_job* jobs, int clp_min, int clp_max,
                             int16_t* const dst[3], const int32_t dst_stride[3])
{
  if (!ctx || !jobs || !dst || !dst_stride || n_jobs <= 0) return MMGPU_ERR_ARG;
  cudaSetDevice(ctx->cfg.device);
  bool used[3] = {false, false, false};
  long area[3] = {0, 0, 0};
  int rc = checkJobs(ctx, n_jobs, jobs, used, area);
  if (rc) return rc;
  int pitch[3] = {0, 0, 0}, cw[3], ch[3];
  for (int c = 0; c < numComps(ctx); ++c)
  {
    int mx, my;
    compDims(ctx, c, &cw[c], &ch[c], &mx, &my);
    pitch[c] = (cw[c] + 63) / 64 * 64;
    if (!used[c]) continue;
    if (!dst[c]) return fail(ctx, MMGPU_ERR_ARG, "null destination plane");
    if ((rc = ensureCap(ctx, &ctx->dPic[c], &ctx->capPic[c], (size_t)pitch[c] * ch[c]))) return rc;
    // the caller's picture already holds the samples of blocks this call does not cover; when the jobs tile the whole
    // plane (blocks of one reconstruction never overlap) there is nothing to preserve and the upload is skipped
    if (area[c] < (long)cw[c] * ch[c])
      CU(cudaMemcpy2DAsync(ctx->dPic[c], 2 * (size_t)pitch[c], dst[c], 2 * (size_t)dst_stride[c], 2 * (size_t)cw[c], ch[c],
                           cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = launchJobs(ctx, n_jobs, jobs, clp_min, clp_max, ctx->dPic, pitch))) return rc;
  for (int c = 0; c < numComps(ctx); ++c)
    if (used[c])
      CU(cudaMemcpy2DAsync(dst[c], 2 * (size_t)dst_stride[c], ctx->dPic[c], 2 * (size_t)pitch[c], 2 * (size_t)cw[c], ch[c],
                           cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MMGPU_OK;
}
