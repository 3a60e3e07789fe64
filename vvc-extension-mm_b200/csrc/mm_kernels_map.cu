// mm_kernels_map.cu -- float side of the path: per-sequence caches, per-candidate scalars and the positions of every
// sub-block of every candidate.  See mm_common.cuh for the kernel list.
#include "mm_sbpos.cuh"

// -------------------------------------------------------------------------------------------------
// k_mpa_cache: picture-wide motion-plane cache (always the vectorised Eigen path: n = W/4*H/4)
// -------------------------------------------------------------------------------------------------
__global__ void k_mpa_cache(MmGeom g, int model, float2* out, uint8_t* vip)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.rows4 * g.cols4) return;
  const int row = idx / g.cols4, col = idx - row * g.cols4;
  const float X = MM_FADD((float)(4 * col), g.off), Y = MM_FADD((float)(4 * row), g.off);
  const long n = (long)g.rows4 * g.cols4;
  const bool pk = ((long)col * g.rows4 + row) < (n & ~3L);
  float PX, PY;
  bool v;
  MmTrig none; none.valid = false; none.st = none.ct = none.sp = none.cp = 0.f;
  mm_mpa_to_perspective(g, model, X, Y, pk, none, &PX, &PY, &v);
  out[idx] = make_float2(PX, PY);
  vip[idx] = v ? 1 : 0;
}

// ERP toSphere angles of the picture grid, scalar-libm and Eigen-packet variants (Projection.cpp:213-221,
// Coordinate.cpp:53-60): tab[0] = scalar, tab[1] = packet; th over rows, ph over columns.
__global__ void k_trig_tables(MmGeom g, float2* th0, float2* th1, float2* ph0, float2* ph1)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool fma = g.fmaBuild != 0;
  if (i < g.rows4)
  {
    float th, ph, s, c;
    mm_erp_angles(g, 0.0f, MM_FADD((float)(4 * i), g.off), &th, &ph);
    mm_sincos_pk(th, false, fma, &s, &c); th0[i] = make_float2(s, c);
    mm_sincos_pk(th, true, fma, &s, &c);  th1[i] = make_float2(s, c);
  }
  if (i < g.cols4)
  {
    float th, ph, s, c;
    mm_erp_angles(g, MM_FADD((float)(4 * i), g.off), 0.0f, &th, &ph);
    mm_sincos_pk(ph, false, fma, &s, &c); ph0[i] = make_float2(s, c);
    mm_sincos_pk(ph, true, fma, &s, &c);  ph1[i] = make_float2(s, c);
  }
}

// -------------------------------------------------------------------------------------------------
// k_build_items: the work list of a batch, one warp per block.  itemBase[b] = first work item of block b
// (host-computed, grouped by motion model); also fills the candidate -> block map used by k_cand_setup.
// -------------------------------------------------------------------------------------------------
__global__ void k_build_items(const mmgpu_block* __restrict__ blocks, const int32_t* __restrict__ itemBase, int nBlocks,
                              MmItem* __restrict__ items, int32_t* __restrict__ candBlock)
{
  const int bi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bi >= nBlocks) return;
  const mmgpu_block b = blocks[bi];
  const int n = (b.w >> 2) * (b.h >> 2);
  const int base = itemBase[bi];
  for (int j = lane; j < b.cand_count; j += 32) candBlock[b.cand_first + j] = bi;
  // a work item = 32 consecutive (candidate, sub-block) entries of the block: one candidate's 32-sub-block chunk
  // (n >= 32) or 32/n candidates (n < 32)
  MmItem it;
  it.x = b.x; it.y = b.y;
  it.wh = (uint32_t)b.w | ((uint32_t)b.h << 16);
  it.rmr = (uint32_t)b.ref_slot | ((uint32_t)b.model << 16) | ((uint32_t)b.rot_slot << 24);
  it.orgxy = (uint32_t)b.org_x | ((uint32_t)b.org_y << 16);
  it.candEnd = b.cand_first + b.cand_count;
  if (n >= 32)
  {
    // chunk-major: the candidates of one 32-sub-block chunk are consecutive items (k_map prepares the chunk once)
    const int chunks = n >> 5, total = b.cand_count * chunks;
    for (int t = lane; t < total; t += 32)
    {
      const int ch = t / b.cand_count, c0 = t - ch * b.cand_count;
      it.cand = b.cand_first + c0;
      it.sb = ch << 5;
      items[base + t] = it;
    }
  }
  else
  {
    const int per = 32 / n, total = (b.cand_count + per - 1) / per;
    for (int t = lane; t < total; t += 32)
    {
      it.cand = b.cand_first + t * per;
      it.sb = 0;
      items[base + t] = it;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// k_cand_setup: one thread per candidate MV
// -------------------------------------------------------------------------------------------------
__global__ void k_cand_setup(MmGeom g, const mmgpu_block* __restrict__ blocks, const int32_t* __restrict__ candBlock,
                             const mmgpu_mv* __restrict__ mvs, int mvShift, const float* __restrict__ rot, int rotOverride,
                             int nCands, MmCand* __restrict__ out, unsigned long long* __restrict__ dist)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nCands) return;
  if (dist) dist[i] = 0ull;   // k_fir accumulates into it (saves a separate memset launch)
  const mmgpu_block b = blocks[candBlock[i]];
  const mmgpu_mv mv = mvs[i];
  MmCand c;
  // Mv::changePrecision(mvPrec, MV_PRECISION_INTERNAL) is a left shift (Mv.h)
  const int mh = mv.hor * (1 << mvShift), mvv = mv.ver * (1 << mvShift);
  if (b.model == MM_CLASSIC)
  {
    c.mvX = 0.f; c.mvY = 0.f; c.mvZero = 0;
    for (int k = 0; k < 9; ++k) c.p[k] = 0.f;
    c.p[0] = __int_as_float(mh);
    c.p[1] = __int_as_float(mvv);
  }
  else
  {
    mm_cand_setup(g, b.model, true, b.x, b.y, b.w, b.h, mh, mvv, mm_rot_ptr(rot, b.model, rotOverride >= 0 ? rotOverride : b.rot_slot), &c);
  }
  out[i] = c;
}

// -------------------------------------------------------------------------------------------------
// k_reproject: positions of one block (parity entry point a1)
// -------------------------------------------------------------------------------------------------
__global__ void k_reproject(MmParams P, int comp, int csx, int csy, int x, int y, int w, int h, mmgpu_mv mv, int model,
                            int rotSlot, int32_t* outX, int32_t* outY)
{
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = h / sbh, cols = w / sbw;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int col = i / rows, row = i - col * rows;
  const float* R = mm_rot_ptr(P.rot, model, rotSlot);
  MmCand c;
  mm_cand_setup(P.g, model, comp == 0, x, y, w, h, mv.hor, mv.ver, R, &c);
  int fx, fy;
  mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, comp, csx, csy, x, y, w, h, c, R, row, col, &fx, &fy);
  outX[i] = fx;
  outY[i] = fy;
}

// -------------------------------------------------------------------------------------------------
// k_map: one lane per (candidate, sub-block) entry of the batch; writes its 1/16-pel position, or MM_POS_OOB in .x
// when the sub-block falls outside the picture by the caller's OOB rule (InterPrediction.cpp:776-780 /
// InterSearch.cpp:6329-6332: the prediction of such a sub-block is 0).
// A warp walks MM_MAP_K consecutive work items.  Items of one 32-sub-block chunk (or of one small block) are consecutive
// and differ only in the candidate, so the candidate-independent part of the mapping (grid look-ups, toSphere, the
// model's prepare step: mm_subblock_prepare) is computed once and only mm_subblock_apply runs per candidate; it is
// recomputed whenever the next item belongs to another chunk / block / model.
// MODEL >= 0: every item of the launch uses that motion model (items are grouped by model); MODEL < 0: generic.
// -------------------------------------------------------------------------------------------------
#ifndef MM_MAP_WARPS
#define MM_MAP_WARPS 4       // warps per CTA
#endif
#ifndef MM_MAP_MINBLOCKS
#define MM_MAP_MINBLOCKS (32 / MM_MAP_WARPS)
#endif
#ifndef MM_MAP_K
#define MM_MAP_K 4
#endif
template <int MODEL>
__global__ void __launch_bounds__(32 * MM_MAP_WARPS, MM_MAP_MINBLOCKS) k_map(const MmParams P, int first, int count, int2* __restrict__ pos)
{
  const int w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * MM_MAP_K;
  const int lane = threadIdx.x & 31;
  if (w0 >= count) return;
  const int nIt = min(MM_MAP_K, count - w0);
  const int mc = (P.oobMode == MMGPU_OOB_ME) ? 0 : P.maxCU;
  MmSbState st;
  MmSub s;
  int4 key = make_int4(-1, -1, -1, -1);   // (x, y, wh, rmr) of the prepared chunk; sb is checked separately
  int keySb = -1, n = 0, bw = 0, bh = 0, model = 0;
  const float* R = P.rot;
#pragma unroll 1
  for (int j = 0; j < nIt; ++j)
  {
    const int warp = first + w0 + j;
    const MmItem it = P.items[warp];
    const bool fresh = it.x != key.x || it.y != key.y || (int)it.wh != key.z || (int)it.rmr != key.w || it.sb != keySb;
    if (fresh)
    {
      bw = it.wh & 0xffff; bh = it.wh >> 16;
      n = (bw >> 2) * (bh >> 2);
      s = mm_sub_from_index(bw, bh, (n >= 32) ? it.sb + lane : lane % n);
      model = (MODEL >= 0) ? MODEL : (int)((it.rmr >> 16) & 0xff);
      R = mm_rot_ptr(P.rot, model, (P.rotOverride >= 0) ? P.rotOverride : (int)(it.rmr >> 24));
    }
    const int cand = (n >= 32) ? it.cand : it.cand + lane / n;
    const int candc = (cand < it.candEnd) ? cand : it.cand;  // lanes past the last candidate redo a valid one, result unused
    const MmCand c = P.cands[candc];
    int fx, fy;
    if (model == MM_CLASSIC)
    {
      fx = ((it.x + s.col * 4) << 4) + __float_as_int(c.p[0]);
      fy = ((it.y + s.row * 4) << 4) + __float_as_int(c.p[1]);
      // xPredInterBlk has no zero-fill rule: the reference clips the MV beforehand (clipMv) so that the block stays inside
      // the padded picture.  Positions of unclipped MVs are held inside the padded plane here, for memory safety only.
      const int mg = P.tmaOffX - 16;
      fx = min(max(fx, -mg << 4), (P.picW + mg) << 4);
      fy = min(max(fy, -mg << 4), (P.picH + mg) << 4);
    }
    else
    {
      if (fresh)
        mm_subblock_prepare(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, 0, 0, 0, it.x, it.y, bw, bh, c, R, s.row, s.col, &st);
      mm_subblock_apply(P.g, model, 0, 0, 0, c, R, st, &fx, &fy);
      const int xP = fx >> 4, yP = fy >> 4;
      if (mm_oob(xP, yP, P.picW, P.picH, mc, mc, 4, 4)) fx = MM_POS_OOB;
    }
    key = make_int4(it.x, it.y, (int)it.wh, (int)it.rmr);
    keySb = it.sb;
    pos[(long)warp * 32 + lane] = make_int2(fx, fy);
  }
}

// -------------------------------------------------------------------------------------------------
// k_tan_selftest: checksum of (x, float(sin((double)x)), float(cos((double)x))) over every `stride`-th float of
// |x| <= float(pi/2) -- the same sum oracle/tools/tan_exceptions.cpp prints for glibc (MM_TAN_SWEEP_CHECKSUM for stride 1).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mm_tan_mix(uint32_t x, uint32_t s, uint32_t c)
{
  unsigned long long h = 0x9e3779b97f4a7c15ull * (unsigned long long)(x + 1u);
  h ^= (unsigned long long)s * 0xc2b2ae3d27d4eb4full;
  h = (h << 31) | (h >> 33);
  h ^= (unsigned long long)c * 0x165667b19e3779f9ull;
  return h * 0xff51afd7ed558ccdull;
}
__global__ void k_tan_selftest(uint32_t stride, unsigned long long* out)
{
  const unsigned long long maxBits = 0x3fc90fdbull;
  const unsigned long long n = maxBits / stride + 1;
  unsigned long long sum = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += (unsigned long long)gridDim.x * blockDim.x)
  {
    const int sign = i >= n;
    const unsigned long long m = (sign ? i - n : i) * stride;
    if (sign && m == 0) continue;
    const uint32_t xb = (uint32_t)m | (sign ? 0x80000000u : 0u);
    const float x = __uint_as_float(xb);
    sum += mm_tan_mix(xb, __float_as_uint(mm_sin_d2f(x)), __float_as_uint(mm_cos_d2f(x)));
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, sum);
}
void mmk_tan_selftest(cudaStream_t s, uint32_t stride, unsigned long long* out)
{
  k_tan_selftest<<<148 * 8, 256, 0, s>>>(stride, out);
}

#ifndef MM_CS_THREADS
#define MM_CS_THREADS 128
#endif
// ---- launch wrappers -----------------------------------------------------------------------------
void mmk_mpa_cache(cudaStream_t s, const MmGeom& g, int model, float2* out, uint8_t* vip)
{
  const int n4 = g.rows4 * g.cols4;
  k_mpa_cache<<<(n4 + 127) / 128, 128, 0, s>>>(g, model, out, vip);
}
void mmk_trig_tables(cudaStream_t s, const MmGeom& g, float2* th0, float2* th1, float2* ph0, float2* ph1)
{
  const int nt = g.rows4 > g.cols4 ? g.rows4 : g.cols4;
  k_trig_tables<<<(nt + 127) / 128, 128, 0, s>>>(g, th0, th1, ph0, ph1);
}
void mmk_build_items(cudaStream_t s, const mmgpu_block* blocks, const int32_t* itemBase, int nBlocks, MmItem* items,
                     int32_t* candBlock)
{
  k_build_items<<<(nBlocks + 3) / 4, 128, 0, s>>>(blocks, itemBase, nBlocks, items, candBlock);
}
void mmk_cand_setup(cudaStream_t s, const MmGeom& g, const mmgpu_block* blocks, const int32_t* candBlock, const mmgpu_mv* mvs,
                    int mvShift, const float* rot, int rotOverride, int nCands, MmCand* out, unsigned long long* dist)
{
  k_cand_setup<<<(nCands + MM_CS_THREADS - 1) / MM_CS_THREADS, MM_CS_THREADS, 0, s>>>(g, blocks, candBlock, mvs, mvShift, rot, rotOverride, nCands,
                                                                                      out, dist);
}
void mmk_reproject(cudaStream_t s, const MmParams& P, int comp, int csx, int csy, int x, int y, int w, int h, mmgpu_mv mv,
                   int model, int rotSlot, int32_t* outX, int32_t* outY)
{
  const int n = (w / (4 >> csx)) * (h / (4 >> csy));
  k_reproject<<<(n + 63) / 64, 64, 0, s>>>(P, comp, csx, csy, x, y, w, h, mv, model, rotSlot, outX, outY);
}
void mmk_map(cudaStream_t s, const MmParams& P, int first, int count, int model, int2* pos)
{
  if (count <= 0) return;
  const int grid = ((count + MM_MAP_K - 1) / MM_MAP_K + MM_MAP_WARPS - 1) / MM_MAP_WARPS;
  switch (model)
  {
#ifdef MM_MAP_PER_MODEL
    case MM_CLASSIC: k_map<MM_CLASSIC><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_MPA_FRONT_BACK: k_map<MM_MPA_FRONT_BACK><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_MPA_LEFT_RIGHT: k_map<MM_MPA_LEFT_RIGHT><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_MPA_TOP_BOTTOM: k_map<MM_MPA_TOP_BOTTOM><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_TANGENTIAL: k_map<MM_TANGENTIAL><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_THREE_D_TRANSLATIONAL: k_map<MM_THREE_D_TRANSLATIONAL><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_ROTATIONAL: k_map<MM_ROTATIONAL><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_GEODESIC_X: k_map<MM_GEODESIC_X><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_GEODESIC_Y: k_map<MM_GEODESIC_Y><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_GEODESIC_Z: k_map<MM_GEODESIC_Z><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
    case MM_GEODESIC_CAMPOSE: k_map<MM_GEODESIC_CAMPOSE><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
#endif
    default: k_map<-1><<<grid, 32 * MM_MAP_WARPS, 0, s>>>(P, first, count, pos); break;
  }
}
