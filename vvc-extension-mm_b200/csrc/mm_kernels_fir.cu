// mm_kernels_fir.cu -- integer side of the path: 8-tap separable interpolation at the positions k_map wrote, then SAD /
// Hadamard SATD per candidate or prediction write-out (a5); border extension and addAvg.
#include "mm_fir.cuh"

// -------------------------------------------------------------------------------------------------
// k_fir: one warp per work item (32 sub-block-candidates of one block), one 4x4 sub-block per lane.
// -------------------------------------------------------------------------------------------------
#ifndef MM_FIR_WARPS
#define MM_FIR_WARPS 1       // warps per CTA: one, so a finished warp frees its slot at once (1/2/4/8 measured: 0.0967/0.0976/0.0982/0.1020 ms)
#endif
#ifndef MM_FIR_MINBLOCKS
#define MM_FIR_MINBLOCKS (32 / MM_FIR_WARPS)
#endif
template <int MODE>
__global__ void __launch_bounds__(32 * MM_FIR_WARPS, MM_FIR_MINBLOCKS) k_fir(const MmParams P, int first, int count, const int2* __restrict__ pos)
{
#ifndef MM_NO_TMA
  __shared__ __align__(128) unsigned char s_box[MM_FIR_WARPS][MM_TMA_BOX_BYTES];
  __shared__ __align__(8) unsigned long long s_bar[MM_FIR_WARPS];
#endif
  const uint2* btaps = reinterpret_cast<const uint2*>(P.hTaps + MM_HTAPS_WORDS);   // (C0, C1) of the 16 phases: one 128-byte line
  const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w0 >= count) return;
  const int warp = first + w0;
  const MmItem it = P.items[warp];
  const int2 fp = __ldg(&pos[(long)warp * 32 + lane]);
  const int bw = it.wh & 0xffff, bh = it.wh >> 16;
  const int n = (bw >> 2) * (bh >> 2);
  int cand, sb;
  if (n >= 32) { cand = it.cand; sb = it.sb + lane; }
  else { cand = it.cand + lane / n; sb = lane % n; }
  const bool valid = cand < it.candEnd;
  const MmSub s = mm_sub_from_index(bw, bh, sb);
  const int refSlot = (P.refOverride >= 0) ? P.refOverride : (int)(it.rmr & 0xffff);
  const MmRef ref = P.refs[refSlot];
  const int16_t* orgp = P.org + (long)((int)(it.orgxy >> 16) + s.row * 4) * P.orgStride + ((int)(it.orgxy & 0xffff) + s.col * 4);

  // Reference samples.  The 32 windows of a warp (11 rows x 14 samples each, neighbours overlapping by 2/3) lie inside a
  // small box of the plane: one lane asks the TMA unit to copy that box into shared memory and the lanes gather their
  // windows from there (per-lane 8-byte reads of shared memory instead of scattered global requests).  Warps whose
  // windows do not fit a box shape (strongly diverging motion, candidates of small blocks sharing a warp, planes
  // without descriptors) gather from global memory directly.
  const bool oob = fp.x == MM_POS_OOB;
  const int xP = fp.x >> 4, yP = fp.y >> 4;
  const int ax = (xP - 3) & ~3, ay = yP - 3;
  int boxType = -1, x0 = 0, y0 = 0;
  // Full-pel warp (classic candidates at integer MVs: the bulk of an integer TZ search): both passes are
  // filterCopy<true,true> (InterpolationFilter.cpp:392-470), the prediction is the reference block itself.
  const bool fullPel = __all_sync(0xffffffffu, oob || ((fp.x | fp.y) & 15) == 0);
#ifndef MM_NO_TMA
  if (!fullPel && n >= 32 && ref.tma)   // warp-uniform
  {
    const int BIG = 1 << 28;
    x0 = __reduce_min_sync(0xffffffffu, oob ? BIG : ax) & ~7;   // TMA: the box must start on a 16-byte boundary
    y0 = __reduce_min_sync(0xffffffffu, oob ? BIG : ay);
    const int wbx = __reduce_max_sync(0xffffffffu, oob ? -BIG : ax + 16) - x0;
    const int hbx = __reduce_max_sync(0xffffffffu, oob ? -BIG : ay + 11) - y0;
    const int dims[MM_TMA_TYPES][2] = MM_TMA_BOX_DIMS;
    if (wbx > 0)
    {
#pragma unroll
      for (int t = MM_TMA_TYPES - 1; t >= 0; --t)
        if (wbx <= dims[t][0] && hbx <= dims[t][1]) boxType = t;
    }
  }
  int pitchQ = 0;   // box row pitch in 8-byte words
  if (boxType >= 0)
  {
    const int dims[MM_TMA_TYPES][2] = MM_TMA_BOX_DIMS;
    int bxw = dims[0][0], bxh = dims[0][1];
#pragma unroll
    for (int t = 1; t < MM_TMA_TYPES; ++t)
      if (boxType == t) { bxw = dims[t][0]; bxh = dims[t][1]; }
    pitchQ = bxw >> 2;
    unsigned long long* bar = &s_bar[threadIdx.x >> 5];
    if (lane == 0)
    {
      mm_mbar_init(bar, 1);
      mm_mbar_expect_tx(bar, (uint32_t)(bxw * bxh * 2));
      const char* tmap = reinterpret_cast<const char*>(P.tmaps) + ((size_t)refSlot * MM_TMA_TYPES + boxType) * 128;
      mm_tma_load_2d(s_box[threadIdx.x >> 5], tmap, x0 + P.tmaOffX, y0 + P.tmaOffY, bar);
    }
    __syncwarp();
    mm_mbar_wait(bar, 0);
  }
#endif
  int pred[16];
  if (oob)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i) pred[i] = 0;
  }
  else if (fullPel)
  {
    const int16_t* p = ref.plane00[0] + (long)yP * ref.stride[0] + xP;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) pred[r * 4 + c] = __ldg(p + (long)r * ref.stride[0] + c);
  }
#ifndef MM_NO_TMA
  else if (boxType >= 0)
  {
    const uint2* base = reinterpret_cast<const uint2*>(s_box[threadIdx.x >> 5]) + (ay - y0) * pitchQ + ((ax - x0) >> 2);
    mm_fir_luma4x4_wide_core<false, true>(P.fir, __ldg(btaps + (fp.x & 15)), __ldg(btaps + (fp.y & 15)), base, pitchQ, xP - 3, fp.x & 15,
                                          fp.y & 15, P.clpMin, P.clpMax, pred);
  }
#endif
  else
  {
    mm_fir_luma4x4_wide<false>(P.fir, __ldg(btaps + (fp.x & 15)), __ldg(btaps + (fp.y & 15)), ref.plane00[0], ref.stride[0], xP, yP,
                               fp.x & 15, fp.y & 15, P.clpMin, P.clpMax, pred);
  }

  if (MODE == MM_MODE_PRED)
  {
    if (!valid) return;
    int16_t* d = P.dst + (long)(it.y + s.row * 4) * P.dstStride + (it.x + s.col * 4);
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
      const uint32_t lo = (uint32_t)(pred[r * 4 + 0] & 0xffff) | ((uint32_t)pred[r * 4 + 1] << 16);
      const uint32_t hi = (uint32_t)(pred[r * 4 + 2] & 0xffff) | ((uint32_t)pred[r * 4 + 3] << 16);
      *reinterpret_cast<uint2*>(d + (long)r * P.dstStride) = make_uint2(lo, hi);
    }
    return;
  }

  // differences org - pred
  const int16_t* o = orgp;
  int d[16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
  {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(o + (long)r * P.orgStride));
    d[r * 4 + 0] = (int)(int16_t)(v.x & 0xffff) - pred[r * 4 + 0];
    d[r * 4 + 1] = (int)(int16_t)(v.x >> 16) - pred[r * 4 + 1];
    d[r * 4 + 2] = (int)(int16_t)(v.y & 0xffff) - pred[r * 4 + 2];
    d[r * 4 + 3] = (int)(int16_t)(v.y >> 16) - pred[r * 4 + 3];
  }
  unsigned int part;
  if (P.distKind == MMGPU_DIST_SAD)
  {
    // RdCost::xGetSAD*: rows stepped by 1 << subShift, sum << subShift (block rows are 4-aligned)
    unsigned int sad = 0;
    const int step = 1 << P.subShift;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if ((r & (step - 1)) == 0)
        sad += abs(d[r * 4 + 0]) + abs(d[r * 4 + 1]) + abs(d[r * 4 + 2]) + abs(d[r * 4 + 3]);
    part = sad << P.subShift;
  }
  else
  {
    float df[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) df[i] = (float)d[i];
    mm_had4x4_f(df);
    // tile transform across the T adjacent lanes (q = lane & (T-1)); uniform per warp (same block)
    if (s.T >= 2) mm_had_xlane_f(df, 1, (lane & 1) ? -1.0f : 1.0f);
    if (s.T >= 4) mm_had_xlane_f(df, 2, (lane & 2) ? -1.0f : 1.0f);
    if (s.T >= 8) mm_had_xlane_f(df, 4, (lane & 4) ? -1.0f : 1.0f);
    float fsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) fsum += fabsf(df[i]);
    unsigned int sum = (unsigned int)fsum;
    if (s.q == 0)
    {
      const unsigned int dc = (unsigned int)fabsf(df[0]);  // JVET_R0164_MEAN_SCALED_SATD
      sum = sum - dc + (dc >> 2);
    }
    if (s.T >= 2) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (s.T >= 4) sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (s.T >= 8) sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    unsigned int t;
    if (s.tileKind == 0) t = (sum + 1) >> 1;
    else if (s.tileKind == 1) t = (sum + 2) >> 2;
    else if (P.hadMode == MMGPU_HAD_SIMD)
    {
      t = (unsigned int)(((unsigned long long)sum * 0xb504f334ULL) >> 32);
      t = (s.tileKind == 2) ? (t >> 2) : (t >> 1);
    }
    else
    {
      const double div = (s.tileKind == 2) ? 0x1.6a09e667f3bcdp+3 /* sqrt(16.0*8) */ : 0x1.6a09e667f3bcdp+2 /* sqrt(4.0*8) */;
      t = (unsigned int)(int)__dmul_rn(__ddiv_rn((double)sum, div), 2.0);
    }
    part = (s.q == 0) ? t : 0u;
  }
  // candidate reduction
  if (n >= 32)
  {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) atomicAdd(&P.dist[cand], (unsigned long long)part);
  }
  else
  {
    for (int off = n >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (valid && sb == 0) P.dist[cand] = (unsigned long long)part;
  }
}

// -------------------------------------------------------------------------------------------------
// Picture::extendPicBorder (Picture.cpp:1008-1046) and Picture::extendWrapBorder (:1049-1092) on the device: one thread
// per margin sample.  xoff = 0: replicate the border sample; xoff > 0 (PPS wrap-around offset in units of the
// component): the first xoff margin columns on either side continue the picture from the opposite side, the rest
// replicate; rows above / below repeat the (horizontally extended) first / last row in both variants.
// -------------------------------------------------------------------------------------------------
__global__ void k_extend_border(int16_t* plane00, int stride, int w, int h, int mx, int my, int xoff)
{
  const int X = blockIdx.x * blockDim.x + threadIdx.x - mx;
  const int Y = blockIdx.y * blockDim.y + threadIdx.y - my;
  if (X >= w + mx || Y >= h + my) return;
  if (X >= 0 && X < w && Y >= 0 && Y < h) return;
  int sx;
  if (X < 0) sx = (-1 - X < xoff) ? X + xoff : 0;
  else if (X >= w) sx = (X - w < xoff) ? X - xoff : w - 1;
  else sx = X;
  const int sy = min(max(Y, 0), h - 1);
  plane00[(long)Y * stride + X] = plane00[(long)sy * stride + sx];
}

// -------------------------------------------------------------------------------------------------
// k_pad_planes: un-padded planes (e.g. the landing buffer of an all-gather) -> padded reference planes, copy and border
// extension (Picture::extendPicBorder / extendWrapBorder) in ONE launch for up to MM_PAD_MAX planes: blockIdx.z = plane,
// one thread per 8 samples (16 bytes) of a padded row; interior groups are one 16-byte load + store.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pad_planes(MmPadJobs J)
{
  const MmPadPlane p = J.p[blockIdx.z];
  const int gpr = (p.w + 2 * p.mx) >> 3;                       // 8-sample groups per padded row (w, mx multiples of 8)
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = (int)blockIdx.y - p.my;
  if (g >= gpr || (int)blockIdx.y >= p.h + 2 * p.my) return;
  const int X0 = g * 8 - p.mx;
  const int sy = min(max(Y, 0), p.h - 1);
  const int16_t* srow = p.src + (long)sy * p.srcStride;
  int16_t* drow = p.dst00 + (long)Y * p.dstStride + X0;
  uint4 v;
  if (X0 >= 0 && X0 + 8 <= p.w)
    v = *reinterpret_cast<const uint4*>(srow + X0);
  else
  {
    int16_t t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
    {
      const int X = X0 + i;
      int sx;
      if (X < 0) sx = (-1 - X < p.xoff) ? X + p.xoff : 0;
      else if (X >= p.w) sx = (X - p.w < p.xoff) ? X - p.xoff : p.w - 1;
      else sx = X;
      t[i] = srow[sx];
    }
    v.x = (uint16_t)t[0] | ((uint32_t)(uint16_t)t[1] << 16); v.y = (uint16_t)t[2] | ((uint32_t)(uint16_t)t[3] << 16);
    v.z = (uint16_t)t[4] | ((uint32_t)(uint16_t)t[5] << 16); v.w = (uint16_t)t[6] | ((uint32_t)(uint16_t)t[7] << 16);
  }
  *reinterpret_cast<uint4*>(drow) = v;
}

// -------------------------------------------------------------------------------------------------
// k_pull_rows / k_pad_tb: the multi-GPU reference-picture exchange fused with the border extension.  The source planes are the
// OWNERS' un-padded pictures in peer memory (NVLink / NVSwitch mapped addresses; the rank's own picture is a local address):
// each source row crosses NVLink exactly once, as 16-byte loads of a few persistent CTAs (gridDim.x is the SM budget of the
// exchange: the kernels of the current step keep the other SMs), and is written into the slot's padded plane together with its
// left / right margin (Picture::extendPicBorder, replicate).  One warp per (plane, row): 8 loads in flight per lane.  k_pad_tb then
// replicates the first / last padded row into the top / bottom margins from local memory.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 mm_ld_peer(const uint4* p)
{
  uint4 v;
  asm volatile("ld.global.relaxed.sys.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(512) k_pull_rows(MmPadJobs J, int nPlanes, int totalRows)
{
  const int lane = threadIdx.x & 31;
  const int warpsPerGrid = gridDim.x * (blockDim.x >> 5);
  for (int unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < totalRows; unit += warpsPerGrid)
  {
    int k = 0, y = unit;
    while (k < nPlanes - 1 && y >= J.p[k].h) { y -= J.p[k].h; ++k; }
    const MmPadPlane& p = J.p[k];
    const int G = p.w >> 3;
    const uint4* srow = reinterpret_cast<const uint4*>(p.src + (long)y * p.srcStride);
    int16_t* drow = p.dst00 + (long)y * p.dstStride;
    uint32_t first = 0, last = 0;
    for (int base = 0; base < G; base += 256)
    {
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
      {
        const int g = base + lane + 32 * i;
        if (g < G) v[i] = mm_ld_peer(srow + g);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
      {
        const int g = base + lane + 32 * i;
        if (g < G)
        {
          *reinterpret_cast<uint4*>(drow + 8 * g) = v[i];
          if (g == 0) first = v[i].x & 0xffffu;
          if (g == G - 1) last = v[i].w >> 16;
        }
      }
    }
    first = __shfl_sync(0xffffffffu, first, 0);
    last = __shfl_sync(0xffffffffu, last, (G - 1) & 31);
    const uint32_t f2 = first | (first << 16), l2 = last | (last << 16);
    const uint4 fv = make_uint4(f2, f2, f2, f2), lv = make_uint4(l2, l2, l2, l2);
    const int mg = p.mx >> 3;
    for (int m = lane; m < mg; m += 32)
    {
      *reinterpret_cast<uint4*>(drow - p.mx + 8 * m) = fv;
      *reinterpret_cast<uint4*>(drow + p.w + 8 * m) = lv;
    }
  }
}
#define MM_TB_ROWS 16   // margin rows written per unit of k_pad_tb: the source row is loaded once and stored 16 times
__global__ void __launch_bounds__(512) k_pad_tb(MmPadJobs J, int nPlanes, int totalUnits)
{
  const int lane = threadIdx.x & 31;
  const int warpsPerGrid = gridDim.x * (blockDim.x >> 5);
  for (int unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < totalUnits; unit += warpsPerGrid)
  {
    int k = 0, u = unit;
    for (;;)
    {
      const int nu = 2 * ((J.p[k].my + MM_TB_ROWS - 1) / MM_TB_ROWS);
      if (k == nPlanes - 1 || u < nu) break;
      u -= nu;
      ++k;
    }
    const MmPadPlane& p = J.p[k];
    const int chunks = (p.my + MM_TB_ROWS - 1) / MM_TB_ROWS;
    const bool bottom = u >= chunks;
    const int r0 = (bottom ? u - chunks : u) * MM_TB_ROWS, r1 = min(r0 + MM_TB_ROWS, p.my);
    const int sy = bottom ? (p.h - 1) : 0;
    const int Y0 = bottom ? p.h : -p.my;                       // first margin row of the side
    const int G = (p.w + 2 * p.mx) >> 3;
    const uint4* srow = reinterpret_cast<const uint4*>(p.dst00 + (long)sy * p.dstStride - p.mx);
    for (int base = 0; base < G; base += 256)
    {
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
      {
        const int g = base + lane + 32 * i;
        if (g < G) v[i] = srow[g];
      }
      for (int r = r0; r < r1; ++r)
      {
        uint4* drow = reinterpret_cast<uint4*>(p.dst00 + (long)(Y0 + r) * p.dstStride - p.mx);
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
          const int g = base + lane + 32 * i;
          if (g < G) drow[g] = v[i];
        }
      }
    }
  }
}
// k_pull_tma: the same pull with the bulk-copy engine of the SM (TMA, 1-D cp.async.bulk) instead of register loads.  One thread
// per warp runs a ring of `stages` shared-memory rows: bulk load of a row from PEER memory (completion on the stage's mbarrier) ->
// bulk store of the arrived row to its place in the padded plane -> refill once the store has read the stage.  No block-wide
// synchronisation, tens to hundreds of KB in flight per CTA: a handful of CTAs (SMs) carries the whole exchange.  The left /
// right margins are written afterwards from local memory (k_pad_lr), then the top / bottom rows (k_pad_tb).
#define MM_PULL_LAG 2        // bulk stores that may still be reading their stage when the next load is issued
__device__ __forceinline__ void mm_bulk_load(void* smemDst, const void* gsrc, uint32_t bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(mm_smem_u32(smemDst)), "l"(gsrc), "r"(bytes), "r"(mm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mm_bulk_store(void* gdst, const void* smemSrc, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(mm_smem_u32(smemSrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
struct MmRowCursor   // walks the rows blockIdx.x, + gridDim.x, ... of the concatenated planes
{
  int k, y;
  __device__ __forceinline__ void init(const MmPadJobs& J, int nPlanes, int unit)
  {
    k = 0; y = unit;
    while (k < nPlanes - 1 && y >= J.p[k].h) { y -= J.p[k].h; ++k; }
  }
  __device__ __forceinline__ void advance(const MmPadJobs& J, int nPlanes, int step)
  {
    y += step;
    while (k < nPlanes - 1 && y >= J.p[k].h) { y -= J.p[k].h; ++k; }
  }
};
#define MM_PULL_RINGS 8      // independent rings (warps, lane 0 active) per CTA: the per-thread issue latency of the bulk operations
                             // (~1000 cycles per row measured with one ring), not the copy engine, limits a single ring
__global__ void __launch_bounds__(32 * MM_PULL_RINGS) k_pull_tma(const __grid_constant__ MmPadJobs J, int nPlanes, int totalRows, int stages,
                                                                  int stageBytes)
{
  extern __shared__ __align__(128) unsigned char s_dyn[];
  if ((threadIdx.x & 31) != 0) return;
  const int ring = threadIdx.x >> 5;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(s_dyn) + ring * stages;          // [rings][stages]
  unsigned char* buf = s_dyn + ((MM_PULL_RINGS * stages * 8 + 127) & ~127) + (size_t)ring * stages * stageBytes;
  for (int i = 0; i < stages; ++i) mm_mbar_init(&bars[i], 1);
  const int first = blockIdx.x * MM_PULL_RINGS + ring, step = gridDim.x * MM_PULL_RINGS;
  const int nMine = (totalRows - first + step - 1) / step;
  if (nMine <= 0) return;
  MmRowCursor ld, st;
  ld.init(J, nPlanes, first);
  st = ld;
  int issued = 0;
  for (; issued < min(stages, nMine); ++issued)
  {
    const uint32_t bytes = (uint32_t)J.p[ld.k].w * 2u;
    mm_mbar_expect_tx(&bars[issued], bytes);
    mm_bulk_load(buf + (size_t)issued * stageBytes, J.p[ld.k].src + (long)ld.y * J.p[ld.k].srcStride, bytes, &bars[issued]);
    ld.advance(J, nPlanes, step);
  }
  for (int j = 0; j < nMine; ++j)
  {
    const int sg = j % stages;
    mm_mbar_wait(&bars[sg], (uint32_t)((j / stages) & 1));
    mm_bulk_store(J.p[st.k].dst00 + (long)st.y * J.p[st.k].dstStride, buf + (size_t)sg * stageBytes, (uint32_t)J.p[st.k].w * 2u);
    st.advance(J, nPlanes, step);
    // refill the stage of row j - LAG: its bulk store has read it once at most LAG younger stores are pending
    if (j >= MM_PULL_LAG && issued < nMine)
    {
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(MM_PULL_LAG) : "memory");
      const int s2 = (j - MM_PULL_LAG) % stages;
      const uint32_t bytes = (uint32_t)J.p[ld.k].w * 2u;
      mm_mbar_expect_tx(&bars[s2], bytes);
      mm_bulk_load(buf + (size_t)s2 * stageBytes, J.p[ld.k].src + (long)ld.y * J.p[ld.k].srcStride, bytes, &bars[s2]);
      ld.advance(J, nPlanes, step);
      ++issued;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA (and its shared memory) goes away
}
// left / right margins of every row from the row's own edge samples (local memory): a warp takes 32 rows, lane r reads the two
// edge samples of row r, then the warp writes the 32 x 2 x (mx / 8) margin groups
__global__ void __launch_bounds__(512) k_pad_lr(MmPadJobs J, int nPlanes, int totalUnits)
{
  const int lane = threadIdx.x & 31;
  const int warpsPerGrid = gridDim.x * (blockDim.x >> 5);
  for (int unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); unit < totalUnits; unit += warpsPerGrid)
  {
    int k = 0, u = unit;
    for (;;)
    {
      const int nu = (J.p[k].h + 31) >> 5;
      if (k == nPlanes - 1 || u < nu) break;
      u -= nu;
      ++k;
    }
    const MmPadPlane& p = J.p[k];
    const int y = u * 32 + lane;
    uint32_t first = 0, last = 0;
    if (y < p.h)
    {
      const int16_t* row = p.dst00 + (long)y * p.dstStride;
      first = (uint16_t)row[0];
      last = (uint16_t)row[p.w - 1];
    }
    const int mg = p.mx >> 3, rows = min(32, p.h - u * 32);
    for (int r = 0; r < rows; ++r)
    {
      const uint32_t f = __shfl_sync(0xffffffffu, first, r), l = __shfl_sync(0xffffffffu, last, r);
      const uint32_t f2 = f | (f << 16), l2 = l | (l << 16);
      int16_t* drow = p.dst00 + (long)(u * 32 + r) * p.dstStride;
      for (int m = lane; m < 2 * mg; m += 32)
      {
        if (m < mg) *reinterpret_cast<uint4*>(drow - p.mx + 8 * m) = make_uint4(f2, f2, f2, f2);
        else *reinterpret_cast<uint4*>(drow + p.w + 8 * (m - mg)) = make_uint4(l2, l2, l2, l2);
      }
    }
  }
}

// k_multicast_copy: the PUSH form of the exchange.  dst is a MULTICAST address of symmetric memory (NVLink SHARP / NVLS: one store
// is replicated by the NVSwitch into the same offset of every GPU's buffer, the writer's included): the owner of a picture reads
// it once from local memory and writes it once; the switch fans it out.  Per GPU and step that is one picture of SM work and
// NVLink egress instead of N - 1 pictures pulled or all-gathered.
__global__ void __launch_bounds__(512) k_multicast_copy(const uint4* __restrict__ src, uint4* dst, long n16)
{
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += 4 * stride)
  {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i + k * stride < n16) v[k] = __ldg(src + i + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i + k * stride < n16)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i + k * stride), "f"(__uint_as_float(v[k].x)),
                     "f"(__uint_as_float(v[k].y)), "f"(__uint_as_float(v[k].z)), "f"(__uint_as_float(v[k].w)) : "memory");
  }
  __threadfence_system();
}
void mmk_multicast_copy(cudaStream_t s, const void* src, void* dstMulticast, size_t bytes, int maxCtas)
{
  k_multicast_copy<<<maxCtas, 512, 0, s>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dstMulticast), (long)(bytes >> 4));
}

void mmk_pull_pad_planes(cudaStream_t s, const MmPadJobs& J, int nPlanes, int maxCtas)
{
  int rows = 0, units = 0, maxW = 0;
  for (int k = 0; k < nPlanes; ++k)
  {
    rows += J.p[k].h;
    units += 2 * ((J.p[k].my + MM_TB_ROWS - 1) / MM_TB_ROWS);
    maxW = max(maxW, J.p[k].w);
  }
  static int mode = -1;   // MMGPU_PULL_KERNEL=ld: register-load variant (k_pull_rows); default: TMA bulk copies
  if (mode < 0) { const char* e = getenv("MMGPU_PULL_KERNEL"); mode = (e && e[0] == 'l') ? 1 : 0; }
  if (mode == 1) k_pull_rows<<<maxCtas, 512, 0, s>>>(J, nPlanes, rows);
  else
  {
    const int stageBytes = (maxW * 2 + 127) & ~127;
    static int wantStages = 0;   // MMGPU_PULL_STAGES: rows in flight per CTA (default 48: ~200 KB of shared memory for 2048-wide luma)
    if (!wantStages) { const char* e = getenv("MMGPU_PULL_STAGES"); wantStages = e ? atoi(e) : 48; if (wantStages <= 0) wantStages = 48; }
    int stages = min(wantStages, (200 * 1024) / stageBytes) / MM_PULL_RINGS;     // per ring
    stages = max(2 * MM_PULL_LAG, min(stages, 16));
    const int smem = ((MM_PULL_RINGS * stages * 8 + 127) & ~127) + MM_PULL_RINGS * stages * stageBytes;
    cudaFuncSetAttribute(k_pull_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);   // per device: set on every call
    k_pull_tma<<<maxCtas, 32 * MM_PULL_RINGS, smem, s>>>(J, nPlanes, rows, stages, stageBytes);
    int lrUnits = 0;
    for (int k = 0; k < nPlanes; ++k) lrUnits += (J.p[k].h + 31) >> 5;
    k_pad_lr<<<maxCtas, 512, 0, s>>>(J, nPlanes, lrUnits);
  }
  k_pad_tb<<<maxCtas, 512, 0, s>>>(J, nPlanes, units);
}
__global__ void k_add_avg(const int16_t* s0, int st0, const int16_t* s1, int st1, int16_t* dst, int ds, int w, int h,
                          int shift, int offset, int lo, int hi)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(long)y * ds + x] = (int16_t)mm_clip(((int)s0[(long)y * st0 + x] + (int)s1[(long)y * st1 + x] + offset) >> shift, lo, hi);
}

// ---- launch wrappers -----------------------------------------------------------------------------
void mmk_fir(cudaStream_t s, const MmParams& P, int first, int count, int mode, const int2* pos)
{
  if (count <= 0) return;
  const int grid = (count + MM_FIR_WARPS - 1) / MM_FIR_WARPS;
  if (mode == MM_MODE_DIST) k_fir<MM_MODE_DIST><<<grid, 32 * MM_FIR_WARPS, 0, s>>>(P, first, count, pos);
  else k_fir<MM_MODE_PRED><<<grid, 32 * MM_FIR_WARPS, 0, s>>>(P, first, count, pos);
}
void mmk_extend_border(cudaStream_t s, int16_t* plane00, int stride, int w, int h, int mx, int my, int xoff)
{
  dim3 tb(32, 8), gr((w + 2 * mx + 31) / 32, (h + 2 * my + 7) / 8);
  k_extend_border<<<gr, tb, 0, s>>>(plane00, stride, w, h, mx, my, xoff);
}
void mmk_pad_planes(cudaStream_t s, const MmPadJobs& J, int nPlanes, int maxGroups, int maxRows)
{
  dim3 gr((maxGroups + 255) / 256, maxRows, nPlanes);
  k_pad_planes<<<gr, 256, 0, s>>>(J);
}
void mmk_add_avg(cudaStream_t s, const int16_t* s0, int st0, const int16_t* s1, int st1, int16_t* dst, int ds, int w, int h,
                 int shift, int offset, int lo, int hi)
{
  dim3 tb(32, 4), gr((w + 31) / 32, (h + 3) / 4);
  k_add_avg<<<gr, tb, 0, s>>>(s0, st0, s1, st1, dst, ds, w, h, shift, offset, lo, hi);
}
