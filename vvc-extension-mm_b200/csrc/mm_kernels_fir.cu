// mm_kernels_fir.cu -- integer side of the path: 8-tap separable interpolation at the positions k_map wrote, then SAD /
// Hadamard SATD per candidate or prediction write-out (a5); border extension and addAvg.
#include "mm_fir.cuh"

// -------------------------------------------------------------------------------------------------
// k_fir: one warp per work item (32 sub-block-candidates of one block), one 4x4 sub-block per lane.
// -------------------------------------------------------------------------------------------------
#ifndef MM_FIR_MINBLOCKS
#define MM_FIR_MINBLOCKS 8
#endif
#ifndef MM_STAGE_CAP
#define MM_STAGE_CAP 0     // samples of the per-warp shared-memory boxes (2560 = 5 KB); 0 = gather from global memory.
                           // Measured: staging cuts L1 wavefronts but its extra instructions cost more.
#endif
#ifndef MM_STAGE_MIN_LGN
#define MM_STAGE_MIN_LGN 4  // smallest group (2^4 = 16 sub-blocks of one candidate) that is staged
#endif
template <int MODE>
__global__ void __launch_bounds__(128, MM_FIR_MINBLOCKS) k_fir(const MmParams P, const int2* __restrict__ pos)
{
  __shared__ MmTaps s_taps[16];
#if MM_STAGE_CAP > 0
  __shared__ uint4 s_stage[4][MM_STAGE_CAP / 8];
#endif
#ifdef MM_HTAPS_SMEM
  __shared__ __align__(16) uint32_t s_htaps[MM_HTAPS_WORDS];
  for (int i = threadIdx.x; i < MM_HTAPS_WORDS / 4; i += blockDim.x)
    reinterpret_cast<uint4*>(s_htaps)[i] = __ldg(reinterpret_cast<const uint4*>(P.hTaps) + i);
#else
  const uint32_t* s_htaps = P.hTaps;   // 3 KB table, L1 resident: no per-CTA prologue, no barrier
#endif
#if MM_STAGE_CAP > 0 || defined(MM_HTAPS_SMEM)
  mm_build_taps(s_taps);
  __syncthreads();
#endif
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= P.nItems) return;
  const MmItem it = P.items[warp];
  const int2 fp = __ldg(&pos[(long)warp * 32 + lane]);
  const int bw = it.wh & 0xffff, bh = it.wh >> 16;
  const int n = (bw >> 2) * (bh >> 2);
  int cand, sb;
  if (n >= 32) { cand = it.cand; sb = it.sb + lane; }
  else { cand = it.cand + lane / n; sb = lane % n; }
  const bool valid = cand < it.candEnd;
  const MmSub s = mm_sub_from_index(bw, bh, sb);
  const MmRef ref = P.refs[it.rmr & 0xffff];

  // The windows of a group of lanes (11 rows x 12 samples each, neighbours overlapping by 2/3) are read once from global
  // memory into a shared-memory box with 16-byte loads and gathered from there.  A group = the 32 sub-blocks of the warp,
  // or the n sub-blocks of one candidate of a small block (n = 16, 8, ..).  Groups whose windows do not fit their box
  // (strongly diverging motion) gather from global memory directly.
  const bool oob = fp.x == MM_POS_OOB;
  const int xP = fp.x >> 4, yP = fp.y >> 4;
  const int ax = (xP - 3) & ~1, ay = yP - 3;
  bool staged = false;
  int x0 = 0, y0 = 0, pitchW = 0;
#if MM_STAGE_CAP > 0
  const int lgn = (n >= 32) ? 5 : (31 - __clz(n));          // lanes per group = 2^lgn
  const int li = lane & ((1 << lgn) - 1);
  uint32_t* sbox = reinterpret_cast<uint32_t*>(s_stage[threadIdx.x >> 5]) + (lane >> lgn) * ((MM_STAGE_CAP / 2) >> (5 - lgn));
  if (lgn >= MM_STAGE_MIN_LGN)                                // warp-uniform
  {
    const unsigned gmask = (lgn == 5) ? 0xffffffffu : (((1u << (1 << lgn)) - 1u) << (lane & ~((1 << lgn) - 1)));
    const int BIG = 1 << 28;
    x0 = __reduce_min_sync(gmask, oob ? BIG : ax) & ~7;
    y0 = __reduce_min_sync(gmask, oob ? BIG : ay);
    const int x1 = __reduce_max_sync(gmask, oob ? -BIG : ax + 12);
    const int y1 = __reduce_max_sync(gmask, oob ? -BIG : ay + 11);
    const int cpr = (x1 - x0 + 7) >> 3;        // 16-byte chunks per box row
    const int hb = y1 - y0;
    pitchW = ((4 * cpr + 5) & ~7) + 2;         // words; = 2 mod 8, >= 4*cpr
    const int lg = 32 - __clz(cpr - 1);        // lanes form rows of 2^lg >= cpr chunks
    const bool aligned = ((reinterpret_cast<uintptr_t>(ref.plane00[0]) | (uintptr_t)(ref.stride[0] * 2)) & 15) == 0;
    staged = aligned && cpr > 0 && hb > 0 && lg <= lgn && mm_stage_row(hb, pitchW) <= ((MM_STAGE_CAP / 2) >> (5 - lgn));
    if (staged)
    {
      const int lc = li & ((1 << lg) - 1), lr = li >> lg, rpp = (1 << lgn) >> lg;
      const int16_t* g = ref.plane00[0] + (long)y0 * ref.stride[0] + x0;
      if (lc < cpr)
        for (int row = lr; row < hb; row += rpp)
        {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(g + (long)row * ref.stride[0]) + lc);
          uint32_t* d = sbox + mm_stage_row(row, pitchW) + 4 * lc;
          d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    __syncwarp();
  }
#endif
  int pred[16];
  if (oob)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i) pred[i] = 0;
  }
#if MM_STAGE_CAP > 0
  else if (staged)
  {
    mm_fir_luma4x4_core<false, true>(P.fir, s_taps, sbox + ((ax - x0) >> 1), pitchW, ay - y0, ((xP - 3) & 1) << 4, fp.x & 15,
                                     fp.y & 15, P.clpMin, P.clpMax, pred);
  }
#endif
  else
  {
    mm_fir_luma4x4_wide<false>(P.fir, s_htaps, ref.plane00[0], ref.stride[0], xP, yP, fp.x & 15, fp.y & 15, P.clpMin, P.clpMax,
                               pred);
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
  const int16_t* o = P.org + (long)((int)(it.orgxy >> 16) + s.row * 4) * P.orgStride + ((int)(it.orgxy & 0xffff) + s.col * 4);
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
#ifndef MM_ABL_NOHAD   // ablation build only: no 4x4 transform / cross-lane butterflies
    mm_had4x4(d);
#endif
    // tile transform across the T adjacent lanes (q = lane & (T-1)); uniform per warp (same block)
#ifndef MM_ABL_NOHAD
    if (s.T >= 2) mm_had_xlane(d, 1, (lane & 1) != 0);
    if (s.T >= 4) mm_had_xlane(d, 2, (lane & 2) != 0);
    if (s.T >= 8) mm_had_xlane(d, 4, (lane & 4) != 0);
#endif
    unsigned int sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) sum += (unsigned int)abs(d[i]);
    if (s.q == 0)
    {
      const unsigned int dc = (unsigned int)abs(d[0]);  // JVET_R0164_MEAN_SCALED_SATD
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
// Picture::extendPicBorder on the device (Picture.cpp:1008-1046): one thread per margin sample
// -------------------------------------------------------------------------------------------------
__global__ void k_extend_border(int16_t* plane00, int stride, int w, int h, int mx, int my)
{
  const int X = blockIdx.x * blockDim.x + threadIdx.x - mx;
  const int Y = blockIdx.y * blockDim.y + threadIdx.y - my;
  if (X >= w + mx || Y >= h + my) return;
  if (X >= 0 && X < w && Y >= 0 && Y < h) return;
  const int sx = min(max(X, 0), w - 1), sy = min(max(Y, 0), h - 1);
  plane00[(long)Y * stride + X] = plane00[(long)sy * stride + sx];
}

__global__ void k_add_avg(const int16_t* s0, int st0, const int16_t* s1, int st1, int16_t* dst, int ds, int w, int h,
                          int shift, int offset, int lo, int hi)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  dst[(long)y * ds + x] = (int16_t)mm_clip(((int)s0[(long)y * st0 + x] + (int)s1[(long)y * st1 + x] + offset) >> shift, lo, hi);
}

// ---- launch wrappers -----------------------------------------------------------------------------
void mmk_fir(cudaStream_t s, const MmParams& P, int mode, const int2* pos)
{
  if (P.nItems <= 0) return;
  const int grid = (P.nItems + 3) / 4;
  if (mode == MM_MODE_DIST) k_fir<MM_MODE_DIST><<<grid, 128, 0, s>>>(P, pos);
  else k_fir<MM_MODE_PRED><<<grid, 128, 0, s>>>(P, pos);
}
void mmk_extend_border(cudaStream_t s, int16_t* plane00, int stride, int w, int h, int mx, int my)
{
  dim3 tb(32, 8), gr((w + 2 * mx + 31) / 32, (h + 2 * my + 7) / 8);
  k_extend_border<<<gr, tb, 0, s>>>(plane00, stride, w, h, mx, my);
}
void mmk_add_avg(cudaStream_t s, const int16_t* s0, int st0, const int16_t* s1, int st1, int16_t* dst, int ds, int w, int h,
                 int shift, int offset, int lo, int hi)
{
  dim3 tb(32, 4), gr((w + 31) / 32, (h + 3) / 4);
  k_add_avg<<<gr, tb, 0, s>>>(s0, st0, s1, st1, dst, ds, w, h, shift, offset, lo, hi);
}
