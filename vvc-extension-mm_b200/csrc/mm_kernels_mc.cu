// mm_kernels_mc.cu -- block-level entry points that map and interpolate in one kernel: single-block MC of any
// component (parity twin of xPredInterBlkMM / xPredInterBlk), whole-picture reconstruction jobs, DMVR costs.
#include "mm_sbpos.cuh"
#include "mm_fir.cuh"

// Classic MVs come clipped from the reference (clipMv before xPredInterBlk); an unclipped one is held inside the padded
// plane, for memory safety only (margin minus the filter reach).
__device__ __forceinline__ void mm_classic_keep_inside(const MmParams& P, int csx, int csy, int shH, int shV, int* fx, int* fy)
{
  const int mgx = (P.tmaOffX >> csx) - 8, mgy = (P.tmaOffY >> csy) - 8;
  *fx = min(max(*fx, -mgx << shH), ((P.picW >> csx) + mgx) << shH);
  *fy = min(max(*fy, -mgy << shV), ((P.picH >> csy) + mgy) << shV);
}

// xPredInterBlk (InterPrediction.cpp:968-971): chroma phase = mv * (1 << (1 - scale)) & 31 -- the chroma filter table always has
// 32 phases, so the 1/16 fraction of an un-subsampled direction (4:4:4, 4:2:2 vertical) is doubled.  xPredInterBlkMM passes
// its raw fraction (InterPrediction.cpp:730-733), so this applies to the classic model only.
__device__ __forceinline__ void mm_classic_chroma_phase(int model, int comp, int csx, int csy, int* xF, int* yF)
{
  if (model == MM_CLASSIC && comp != 0)
  {
    *xF <<= (1 - csx);
    *yF <<= (1 - csy);
  }
}

// One thread per sub-block of ONE block (mmgpu_mc).  model == CLASSIC: one MV for the block (xPredInterBlk).
__global__ void k_mc_generic(MmParams P, int refSlot, int comp, int csx, int csy, int x, int y, int w, int h, mmgpu_mv mv,
                             int model, int rotSlot, int bi, int16_t* dst, int dstStride)
{
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = h / sbh, cols = w / sbw;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int col = i / rows, row = i - col * rows;
  const MmRef ref = P.refs[refSlot];
  const int shH = 4 + csx, shV = 4 + csy;
  int fx, fy;
  int16_t* d = dst + (long)row * sbh * dstStride + col * sbw;
  if (model == MM_CLASSIC)
  {
    // xPredInterBlk (InterPrediction.cpp:928-1096, no wrap / RPR): every sample uses the block MV
    fx = ((x + col * sbw) << shH) + mv.hor;
    fy = ((y + row * sbh) << shV) + mv.ver;
    mm_classic_keep_inside(P, csx, csy, shH, shV, &fx, &fy);
  }
  else
  {
    const float* R = mm_rot_ptr(P.rot, model, rotSlot);
    MmCand c;
    mm_cand_setup(P.g, model, comp == 0, x, y, w, h, mv.hor, mv.ver, R, &c);
    mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, comp, csx, csy, x, y, w, h, c, R, row, col, &fx, &fy);
    const int xP = fx >> shH, yP = fy >> shV;
    const int picw = P.picW >> csx, pich = P.picH >> csy;
    const int mcw = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csx);
    const int mch = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csy);
    if (mm_oob(xP, yP, picw, pich, mcw, mch, sbw, sbh))
    {
      for (int r = 0; r < sbh; ++r)
        for (int q = 0; q < sbw; ++q) d[r * dstStride + q] = 0;
      return;
    }
  }
  const int xP = fx >> shH, yP = fy >> shV;
  int xF = fx & ((1 << shH) - 1), yF = fy & ((1 << shV) - 1);
  mm_classic_chroma_phase(model, comp, csx, csy, &xF, &yF);
  if (comp == 0)
    mm_fir_generic<8>(P.fir, ref.plane00[0], ref.stride[0], xP, yP, xF, yF, sbw, sbh, !bi, P.clpMin, P.clpMax, d, dstStride);
  else
    mm_fir_generic<4>(P.fir, ref.plane00[comp], ref.stride[comp], xP, yP, xF, yF, sbw, sbh, !bi, P.clpMin, P.clpMax, d,
                      dstStride);
}

// One sub-block of one component into `out` (row stride 4): mapping (or classic MV) + OOB rule + FIR.
__device__ __forceinline__ void mm_mc_subblock(const MmParams& P, int refSlot, int comp, int csx, int csy, int x, int y, int w,
                                               int h, mmgpu_mv mv, int model, int rotSlot, bool bi, int row, int col,
                                               const MmCand& c, int16_t* out)
{
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const MmRef ref = P.refs[refSlot];
  const int shH = 4 + csx, shV = 4 + csy;
  int fx, fy;
  if (model == MM_CLASSIC)
  {
    fx = ((x + col * sbw) << shH) + mv.hor;
    fy = ((y + row * sbh) << shV) + mv.ver;
    mm_classic_keep_inside(P, csx, csy, shH, shV, &fx, &fy);
  }
  else
  {
    const float* R = mm_rot_ptr(P.rot, model, rotSlot);
    mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, model, comp, csx, csy, x, y, w, h, c, R, row, col, &fx, &fy);
    const int xP = fx >> shH, yP = fy >> shV;
    const int picw = P.picW >> csx, pich = P.picH >> csy;
    const int mcw = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csx);
    const int mch = (P.oobMode == MMGPU_OOB_ME) ? 0 : (P.maxCU >> csy);
    if (mm_oob(xP, yP, picw, pich, mcw, mch, sbw, sbh))
    {
      for (int i = 0; i < 16; ++i) out[i] = 0;
      return;
    }
  }
  const int xP = fx >> shH, yP = fy >> shV;
  int xF = fx & ((1 << shH) - 1), yF = fy & ((1 << shV) - 1);
  mm_classic_chroma_phase(model, comp, csx, csy, &xF, &yF);
  if (comp == 0)
  {
    // luma: packed dp2a two-pass FIR (same identity as k_fir; the bi variant keeps the 14-bit intermediate like k_dmvr)
    const uint2* btaps = reinterpret_cast<const uint2*>(P.hTaps + MM_HTAPS_WORDS);
    int pred[16];
    if (bi) mm_fir_luma4x4_wide<true>(P.fir, __ldg(btaps + xF), __ldg(btaps + yF), ref.plane00[0], ref.stride[0], xP, yP, xF, yF, P.clpMin, P.clpMax, pred);
    else mm_fir_luma4x4_wide<false>(P.fir, __ldg(btaps + xF), __ldg(btaps + yF), ref.plane00[0], ref.stride[0], xP, yP, xF, yF, P.clpMin, P.clpMax, pred);
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = (int16_t)pred[i];
  }
  else
    mm_fir_generic<4>(P.fir, ref.plane00[comp], ref.stride[comp], xP, yP, xF, yF, sbw, sbh, !bi, P.clpMin, P.clpMax, out, 4);
}

// Whole-picture reconstruction shape (DecCu::xReconInter -> motionCompensation -> xPredInterUni / xPredInterBi,
// InterPrediction.cpp:455-681): one CTA per job (one component of one PU); uni-prediction writes the clipped
// prediction, bi-prediction interpolates both lists at 14 bit and applies PelBuf::addAvg (Buffer.cpp:551-587) in
// registers before the store.  dst planes are un-padded device pictures.
// items[i] = (job index, first sub-block): one CTA of MM_JOB_THREADS threads per chunk of MM_JOB_THREADS sub-blocks of a job, so
// that a 128x128 PU is spread over 16 CTAs instead of being walked by one.
#define MM_JOB_THREADS 64
__global__ void __launch_bounds__(MM_JOB_THREADS) k_mc_jobs(MmParams P, const mmgpu_mc_job* __restrict__ jobs, const int2* __restrict__ items,
                                                            int16_t* dstY, int16_t* dstCb, int16_t* dstCr, int strideY, int strideC, int csxC,
                                                            int csyC, int avgShift, int avgOffset)
{
  const int2 it = items[blockIdx.x];
  const mmgpu_mc_job j = jobs[it.x];
  const int csx = j.comp ? csxC : 0, csy = j.comp ? csyC : 0;
  const int sbw = 4 >> csx, sbh = 4 >> csy;
  const int rows = j.h / sbh, cols = j.w / sbw;
  int16_t* plane = j.comp == 0 ? dstY : (j.comp == 1 ? dstCb : dstCr);
  const int stride = j.comp == 0 ? strideY : strideC;
  // per-(block, MV) scalars of the motion model: once per job and list, not once per sub-block
  __shared__ MmCand s_c[2];
  if (j.model != MM_CLASSIC && threadIdx.x < (j.bi ? 2 : 1))
  {
    const int l = threadIdx.x;
    const mmgpu_mv mv = l ? j.mv1 : j.mv0;
    mm_cand_setup(P.g, j.model, j.comp == 0, j.x, j.y, j.w, j.h, mv.hor, mv.ver, mm_rot_ptr(P.rot, j.model, l ? j.rot_slot1 : j.rot_slot0), &s_c[l]);
  }
  __syncthreads();
  const int i = it.y + threadIdx.x;
  if (i < rows * cols)
  {
    const int col = i / rows, row = i - col * rows;
    int16_t a[16], b[16];
    mm_mc_subblock(P, j.ref_slot0, j.comp, csx, csy, j.x, j.y, j.w, j.h, j.mv0, j.model, j.rot_slot0, j.bi != 0, row, col, s_c[0], a);
    if (j.bi)
    {
      mm_mc_subblock(P, j.ref_slot1, j.comp, csx, csy, j.x, j.y, j.w, j.h, j.mv1, j.model, j.rot_slot1, true, row, col, s_c[1], b);
      for (int r = 0; r < sbh; ++r)
        for (int q = 0; q < sbw; ++q)
          a[r * 4 + q] = (int16_t)mm_clip(((int)a[r * 4 + q] + (int)b[r * 4 + q] + avgOffset) >> avgShift, P.clpMin, P.clpMax);
    }
    int16_t* d = plane + (long)(j.y + row * sbh) * stride + (j.x + col * sbw);
    for (int r = 0; r < sbh; ++r)
      for (int q = 0; q < sbw; ++q) d[(long)r * stride + q] = a[r * 4 + q];
  }
}

// -------------------------------------------------------------------------------------------------
// k_dmvr: xProcessDMVRProjected's cost evaluations (InterPrediction.cpp:2510-2567).  One group of 2n lanes per
// (sub-PU, search offset): n lanes interpolate the L0 prediction of the sub-PU's n sub-blocks at mv0 + offset,
// n lanes the L1 prediction at mv1 - offset (bi = true: 14-bit intermediates, +-maxCU OOB rule, the sub-PU is
// "the block": quirk Q6); xDMVRCost = SAD over even rows (subShift 1), << 1, >> 1.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dmvr(const MmParams P, const mmgpu_dmvr_subpu* __restrict__ subpus,
                                              const MmDmvrItem* __restrict__ items, int nItems, int simdSad,
                                              unsigned long long* __restrict__ costs)
{
  __shared__ MmTaps s_taps[16];
  mm_build_taps(s_taps);
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nItems) return;
  const MmDmvrItem it = items[warp];
  const mmgpu_dmvr_subpu sp = subpus[it.subpu];
  const int cols = sp.w >> 2, rows = sp.h >> 2;
  const int n = cols * rows;                    // 4, 8 or 16 sub-blocks
  const int g = lane / (2 * n), within = lane - g * 2 * n;
  const int list = within / n, sb = within - list * n;
  const int off = it.off + g;
  const bool valid = off < 25;
  const int offc = valid ? off : it.off;
  const int oh = (offc % 5 - 2) * 16, ov = (offc / 5 - 2) * 16;   // m_pSearchOffset << MV_FRACTIONAL_BITS_INTERNAL
  const int mh = list ? (sp.mv1.hor - oh) : (sp.mv0.hor + oh);
  const int mv = list ? (sp.mv1.ver - ov) : (sp.mv0.ver + ov);
  const int row = sb / cols, col = sb - row * cols;
  const MmRef ref = P.refs[list ? sp.ref_slot1 : sp.ref_slot0];
  const float* R = mm_rot_ptr(P.rot, sp.model, list ? sp.rot_slot1 : sp.rot_slot0);
  MmCand c;
  mm_cand_setup(P.g, sp.model, true, sp.x, sp.y, sp.w, sp.h, mh, mv, R, &c);
  int fx, fy;
  mm_subblock_pos(P.g, P.mpa, P.mpaVip, P.thTab, P.phTab, sp.model, 0, 0, 0, sp.x, sp.y, sp.w, sp.h, c, R, row, col, &fx, &fy);
  const int xP = fx >> 4, yP = fy >> 4;
  const bool oob = mm_oob(xP, yP, P.picW, P.picH, P.maxCU, P.maxCU, 4, 4);
  int pred[16];
  if (oob)
  {
#pragma unroll
    for (int i = 0; i < 16; ++i) pred[i] = 0;
  }
  else
  {
    mm_fir_luma4x4<true>(P.fir, s_taps, ref.plane00[0], ref.stride[0], xP, yP, fx & 15, fy & 15, P.clpMin, P.clpMax, pred);
  }
  // rows 0 and 2 of the sub-block (even rows of the sub-PU): L0 lanes receive the L1 samples of lane ^ n
  unsigned int sad = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
  {
    const int mine = pred[(i >> 2) * 8 + (i & 3)];
    const int other = __shfl_xor_sync(0xffffffffu, mine, n);
    int d = mine - other;
    if (simdSad) { d = (int)(int16_t)d; d = abs(d) & 0xffff; }   // _mm_sub_epi16 + _mm_abs_epi16, zero-extended
    else d = abs(d);
    sad += (unsigned int)d;
  }
  if (list) sad = 0;
  for (int o = n >> 1; o > 0; o >>= 1) sad += __shfl_xor_sync(0xffffffffu, sad, o);
  if (valid && within == 0) costs[(long)it.subpu * 25 + off] = ((unsigned long long)sad << 1) >> 1;
}

// Ring of integer-position samples around a classic bi-predicted luma block (xPredInterBlk, InterPrediction.cpp:1054-1093).
// tmp: (w + 4) x (h + 4) buffer; ring sample (i, j), i in [0, w+1], j in [0, h+1] on the border of the (w+2) x (h+2) box, is
// stored at tmp[(j + 1) * (w + 4) + (i + 1)] and comes from reference sample (x0 + i - xOff, y0 + j - yOff).
__global__ void k_bdof_ring(const int16_t* __restrict__ ref, int stride, int x0, int y0, int xOff, int yOff, int w, int h, int shift,
                            int16_t* __restrict__ tmp)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 2 * (w + 2) + 2 * h;
  if (t >= n) return;
  int i, j;
  if (t < w + 2) { i = t; j = 0; }
  else if (t < 2 * (w + 2)) { i = t - (w + 2); j = h + 1; }
  else { const int k = t - 2 * (w + 2); j = 1 + (k >> 1); i = (k & 1) ? (w + 1) : 0; }
  const int v = ref[(long)(y0 + j - yOff) * stride + (x0 + i - xOff)];
  tmp[(j + 1) * (w + 4) + (i + 1)] = (int16_t)((int16_t)(v << shift) - (int16_t)8192);
}
void mmk_bdof_ring(cudaStream_t s, const int16_t* ref, int stride, int x0, int y0, int xOff, int yOff, int w, int h, int shift, int16_t* tmp)
{
  const int n = 2 * (w + 2) + 2 * h;
  k_bdof_ring<<<(n + 127) / 128, 128, 0, s>>>(ref, stride, x0, y0, xOff, yOff, w, h, shift, tmp);
}

// ---- launch wrappers -----------------------------------------------------------------------------
void mmk_mc_generic(cudaStream_t s, const MmParams& P, int refSlot, int comp, int csx, int csy, int x, int y, int w, int h,
                    mmgpu_mv mv, int model, int rotSlot, int bi, int16_t* dst, int dstStride)
{
  const int n = (w / (4 >> csx)) * (h / (4 >> csy));
  k_mc_generic<<<(n + 63) / 64, 64, 0, s>>>(P, refSlot, comp, csx, csy, x, y, w, h, mv, model, rotSlot, bi, dst, dstStride);
}
void mmk_mc_jobs(cudaStream_t s, const MmParams& P, int nItems, const mmgpu_mc_job* jobs, const int2* items, int16_t* dstY, int16_t* dstCb,
                 int16_t* dstCr, int strideY, int strideC, int csxC, int csyC, int avgShift, int avgOffset)
{
  k_mc_jobs<<<nItems, MM_JOB_THREADS, 0, s>>>(P, jobs, items, dstY, dstCb, dstCr, strideY, strideC, csxC, csyC, avgShift, avgOffset);
}
void mmk_dmvr(cudaStream_t s, const MmParams& P, const mmgpu_dmvr_subpu* subpus, const MmDmvrItem* items, int nItems,
              int simdSad, unsigned long long* costs)
{
  k_dmvr<<<(nItems + 3) / 4, 128, 0, s>>>(P, subpus, items, nItems, simdSad, costs);
}
