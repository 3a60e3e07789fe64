// mm_kernels_dmvr.cu -- classic DMVR (InterPrediction::xProcessDMVR, InterPrediction.cpp:2239-2439) on the device.
//
// The reference prefetches a (w + taps - 1) x (h + taps - 1) window of the reference picture into a small buffer
// (xPrefetch :1891-1936), pads it by replication (xPad :1938-1956) and then filters FROM THAT BUFFER: the bilinear search
// predictions (xinitMC :2177-2223) and, after the decision, the final 8-tap / 4-tap MC of the moved sub-PU
// (xFinalPaddedMCForDMVR :2075-2145).  Here the buffer never exists: a sample of it is the reference plane read at
// coordinates clamped to the window, staged once per CTA in shared memory.
#include "mm_common.cuh"

static __constant__ int8_t c_lumaAltHpel[8] = {0, 3, 9, 20, 20, 9, 3, 0};   // InterpolationFilter.cpp:186

// ---- integer search: 25 costs per sub-PU ------------------------------------------------------------------
// One CTA per sub-PU.  S[l]: window rows/columns 1 .. of list l ((dx + 5) x (dy + 5) samples: the (dx + 4) x (dy + 4) bilinear
// block + one more row and column for the second tap), pred[l]: 10-bit bilinear predictions (FILTER_DMVR,
// InterpolationFilter.cpp:416-447, 588-600), then xDMVRCost (:2147-2155) for the 25 offsets of m_pSearchOffset: L0 at +offset,
// L1 at -offset (xBIPMVRefine :2050-2073), SAD over even rows << 1 >> 1.
__global__ void __launch_bounds__(128) k_dmvr_classic(const MmRef* __restrict__ refs, const mmgpu_dmvr_classic_subpu* __restrict__ subpus,
                                                      int bd, unsigned long long* __restrict__ costs)
{
  __shared__ int16_t S[2][21 * 21];
  __shared__ int16_t pred[2][20 * 20];
  const mmgpu_dmvr_classic_subpu sp = subpus[blockIdx.x];
  const int dx = sp.w, dy = sp.h;
  const int sw = dx + 5, sh = dy + 5, pw = dx + 4, ph = dy + 4;
  for (int l = 0; l < 2; ++l)
  {
    const MmRef ref = refs[l ? sp.ref_slot1 : sp.ref_slot0];
    const int ox = (l ? sp.org_x1 : sp.org_x0) + 1, oy = (l ? sp.org_y1 : sp.org_y0) + 1;   // xinitMC: offset 3 * (stride + 1) in a buffer whose window starts at (2,2)
    for (int t = threadIdx.x; t < sw * sh; t += blockDim.x)
    {
      const int y = t / sw, x = t - y * sw;
      S[l][y * 21 + x] = ref.plane00[0][(long)(oy + y) * ref.stride[0] + ox + x];
    }
  }
  __syncthreads();
  const int shift1 = 4 - (10 - bd), off1 = 1 << (shift1 - 1);   // isFirst: IF_FILTER_PREC_BILINEAR - (IF_INTERNAL_PREC_BILINEAR - bd)
  for (int l = 0; l < 2; ++l)
  {
    const int fx = l ? sp.frac_x1 : sp.frac_x0, fy = l ? sp.frac_y1 : sp.frac_y0;
    const int16_t* s = S[l];
    for (int t = threadIdx.x; t < pw * ph; t += blockDim.x)
    {
      const int y = t / pw, x = t - y * pw;
      const int16_t* p = s + y * 21 + x;
      int v;
      if (fy == 0)
      {
        if (fx == 0) v = (bd > 10) ? (int16_t)((p[0] + (1 << (bd - 11))) >> (bd - 10)) : (int16_t)(p[0] << (10 - bd));   // filterCopy<true,false>
        else v = (int16_t)((p[0] * (16 - fx) + p[1] * fx + off1) >> shift1);
      }
      else if (fx == 0) v = (int16_t)((p[0] * (16 - fy) + p[21] * fy + off1) >> shift1);
      else
      {
        const int a = (int16_t)((p[0] * (16 - fx) + p[1] * fx + off1) >> shift1);
        const int b = (int16_t)((p[21] * (16 - fx) + p[22] * fx + off1) >> shift1);
        v = (int16_t)((a * (16 - fy) + b * fy + 8) >> 4);
      }
      pred[l][y * 20 + x] = (int16_t)v;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = dx * (dy >> 1);
  for (int k = warp; k < 25; k += 4)
  {
    const int oh = k % 5 - 2, ov = k / 5 - 2;
    const int16_t* a = pred[0] + (2 + ov) * 20 + 2 + oh;
    const int16_t* b = pred[1] + (2 - ov) * 20 + 2 - oh;
    unsigned int sad = 0;
    for (int t = lane; t < n; t += 32)
    {
      const int r = t / dx, x = t - r * dx;
      sad += (unsigned int)abs(a[2 * r * 20 + x] - b[2 * r * 20 + x]);
    }
    for (int o = 16; o > 0; o >>= 1) sad += __shfl_xor_sync(0xffffffffu, sad, o);
    if (lane == 0) costs[(long)blockIdx.x * 25 + k] = ((unsigned long long)sad << 1) >> 1;
  }
}

// ---- final MC from the padded window -------------------------------------------------------------------------
// One CTA per job = one component of one list of one sub-PU.  B: the padded buffer ((w + taps - 1 + 2 pad) x (h + taps - 1 + 2 pad)
// samples, clamped reads of the window), T: first-pass rows.  Output as xPredInterBlk (:1019-1049): bi = true (rnd 0): 14-bit
// intermediates, rnd 1: rounded and clipped; with bio (luma) into (2,2) of a (w + 4)-stride buffer plus the ring of integer samples (:1054-1093).
#define MM_DMVR_BMAX (16 + 7 + 4)
__global__ void __launch_bounds__(128) k_dmvr_final_mc(const MmRef* __restrict__ refs, MmFir f, const mmgpu_dmvr_mc_job* __restrict__ jobs,
                                                       int rnd, int lo, int hi, int16_t* __restrict__ dst)
{
  __shared__ int16_t B[MM_DMVR_BMAX * MM_DMVR_BMAX];
  __shared__ int16_t T[(16 + 7) * 16];
  const mmgpu_dmvr_mc_job j = jobs[blockIdx.x];
  const MmRef ref = refs[j.ref_slot];
  const bool luma = j.comp == 0;
  const int taps = luma ? 8 : 4, half = taps / 2 - 1;
  const int w = j.w, h = j.h, ww = w + taps - 1, wh = h + taps - 1, pad = j.pad;
  const int bw = ww + 2 * pad, bh = wh + 2 * pad;
  const int16_t* plane = ref.plane00[j.comp];
  const int stride = ref.stride[j.comp];
  for (int t = threadIdx.x; t < bw * bh; t += blockDim.x)
  {
    const int y = t / bw, x = t - y * bw;
    const int cy = min(max(y - pad, 0), wh - 1), cx = min(max(x - pad, 0), ww - 1);
    B[y * MM_DMVR_BMAX + x] = plane[(long)(j.org_y + cy) * stride + j.org_x + cx];
  }
  __syncthreads();
  // block sample (0,0) in B: (pad + half + dint) -- buffer offset (2 + leftPixelExtra) * (stride + 1) + deltaInt (:2131-2133)
  const int16_t* p = B + (pad + half + j.dint_y) * MM_DMVR_BMAX + pad + half + j.dint_x;
  const int xF = j.frac_x, yF = j.frac_y;
  const int8_t* ch = luma ? ((j.alt_hpel && xF == 8) ? c_lumaAltHpel : c_lumaFilter[xF]) : c_chromaFilter[xF];
  const int8_t* cv = luma ? ((j.alt_hpel && yF == 8) ? c_lumaAltHpel : c_lumaFilter[yF]) : c_chromaFilter[yF];
  const bool bio = j.bio && luma;
  int16_t* out = dst + j.dst_offset + (bio ? 2 * (w + 4) + 2 : 0);
  const int os = bio ? (w + 4) : j.dst_stride;
  if (xF == 0 || yF == 0)
  {
    for (int t = threadIdx.x; t < w * h; t += blockDim.x)
    {
      const int y = t / w, x = t - y * w;
      const int16_t* q = p + y * MM_DMVR_BMAX + x;
      int v;
      if (xF == 0 && yF == 0) v = rnd ? (int)q[0] : (int)(int16_t)((int16_t)(q[0] << f.cpShift) - (int16_t)8192);   // filterCopy<true,isLast>
      else
      {
        int sum = 0;
        if (yF == 0)
          for (int k = 0; k < taps; ++k) sum += q[k - half] * ch[k];
        else
          for (int k = 0; k < taps; ++k) sum += q[(k - half) * MM_DMVR_BMAX] * cv[k];
        if (rnd) v = min(max((int)(int16_t)((sum + f.off1l) >> f.shift1l), lo), hi);
        else v = (int)(int16_t)((sum + f.off1) >> f.shift1);
      }
      out[y * os + x] = (int16_t)v;
    }
  }
  else
  {
    for (int t = threadIdx.x; t < w * wh; t += blockDim.x)
    {
      const int y = t / w, x = t - y * w;
      const int16_t* q = p + (y - half) * MM_DMVR_BMAX + x;
      int sum = 0;
      for (int k = 0; k < taps; ++k) sum += q[k - half] * ch[k];
      T[y * 16 + x] = (int16_t)((sum + f.off1) >> f.shift1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < w * h; t += blockDim.x)
    {
      const int y = t / w, x = t - y * w;
      int sum = 0;
      for (int k = 0; k < taps; ++k) sum += T[(y + k) * 16 + x] * cv[k];
      if (rnd) out[y * os + x] = (int16_t)min(max((int)(int16_t)((sum + f.off2) >> f.shift2), lo), hi);
      else out[y * os + x] = (int16_t)(sum >> f.shift2bi);
    }
  }
  if (bio)
  {
    const int xOff = (xF < 8) ? 1 : 0, yOff = (yF < 8) ? 1 : 0;
    int16_t* tmp = dst + j.dst_offset;
    const int n = 2 * (w + 2) + 2 * h;
    for (int t = threadIdx.x; t < n; t += blockDim.x)
    {
      int i, r;
      if (t < w + 2) { i = t; r = 0; }
      else if (t < 2 * (w + 2)) { i = t - (w + 2); r = h + 1; }
      else { const int k = t - 2 * (w + 2); r = 1 + (k >> 1); i = (k & 1) ? (w + 1) : 0; }
      const int v = p[(r - yOff) * MM_DMVR_BMAX + (i - xOff)];
      tmp[(r + 1) * (w + 4) + (i + 1)] = (int16_t)((int16_t)(v << f.cpShift) - (int16_t)8192);
    }
  }
}

void mmk_dmvr_classic(cudaStream_t s, const MmRef* refs, const mmgpu_dmvr_classic_subpu* subpus, int n, int bd, unsigned long long* costs)
{
  k_dmvr_classic<<<n, 128, 0, s>>>(refs, subpus, bd, costs);
}
void mmk_dmvr_final_mc(cudaStream_t s, const MmRef* refs, const MmFir& f, const mmgpu_dmvr_mc_job* jobs, int n, int rnd, int lo, int hi,
                       int16_t* dst)
{
  k_dmvr_final_mc<<<n, 128, 0, s>>>(refs, f, jobs, rnd, lo, hi, dst);
}
