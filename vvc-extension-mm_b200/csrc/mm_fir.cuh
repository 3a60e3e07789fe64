// mm_fir.cuh -- device functions of the integer part: VTM separable interpolation and the distortion pieces.
#pragma once
#include "mm_common.cuh"

// -------------------------------------------------------------------------------------------------
// Generic FIR of one sub-block (any component / tap count / uni or bi): the literal two-pass
// structure of xPredInterBlkMM (InterPrediction.cpp:786-828) with filter<>'s arithmetic.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mm_clip(int v, int lo, int hi) { return min(max(v, lo), hi); }

template <int NTAPS>
__device__ void mm_fir_generic(const MmFir& f, const int16_t* __restrict__ ref, int stride, int xPos, int yPos, int xFrac,
                               int yFrac, int sbw, int sbh, bool rnd, int lo, int hi, int16_t* dst, int dstStride)
{
  const int8_t* ch = (NTAPS == 8) ? c_lumaFilter[xFrac] : c_chromaFilter[xFrac];
  const int8_t* cv = (NTAPS == 8) ? c_lumaFilter[yFrac] : c_chromaFilter[yFrac];
  const int16_t* p = ref + (long)yPos * stride + xPos;
  const int half = NTAPS / 2 - 1;
  if (yFrac == 0 || xFrac == 0)
  {
    for (int r = 0; r < sbh; ++r)
      for (int x = 0; x < sbw; ++x)
      {
        int val;
        if (xFrac == 0 && yFrac == 0)
        {
          // filterCopy<true,isLast> (:392-470)
          const int s = p[r * stride + x];
          val = rnd ? s : (int)(int16_t)((int16_t)(s << f.cpShift) - (int16_t)8192);
        }
        else
        {
          int sum = 0;
          if (yFrac == 0)
            for (int k = 0; k < NTAPS; ++k) sum += p[r * stride + x + k - half] * ch[k];
          else
            for (int k = 0; k < NTAPS; ++k) sum += p[(r + k - half) * stride + x] * cv[k];
          if (rnd) val = mm_clip((int)(int16_t)((sum + f.off1l) >> f.shift1l), lo, hi);
          else val = (int)(int16_t)((sum + f.off1) >> f.shift1);
        }
        dst[r * dstStride + x] = (int16_t)val;
      }
    return;
  }
  int16_t tmp[(4 + 7) * 4];
  for (int r = 0; r < sbh + NTAPS - 1; ++r)
    for (int x = 0; x < sbw; ++x)
    {
      int sum = 0;
      for (int k = 0; k < NTAPS; ++k) sum += p[(r - half) * stride + x + k - half] * ch[k];
      tmp[r * 4 + x] = (int16_t)((sum + f.off1) >> f.shift1);
    }
  for (int r = 0; r < sbh; ++r)
    for (int x = 0; x < sbw; ++x)
    {
      int sum = 0;
      for (int k = 0; k < NTAPS; ++k) sum += tmp[(r + k) * 4 + x] * cv[k];
      int val;
      if (rnd) val = mm_clip((int)(int16_t)((sum + f.off2) >> f.shift2), lo, hi);
      else val = (int)(int16_t)(sum >> f.shift2bi);
      dst[r * dstStride + x] = (int16_t)val;
    }
}

// -------------------------------------------------------------------------------------------------
// Fast luma path: 4x4 sub-block, 8-tap separable FIR on packed 16-bit pairs with dp2a.
// All (xFrac, yFrac) cases, including 0, run the two-pass form: with the phase-0 kernel {0,0,0,64,0,0,0,0}
// it equals the reference's single-pass / copy results exactly for rndRes = true ("FIR identity");
// the copy case skips the clip like filterCopy<true,true>.
// -------------------------------------------------------------------------------------------------
struct MmTaps  // packed coefficient words of one phase, built once per CTA in shared memory
{
  int32_t a0, a1;      // (c0,c1,c2,c3) (c4,c5,c6,c7): output aligned with the sample pair
  int32_t b0, b1, b2;  // (0,c0,c1,c2) (c3,c4,c5,c6) (c7,0,0,0): output starts on the odd sample
};

__device__ __forceinline__ int mm_pack4(int c0, int c1, int c2, int c3)
{
  return (c0 & 0xff) | ((c1 & 0xff) << 8) | ((c2 & 0xff) << 16) | ((c3 & 0xff) << 24);
}
__device__ __forceinline__ void mm_build_taps(MmTaps* t /* [16] in shared memory */)
{
  for (int ph = threadIdx.x; ph < 16; ph += blockDim.x)
  {
    const int8_t* c = c_lumaFilter[ph];
    t[ph].a0 = mm_pack4(c[0], c[1], c[2], c[3]);
    t[ph].a1 = mm_pack4(c[4], c[5], c[6], c[7]);
    t[ph].b0 = mm_pack4(0, c[0], c[1], c[2]);
    t[ph].b1 = mm_pack4(c[3], c[4], c[5], c[6]);
    t[ph].b2 = mm_pack4(c[7], 0, 0, 0);
  }
}
// four outputs of an 8-tap filter over 12 packed samples w0..w5 (output j starts at sample j)
__device__ __forceinline__ void mm_fir4(const int w0, const int w1, const int w2, const int w3, const int w4,
                                        const int w5, const MmTaps& t, int init, int& o0, int& o1, int& o2, int& o3)
{
  o0 = __dp2a_lo(w0, t.a0, __dp2a_hi(w1, t.a0, __dp2a_lo(w2, t.a1, __dp2a_hi(w3, t.a1, init))));
  o1 = __dp2a_lo(w0, t.b0, __dp2a_hi(w1, t.b0, __dp2a_lo(w2, t.b1, __dp2a_hi(w3, t.b1, __dp2a_lo(w4, t.b2, init)))));
  o2 = __dp2a_lo(w1, t.a0, __dp2a_hi(w2, t.a0, __dp2a_lo(w3, t.a1, __dp2a_hi(w4, t.a1, init))));
  o3 = __dp2a_lo(w1, t.b0, __dp2a_hi(w2, t.b0, __dp2a_lo(w3, t.b1, __dp2a_hi(w4, t.b1, __dp2a_lo(w5, t.b2, init)))));
}

// pred[r*4+c]; BI = false: clipped prediction, BI = true: 14-bit intermediate
// Second (vertical) pass over the packed first-pass rows P[pair][column]; pred[r*4+c].
template <bool BI>
__device__ __forceinline__ void mm_fir_vpass(const MmFir& f, const int (*P)[4], const MmTaps& tv, bool copy, int lo, int hi,
                                             int* pred)
{
  // rndRes = true: the reference stores (sum + offset) >> shift into a Pel (int16) and then clips (except in the pure copy
  // case, filterCopy<true,true>).  The int16 store never wraps here: the first-pass rows are int16, sum |c| <= 112, so
  // |(sum + off2) >> shift2| <= (32768 * 112 + 524800) >> 10 < 2^13 (and the copy case returns the sample itself); the
  // copy case is folded into per-lane clip bounds.
  const int clo = copy ? -32768 : lo, chi = copy ? 32767 : hi;
#pragma unroll
  for (int c = 0; c < 4; ++c)
  {
    int o0, o1, o2, o3;
    mm_fir4(P[0][c], P[1][c], P[2][c], P[3][c], P[4][c], P[5][c], tv, BI ? 0 : f.off2, o0, o1, o2, o3);
    if (BI)
    {
      pred[0 * 4 + c] = (int)(int16_t)(o0 >> f.shift2bi);
      pred[1 * 4 + c] = (int)(int16_t)(o1 >> f.shift2bi);
      pred[2 * 4 + c] = (int)(int16_t)(o2 >> f.shift2bi);
      pred[3 * 4 + c] = (int)(int16_t)(o3 >> f.shift2bi);
    }
    else
    {
      pred[0 * 4 + c] = mm_clip(o0 >> f.shift2, clo, chi);
      pred[1 * 4 + c] = mm_clip(o1 >> f.shift2, clo, chi);
      pred[2 * 4 + c] = mm_clip(o2 >> f.shift2, clo, chi);
      pred[3 * 4 + c] = mm_clip(o3 >> f.shift2, clo, chi);
    }
  }
}

// 4-byte gather variant (k_dmvr): `base` points at the 32-bit word holding sample (xPos-3)&~1 of row yPos-3 (read-only
// path, word stride `wstride`); sh = 16 if xPos-3 is odd (funnel-shift realignment).
template <bool BI>
__device__ __forceinline__ void mm_fir_luma4x4_core(const MmFir& f, const MmTaps* taps, const uint32_t* base, int wstride,
                                                    int sh, int xFrac, int yFrac, int lo, int hi, int* pred)
{
  const MmTaps th = taps[xFrac];
  const MmTaps tv = taps[yFrac];
  int P[6][4];  // vertical pairs of first-pass rows, per column
  int hprev[4];
#pragma unroll
  for (int r = 0; r < 11; ++r)
  {
    const uint32_t* rowp = base + (long)r * wstride;
    uint32_t w0 = __ldg(rowp + 0), w1 = __ldg(rowp + 1), w2 = __ldg(rowp + 2), w3 = __ldg(rowp + 3), w4 = __ldg(rowp + 4),
             w5 = __ldg(rowp + 5);
    w0 = __funnelshift_r(w0, w1, sh);
    w1 = __funnelshift_r(w1, w2, sh);
    w2 = __funnelshift_r(w2, w3, sh);
    w3 = __funnelshift_r(w3, w4, sh);
    w4 = __funnelshift_r(w4, w5, sh);
    w5 = __funnelshift_r(w5, w5, sh);
    int h0, h1, h2, h3;
    mm_fir4((int)w0, (int)w1, (int)w2, (int)w3, (int)w4, (int)w5, th, f.off1, h0, h1, h2, h3);
    h0 >>= f.shift1; h1 >>= f.shift1; h2 >>= f.shift1; h3 >>= f.shift1;
    if (r & 1)
    {
      P[r >> 1][0] = __byte_perm(hprev[0], h0, 0x5410);
      P[r >> 1][1] = __byte_perm(hprev[1], h1, 0x5410);
      P[r >> 1][2] = __byte_perm(hprev[2], h2, 0x5410);
      P[r >> 1][3] = __byte_perm(hprev[3], h3, 0x5410);
    }
    else if (r == 10)
    {
      P[5][0] = h0 & 0xffff; P[5][1] = h1 & 0xffff; P[5][2] = h2 & 0xffff; P[5][3] = h3 & 0xffff;
    }
    else
    {
      hprev[0] = h0; hprev[1] = h1; hprev[2] = h2; hprev[3] = h3;
    }
  }
  mm_fir_vpass<BI>(f, P, tv, (xFrac | yFrac) == 0, lo, hi, pred);
}

template <bool BI>
__device__ __forceinline__ void mm_fir_luma4x4(const MmFir& f, const MmTaps* taps, const int16_t* __restrict__ ref,
                                               int stride, int xPos, int yPos, int xFrac, int yFrac, int lo, int hi,
                                               int* pred)
{
  const int a = xPos - 3;
  const uint32_t* base = reinterpret_cast<const uint32_t*>(ref + (long)(yPos - 3) * stride + (a & ~1));
  mm_fir_luma4x4_core<BI>(f, taps, base, stride >> 1, (a & 1) << 4, xFrac, yFrac, lo, hi, pred);
}

// -------------------------------------------------------------------------------------------------
// Wide-load variant of the luma path (k_fir).  A row is read as four aligned 8-byte words (16 samples starting at
// (xPos-3)&~3): half the load requests of the 4-byte gather and fewer L1 wavefronts.  The horizontal alignment
// o = (xPos-3)&3 is folded into per-lane coefficient words, so no realignment shifts are needed:
//   h_j = sum_{k=0..5} dp2a(v[(j>>1)+k], C[o][xFrac][j][k]),  C = the 8 taps placed at sample offset o+j, zeros elsewhere.
// Table (built on the host, mm_build_htaps): t[((o*16 + phase)*4 + j)*3 + m] packs pairs k=2m (low half) and 2m+1 (high half).
// Entry o = 0 also holds the vertical pass' words: j=0 -> (a0,a1,0), j=1 -> (b0,b1,b2) of MmTaps.
// -------------------------------------------------------------------------------------------------
#include "mm_htaps.h"
__device__ __forceinline__ uint32_t mm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int mm_fir6(int v0, int v1, int v2, int v3, int v4, int v5, uint32_t c0, uint32_t c1, uint32_t c2,
                                       int init)
{
  return __dp2a_lo(v0, (int)c0, __dp2a_hi(v1, (int)c0, __dp2a_lo(v2, (int)c1, __dp2a_hi(v3, (int)c1,
         __dp2a_lo(v4, (int)c2, __dp2a_hi(v5, (int)c2, init))))));
}

// cx, cy: the base words (C0, C1) of the horizontal / vertical phase (compact 128-byte table behind the alignment-folded one:
// one 8-byte load per lane each); the six alignment-folded words of the lane are funnel shifts of cx (mm_htaps.h:
// mm_htaps_shift; even outputs sh = o, odd outputs sh = o + 1).  `base` = 8-byte word holding sample (xPos-3)&~3 of row
// yPos-3, `qstride` = row stride in 8-byte words.  SMEM = false: global memory (read-only path); SMEM = true: a TMA-filled
// box in shared memory.  Samples 12,13 of the aligned row are touched only for alignments 2 and 3: the fourth load of a row
// is predicated on that.
template <bool BI, bool SMEM>
__device__ __forceinline__ void mm_fir_luma4x4_wide_core(const MmFir& f, uint2 cx, uint2 cy, const uint2* base, int qstride,
                                                         int a, int xFrac, int yFrac, int lo, int hi, int* pred)
{
  const unsigned she = 8u * (unsigned)(a & 3), sho = she + 8u;
  const uint32_t e0 = __funnelshift_lc(0u, cx.x, she), e1 = __funnelshift_lc(cx.x, cx.y, she), e2 = __funnelshift_lc(cx.y, 0u, she);
  const uint32_t o0 = __funnelshift_lc(0u, cx.x, sho), o1 = __funnelshift_lc(cx.x, cx.y, sho), o2 = __funnelshift_lc(cx.y, 0u, sho);
  const int needTail = (a & 3) >= 2;
  int P[6][4];  // vertical pairs of first-pass rows, per column
  int hprev[4];
#pragma unroll
  for (int r = 0; r < 11; ++r)
  {
    const uint2* rowp = SMEM ? (base + r * qstride) : (base + (long)r * qstride);
    uint2 q0, q1, q2;
    uint32_t q3 = 0;
#ifdef MM_ABL_NOLOAD   // ablation build only (tools/): no reference loads
    q0 = make_uint2(r + a, yFrac); q1 = make_uint2(a, r ^ xFrac); q2 = q0; q3 = q1.x;
    (void)rowp;
#else
    if (SMEM)
    {
      q0 = rowp[0]; q1 = rowp[1]; q2 = rowp[2];
      asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %2, 0;\n@p ld.shared.b32 %0, [%1];\n}" : "+r"(q3) : "r"(mm_smem_u32(rowp + 3)), "r"(needTail));
    }
    else
    {
      q0 = __ldg(rowp + 0); q1 = __ldg(rowp + 1); q2 = __ldg(rowp + 2);
      asm volatile("{\n.reg .pred p;\nsetp.ne.s32 p, %2, 0;\n@p ld.global.nc.b32 %0, [%1];\n}" : "+r"(q3) : "l"(rowp + 3), "r"(needTail));
    }
#endif
    int h0 = mm_fir6((int)q0.x, (int)q0.y, (int)q1.x, (int)q1.y, (int)q2.x, (int)q2.y, e0, e1, e2, f.off1);
    int h1 = mm_fir6((int)q0.x, (int)q0.y, (int)q1.x, (int)q1.y, (int)q2.x, (int)q2.y, o0, o1, o2, f.off1);
    int h2 = mm_fir6((int)q0.y, (int)q1.x, (int)q1.y, (int)q2.x, (int)q2.y, (int)q3, e0, e1, e2, f.off1);
    int h3 = mm_fir6((int)q0.y, (int)q1.x, (int)q1.y, (int)q2.x, (int)q2.y, (int)q3, o0, o1, o2, f.off1);
    h0 >>= f.shift1; h1 >>= f.shift1; h2 >>= f.shift1; h3 >>= f.shift1;
    if (r & 1)
    {
      P[r >> 1][0] = __byte_perm(hprev[0], h0, 0x5410);
      P[r >> 1][1] = __byte_perm(hprev[1], h1, 0x5410);
      P[r >> 1][2] = __byte_perm(hprev[2], h2, 0x5410);
      P[r >> 1][3] = __byte_perm(hprev[3], h3, 0x5410);
    }
    else if (r == 10)
    {
      P[5][0] = h0 & 0xffff; P[5][1] = h1 & 0xffff; P[5][2] = h2 & 0xffff; P[5][3] = h3 & 0xffff;
    }
    else
    {
      hprev[0] = h0; hprev[1] = h1; hprev[2] = h2; hprev[3] = h3;
    }
  }
  MmTaps tv;   // vertical pass: (a0, a1) = the base words, (b0, b1, b2) = the same moved up by one byte
  tv.a0 = (int)cy.x; tv.a1 = (int)cy.y;
  tv.b0 = (int)__funnelshift_lc(0u, cy.x, 8u); tv.b1 = (int)__funnelshift_lc(cy.x, cy.y, 8u); tv.b2 = (int)__funnelshift_lc(cy.y, 0u, 8u);
  mm_fir_vpass<BI>(f, P, tv, (xFrac | yFrac) == 0, lo, hi, pred);
}

// ref plane 8-byte aligned with a stride that is a multiple of 4 samples
template <bool BI>
__device__ __forceinline__ void mm_fir_luma4x4_wide(const MmFir& f, uint2 cx, uint2 cy, const int16_t* __restrict__ ref,
                                                    int stride, int xPos, int yPos, int xFrac, int yFrac, int lo, int hi,
                                                    int* pred)
{
  const int a = xPos - 3;
  const uint2* base = reinterpret_cast<const uint2*>(ref + (long)(yPos - 3) * stride + (a & ~3));
  mm_fir_luma4x4_wide_core<BI, false>(f, cx, cy, base, stride >> 2, a, xFrac, yFrac, lo, hi, pred);
}

// -------------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (sm_100a) used by k_fir to stage a box of the reference plane in shared memory
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mm_mbar_init(unsigned long long* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mm_smem_u32(bar)), "r"(count));
  // make the initialised barrier visible to the async proxy (TMA); the cluster-scope fence.mbarrier_init would
  // invalidate L1 (CCTL.IVALL) once per warp
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mm_mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mm_mbar_wait(unsigned long long* bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(mm_smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tile load: box of the tensor map at element coordinates (x, y) -> shared memory, completion on `bar`
__device__ __forceinline__ void mm_tma_load_2d(void* dst, const void* tmap, int x, int y, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(mm_smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(mm_smem_u32(bar)) : "memory");
}

// -------------------------------------------------------------------------------------------------
// Distortion pieces
// -------------------------------------------------------------------------------------------------
// 4x4 2-D Hadamard in place (d[r*4+c]) and the cross-lane butterflies of larger tiles, in FP32: every value is an integer
// below 2^24 (|coefficient| <= 128 * 1023, sum of magnitudes of a tile <= 128 * sqrt(128) * 1023), so float add/sub/fma are
// exact, and they issue on the FP32 pipes that the integer FIR leaves idle instead of the integer ALU / IMAD pipes that
// it loads.  Coefficient order/sign is irrelevant for the sum of magnitudes; d[0] ends up as the DC term.
__device__ __forceinline__ void mm_had4x4_f(float* d)
{
#pragma unroll
  for (int r = 0; r < 4; ++r)
  {
    const float a = d[r * 4 + 0] + d[r * 4 + 1], b = d[r * 4 + 0] - d[r * 4 + 1];
    const float c = d[r * 4 + 2] + d[r * 4 + 3], e = d[r * 4 + 2] - d[r * 4 + 3];
    d[r * 4 + 0] = a + c; d[r * 4 + 1] = b + e; d[r * 4 + 2] = a - c; d[r * 4 + 3] = b - e;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
  {
    const float a = d[0 * 4 + c] + d[1 * 4 + c], b = d[0 * 4 + c] - d[1 * 4 + c];
    const float g = d[2 * 4 + c] + d[3 * 4 + c], e = d[2 * 4 + c] - d[3 * 4 + c];
    d[0 * 4 + c] = a + g; d[1 * 4 + c] = b + e; d[2 * 4 + c] = a - g; d[3 * 4 + c] = b - e;
  }
}
// cross-lane butterfly: every lane ends with d[i] (lower lane) + / - d[i] (upper lane): the lower lane keeps the sums, the
// upper lane the differences (sign irrelevant for the sum of magnitudes): one shuffle and one fma per element, no selects.
// d[0] of lane q = 0 stays the DC term.
__device__ __forceinline__ void mm_had_xlane_f(float* d, int bit, float sgn /* +1 lower lane, -1 upper lane */)
{
#pragma unroll
  for (int i = 0; i < 16; ++i)
  {
    const float recv = __shfl_xor_sync(0xffffffffu, d[i], bit);
    d[i] = fmaf(sgn, d[i], recv);
  }
}
