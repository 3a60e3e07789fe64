// pipe_bench.cu -- issue-rate microbenchmark of the integer/FP32 instructions the FIR/Hadamard kernels are built
// from (sm_100a).  Prints warp-instructions per cycle per SM sub-partition for single instructions and pairs, which
// tells which instructions share an execution pipe.  nvcc -arch=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

#define OP_IDP(i)  asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
#define OP_IMAD(i) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
#define OP_IADD(i) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
#define OP_SHF(i)  asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
#define OP_PRMT(i) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
#define OP_LOP(i)  asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
#define OP_MNMX(i) asm volatile("max.s32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(a[(i + 3) & 7]));
#define OP_SHR(i)  asm volatile("shr.s32 %0, %1, 2;" : "=r"(a[i]) : "r"(a[(i + 1) & 7]));
#define OP_SEL(i)  asm volatile("slct.s32.s32 %0, %1, %2, %3;" : "=r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(b), "r"(a[(i + 3) & 7]));
#define OP_ABS(i)  asm volatile("abs.s32 %0, %1;" : "=r"(a[i]) : "r"(a[(i + 1) & 7]));
#define OP_SAD(i)  asm volatile("sad.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
#define OP_FADD(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g));
#define OP_FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g), "f"(h));
#define OP_FMUL(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g));
#define OP_SHFL(i) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(a[i]));
#define OP_I2F(i)  asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(a[(i + 1) & 7])); asm volatile("mov.b32 %0, %1;" : "=r"(a[i]) : "f"(f[i]));
#define OP_F2I(i)  asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(a[i]) : "f"(f[i]));
#define OP_HFMA2(i) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
#define OP_DADD(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(dg));
#define OP_DFMA(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dg), "d"(dg));
#define OP_NONE(i)

#define REP8(OP) OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7)

#define KERNEL(NAME, OPA, OPB, NA, NB)                                                                       \
  __global__ void __launch_bounds__(1024) NAME(int* out, long long* cyc, int b, int c, float g, float h, int iters)     \
  {                                                                                                          \
    int a[8];                                                                                                \
    float f[8];                                                                                              \
    double d[8];                                                                                             \
    double dg = (double)g;                                                                                   \
    bool p = (b & 1) != 0;                                                                                   \
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; f[i] = (float)(threadIdx.x + i); d[i] = f[i]; }  \
    __syncthreads();                                                                                         \
    long long t0 = clock64();                                                                                \
    _Pragma("unroll 1") for (int it = 0; it < iters; ++it)                                                                      \
    {                                                                                                        \
      REP8(OPA) REP8(OPB) REP8(OPA) REP8(OPB)                                                                \
    }                                                                                                        \
    long long t1 = clock64();                                                                                \
    __syncthreads();                                                                                         \
    int s = 0;                                                                                               \
    for (int i = 0; i < 8; ++i) s += a[i] + (int)f[i] + (int)d[i];                                           \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (int)p;                                                 \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                         \
  }                                                                                                          \
  static void run_##NAME(int* out, long long* cyc)                                                           \
  {                                                                                                          \
    for (int threads = 256; threads <= 1024; threads *= 2)                                                   \
    {                                                                                                        \
      NAME<<<148, threads>>>(out, cyc, 3, 5, 1.0001f, 0.5f, ITER);                                                 \
      cudaDeviceSynchronize();                                                                               \
      long long hc[148];                                                                                     \
      cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);                                               \
      double avg = 0;                                                                                        \
      for (int i = 0; i < 148; ++i) avg += hc[i];                                                            \
      avg /= 148;                                                                                            \
      const double warpsPerSmsp = threads / 32 / 4.0;                                                        \
      const double na = 16.0 * (NA) * ITER * warpsPerSmsp, nb = 16.0 * (NB) * ITER * warpsPerSmsp;          \
      printf("%-14s warps/SMSP %4.0f  cycles %10.0f  A/clk %.3f  B/clk %.3f  total/clk %.3f\n", #NAME,       \
             warpsPerSmsp, avg, na / avg, nb / avg, (na + nb) / avg);                                        \
    }                                                                                                        \
  }

KERNEL(k_idp, OP_IDP, OP_NONE, 1, 0)
KERNEL(k_imad, OP_IMAD, OP_NONE, 1, 0)
KERNEL(k_iadd, OP_IADD, OP_NONE, 1, 0)
KERNEL(k_shf, OP_SHF, OP_NONE, 1, 0)
KERNEL(k_prmt, OP_PRMT, OP_NONE, 1, 0)
KERNEL(k_lop, OP_LOP, OP_NONE, 1, 0)
KERNEL(k_mnmx, OP_MNMX, OP_NONE, 1, 0)
KERNEL(k_shr, OP_SHR, OP_NONE, 1, 0)
KERNEL(k_sel, OP_SEL, OP_NONE, 1, 0)
KERNEL(k_abs, OP_ABS, OP_NONE, 1, 0)
KERNEL(k_sad, OP_SAD, OP_NONE, 1, 0)
KERNEL(k_fadd, OP_FADD, OP_NONE, 1, 0)
KERNEL(k_ffma, OP_FFMA, OP_NONE, 1, 0)
KERNEL(k_fmul, OP_FMUL, OP_NONE, 1, 0)
KERNEL(k_shfl, OP_SHFL, OP_NONE, 1, 0)
KERNEL(k_i2f, OP_I2F, OP_NONE, 1, 0)
KERNEL(k_f2i, OP_F2I, OP_NONE, 1, 0)
KERNEL(k_hfma2, OP_HFMA2, OP_NONE, 1, 0)
KERNEL(k_dadd, OP_DADD, OP_NONE, 1, 0)
KERNEL(k_dfma, OP_DFMA, OP_NONE, 1, 0)
KERNEL(k_idp_iadd, OP_IDP, OP_IADD, 1, 1)
KERNEL(k_idp_shf, OP_IDP, OP_SHF, 1, 1)
KERNEL(k_idp_fadd, OP_IDP, OP_FADD, 1, 1)
KERNEL(k_idp_ffma, OP_IDP, OP_FFMA, 1, 1)
KERNEL(k_idp_imad, OP_IDP, OP_IMAD, 1, 1)
KERNEL(k_iadd_fadd, OP_IADD, OP_FADD, 1, 1)
KERNEL(k_iadd_imad, OP_IADD, OP_IMAD, 1, 1)
KERNEL(k_iadd_shf, OP_IADD, OP_SHF, 1, 1)
KERNEL(k_fadd_ffma, OP_FADD, OP_FFMA, 1, 1)
KERNEL(k_imad_fadd, OP_IMAD, OP_FADD, 1, 1)
KERNEL(k_shfl_iadd, OP_SHFL, OP_IADD, 1, 1)
KERNEL(k_shfl_idp, OP_SHFL, OP_IDP, 1, 1)

int main()
{
  int* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(int));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run_k_idp(out, cyc); run_k_imad(out, cyc); run_k_iadd(out, cyc); run_k_shf(out, cyc); run_k_prmt(out, cyc);
  run_k_lop(out, cyc); run_k_mnmx(out, cyc); run_k_shr(out, cyc); run_k_sel(out, cyc); run_k_abs(out, cyc);
  run_k_sad(out, cyc); run_k_fadd(out, cyc); run_k_ffma(out, cyc); run_k_fmul(out, cyc); run_k_shfl(out, cyc);
  run_k_i2f(out, cyc); run_k_f2i(out, cyc); run_k_hfma2(out, cyc); run_k_dadd(out, cyc); run_k_dfma(out, cyc);
  run_k_idp_iadd(out, cyc); run_k_idp_shf(out, cyc); run_k_idp_fadd(out, cyc); run_k_idp_ffma(out, cyc);
  run_k_idp_imad(out, cyc); run_k_iadd_fadd(out, cyc); run_k_iadd_imad(out, cyc); run_k_iadd_shf(out, cyc);
  run_k_fadd_ffma(out, cyc); run_k_imad_fadd(out, cyc); run_k_shfl_iadd(out, cyc); run_k_shfl_idp(out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
