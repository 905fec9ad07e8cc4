// Micro-benchmark 3: what exactly limits the 64-bit integer multiply-add pipe on
// the B200, and how fast is the FP64 pipe next to it?  (Feeds the choice of
// multiplier for the next round: DESIGN.md section 4.)
//   * IMAD.WIDE without carries (register / constant-bank multiplier) and as
//     predicate carry chains of 2 and 4
//   * DFMA alone, and DFMA with the two 64-bit integer adds per product that a
//     52-bit-limb FP64 multiplier needs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench3 microbench3.cu
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct Consts { uint32_t c[8]; };

// Plain (carry-free) IMAD.WIDE accumulation.  One multiplier operand is the high word of a
// NEIGHBOURING accumulator, so the products change every iteration: ptxas hoists loop-
// invariant products out of the loop and turns the multiply-adds into IADD3s -- which is
// what it did to the first version of this test (and to microbench.cu's k_imad_wide): the
// "60 /clk/SM" those reported was the integer ADD rate.
// mode 0: second operand in a register   1: second operand from the constant bank
template <int MODE>
__global__ void __launch_bounds__(256) k_wide(uint32_t* o, int iters, uint32_t m0, const __grid_constant__ Consts cc) {
  uint64_t acc[8];
  uint32_t b[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    acc[c] = ((uint64_t)(threadIdx.x * 2654435761u + 1 + c * 40503u) << 32) | (c + threadIdx.x);
    b[c] = m0 + c * 7919u;
  }
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const uint32_t a = (uint32_t)(acc[(c + 1) & 7] >> 32);  // changes every iteration
      if (MODE == 1)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(cc.c[c]));
      else
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b[c]));
    }
  }
  uint64_t r = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) r ^= acc[c];
  o[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(r ^ (r >> 32));
}

// carry chains over distinct registers: LEN products per chain (LEN = 2 or 4), 16 products per step
// CONSTB: the common multiplier comes from the constant bank
template <int LEN, bool CONSTB>
__global__ void __launch_bounds__(256) k_chain(uint32_t* o, int iters, uint32_t m0, const __grid_constant__ Consts cc) {
  uint32_t acc[4][8], a[4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    a[c] = threadIdx.x * 2654435761u + 1 + c * 40503u;
#pragma unroll
    for (int j = 0; j < 8; j++) acc[c][j] = c + j + threadIdx.x;
  }
  uint32_t top = 0;
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const uint32_t b = CONSTB ? cc.c[c] : (m0 + c);
      uint32_t* x = acc[c];
      if (LEN == 4) {
        asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(top)
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b));
      } else {
        asm volatile("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(top)
            : "r"(a[0]), "r"(a[1]), "r"(b));
        asm volatile("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]), "+r"(top)
            : "r"(a[2]), "r"(a[3]), "r"(b));
      }
    }
  }
  uint32_t r = top;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= acc[c][j];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// FP64: 8 independent chains of fma(a[c], b[c], x[c]); IADDS = 64-bit integer adds issued per DFMA
template <int IADDS>
__global__ void __launch_bounds__(256) k_dfma(uint32_t* o, int iters, double m0) {
  double x[8], a[8], b[8];
  unsigned long long s[8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    x[c] = 1.0 + c + threadIdx.x;
    a[c] = 1.0 + 1e-9 * (c + threadIdx.x);
    b[c] = m0 * (1 + c);
    s[c] = c;
  }
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 8; c++) {
      asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(x[c]) : "d"(a[c]), "d"(b[c]));
      if (IADDS >= 1) asm volatile("add.u64 %0, %0, %1;" : "+l"(s[c]) : "l"(__double_as_longlong(a[c])));
      if (IADDS >= 2) asm volatile("add.u64 %0, %0, %1;" : "+l"(s[(c + 3) & 7]) : "l"(__double_as_longlong(b[c])));
    }
  }
  double r = 0;
  unsigned long long q = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) { r += x[c]; q ^= s[c]; }
  o[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double_as_longlong(r) ^ (uint32_t)q ^ (uint32_t)(q >> 32);
}

// IMAD.WIDE chains (the present multiplier) and DFMA issued by the same warp: do the two pipes overlap?
__global__ void __launch_bounds__(256) k_mix(uint32_t* o, int iters, uint32_t m0, double d0) {
  uint32_t acc[2][8], a[4];
  double x[8], da[8];
#pragma unroll
  for (int c = 0; c < 4; c++) a[c] = threadIdx.x * 2654435761u + 1 + c * 40503u;
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[c][j] = c + j + threadIdx.x;
#pragma unroll
  for (int c = 0; c < 8; c++) { x[c] = 1.0 + c + threadIdx.x; da[c] = 1.0 + 1e-9 * (c + threadIdx.x); }
  uint32_t top = 0;
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      uint32_t* y = acc[c];
      const uint32_t b = m0 + c;
      asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
          "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
          "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
          "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
          "addc.u32 %8, %8, 0;"
          : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7]), "+r"(top)
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b));
#pragma unroll
      for (int q = 0; q < 4; q++)
        asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(x[4 * c + q]) : "d"(da[4 * c + q]), "d"(d0));
    }
  }
  uint32_t r = top;
  double rr = 0;
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= acc[c][j];
#pragma unroll
  for (int c = 0; c < 8; c++) rr += x[c];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r ^ (uint32_t)__double_as_longlong(rr);
}

template <class K>
float time_ms(K launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double clk = clk_khz * 1e3;
  printf("device %s, %d SMs, clock %.0f MHz\n", prop.name, sms, clk / 1e6);
  uint32_t* o; CK(cudaMalloc(&o, (size_t)sms * 8 * 256 * 4));
  Consts cc; for (int i = 0; i < 8; i++) cc.c[i] = 0x9e3779b9u * (i + 1);
  const int iters = 4096;
  for (int bps : {1, 2, 4}) {
    const int blocks = sms * bps;
    printf("--- %d blocks/SM x 256 threads (warps/SM = %d)\n", bps, bps * 8);
    auto rep = [&](const char* name, double ops_per_thread_iter, float ms) {
      double ops = ops_per_thread_iter * iters * (double)blocks * 256;
      printf("%-44s %9.1f Gop/s  (%6.2f /clk/SM)\n", name, ops / ms / 1e6, ops / (ms * 1e-3) / clk / sms);
    };
    rep("IMAD.WIDE + 64-bit ALU add, no carry chain (reg b)", 8, time_ms([&] { k_wide<0><<<blocks, 256>>>(o, iters, 12345u, cc); }));
    rep("IMAD.WIDE + 64-bit ALU add, no carry chain (ur b)", 8, time_ms([&] { k_wide<1><<<blocks, 256>>>(o, iters, 12345u, cc); }));
    rep("carry chains of 4 (reg b)   [imad.wide]", 16, time_ms([&] { k_chain<4, false><<<blocks, 256>>>(o, iters, 12345u, cc); }));
    rep("carry chains of 4 (const b) [imad.wide]", 16, time_ms([&] { k_chain<4, true><<<blocks, 256>>>(o, iters, 12345u, cc); }));
    rep("carry chains of 2 (reg b)   [imad.wide]", 16, time_ms([&] { k_chain<2, false><<<blocks, 256>>>(o, iters, 12345u, cc); }));
    rep("carry chains of 2 (const b) [imad.wide]", 16, time_ms([&] { k_chain<2, true><<<blocks, 256>>>(o, iters, 12345u, cc); }));
    rep("DFMA.RZ                     [dfma]", 8, time_ms([&] { k_dfma<0><<<blocks, 256>>>(o, iters, 1.0000001); }));
    rep("DFMA.RZ + 1 add.u64 each    [dfma]", 8, time_ms([&] { k_dfma<1><<<blocks, 256>>>(o, iters, 1.0000001); }));
    rep("DFMA.RZ + 2 add.u64 each    [dfma]", 8, time_ms([&] { k_dfma<2><<<blocks, 256>>>(o, iters, 1.0000001); }));
    rep("8 IMAD.WIDE.X + 8 DFMA mixed [imad.wide]", 8, time_ms([&] { k_mix<<<blocks, 256>>>(o, iters, 12345u, 1.0000001); }));
  }
  CK(cudaFree(o));
  return 0;
}
