// Integer-pipe micro-benchmarks for the B200 (sm_100a): how fast are
// IMAD / IMAD.WIDE and the 256-bit Montgomery multiplier variants?  These set
// the compute roofline of the share-reconstruction kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../honeybadgermpc_b200/csrc/fp256.cuh"
using namespace hb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <class F, int CH>
__global__ void __launch_bounds__(256) k_mul(const Fe* a, const Fe* b, Fe* o, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fe x[CH], y = b[i];
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = a[i]; x[c].w[0] += c; }
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = mont_mul<F>(x[c], y);
  }
  Fe r = x[0];
#pragma unroll
  for (int c = 1; c < CH; c++) r = fe_add<F>(r, x[c]);
  o[i] = r;
}

template <class F, int CH>
__global__ void __launch_bounds__(256) k_bfly(const Fe* a, const Fe* b, Fe* o, int iters) {
  // NTT-like mix: one mulmod + one add + one sub per step
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fe x[CH], z[CH], w = b[i];
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = a[i]; z[c] = b[i]; x[c].w[0] += c; }
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      Fe tw = mont_mul<F>(z[c], w);
      z[c] = fe_sub<F>(x[c], tw);
      x[c] = fe_add<F>(x[c], tw);
    }
  }
  Fe r = x[0];
#pragma unroll
  for (int c = 0; c < CH; c++) r = fe_add<F>(r, z[c]);
  o[i] = r;
}

__global__ void __launch_bounds__(256) k_imad_wide(uint32_t* o, int iters, uint32_t m0) {
  uint64_t acc[8];
  uint32_t a = threadIdx.x * 2654435761u + 1, b = m0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc[c] = c + threadIdx.x;
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 8; c++)
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
  }
  uint64_t r = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) r ^= acc[c];
  o[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(r ^ (r >> 32));
}

__global__ void __launch_bounds__(256) k_imad(uint32_t* o, int iters, uint32_t m0) {
  uint32_t acc[8];
  uint32_t a = threadIdx.x * 2654435761u + 1, b = m0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc[c] = c + threadIdx.x;
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 8; c++)
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) r ^= acc[c];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void __launch_bounds__(256) k_iadd3(uint32_t* o, int iters, uint32_t m0) {
  uint32_t acc[8];
  uint32_t a = threadIdx.x * 2654435761u + 1, b = m0;
#pragma unroll
  for (int c = 0; c < 8; c++) acc[c] = c + threadIdx.x;
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 8; c++)
      asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(acc[c]) : "r"(a), "r"(b));
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) r ^= acc[c];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
float time_ms(K launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
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
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d MHz\n", prop.name, sms, prop.clockRate / 1000);
  // BLS12-381 r parameters for the __constant__ bank
  FieldParams fp = {};
  const uint32_t P[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
  for (int i = 0; i < 8; i++) fp.p[i] = P[i];
  fp.n0inv = 0xffffffffu;
  CK(cudaMemcpyToSymbol(c_field, &fp, sizeof(fp)));

  const int threads = 256;
  for (int bps : {1, 2, 4}) {
    int blocks = sms * bps;
    size_t n = (size_t)blocks * threads;
    std::vector<Fe> ha(n), hb_(n);
    for (size_t i = 0; i < n; i++) for (int j = 0; j < 8; j++) { ha[i].w[j] = (uint32_t)(i * 2654435761u + j * 40503u) & (j == 7 ? 0x3fffffffu : 0xffffffffu); hb_[i].w[j] = (uint32_t)(i * 40503u + j * 2654435761u + 7) & (j == 7 ? 0x3fffffffu : 0xffffffffu); }
    Fe *da, *db, *dout; uint32_t* d32;
    CK(cudaMalloc(&da, n * 32)); CK(cudaMalloc(&db, n * 32)); CK(cudaMalloc(&dout, n * 32)); CK(cudaMalloc(&d32, n * 4));
    CK(cudaMemcpy(da, ha.data(), n * 32, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb_.data(), n * 32, cudaMemcpyHostToDevice));
    int iters = 4000;
    printf("--- %d blocks/SM x %d threads (warps/SM = %d)\n", bps, threads, bps * threads / 32);
    {
      float ms = time_ms([&] { k_imad_wide<<<blocks, threads>>>(d32, iters, 12345u); });
      printf("IMAD.WIDE.U32 : %8.1f Gop/s  (%.2f /clk/SM @%d MHz)\n", n * 8.0 * iters / ms / 1e6, n * 8.0 * iters / ms / 1e3 / sms / (prop.clockRate / 1000.0) , prop.clockRate / 1000);
      ms = time_ms([&] { k_imad<<<blocks, threads>>>(d32, iters, 12345u); });
      printf("IMAD (32)     : %8.1f Gop/s  (%.2f /clk/SM)\n", n * 8.0 * iters / ms / 1e6, n * 8.0 * iters / ms / 1e3 / sms / (prop.clockRate / 1000.0));
      ms = time_ms([&] { k_iadd3<<<blocks, threads>>>(d32, iters, 12345u); });
      printf("IADD x2       : %8.1f Gop/s  (%.2f /clk/SM)\n", n * 16.0 * iters / ms / 1e6, n * 16.0 * iters / ms / 1e3 / sms / (prop.clockRate / 1000.0));
    }
#define RUN(NAME, KERN, CH) { float ms = time_ms([&] { KERN<<<blocks, threads>>>(da, db, dout, iters / CH); }); \
      printf("%-28s ch=%d: %8.2f G mulmod/s  (%.1f ms)\n", NAME, CH, n * (double)(iters / CH) * CH / ms / 1e6, ms); }
    RUN("mul FieldAny(const bank)", (k_mul<FieldAny, 1>), 1);
    RUN("mul FieldAny(const bank)", (k_mul<FieldAny, 2>), 2);
    RUN("mul FieldAny(const bank)", (k_mul<FieldAny, 4>), 4);
    RUN("mul FieldBLS(imm,lowones)", (k_mul<FieldBLS, 1>), 1);
    RUN("mul FieldBLS(imm,lowones)", (k_mul<FieldBLS, 2>), 2);
    RUN("mul FieldBLS(imm,lowones)", (k_mul<FieldBLS, 4>), 4);
    RUN("mul FieldBLSConst(lowones)", (k_mul<FieldBLSConst, 1>), 1);
    RUN("mul FieldBLSConst(lowones)", (k_mul<FieldBLSConst, 2>), 2);
    RUN("mul FieldBLSConst(lowones)", (k_mul<FieldBLSConst, 4>), 4);
    RUN("bfly FieldAny", (k_bfly<FieldAny, 2>), 2);
    RUN("bfly FieldBLS", (k_bfly<FieldBLS, 2>), 2);
    RUN("bfly FieldBLSConst", (k_bfly<FieldBLSConst, 2>), 2);
    RUN("bfly FieldBLSConst", (k_bfly<FieldBLSConst, 4>), 4);
    // correctness cross-check between variants (same inputs, same iteration count)
    std::vector<Fe> r1(n), r2(n), r3(n);
    k_mul<FieldAny, 1><<<blocks, threads>>>(da, db, dout, 16); CK(cudaMemcpy(r1.data(), dout, n * 32, cudaMemcpyDeviceToHost));
    k_mul<FieldBLS, 1><<<blocks, threads>>>(da, db, dout, 16); CK(cudaMemcpy(r2.data(), dout, n * 32, cudaMemcpyDeviceToHost));
    k_mul<FieldBLSConst, 1><<<blocks, threads>>>(da, db, dout, 16); CK(cudaMemcpy(r3.data(), dout, n * 32, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n; i++) for (int j = 0; j < 8; j++) if (r1[i].w[j] != r2[i].w[j] || r1[i].w[j] != r3[i].w[j]) { bad++; break; }
    printf("variant agreement: %zu mismatches of %zu\n", bad, n);
    CK(cudaFree(da)); CK(cudaFree(db)); CK(cudaFree(dout)); CK(cudaFree(d32));
  }
  return 0;
}
