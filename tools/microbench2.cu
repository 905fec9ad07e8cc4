// Micro-benchmark 2: is the predicate-carry IMAD.WIDE.X half rate, and how fast
// is a carry-free radix-2^29 Montgomery multiplier (9 limbs, R = 2^261)?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../honeybadgermpc_b200/csrc/fp256.cuh"
using namespace hb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_chain(uint32_t* o, int iters, uint32_t m0) {
  // 4 independent accumulators of 8 words; each step = cmad4 (4 IMAD.WIDE, 3 with .X)
  uint32_t acc[4][8]; uint32_t a = threadIdx.x * 2654435761u + 1, b = m0;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[c][j] = c + j + threadIdx.x;
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < 4; c++) cmad4(acc[c], a, b, a, b, a ^ b);
  }
  uint32_t r = 0;
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= acc[c][j];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

constexpr uint32_t MASK29 = (1u << 29) - 1;
struct Fe29 { uint32_t l[9]; };
struct P29 { uint32_t p[9]; uint32_t n0; };
__constant__ P29 c_p29;

template <bool LOWONES>
__device__ __forceinline__ Fe29 mul29(const Fe29& a, const Fe29& b) {
  uint64_t c[9];
#pragma unroll
  for (int i = 0; i < 9; i++) {
    if (i == 0) {
#pragma unroll
      for (int j = 0; j < 9; j++) c[j] = (uint64_t)a.l[j] * b.l[0];
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) c[j] += (uint64_t)a.l[j] * b.l[i];
      c[8] = (uint64_t)a.l[8] * b.l[i];
    }
    uint64_t carry;
    if (LOWONES) {
      // p = 1 mod 2^29, p1 = 2^29 - 8
      uint32_t m = (0u - (uint32_t)c[0]) & MASK29;
      carry = (c[0] + m) >> 29;
      c[1] += ((uint64_t)m << 29) - ((uint64_t)m << 3);
#pragma unroll
      for (int j = 2; j < 9; j++) c[j] += (uint64_t)m * c_p29.p[j];
    } else {
      uint32_t m = ((uint32_t)c[0] * c_p29.n0) & MASK29;
#pragma unroll
      for (int j = 0; j < 9; j++) c[j] += (uint64_t)m * c_p29.p[j];
      carry = c[0] >> 29;
    }
    c[0] = c[1] + carry;
#pragma unroll
    for (int j = 1; j < 8; j++) c[j] = c[j + 1];
  }
  Fe29 r;
#pragma unroll
  for (int j = 0; j < 7; j++) { c[j + 1] += c[j] >> 29; r.l[j] = (uint32_t)c[j] & MASK29; }
  r.l[7] = (uint32_t)c[7] & MASK29;
  r.l[8] = (uint32_t)(c[7] >> 29);
  return r;
}

template <bool LOWONES, int CH>
__global__ void __launch_bounds__(256) k_mul29(const Fe29* a, const Fe29* b, Fe29* o, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fe29 x[CH], y = b[i];
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = a[i]; x[c].l[0] ^= c; }
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < CH; c++) x[c] = mul29<LOWONES>(x[c], y);
  }
  Fe29 r = x[0];
#pragma unroll
  for (int c = 1; c < CH; c++)
#pragma unroll
    for (int j = 0; j < 9; j++) r.l[j] ^= x[c].l[j];
  o[i] = r;
}

// NTT-like step with lazy limb-wise add / sub (sub adds a redundant-form multiple of p)
template <bool LOWONES, int CH>
__global__ void __launch_bounds__(256) k_bfly29(const Fe29* a, const Fe29* b, Fe29* o, int iters, Fe29 D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fe29 x[CH], z[CH], w = b[i];
#pragma unroll
  for (int c = 0; c < CH; c++) { x[c] = a[i]; z[c] = b[i]; x[c].l[0] ^= c; }
  for (int t = 0; t < iters; t++) {
#pragma unroll
    for (int c = 0; c < CH; c++) {
      Fe29 tw = mul29<LOWONES>(z[c], w);
      Fe29 s, d;
#pragma unroll
      for (int j = 0; j < 9; j++) { s.l[j] = x[c].l[j] + tw.l[j]; d.l[j] = x[c].l[j] + D.l[j] - tw.l[j]; }
      // carry pass (limbs back under 2^29, value unchanged)
#pragma unroll
      for (int j = 0; j < 8; j++) { s.l[j + 1] += s.l[j] >> 29; s.l[j] &= MASK29; d.l[j + 1] += d.l[j] >> 29; d.l[j] &= MASK29; }
      s.l[8] &= MASK29; d.l[8] &= MASK29;  // (microbench only: keep values bounded)
      x[c] = s; z[c] = d;
    }
  }
  Fe29 r = x[0];
#pragma unroll
  for (int c = 0; c < CH; c++)
#pragma unroll
    for (int j = 0; j < 9; j++) r.l[j] ^= z[c].l[j];
  o[i] = r;
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

typedef unsigned __int128 u128;
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  // p in radix 2^29
  const uint64_t P64[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
  P29 hp = {};
  for (int j = 0; j < 9; j++) {
    int bit = 29 * j; uint64_t v = 0;
    for (int k = 0; k < 29 && bit + k < 256; k++) v |= ((P64[(bit + k) / 64] >> ((bit + k) % 64)) & 1ull) << k;
    hp.p[j] = (uint32_t)v;
  }
  uint32_t inv = 1; for (int i = 0; i < 5; i++) inv *= 2u - hp.p[0] * inv;
  hp.n0 = (0u - inv) & MASK29;
  printf("p29 = "); for (int j = 0; j < 9; j++) printf("%08x ", hp.p[j]); printf(" n0=%08x\n", hp.n0);
  CK(cudaMemcpyToSymbol(c_p29, &hp, sizeof(hp)));
  Fe29 D;  // 4p in redundant form, every limb >= 2^29 (except top)
  {
    uint64_t v[9]; u128 carry = 0;
    for (int j = 0; j < 9; j++) { u128 t = (u128)hp.p[j] * 4 + carry; v[j] = (uint64_t)(t & MASK29); carry = t >> 29; }
    D.l[0] = (uint32_t)v[0] + (1u << 29);
    for (int j = 1; j < 8; j++) D.l[j] = (uint32_t)v[j] + MASK29;
    D.l[8] = (uint32_t)v[8] - 1;
  }
  const int threads = 256;
  for (int bps : {1, 2, 4}) {
    int blocks = sms * bps; size_t n = (size_t)blocks * threads;
    std::vector<Fe29> ha(n), hb_(n);
    for (size_t i = 0; i < n; i++) for (int j = 0; j < 9; j++) { ha[i].l[j] = (uint32_t)(i * 2654435761u + j * 40503u) & (j == 8 ? 0x3fffffu : MASK29); hb_[i].l[j] = (uint32_t)(i * 40503u + j * 2654435761u + 7) & (j == 8 ? 0x3fffffu : MASK29); }
    Fe29 *da, *db, *dout; uint32_t* d32;
    CK(cudaMalloc(&da, n * 36)); CK(cudaMalloc(&db, n * 36)); CK(cudaMalloc(&dout, n * 36)); CK(cudaMalloc(&d32, n * 4));
    CK(cudaMemcpy(da, ha.data(), n * 36, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb_.data(), n * 36, cudaMemcpyHostToDevice));
    int iters = 4000;
    printf("--- %d blocks/SM x %d threads (warps/SM = %d)\n", bps, threads, bps * threads / 32);
    { float ms = time_ms([&] { k_chain<<<blocks, threads>>>(d32, iters, 12345u); });
      printf("cmad4 chains (IMAD.WIDE.X): %8.1f G imad.wide/s (%.2f /clk/SM @1965)\n", n * 16.0 * iters / ms / 1e6, n * 16.0 * iters / ms / 1e3 / sms / 1965.0); }
#define RUN(NAME, KERN, CH, ...) { float ms = time_ms([&] { KERN<<<blocks, threads>>>(da, db, dout, iters / CH, ##__VA_ARGS__); }); \
      printf("%-28s ch=%d: %8.2f G mulmod/s  (%.1f ms)\n", NAME, CH, n * (double)(iters / CH) * CH / ms / 1e6, ms); }
    RUN("mul29 generic", (k_mul29<false, 1>), 1);
    RUN("mul29 generic", (k_mul29<false, 2>), 2);
    RUN("mul29 lowones", (k_mul29<true, 1>), 1);
    RUN("mul29 lowones", (k_mul29<true, 2>), 2);
    RUN("bfly29 generic", (k_bfly29<false, 1>), 1, D);
    RUN("bfly29 generic", (k_bfly29<false, 2>), 2, D);
    RUN("bfly29 lowones", (k_bfly29<true, 1>), 1, D);
    RUN("bfly29 lowones", (k_bfly29<true, 2>), 2, D);
    // correctness of mul29: compare generic vs lowones and vs host u128 reference on a few lanes
    std::vector<Fe29> r1(n), r2(n);
    k_mul29<false, 1><<<blocks, threads>>>(da, db, dout, 3); CK(cudaMemcpy(r1.data(), dout, n * 36, cudaMemcpyDeviceToHost));
    k_mul29<true, 1><<<blocks, threads>>>(da, db, dout, 3); CK(cudaMemcpy(r2.data(), dout, n * 36, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n; i++) for (int j = 0; j < 9; j++) if (r1[i].l[j] != r2[i].l[j]) { bad++; break; }
    printf("generic vs lowones: %zu mismatches of %zu\n", bad, n);
    CK(cudaFree(da)); CK(cudaFree(db)); CK(cudaFree(dout)); CK(cudaFree(d32));
  }
  return 0;
}
