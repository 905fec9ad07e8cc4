// Stand-alone probe of the tensor-core constant-matrix kernel (csrc/tc_kernels.cuh):
// bit-exactness against a host computation with the field code of host_math.hpp, raw
// accumulator dump against a plain u8 dot product, and device timing.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//        -diag-suppress 20011,20014 -o tools/tc_probe tools/tc_probe.cu
//   tools/tc_probe <d> <n_out> <ob> <batch> [a_pad=16] [hyp=0] [stages=4] [reps=20]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <random>
#include <vector>

#include "../honeybadgermpc_b200/csrc/host_math.hpp"
#include "../honeybadgermpc_b200/csrc/tc_kernels.cuh"

using namespace hb;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

static const uint64_t kBlsP[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull,
                                  0x73eda753299d7d48ull};

int main(int argc, char** argv) {
  const unsigned d = argc > 1 ? atoi(argv[1]) : 6;
  const unsigned n_out = argc > 2 ? atoi(argv[2]) : 6;
  const unsigned ob = argc > 3 ? atoi(argv[3]) : 6;
  const size_t batch = argc > 4 ? atoll(argv[4]) : 65536;
  const unsigned a_pad = argc > 5 ? atoi(argv[5]) : 16;
  const unsigned hyp = argc > 6 ? atoi(argv[6]) : 0;
  const unsigned stages = argc > 7 ? atoi(argv[7]) : 4;
  const int reps = argc > 8 ? atoi(argv[8]) : 20;
  const int ew = argc > 9 ? atoi(argv[9]) : 12;
  const int staged = argc > 10 ? atoi(argv[10]) : (ew == 16);  // staged full-line stores (EW = 16 only)

  FieldParams fp;
  if (!field_params_init(kBlsP, &fp)) return 1;
  HostField f(fp);
  std::mt19937_64 rng(0xB200 + d * 131 + n_out);
  auto rand_fe = [&]() {
    uint64_t l[4] = {rng(), rng(), rng(), rng() >> 2};
    return f.reduce(fe_from_u64(l));
  };

  const unsigned K = 32 * d, NB = 32 * ob, n_blocks = (n_out + ob - 1) / ob;
  // matrix in standard form; B[(i,c)][(j,a)] = byte c of M[i][j] * 256^a mod p
  std::vector<Fe> m((size_t)n_out * d);
  for (auto& v : m) v = rand_fe();
  std::vector<uint8_t> bmat((size_t)n_blocks * NB * K, 0), bplain((size_t)n_blocks * NB * K, 0);
  const Fe c256 = f.from_small(256);
  for (unsigned i = 0; i < n_out; i++)
    for (unsigned j = 0; j < d; j++) {
      Fe cur = f.to_mont(m[(size_t)i * d + j]);
      for (unsigned a = 0; a < 32; a++) {
        Fe s = f.from_mont(cur);
        const uint8_t* bytes = (const uint8_t*)s.w;
        const unsigned nb = i / ob, o = i % ob, kb = j * 32 + a;
        for (unsigned c = 0; c < 32; c++) {
          const unsigned nl = o * 32 + c;
          bmat[(size_t)nb * NB * K + (size_t)(kb / 16) * (NB * 16) + nl * 16 + kb % 16] = bytes[c];
          bplain[((size_t)nb * NB + nl) * K + kb] = bytes[c];
        }
        cur = f.mul(cur, c256);
      }
    }
  // mu = floor(2^280 / p) by long division on bits
  unsigned mu = 0;
  {
    // 2^280 / p: p > 2^254, so the quotient has < 27 bits
    Fe r = fe_zero();
    Fe p;
    memcpy(p.w, fp.p, 32);
    // r = 2^280 mod-steps: shift in bits of 2^280 from the top
    unsigned long long q = 0;
    uint32_t rem[9] = {0};
    for (int bit = 280; bit >= 0; bit--) {
      // rem = rem*2 + (bit == 280)
      uint32_t carry = bit == 280 ? 1 : 0;
      for (int w = 0; w < 9; w++) {
        uint32_t nc = rem[w] >> 31;
        rem[w] = (rem[w] << 1) | carry;
        carry = nc;
      }
      // if rem >= p: rem -= p, q bit = 1
      uint32_t t[9];
      long long b = 0;
      for (int w = 0; w < 9; w++) {
        long long dd = (long long)rem[w] - (long long)(w < 8 ? p.w[w] : 0) + b;
        t[w] = (uint32_t)dd;
        b = dd >> 32;
      }
      q <<= 1;
      if (b == 0) {
        memcpy(rem, t, sizeof t);
        q |= 1;
      }
    }
    mu = (unsigned)q;
    (void)r;
  }
  printf("d=%u n_out=%u ob=%u blocks=%u batch=%zu K=%u NB=%u a_pad=%u hyp=%u stages=%u ew=%d mu=%u\n", d, n_out, ob,
         n_blocks, batch, K, NB, a_pad, hyp, stages, ew, mu);

  // inputs (a few rotating copies so the timed launches read cold data)
  const int copies = 4;
  std::vector<Fe> in((size_t)batch * d);
  for (auto& v : in) v = rand_fe();
  if (batch > 4) {  // edge rows: zeros, p-1, 2^256-1 (non-canonical input), alternating
    Fe pm1;
    memcpy(pm1.w, fp.p, 32);
    pm1.w[0] -= 1;
    for (unsigned j = 0; j < d; j++) {
      in[0 * d + j] = fe_zero();
      in[1 * d + j] = pm1;
      memset(in[2 * d + j].w, 0xff, 32);
      in[3 * d + j] = (j & 1) ? pm1 : fe_zero();
    }
  }
  uint8_t *d_in, *d_b, *d_out;
  uint32_t* d_dbg;
  unsigned* d_err;
  const size_t in_bytes = batch * d * 32, out_bytes = batch * n_out * 32;
  CK(cudaMalloc(&d_in, in_bytes * copies));
  CK(cudaMalloc(&d_out, out_bytes * copies));
  CK(cudaMalloc(&d_b, bmat.size()));
  CK(cudaMalloc(&d_dbg, (size_t)128 * n_blocks * NB * 4));
  CK(cudaMalloc(&d_err, 4));
  CK(cudaMemset(d_err, 0, 4));
  CK(cudaMemset(d_out, 0xAB, out_bytes * copies));
  for (int c = 0; c < copies; c++) CK(cudaMemcpy(d_in + c * in_bytes, in.data(), in_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, bmat.data(), bmat.size(), cudaMemcpyHostToDevice));

  TcArgs a;
  memset(&a, 0, sizeof a);
  a.in = d_in;
  a.bmat = d_b;
  a.out = d_out;
  a.batch = batch;
  a.K = K;
  a.n_out = n_out;
  a.ob = ob;
  a.n_blocks = n_blocks;
  a.in_pitch = K;
  a.out_pitch = n_out * 32;
  a.stages = stages;
  (void)a_pad;
  a.mu = mu;
  a.hyp = hyp;
  a.debug = d_dbg;
  a.error = d_err;
  a.staged = staged && ew == 16;
  const size_t smem = tc_smem_bytes(K, n_blocks, ob, stages) + (a.staged ? kTcStoreStaging : 0);
  printf("smem %zu bytes\n", smem);
  auto launch = [&](const TcArgs& args, unsigned grid_) {
    CUtensorMap tm;
    if (!tc_make_tmap(&tm, args.in, args.batch, args.K, args.in_pitch)) {
      printf("tensor map creation failed\n");
      exit(4);
    }
    if (ew == 8) tc_apply_kernel<FieldBLS, 8, false, true><<<grid_, (8 + kTcLoadWarps + 1) * 32, smem>>>(tm, tm, args);
    else if (ew == 12) tc_apply_kernel<FieldBLS, 12, false, true><<<grid_, (12 + kTcLoadWarps + 1) * 32, smem>>>(tm, tm, args);
    else tc_apply_kernel<FieldBLS, 16, false, true><<<grid_, (16 + kTcLoadWarps + 1) * 32, smem>>>(tm, tm, args);
  };
  CK(cudaFuncSetAttribute(tc_apply_kernel<FieldBLS, 8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(tc_apply_kernel<FieldBLS, 12, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(tc_apply_kernel<FieldBLS, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t tiles = (batch + 127) / 128;
  const unsigned grid = (unsigned)(tiles < (size_t)sms ? tiles : sms);

  launch(a, grid);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 3;
  }
  // raw accumulators of the first tile vs a plain u8 dot product
  {
    std::vector<uint32_t> dbg((size_t)128 * n_blocks * NB);
    CK(cudaMemcpy(dbg.data(), d_dbg, dbg.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    const size_t rows = batch < 128 ? batch : 128;
    for (size_t r = 0; r < rows; r++) {
      const uint8_t* arow = (const uint8_t*)&in[r * d];
      for (unsigned n = 0; n < n_blocks * NB; n++) {
        if (n / 32 >= n_out) continue;
        uint32_t s = 0;
        for (unsigned k = 0; k < K; k++) s += (uint32_t)arow[k] * bplain[(size_t)n * K + k];
        if (s != dbg[r * n_blocks * NB + n]) {
          if (bad < 8) printf("  raw mismatch row %zu col %u: got %u want %u\n", r, n, dbg[r * n_blocks * NB + n], s);
          bad++;
        }
      }
    }
    printf("raw accumulators (first tile): %s (%zu mismatches)\n", bad ? "FAIL" : "ok", bad);
  }
  // full outputs vs the host field code (sampled rows), from a launch without the debug dump
  // (= the production epilogue)
  if (!(hyp & ~0u)) {
    a.debug = nullptr;
    CK(cudaMemset(d_out, 0xAB, out_bytes));
    launch(a, grid);
    CK(cudaDeviceSynchronize());
  }
  {
    std::vector<Fe> out((size_t)batch * n_out);
    CK(cudaMemcpy(out.data(), d_out, out_bytes, cudaMemcpyDeviceToHost));
    size_t bad = 0, checked = 0;
    const size_t step = batch > 4096 ? batch / 1024 : 1;
    for (size_t r = 0; r < batch; r++) {
      if (!(r < 256 || r % step == 0 || r + 130 > batch)) continue;
      for (unsigned i = 0; i < n_out; i++) {
        Fe acc = fe_zero();
        for (unsigned j = 0; j < d; j++)
          acc = f.add(acc, f.mul(f.to_mont(in[r * d + j]), m[(size_t)i * d + j]));
        checked++;
        if (!fe_eq(acc, out[r * n_out + i])) {
          if (bad < 4) printf("  output mismatch row %zu out %u\n", r, i);
          bad++;
        }
      }
    }
    printf("outputs: %s (%zu of %zu sampled wrong)\n", bad ? "FAIL" : "ok", bad, checked);
  }
  // per-role time stamps of CTA 0 (steady state: after the timing runs are warm)
  {
    long long* d_tr;
    CK(cudaMalloc(&d_tr, 3 * 64 * 8 * 8));
    CK(cudaMemset(d_tr, 0, 3 * 64 * 8 * 8));
    a.debug = nullptr;
    a.trace = d_tr;
    launch(a, grid);
    CK(cudaDeviceSynchronize());
    std::vector<long long> tr(3 * 64 * 8);
    CK(cudaMemcpy(tr.data(), d_tr, tr.size() * 8, cudaMemcpyDeviceToHost));
    a.trace = nullptr;
    const long long t0 = tr[(0 * 64 + 0) * 8 + 0];
    if (getenv("TC_TRACE")) {
      printf("tile | loader: empty issued | mma: full tempty0 released tempty1 issued0 issued1 | epi(warp 0): tfull0 done0 arrived0 tfull1 done1 arrived1\n");
      for (int t = 0; t < 24; t++) {
        printf("%3d |", t);
        for (int e = 0; e < 2; e++) printf(" %7lld", tr[(0 * 64 + t) * 8 + e] ? tr[(0 * 64 + t) * 8 + e] - t0 : 0);
        printf(" |");
        for (int e = 0; e < 6; e++) printf(" %7lld", tr[(1 * 64 + t) * 8 + e] ? tr[(1 * 64 + t) * 8 + e] - t0 : 0);
        printf(" |");
        for (int e = 0; e < 6; e++) printf(" %7lld", tr[(2 * 64 + t) * 8 + e] ? tr[(2 * 64 + t) * 8 + e] - t0 : 0);
        printf("\n");
      }
    }
  }
  // timing (no debug dump)
  a.debug = nullptr;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) {
    a.in = d_in + (w % copies) * in_bytes;
    a.out = d_out + (w % copies) * out_bytes;
    launch(a, grid);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int w = 0; w < reps; w++) {
    a.in = d_in + (w % copies) * in_bytes;
    a.out = d_out + (w % copies) * out_bytes;
    launch(a, grid);
  }
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / reps;
  printf("time %.2f us per launch; %.1f GB/s algorithmic (in+out); %.3g rows/s\n", us,
         (in_bytes + out_bytes) / us * 1e-3, batch / us * 1e6);
  unsigned err = 0;
  CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost));
  if (err) printf("barrier timeout flag set\n");
  return 0;
}
