// Micro-benchmark of the tensor-core epilogue's two ingredients on one SM:
//   * tcgen05.ld.32x32b.x32 (one output's 32 column sums per row, 4 KB per warp), and
//   * the fold + Barrett reduction of tc_kernels.cuh (tc_fold_reduce),
// alone, back to back, and software-pipelined, for 1..4 warps per SM sub-partition.
// Answers: how many cycles a TMEM read costs per sub-partition, whether it overlaps the
// integer work of the same warp / of other warps, and what the epilogue floor per output is.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//        -diag-suppress 20011,20013,20014 -o tools/tmem_probe tools/tmem_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../honeybadgermpc_b200/csrc/tc_kernels.cuh"

using namespace hb;

// mode: 0 = loads only, 1 = fold only, 2 = load, wait, fold (synchronous), 3 = pipelined (two
// register sets), 4 = like 2 with x64 loads (two outputs per instruction), 5 = loads only, x64
template <bool NARROW>
__global__ void __launch_bounds__(544, 1) probe(int mode, int iters, unsigned mu, uint32_t* sink, long long* cycles,
                                                uint8_t* out) {
  __shared__ unsigned slot;
  __shared__ __align__(1024) uint8_t staging[32768];
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned n_warps = blockDim.x / 32 - 1;
  if (warp == n_warps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = slot;
  if (warp < n_warps) {
    const unsigned tq = tmem + (((warp & 3) * 32) << 16);
    uint32_t c0[32], c1[32];
    uint32_t acc = 0;
    // contents of TMEM are whatever the previous kernel left: mask to 24 bits for the fold
    auto use = [&](uint32_t* c) {
      Fe r;
#pragma unroll
      for (int i = 0; i < 32; i++) c[i] &= 0xffffffu;
      tc_fold_reduce<FieldBLS, NARROW>(c, mu, r);
#pragma unroll
      for (int i = 0; i < 8; i++) acc ^= r.w[i];
    };
    const long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < iters; i++) {
        tc_ld32_async(tq + ((i * 32 + (warp >> 2) * 64) & 511 & ~31), c0);
        tc_ld_wait(c0);
        acc ^= c0[0] ^ c0[31];
      }
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 32; i++) c0[i] = threadIdx.x * 2654435761u + i * 40503u;
      for (int i = 0; i < iters; i++) {
        use(c0);
#pragma unroll
        for (int k = 0; k < 32; k++) c0[k] += acc + k;
      }
    } else if (mode == 2) {
      for (int i = 0; i < iters; i++) {
        tc_ld32_async(tq + ((i * 32 + (warp >> 2) * 64) & 511 & ~31), c0);
        tc_ld_wait(c0);
        use(c0);
      }
    } else if (mode == 3) {
      tc_ld32_async(tq, c0);
      for (int i = 0; i < iters; i += 2) {
        tc_ld_wait(c0);
        tc_ld32_async(tq + (((i + 1) * 32) & 511), c1);
        use(c0);
        tc_ld_wait(c1);
        tc_ld32_async(tq + (((i + 2) * 32) & 511), c0);
        use(c1);
      }
      tc_ld_wait(c0);
      acc ^= c0[0];
    } else if (mode == 4 || mode == 5 || mode == 6) {
      // the staged epilogue of tc_apply_kernel for a block of 8 outputs (2 per warp), without the
      // MMA side: 4 = as is, 5 = without the global stores, 6 = direct 32-byte stores instead
      const unsigned quarter = warp & 3, group = warp >> 2;
      const unsigned st_w0 = lane * 128 + (((2 * group) ^ (lane & 7)) << 4);
      const unsigned st_w1 = lane * 128 + (((2 * group + 1) ^ (lane & 7)) << 4);
      const unsigned rd_chunk = lane & 7, rd_row0 = group * 8 + (lane >> 3);
      const unsigned st_r0 = rd_row0 * 128 + ((rd_chunk ^ (rd_row0 & 7)) << 4);
      const unsigned st_r1 = (rd_row0 + 4) * 128 + ((rd_chunk ^ ((rd_row0 + 4) & 7)) << 4);
      unsigned wave = 0;
      for (int i = 0; i < iters; i += 2) {  // one block = 2 items per warp
        tc_ld32_async(tq + (group * 32), c0);
        tc_ld32_async(tq + ((group + 4) * 32), c1);
        tc_ld_wait(c0);
        tc_ld_wait(c1);
        const size_t tile = (size_t)(i / 4) % 64;
        for (int j = 0; j < 2; j++, wave++) {
          uint32_t* c = j ? c1 : c0;
          Fe r;
#pragma unroll
          for (int k = 0; k < 32; k++) c[k] &= 0xffffffu;
          tc_fold_reduce<FieldBLS, NARROW>(c, mu, r);
          if (mode == 6) {
            const size_t row = tile * 128 + quarter * 32 + lane;
            tc_st256(out + row * 512 + ((i / 2) & 1) * 256 + (4 * j + group) * 32, r);
            continue;
          }
          const unsigned stg = tc_smem_u32(staging) + (wave & 1) * 16384 + quarter * 4096;
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + st_w0), "r"(r.w[0]), "r"(r.w[1]),
                       "r"(r.w[2]), "r"(r.w[3]) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + st_w1), "r"(r.w[4]), "r"(r.w[5]),
                       "r"(r.w[6]), "r"(r.w[7]) : "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
          if (mode == 4) {
            const size_t row = tile * 128 + quarter * 32 + rd_row0;
            uint8_t* dst = out + row * 512 + ((i / 2) & 1) * 256 + j * 128 + rd_chunk * 16;
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(stg + st_r0) : "memory");
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(stg + st_r1) : "memory");
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * 512), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
          }
        }
      }
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (lane == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == n_warps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  uint32_t* sink;
  long long* cyc;
  cudaMalloc(&sink, 1 << 20);
  cudaMalloc(&cyc, 148 * 32 * 8);
  const unsigned mu = 37048468u;
  uint8_t* out;
  cudaMalloc(&out, 64 * 128 * 512);
  const char* names[] = {"tmem read only", "fold only", "read, wait, fold", "pipelined read / fold",
                         "staged epilogue", "staged, no global store", "direct 32-byte stores"};
  for (int narrow = 0; narrow < 2; narrow++)
    for (int mode = 0; mode < 7; mode++)
      for (int wps = (mode >= 4 ? 4 : 1); wps <= 4; wps++) {
        const int warps = 4 * wps;
        for (int rep = 0; rep < 2; rep++) {
          if (narrow)
            probe<true><<<1, (warps + 1) * 32>>>(mode, iters, mu, sink, cyc, out);
          else
            probe<false><<<1, (warps + 1) * 32>>>(mode, iters, mu, sink, cyc, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("CUDA error %s\n", cudaGetErrorString(e));
            return 1;
          }
        }
        long long h[32];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
        // items per sub-partition = wps * iters
        printf("%-24s %s warps/subpartition %d: %7.1f cycles per item per warp, %6.1f per item per sub-partition\n",
               names[mode], narrow ? "narrow" : "wide  ", wps, (double)mx / iters, (double)mx / iters / wps);
      }
  return 0;
}
