// Batch kernels of the share-reconstruction path (sm_100a).
//
// Data layout (HBM): an element is 32 bytes (8 x u32 little endian = 4 x u64),
// canonical residue; batch arrays are dense row-major [batch][width], so one
// polynomial's coefficients / one row of the (batch x n) share matrix is one
// contiguous run of 128-bit vectors.  All constants (matrices, twiddles) are
// in Montgomery form, data stays in standard form: mont_mul(data, constant)
// is the plain product, so no conversion pass ever touches the batch.
#pragma once
#include "fp256.cuh"
#include "rowmath.cuh"

namespace hb {

HB_D Fe ld_fe(const uint4* p) {
  uint4 lo = p[0], hi = p[1];
  Fe r;
  r.w[0] = lo.x; r.w[1] = lo.y; r.w[2] = lo.z; r.w[3] = lo.w;
  r.w[4] = hi.x; r.w[5] = hi.y; r.w[6] = hi.z; r.w[7] = hi.w;
  return r;
}

HB_D void st_fe(uint4* p, const Fe& r) {
  p[0] = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
  p[1] = make_uint4(r.w[4], r.w[5], r.w[6], r.w[7]);
}

// Shared-memory element load that the compiler may neither hoist nor merge with an
// earlier load of the same address: lane-per-row kernels RE-READ their inputs from the
// TMA-staged tile for every use instead of keeping them in registers (8 registers per
// element), which is what keeps every row-warp of the batch resident at once.
HB_D Fe lds_fe_reload(const uint4* p) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(p);
  Fe r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
               : "r"(a));
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];"
               : "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "r"(a));
  return r;
}

// One element = one 32-byte sector = ONE 256-bit global store (STG.E.256, sm_100+).
// Lane-per-row kernels store their results straight from registers with it: every
// store fills a whole sector, so no shared-memory transpose is needed to keep the
// DRAM/L2 traffic at the algorithmic bytes.
HB_D void st_fe_global256(uint4* p, const Fe& r) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]),
               "r"(r.w[2]), "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
               : "memory");
}

// ---------------------------------------------------------------------------
// apply_matrix: out[b][i] = sum_j M[i][j] * in[b][col(j)]
//
// Serves vandermonde_batch_evaluate (M = V(x)), vandermonde_batch_interpolate
// and fft_batch_interpolate (M = V(x)^-1), and the hyper-invertible-matrix
// step of RanDouSha (rsdecode_impl.h:23-36, :97-122; pyx:183, :237 -- the
// reference runs these as one NTL mat_ZZ_p mul).
// One thread per output element; the row dot product is accumulated lazily
// (acc_mac) and reduced once.  M is stored word-interleaved ([j][word][i]) so
// the 32 lanes of a warp read 32 consecutive words.
// ---------------------------------------------------------------------------
struct MatvecArgs {
  const uint32_t* mt;        // [d][8][n_out]
  const uint4* in;           // [batch][in_stride] elements
  uint4* out;                // [batch][out_stride] elements
  const int* in_cols;        // optional gather of input columns (device), or null
  unsigned long long batch;
  int n_out, d;
  int in_stride, out_stride;
  int rows_per_cta;          // max(1, 256 / n_out)
  unsigned magic;            // ceil(2^20 / n_out): t / n_out == (t * magic) >> 20 for t < 256
};

template <class F>
HB_D Fe row_dot(const MatvecArgs& a, const uint4* row, int i) {
  const uint32_t* mcol = a.mt + i;
  Acc acc;
  acc_zero(acc);
  int pending = 0;
  for (int j = 0; j < a.d; j++) {
    int c = a.in_cols ? a.in_cols[j] : j;
    Fe x = ld_fe(row + 2 * c);
    Fe m;
#pragma unroll
    for (int w = 0; w < 8; w++) m.w[w] = mcol[(size_t)(j * 8 + w) * a.n_out];
    acc_mac(acc, x, m);
    if (++pending == F::kFold) {
      acc_fold<F>(acc);
      pending = 0;
    }
  }
  if (pending) acc_fold<F>(acc);
  return acc_redc<F>(acc);
}

template <class F>
__global__ void __launch_bounds__(256) apply_matrix_kernel(MatvecArgs a) {
  if (a.n_out <= 256) {
    int rl = (int)((threadIdx.x * a.magic) >> 20);
    int i = (int)threadIdx.x - rl * a.n_out;
    unsigned long long b = (unsigned long long)blockIdx.x * a.rows_per_cta + rl;
    if (rl >= a.rows_per_cta || b >= a.batch) return;
    Fe r = row_dot<F>(a, a.in + 2ull * b * (unsigned)a.in_stride, i);
    st_fe(a.out + 2ull * (b * (unsigned)a.out_stride + i), r);
  } else {
    unsigned long long b = blockIdx.x;
    for (int i = threadIdx.x; i < a.n_out; i += 256) {
      Fe r = row_dot<F>(a, a.in + 2ull * b * (unsigned)a.in_stride, i);
      st_fe(a.out + 2ull * (b * (unsigned)a.out_stride + i), r);
    }
  }
}

// ---------------------------------------------------------------------------
// apply_matrix, shared-memory variant (the hot interpolate / small-matrix path):
// the matrix (<= 48 KB) and the CTA's tile of input rows live in shared memory.
// The input tile -- rows_per_cta consecutive rows = one contiguous byte range of
// the batch array -- is fetched by ONE TMA bulk copy (cp.async.bulk, completion
// on an mbarrier) issued by thread 0 while all threads stage the matrix.
// ---------------------------------------------------------------------------
HB_D void mbar_init(uint64_t* bar, unsigned count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

HB_D void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
      "l"(gmem_src), "r"(bytes), "r"(b)
      : "memory");
}

HB_D void mbar_wait(uint64_t* bar, unsigned phase) {
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(b),
      "r"(phase)
      : "memory");
}

template <class F>
__global__ void __launch_bounds__(256) apply_matrix_smem_kernel(MatvecArgs a) {
  extern __shared__ uint4 smem[];
  __shared__ alignas(8) uint64_t bar;
  const int n_out = a.n_out, d = a.d;
  uint4* xt = smem;                                         // [rows_per_cta][d] elements
  uint32_t* sm = (uint32_t*)(smem + 2 * a.rows_per_cta * d);  // [d][8][n_out] words
  const unsigned long long row0 = (unsigned long long)blockIdx.x * a.rows_per_cta;
  unsigned long long rows_here = a.batch - row0;
  if (rows_here > (unsigned long long)a.rows_per_cta) rows_here = a.rows_per_cta;
  const unsigned tile_bytes = (unsigned)rows_here * d * 32u;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) tma_load_1d(xt, a.in + 2ull * row0 * d, tile_bytes, &bar);
  for (int q = threadIdx.x; q < d * 8 * n_out; q += 256) sm[q] = a.mt[q];
  __syncthreads();
  mbar_wait(&bar, 0);

  int rl = (int)((threadIdx.x * a.magic) >> 20);
  int i = (int)threadIdx.x - rl * n_out;
  if (rl >= (int)rows_here) return;
  const uint4* row = xt + 2 * rl * d;
  const uint32_t* mcol = sm + i;
  Acc acc;
  acc_zero(acc);
  int pending = 0;
#pragma unroll 2
  for (int j = 0; j < d; j++) {
    Fe x = ld_fe(row + 2 * j);
    Fe m;
#pragma unroll
    for (int w = 0; w < 8; w++) m.w[w] = mcol[(j * 8 + w) * n_out];
    acc_mac(acc, x, m);
    if (++pending == F::kFold) {
      acc_fold<F>(acc);
      pending = 0;
    }
  }
  if (pending) acc_fold<F>(acc);
  Fe r = acc_redc<F>(acc);
  st_fe(a.out + 2ull * ((row0 + rl) * (unsigned)a.out_stride + i), r);
}

// ---------------------------------------------------------------------------
// 16-point NTT in registers (the n = 16 encode of the headline config;
// fft_batch_evaluate with n == 16).  Decimation in frequency, fully unrolled.
// A polynomial is shared by TWO threads in two different warps: after the first
// DIF stage the transform splits into two independent 8-point transforms
//   U_i = c_i + c_{i+8}            -> the even outputs
//   L_i = (c_i - c_{i+8}) omega^i  -> the odd outputs
// so warp 2w computes the U halves and warp 2w+1 the L halves of the same 32
// polynomials (no divergence, no redundant multiplies, twice the parallelism of
// a thread-per-polynomial mapping -- the batch alone is only ~14 warps per SM).
// D = number of input coefficients is a template parameter, structurally-zero
// operands are dropped at compile time (d = 6: 15 modular multiplications per
// polynomial instead of 32).  The twiddles come from the kernel-parameter
// constant bank.  Rows are staged through shared memory: TMA bulk copy in,
// odd-stride tile + coalesced 128-bit stores out.
// ---------------------------------------------------------------------------
struct Ntt16Args {
  const uint4* in;    // [batch][stride]
  uint4* out;         // [batch][k_out]
  unsigned long long batch;
  int d, k_out, stride;  // d <= 16 coefficients used, rows are `stride` elements apart
  uint32_t tw[16][8];     // omega^i, i < 16, Montgomery form
};

// Row-per-thread kernels keep global traffic coalesced by staging through shared
// memory: the CTA's input rows (one contiguous byte range) arrive with one TMA
// bulk copy; results are parked in a tile whose row stride is an ODD number of
// 16-byte chunks (conflict-free for lane-per-row writes) and then streamed out
// by all threads with consecutive 128-bit stores.
template <int THREADS>
HB_D void tile_store_rows(uint4* tile, uint4* gout, int rows_here, int chunks_per_row) {
  const int total = rows_here * chunks_per_row;
  const int pstride = chunks_per_row | 1;  // odd
  for (int q = threadIdx.x; q < total; q += THREADS) {
    int r = q / chunks_per_row, c = q - r * chunks_per_row;
    gout[q] = tile[r * pstride + c];
  }
}

// Fused all-gather: instead of one local destination the result tile is stored
// into every rank's gather buffer through NVLink-mapped peer pointers, or -- when
// the buffers are bound to an NVSwitch multicast object -- with ONE multimem.st
// per 16 bytes that the switch replicates to all ranks.
struct GatherDst {
  uint4* peers[8];   // peer-mapped base pointers of every rank's [world*batch][k] array
  uint4* mc;         // multicast address of the same buffers, or null
  int world;
};

HB_D void st_multimem(uint4* p, const uint4& v) {
  asm volatile("multimem.st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

template <int THREADS>
HB_D void tile_store_rows_gather(uint4* tile, const GatherDst& g, unsigned long long chunk0, int rows_here,
                                 int chunks_per_row) {
  const int total = rows_here * chunks_per_row;
  const int pstride = chunks_per_row | 1;
  for (int q = threadIdx.x; q < total; q += THREADS) {
    int r = q / chunks_per_row, c = q - r * chunks_per_row;
    uint4 v = tile[r * pstride + c];
    if (g.mc) {
      st_multimem(g.mc + chunk0 + q, v);
    } else {
      for (int w = 0; w < g.world; w++) g.peers[w][chunk0 + q] = v;
    }
  }
}

// All-gather of an already decoded block as a copy kernel: a handful of CTAs read
// the local block and store it into every rank's gather buffer (one multimem.st
// per 16 bytes when a multicast address is available, so each rank's egress is 1x
// its block and the NVSwitch does the replication).  Launched on a side stream it
// overlaps the next step's kernels while occupying only a few SMs.
__global__ void __launch_bounds__(256) gather_copy_kernel(const uint4* src, GatherDst g,
                                                          unsigned long long chunk0,
                                                          unsigned long long chunks) {
  for (unsigned long long q = (unsigned long long)blockIdx.x * 256 + threadIdx.x; q < chunks;
       q += (unsigned long long)gridDim.x * 256) {
    uint4 v = src[q];
    if (g.mc) {
      st_multimem(g.mc + chunk0 + q, v);
    } else {
      for (int w = 0; w < g.world; w++) g.peers[w][chunk0 + q] = v;
    }
  }
}

// The same copy with the slot hand-over done by flags in symmetric memory instead of
// host-issued barriers.  Every rank owns a flag array (uint32, zero-initialised), mapped
// into every process:
//   arrived [slot][r]  = how many blocks rank r has delivered into MY buffer of that slot,
//                        counted over the whole run (monotonic)
//   released[slot][r]  = how many fills of that slot rank r has finished reading
// plus two words per slot that only the owner touches: sent (blocks this rank has copied
// out for the slot) and waited (fills this rank has consumed).  Every value a kernel waits
// for or publishes is derived from those device-side counters, never from a kernel
// argument, so a CUDA graph of these launches can be replayed any number of times.
//   copy kernel:   first part of a fill: wait until every rank has released the previous
//                  fill; copy; all CTAs fence.sys; the last CTA (device counter) bumps
//                  `sent` and publishes it in arrived[slot][me] on every rank (st.release.sys)
//   wait kernel:   until arrived[slot][r] >= (waited + 1) * parts for every r
//   release kernel: waited += 1, published in released[slot][me] on every rank
// Waits are bounded: a dead peer sets *error instead of hanging the GPU.
struct GatherSignal {
  unsigned* flags[8];      // every rank's flag array as mapped here; flags[rank] is the local one
  unsigned* counter;       // device word, 0 between launches (last-CTA detection)
  unsigned* error;
  unsigned slot_base;      // index of arrived[slot][0]; released[slot][0] is at slot_base + world
  unsigned local_base;     // index of this slot's {sent, waited} pair
  unsigned parts;          // blocks per rank and fill
  int first_part;          // copy kernel: wait for the release of the previous fill
  int world, rank;
};

HB_D unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

HB_D void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// spin until *p >= want (32-bit counters compared modulo 2^32); false on time-out (~2 s)
HB_D bool spin_until(const unsigned* p, unsigned want, unsigned* error) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(p) - want) < 0) {
    if (clock64() - t0 > 4000000000ll) {
      if (error) atomicExch(error, 2u);
      return false;
    }
    __nanosleep(64);
  }
  return true;
}

__global__ void __launch_bounds__(256) gather_copy_signal_kernel(const uint4* src, GatherDst g,
                                                                 unsigned long long chunk0,
                                                                 unsigned long long chunks, GatherSignal s) {
  unsigned* local = s.flags[s.rank];
  // `sent` only changes when the last CTA of this launch finishes, i.e. after every CTA read it
  const unsigned sent = ld_acquire_sys(local + s.local_base);
  if (s.first_part) {
    if (threadIdx.x < (unsigned)s.world)
      spin_until(local + s.slot_base + s.world + threadIdx.x, sent / s.parts, s.error);
    __syncthreads();
  }
  // four independent 16-byte loads in flight per thread before the (remote) stores
  const unsigned long long stride = (unsigned long long)gridDim.x * 256;
  for (unsigned long long q = (unsigned long long)blockIdx.x * 256 + threadIdx.x; q < chunks; q += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (q + u * stride < chunks) v[u] = src[q + u * stride];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const unsigned long long qq = q + u * stride;
      if (qq >= chunks) break;
      if (g.mc) {
        st_multimem(g.mc + chunk0 + qq, v[u]);
      } else {
        for (int w = 0; w < g.world; w++) g.peers[w][chunk0 + qq] = v[u];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(s.counter, 1u);
    if (done == gridDim.x - 1) {  // every CTA's stores are fenced: publish
      *s.counter = 0;
      st_release_sys(local + s.local_base, sent + 1);
      __threadfence_system();
      for (int w = 0; w < s.world; w++) st_release_sys(s.flags[w] + s.slot_base + s.rank, sent + 1);
    }
  }
}

// The same hand-over around a copy done by the TMA unit: per CTA one thread streams 16 KB chunks of the
// local block into a ring of shared-memory stages (cp.async.bulk global -> shared) and another
// thread forwards every chunk to each PEER's buffer (cp.async.bulk shared -> global over NVLink),
// so a handful of SMs keeps megabytes in flight without occupying registers or load/store slots.
// Unicast only (bulk copies do not take multicast addresses): egress = (world - 1) blocks, ingress =
// (world - 1) blocks -- the multicast kernel above sends 1 block but receives world blocks (its own
// comes back through the switch), which is the larger of the two floors from 4 ranks on.
constexpr unsigned kGbStages = 4, kGbChunk = 16384;

__global__ void __launch_bounds__(64) gather_bulk_signal_kernel(const uint8_t* src, GatherDst g,
                                                                unsigned long long offset_bytes,
                                                                unsigned long long bytes, GatherSignal s) {
  extern __shared__ __align__(128) uint8_t gb_smem[];
  __shared__ uint64_t gb_full[kGbStages], gb_empty[kGbStages];
  unsigned* local = s.flags[s.rank];
  const unsigned sent = ld_acquire_sys(local + s.local_base);
  if (threadIdx.x == 0) {
    for (unsigned i = 0; i < kGbStages; i++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&gb_full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&gb_empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (s.first_part && threadIdx.x < (unsigned)s.world)
    spin_until(local + s.slot_base + s.world + threadIdx.x, sent / s.parts, s.error);
  __syncthreads();
  const unsigned long long n_chunks = (bytes + kGbChunk - 1) / kGbChunk;
  auto wait = [&](uint64_t* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    const long long t0 = clock64();
    for (;;) {
      unsigned done;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(a), "r"(parity)
          : "memory");
      if (done) return;
      if (clock64() - t0 > 4000000000ll) {
        if (s.error) atomicExch(s.error, 3u);
        __trap();
      }
    }
  };
  if (threadIdx.x == 0) {
    // loader
    unsigned it = 0;
    for (unsigned long long i = blockIdx.x; i < n_chunks; i += gridDim.x, it++) {
      const unsigned st = it % kGbStages, ph = (it / kGbStages) & 1;
      wait(&gb_empty[st], ph ^ 1);
      const unsigned long long o = i * kGbChunk;
      const unsigned size = (unsigned)(bytes - o < kGbChunk ? bytes - o : kGbChunk);
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&gb_full[st]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(size) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              (unsigned)__cvta_generic_to_shared(gb_smem + st * kGbChunk)),
          "l"(src + o), "r"(size), "r"(bar)
          : "memory");
    }
  } else if (threadIdx.x == 32) {
    // forwarder
    unsigned it = 0;
    for (unsigned long long i = blockIdx.x; i < n_chunks; i += gridDim.x, it++) {
      const unsigned st = it % kGbStages, ph = (it / kGbStages) & 1;
      wait(&gb_full[st], ph);
      const unsigned long long o = i * kGbChunk;
      const unsigned size = (unsigned)(bytes - o < kGbChunk ? bytes - o : kGbChunk);
      const unsigned sm = (unsigned)__cvta_generic_to_shared(gb_smem + st * kGbChunk);
      for (int k = 1; k < g.world; k++) {
        const int w = (s.rank + k) % g.world;  // staggered: the ranks target different peers at any moment
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                         (uint8_t*)g.peers[w] + offset_bytes + o),
                     "r"(sm), "r"(size)
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (it > 0) {  // the previous chunk's copies have finished reading their stage
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&gb_empty[(it - 1) % kGbStages]))
                     : "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every copy has been performed
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 32) {
    const unsigned done = atomicAdd(s.counter, 1u);
    if (done == gridDim.x - 1) {  // every CTA's copies are complete and fenced: publish
      *s.counter = 0;
      st_release_sys(local + s.local_base, sent + 1);
      __threadfence_system();
      for (int w = 0; w < s.world; w++) st_release_sys(s.flags[w] + s.slot_base + s.rank, sent + 1);
    }
  }
}

// The hand-over around a copy done by the copy engines (cudaMemcpyAsync to the peers'
// buffers, stream-ordered between these two one-warp kernels):
// before -- first part of a fill: wait until every rank has released the previous fill
__global__ void gather_wait_released_kernel(GatherSignal s) {
  unsigned* local = s.flags[s.rank];
  const unsigned sent = ld_acquire_sys(local + s.local_base);
  if (threadIdx.x < (unsigned)s.world)
    spin_until(local + s.slot_base + s.world + threadIdx.x, sent / s.parts, s.error);
}

// after -- one more block of this rank has landed everywhere: count it and tell every rank
__global__ void gather_signal_arrived_kernel(GatherSignal s) {
  unsigned* local = s.flags[s.rank];
  const unsigned sent = ld_acquire_sys(local + s.local_base) + 1;
  __syncwarp();
  if (threadIdx.x == 0) st_release_sys(local + s.local_base, sent);
  __threadfence_system();
  if (threadIdx.x < (unsigned)s.world) st_release_sys(s.flags[threadIdx.x] + s.slot_base + s.rank, sent);
}

// consumer side: wait until every rank's blocks of the next fill have landed here
__global__ void gather_wait_kernel(GatherSignal s) {
  unsigned* local = s.flags[s.rank];
  const unsigned waited = ld_acquire_sys(local + s.local_base + 1);
  if (threadIdx.x < (unsigned)s.world)
    spin_until(local + s.slot_base + threadIdx.x, (waited + 1) * s.parts, s.error);
}

// consumer side: this rank is done reading the fill
__global__ void gather_release_kernel(GatherSignal s) {
  unsigned* local = s.flags[s.rank];
  const unsigned waited = ld_acquire_sys(local + s.local_base + 1) + 1;
  __syncwarp();
  if (threadIdx.x == 0) st_release_sys(local + s.local_base + 1, waited);
  if (threadIdx.x < (unsigned)s.world)
    st_release_sys(s.flags[threadIdx.x] + s.slot_base + s.world + s.rank, waited);
}

// ---------------------------------------------------------------------------
// device-resident IncrementalDecoder (reed_solomon.py:305-331): columns live as
// colbuf[n][batch]; these two kernels are pure data movement / comparison.
// ---------------------------------------------------------------------------
struct ColumnIdx {
  int idx[256];
};

// rows[b][j] = colbuf[idx[j]][b]: thread = (row, column), 32 bytes each way; a warp reads
// 32 consecutive elements of one column (coalesced) when k divides the warp evenly or not.
__global__ void __launch_bounds__(256) columns_to_rows_kernel(const uint4* colbuf, unsigned long long batch,
                                                              ColumnIdx ci, int k, uint4* rows) {
  const unsigned long long total = batch * (unsigned)k;
  for (unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * 256) {
    const unsigned j = (unsigned)(t / batch);       // column-major sweep: coalesced reads
    const unsigned long long b = t - (unsigned long long)j * batch;
    const uint4* src = colbuf + 2ull * ((unsigned long long)ci.idx[j] * batch + b);
    uint4* dst = rows + 2ull * (b * (unsigned)k + j);
    dst[0] = src[0];
    dst[1] = src[1];
  }
}

// flags[j] = 1 when column idx[j] differs anywhere from rows[.][col_offset + idx[j]]
__global__ void __launch_bounds__(256) compare_columns_kernel(const uint4* rows, int row_width, int col_offset,
                                                              const uint4* colbuf, unsigned long long batch,
                                                              ColumnIdx ci, int m, int* flags) {
  const unsigned long long total = batch * (unsigned)m;
  for (unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * 256) {
    const unsigned j = (unsigned)(t / batch);
    const unsigned long long b = t - (unsigned long long)j * batch;
    const int col = ci.idx[j];
    const uint4* x = colbuf + 2ull * ((unsigned long long)col * batch + b);
    const uint4* y = rows + 2ull * (b * (unsigned)row_width + (unsigned)(col_offset + col));
    const uint4 x0 = x[0], x1 = x[1], y0 = y[0], y1 = y[1];
    const bool same = x0.x == y0.x && x0.y == y0.y && x0.z == y0.z && x0.w == y0.w && x1.x == y1.x &&
                      x1.y == y1.y && x1.z == y1.z && x1.w == y1.w;
    if (!same) flags[j] = 1;
  }
}

// ---------------------------------------------------------------------------
// NTT-structured interpolation, device side of fnt_decode_step2
// (rsdecode_impl.h:226-265): elementwise pieces around the NTT launches.
// ---------------------------------------------------------------------------
struct FntArgs {
  const uint4* in;     // [batch][k]   or  [batch][m]
  uint4* out;          // see each kernel
  const uint4* cst;    // Montgomery-form constants (per column)
  const int* zs;       // scatter positions (device), k entries
  unsigned long long batch;
  int k, n, m;
};

// N[b][zs[i]] = ys[b][i] * (1 / A'(x_i)); `out` ([batch][n]) was zeroed beforehand
template <class F>
__global__ void __launch_bounds__(256) fnt_scale_scatter_kernel(FntArgs a) {
  const unsigned long long total = a.batch * (unsigned)a.k;
  for (unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * 256) {
    const unsigned long long b = t / (unsigned)a.k;
    const int i = (int)(t - b * (unsigned)a.k);
    const Fe y = ld_fe(a.in + 2ull * t);
    const Fe d = ld_fe(a.cst + 2 * i);
    st_fe(a.out + 2ull * (b * (unsigned)a.n + (unsigned)a.zs[i]), mont_mul<F>(y, d));
  }
}

// Q[b][j] = -R[b][(j + 1) % n], j < k; R has kr = min(k + 1, n) entries per row
template <class F>
__global__ void __launch_bounds__(256) fnt_shift_negate_kernel(FntArgs a) {
  const unsigned long long total = a.batch * (unsigned)a.k;
  const int kr = a.k + 1 < a.n ? a.k + 1 : a.n;
  for (unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * 256) {
    const unsigned long long b = t / (unsigned)a.k;
    const int j = (int)(t - b * (unsigned)a.k);
    const Fe r = ld_fe(a.in + 2ull * (b * (unsigned)kr + (unsigned)((j + 1) % a.n)));
    st_fe(a.out + 2ull * t, fe_neg<F>(r));
  }
}

// buf[b][i] *= cst[i], i < m (in place)
template <class F>
__global__ void __launch_bounds__(256) fnt_pointwise_kernel(FntArgs a) {
  const unsigned long long total = a.batch * (unsigned)a.m;
  for (unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * 256) {
    const int i = (int)(t % (unsigned)a.m);
    const Fe v = ld_fe(a.in + 2ull * t);
    st_fe(a.out + 2ull * t, mont_mul<F>(v, ld_fe(a.cst + 2 * i)));
  }
}

// within the 8-point transform: after s stages slot idx depends on the inputs
// j = idx (mod 8 >> s); with the first D8 inputs non-zero it is non-zero iff that
// residue is < D8
#define HB_NTT8_NZ(s, idx, D8) ((((idx) & ((8 >> (s)) - 1))) < (D8))

template <class F, int D, int POLYS>
__global__ void __launch_bounds__(2 * POLYS) ntt16_split_kernel(const __grid_constant__ Ntt16Args a) {
  extern __shared__ uint4 smem[];
  __shared__ alignas(8) uint64_t bar;
  constexpr int D8 = D < 8 ? D : 8;
  const unsigned long long row0 = (unsigned long long)blockIdx.x * POLYS;
  unsigned long long left = a.batch - row0;
  const int rows_here = left < (unsigned long long)POLYS ? (int)left : POLYS;
  const bool dense = a.stride == a.d;  // contiguous input rows: one bulk copy
  if (dense) {
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0)
      tma_load_1d(smem, a.in + 2ull * row0 * a.stride, (unsigned)rows_here * a.d * 32u, &bar);
    mbar_wait(&bar, 0);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp & 1;                  // 0: U (even outputs), 1: L (odd outputs)
  const int poly = (warp >> 1) * 32 + lane;   // row within the CTA
  const bool active = poly < rows_here;
  Fe v[8];
  {
    const uint4* src = dense ? smem + 2 * poly * a.d : a.in + 2ull * (row0 + poly) * a.stride;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (i >= D8) continue;
      Fe lo = (active && i < a.d) ? ld_fe(src + 2 * i) : fe_zero();
      if (i + 8 < D) {
        Fe hi = (active && i + 8 < a.d) ? ld_fe(src + 2 * (i + 8)) : fe_zero();
        v[i] = half ? fe_sub<F>(lo, hi) : fe_add<F>(lo, hi);
      } else {
        v[i] = lo;
      }
    }
  }
  __syncthreads();  // the input tile is dead: its space becomes the output tile
  if (half) {
#pragma unroll
    for (int i = 1; i < 8; i++) {
      if (i >= D8) continue;
      Fe w;
#pragma unroll
      for (int q = 0; q < 8; q++) w.w[q] = a.tw[i][q];
      v[i] = mont_mul<F>(v[i], w);
    }
  }
  // 8-point DIF on omega^2
#pragma unroll
  for (int s = 1; s <= 3; s++) {
    const int h = 8 >> s;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int pos = t & (h - 1);
      const int i0 = ((t & ~(h - 1)) << 1) | pos;
      const int i1 = i0 + h;
      const bool nz0 = HB_NTT8_NZ(s - 1, i0, D8), nz1 = HB_NTT8_NZ(s - 1, i1, D8);
      if (!nz0 && !nz1) continue;
      const int tw_idx = (pos << (s - 1)) * 2;  // (omega^2)^(pos * 2^(s-1))
      Fe lo, hi;
      if (!nz1) {
        lo = v[i0];
        hi = v[i0];
      } else if (!nz0) {
        lo = v[i1];
        hi = fe_neg<F>(v[i1]);
      } else {
        lo = fe_add<F>(v[i0], v[i1]);
        hi = fe_sub<F>(v[i0], v[i1]);
      }
      if (tw_idx != 0) {
        Fe w;
#pragma unroll
        for (int q = 0; q < 8; q++) w.w[q] = a.tw[tw_idx][q];
        hi = mont_mul<F>(hi, w);
      }
      v[i0] = lo;
      v[i1] = hi;
    }
  }
  // slot idx of half h holds X[2 * bitrev3(idx) + h]
  const int chunks = 2 * a.k_out, pstride = chunks | 1;
  uint4* mine = smem + poly * pstride;
#pragma unroll
  for (int idx = 0; idx < 8; idx++) {
    const int i = 2 * (((idx & 1) << 2) | (idx & 2) | ((idx & 4) >> 2)) + half;
    if (i < a.k_out) st_fe(mine + 2 * i, v[idx]);
  }
  __syncthreads();
  tile_store_rows<2 * POLYS>(smem, a.out + 2ull * row0 * a.k_out, rows_here, chunks);
}

// ---------------------------------------------------------------------------
// 16-point NTT of d <= 8 coefficients (the headline encode: d = t+1 = 6), the
// default n = 16 kernel.  Same two-threads-per-polynomial split as
// ntt16_split_kernel (warp 2w: even outputs, warp 2w+1: odd outputs), but each
// thread works through its 8-point transform as two 4-point groups
// (rowmath.cuh: ntt16_half), re-reading the coefficients from the TMA-staged
// tile instead of holding all eight values: ~2/3 of the registers, so every
// row-warp of a 65 536-polynomial batch is resident at once.  Results leave
// straight from registers, one 256-bit store per element (whole 32-byte
// sectors), so there is no output tile and no barrier after the TMA wait.
// ---------------------------------------------------------------------------
// The multiplier is a real function (one copy, ~230 instructions) instead of 15 inlined
// copies: the straight-line kernel shrinks from 77 KB to 28 KB of code and fits the 32 KB
// instruction cache -- with everything inlined 28 % of the issue slots were instruction-
// fetch stalls (profiles/).  The twiddle comes from a 512-byte shared-memory table.
template <class F>
__device__ __noinline__ Fe mul_tw_call(const uint4* twp, Fe x) {
  return mont_mul<F>(ld_fe(twp), x);
}

// Hand-over slots between the two threads of a polynomial (rowmath.cuh: ntt16_half):
// element-major, 16-byte halves apart, so lane-per-row accesses are conflict free.  The
// two warps of a pair meet on a named barrier (id 1 + pair): the even warp only arrives,
// the odd warp waits -- long after the even warp has passed.
template <int POLYS>
struct Ntt16Xch {
  uint4* base;   // [4][2][POLYS] uint4
  int poly, bar_id;
  HB_D void put(int i, const Fe& v) {
    base[(2 * i) * POLYS + poly] = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
    base[(2 * i + 1) * POLYS + poly] = make_uint4(v.w[4], v.w[5], v.w[6], v.w[7]);
  }
  HB_D Fe get(int i) const {
    const uint4 lo = base[(2 * i) * POLYS + poly], hi = base[(2 * i + 1) * POLYS + poly];
    Fe r;
    r.w[0] = lo.x; r.w[1] = lo.y; r.w[2] = lo.z; r.w[3] = lo.w;
    r.w[4] = hi.x; r.w[5] = hi.y; r.w[6] = hi.z; r.w[7] = hi.w;
    return r;
  }
  HB_D void signal() {
    __threadfence_block();
    asm volatile("barrier.cta.arrive %0, 64;" ::"r"(bar_id) : "memory");
  }
  HB_D void wait() { asm volatile("barrier.cta.sync %0, 64;" ::"r"(bar_id) : "memory"); }
};

template <class F, int D, int POLYS, bool BAL = false>
__global__ void __launch_bounds__(2 * POLYS, (D <= 6 ? 7 : 5))
ntt16_g4_kernel(const __grid_constant__ Ntt16Args a) {
  extern __shared__ uint4 smem[];   // input tile [POLYS][d], then the hand-over slots
  __shared__ alignas(8) uint64_t bar;
  __shared__ uint4 twsm[32];  // omega^j, j < 16 (Montgomery form), two 16-byte halves each
  const unsigned long long row0 = (unsigned long long)blockIdx.x * POLYS;
  unsigned long long left = a.batch - row0;
  const int rows_here = left < (unsigned long long)POLYS ? (int)left : POLYS;
  // rows are dense here (stride == d: the launcher sends everything else to the split kernel)
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  if (threadIdx.x < 32) {
    const int j = threadIdx.x >> 1, h = 4 * (threadIdx.x & 1);
    twsm[threadIdx.x] = make_uint4(a.tw[j][h], a.tw[j][h + 1], a.tw[j][h + 2], a.tw[j][h + 3]);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    tma_load_1d(smem, a.in + 2ull * row0 * a.d, (unsigned)rows_here * a.d * 32u, &bar);
  mbar_wait(&bar, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = warp & 1;                  // 0: even outputs, 1: odd outputs
  const int poly = (warp >> 1) * 32 + lane;   // row within the CTA
  // rows past the end of the batch run on row 0's data and store nothing (with BAL both
  // warps of a pair must reach the hand-over barrier)
  const bool active = poly < rows_here;
  if (!BAL && !active) return;
  const int d = a.d, k_out = a.k_out;
  const uint4* src = smem + 2 * (active ? poly : 0) * d;
  uint4* dst = a.out + 2ull * (row0 + poly) * k_out;
  Ntt16Xch<POLYS> xch{smem + 2 * POLYS * d, poly, 1 + (warp >> 1)};
  auto ld = [&](int i) { return i < d ? ld_fe(src + 2 * i) : fe_zero(); };
  auto mul = [&](int j, const Fe& x) { return mul_tw_call<F>(twsm + 2 * j, x); };  // omega^j * x
  auto st = [&](int k, const Fe& v) {
    if (active && k < k_out) st_fe_global256(dst + 2 * k, v);
  };
  if (half)
    ntt16_half<F, D, 1, BAL>(ld, mul, st, xch);
  else
    ntt16_half<F, D, 0, BAL>(ld, mul, st, xch);
}

// ---------------------------------------------------------------------------
// Small square interpolation (k <= 8), one row per thread: the k x k matrix
// V(x)^-1 sits in the kernel-parameter constant bank, so its limbs are direct
// IMAD.WIDE operands (no loads, no registers); the k inputs of the row stay in
// registers for all k outputs.  vandermonde_batch_interpolate and
// fft_batch_interpolate of the headline config (k = t+1 = 6) run here.
// ---------------------------------------------------------------------------
template <int K>
struct SmallInterpArgs {
  const uint4* in;    // [batch][K]
  uint4* out;         // [batch][K]
  unsigned long long batch;
  unsigned long long gather_row0;  // first row of this rank inside the gathered array
  GatherDst gather;                // world == 0: plain local output
  // M[i][j] in the form the arithmetic wants (9 words reserved per entry):
  //   ARITH 0: 8 words, Montgomery form (R = 2^256)
  //   ARITH 1: 9 limbs of 29 bits of M * 2^261 mod p
  uint32_t mc[K][K][9];
};

// ROWS rows per CTA, SPLIT warps share a row (warp w handles the outputs
// i = w % SPLIT, w % SPLIT + SPLIT, ...: no divergence inside a warp); the small
// batch of the headline config is only ~14 warps per SM at one thread per row, so
// SPLIT = 2 doubles the warps available to hide the IMAD dependency latency.
// GATHER = false: every result leaves straight from registers as one 256-bit
// store (a whole 32-byte sector).  GATHER = true (fused all-gather): results are
// parked in a shared tile and streamed to every rank with consecutive 16-byte
// multimem / peer stores, which keeps the NVLink packets large.
// ARITH = 0: the 32-bit-limb lazy accumulator of fp256.cuh (Acc): 64 IMAD.WIDE per term.
// ARITH = 1: carry-free radix-2^29 accumulation (rowmath.cuh: mac29 / redc29): 81 plain
//            IMAD.WIDE per term, far fewer ALU instructions; same speed as 0 on the B200
//            (IMAD.WIDE costs 4 pipe cycles with or without a carry chain).
// Both are bit-identical (tests/test_gpu_ntl.py runs them against the oracle).  A third,
// Karatsuba (48 IMAD.WIDE per term, three accumulators combined once per output), was
// bit-exact too but needed 96 registers and 1 992 instructions per output: 46 us against
// 36 us at 65 536 rows; it is not kept.
template <class F, int K, int ROWS, int SPLIT, bool GATHER, int ARITH>
__global__ void __launch_bounds__(ROWS * SPLIT, (ROWS * SPLIT == 128 ? 7 : 1)) interp_small_kernel(const __grid_constant__ SmallInterpArgs<K> a) {
  constexpr int THREADS = ROWS * SPLIT;
  extern __shared__ uint4 smem[];
  __shared__ alignas(8) uint64_t bar;
  const unsigned long long row0 = (unsigned long long)blockIdx.x * ROWS;
  unsigned long long left = a.batch - row0;
  const int rows_here = left < (unsigned long long)ROWS ? (int)left : ROWS;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) tma_load_1d(smem, a.in + 2ull * row0 * K, (unsigned)rows_here * K * 32u, &bar);
  mbar_wait(&bar, 0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int part = warp % SPLIT;
  const int row = (warp / SPLIT) * 32 + lane;
  const bool active = row < rows_here;
  if (!GATHER && !active) return;  // no barrier follows on this path
  // the inputs stay in the shared tile (re-read per term: two LDS.128) instead of in
  // 8*K registers, which buys resident warps -- the kernel is latency bound
  const uint4* yrow = smem + 2 * (active ? row : 0) * K;
  constexpr int pstride = (2 * K) | 1;
  uint4* otile = smem + 2 * ROWS * K;
  uint4* mine = otile + row * pstride;
  uint4* grow = a.out + 2ull * (row0 + row) * K;
  // not unrolled: one output's accumulator live at a time (72 registers), a third of the code
#pragma unroll 1
  for (int i0 = 0; i0 < K; i0 += SPLIT) {
    const int i = i0 + part;
    if (i < K) {
      Fe r;
      if (ARITH == 1) {
        uint64_t col[17];
#pragma unroll
        for (int c = 0; c < 17; c++) col[c] = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
          uint32_t y[9], m[9];
          to_limbs29(lds_fe_reload(yrow + 2 * j), y);
#pragma unroll
          for (int q = 0; q < 9; q++) m[q] = a.mc[i][j][q];
          mac29(col, y, m);
        }
        if (K > 6) norm29(col);  // 7 or 8 terms: make room for the reduction products
        r = redc29<F>(col);
      } else {
        Acc acc;
        acc_zero(acc);
#pragma unroll
        for (int j = 0; j < K; j++) {
          Fe m;
#pragma unroll
          for (int q = 0; q < 8; q++) m.w[q] = a.mc[i][j][q];
          acc_mac<true>(acc, lds_fe_reload(yrow + 2 * j), m);
          if ((j + 1) % F::kFold == 0 || j == K - 1) acc_fold<F>(acc);
        }
        r = acc_redc<F>(acc);
      }
      if (GATHER)
        st_fe(mine + 2 * i, r);
      else
        st_fe_global256(grow + 2 * i, r);
    }
  }
  if (GATHER) {
    __syncthreads();
    tile_store_rows_gather<THREADS>(otile, a.gather, 2ull * (a.gather_row0 + row0) * K, rows_here, 2 * K);
  }
}

// ---------------------------------------------------------------------------
// Radix-2 NTT in shared memory (fft / partial_fft / fft_batch_evaluate,
// rsdecode_impl.h:125-192).  n <= 1024.  A CTA of 256 threads transforms
// max(1, 512/n) polynomials at a time; element planes are split in two uint4
// halves so consecutive lanes touch consecutive 16-byte slots.
// DIT: inputs are zero-padded to n and scattered to bit-reversed slots, the
// first stage (all twiddles 1) skips the multiply.
// ---------------------------------------------------------------------------
struct NttArgs {
  const uint4* in;     // [batch][d]
  uint4* out;          // [batch][k_out]
  const uint4* tw;     // [n/2] omega^i, Montgomery form
  unsigned long long batch;
  int n, log_n, d, k_out;
};

HB_D Fe lds_fe(const uint4* lo, const uint4* hi, int e) {
  uint4 a = lo[e], b = hi[e];
  Fe r;
  r.w[0] = a.x; r.w[1] = a.y; r.w[2] = a.z; r.w[3] = a.w;
  r.w[4] = b.x; r.w[5] = b.y; r.w[6] = b.z; r.w[7] = b.w;
  return r;
}

HB_D void sts_fe(uint4* lo, uint4* hi, int e, const Fe& r) {
  lo[e] = make_uint4(r.w[0], r.w[1], r.w[2], r.w[3]);
  hi[e] = make_uint4(r.w[4], r.w[5], r.w[6], r.w[7]);
}

template <class F>
__global__ void __launch_bounds__(256) ntt_smem_kernel(NttArgs a) {
  extern __shared__ uint4 smem[];
  const int n = a.n, log_n = a.log_n;
  const int per_cta = n >= 512 ? 1 : 512 / n;
  const int elems = per_cta * n;
  uint4* lo = smem;
  uint4* hi = smem + elems;
  const unsigned long long first = (unsigned long long)blockIdx.x * per_cta;
  const int d = a.d < n ? a.d : n;  // coefficients beyond n are dropped (rsdecode_impl.h:173-175)

  // load + bit-reverse scatter (one uint4 half per step: fully coalesced)
  for (int q = threadIdx.x; q < 2 * elems; q += 256) {
    int e = q >> 1, half = q & 1;
    int poly = e >> log_n, j = e & (n - 1);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j < d && first + poly < a.batch) v = a.in[2ull * ((first + poly) * a.d + j) + half];
    int rj = (int)(__brev((unsigned)j) >> (32 - log_n));
    (half ? hi : lo)[poly * n + rj] = v;
  }
  __syncthreads();

  const int bflies = elems >> 1;
  for (int s = 1; s <= log_n; s++) {
    const int h = 1 << (s - 1);
    const int tw_stride = n >> s;
    for (int t = threadIdx.x; t < bflies; t += 256) {
      int poly = t >> (log_n - 1), bt = t & ((n >> 1) - 1);
      int pos = bt & (h - 1);
      int i0 = poly * n + ((bt >> (s - 1)) << s) + pos;
      int i1 = i0 + h;
      Fe u = lds_fe(lo, hi, i0);
      Fe v = lds_fe(lo, hi, i1);
      if (s > 1) {
        Fe w = ld_fe(a.tw + 2 * (pos * tw_stride));
        v = mont_mul<F>(v, w);
      }
      sts_fe(lo, hi, i0, fe_add<F>(u, v));
      sts_fe(lo, hi, i1, fe_sub<F>(u, v));
    }
    __syncthreads();
  }

  const int k_out = a.k_out;
  for (int q = threadIdx.x; q < 2 * elems; q += 256) {
    int e = q >> 1, half = q & 1;
    int poly = e >> log_n, i = e & (n - 1);
    if (i < k_out && first + poly < a.batch)
      a.out[2ull * ((first + poly) * k_out + i) + half] = (half ? hi : lo)[e];
  }
}

// ---------------------------------------------------------------------------
// Large transforms (n > 1024): one global-memory pass per stage.
// ---------------------------------------------------------------------------
struct NttBigArgs {
  uint4* work;        // [batch][n]
  const uint4* in;    // [batch][d]
  uint4* out;         // [batch][k_out]
  const uint4* tw;    // [n/2]
  unsigned long long batch;
  int n, log_n, d, k_out, stage;
};

__global__ void __launch_bounds__(256) ntt_big_scatter_kernel(NttBigArgs a) {
  unsigned long long g = (unsigned long long)blockIdx.x * 256ull + threadIdx.x;
  unsigned long long total = a.batch * (unsigned long long)a.n;
  if (g >= total) return;
  unsigned long long b = g >> a.log_n;
  int j = (int)(g & (unsigned)(a.n - 1));
  int d = a.d < a.n ? a.d : a.n;
  uint4 z = make_uint4(0, 0, 0, 0), v0 = z, v1 = z;
  if (j < d) {
    v0 = a.in[2ull * (b * a.d + j)];
    v1 = a.in[2ull * (b * a.d + j) + 1];
  }
  int rj = (int)(__brev((unsigned)j) >> (32 - a.log_n));
  uint4* dst = a.work + 2ull * (b * a.n + rj);
  dst[0] = v0;
  dst[1] = v1;
}

template <class F>
__global__ void __launch_bounds__(256) ntt_big_stage_kernel(NttBigArgs a) {
  unsigned long long g = (unsigned long long)blockIdx.x * 256ull + threadIdx.x;
  unsigned long long total = a.batch * (unsigned long long)(a.n >> 1);
  if (g >= total) return;
  const int s = a.stage, h = 1 << (s - 1);
  unsigned long long b = g >> (a.log_n - 1);
  int bt = (int)(g & (unsigned)((a.n >> 1) - 1));
  int pos = bt & (h - 1);
  unsigned long long i0 = b * a.n + (((unsigned)bt >> (s - 1)) << s) + pos;
  unsigned long long i1 = i0 + h;
  Fe u = ld_fe(a.work + 2 * i0);
  Fe v = ld_fe(a.work + 2 * i1);
  if (s > 1) {
    Fe w = ld_fe(a.tw + 2ull * ((unsigned long long)pos * (a.n >> s)));
    v = mont_mul<F>(v, w);
  }
  st_fe(a.work + 2 * i0, fe_add<F>(u, v));
  st_fe(a.work + 2 * i1, fe_sub<F>(u, v));
}

__global__ void __launch_bounds__(256) ntt_big_gather_kernel(NttBigArgs a) {
  unsigned long long g = (unsigned long long)blockIdx.x * 256ull + threadIdx.x;
  unsigned long long total = a.batch * (unsigned long long)a.k_out;
  if (g >= total) return;
  unsigned long long b = g / (unsigned)a.k_out;
  int i = (int)(g - b * (unsigned)a.k_out);
  const uint4* src = a.work + 2ull * (b * a.n + i);
  uint4* dst = a.out + 2ull * g;
  dst[0] = src[0];
  dst[1] = src[1];
}

}  // namespace hb
