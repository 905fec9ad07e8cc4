// Constant-matrix products on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   out[b][i] = sum_j M[i][j] * in[b][j]  (mod p),   M fixed for the whole batch
//
// is what every non-robust primitive of the reconstruction path computes
// (vandermonde_batch_evaluate / _interpolate, fft_batch_interpolate and the matrix
// form of fft_batch_evaluate: rsdecode_impl.h:23-36, :97-122, :125-192, pyx:183,237).
// The IMAD kernels (kernels.cuh) spend 64 IMAD.WIDE per 256-bit product and are bound
// by the integer multiplier.  Here the products run as an EXACT unsigned 8-bit integer
// GEMM with 32-bit accumulation:
//
//   * a field element is 32 bytes; row b of the batch is already a K = 32*d byte
//     vector A[b][(j,a)] = byte a of in[b][j]           -- the HBM layout as it is;
//   * the constant operand absorbs the modular reduction:
//         B[(i,c)][(j,a)] = byte c of ( M[i][j] * 2^(8a) mod p ),
//     because in[b][j] = sum_a A[b][(j,a)] 2^(8a), so
//         V[b][i] := sum_c 2^(8c) * ( sum_(j,a) A[b][(j,a)] * B[(i,c)][(j,a)] )
//                 == out[b][i]  (mod p),     V < 2^280 for d <= 1032;
//   * D = A * B^T is one tcgen05.mma.kind::i8 chain per 128-row tile (u8 x u8 -> s32 in
//     TMEM); the epilogue reads each output's 32 column sums with tcgen05.ld, carries
//     them into a 288-bit integer and reduces it with one small-quotient Barrett step
//     (q < 2^26) and one conditional subtraction (tc_fold_reduce).  No Montgomery form
//     anywhere: inputs, constants and outputs are plain residues.
//
// Warp roles of the persistent CTA (one per SM): EW = 8, 12 or 16 epilogue warps (TMEM
// lane quarter = warp % 4, the outputs of a block dealt round-robin over warp / 4), one
// thread streams the 128-row A tiles into a ring of shared-memory stages with TMA tensor
// copies (128-byte-wide boxes, SWIZZLE_128B = the UMMA K-major swizzled layout; rows past the
// batch and bytes past K are zero-filled by the TMA unit; measured: 16-byte cp.async copies
// from 128 threads could not keep more than ~2 TB/s of loads in flight), one warp owns
// TMEM and issues the MMAs (warp-uniform loops, one elected lane issues).  The constant operand
// is fetched once per CTA -- one TMA bulk copy per 8-output block, queued behind the first input
// tile -- and stays resident in shared memory; accumulators ping-pong between two 256-column
// TMEM buffers, and an epilogue warp hands a buffer back as soon as its TMEM reads have landed,
// before the arithmetic, so the MMAs of one block overlap the epilogue of the previous.
// What bounds it (tools/tmem_probe.cu, tools/tc_probe.cu): the ~100-instruction
// epilogue per output and, for large batches, the write side of HBM; not the MMAs, not TMEM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fp256.cuh"

namespace hb {

struct TcArgs {
  const uint8_t* in;         // rows of K bytes, pitch in_pitch bytes (multiple of 16)
  const uint8_t* bmat;       // n_blocks blocks of (32*ob) x K bytes, UMMA canonical K-major layout
  uint8_t* out;              // output o of row r at r*out_pitch + o*32
  unsigned long long batch;
  unsigned K;                // 32 * d
  unsigned n_out;            // outputs per row
  unsigned ob;               // outputs per accumulator block (1..8)
  unsigned n_blocks;         // ceil(n_out / ob)
  unsigned in_pitch, out_pitch;
  unsigned stages;           // A stages in shared memory (>= 2); one stage = ceil(K/128) boxes of 16 KB
  // split != 0: the matrix is a DFT (out[i] = sum_j in[j] w^(ij), i < n_out <= 2*half) evaluated as
  // one radix-2 step on top of two half-size products (the decomposition of rsdecode_impl.h:138-157
  // by input parity): E[i] = sum_m in[2m] w^(2mi), O[i] = sum_m in[2m+1] w^((2m+1)i), i < half, and
  // out[i] = E[i] + O[i], out[i + half] = E[i] - O[i].  A block then holds `ob` PAIRS: columns
  // [0, 32 ob) accumulate E over the even K steps, [32 ob, 64 ob) accumulate O over the odd ones
  // -- half the multiply-accumulates and half the constant operand of the plain form.
  unsigned split;
  unsigned half;             // n / 2 of the DFT (split mode)
  unsigned mu;               // floor(2^280 / p)
  // staged != 0 (opt-in, needs EW = 16): results go through shared memory (two buffers of four
  // 32-row x 128-byte boxes, XOR-swizzled) so that every global store instruction writes full
  // 128-byte lines instead of one 32-byte sector per thread.  Measured SLOWER than the per-thread
  // stores on the cfg2 shapes (the extra shared-memory passes and 128-thread barriers cost more
  // than the store path saves); kept as an option, the tests force both.
  unsigned staged;
  // fused all-gather (hbg_fft_batch_interpolate_allgather): when gather_world > 0 the result of
  // row r is stored at row gather_row0 + r of EVERY rank's buffer -- one multimem.st per 16
  // bytes through the NVSwitch multicast address gather_mc, or gather_world peer stores
  uint8_t* gather_peers[8];
  uint8_t* gather_mc;
  unsigned long long gather_row0;
  unsigned gather_world;
  unsigned hyp;              // probe only: descriptor field hypotheses
  uint32_t* debug;           // probe only: raw column sums of the first tile, or null
  long long* trace;          // probe only: clock64 stamps of CTA 0, [role][tile < 64][event < 8]
  unsigned* error;           // set to non-zero when a barrier wait times out
};

constexpr int kTcMaxStages = 6;
constexpr int kTcBBars = 4;      // the resident operand arrives in up to 4 pieces (block 0, 1, 2, the rest)
constexpr int kTcLoadWarps = 1;  // one elected thread issues the TMA tensor copies

HB_D unsigned tc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

HB_D void tc_mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}

HB_D void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

// Bounded wait: a protocol bug must end in an error code, not in a hung GPU.
HB_D bool tc_mbar_try(unsigned a, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(a), "r"(parity)
      : "memory");
  return done != 0;
}

HB_D bool tc_mbar_wait(uint64_t* bar, unsigned parity, unsigned* error) {
  const unsigned a = tc_smem_u32(bar);
  if (tc_mbar_try(a, parity)) return true;  // the common case costs no clock read
  const long long t0 = clock64();
  for (;;) {
    if (tc_mbar_try(a, parity)) return true;
    if (clock64() - t0 > 4000000000ll) {  // ~2 s
      if (error) atomicExch(error, 1u);
      __trap();
    }
  }
}

HB_D uint64_t tc_smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
  // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46),
  // version 1 [46,48), layout type 0 = no swizzle [61,64)
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}

// K-major operand in the 128-byte-swizzled layout TMA writes (rows of 128 bytes, 8-row groups
// of 1024 bytes): layout type 2, SBO = 1024, LBO unused (1).  A K step of 32 bytes advances
// the start address inside the swizzle span.
HB_D uint64_t tc_smem_desc_sw128(unsigned addr) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

HB_D void tc_mma_i8(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// One lane of a converged warp (the same one every time in practice).
HB_D bool tc_elect() {
  unsigned pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

HB_D void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   tc_smem_u32(bar))
               : "memory");
}

// Asynchronous: the registers are valid after tc_ld_wait(c) only.
HB_D void tc_ld32_async(unsigned taddr, uint32_t* c) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]),
        "=r"(c[8]), "=r"(c[9]), "=r"(c[10]), "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]),
        "=r"(c[15]), "=r"(c[16]), "=r"(c[17]), "=r"(c[18]), "=r"(c[19]), "=r"(c[20]), "=r"(c[21]),
        "=r"(c[22]), "=r"(c[23]), "=r"(c[24]), "=r"(c[25]), "=r"(c[26]), "=r"(c[27]), "=r"(c[28]),
        "=r"(c[29]), "=r"(c[30]), "=r"(c[31])
      : "r"(taddr)
      : "memory");
}

// Waits for every tcgen05.ld of this thread; the registers are operands so that no use of
// them can be scheduled above the wait and no other value can live in them meanwhile.
HB_D void tc_ld_wait(uint32_t* c) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7]),
                 "+r"(c[8]), "+r"(c[9]), "+r"(c[10]), "+r"(c[11]), "+r"(c[12]), "+r"(c[13]), "+r"(c[14]),
                 "+r"(c[15]), "+r"(c[16]), "+r"(c[17]), "+r"(c[18]), "+r"(c[19]), "+r"(c[20]), "+r"(c[21]),
                 "+r"(c[22]), "+r"(c[23]), "+r"(c[24]), "+r"(c[25]), "+r"(c[26]), "+r"(c[27]), "+r"(c[28]),
                 "+r"(c[29]), "+r"(c[30]), "+r"(c[31])
               :
               : "memory");
}

HB_D void tc_ld32(unsigned taddr, uint32_t* c) {
  tc_ld32_async(taddr, c);
  tc_ld_wait(c);
}

// 32 column sums (weights 2^(8c), each < 2^31) -> the canonical residue of
// V = sum_c col[c] 2^(8c) modulo p, for V < 2^279.
//   1. eight independent 64-bit word sums s_i = c[4i] + c[4i+1] 2^8 + c[4i+2] 2^16 + c[4i+3] 2^24
//      (three IMAD.WIDE each, no dependency between words), one 32-bit carry chain -> w[0..8];
//   2. q = floor(floor(V / 2^248) * mu / 2^32) with mu = floor(2^280 / p): q <= V/p and
//      V/p - q < 1 + V/2^280 + 2^248/p < 2, so R = V - q p lies in [0, 2p) and fits 256 bits;
//   3. R is computed modulo 2^256 (eight independent q*p_i products, two carry chains) and
//      reduced with ONE conditional subtraction of p.
//
// NARROW = true (K <= 256: every column sum is < 255^2 * 256 < 2^24 - 2^16): a pair of columns
// c[2i] + 2^8 c[2i+1] fits 32 bits, so V = X + 2^16 Y with X, Y the 256-bit numbers whose words
// ARE the even / odd pair sums -- 16 32-bit multiply-adds, 9 funnel shifts and one carry chain
// instead of 24 IMAD.WIDE (the epilogue is bound by the multiplier pipe: 4 cycles per IMAD.WIDE).
template <class F, bool NARROW>
HB_D void tc_fold_reduce(const uint32_t* c, uint32_t mu, Fe& r) {
  uint32_t w[9];
  if constexpr (NARROW) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x[i] = c[4 * i + 1] * 256u + c[4 * i];
      y[i] = c[4 * i + 3] * 256u + c[4 * i + 2];
    }
    uint32_t ys[9];
    ys[0] = y[0] << 16;
#pragma unroll
    for (int i = 1; i < 8; i++) ys[i] = __funnelshift_l(y[i - 1], y[i], 16);
    ys[8] = y[7] >> 16;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, %25, 0;"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
          "=r"(w[8])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
          "r"(ys[0]), "r"(ys[1]), "r"(ys[2]), "r"(ys[3]), "r"(ys[4]), "r"(ys[5]), "r"(ys[6]), "r"(ys[7]),
          "r"(ys[8]));
  } else {
  uint32_t lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    unsigned long long s = (unsigned long long)c[4 * i + 1] * 256u + c[4 * i];
    s += (unsigned long long)c[4 * i + 2] * 65536u;
    s += (unsigned long long)c[4 * i + 3] * 16777216u;
    lo[i] = (uint32_t)s;
    hi[i] = (uint32_t)(s >> 32);
  }
  w[0] = lo[0];
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, 0;"
      : "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8])
      : "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]), "r"(hi[7]),
        "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]));
  }
  const uint32_t top = __funnelshift_l(w[7], w[8], 8);  // floor(V / 2^248)
  const uint32_t q = __umulhi(top, mu);
  // q*p modulo 2^256: even-limb products on aligned word pairs, odd-limb products one word up
  uint32_t e[8], o[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned long long pe = (unsigned long long)q * F::p(2 * j);
    e[2 * j] = (uint32_t)pe;
    e[2 * j + 1] = (uint32_t)(pe >> 32);
    const unsigned long long po = (unsigned long long)q * F::p(2 * j + 1);
    o[2 * j] = (uint32_t)po;
    o[2 * j + 1] = (uint32_t)(po >> 32);
  }
  // R = w - e - (o << 32)   (mod 2^256)
  uint32_t t[8];
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
      : "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
        "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]));
  asm("sub.cc.u32 %0, %0, %7;\n\t"
      "subc.cc.u32 %1, %1, %8;\n\t"
      "subc.cc.u32 %2, %2, %9;\n\t"
      "subc.cc.u32 %3, %3, %10;\n\t"
      "subc.cc.u32 %4, %4, %11;\n\t"
      "subc.cc.u32 %5, %5, %12;\n\t"
      "subc.u32 %6, %6, %13;"
      : "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7])
      : "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]));
  // one conditional subtraction of p
  uint32_t u[8], borrow;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(borrow)
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(F::p(0)), "r"(F::p(1)), "r"(F::p(2)), "r"(F::p(3)), "r"(F::p(4)), "r"(F::p(5)), "r"(F::p(6)),
        "r"(F::p(7)));
#pragma unroll
  for (int i = 0; i < 8; i++) r.w[i] = borrow ? t[i] : u[i];
}

// one 256-bit store (two 128-bit stores: 21.7 instead of 18.1 us per cfg2 step; .cs: no difference)
HB_D void tc_st256(uint8_t* p, const Fe& r) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]),
               "r"(r.w[2]), "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7])
               : "memory");
}

HB_D void tc_st_multimem(uint8_t* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("multimem.st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}

#define TC_TRACE(role, it_, ev)                                                              \
  do {                                                                                       \
    if (PROBE && a.trace && blockIdx.x == 0 && (it_) < 64) a.trace[((role)*64 + (it_)) * 8 + (ev)] = clock64(); \
  } while (0)

// STREAM = false: the constant operand is resident in shared memory (a.bmat, canonical layout).
// STREAM = true:  it is streamed like the input -- a.bmat is a plain row-major u8 matrix
//   [n_blocks * 32 ob rows][ceil(K/128) * 128 bytes] described by tmap_b; a stage holds one
//   128-byte K chunk of the input tile (16 KB) and of the current block of the operand
//   (32 ob x 128 bytes), both 128-byte swizzled, and the K loop runs over stages.  For matrices
//   whose operand does not fit shared memory (cfg4: 16 x 16, cfg5: 43 x 43, Gao's m x m step).
// PROBE = true (tools/tc_probe only) compiles the hypothesis flags, the accumulator dump and the
// role trace in; the library's instantiations carry none of them.
template <class F, int EW, bool STREAM, bool PROBE = false>
__global__ void __launch_bounds__((EW + kTcLoadWarps + 1) * 32, 1)
tc_apply_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_b, TcArgs a) {
  constexpr int kTcEpiWarps = EW;
  const unsigned hyp = PROBE ? a.hyp : 0u;
  uint32_t* const debug = PROBE ? a.debug : nullptr;
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ uint64_t bar_full[kTcMaxStages], bar_empty[kTcMaxStages], bar_tfull[2], bar_tempty[2], bar_b[kTcBBars];
  __shared__ unsigned tmem_base_slot;

  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned NB = 32 * a.ob;            // accumulator columns per block (per part in split mode)
  const unsigned KBOX = (a.K + 127) >> 7;   // 128-byte-wide TMA boxes per tile
  const unsigned KS = a.K >> 5;             // MMA K steps (32 bytes each)
  const unsigned b_block_bytes = NB * a.K;
  const unsigned b_bytes = STREAM ? 0u : b_block_bytes * a.n_blocks;
  const unsigned b_chunk = NB * 128;        // STREAM: one K chunk of one block of the operand
  const unsigned stage_bytes = STREAM ? 16384u + ((b_chunk + 1023) & ~1023u) : KBOX * 16384;
  uint8_t* smem_b = tc_smem;
  uint8_t* smem_a = tc_smem + ((b_bytes + 1023) & ~1023u);  // swizzle atoms need 1024-byte alignment
  uint8_t* smem_st = smem_a + (size_t)a.stages * stage_bytes;  // store staging (a.staged), 2 x 16 KB
  const unsigned long long tiles = (a.batch + 127) >> 7;

  // --- set-up: barriers, TMEM, the resident constant operand -------------------
  if (threadIdx.x == 0) {
    for (unsigned s = 0; s < a.stages; s++) {
      tc_mbar_init(&bar_full[s], 1);
      tc_mbar_init(&bar_empty[s], 1);
    }
    for (int i = 0; i < 2; i++) {
      tc_mbar_init(&bar_tfull[i], 1);
      tc_mbar_init(&bar_tempty[i], kTcEpiWarps);
    }
    for (int i = 0; i < kTcBBars; i++) tc_mbar_init(&bar_b[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == kTcEpiWarps + kTcLoadWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     tc_smem_u32(&tmem_base_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = tmem_base_slot;

  if (warp == kTcEpiWarps) {
    // ===== loader: one thread, ceil(K/128) TMA box copies per tile, `stages` tiles in flight =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      unsigned it = 0;
      // The resident constant operand: one TMA bulk copy per piece (block 0, 1, 2, the rest), each on
      // its own barrier, queued BEHIND the first input tile -- the MMAs of block 0 start once that
      // tile and a 1 / n_blocks share of the operand are in, not after all of it (98 KB per CTA for
      // the cfg2 encode: most of the kernel's fill time).
      auto issue_b = [&]() {
        for (unsigned g = 0; g < (unsigned)kTcBBars; g++) {
          const unsigned first = g, count = g + 1 < (unsigned)kTcBBars ? 1u : (a.n_blocks > g ? a.n_blocks - g : 0u);
          const unsigned bytes = first < a.n_blocks ? b_block_bytes * count : 0u;
          const unsigned bar = tc_smem_u32(&bar_b[g]);
          if (bytes == 0) {
            tc_mbar_arrive(&bar_b[g]);
            continue;
          }
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  tc_smem_u32(smem_b + (size_t)first * b_block_bytes)),
              "l"(a.bmat + (size_t)first * b_block_bytes), "r"(bytes), "r"(bar)
              : "memory");
        }
      };
      if constexpr (STREAM) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
          const int row0 = (int)(tile << 7);
          for (unsigned nb = 0; nb < a.n_blocks; nb++)
            for (unsigned kb = 0; kb < KBOX; kb++, it++) {
              const unsigned s = it % a.stages, ph = (it / a.stages) & 1;
              tc_mbar_wait(&bar_empty[s], ph ^ 1, a.error);
              const unsigned bar = tc_smem_u32(&bar_full[s]);
              asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                           "r"(16384u + b_chunk)
                           : "memory");
              const unsigned dst = tc_smem_u32(smem_a + (size_t)s * stage_bytes);
              asm volatile(
                  "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                  "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                  "l"(&tmap), "r"((int)(kb * 128)), "r"(row0), "r"(bar)
                  : "memory");
              asm volatile(
                  "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                  "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst + 16384),
                  "l"(&tmap_b), "r"((int)(kb * 128)), "r"((int)(nb * NB)), "r"(bar)
                  : "memory");
            }
        }
      } else {
      for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
        const unsigned s = it % a.stages, ph = (it / a.stages) & 1;
        tc_mbar_wait(&bar_empty[s], ph ^ 1, a.error);
        TC_TRACE(0, it, 0);
        const unsigned bar = tc_smem_u32(&bar_full[s]);
        if (hyp & 8) {  // probe: no loads
          if (it == 0) issue_b();
          tc_mbar_arrive(&bar_full[s]);
          continue;
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes)
                     : "memory");
        const unsigned dst = tc_smem_u32(smem_a + (size_t)s * stage_bytes);
        const int row0 = (int)(tile << 7);
        for (unsigned kb = 0; kb < KBOX; kb++)
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
              "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst + kb * 16384),
              "l"(&tmap), "r"((int)(kb * 128)), "r"(row0), "r"(bar)
              : "memory");
        if (it == 0) issue_b();
        TC_TRACE(0, it, 1);
      }
      }
    }
  } else if (warp == kTcEpiWarps + kTcLoadWarps) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loops so that the descriptors live
    // in uniform registers; one elected lane issues the MMAs and the commits.  (With the loops
    // under `lane == 0` every MMA cost ~200 cycles of register -> uniform-register moves.) =====
    {
      // cute::UMMA::InstrDescriptor: c_format S32 (2) [4,6), a/b format UINT8 (0), K-major both,
      // N >> 3 at [17,23), M >> 4 at [24,29)
      const unsigned idesc = (2u << 4) | ((NB >> 3) << 17) | ((128u >> 4) << 24);
      const unsigned b_lbo = NB * 16, b_sbo = 128u;
      unsigned it = 0, acc_it = 0;
      if constexpr (STREAM) {
        for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
          for (unsigned nb = 0; nb < a.n_blocks; nb++, acc_it++) {
            const unsigned buf = acc_it & 1, aph = (acc_it >> 1) & 1;
            tc_mbar_wait(&bar_tempty[buf], aph ^ 1, a.error);
            for (unsigned kb = 0; kb < KBOX; kb++, it++) {
              const unsigned s = it % a.stages, ph = (it / a.stages) & 1;
              tc_mbar_wait(&bar_full[s], ph, a.error);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const unsigned a_base = tc_smem_u32(smem_a + (size_t)s * stage_bytes);
              const unsigned steps = KS - 4 * kb < 4 ? KS - 4 * kb : 4;
              const uint64_t ad0 = tc_smem_desc_sw128(a_base), bd0 = tc_smem_desc_sw128(a_base + 16384);
              if (tc_elect()) {
                for (unsigned ks = 0; ks < steps; ks++)
                  tc_mma_i8(tmem_base + buf * 256, ad0 + 2 * ks, bd0 + 2 * ks, idesc, (kb | ks) > 0);
                tc_commit(&bar_empty[s]);
              }
              __syncwarp();
            }
            if (tc_elect()) tc_commit(&bar_tfull[buf]);
            __syncwarp();
          }
        }
      } else {
      const unsigned b_step = (2 * (NB * 16)) >> 4;  // descriptor units (16 bytes) per K step of the operand
      for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
        const unsigned s = it % a.stages, ph = (it / a.stages) & 1;
        tc_mbar_wait(&bar_full[s], ph, a.error);
        if (lane == 0) TC_TRACE(1, it, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned a_base = tc_smem_u32(smem_a + (size_t)s * stage_bytes);
        const uint64_t ad0 = tc_smem_desc_sw128(a_base);
        for (unsigned nb = 0; nb < a.n_blocks; nb++, acc_it++) {
          const unsigned buf = acc_it & 1, aph = (acc_it >> 1) & 1;
          tc_mbar_wait(&bar_tempty[buf], aph ^ 1, a.error);
          if (lane == 0 && nb < 2) TC_TRACE(1, it, 1 + 2 * nb);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (it == 0) tc_mbar_wait(&bar_b[nb < (unsigned)kTcBBars ? nb : kTcBBars - 1], 0, a.error);
          const unsigned b_base = tc_smem_u32(smem_b + (size_t)nb * b_block_bytes);
          const uint64_t bd0 = tc_smem_desc(b_base, b_lbo, b_sbo);
          const unsigned d0 = tmem_base + buf * 256;
          if (tc_elect()) {
            if (!a.split) {
              for (unsigned ks = 0; ks < KS && !(hyp & 16); ks++) {
                // a K step advances 32 bytes inside the 128-byte swizzle span, 16 KB per box
                const uint64_t ad = ad0 + (ks >> 2) * (16384u >> 4) + (ks & 3) * 2;
                tc_mma_i8(d0, ad, bd0 + (uint64_t)ks * b_step, idesc, ks > 0);
              }
            } else {
              // part 0 (E): even elements = even K steps; part 1 (O): odd ones.  The block's
              // constant operand stores the even steps first, then the odd steps.
              const unsigned n_even = (KS + 1) >> 1;
              for (unsigned part = 0; part < 2; part++) {
                unsigned bstep = part ? n_even : 0;
                for (unsigned ks = part; ks < KS; ks += 2, bstep++) {
                  const uint64_t ad = ad0 + (ks >> 2) * (16384u >> 4) + (ks & 3) * 2;
                  tc_mma_i8(d0 + part * NB, ad, bd0 + (uint64_t)bstep * b_step, idesc, ks > part);
                }
              }
            }
            tc_commit(&bar_tfull[buf]);
          }
          __syncwarp();
          if (lane == 0 && nb < 2) TC_TRACE(1, it, 4 + nb);
        }
        if (tc_elect()) tc_commit(&bar_empty[s]);
        __syncwarp();
        if (lane == 0) TC_TRACE(1, it, 2);
      }
      }
    }
  } else {
    // ===== epilogue: thread = row (TMEM lane); warp / 4 deals out the outputs of a block =====
    const unsigned quarter = warp & 3, group = warp >> 2;
    constexpr unsigned G = kTcEpiWarps / 4;
    if (!a.split && !(hyp & 6) && !debug) {
      // Per block: read the warp's outputs (up to two at a time, back to back) from TMEM, hand the
      // accumulator back to the MMA warp as soon as the reads have landed, THEN fold, reduce and
      // store.  Measured (tools/tmem_probe, tools/tc_probe traces): a tcgen05.ld.x32 costs ~21
      // cycles per SM sub-partition, the fold + reduce ~110; with the release after the arithmetic
      // the two accumulator buffers each ran a strictly serial MMA -> epilogue -> MMA chain and
      // every hand-over (~400 cycles) was exposed.
      const unsigned n_o = group < a.ob ? (a.ob - group + G - 1) / G : 0;
      const bool narrow = a.K <= 256;  // column sums < 2^24 - 2^16 (see tc_fold_reduce)
      const unsigned tq = tmem_base + ((quarter * 32) << 16);
      auto finish = [&](const uint32_t* c, unsigned long long row, unsigned out_idx) {
        if (row < a.batch && out_idx < a.n_out) {
          Fe r;
          if (hyp & 128) {  // probe: no arithmetic
#pragma unroll
            for (int i = 0; i < 8; i++) r.w[i] = c[i] ^ c[i + 8] ^ c[i + 16] ^ c[i + 24];
          } else if (narrow)
            tc_fold_reduce<F, true>(c, a.mu, r);
          else
            tc_fold_reduce<F, false>(c, a.mu, r);
          if ((hyp & 64) && !(r.w[0] == 0xdeadbeefu && r.w[7] == 0x12345u)) return;  // probe: no store
          if (a.gather_world == 0) {
            tc_st256(a.out + row * a.out_pitch + out_idx * 32, r);
          } else {
            const unsigned long long off = (a.gather_row0 + row) * a.out_pitch + out_idx * 32;
            if (a.gather_mc) {
              tc_st_multimem(a.gather_mc + off, r.w[0], r.w[1], r.w[2], r.w[3]);
              tc_st_multimem(a.gather_mc + off + 16, r.w[4], r.w[5], r.w[6], r.w[7]);
            } else {
#pragma unroll 1
              for (unsigned w = 0; w < a.gather_world; w++) tc_st256(a.gather_peers[w] + off, r);
            }
          }
        }
      };
      if (kTcEpiWarps == 16 && a.staged) {
        // G = 4: wave j of a block = its outputs 4j .. 4j+3 = 128 contiguous bytes of every row.  The
        // four warps of a lane quarter fill a box of 32 rows x 128 bytes (thread = row, 16-byte chunk
        // c of row r at r*128 + ((c ^ (r & 7)) << 4): conflict-free), meet at a 128-thread barrier,
        // and each warp writes 8 of the rows back to global memory, 4 full lines per instruction.
        // Two buffers: a thread's reads of wave w are complete (its stores consumed them) before it
        // reaches the barrier of wave w+1, after which wave w+2 may overwrite the buffer.
        const unsigned n_waves = (a.ob + 3) >> 2;
        const unsigned st_w0 = lane * 128 + (((2 * group) ^ (lane & 7)) << 4);      // my result, low half
        const unsigned st_w1 = lane * 128 + (((2 * group + 1) ^ (lane & 7)) << 4);  // high half
        const unsigned rd_chunk = lane & 7, rd_row0 = group * 8 + (lane >> 3);      // rows rd_row0, rd_row0 + 4
        const unsigned st_r0 = rd_row0 * 128 + ((rd_chunk ^ (rd_row0 & 7)) << 4);
        const unsigned st_r1 = (rd_row0 + 4) * 128 + ((rd_chunk ^ ((rd_row0 + 4) & 7)) << 4);
        // 16 bytes to the output: the local buffer, or (fused all-gather) every rank's buffer -- one
        // multimem.st through the NVSwitch multicast address, or one store per peer; either way a
        // warp-wide instruction covers 4 full 128-byte lines
        auto put16 = [&](size_t off, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
          if (a.gather_world == 0) {
            asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(a.out + off), "r"(v0), "r"(v1), "r"(v2), "r"(v3)
                         : "memory");
          } else if (a.gather_mc) {
            tc_st_multimem(a.gather_mc + off, v0, v1, v2, v3);
          } else {
#pragma unroll 1
            for (unsigned w = 0; w < a.gather_world; w++)
              asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(a.gather_peers[w] + off), "r"(v0), "r"(v1),
                           "r"(v2), "r"(v3)
                           : "memory");
          }
        };
        const unsigned stg_base = tc_smem_u32(smem_st) + quarter * 4096, bar_id = 1 + quarter;
        const size_t pitch4 = 4 * (size_t)a.out_pitch;
        const bool has0 = group < a.ob, has1 = group + 4 < a.ob;
        unsigned acc = 0, wave = 0, it = 0;
        for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
          const unsigned long long grow = (tile << 7) + quarter * 32 + rd_row0;  // global row of my first read
          const bool row_ok0 = grow < a.batch, row_ok1 = grow + 4 < a.batch;
          const size_t ooff = (a.gather_row0 + grow) * a.out_pitch + (rd_chunk & 1) * 16;  // byte offset of my chunk
          for (unsigned nb = 0; nb < a.n_blocks; nb++, acc++) {
            const unsigned buf = acc & 1;
            tc_mbar_wait(&bar_tfull[buf], (acc >> 1) & 1, a.error);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (threadIdx.x == 0 && nb < 2) TC_TRACE(2, it, 3 * nb + 0);
            uint32_t c0[32], c1[32];
            if (has0) tc_ld32_async(tq + buf * 256 + group * 32, c0);
            if (has1) tc_ld32_async(tq + buf * 256 + (group + 4) * 32, c1);
            if (has0) tc_ld_wait(c0);
            if (has1) tc_ld_wait(c1);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&bar_tempty[buf]);
            if (threadIdx.x == 0 && nb < 2) TC_TRACE(2, it, 3 * nb + 2);
            // (folding both outputs first, to give the scheduler two independent carry-chain streams,
            // was measured 50 % SLOWER: the two 32-register inputs plus both results exceed the 96
            // registers a 576-thread CTA gets and the spills land on the critical path)
            auto wave_body = [&](const uint32_t* c, bool has, unsigned j) {
              const unsigned stg = stg_base + (wave & 1) * 16384;
              if (has) {
                Fe r;
                if (PROBE && (hyp & 128)) {  // probe: no arithmetic
#pragma unroll
                  for (int i = 0; i < 8; i++) r.w[i] = c[i] ^ c[i + 8] ^ c[i + 16] ^ c[i + 24];
                } else if (narrow)
                  tc_fold_reduce<F, true>(c, a.mu, r);
                else
                  tc_fold_reduce<F, false>(c, a.mu, r);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + st_w0), "r"(r.w[0]), "r"(r.w[1]),
                             "r"(r.w[2]), "r"(r.w[3])
                             : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stg + st_w1), "r"(r.w[4]), "r"(r.w[5]),
                             "r"(r.w[6]), "r"(r.w[7])
                             : "memory");
              }
              asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
              // chunk rd_chunk of the box = half of output 4j + rd_chunk / 2 of the block
              const unsigned o = 4 * j + (rd_chunk >> 1), out_idx = nb * a.ob + o;
              if (o < a.ob && out_idx < a.n_out && !(PROBE && (hyp & 64))) {
                const size_t off = ooff + out_idx * 32;
                uint32_t v0, v1, v2, v3;
                if (row_ok0) {
                  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                               : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                               : "r"(stg + st_r0)
                               : "memory");
                  put16(off, v0, v1, v2, v3);
                }
                if (row_ok1) {
                  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                               : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                               : "r"(stg + st_r1)
                               : "memory");
                  put16(off + pitch4, v0, v1, v2, v3);
                }
              }
              wave++;
            };
            wave_body(c0, has0, 0);
            if (n_waves > 1) wave_body(c1, has1, 1);
            if (threadIdx.x == 0 && nb < 2) TC_TRACE(2, it, 3 * nb + 1);
          }
        }
      } else {
      unsigned it = 0, acc = 0;
      for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
        const unsigned long long row = (tile << 7) + quarter * 32 + lane;
        for (unsigned nb = 0; nb < a.n_blocks; nb++, acc++) {
          const unsigned buf = acc & 1;
          tc_mbar_wait(&bar_tfull[buf], (acc >> 1) & 1, a.error);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (threadIdx.x == 0 && nb < 2) TC_TRACE(2, it, 3 * nb + 0);
          unsigned j = 0;
          do {
            uint32_t c0[32], c1[32];
            const bool two = j + 1 < n_o, last = j + 2 >= n_o;
            if (j < n_o) tc_ld32_async(tq + buf * 256 + (group + j * G) * 32, c0);
            if (two) tc_ld32_async(tq + buf * 256 + (group + (j + 1) * G) * 32, c1);
            if (j < n_o) tc_ld_wait(c0);
            if (two) tc_ld_wait(c1);
            if (last) {
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) tc_mbar_arrive(&bar_tempty[buf]);
              if (threadIdx.x == 0 && nb < 2) TC_TRACE(2, it, 3 * nb + 2);
            }
            if (j < n_o) finish(c0, row, nb * a.ob + group + j * G);
            if (two) finish(c1, row, nb * a.ob + group + (j + 1) * G);
            j += 2;
          } while (j < n_o);
          if (threadIdx.x == 0 && nb < 2) TC_TRACE(2, it, 3 * nb + 1);
        }
      }
      }
    } else {
    unsigned it = 0, acc_it = 0;
    for (unsigned long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
      const unsigned long long row = (tile << 7) + quarter * 32 + lane;
      for (unsigned nb = 0; nb < a.n_blocks; nb++, acc_it++) {
        const unsigned buf = acc_it & 1, aph = (acc_it >> 1) & 1;
        tc_mbar_wait(&bar_tfull[buf], aph, a.error);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (a.split) {
          for (unsigned o = group; o < a.ob; o += G) {
            uint32_t c[32];
            Fe ev, od;
            const unsigned tbase = tmem_base + ((quarter * 32) << 16) + buf * 256 + o * 32;
            tc_ld32(tbase, c);
            tc_fold_reduce<F, false>(c, a.mu, ev);
            tc_ld32(tbase + NB, c);
            tc_fold_reduce<F, false>(c, a.mu, od);
            const unsigned i = nb * a.ob + o;
            if (row < a.batch && i < a.half) {
              uint8_t* orow = a.out + row * a.out_pitch;
              if (i < a.n_out) tc_st256(orow + i * 32, fe_add<F>(ev, od));
              if (i + a.half < a.n_out) tc_st256(orow + (i + a.half) * 32, fe_sub<F>(ev, od));
            }
          }
        } else {
          // probe / debug path
          for (unsigned o = group; o < a.ob; o += G) {
            uint32_t c[32];
            if (hyp & 4) continue;  // probe: no TMEM read, no arithmetic, no store
            tc_ld32(tmem_base + ((quarter * 32) << 16) + buf * 256 + o * 32, c);
            if (hyp & 2) {          // probe: TMEM read only
              if (c[0] == 0xdeadbeefu && c[31] == 0x12345u) a.out[0] = 1;
              continue;
            }
            const unsigned out_idx = nb * a.ob + o;
            if (debug) {
              if (tile == 0 && row < 128)
                for (int i = 0; i < 32; i++) debug[(row * a.n_blocks * NB) + nb * NB + o * 32 + i] = c[i];
            }
            if (row < a.batch && out_idx < a.n_out) {
              Fe r;
              tc_fold_reduce<F, false>(c, a.mu, r);
              tc_st256(a.out + row * a.out_pitch + out_idx * 32, r);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&bar_tempty[buf]);
      }
    }
    }
  }

  // --- teardown -----------------------------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kTcEpiWarps + kTcLoadWarps) {
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// Host: the TMA descriptor of the input rows -- a 2-D u8 tensor (K bytes x batch rows, row pitch
// in_pitch), boxes of 128 bytes x 128 rows written to shared memory with the 128-byte swizzle.
// cuTensorMapEncodeTiled is a pure host routine of the driver; it is looked up through the
// runtime so that the library does not link libcuda.
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);

inline bool tc_make_tmap(CUtensorMap* m, const void* in, unsigned long long batch, unsigned K,
                         unsigned long long in_pitch) {
  static tc_encode_fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (tc_encode_fn)p;
  }();
  if (!fn || ((uintptr_t)in & 15) || (in_pitch & 15)) return false;
  const cuuint64_t dims[2] = {K, batch};
  const cuuint64_t strides[1] = {in_pitch};
  const cuuint32_t box[2] = {128, 128};
  const cuuint32_t es[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(in), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Host: the TMA descriptor of a streamed constant operand: plain row-major u8 [rows][kpad], boxes of
// 128 bytes x box_rows (= 32 ob) rows, 128-byte swizzle.
inline bool tc_make_tmap_b(CUtensorMap* m, const void* bmat, unsigned long long rows, unsigned kpad,
                           unsigned box_rows) {
  CUtensorMap tmp;
  if (!tc_make_tmap(&tmp, bmat, rows, kpad, kpad)) return false;  // resolves the entry point, checks alignment
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  const cuuint64_t dims[2] = {kpad, rows};
  const cuuint64_t strides[1] = {kpad};
  const cuuint32_t box[2] = {128, box_rows};
  const cuuint32_t es[2] = {1, 1};
  return ((tc_encode_fn)p)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(bmat), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr size_t kTcStoreStaging = 32768;  // store staging: two buffers of four 32-row x 128-byte boxes

inline size_t tc_stream_stage_bytes(unsigned ob) { return 16384 + (((size_t)32 * ob * 128 + 1023) & ~(size_t)1023); }

// Host: shared memory the kernel needs for (K, n_blocks, ob, stages) -- plus up to 1008 bytes
// the kernel may skip to align its dynamic shared memory window to 1024.
inline size_t tc_stage_bytes(unsigned K) { return (size_t)((K + 127) / 128) * 16384; }
inline size_t tc_smem_bytes(unsigned K, unsigned n_blocks, unsigned ob, unsigned stages) {
  const size_t b_bytes = (size_t)32 * ob * K * n_blocks;
  return ((b_bytes + 1023) & ~(size_t)1023) + (size_t)stages * tc_stage_bytes(K);
}

}  // namespace hb
