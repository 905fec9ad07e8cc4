// 256-bit prime-field arithmetic for sm_100a (8 x 32-bit limbs, Montgomery
// form with R = 2^256).  Replaces what the reference gets from NTL's ZZ_p
// (honeybadgermpc/ntl/ntlwrapper.pxd:1-42): add/sub/mul/inverse mod p.
//
// Design notes
//  * The multiplier is the "even/odd column" CIOS: products a_j*b_i with j even
//    land on 64-bit aligned column pairs, those with j odd on pairs shifted by
//    one word, so every 32x32->64 multiply-add is one IMAD.WIDE.U32 whose
//    carry rides the predicate carry chain (mad.lo.cc / madc.hi.cc pairs) and
//    the ALU pipe stays almost idle: 16 IMAD.WIDE + 1 IMAD per row, 8 rows.
//  * Data lives in HBM in STANDARD form (canonical residues).  Constants
//    (twiddles, Vandermonde entries) are kept in Montgomery form, so
//    mont_mul(data, constant) is already the standard-form product and no
//    to/from-Montgomery pass over the data is ever needed.
//  * Two field policies: FieldBLS (BLS12-381 scalar field, limbs are
//    immediates) and FieldAny (any odd modulus < 2^256, limbs in __constant__).
//  * Every primitive has a host twin (emulated carry flag) so the exact same
//    algorithm text is unit-tested on the CPU (tests/test_host_field.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#define HB_D __device__ __forceinline__
#else
#define HB_HD inline
#define HB_D inline
#endif

namespace hb {

struct Fe {
  uint32_t w[8];
};

struct FieldParams {
  uint32_t p[8];     // modulus
  uint32_t r2[8];    // R^2 mod p   (to Montgomery: mont_mul(a, r2))
  uint32_t one[8];   // R mod p     (Montgomery form of 1)
  uint32_t n0inv;    // -p^{-1} mod 2^32
  uint32_t pad[7];
  // radix-2^29 view of the same modulus (carry-free lazy dot products, rowmath.cuh)
  uint32_t p29[9];   // p = sum p29[i] 2^(29 i)
  uint32_t n0inv29;  // -p^{-1} mod 2^29
  uint32_t r261[8];  // 2^261 mod p, plain residue: mont_mul(x*R, r261) = x * 2^261 mod p
  uint32_t pad2[6];
};

#if defined(__CUDACC__)
__constant__ FieldParams c_field;  // FieldAny reads this (set per context)
#endif

// BLS12-381 scalar field r (honeybadgermpc/elliptic_curve.py:5).
struct FieldBLS {
  static HB_HD constexpr uint32_t p(int i) {
    return i == 0 ? 0x00000001u : i == 1 ? 0xffffffffu : i == 2 ? 0xfffe5bfeu
         : i == 3 ? 0x53bda402u : i == 4 ? 0x09a1d805u : i == 5 ? 0x3339d808u
         : i == 6 ? 0x299d7d48u : 0x73eda753u;
  }
  static HB_HD constexpr uint32_t n0inv() { return 0xffffffffu; }
  static HB_HD constexpr uint32_t p29(int i) {
    return i == 0 ? 0x00000001u : i == 1 ? 0x1ffffff8u : i == 2 ? 0x1f96ffbfu
         : i == 3 ? 0x1b4805ffu : i == 4 ? 0x1d80553bu : i == 5 ? 0x0c0404d0u
         : i == 6 ? 0x1520cce7u : i == 7 ? 0x0a6533afu : 0x0073eda7u;
  }
  static HB_HD constexpr uint32_t n0inv29() { return 0x1fffffffu; }
  // p = 1 (mod 2^32) and p[1] = 2^32-1: the two lowest columns of m*p are
  // adds/subs of m, not multiplies (see redc_row_low_ones).
  static constexpr bool kLowOnes = true;
  // macs between two acc_fold calls: p*R + 2*p^2 < 2^512 for this p
  static constexpr int kFold = 2;
};

// Same field, but the limbs come from the constant bank instead of immediates
// (lets ptxas keep single IMAD.WIDE ops in the reduction rows).
struct FieldBLSConst;

#if defined(__CUDACC__)
struct FieldAny {
  static HB_D uint32_t p(int i) { return c_field.p[i]; }
  static HB_D uint32_t n0inv() { return c_field.n0inv; }
  static HB_D uint32_t p29(int i) { return c_field.p29[i]; }
  static HB_D uint32_t n0inv29() { return c_field.n0inv29; }
  static constexpr bool kLowOnes = false;
  static constexpr int kFold = 1;
};
struct FieldBLSConst {
  static HB_D uint32_t p(int i) { return c_field.p[i]; }
  static HB_D uint32_t n0inv() { return 0xffffffffu; }
  static HB_D uint32_t p29(int i) { return c_field.p29[i]; }
  static HB_D uint32_t n0inv29() { return 0x1fffffffu; }
  static constexpr bool kLowOnes = true;
  static constexpr int kFold = 2;
};
#endif

// Host-side policy with the parameters in a plain struct (used by the host
// precompute code and by the CPU unit tests of this header).
struct FieldHost {
  static inline const FieldParams*& cur() {
    static thread_local const FieldParams* f = nullptr;
    return f;
  }
  static inline uint32_t p(int i) { return cur()->p[i]; }
  static inline uint32_t n0inv() { return cur()->n0inv; }
  static inline uint32_t p29(int i) { return cur()->p29[i]; }
  static inline uint32_t n0inv29() { return cur()->n0inv29; }
  static constexpr bool kLowOnes = false;
  static constexpr int kFold = 1;
};
// Host twin of the BLS fast reduction (unit tests only).
struct FieldHostLowOnes : FieldHost {
  static constexpr bool kLowOnes = true;
};

// ---------------------------------------------------------------------------
// carry-chain primitives.  Device: one asm statement per chain (the carry flag
// never crosses an asm boundary).  Host: the same dataflow in 64-bit C.
// ---------------------------------------------------------------------------

// acc[2j],acc[2j+1] = a_j * b   (four independent 32x32->64 products)
HB_HD void mul4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mul.lo.u32 %0, %8, %12;\n\t mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t mul.hi.u32 %7, %11, %12;"
      : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]),
        "=r"(acc[6]), "=r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
  const uint32_t a[4] = {a0, a1, a2, a3};
  for (int j = 0; j < 4; j++) {
    uint64_t t = (uint64_t)a[j] * b;
    acc[2 * j] = (uint32_t)t;
    acc[2 * j + 1] = (uint32_t)(t >> 32);
  }
#endif
}

// acc (8 words) += sum_j a_j * b * 2^(64 j); the carry out of word 7 is added to `top`.
HB_HD void cmad4_top(uint32_t* acc, uint32_t& top, uint32_t a0, uint32_t a1, uint32_t a2,
                     uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
  const uint32_t a[4] = {a0, a1, a2, a3};
  uint64_t c = 0;
  for (int j = 0; j < 4; j++) {
    uint64_t lo = (uint64_t)acc[2 * j] + (uint32_t)((uint64_t)a[j] * b) + c;
    acc[2 * j] = (uint32_t)lo;
    uint64_t hi = (uint64_t)acc[2 * j + 1] + (uint32_t)(((uint64_t)a[j] * b) >> 32) + (lo >> 32);
    acc[2 * j + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  top += (uint32_t)c;
#endif
}

// Same chain, for the deferred-carry counters of the lazy accumulator (acc_mac): the
// carry-out is first materialised in a fresh register and then added to the counter.
// With "addc top, top, 0" ptxas postpones the counter updates (nothing needs them before
// acc_fold), runs out of predicate registers for the pending carries and packs them bit
// by bit into a general register: ~250 LOP3 per 6-term dot product, a fifth of the
// kernel's instructions.  The fresh register ends the predicate's life at once.
HB_HD void cmad4_cnt(uint32_t* acc, uint32_t& cnt, uint32_t a0, uint32_t a1, uint32_t a2,
                     uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t cy;
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
      "addc.u32 %8, 0, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7]), "=r"(cy)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
  cnt += cy;
#else
  cmad4_top(acc, cnt, a0, a1, a2, a3, b);
#endif
}

// Same, when the mathematical bound guarantees no carry out of word 7.
HB_HD void cmad4(uint32_t* acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.u32 %7, %11, %12, %7;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),
        "+r"(acc[6]), "+r"(acc[7])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
  uint32_t sink = 0;
  cmad4_top(acc, sink, a0, a1, a2, a3, b);
#endif
}

// The row-entry step of the even/odd CIOS (rows 1..7):
//   e[0] += o[1]                      (the word that falls onto column 0 after the shift)
//   o    = (o >> 64) + a_odd * b      (carry of the add above enters at column 1 = o[0])
HB_HD void shift_madc4(uint32_t* e, uint32_t* o, uint32_t a1, uint32_t a3, uint32_t a5,
                       uint32_t a7, uint32_t b) {
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %8, %8, %1;\n\t"
      "madc.lo.cc.u32 %0, %9, %13, %2;\n\t madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %4;\n\t madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %6;\n\t madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, 0;\n\t madc.hi.u32 %7, %12, %13, 0;"
      : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]),
        "+r"(o[7]), "+r"(e[0])
      : "r"(a1), "r"(a3), "r"(a5), "r"(a7), "r"(b));
#else
  uint64_t s = (uint64_t)e[0] + o[1];
  e[0] = (uint32_t)s;
  uint64_t c = s >> 32;
  const uint32_t a[4] = {a1, a3, a5, a7};
  for (int j = 0; j < 4; j++) {
    uint32_t in_lo = (j < 3) ? o[2 * j + 2] : 0u;
    uint32_t in_hi = (j < 3) ? o[2 * j + 3] : 0u;
    uint64_t prod = (uint64_t)a[j] * b;
    uint64_t lo = (uint64_t)in_lo + (uint32_t)prod + c;
    uint64_t hi = (uint64_t)in_hi + (uint32_t)(prod >> 32) + (lo >> 32);
    o[2 * j] = (uint32_t)lo;
    o[2 * j + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
#endif
}


// Reduction row for moduli with p[0] = 1, p[1] = 2^32-1 (then n0inv = -1 and
// m = -e[0]):  e += m*(p0 + p2 2^64 + p4 2^128 + p6 2^192) where m*p0 = m is a
// plain add that zeroes e[0];  o += m*(p1 + p3 2^64 + ...) where
// m*p1 = m*2^32 - m is a 64-bit subtract.  6 multiplies per row instead of 8,
// and no multiply to form m.
HB_HD void redc_row_low_ones(uint32_t* e, uint32_t* o, uint32_t p2, uint32_t p3, uint32_t p4,
                             uint32_t p5, uint32_t p6, uint32_t p7) {
#if defined(__CUDA_ARCH__)
  uint32_t m, lo, hi;
  asm("sub.u32 %0, 0, %3;\n\t"            // m = -e0
      "sub.cc.u32 %1, 0, %0;\n\t"         // (hi:lo) = m*2^32 - m
      "subc.u32 %2, %0, 0;"
      : "=r"(m), "=r"(lo), "=r"(hi) : "r"(e[0]));
  asm("add.cc.u32 %0, %0, %8;\n\t addc.cc.u32 %1, %1, %9;\n\t"
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t madc.hi.u32 %7, %12, %13, %7;"
      : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]),
        "+r"(o[7])
      : "r"(lo), "r"(hi), "r"(p3), "r"(p5), "r"(p7), "r"(m));
  asm("add.cc.u32 %0, %0, %12;\n\t addc.cc.u32 %1, %1, 0;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.cc.u32 %7, %11, %12, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]),
        "+r"(e[7]), "+r"(o[7])
      : "r"(p2), "r"(p4), "r"(p6), "r"(m));
#else
  uint32_t m = 0u - e[0];
  cmad4(o, 0xffffffffu, p3, p5, p7, m);
  cmad4_top(e, o[7], 1u, p2, p4, p6, m);
#endif
}

// r = x + y (8 words), returns carry out
HB_HD uint32_t add8(uint32_t* r, const uint32_t* x, const uint32_t* y) {
#if defined(__CUDA_ARCH__)
  uint32_t c;
  asm("add.cc.u32 %0, %9, %17;\n\t addc.cc.u32 %1, %10, %18;\n\t addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t addc.cc.u32 %4, %13, %21;\n\t addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t addc.cc.u32 %7, %16, %24;\n\t addc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(c)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
  return c;
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)x[i] + y[i] + c;
    r[i] = (uint32_t)s;
    c = s >> 32;
  }
  return (uint32_t)c;
#endif
}

// r = x - y (8 words), returns borrow (1 if x < y)
HB_HD uint32_t sub8(uint32_t* r, const uint32_t* x, const uint32_t* y) {
#if defined(__CUDA_ARCH__)
  uint32_t b;
  asm("sub.cc.u32 %0, %9, %17;\n\t subc.cc.u32 %1, %10, %18;\n\t subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t subc.cc.u32 %4, %13, %21;\n\t subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t subc.cc.u32 %7, %16, %24;\n\t subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(b)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
  return b & 1u;
#else
  uint64_t bw = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)x[i] - y[i] - bw;
    r[i] = (uint32_t)d;
    bw = (d >> 32) & 1u;
  }
  return (uint32_t)bw;
#endif
}

template <class F>
HB_HD void load_p(uint32_t* p) {
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = F::p(i);
}

// x in [0, 2p) -> [0, p)   (carry_in: a 257th bit of x, when the caller has one)
template <class F>
HB_HD void cond_sub_p(Fe& x, uint32_t carry_in = 0) {
  uint32_t p[8], d[8];
  load_p<F>(p);
  uint32_t borrow = sub8(d, x.w, p);
  bool take = (carry_in != 0) || (borrow == 0);
#pragma unroll
  for (int i = 0; i < 8; i++) x.w[i] = take ? d[i] : x.w[i];
}

template <class F>
HB_HD Fe fe_add(const Fe& a, const Fe& b) {
  Fe r;
  uint32_t c = add8(r.w, a.w, b.w);
  cond_sub_p<F>(r, c);
  return r;
}

template <class F>
HB_HD Fe fe_sub(const Fe& a, const Fe& b) {
  Fe r;
  uint32_t p[8], s[8];
  load_p<F>(p);
  uint32_t borrow = sub8(r.w, a.w, b.w);
  add8(s, r.w, p);
#pragma unroll
  for (int i = 0; i < 8; i++) r.w[i] = borrow ? s[i] : r.w[i];
  return r;
}

template <class F>
HB_HD Fe fe_neg(const Fe& a) {
  Fe r;
  uint32_t p[8];
  load_p<F>(p);
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) nz |= a.w[i];
  sub8(r.w, p, a.w);
#pragma unroll
  for (int i = 0; i < 8; i++) r.w[i] = nz ? r.w[i] : 0u;
  return r;
}

HB_HD bool fe_is_zero(const Fe& a) {
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) nz |= a.w[i];
  return nz == 0;
}

HB_HD bool fe_eq(const Fe& a, const Fe& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.w[i] ^ b.w[i];
  return d == 0;
}

HB_HD Fe fe_zero() {
  Fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.w[i] = 0;
  return r;
}

// Montgomery product a*b/R mod p, canonical output.  a is the multiplicand (all
// of its limbs enter every row), b supplies the row digits.  Requires a < p; b
// may be any value < 2^256 (every row keeps t < a + p, the result is < 2p
// before the final conditional subtraction).
template <class F>
HB_HD Fe mont_mul(const Fe& a, const Fe& b) {
  uint32_t x[8], y[8];  // two column accumulators whose even/odd roles swap every row
  uint32_t p[8];
  load_p<F>(p);
  const uint32_t n0 = F::n0inv();

  // row 0: x = even columns, y = odd columns
  mul4(x, a.w[0], a.w[2], a.w[4], a.w[6], b.w[0]);
  mul4(y, a.w[1], a.w[3], a.w[5], a.w[7], b.w[0]);
  if (F::kLowOnes) {
    redc_row_low_ones(x, y, p[2], p[3], p[4], p[5], p[6], p[7]);
  } else {
    uint32_t m = x[0] * n0;
    cmad4(y, p[1], p[3], p[5], p[7], m);
    cmad4_top(x, y[7], p[0], p[2], p[4], p[6], m);
  }
#pragma unroll
  for (int i = 1; i < 8; i++) {
    uint32_t* e = (i & 1) ? y : x;  // even-aligned this row
    uint32_t* o = (i & 1) ? x : y;  // odd-aligned this row (holds last row's even part)
    shift_madc4(e, o, a.w[1], a.w[3], a.w[5], a.w[7], b.w[i]);
    cmad4_top(e, o[7], a.w[0], a.w[2], a.w[4], a.w[6], b.w[i]);
    if (F::kLowOnes) {
      redc_row_low_ones(e, o, p[2], p[3], p[4], p[5], p[6], p[7]);
    } else {
      uint32_t m = e[0] * n0;
      cmad4(o, p[1], p[3], p[5], p[7], m);
      cmad4_top(e, o[7], p[0], p[2], p[4], p[6], m);
    }
  }
  // after row 7: even-aligned = y, odd-aligned = x;  result = (y >> 32) + x
  Fe r;
  uint32_t hi[8];
#pragma unroll
  for (int i = 0; i < 7; i++) hi[i] = y[i + 1];
  hi[7] = 0;
  add8(r.w, x, hi);
  cond_sub_p<F>(r);
  return r;
}


// r = x + y + cin (8 words, cin in {0,1}), returns carry out
HB_HD uint32_t add8c(uint32_t* r, const uint32_t* x, const uint32_t* y, uint32_t cin) {
#if defined(__CUDA_ARCH__)
  uint32_t c;
  asm("{\n\t.reg .u32 t;\n\t"
      "add.cc.u32 t, %25, 0xffffffff;\n\t"  // carry flag := cin
      "addc.cc.u32 %0, %9, %17;\n\t addc.cc.u32 %1, %10, %18;\n\t addc.cc.u32 %2, %11, %19;\n\t"
      "addc.cc.u32 %3, %12, %20;\n\t addc.cc.u32 %4, %13, %21;\n\t addc.cc.u32 %5, %14, %22;\n\t"
      "addc.cc.u32 %6, %15, %23;\n\t addc.cc.u32 %7, %16, %24;\n\t addc.u32 %8, 0, 0;\n\t}"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(c)
      : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]),
        "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]),
        "r"(cin));
  return c;
#else
  uint64_t c = cin;
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)x[i] + y[i] + c;
    r[i] = (uint32_t)s;
    c = s >> 32;
  }
  return (uint32_t)c;
#endif
}

// Row entry of the even/odd CIOS when the multiplier digit is zero:
//   e[0] += o[1];  o = (o >> 64) + carry
HB_HD void shift_row(uint32_t* e, uint32_t* o) {
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %8, %8, %1;\n\t"
      "addc.cc.u32 %0, %2, 0;\n\t addc.cc.u32 %1, %3, 0;\n\t addc.cc.u32 %2, %4, 0;\n\t"
      "addc.cc.u32 %3, %5, 0;\n\t addc.cc.u32 %4, %6, 0;\n\t addc.cc.u32 %5, %7, 0;\n\t"
      "addc.u32 %6, 0, 0;\n\t mov.u32 %7, 0;"
      : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]),
        "+r"(o[7]), "+r"(e[0]));
#else
  uint64_t s = (uint64_t)e[0] + o[1];
  e[0] = (uint32_t)s;
  uint64_t c = s >> 32;
  for (int j = 0; j < 6; j++) {
    uint64_t v = (uint64_t)o[j + 2] + c;
    o[j] = (uint32_t)v;
    c = v >> 32;
  }
  o[6] = (uint32_t)c;
  o[7] = 0;
#endif
}

template <class F>
HB_HD void redc_row(uint32_t* e, uint32_t* o, const uint32_t* p) {
  if (F::kLowOnes) {
    redc_row_low_ones(e, o, p[2], p[3], p[4], p[5], p[6], p[7]);
  } else {
    uint32_t m = e[0] * F::n0inv();
    cmad4(o, p[1], p[3], p[5], p[7], m);
    cmad4_top(e, o[7], p[0], p[2], p[4], p[6], m);
  }
}

// a / R mod p for a < 2^256 (a need not be reduced): mont_mul(a, 1) without the
// multiply rows.  Canonical output.
template <class F>
HB_HD Fe mont_redc_lo(const uint32_t* a) {
  uint32_t x[8], y[8], p[8];
  load_p<F>(p);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    x[2 * j] = a[2 * j];
    x[2 * j + 1] = 0;
    y[2 * j] = a[2 * j + 1];
    y[2 * j + 1] = 0;
  }
  redc_row<F>(x, y, p);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    uint32_t* e = (i & 1) ? y : x;
    uint32_t* o = (i & 1) ? x : y;
    shift_row(e, o);
    redc_row<F>(e, o, p);
  }
  Fe r;
  uint32_t hi[8];
#pragma unroll
  for (int i = 0; i < 7; i++) hi[i] = y[i + 1];
  hi[7] = 0;
  add8(r.w, x, hi);
  cond_sub_p<F>(r);
  return r;
}

// ---------------------------------------------------------------------------
// Lazy-reduction dot products.  A row dot product sum_j a_j*b_j is accumulated
// as a 512-bit integer and Montgomery-reduced ONCE (instead of once per
// product): 64 IMAD.WIDE per term + 48 (BLS) / 64 per output.
//
// Representation: value = E + (O << 32) + deferred carries, with E = e[0..16)
// (word w has weight 2^(32w)) and O = o[0..16) (word w has weight 2^(32(w+1))).
// Products a_j*b_i land on E when i+j is even and on O when it is odd, so every
// 32x32->64 multiply-add works on a 64-bit ALIGNED register pair and fuses into
// one IMAD.WIDE.  A 4-product chain that starts at word s drops its carry-out
// (destined for word s+8) into ke[s/2] / ko[s/2] instead of rippling upwards,
// so each chain is one short asm statement; only words >= 8 receive deferred
// carries, the low half is always exact.
// ---------------------------------------------------------------------------
struct Acc {
  uint32_t e[16];
  uint32_t o[16];
  uint32_t ke[4], ko[4];
};

HB_HD void acc_zero(Acc& t) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    t.e[i] = 0;
    t.o[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    t.ke[i] = 0;
    t.ko[i] = 0;
  }
}

// t += a*b.  Caller keeps the true value below 2^512 (see acc_fold).
// FRESH: count the chain carries through a fresh register (cmad4_cnt) -- for kernels that
// unroll several macs back to back (interp_small_kernel); loops of one mac per iteration
// are better off with the direct form.
template <bool FRESH = false>
HB_HD void acc_mac(Acc& t, const Fe& a, const Fe& b) {
  auto chain = [](uint32_t* acc, uint32_t& cnt, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                  uint32_t d) {
    if (FRESH)
      cmad4_cnt(acc, cnt, a0, a1, a2, a3, d);
    else
      cmad4_top(acc, cnt, a0, a1, a2, a3, d);
  };
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    // even digit b_i: a_even*b_i -> E[i..i+8), a_odd*b_i -> O[i..i+8)
    chain(t.e + i, t.ke[i / 2], a.w[0], a.w[2], a.w[4], a.w[6], b.w[i]);
    chain(t.o + i, t.ko[i / 2], a.w[1], a.w[3], a.w[5], a.w[7], b.w[i]);
    // odd digit b_{i+1}: a_odd*b -> E[i+2..i+10), a_even*b -> O[i..i+8)
    if (i + 2 < 8) {
      chain(t.e + i + 2, t.ke[i / 2 + 1], a.w[1], a.w[3], a.w[5], a.w[7], b.w[i + 1]);
    } else {
      cmad4(t.e + 8, a.w[1], a.w[3], a.w[5], a.w[7], b.w[7]);  // top chain: value < 2^512
    }
    chain(t.o + i, t.ko[i / 2], a.w[0], a.w[2], a.w[4], a.w[6], b.w[i + 1]);
  }
}

// Collapse E, O and the deferred carries into e[0..16) (o, ke, ko := 0) and bring
// the value below p*2^256 by one conditional subtraction of p*2^256.
// Precondition: value < 2*p*2^256 and < 2^512.  With inputs a < p, b < p each
// product is < p^2 < 0.46*p*2^256, so for the BLS field (p ~ 0.453*2^256) one
// fold per TWO macs keeps both bounds (p*R + 2*p^2 < 0.87 * 2^512).  Generic
// fields fold after every mac (p*R + p^2 < 2*p*R; < 2^512 needs p < 2^255, which
// hbg_ctx_create enforces).
template <class F>
HB_HD void acc_fold(Acc& t) {
  uint32_t p[8], lo[8], hi[8], d[8], sh[8], kk[8];
  load_p<F>(p);
  // low half: e[0..8) + (o[0..7) << 32)
  sh[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) sh[i] = t.o[i - 1];
  uint32_t c = add8(lo, t.e, sh);
  // high half: e[8..16) + o[7..15) + carries
#pragma unroll
  for (int i = 0; i < 8; i++) sh[i] = t.o[7 + i];
  add8c(hi, t.e + 8, sh, c);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    kk[2 * i] = t.ke[i];
    kk[2 * i + 1] = t.ko[i];
  }
  add8(hi, hi, kk);
  uint32_t borrow = sub8(d, hi, p);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    t.e[i] = lo[i];
    t.e[8 + i] = borrow ? hi[i] : d[i];
    t.o[i] = 0;
    t.o[8 + i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    t.ke[i] = 0;
    t.ko[i] = 0;
  }
}

// Montgomery reduction of a folded accumulator (value = e[0..16) < p*2^256):
// value/R mod p, canonical.   T/R = mont_redc_lo(T_lo) + T_hi  (mod p).
template <class F>
HB_HD Fe acc_redc(Acc& t) {
  Fe lo = mont_redc_lo<F>(t.e);
  Fe hi;
#pragma unroll
  for (int i = 0; i < 8; i++) hi.w[i] = t.e[8 + i];
  return fe_add<F>(lo, hi);
}

template <class F>
HB_HD Fe mont_sqr(const Fe& a) {
  return mont_mul<F>(a, a);
}

}  // namespace hb
