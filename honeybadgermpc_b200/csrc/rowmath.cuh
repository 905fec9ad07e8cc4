// Per-thread row arithmetic shared by the kernels and their host twins
// (host_selftest.cpp compiles this file with g++: tests/test_host_field.py).
//
//  * ntt16_half: one thread's half of a 16-point NTT of <= 8 coefficients,
//    organised as two 4-point groups so that only four values are live at a
//    time (the n = 16 encode of the headline config; rsdecode_impl.h:125-192
//    computes the same DFT recursively).
//  * mac29 / redc29: carry-free lazy dot products in radix 2^29, the row math of
//    the small-k interpolation kernel.
#pragma once
#include "fp256.cuh"

namespace hb {

// Lazy subtraction: inputs < p, result a + (p - b) in [0, 2p), which fits 256 bits because
// p < 2^255 (hbg_ctx_create enforces it).  Only ever used as the DIGIT operand
// (second argument) of mont_mul: with the multiplicand a < p every CIOS row keeps
// t < a + p < 2p whatever the digits are, and the output is canonical again.
// (The multiplicand side must stay below 2^256 - p.)
template <class F>
HB_HD Fe fe_sub_lazy(const Fe& a, const Fe& b) {
  uint32_t p[8], t[8];
  load_p<F>(p);
  sub8(t, p, b.w);  // p - b in (0, p]
  Fe r;
  add8(r.w, a.w, t);
  return r;
}

// 4-point DFT on zeta = omega^4 (zeta^2 = -1) of canonical a[0..4):
//   Z[r] = sum_i a_i zeta^(i r);  emit(r, Z[r]) for r = 0..3.
template <class F, class Mul, class Emit>
HB_HD void dft4(const Fe* a, Mul mul, Emit emit) {
  Fe b0 = fe_add<F>(a[0], a[2]);
  Fe b1 = fe_add<F>(a[1], a[3]);
  emit(0, fe_add<F>(b0, b1));
  emit(2, fe_sub<F>(b0, b1));
  Fe b2 = fe_sub<F>(a[0], a[2]);
  Fe b3 = mul(4, fe_sub_lazy<F>(a[1], a[3]));
  emit(1, fe_add<F>(b2, b3));
  emit(3, fe_sub<F>(b2, b3));
}

// X[k] = sum_{i<D} c_i omega^(i k), k < 16, D <= 8, for the outputs of parity H:
//   k = 4 r + 2 g + H,  g = 0, 1 (group), r = 0..3.
// With s_i = c_i omega^(H i):
//   group 0: a_i = s_i + s_{i+4}
//   group 1: a_i = (s_i - s_{i+4}) omega^(2 i)
// and X[4 r + 2 g + H] = sum_i a_i (omega^4)^(i r).
// ld(i) -> c_i (zero beyond the row's real length); mul(j, x) -> omega^j * x, canonical,
// for j < 16 and any x < 2^256 (mont_mul with omega^j in Montgomery form as the
// multiplicand and x on the digit side); st(k, X[k]).  D is a compile-time bound on the
// number of coefficients (4 <= D <= 8): operands that are structurally zero cost nothing.
//
// Load balance: the odd half needs twice the multiplications of the even half (D = 6:
// 10 against 5).  The group-1 inputs of the odd half that come from a single coefficient,
// a_i = c_i omega^(3 i) for i >= 4 - OFF, depend on nothing else, so the EVEN thread
// computes them first and hands them over: xch.put(i, v) + xch.signal() on the even
// side, xch.wait() + xch.get(i) on the odd side.  D = 6, OFF = 2: 7 against 8.
// Measured on the B200 (n = 16, d = 6): no gain at 65 536 polynomials (29.5 us either way)
// and 9 % slower at 1 Mi (the hand-over costs more than the idle half-warps, which the
// next CTAs' warps fill anyway), so the kernels run with BAL = false; the balanced form
// stays as a tested option.
constexpr int ntt16_offload(int D, bool bal) { return !bal ? 0 : D == 6 ? 2 : D <= 5 ? 1 : 0; }

template <class F, int D, int H, bool BAL, class Ld, class Mul, class St, class Xch>
HB_HD void ntt16_half(Ld ld, Mul mul, St st, Xch& xch) {
  static_assert(D >= 4 && D <= 8, "D in [4, 8]");
  constexpr int OFF = ntt16_offload(D, BAL);
  static_assert(8 - OFF >= D, "offloaded inputs must be single-coefficient ones");
  if (H == 0 && OFF > 0) {
#pragma unroll
    for (int i = 4 - OFF; i < 4; i++) xch.put(i, mul(3 * i, ld(i)));
    xch.signal();
  }
  Fe a[4];
  Fe dif[4];  // s_i - s_{i+4} for the pairs that exist: canonical for i = 0, lazy otherwise
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (i + 4 < D) {
      Fe si = (H && i > 0) ? mul(i, ld(i)) : ld(i);
      Fe sj = H ? mul(i + 4, ld(i + 4)) : ld(i + 4);
      a[i] = fe_add<F>(si, sj);
      dif[i] = (i == 0) ? fe_sub<F>(si, sj) : fe_sub_lazy<F>(si, sj);
    } else {
      a[i] = (H && i > 0) ? mul(i, ld(i)) : ld(i);
    }
  }
  dft4<F>(a, mul, [&](int r, const Fe& v) { st(4 * r + H, v); });
  if (H == 1 && OFF > 0) xch.wait();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (i + 4 < D) {
      a[i] = (i == 0) ? dif[0] : mul(2 * i, dif[i]);
    } else if (i == 0) {
      a[i] = ld(0);
    } else if (H == 1 && i >= 4 - OFF) {
      a[i] = xch.get(i);
    } else {
      a[i] = mul(H ? 3 * i : 2 * i, ld(i));
    }
  }
  dft4<F>(a, mul, [&](int r, const Fe& v) { st(4 * r + 2 + H, v); });
}

// ---------------------------------------------------------------------------
// Carry-free lazy dot products in radix 2^29 (R' = 2^261).
//
// Measured on the B200 (tools/microbench3.cu): IMAD.WIDE issues at ~58 per clock
// per SM, but only ~31 when it is part of a predicate carry chain
// (mad.lo.cc / madc.hi.cc).  With 9 limbs of 29 bits a product is < 2^58, so a
// 64-bit column accumulator takes the 9 products of a column for up to 7 terms
// (8 with canonical operands, see mac29) with NO carry at all: 81 plain IMAD.WIDE
// per term instead of 64 chained ones, and the Montgomery reduction (9 rounds of
// m * p, R' = 2^261) is carry-free as well.  R' = 32 R leaves 5 bits of headroom:
// sum of K products < K p^2 reduces to < (1 + K * 0.0142) p, one conditional
// subtraction for any K <= 64, so there is no periodic fold either.
// Constants are stored as 9 x 29-bit limbs of M * 2^261 mod p; data stays in
// standard form and is re-limbed on the fly (one funnel shift + mask per limb).
// ---------------------------------------------------------------------------
constexpr uint32_t kMask29 = (1u << 29) - 1;

HB_HD uint32_t shr_pair(uint32_t lo, uint32_t hi, int sh) {  // bits [sh, sh+32) of hi:lo, 0 < sh < 32
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}

// x (8 x 32-bit words, any value < 2^256) -> 9 limbs of 29 bits
HB_HD void to_limbs29(const Fe& x, uint32_t* l) {
  l[0] = x.w[0] & kMask29;
  l[1] = shr_pair(x.w[0], x.w[1], 29) & kMask29;
  l[2] = shr_pair(x.w[1], x.w[2], 26) & kMask29;
  l[3] = shr_pair(x.w[2], x.w[3], 23) & kMask29;
  l[4] = shr_pair(x.w[3], x.w[4], 20) & kMask29;
  l[5] = shr_pair(x.w[4], x.w[5], 17) & kMask29;
  l[6] = shr_pair(x.w[5], x.w[6], 14) & kMask29;
  l[7] = shr_pair(x.w[6], x.w[7], 11) & kMask29;
  l[8] = x.w[7] >> 8;
}

// 9 limbs (each < 2^29, value < 2^256) -> 8 words
HB_HD Fe from_limbs29(const uint32_t* l) {
  Fe r;
  r.w[0] = l[0] | (l[1] << 29);
  r.w[1] = (l[1] >> 3) | (l[2] << 26);
  r.w[2] = (l[2] >> 6) | (l[3] << 23);
  r.w[3] = (l[3] >> 9) | (l[4] << 20);
  r.w[4] = (l[4] >> 12) | (l[5] << 17);
  r.w[5] = (l[5] >> 15) | (l[6] << 14);
  r.w[6] = (l[6] >> 18) | (l[7] << 11);
  r.w[7] = (l[7] >> 21) | (l[8] << 8);
  return r;
}

// col[0..17) += a * b (limb products by column).  No carries: the caller keeps every
// column below 2^64.  With a, b < 2^256 in 29-bit limbs (top limbs < 2^24) one term
// adds less than 9 * 2^58 to a column, so 7 terms always fit; with canonical
// operands of a field with p < 2^255 the 9-product column is smaller and 8 terms fit.
HB_HD void mac29(uint64_t* col, const uint32_t* a, const uint32_t* b) {
#pragma unroll
  for (int x = 0; x < 9; x++) {
#pragma unroll
    for (int y = 0; y < 9; y++) col[x + y] += (uint64_t)a[x] * b[y];
  }
}

// Column carry pass: every column but the last below 2^29 afterwards (value unchanged).
HB_HD void norm29(uint64_t* col) {
#pragma unroll
  for (int c = 0; c < 16; c++) {
    col[c + 1] += col[c] >> 29;
    col[c] &= kMask29;
  }
}

// Montgomery reduction of T = sum col[c] 2^(29 c) by R' = 2^261: T / R' mod p,
// canonical.  Requires T < 64 p^2-ish (result before the final conditional
// subtraction < 2p) and every column, plus the 9 reduction products it receives
// (< 9 * 2^58) and a carry (< 2^36), below 2^64: callers with more than 6 terms run
// norm29 first.
template <class F>
HB_HD Fe redc29(uint64_t* col) {
  uint64_t carry = 0;
#pragma unroll
  for (int r = 0; r < 9; r++) {
    uint64_t t = col[r] + carry;
    const uint32_t m = ((uint32_t)t * F::n0inv29()) & kMask29;
    t += (uint64_t)m * F::p29(0);  // low 29 bits are now zero
    carry = t >> 29;
#pragma unroll
    for (int j = 1; j < 9; j++) col[r + j] += (uint64_t)m * F::p29(j);
  }
  uint32_t l[9];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const uint64_t t = col[9 + c] + carry;
    l[c] = (uint32_t)t & kMask29;
    carry = t >> 29;
  }
  l[8] = (uint32_t)carry;
  Fe r = from_limbs29(l);
  cond_sub_p<F>(r);
  return r;
}

}  // namespace hb
