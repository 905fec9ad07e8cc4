// Host-side field helpers of the product library: parameter derivation and the
// O(n^2) per-call constant tables (Vandermonde matrices and inverses, twiddle
// tables, the x-only part of FFT interpolation).  These depend only on the
// evaluation points, never on the batch; the batch data path is CUDA only.
// Uses the same fp256.cuh algorithms as the kernels (host twins).
#pragma once
#include <stdint.h>
#include <string.h>

#include <stdexcept>
#include <vector>

#include "fp256.cuh"

namespace hb {

inline Fe fe_from_u64(const uint64_t* l) {
  Fe r;
  for (int i = 0; i < 4; i++) {
    r.w[2 * i] = (uint32_t)l[i];
    r.w[2 * i + 1] = (uint32_t)(l[i] >> 32);
  }
  return r;
}

inline void fe_to_u64(const Fe& a, uint64_t* l) {
  for (int i = 0; i < 4; i++) l[i] = (uint64_t)a.w[2 * i] | ((uint64_t)a.w[2 * i + 1] << 32);
}

inline Fe fe_from_u32(uint32_t v) {
  Fe r = fe_zero();
  r.w[0] = v;
  return r;
}

inline int fe_cmp(const Fe& a, const Fe& b) {
  for (int i = 7; i >= 0; i--) {
    if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
  }
  return 0;
}

// Derive R mod p, R^2 mod p and -p^-1 mod 2^32.  Returns false unless p is odd
// and >= 3 (Montgomery needs gcd(p, 2^32) = 1).
inline bool field_params_init(const uint64_t p64[4], FieldParams* fp) {
  memset(fp, 0, sizeof(*fp));
  Fe p = fe_from_u64(p64);
  memcpy(fp->p, p.w, sizeof(p.w));
  if ((p.w[0] & 1u) == 0) return false;
  bool ge3 = p.w[0] >= 3;
  for (int i = 1; i < 8; i++) ge3 = ge3 || p.w[i] != 0;
  if (!ge3) return false;
  uint32_t inv = 1;  // Newton: inv = p^-1 mod 2^32
  for (int i = 0; i < 5; i++) inv *= 2u - p.w[0] * inv;
  fp->n0inv = 0u - inv;
  // x = 2^i mod p by repeated doubling
  Fe x = fe_zero();
  x.w[0] = 1;
  // p may be tiny: reduce 1 mod p is still 1 since p >= 3
  for (int i = 0; i < 512; i++) {
    uint32_t d[8], s[8];
    uint32_t carry = add8(d, x.w, x.w);
    uint32_t borrow = sub8(s, d, p.w);
    bool take = carry || !borrow;
    for (int j = 0; j < 8; j++) x.w[j] = take ? s[j] : d[j];
    if (i == 255) memcpy(fp->one, x.w, sizeof(x.w));
    if (i == 260) memcpy(fp->r261, x.w, sizeof(x.w));  // 2^261 mod p, plain residue
  }
  memcpy(fp->r2, x.w, sizeof(x.w));
  // radix-2^29 view: limbs of p and -p^-1 mod 2^29
  for (int i = 0; i < 9; i++) {
    int bit = 29 * i, wd = bit / 32, sh = bit % 32;
    uint64_t v = p.w[wd] >> sh;
    if (sh > 3 && wd + 1 < 8) v |= (uint64_t)p.w[wd + 1] << (32 - sh);
    fp->p29[i] = (uint32_t)v & 0x1fffffffu;
  }
  fp->n0inv29 = fp->n0inv & 0x1fffffffu;
  return true;
}

// floor(2^280 / p) for 2^248 < p < 2^256 (the small-quotient Barrett constant of
// tc_kernels.cuh: tc_fold_reduce); 0 when it does not fit 32 bits.
inline uint32_t barrett_mu280(const FieldParams& fp) {
  uint32_t rem[9] = {0};
  unsigned long long q = 0;
  for (int bit = 280; bit >= 0; bit--) {  // long division of 2^280, bit by bit
    uint32_t carry = bit == 280 ? 1u : 0u;
    for (int w = 0; w < 9; w++) {
      const uint32_t nc = rem[w] >> 31;
      rem[w] = (rem[w] << 1) | carry;
      carry = nc;
    }
    uint32_t t[9];
    long long b = 0;
    for (int w = 0; w < 9; w++) {
      const long long dd = (long long)rem[w] - (long long)(w < 8 ? fp.p[w] : 0u) + b;
      t[w] = (uint32_t)dd;
      b = dd >> 32;
    }
    q <<= 1;
    if (b == 0) {
      memcpy(rem, t, sizeof t);
      q |= 1;
    }
    if (q >> 32) return 0;
  }
  return (uint32_t)q;
}

// A field bound to its parameters; all values held by users of this class are
// in MONTGOMERY form unless a name says "std".
class HostField {
 public:
  explicit HostField(const FieldParams& fp) : fp_(fp) {
    memcpy(one_.w, fp.one, 32);
    memcpy(r2_.w, fp.r2, 32);
    memcpy(p_.w, fp.p, 32);
  }
  const FieldParams& params() const { return fp_; }
  struct Scope {
    const FieldParams* prev;
    explicit Scope(const FieldParams* f) : prev(FieldHost::cur()) { FieldHost::cur() = f; }
    ~Scope() { FieldHost::cur() = prev; }
  };
  Fe one() const { return one_; }
  Fe mul(const Fe& a, const Fe& b) const {
    Scope s(&fp_);
    return mont_mul<FieldHost>(a, b);
  }
  Fe add(const Fe& a, const Fe& b) const {
    Scope s(&fp_);
    return fe_add<FieldHost>(a, b);
  }
  Fe sub(const Fe& a, const Fe& b) const {
    Scope s(&fp_);
    return fe_sub<FieldHost>(a, b);
  }
  Fe neg(const Fe& a) const {
    Scope s(&fp_);
    return fe_neg<FieldHost>(a);
  }
  // standard-form value (must be < 2^256; reduced here) -> Montgomery form
  Fe to_mont(const Fe& std_val) const { return mul(reduce(std_val), r2_); }
  Fe from_mont(const Fe& m) const {
    Fe o = fe_zero();
    o.w[0] = 1;
    return mul(m, o);
  }
  Fe reduce(Fe v) const {  // v < 2^256 -> v mod p (p may be much smaller than 2^256)
    if (fe_cmp(v, p_) < 0) return v;
    // shift-subtract long reduction
    Fe r = fe_zero();
    for (int bit = 255; bit >= 0; bit--) {
      uint32_t d[8], s[8];
      uint32_t carry = add8(d, r.w, r.w);
      d[0] |= (v.w[bit / 32] >> (bit % 32)) & 1u;
      uint32_t borrow = sub8(s, d, p_.w);
      bool take = carry || !borrow;
      for (int j = 0; j < 8; j++) r.w[j] = take ? s[j] : d[j];
    }
    return r;
  }
  Fe pow(Fe base, const Fe& exp_std) const {
    Fe acc = one_;
    for (int bit = 255; bit >= 0; bit--) {
      acc = mul(acc, acc);
      if ((exp_std.w[bit / 32] >> (bit % 32)) & 1u) acc = mul(acc, base);
    }
    return acc;
  }
  Fe pow_u64(Fe base, uint64_t e) const {
    Fe acc = one_;
    while (e) {
      if (e & 1) acc = mul(acc, base);
      base = mul(base, base);
      e >>= 1;
    }
    return acc;
  }
  // Fermat inverse (p prime).  Returns zero for zero input; callers check.
  Fe inv(const Fe& a) const {
    Fe e = p_;
    Fe two = fe_zero();
    two.w[0] = 2;
    uint32_t tmp[8];
    sub8(tmp, e.w, two.w);
    memcpy(e.w, tmp, 32);
    return pow(a, e);
  }
  // Montgomery batch inversion; returns false if any element is zero.
  bool batch_inv(std::vector<Fe>& v) const {
    size_t n = v.size();
    if (n == 0) return true;
    std::vector<Fe> pre(n);
    Fe acc = one_;
    for (size_t i = 0; i < n; i++) {
      if (fe_is_zero(v[i])) return false;
      pre[i] = acc;
      acc = mul(acc, v[i]);
    }
    Fe ia = inv(acc);
    for (size_t i = n; i-- > 0;) {
      Fe t = mul(ia, pre[i]);
      ia = mul(ia, v[i]);
      v[i] = t;
    }
    return true;
  }
  Fe from_small(uint32_t v) const { return to_mont(fe_from_u32(v)); }

 private:
  FieldParams fp_;
  Fe one_, r2_, p_;
};

// prod (X - x_i), monic, Montgomery-form coefficients, length k+1.
inline std::vector<Fe> build_from_roots(const HostField& f, const std::vector<Fe>& xs) {
  std::vector<Fe> a(1, f.one());
  for (const Fe& x : xs) {
    std::vector<Fe> nxt(a.size() + 1, fe_zero());
    for (size_t i = 0; i < a.size(); i++) {
      nxt[i + 1] = f.add(nxt[i + 1], a[i]);
      nxt[i] = f.sub(nxt[i], f.mul(a[i], x));
    }
    a.swap(nxt);
  }
  return a;
}

// In-place radix-2 NTT on the host (Montgomery-form values): a[i] <- sum_j a[j] w^(ij),
// len(a) = n a power of two, w a primitive n-th root (Montgomery form).  Constants only.
inline void host_ntt(const HostField& f, std::vector<Fe>& a, const Fe& w) {
  const size_t n = a.size();
  for (size_t i = 1, j = 0; i < n; i++) {  // bit reversal
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(a[i], a[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const Fe wl = f.pow_u64(w, (uint64_t)(n / len));
    for (size_t i = 0; i < n; i += len) {
      Fe cur = f.one();
      for (size_t j = 0; j < len / 2; j++) {
        const Fe u = a[i + j], v = f.mul(a[i + j + len / 2], cur);
        a[i + j] = f.add(u, v);
        a[i + j + len / 2] = f.sub(u, v);
        cur = f.mul(cur, wl);
      }
    }
  }
}

inline Fe horner(const HostField& f, const std::vector<Fe>& a, const Fe& x) {
  Fe acc = fe_zero();
  for (size_t i = a.size(); i-- > 0;) acc = f.add(f.mul(acc, x), a[i]);
  return acc;
}

// Row-major k x k inverse of the Vandermonde matrix V[i][j] = x_i^j
// (rsdecode_impl.h:97-122 computes the same matrix with NTL's generic inv()):
// column i of V^-1 holds the coefficients of the Lagrange basis polynomial
// L_i(X) = A(X) / ((X - x_i) A'(x_i)), A = prod (X - x_j).  O(k^2).
// Returns false when two points coincide (singular, det = 0).
inline bool vandermonde_inverse(const HostField& f, const std::vector<Fe>& xs,
                                std::vector<Fe>& out) {
  size_t k = xs.size();
  out.assign(k * k, fe_zero());
  if (k == 0) return true;
  std::vector<Fe> a = build_from_roots(f, xs);
  std::vector<Fe> q(k * k);     // q[i*k + j]: coefficient j of A/(X - x_i)
  std::vector<Fe> denom(k);
  for (size_t i = 0; i < k; i++) {
    Fe carry = fe_zero();
    for (size_t j = k; j >= 1; j--) {
      carry = f.add(a[j], f.mul(carry, xs[i]));
      q[i * k + (j - 1)] = carry;
    }
    std::vector<Fe> qi(q.begin() + i * k, q.begin() + (i + 1) * k);
    denom[i] = horner(f, qi, xs[i]);
  }
  if (!f.batch_inv(denom)) return false;
  for (size_t i = 0; i < k; i++)
    for (size_t j = 0; j < k; j++) out[j * k + i] = f.mul(q[i * k + j], denom[i]);
  return true;
}

}  // namespace hb
