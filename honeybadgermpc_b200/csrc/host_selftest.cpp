// Host twins of the device field arithmetic, exported for the CPU unit tests
// (tests/test_host_field.py).  Not used by the product path: the batch data
// path is CUDA only.  Built by __graft_entry__.build_host_selftest() with g++.
#include <vector>

#include "host_math.hpp"
#include "rowmath.cuh"

using namespace hb;

namespace {
template <class Pol>
void mulmod_with(const FieldParams& fp, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  HostField f(fp);
  HostField::Scope s(&fp);
  Fe r2;
  memcpy(r2.w, fp.r2, 32);
  Fe x = mont_mul<Pol>(f.reduce(fe_from_u64(a)), r2);   // a*R
  Fe y = f.reduce(fe_from_u64(b));                     // b (standard)
  Fe r = mont_mul<Pol>(y, x);                          // a*b, standard form
  fe_to_u64(r, out);
}
}  // namespace

extern "C" {

// out = a*b mod p through mont_mul (generic reduction rows)
int hbt_mulmod(const uint64_t* p, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  mulmod_with<FieldHost>(fp, a, b, out);
  return 0;
}

// same through the p[0]=1, p[1]=2^32-1 fast reduction rows (BLS12-381 r only)
int hbt_mulmod_lowones(const uint64_t* p, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  if (fp.p[0] != 1u || fp.p[1] != 0xffffffffu) return 2;
  mulmod_with<FieldHostLowOnes>(fp, a, b, out);
  return 0;
}

int hbt_addsub(const uint64_t* p, const uint64_t* a, const uint64_t* b, uint64_t* sum,
               uint64_t* diff, uint64_t* neg) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  HostField f(fp);
  fe_to_u64(f.add(fe_from_u64(a), fe_from_u64(b)), sum);
  fe_to_u64(f.sub(fe_from_u64(a), fe_from_u64(b)), diff);
  fe_to_u64(f.neg(fe_from_u64(a)), neg);
  return 0;
}

// out = sum_j a[j]*b[j] mod p with the lazy accumulator (acc_mac / acc_fold /
// acc_redc), folding every `fold` macs.  a standard form, b converted to
// Montgomery form here -- exactly what apply_matrix_kernel does.
int hbt_dot(const uint64_t* p, int n, const uint64_t* a, const uint64_t* b, int fold,
            uint64_t* out) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  HostField f(fp);
  HostField::Scope s(&fp);
  Acc acc;
  acc_zero(acc);
  int pending = 0;
  for (int j = 0; j < n; j++) {
    Fe x = fe_from_u64(a + 4 * j);
    Fe m = f.to_mont(fe_from_u64(b + 4 * j));
    acc_mac(acc, x, m);
    if (++pending == fold) {
      acc_fold<FieldHost>(acc);
      pending = 0;
    }
  }
  if (pending) acc_fold<FieldHost>(acc);
  fe_to_u64(acc_redc<FieldHost>(acc), out);
  return 0;
}

// Row-major k x k inverse Vandermonde matrix (standard form); returns 2 if singular.
int hbt_vandermonde_inverse(const uint64_t* p, int k, const uint64_t* xs, uint64_t* out) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  HostField f(fp);
  std::vector<Fe> x(k), inv;
  for (int i = 0; i < k; i++) x[i] = f.to_mont(fe_from_u64(xs + 4 * i));
  if (!vandermonde_inverse(f, x, inv)) return 2;
  for (int i = 0; i < k * k; i++) fe_to_u64(f.from_mont(inv[i]), out + 4 * i);
  return 0;
}

int hbt_pow_inv(const uint64_t* p, const uint64_t* a, uint64_t e, uint64_t* pw, uint64_t* inv) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  HostField f(fp);
  Fe am = f.to_mont(fe_from_u64(a));
  fe_to_u64(f.from_mont(f.pow_u64(am, e)), pw);
  fe_to_u64(f.from_mont(f.inv(am)), inv);
  return 0;
}

// 16-point NTT of d <= 8 coefficients through ntt16_half (both parities), the row math of
// ntt16_g4_kernel.  omega must be a primitive 16th root of unity mod p.
int hbt_ntt16(const uint64_t* p, int d, const uint64_t* coeffs, const uint64_t* omega, int lowones,
              int bal, uint64_t* out) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  if (d < 0 || d > 8) return 3;
  HostField f(fp);
  HostField::Scope s(&fp);
  Fe tw[16];
  tw[0] = f.one();
  Fe w = f.to_mont(fe_from_u64(omega));
  for (int i = 1; i < 16; i++) tw[i] = f.mul(tw[i - 1], w);
  auto ld = [&](int i) { return i < d ? fe_from_u64(coeffs + 4 * i) : fe_zero(); };
  auto st = [&](int k, const Fe& v) { fe_to_u64(v, out + 4 * k); };
  struct HostXch {  // the even thread's hand-over to the odd thread (shared memory on the device)
    Fe slot[4];
    void put(int i, const Fe& v) { slot[i] = v; }
    Fe get(int i) const { return slot[i]; }
    void signal() {}
    void wait() {}
  } xch;
  auto run = [&](auto pol) {
    using Pol = decltype(pol);
    auto twf = [&](int j, const Fe& x) { return mont_mul<Pol>(tw[j], x); };
    if (d <= 4) {
      if (bal) ntt16_half<Pol, 4, 0, true>(ld, twf, st, xch); else ntt16_half<Pol, 4, 0, false>(ld, twf, st, xch);
      if (bal) ntt16_half<Pol, 4, 1, true>(ld, twf, st, xch); else ntt16_half<Pol, 4, 1, false>(ld, twf, st, xch);
    } else if (d <= 6) {
      if (bal) ntt16_half<Pol, 6, 0, true>(ld, twf, st, xch); else ntt16_half<Pol, 6, 0, false>(ld, twf, st, xch);
      if (bal) ntt16_half<Pol, 6, 1, true>(ld, twf, st, xch); else ntt16_half<Pol, 6, 1, false>(ld, twf, st, xch);
    } else {
      if (bal) ntt16_half<Pol, 8, 0, true>(ld, twf, st, xch); else ntt16_half<Pol, 8, 0, false>(ld, twf, st, xch);
      if (bal) ntt16_half<Pol, 8, 1, true>(ld, twf, st, xch); else ntt16_half<Pol, 8, 1, false>(ld, twf, st, xch);
    }
  };
  if (lowones) {
    if (fp.p[0] != 1u || fp.p[1] != 0xffffffffu) return 2;
    run(FieldHostLowOnes{});
  } else {
    run(FieldHost{});
  }
  return 0;
}

// out = sum_j a[j]*b[j] mod p through the radix-2^29 carry-free accumulator (to_limbs29 /
// mac29 / norm29 / redc29): a standard form, b converted to limbs of b * 2^261 here --
// exactly what interp_small_kernel does.  prenorm != 0 runs norm29 before the reduction.
int hbt_dot29(const uint64_t* p, int n, const uint64_t* a, const uint64_t* b, int lowones,
              int prenorm, uint64_t* out) {
  FieldParams fp;
  if (!field_params_init(p, &fp)) return 1;
  if (n > 8) return 3;
  HostField f(fp);
  HostField::Scope s(&fp);
  Fe r261;
  memcpy(r261.w, fp.r261, 32);
  uint64_t col[17];
  for (int c = 0; c < 17; c++) col[c] = 0;
  for (int j = 0; j < n; j++) {
    uint32_t la[9], lb[9];
    to_limbs29(fe_from_u64(a + 4 * j), la);
    to_limbs29(f.mul(f.to_mont(fe_from_u64(b + 4 * j)), r261), lb);
    mac29(col, la, lb);
  }
  if (prenorm) norm29(col);
  Fe r;
  if (lowones) {
    if (fp.p29[0] != 1u) return 2;
    r = redc29<FieldHostLowOnes>(col);
  } else {
    r = redc29<FieldHost>(col);
  }
  fe_to_u64(r, out);
  return 0;
}

}  // extern "C"
