// Host twins of the device field arithmetic, exported for the CPU unit tests
// (tests/test_host_field.py).  Not used by the product path: the batch data
// path is CUDA only.  Built by __graft_entry__.build_host_selftest() with g++.
#include <vector>

#include "host_math.hpp"

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

}  // extern "C"
