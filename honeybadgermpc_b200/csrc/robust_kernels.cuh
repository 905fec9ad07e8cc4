// Robust (error-correcting) Reed-Solomon decoders on the GPU.
//
//  * gao_kernel:  Gao's decoder, one received word per warp, polynomials in
//    shared memory, lanes strided over coefficients.  Follows
//    rsdecode_impl.h:281-363 (partial_gcd + gao_interpolate): the extended
//    Euclid sequence starts from (g0, g1), stops at the first remainder of
//    degree < (m+k)/2, and the returned locator is the UN-normalised Bezout
//    cofactor.  Field inversions are avoided inside the loop: every division
//    step is a pseudo-division (rows scaled by the divisor's leading
//    coefficient), the common scale alpha of (r_j, t_j) is carried along and
//    removed with ONE inversion at the end, which yields exactly the
//    reference's (g, v).
//  * wb_kernel:   Welch-Berlekamp, one received word per CTA, the linear system
//    in shared memory.  Follows reed_solomon_wb.py:79-127 (solve_system),
//    :157-197 (rref), :240-273 (some_solution) including its syntactic
//    pivot-column test.  Elimination is fraction free (rows are scaled instead
//    of normalised -- same zero pattern, hence same pivot / free columns as
//    the reference's RREF); the pivots are inverted once, in parallel.
//
// All polynomial / matrix entries inside these kernels are in Montgomery form.
#pragma once
#include "kernels.cuh"

namespace hb {

template <class F>
HB_D Fe fe_one_mont() {
  // R mod p: for FieldBLS computed from the immediates would need a table; both
  // policies read it from the constant bank (bound for every robust launch).
  Fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.w[i] = c_field.one[i];
  return r;
}

HB_D Fe fe_small(uint32_t v) {
  Fe r = fe_zero();
  r.w[0] = v;
  return r;
}

// a*b - c*d  (all Montgomery form in, Montgomery form out), one reduction
template <class F>
HB_D Fe mul_sub(const Fe& a, const Fe& b, const Fe& neg_c, const Fe& d) {
  Acc acc;
  acc_zero(acc);
  acc_mac(acc, b, a);
  if (F::kFold == 1) acc_fold<F>(acc);
  acc_mac(acc, d, neg_c);
  acc_fold<F>(acc);
  return acc_redc<F>(acc);
}

// x^(p-2); uniform control flow (exponent = modulus - 2 from the constant bank)
template <class F>
HB_D Fe fe_inv(const Fe& x) {
  uint32_t e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = c_field.p[i];
  // p is odd and >= 3: p - 2 never borrows beyond word 0 unless p[0] == 1
  uint32_t borrow = e[0] < 2u;
  e[0] -= 2u;
  for (int i = 1; i < 8 && borrow; i++) {
    borrow = e[i] == 0;
    e[i] -= 1u;
  }
  Fe acc = fe_one_mont<F>();
  bool started = false;
  for (int bit = 255; bit >= 0; bit--) {
    if (started) acc = mont_mul<F>(acc, acc);
    if ((e[bit >> 5] >> (bit & 31)) & 1u) {
      acc = started ? mont_mul<F>(acc, x) : x;
      started = true;
    }
  }
  return acc;
}

struct Planes {
  uint4* lo;
  uint4* hi;
  HB_D Fe get(int j) const { return lds_fe(lo, hi, j); }
  HB_D void set(int j, const Fe& v) const { sts_fe(lo, hi, j, v); }
};

// ---------------------------------------------------------------------------
// Gao
// ---------------------------------------------------------------------------
struct GaoArgs {
  const uint4* g0;      // [m+1] prod (X - x_i), Montgomery form
  const uint4* g1;      // [batch][m] interpolants of the received words, Montgomery form
  uint4* coeffs;        // [batch][k]  decoded message, standard form, zero padded
  uint4* locator;       // [batch][loc_stride] un-normalised error locator, standard form
  int* loc_len;         // [batch] deg(v) + 1
  int* status;          // [batch] 0 decoded, 1 failed (the reference returns (None, None))
  unsigned long long batch;
  int m, k, thr, loc_stride, warps_per_cta;
};

// highest j < bound with a[j] != 0, or -1 (warp-uniform result)
HB_D int warp_degree(const Planes& a, int bound, int lane) {
  int best = -1;
  for (int j = lane; j < bound; j += 32) {
    Fe v = a.get(j);
    if (!fe_is_zero(v)) best = j;
  }
  return __reduce_max_sync(0xffffffffu, best);
}

template <class F>
__global__ void __launch_bounds__(256) gao_kernel(GaoArgs a) {
  extern __shared__ uint4 smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= a.warps_per_cta) return;
  const int m = a.m, len = m + 1;
  uint4* base = smem + (size_t)warp * 8 * len;
  Planes r0{base, base + len}, r1{base + 2 * len, base + 3 * len};
  Planes t0{base + 4 * len, base + 5 * len}, t1{base + 6 * len, base + 7 * len};
  const Fe one = fe_one_mont<F>();
  const Fe std_one = fe_small(1);

  for (unsigned long long row = (unsigned long long)blockIdx.x * a.warps_per_cta + warp;
       row < a.batch; row += (unsigned long long)gridDim.x * a.warps_per_cta) {
    __syncwarp();
    for (int j = lane; j < len; j += 32) {
      r0.set(j, ld_fe(a.g0 + 2 * j));
      r1.set(j, j < m ? ld_fe(a.g1 + 2ull * (row * m + j)) : fe_zero());
      t0.set(j, fe_zero());
      t1.set(j, j == 0 ? one : fe_zero());
    }
    __syncwarp();
    int d0 = m, d1 = warp_degree(r1, m, lane);
    int b0 = -1, b1 = 0;           // degree bounds of t0, t1
    Fe alpha0 = one, alpha1 = one;  // common scale of (r0,t0) and of (r1,t1)
    int st = 0;
    bool trivial = false;
    uint4* out_c = a.coeffs + 2ull * row * a.k;
    uint4* out_l = a.locator + 2ull * row * a.loc_stride;

    if (d0 < a.thr) {
      st = 1;  // v = 0 (only when k > m): DivRem by zero in the reference
    } else if (d1 < a.thr) {
      trivial = true;  // no errors: (g, v) = (g1, 1), rsdecode_impl.h:296-301
    } else {
      while (true) {
        // pseudo-divide r0 by r1: delta+1 row operations r0 <- L*r0 - c*X^s*r1
        const Fe lead = r1.get(d1);
        for (int s = d0 - d1; s >= 0; s--) {
          const Fe c = r0.get(d1 + s);
          const Fe negc = fe_neg<F>(c);
          __syncwarp();
          for (int j = lane; j <= d1 + s; j += 32) {
            Fe x = r0.get(j);
            Fe v = j >= s ? mul_sub<F>(lead, x, negc, r1.get(j - s)) : mont_mul<F>(x, lead);
            r0.set(j, v);
          }
          int nb = b1 + s > b0 ? b1 + s : b0;
          for (int j = lane; j <= nb; j += 32) {
            Fe x = j <= b0 ? t0.get(j) : fe_zero();
            Fe y = (j >= s && j - s <= b1) ? t1.get(j - s) : fe_zero();
            t0.set(j, mul_sub<F>(lead, x, negc, y));
          }
          b0 = nb;
          alpha0 = mont_mul<F>(alpha0, lead);
          __syncwarp();
        }
        int d2 = warp_degree(r0, d1, lane);
        // rotate: (r0,t0) <- (r1,t1), (r1,t1) <- remainder
        Planes tp = r0; r0 = r1; r1 = tp;
        tp = t0; t0 = t1; t1 = tp;
        Fe ta = alpha0; alpha0 = alpha1; alpha1 = ta;
        int tb = b0; b0 = b1; b1 = tb;
        d0 = d1;
        d1 = d2;
        if (d1 < a.thr) break;
      }
    }

    if (st == 0 && trivial) {
      if (d1 >= a.k) {
        st = 1;
      } else {
        for (int j = lane; j < a.k; j += 32)
          st_fe(out_c + 2 * j, j <= d1 ? mont_mul<F>(r1.get(j), std_one) : fe_zero());
        if (lane == 0) {
          st_fe(out_l, std_one);
          a.loc_len[row] = 1;
        }
      }
    } else if (st == 0) {
      // (g, v) = (r1, t1) / alpha1
      const int dg = d1;
      const int dv = warp_degree(t1, b1 + 1, lane);
      if (dv < 0) {
        st = 1;
      } else {
        const Fe lv = t1.get(dv);
        const Fe inv = fe_inv<F>(mont_mul<F>(alpha1, lv));
        const Fe inv_alpha_std = mont_mul<F>(mont_mul<F>(inv, lv), std_one);
        const Fe inv_lv = mont_mul<F>(inv, alpha1);
        const Fe inv_lv_std = mont_mul<F>(inv_lv, std_one);
        for (int j = lane; j <= dv; j += 32) st_fe(out_l + 2 * j, mont_mul<F>(t1.get(j), inv_alpha_std));
        if (lane == 0) a.loc_len[row] = dv + 1;
        // f = g / v must be exact and of degree < k
        if (dg < 0) {
          for (int j = lane; j < a.k; j += 32) st_fe(out_c + 2 * j, fe_zero());
        } else if (dg < dv || dg - dv >= a.k) {
          st = 1;
        } else {
          const int delta = dg - dv;
          for (int s = delta; s >= 0; s--) {
            const Fe c = r1.get(dv + s);
            const Fe q = mont_mul<F>(c, inv_lv);
            if (lane == 0) st_fe(out_c + 2 * s, mont_mul<F>(c, inv_lv_std));
            const Fe negq = fe_neg<F>(q);
            __syncwarp();
            for (int j = lane; j <= dv; j += 32) {
              Fe x = r1.get(j + s);
              r1.set(j + s, fe_add<F>(x, mont_mul<F>(t1.get(j), negq)));
            }
            __syncwarp();
          }
          if (warp_degree(r1, dv, lane) >= 0) st = 1;  // remainder != 0
          for (int j = delta + 1 + lane; j < a.k; j += 32) st_fe(out_c + 2 * j, fe_zero());
        }
      }
    }
    if (lane == 0) a.status[row] = st;
  }
}

// ---------------------------------------------------------------------------
// Welch-Berlekamp
// ---------------------------------------------------------------------------
struct WbArgs {
  const uint4* pw;      // [m][pw_stride] a_i^j, Montgomery form, j < e_max + k
  const uint4* ys;      // [batch][m] received values, standard form
  uint4* coeffs;        // [batch][k] decoded message, standard form, zero padded
  int* out_len;         // [batch] length of the stripped coefficient list
  int* status;          // [batch] 0 decoded, 1 "found no divisors!", 2 "No solution", 3 E == 0
  uint4* work;          // per-CTA global workspace when the system exceeds shared memory
  unsigned long long batch;
  unsigned long long work_stride;  // uint4 per CTA
  int m, k, e_max, pw_stride;
  int in_smem;
};

constexpr int kWbThreads = 512;  // 16 warps per CTA: one CTA per SM when the system fills shared memory

template <class F>
__global__ void __launch_bounds__(kWbThreads) wb_kernel(WbArgs a) {
  extern __shared__ uint4 smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int m = a.m, k = a.k;
  const int nrows = m + 1;
  const int max_cols = 2 * a.e_max + k + 2;
  // shared scratch: factor column, per-variable values, bookkeeping
  uint4* sp = smem;
  Planes fcol{sp, sp + nrows};
  sp += 2 * nrows;
  Planes val{sp, sp + max_cols};   // solution vector / numerators
  sp += 2 * max_cols;
  Planes bm{sp, sp + m};           // received values in Montgomery form
  sp += 2 * m;
  int* ip = (int*)sp;
  int* piv_row = ip;               // [max_cols] pivot row of a column, -1 free, -2 pseudo pivot
  int* act = ip + max_cols;        // [max_cols] columns touched by the current row operation
  int* pseudo_row = act + max_cols;  // [max_cols]
  int* misc = pseudo_row + max_cols;  // [8]
  int* row_pc = misc + 8;             // [nrows] pivot column of a pivot row
  sp = (uint4*)(ip + ((3 * max_cols + 8 + nrows + 3) & ~3));
  Planes M;
  if (a.in_smem) {
    M = Planes{sp, sp + (size_t)nrows * max_cols};
  } else {
    uint4* g = a.work + (size_t)blockIdx.x * a.work_stride;
    M = Planes{g, g + (size_t)nrows * max_cols};
  }
  const Fe one = fe_one_mont<F>();
  const Fe std_one = fe_small(1);
  Fe r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.w[i] = c_field.r2[i];

  for (unsigned long long row = blockIdx.x; row < a.batch; row += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < m; i += kWbThreads) bm.set(i, mont_mul<F>(ld_fe(a.ys + 2ull * (row * m + i)), r2));
    int result = 1;  // "found no divisors!" unless some e works
    int out_len = 0;
    for (int e = a.e_max; e >= 1; e--) {
      const int ncols = 2 * e + k + 2, rhs = ncols - 1, nvars = ncols - 1;
      __syncthreads();
      // ---- build the system (reed_solomon_wb.py:92-102)
      for (int idx = tid; idx < nrows * ncols; idx += kWbThreads) {
        int r = idx / ncols, c = idx - r * ncols;
        Fe v = fe_zero();
        if (r < m) {
          if (c <= e) v = mont_mul<F>(bm.get(r), ld_fe(a.pw + 2ull * ((size_t)r * a.pw_stride + c)));
          else if (c < rhs) v = fe_neg<F>(ld_fe(a.pw + 2ull * ((size_t)r * a.pw_stride + (c - e - 1))));
        } else if (c == e || c == rhs) {
          v = one;
        }
        M.set(idx, v);
      }
      for (int c = tid; c < ncols; c += kWbThreads) piv_row[c] = -1;
      __syncthreads();
      // ---- fraction-free Gauss-Jordan over all columns incl. the constants (rref, :157-197)
      int prow = 0, nfree = 0;
      for (int col = 0; col < ncols && prow < nrows; col++) {
        if (tid == 0) misc[0] = nrows;
        __syncthreads();
        for (int r = prow + tid; r < nrows; r += kWbThreads)
          if (!fe_is_zero(M.get(r * ncols + col))) atomicMin(&misc[0], r);
        __syncthreads();
        const int pr = misc[0];
        if (pr == nrows) {  // free column
          if (tid == 0) act[nfree] = col;
          nfree++;
          __syncthreads();
          continue;
        }
        if (pr != prow) {
          for (int c = tid; c < ncols; c += kWbThreads) {
            Fe x = M.get(prow * ncols + c), y = M.get(pr * ncols + c);
            M.set(prow * ncols + c, y);
            M.set(pr * ncols + c, x);
          }
        }
        if (tid == 0) {
          piv_row[col] = prow;
          row_pc[prow] = col;
        }
        __syncthreads();
        for (int r = tid; r < nrows; r += kWbThreads) fcol.set(r, M.get(r * ncols + col));
        // columns that can change: free columns seen so far and everything from col on
        for (int c = col + tid; c < ncols; c += kWbThreads) act[nfree + (c - col)] = c;
        __syncthreads();
        // every other row q becomes piv*row_q - f_q*row_prow.  Columns that can change:
        // the free columns seen so far, everything from `col` on, and -- for the
        // pivot rows above -- the row's own pivot entry (scaled by piv; the pivot
        // row is zero there).  Earlier pivot columns are zero in both rows.
        const int cnt = nfree + (ncols - col) + 1;
        const Fe piv = fcol.get(prow);
        for (int idx = tid; idx < nrows * cnt; idx += kWbThreads) {
          int r = idx / cnt, ci = idx - r * cnt;
          if (r == prow) continue;
          Fe f = fcol.get(r);
          if (fe_is_zero(f)) continue;
          if (ci == cnt - 1) {
            if (r < prow) {
              int c = row_pc[r];
              M.set(r * ncols + c, mont_mul<F>(M.get(r * ncols + c), piv));
            }
            continue;
          }
          int c = act[ci];
          Fe v = mul_sub<F>(piv, M.get(r * ncols + c), fe_neg<F>(f), M.get(prow * ncols + c));
          M.set(r * ncols + c, v);
        }
        prow++;
        __syncthreads();
      }
      // columns never visited because the rows ran out are free as well
      __syncthreads();
      // ---- some_solution (:240-273)
      if (piv_row[rhs] >= 0) {  // a row 0 ... 0 | c: "No solution" (not caught by the caller)
        result = 2;
        break;
      }
      // syntactic pivot test of is_pivot_column for the free columns: exactly one
      // non-zero entry, equal (after normalisation) to 1
      for (int c = tid; c < nvars; c += kWbThreads) {
        pseudo_row[c] = -1;
        if (piv_row[c] >= 0) continue;
        int hits = 0, at = -1;
        for (int r = 0; r < prow; r++)
          if (!fe_is_zero(M.get(r * ncols + c))) { hits++; at = r; }
        if (hits == 1) {
          // the pivot entry of row `at`
          int pc = 0;
          while (piv_row[pc] != at) pc++;
          if (fe_eq(M.get(at * ncols + c), M.get(at * ncols + pc))) pseudo_row[c] = at;
        }
      }
      __syncthreads();
      // value of every variable
      for (int c = tid; c < nvars; c += kWbThreads) {
        int r = piv_row[c] >= 0 ? piv_row[c] : pseudo_row[c];
        if (r < 0) {
          val.set(c, one);  // free variable := 1
          continue;
        }
        Fe num = M.get(r * ncols + rhs);
        for (int f = 0; f < nvars; f++)
          if (piv_row[f] < 0 && pseudo_row[f] < 0) num = fe_sub<F>(num, M.get(r * ncols + f));
        int pc = c;
        if (piv_row[c] < 0) {
          pc = 0;
          while (piv_row[pc] != r) pc++;
        }
        Fe den = M.get(r * ncols + pc);
        val.set(c, mont_mul<F>(num, fe_inv<F>(den)));
      }
      __syncthreads();
      // ---- Q mod E == 0 ?  (E = val[0..e], Q = val[e+1..]; polynomial.py:219-234)
      if (tid < 32) {
        Planes Q{val.lo + e + 1, val.hi + e + 1};
        const int de = warp_degree(val, e + 1, lane);
        const int dq = warp_degree(Q, e + k, lane);
        int ok = 1, plen = 0;
        if (de < 0) {
          ok = 2;  // division by the zero polynomial
        } else if (dq >= 0) {
          if (dq < de) {
            ok = 0;
          } else {
            const Fe lead = val.get(de);
            const Fe inv_lead = fe_eq(lead, one) ? one : fe_inv<F>(lead);
            for (int s = dq - de; s >= 0; s--) {
              const Fe q = mont_mul<F>(Q.get(de + s), inv_lead);
              const Fe negq = fe_neg<F>(q);
              __syncwarp();
              for (int j = lane; j < de; j += 32)
                Q.set(j + s, fe_add<F>(Q.get(j + s), mont_mul<F>(val.get(j), negq)));
              if (lane == 0) Q.set(de + s, q);
              __syncwarp();
            }
            if (warp_degree(Q, de, lane) >= 0) ok = 0;
            plen = dq - de + 1;
          }
        }
        if (lane == 0) {
          misc[1] = ok;
          misc[2] = plen;
        }
        if (ok == 1) {
          uint4* out_c = a.coeffs + 2ull * row * k;
          for (int j = lane; j < k; j += 32)
            st_fe(out_c + 2 * j, j < plen ? mont_mul<F>(Q.get(de + j), std_one) : fe_zero());
        }
      }
      __syncthreads();
      if (misc[1] == 1) {
        result = 0;
        out_len = misc[2];
        break;
      }
      if (misc[1] == 2) {
        result = 3;
        break;
      }
    }
    if (tid == 0) {
      a.status[row] = result;
      a.out_len[row] = out_len;
    }
  }
}

// ---------------------------------------------------------------------------
// glue of the Welch-Berlekamp unique-decoding shortcut (hbg_wb_decode_batch)
// ---------------------------------------------------------------------------
// out_len[b] = length of coeffs[b] with trailing zeros stripped (polynomial.py:36), for the words the
// Gao kernel decoded (status 0); the others keep status 1 until the exact kernel has seen them.
__global__ void __launch_bounds__(256) wb_strip_kernel(const uint4* coeffs, const int* status, int* out_len,
                                                       unsigned long long batch, int k) {
  for (unsigned long long b = (unsigned long long)blockIdx.x * 256 + threadIdx.x; b < batch;
       b += (unsigned long long)gridDim.x * 256) {
    int len = 0;
    if (status[b] == 0) {
      for (int j = k - 1; j >= 0; j--) {
        const uint4 lo = coeffs[2ull * (b * (unsigned)k + j)], hi = coeffs[2ull * (b * (unsigned)k + j) + 1];
        if (lo.x | lo.y | lo.z | lo.w | hi.x | hi.y | hi.z | hi.w) {
          len = j + 1;
          break;
        }
      }
    }
    out_len[b] = len;
  }
}

// dst[i] = src[idx[i]] (rows of `chunks` uint4)
__global__ void __launch_bounds__(256) rows_gather_kernel(const uint4* src, uint4* dst, const int* idx,
                                                          unsigned long long n, int chunks) {
  const unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= n * (unsigned)chunks) return;
  const unsigned long long i = t / (unsigned)chunks;
  const int c = (int)(t - i * (unsigned)chunks);
  dst[t] = src[(unsigned long long)idx[i] * (unsigned)chunks + c];
}

// dst[idx[i]] = src[i], plus the per-row out_len / status
__global__ void __launch_bounds__(256) rows_scatter_kernel(const uint4* src, uint4* dst, const int* idx,
                                                           unsigned long long n, int chunks, const int* len_src,
                                                           const int* st_src, int* len_dst, int* st_dst) {
  const unsigned long long t = (unsigned long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= n * (unsigned)chunks) return;
  const unsigned long long i = t / (unsigned)chunks;
  const int c = (int)(t - i * (unsigned)chunks);
  dst[(unsigned long long)idx[i] * (unsigned)chunks + c] = src[t];
  if (c == 0) {
    len_dst[idx[i]] = len_src[i];
    st_dst[idx[i]] = st_src[i];
  }
}

}  // namespace hb