"""CPU oracle for the HoneyBadgerMPC share-reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``honeybadgermpc_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg use it, and there only as the
checker / the reported CPU number.

It is a plain-Python-int restatement of the algorithms of the reference's
``honeybadgermpc.ntl`` extension (Cython ``hbmpc_ntl_helpers.pyx`` on top of the
C++ header ``rsdecode_impl.h``, which in turn calls the un-vendored third-party
library NTL, version pinned only by the reference's docker image digest,
``Dockerfile:1-3``), plus ``reed_solomon_wb.py`` (Welch-Berlekamp) and
``polynomial.py`` ``EvalPoint`` / ``get_omega``.  All paths below are relative
to ``/root/reference/honeybadgermpc``.

Parity status: PINNED by the reference's own known-answer tests
(``tests/test_ntl.py``, ``tests/test_reed_solomon.py``,
``tests/test_reed_solomon_wb.py``, ``tests/test_batch_reconstruction.py``;
see ``tests/test_oracle_kats.py``) and, for the pure-Python parts of the
reference that can be imported in the authoring container (Welch-Berlekamp,
``EvalPoint``, the Python FFT / ``fnt_decode``), by golden vectors generated
from the reference itself (``tests/golden/make_golden.py``).  The NTL binary
cannot be built here (no NTL/GMP headers, no network); because every output is
the canonical residue of an exact field computation the results are
implementation independent, except where noted (``sqrt_mod`` root choice).

Conventions: polynomials are coefficient lists, lowest degree first; the zero
polynomial is ``[]`` internally; ``deg([]) == -1`` like NTL.
"""

from random import Random

# --------------------------------------------------------------------------
# marshalling helpers (ntl/hbmpc_ntl_helpers.pyx:20-70)
# --------------------------------------------------------------------------


def _to_int(v):
    """py_obj_to_ZZ (pyx:37-46): ints pass, str/bytes are read as decimal,
    None raises ValueError.  Negative ints raise OverflowError, as
    ``int.to_bytes`` does in intToZZ (pyx:20-22)."""
    if isinstance(v, bool):
        v = int(v)
    if isinstance(v, int):
        if v < 0:
            raise OverflowError("can't convert negative int to unsigned")
        return v
    if v is None:
        raise ValueError(f"Unsupported data type. {type(v)}")
    if isinstance(v, bytes):
        v = v.decode()
    return int(str(v).strip())


def _el(v, p):
    """intToZZp (pyx:31-32): reduce into [0, p)."""
    if not isinstance(v, int):
        raise TypeError(f"expected int, got {type(v)}")
    if v < 0:
        raise OverflowError("can't convert negative int to unsigned")
    return v % p


def _inv(a, p):
    a %= p
    if a == 0:
        raise ZeroDivisionError("inverse of 0 (NTL would abort the process)")
    return pow(a, -1, p)


# --------------------------------------------------------------------------
# polynomial helpers standing in for NTL ZZ_pX
# --------------------------------------------------------------------------


def _strip(a):
    a = list(a)
    while a and a[-1] == 0:
        a.pop()
    return a


def _deg(a):
    return len(_strip(a)) - 1


def poly_mul(a, b, p):
    a, b = _strip(a), _strip(b)
    if not a or not b:
        return []
    out = [0] * (len(a) + len(b) - 1)
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                out[i + j] = (out[i + j] + ai * bj) % p
    return _strip(out)


def poly_sub(a, b, p):
    n = max(len(a), len(b))
    a = list(a) + [0] * (n - len(a))
    b = list(b) + [0] * (n - len(b))
    return _strip([(x - y) % p for x, y in zip(a, b)])


def poly_divrem(a, b, p):
    """NTL DivRem(q, r, a, b): a = q*b + r, deg r < deg b."""
    a, b = _strip(a), _strip(b)
    if not b:
        raise ZeroDivisionError("DivRem by zero polynomial")
    if len(a) < len(b):
        return [], a
    inv_lc = _inv(b[-1], p)
    r = list(a)
    q = [0] * (len(a) - len(b) + 1)
    for i in range(len(a) - len(b), -1, -1):
        c = (r[i + len(b) - 1] * inv_lc) % p
        q[i] = c
        if c:
            for j, bj in enumerate(b):
                r[i + j] = (r[i + j] - c * bj) % p
    return _strip(q), _strip(r[: len(b) - 1])


def build_from_roots(xs, p):
    """NTL BuildFromRoots: prod (X - x_i), monic."""
    a = [1]
    for x in xs:
        nxt = [0] * (len(a) + 1)
        for i, c in enumerate(a):
            nxt[i + 1] = (nxt[i + 1] + c) % p
            nxt[i] = (nxt[i] - c * x) % p
        a = nxt
    return a


def poly_eval(a, x, p):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % p
    return acc


def _interpolate(xs, ys, p):
    """NTL interpolate(P, a, b): unique P of degree < len(xs).  Stripped."""
    k = len(xs)
    if k == 0:
        return []
    a = build_from_roots(xs, p)
    out = [0] * k
    for i in range(k):
        # synthetic division A / (X - x_i)
        q = [0] * k
        carry = 0
        for j in range(k, 0, -1):
            carry = (a[j] + carry * xs[i]) % p
            q[j - 1] = carry
        denom = poly_eval(q, xs[i], p)
        s = (ys[i] * _inv(denom, p)) % p
        if s:
            for j in range(k):
                out[j] = (out[j] + s * q[j]) % p
    return _strip(out)


# --------------------------------------------------------------------------
# exported functions of honeybadgermpc.ntl
# --------------------------------------------------------------------------


class InterpolationError(Exception):
    """ntl/hbmpc_ntl_helpers.pyx:135"""


def lagrange_interpolate(x, y, modulus):
    """pyx:73-99 -> rsdecode_impl.h:67-90.  Result is stripped (loop to deg P)."""
    assert len(x) == len(y)
    p = _to_int(modulus)
    xs = [_to_int(v) % p for v in x]
    ys = [_to_int(v) % p for v in y]
    return _interpolate(xs, ys, p)


def evaluate(polynomial, x, modulus):
    """pyx:101-113"""
    p = _to_int(modulus)
    return poly_eval([_el(c, p) for c in polynomial], _el(x, p), p)


def _mat_inverse(m, p):
    """Gauss-Jordan; returns None when singular (NTL inv(det, X, A), det==0)."""
    n = len(m)
    a = [list(row) + [1 if i == j else 0 for j in range(n)] for i, row in enumerate(m)]
    for c in range(n):
        piv = None
        for r in range(c, n):
            if a[r][c] % p:
                piv = r
                break
        if piv is None:
            return None
        a[c], a[piv] = a[piv], a[c]
        s = _inv(a[c][c], p)
        a[c] = [(v * s) % p for v in a[c]]
        for r in range(n):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(v - f * w) % p for v, w in zip(a[r], a[c])]
    return [row[n:] for row in a]


def _vandermonde(xs, d, p):
    """set_vm_matrix, rsdecode_impl.h:23-36: V[i][j] = x_i^j."""
    out = []
    for x in xs:
        row, acc = [], 1
        for _ in range(d):
            row.append(acc)
            acc = (acc * x) % p
        out.append(row)
    return out


def _vandermonde_inverse(xs, p):
    """rsdecode_impl.h:97-122"""
    return _mat_inverse(_vandermonde(xs, len(xs), p), p)


def vandermonde_inverse(x, modulus):
    """pyx:115-132.  Returns the NTL textual form of the matrix
    (``[[a b]\\n[c d]\\n]``); a singular input leaves NTL's result matrix
    empty (``[]``)."""
    p = _to_int(modulus)
    inv = _vandermonde_inverse([_to_int(v) % p for v in x], p)
    if inv is None:
        return "[]"
    return "[" + "".join("[" + " ".join(str(v) for v in row) + "]\n" for row in inv) + "]"


def vandermonde_batch_interpolate(x, data_list, modulus):
    """pyx:139-197.  Output rows have exactly k = max row length coefficients
    (not stripped); short rows are zero padded (pyx:172-181)."""
    p = _to_int(modulus)
    xs = [_to_int(v) % p for v in x]
    inv = _vandermonde_inverse(xs, p)
    if inv is None:
        raise InterpolationError("Interpolation failed")
    k = max(len(d) for d in data_list)
    if k != len(xs):
        raise ValueError("dimension mismatch (NTL would abort the process)")
    out = []
    for row in data_list:
        y = [_el(v, p) for v in row] + [0] * (k - len(row))
        out.append([sum(inv[j][l] * y[l] for l in range(k)) % p for j in range(k)])
    return out


def vandermonde_batch_evaluate(x, polynomials, modulus):
    """pyx:199-244: result[j][i] = sum_l polys[j][l] * x[i]^l."""
    p = _to_int(modulus)
    if not isinstance(x, (list, tuple)):
        raise ValueError("Invalid arguments")
    xs = [_el(v, p) for v in x]
    d = max(len(poly) for poly in polynomials)
    vm = _vandermonde(xs, d, p)
    out = []
    for poly in polynomials:
        c = [_el(v, p) for v in poly] + [0] * (d - len(poly))
        out.append([sum(vm[i][l] * c[l] for l in range(d)) % p for i in range(len(xs))])
    return out


FFT_VAN_THRESHOLD = 16  # rsdecode_impl.h:16


def _fft_rec(a, omega, n, p, van_matrix, van_threshold):
    """_fft, rsdecode_impl.h:125-169 (all n outputs are computed here; the
    C++ merely skips stores beyond m, which are never read)."""
    if n == 1:
        return a
    if van_matrix is not None and van_threshold == n:
        return [sum(van_matrix[i][j] * a[j] for j in range(n)) % p for i in range(n)]
    a0 = _fft_rec(a[0::2], omega * omega % p, n // 2, p, van_matrix, van_threshold)
    a1 = _fft_rec(a[1::2], omega * omega % p, n // 2, p, van_matrix, van_threshold)
    out = [0] * n
    w = 1
    for k in range(n // 2):
        t2 = (w * a1[k]) % p
        out[k] = (a0[k] + t2) % p
        out[k + n // 2] = (a0[k] - t2) % p
        w = (w * omega) % p
    return out


def _fft(coeffs, omega, n, p, k=-1):
    """fft, rsdecode_impl.h:171-192: coefficients beyond n are dropped (not
    wrapped), short inputs zero padded, base case = 16-point Vandermonde."""
    a = [coeffs[i] if i < len(coeffs) else 0 for i in range(n)]
    van = None
    if n >= FFT_VAN_THRESHOLD:
        omega_pow = pow(omega, n // FFT_VAN_THRESHOLD, p)
        xs = [pow(omega_pow, i, p) for i in range(FFT_VAN_THRESHOLD)]
        van = _vandermonde(xs, FFT_VAN_THRESHOLD, p)
    out = _fft_rec(a, omega, n, p, van, FFT_VAN_THRESHOLD)
    return out if k == -1 else out[:k]


def fft(coeffs, omega, modulus, n):
    """pyx:246-264"""
    p = _to_int(modulus)
    return _fft([_el(c, p) for c in coeffs], _el(omega, p), int(n), p)


def partial_fft(coeffs, omega, modulus, n, k):
    """pyx:266-284"""
    p = _to_int(modulus)
    return _fft([_el(c, p) for c in coeffs], _el(omega, p), int(n), p, int(k))


def fft_batch_evaluate(coeffs, omega, modulus, n, k):
    """pyx:286-316: every row is read to d = len(coeffs[0]) (pyx:295)."""
    p = _to_int(modulus)
    d = len(coeffs[0])
    w = _el(omega, p)
    return [_fft([_el(row[j], p) for j in range(d)], w, int(n), p, int(k)) for row in coeffs]


def _fnt_decode_step1(zs, omega, n, p):
    """rsdecode_impl.h:194-224 -> (A, [1/A'(x_i)])."""
    xs = [pow(omega, z, p) for z in zs]
    a = build_from_roots(xs, p)
    d = len(a) - 1
    ad = [((i + 1) * a[i + 1]) % p for i in range(d)]
    evals = _fft(ad, omega, n, p)
    return a, [_inv(evals[z], p) for z in zs]


def _fnt_decode_step2(a, ad_evals, zs, ys, omega, n, p):
    """rsdecode_impl.h:226-265"""
    k = len(zs)
    nis = [(ys[i] * ad_evals[i]) % p for i in range(k)]
    ncoeffs = [0] * n
    for i in range(k):
        # the C++ swaps nis[i] into place: a repeated z keeps the last write
        ncoeffs[zs[i]] = nis[i]
    omega_inv = _inv(omega, p)
    nrev = _fft(ncoeffs, omega_inv, n, p, k + 1 if k < n else n)
    q = [(-nrev[(i + 1) % n]) % p for i in range(k)]
    # MulTrunc(P, Q, A, k); VectorCopy(P_coeffs, P, k)
    out = [0] * k
    for i, qi in enumerate(q):
        if qi:
            for j, aj in enumerate(a):
                if i + j < k:
                    out[i + j] = (out[i + j] + qi * aj) % p
    return out


def fft_interpolate(zs, ys, omega, modulus, n):
    """pyx:318-340"""
    p = _to_int(modulus)
    w = _el(omega, p)
    zs = [int(z) for z in zs]
    a, ad = _fnt_decode_step1(zs, w, int(n), p)
    return _fnt_decode_step2(a, ad, zs, [_el(y, p) for y in ys[: len(zs)]], w, int(n), p)


def fft_batch_interpolate(zs, ys_list, omega, modulus, n):
    """pyx:342-381: step 1 once, step 2 per row."""
    p = _to_int(modulus)
    w = _el(omega, p)
    zs = [int(z) for z in zs]
    k = len(zs)
    a, ad = _fnt_decode_step1(zs, w, int(n), p)
    return [
        _fnt_decode_step2(a, ad, zs, [_el(row[j], p) for j in range(k)], w, int(n), p)
        for row in ys_list
    ]


def _partial_gcd(p0, p1, threshold, p):
    """rsdecode_impl.h:281-323 -> (r, v).  u is unused by the callers."""
    r0, r1 = _strip(p0), _strip(p1)
    t0, t1 = [], [1]
    if _deg(r0) < threshold:
        return r0, t0
    if _deg(r1) < threshold:
        return r1, t1
    while True:
        q, r2 = poly_divrem(r0, r1, p)
        t2 = poly_sub(t0, poly_mul(q, t1, p), p)
        if _deg(r2) < threshold:
            return r2, t2
        r0, r1 = r1, r2
        t0, t1 = t1, t2


def gao_interpolate(
    x, y, k, modulus, z=None, omega=None, order=None, use_omega_powers=False
):
    """pyx:389-439 -> rsdecode_impl.h:325-405.

    Returns (k coefficients, un-normalised error locator v of length
    deg(v)+1) or (None, None)."""
    assert len(x) == len(y)
    p = _to_int(modulus)
    is_null = [yi is None for yi in y]
    x = [x[i] for i in range(len(x)) if not is_null[i]]
    y = [y[i] for i in range(len(y)) if not is_null[i]]
    if z is not None:
        z = [z[i] for i in range(len(z)) if not is_null[i]]
    n = len(x)
    xs = [_el(v, p) for v in x]
    ys = [_el(v, p) for v in y]
    k = int(k)

    g0 = build_from_roots(xs, p)
    if use_omega_powers is True:
        assert z is not None
        assert len(z) == n
        assert omega is not None
        w = _el(omega, p)
        zs = [int(v) for v in z]
        a, ad = _fnt_decode_step1(zs, w, int(order), p)
        g1 = _strip(_fnt_decode_step2(a, ad, zs, ys, w, int(order), p))
    else:
        g1 = _interpolate(xs, ys, p)

    g, v = _partial_gcd(g0, g1, (n + k) // 2, p)
    if not _strip(v):
        # unreachable for k <= n (deg g0 = n >= threshold); NTL DivRem would abort
        return None, None
    f1, r = poly_divrem(g, v, p)
    if r or _deg(f1) >= k:
        return None, None
    res = [f1[i] if i < len(f1) else 0 for i in range(k)]
    return res, _strip(v)


def sqrt_mod(a, n):
    """pyx:441-444 (NTL SqrRootMod).  Either root is acceptable; the
    reference test only checks x*x == a (tests/test_ntl.py:331-341).
    Tonelli-Shanks, returns the smaller root for determinism."""
    p = _to_int(n)
    a = _to_int(a) % p
    if a == 0 or p == 2:
        return a
    if pow(a, (p - 1) // 2, p) != 1:
        raise ValueError("not a quadratic residue (NTL would abort the process)")
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    zz = 2
    while pow(zz, (p - 1) // 2, p) != p - 1:
        zz += 1
    m, c, t, r = s, pow(zz, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return min(r, p - r)


_ntl_threads = 1
_omp_threads = 1


def SetNTLNumThreads(x):  # noqa: N802  (pyx:383-384)
    global _ntl_threads
    _ntl_threads = int(x)


def AvailableNTLThreads():  # noqa: N802  (pyx:386-387)
    return _ntl_threads


def SetNumThreads(n):  # noqa: N802  (pyx:446-452)
    global _omp_threads
    SetNTLNumThreads(n)
    _omp_threads = int(n)


def GetMaxThreads():  # noqa: N802  (pyx:454-455)
    return _omp_threads


# --------------------------------------------------------------------------
# polynomial.py: get_omega / EvalPoint
# --------------------------------------------------------------------------


def get_omega(modulus, n, seed=None):
    """polynomial.py:253-268 with field.random (field.py:64-65)."""
    assert n & (n - 1) == 0, "n must be a power of 2"
    while True:
        x = Random(seed).randint(0, modulus - 1)
        y = pow(x, (modulus - 1) // n, modulus)
        if y == 1 or pow(y, n // 2, modulus) == 1:
            seed = None
            continue
        return y


class EvalPoint:
    """polynomial.py:385-423, on plain ints: ``point(i)`` returns an int."""

    def __init__(self, modulus, n, use_omega_powers=False):
        self.modulus = modulus
        self.n = n
        self.use_omega_powers = use_omega_powers
        if use_omega_powers:
            self.order = n if n & (n - 1) == 0 else 2 ** n.bit_length()
            self.omega2 = get_omega(modulus, 2 * self.order, seed=0)
            self.omega = self.omega2 * self.omega2 % modulus
        else:
            self.order = n
            self.omega2 = None
            self.omega = None

    def __call__(self, i):
        if self.use_omega_powers:
            return pow(self.omega2, 2 * i, self.modulus)
        return (i + 1) % self.modulus


# --------------------------------------------------------------------------
# reed_solomon_wb.py: Welch-Berlekamp
# --------------------------------------------------------------------------


def _rref(m, p):
    """reed_solomon_wb.py:157-197, in place."""
    if not m:
        return m
    num_rows, num_cols = len(m), len(m[0])
    i = j = 0
    while i < num_rows and j < num_cols:
        if m[i][j] == 0:
            nz = i
            while nz < num_rows and m[nz][j] == 0:
                nz += 1
            if nz == num_rows:
                j += 1
                continue
            m[i], m[nz] = m[nz], m[i]
        s = _inv(m[i][j], p)
        m[i] = [(v * s) % p for v in m[i]]
        for r in range(num_rows):
            if r != i and m[r][j] != 0:
                f = m[r][j]
                m[r] = [(y - f * x) % p for x, y in zip(m[i], m[r])]
        i += 1
        j += 1
    return m


def _some_solution(system, p, free_variable_value=1):
    """reed_solomon_wb.py:202-273"""
    _rref(system, p)
    i = -1
    while all(v == 0 for v in system[i]):
        i -= 1
    if all(v == 0 for v in system[i][:-1]):
        raise Exception("No solution")
    num_vars = len(system[0]) - 1
    values = [0] * num_vars
    free_vars, pivot_row = [], {}
    for j in range(num_vars):
        r = 0
        while r < len(system) and system[r][j] == 0:
            r += 1
        is_pivot = r < len(system) and system[r][j] == 1 and all(
            system[q][j] == 0 for q in range(r + 1, len(system))
        )
        if is_pivot:
            pivot_row[j] = r
        else:
            free_vars.append(j)
    for j in free_vars:
        values[j] = free_variable_value % p
    for j, r in pivot_row.items():
        values[j] = (system[r][-1] - sum(system[r][f] * values[f] for f in free_vars)) % p
    return values


def wb_solve_system(points, k, p, max_e):
    """reed_solomon_wb.py:79-127 -> (Q, E) stripped coefficient lists."""
    for e in range(max_e, 0, -1):
        e_num_vars, q_num_vars = e + 1, e + k
        system = []
        for a, b in points:
            pw = [pow(a, j, p) for j in range(max(e_num_vars, q_num_vars))]
            system.append(
                [(b * pw[j]) % p for j in range(e_num_vars)]
                + [(-pw[j]) % p for j in range(q_num_vars)]
                + [0]
            )
        system.append([0] * (e_num_vars - 1) + [1] + [0] * q_num_vars + [1])
        sol = _some_solution(system, p)
        e_poly = _strip(sol[: e + 1])
        q_poly = _strip(sol[e + 1 :])
        if not e_poly:
            # Polynomial.__divmod__ (polynomial.py:219-229) divides by the leading
            # coefficient of the all-zero E: GFElement inverse of 0 (field.py:133)
            raise ZeroDivisionError("Cannot invert zero")
        _, rem = poly_divrem(q_poly, e_poly, p)
        if not rem:
            return q_poly, e_poly
    raise ValueError("found no divisors!")


def wb_decode(encoded_msg, n, k, p, point):
    """reed_solomon_wb.py:129-151.  ``encoded_msg``: n entries, None =
    erasure.  Returns the stripped coefficient list."""
    assert len(encoded_msg) == n
    t = k - 1
    c = sum(m is None for m in encoded_msg)
    assert 2 * t + 1 + c <= n
    e = (n - c - t) // 2
    pts = [(point(i), m % p) for i, m in enumerate(encoded_msg) if m is not None]
    if e == 0:
        return _interpolate([a for a, _ in pts], [b for _, b in pts], p)
    q_poly, e_poly = wb_solve_system(pts, k, p, e)
    quo, rem = poly_divrem(q_poly, e_poly, p)
    if rem:
        raise Exception("Q is not divisibly by E!")
    return quo


def wb_robust_decode(z, encoded, n, k, p, point):
    """reed_solomon.py:200-225 (WelchBerlekampRobustDecoder.robust_decode)."""
    m = {zi: i for i, zi in enumerate(z)}
    ext = [encoded[m[i]] % p if i in m else None for i in range(n)]
    try:
        coeffs = wb_decode(ext, n, k, p, point)
    except Exception as e:  # noqa: BLE001 - mirrors the reference's catch-all
        if str(e) not in ("Wrong degree", "found no divisors!"):
            raise
        return None, None
    xs = [point(i) for i in range(n)]
    ev = vandermonde_batch_evaluate(xs, [coeffs], p)[0] if coeffs else [0] * n
    errors = [i for i in range(n) if ext[i] is not None and ext[i] != ev[i]]
    return coeffs, errors


def gao_robust_decode(z, encoded, n, k, p, point):
    """reed_solomon.py:160-186 (GaoRobustDecoder.robust_decode)."""
    x = [point(zi) for zi in z]
    if point.use_omega_powers:
        decoded, err = gao_interpolate(
            x, encoded, k, p, z=list(z), omega=point.omega, order=point.order,
            use_omega_powers=True,
        )
    else:
        decoded, err = gao_interpolate(x, encoded, k, p)
    if decoded is None:
        return None, None
    errors = []
    if len(err) > 1:
        errors = [i for i in range(n) if poly_eval(err, point(i), p) == 0]
    return decoded, errors


# --------------------------------------------------------------------------
# reed_solomon.py: IncrementalDecoder (the decode inner loop of batch_reconstruct)
# --------------------------------------------------------------------------


class DecodeValidationError(Exception):
    """reed_solomon.py:228-229"""


class IncrementalDecoder:
    """reed_solomon.py:232-403 restated on plain ints, ONE ROW AT A TIME on the
    Byzantine path exactly like ``_robust_update`` (:334-365): the first row
    that must wait stops the loop, so later rows are never decoded in that
    round, and an exception of the robust decoder surfaces only when the loop
    reaches the row that causes it.  ``algorithm`` is "gao" or
    "welch-berlekamp"; the non-robust decoder / encoder are interpolation and
    evaluation on ``point`` (every reference codec computes the same map)."""

    def __init__(self, point, degree, batch_size, max_errors, algorithm="gao",
                 confirmed_errors=None, validator=None):
        self.point, self.p, self.n = point, point.modulus, point.n
        self.degree, self.batch_size, self.max_errors = degree, batch_size, max_errors
        self.algorithm, self.validator = algorithm, validator
        self._confirmed_errors = confirmed_errors if confirmed_errors is not None else set()
        self._available_points = set()
        self._z = []
        self._available_data = [[] for _ in range(batch_size)]
        self._guess_decoded = self._guess_encoded = None
        self._optimistic = True
        self._num_decoded = 0
        self._partial_result = []
        self._result = None

    def _robust_decode(self, z, row):
        fn = gao_robust_decode if self.algorithm == "gao" else wb_robust_decode
        return fn(z, row, self.n, self.degree + 1, self.p, self.point)

    def _min_points_required(self):  # :302-303
        return self.degree + 1 + self.max_errors - len(self._confirmed_errors)

    def _optimistic_update(self, idx, data):  # :305-332
        success = True
        if len(self._available_points) == self.degree + 1:
            xs = [self.point(i) for i in self._z]
            self._guess_decoded = vandermonde_batch_interpolate(xs, self._available_data, self.p) \
                if self.batch_size else []
            allx = [self.point(i) for i in range(self.n)]
            self._guess_encoded = vandermonde_batch_evaluate(allx, self._guess_decoded, self.p) \
                if self.batch_size else []
        else:
            for i in range(self.batch_size):
                if data[i] % self.p != self._guess_encoded[i][idx]:
                    success = False
                    break
            if not success:
                self._guess_decoded = self._guess_encoded = None
                self._optimistic = False
        if success and len(self._available_points) >= self._min_points_required():
            self._result = self._guess_decoded
        return success

    def _robust_update(self):  # :334-365
        while self._num_decoded < self.batch_size:
            decoded, errors = self._robust_decode(self._z, self._available_data[0])
            if decoded is None:
                break
            if len(self._available_points) - len(errors) < self._min_points_required():
                break
            self._num_decoded += 1
            self._available_data = self._available_data[1:]
            self._partial_result.append(decoded)
            self._confirmed_errors |= set(errors)
            self._available_points -= set(errors)
            for e in errors:
                at = self._z.index(e)
                del self._z[at]
                for row in self._available_data:
                    del row[at]
        if self._num_decoded == self.batch_size:
            self._result = self._partial_result

    def add(self, idx, data):  # :368-395
        if self.done():
            return
        if idx in self._available_points or idx in self._confirmed_errors:
            return
        if len(data) != self.batch_size:
            raise DecodeValidationError("Incorrect length of data")
        if self.validator is not None:
            for d in data:
                self.validator(d)
        self._available_points.add(idx)
        self._z.append(idx)
        for i in range(self._num_decoded, self.batch_size):
            self._available_data[i - self._num_decoded].append(data[i])
        if len(self._available_points) <= self.degree:
            return
        if self._optimistic and self._optimistic_update(idx, data):
            return
        if len(self._available_points) >= self._min_points_required():
            self._robust_update()

    def done(self):
        return self._result is not None

    def get_results(self):
        if self._result is not None:
            return self._result, self._confirmed_errors
        return None, None
