"""The compute steps of the reference's OFFLINE-phase callers of this path, on limb arrays
(``uint64[batch, width, 4]``, no Python ints), batched over many independent instances:

  * ``randousha``            offline_randousha.py:47-53 (share generation), :72-78 (the
                             hyper-invertible-matrix refinement = an n x n Vandermonde
                             product), :99-121 (the checkers' degree / secret test)
  * ``refine_triples``       progs/triple_refinement.py:43-88 (nine ``vandermonde_batch_*``
                             calls on batches of ONE in the reference; here one launch per
                             stage for any number of triple sets)
  * ``_write_polys``         preprocessing.py:222-226

Message passing, Beaver multiplication and file I/O stay with the callers; these functions
replace only what runs inside ``honeybadgermpc.ntl``.  Points are ``EvalPoint(field, n,
use_omega_powers=False)`` (x_i = i + 1) for randousha / _write_polys and the literal integer
points of triple_refinement.py.
"""

import numpy as np

from . import _native, ntl
from .ntl import pack_vec


def _ctx(p):
    return ntl._ctx(p)


# ---------------------------------------------------------------------------
# randousha
# ---------------------------------------------------------------------------


def randousha_share(coeffs, n, p):
    """offline_randousha.py:50-53: evaluate the random degree-t (or 2t) polynomials
    ``coeffs[batch, deg+1, 4]`` at the n party points -> ``[batch, n, 4]``; row b,
    column i is what party i receives for random element b."""
    xs = pack_vec(list(range(1, n + 1)), p)
    return ntl.vandermonde_batch_evaluate_limbs(xs, np.ascontiguousarray(coeffs), p)


def randousha_refine(received, n, t, p):
    """offline_randousha.py:72-78, :148-149: ``received[batch, n, 4]`` -- entry (b, s) is the
    share received from sender s for batch item b -- is read as the coefficient vector of a
    polynomial and evaluated at the n party points (the hyper-invertible matrix M[i][s] =
    (i+1)^s).  Returns ``(kept, to_check)``: the first n - 2t refined shares of every row (the
    outputs) and the last 2t (sent to the checkers)."""
    refined = randousha_share(received, n, p)
    big_t = n - 2 * t
    return refined[:, :big_t, :], refined[:, big_t:, :]


def degree_and_secret(shares, n, p):
    """offline_randousha.py:99-110 (``get_degree_and_secret``): interpolate each row
    ``shares[batch, n, 4]`` (one share from each of the n parties) and return
    ``(degrees int64[batch], secrets uint64[batch, 4])`` with the reference's convention
    that the zero polynomial has degree 0."""
    xs = pack_vec(list(range(1, n + 1)), p)
    polys = ntl.vandermonde_batch_interpolate_limbs(xs, np.ascontiguousarray(shares), p)
    nz = polys.any(axis=2)                                  # [batch, n]
    deg = np.where(nz.any(axis=1), n - 1 - np.argmax(nz[:, ::-1], axis=1), 0)
    return deg.astype(np.int64), np.ascontiguousarray(polys[:, 0, :])


def randousha_check(shares_t, shares_2t, n, t, p):
    """offline_randousha.py:112-121: True iff every row of ``shares_t`` has degree exactly t,
    every row of ``shares_2t`` degree exactly 2t, and the secrets agree row by row."""
    deg_t, sec_t = degree_and_secret(shares_t, n, p)
    deg_2t, sec_2t = degree_and_secret(shares_2t, n, p)
    return bool((deg_t == t).all() and (deg_2t == 2 * t).all() and np.array_equal(sec_t, sec_2t))


# ---------------------------------------------------------------------------
# refine_triples
# ---------------------------------------------------------------------------


def _interp_reencode(xs_k, xs_all, ys, p):
    """one fused launch: coefficients through (xs_k, ys) and their values at xs_all"""
    k, n = len(xs_k), len(xs_all)
    ys = np.ascontiguousarray(ys)
    out = np.empty((ys.shape[0], k + n, 4), dtype=np.uint64)
    try:
        _ctx(p).interpolate_reencode(pack_vec(xs_k, p), pack_vec(xs_all, p), ys, ys.shape[0], out)
    except _native.SingularError as e:
        raise ntl.InterpolationError("Interpolation failed: points are not distinct") from e
    return out[:, :k, :], out[:, k:, :]


def refine_triples_stage1(a_dirty, b_dirty, n, t, p):
    """triple_refinement.py:36-57.  ``a_dirty, b_dirty``: ``[batch, m, 4]`` (m dirty triples per
    instance, n - t <= m <= n).  A and B are the degree-d polynomials (d = (m-1)//2) through
    the first d+1 values at the points 0..d; returns ``(a_coeffs, b_coeffs, a_rest, b_rest)``
    with the rest = their values at the d further points d+1..2d (to be multiplied with the
    Beaver step by the caller)."""
    m = a_dirty.shape[1]
    assert a_dirty.shape == b_dirty.shape and n - t <= m <= n
    d = (m - 1) // 2
    first, more = list(range(d + 1)), list(range(d + 1, 2 * d + 1))
    a_coeffs, a_rest = _interp_reencode(first, more, a_dirty[:, : d + 1, :], p)
    b_coeffs, b_rest = _interp_reencode(first, more, b_dirty[:, : d + 1, :], p)
    return a_coeffs, b_coeffs, a_rest, b_rest


def refine_triples_stage2(a_coeffs, b_coeffs, c_first, c_rest, n, t, p):
    """triple_refinement.py:71-88.  ``c_first[batch, d+1, 4]`` (the dirty c values at points
    0..d) and ``c_rest[batch, d, 4]`` (the Beaver products at d+1..2d) define C of degree 2d;
    returns ``(p, q, pq)``: A, B, C at the k = d + 1 - t fresh points n+1..n+k."""
    d = a_coeffs.shape[1] - 1
    k = d + 1 - t
    fresh = list(range(n + 1, n + 1 + k))
    c_all = np.concatenate([c_first, c_rest], axis=1)
    _, pq = _interp_reencode(list(range(2 * d + 1)), fresh, c_all, p)
    xf = pack_vec(fresh, p)
    pv = ntl.vandermonde_batch_evaluate_limbs(xf, np.ascontiguousarray(a_coeffs), p)
    qv = ntl.vandermonde_batch_evaluate_limbs(xf, np.ascontiguousarray(b_coeffs), p)
    return pv, qv, pq


# ---------------------------------------------------------------------------
# preprocessing._write_polys
# ---------------------------------------------------------------------------


def write_polys_values(polys, n, p):
    """preprocessing.py:222-231: ``polys[batch, t+1, 4]`` -> ``[n, batch, 4]``: row i is the
    list of share values written to party i's preprocessing file."""
    xs = pack_vec(list(range(1, n + 1)), p)
    vals = ntl.vandermonde_batch_evaluate_limbs(xs, np.ascontiguousarray(polys), p)
    return np.ascontiguousarray(vals.transpose(1, 0, 2))
