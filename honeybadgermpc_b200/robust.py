"""Host side of the robust (error-correcting) decoders: marshalling around the
batched CUDA Gao and Welch-Berlekamp kernels (``hbg_gao_decode_batch``,
``hbg_wb_decode_batch``).

Replaces, for this path, ``gao_interpolate`` (ntl/hbmpc_ntl_helpers.pyx:389-439
over rsdecode_impl.h:281-405) and the pure-Python Welch-Berlekamp solver
(reed_solomon_wb.py:79-151).  The reference decodes one received word per call;
here every entry point is batched over received words that share the same set
of present positions, and the single-word API is the batch of one.
"""

import numpy as np

from . import _native
from .ntl import _to_int, pack_rows, pack_vec, unpack_rows

WB_OK, WB_NO_DIVISORS, WB_NO_SOLUTION, WB_ZERO_DIVISOR = 0, 1, 2, 3


def _ctx(p):
    return _native.get_context(p)


# --------------------------------------------------------------------------
# Gao
# --------------------------------------------------------------------------


def gao_decode_batch_limbs(xs, ys, k, p):
    """xs: uint64[m,4] points; ys: uint64[batch,m,4] received words.
    -> (coeffs uint64[batch,k,4], locator uint64[batch,L,4], loc_len int32[batch],
        status int32[batch])   with status 0 = decoded, 1 = failed."""
    m, batch = len(xs), ys.shape[0]
    loc_stride = max(1, m - (m + k) // 2 + 1)
    coeffs = np.zeros((batch, k, 4), dtype=np.uint64)
    locator = np.zeros((batch, loc_stride, 4), dtype=np.uint64)
    loc_len = np.zeros(batch, dtype=np.int32)
    status = np.ones(batch, dtype=np.int32)
    try:
        _ctx(p).gao_decode_batch(np.ascontiguousarray(xs), k, np.ascontiguousarray(ys), batch,
                                 coeffs, locator, loc_stride, loc_len, status)
    except _native.SingularError as e:
        raise ZeroDivisionError("repeated point (NTL would abort the process)") from e
    return coeffs, locator, loc_len, status


def gao_interpolate(x, y, k, modulus, z=None, omega=None, order=None, use_omega_powers=False):
    """pyx:389-439.  ``None`` entries of ``y`` are erasures.  Returns
    ``(k coefficients, error locator)`` or ``(None, None)``.  With
    ``use_omega_powers`` the reference interpolates g1 through the FFT
    (gao_interpolate_fft, rsdecode_impl.h:365-405); the interpolant is the same
    polynomial, so both variants run the same kernel on the points ``x``."""
    p = _to_int(modulus)
    keep = [i for i, v in enumerate(y) if v is not None]
    xs = [x[i] for i in keep]
    ys = [y[i] for i in keep]
    if use_omega_powers is True:
        assert z is not None
        assert len([z[i] for i in keep]) == len(xs)
        assert omega is not None
        int(order)
    k = int(k)
    if len(xs) == 0:
        _ctx(p)
        return None, None
    xl = pack_vec(xs, p)
    yl = pack_rows([ys], len(ys), p)
    coeffs, locator, loc_len, status = gao_decode_batch_limbs(xl, yl, k, p)
    if status[0] != 0:
        return None, None
    return unpack_rows(coeffs)[0], unpack_rows(locator[:, : int(loc_len[0])])[0]


# --------------------------------------------------------------------------
# Welch-Berlekamp
# --------------------------------------------------------------------------


def wb_decode_batch_limbs(xs, ys, k, e_max, p):
    """xs: uint64[m,4] points of the received positions; ys: uint64[batch,m,4].
    -> (coeffs uint64[batch,k,4], out_len int32[batch], status int32[batch])."""
    batch = ys.shape[0]
    coeffs = np.zeros((batch, k, 4), dtype=np.uint64)
    out_len = np.zeros(batch, dtype=np.int32)
    status = np.ones(batch, dtype=np.int32)
    _ctx(p).wb_decode_batch(np.ascontiguousarray(xs), k, e_max, np.ascontiguousarray(ys), batch,
                            coeffs, out_len, status)
    return coeffs, out_len, status


def wb_decode_rows(xs_ints, rows, n_total, n_erased, k, p):
    """Decode received words (lists of ints on the points ``xs_ints``) the way
    ``make_wb_encoder_decoder(...).decode`` does (reed_solomon_wb.py:129-151).
    One entry per row: the stripped coefficient list; ``None`` where the
    reference raises ``ValueError("found no divisors!")`` (swallowed by its
    caller, reed_solomon.py:205-212); or the EXCEPTION INSTANCE the reference
    would raise for that row (``"No solution"``, reed_solomon_wb.py:244; the
    inverse of zero when E(x) vanishes, polynomial.py:219-229 / field.py:133).
    It is returned, not raised: the reference decodes one row per call, so the
    caller must raise it only when it gets to that row (reed_solomon.py:334-348).
    The pre-condition on the number of erasures (reed_solomon_wb.py:132) does not
    depend on the row and is asserted here."""
    t = k - 1
    assert 2 * t + 1 + n_erased <= n_total
    e_max = (n_total - n_erased - t) // 2
    xl = pack_vec(xs_ints, p)
    yl = pack_rows(rows, len(xs_ints), p)
    if e_max == 0:
        # no redundancy left: plain interpolation through all points, no degree
        # check (reed_solomon_wb.py:142-145)
        from .ntl import vandermonde_batch_interpolate_limbs, _strip

        out = unpack_rows(vandermonde_batch_interpolate_limbs(xl, yl, p))
        return [_strip(r) for r in out]
    coeffs, out_len, status = wb_decode_batch_limbs(xl, yl, k, e_max, p)
    ints = unpack_rows(coeffs)
    res = []
    for i in range(len(rows)):
        s = int(status[i])
        if s == WB_OK:
            res.append(ints[i][: int(out_len[i])])
        elif s == WB_NO_DIVISORS:
            res.append(None)
        elif s == WB_NO_SOLUTION:
            res.append(Exception("No solution"))
        else:
            res.append(ZeroDivisionError("Cannot invert zero"))
    return res
