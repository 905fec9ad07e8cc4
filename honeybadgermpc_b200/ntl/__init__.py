"""Drop-in for ``honeybadgermpc.ntl`` (``from ._hbmpc_ntl_helpers import *``,
/root/reference/honeybadgermpc/ntl/__init__.py:1): the same 16 functions and
``InterpolationError``, same positional arguments, list-of-int in, list-of-int
out, ``modulus`` per call -- computed by the sm_100a kernels of
``libhbmpc_b200.so`` instead of NTL.

Install it under the reference with::

    import sys, honeybadgermpc_b200.ntl as m
    sys.modules["honeybadgermpc.ntl._hbmpc_ntl_helpers"] = m

There is no CPU fallback: every batch function raises ``NativeLibraryError``
when the CUDA library or a device is missing.  Host-side Python only marshals
ints <-> 4x64-bit limbs and, for ``sqrt_mod`` (a scalar helper with no batch
axis), runs Tonelli-Shanks on Python ints.

``*_limbs`` variants take/return numpy ``uint64[..., 4]`` arrays and skip the
Python-int marshalling (the reference has no such entry points; they are the
fast path used by ``honeybadgermpc_b200.reed_solomon``).
"""

import numpy as np

from .. import _native
from .._native import NativeLibraryError  # noqa: F401  (re-export)

__all__ = [
    "lagrange_interpolate", "evaluate", "vandermonde_inverse", "InterpolationError",
    "vandermonde_batch_interpolate", "vandermonde_batch_evaluate", "fft", "partial_fft",
    "fft_batch_evaluate", "fft_interpolate", "fft_batch_interpolate", "SetNTLNumThreads",
    "AvailableNTLThreads", "gao_interpolate", "sqrt_mod", "SetNumThreads", "GetMaxThreads",
]


class InterpolationError(Exception):
    """hbmpc_ntl_helpers.pyx:135"""


# --------------------------------------------------------------------------
# marshalling (pyx:20-70)
# --------------------------------------------------------------------------

_ZERO32 = bytes(32)


def _to_int(v):
    """py_obj_to_ZZ (pyx:37-46): ints pass, str/bytes are decimal text, None is
    a ValueError; negative ints fail like ``int.to_bytes`` in intToZZ."""
    if isinstance(v, int):
        if v < 0:
            raise OverflowError("can't convert negative int to unsigned")
        return int(v)
    if v is None:
        raise ValueError(f"Unsupported data type. {type(v)}")
    if isinstance(v, bytes):
        v = v.decode()
    return int(str(v).strip())


def _el_bytes(v, p):
    """intToZZp (pyx:31-32) -> 32 little-endian bytes of v mod p."""
    if v >= p:
        v %= p
    return v.to_bytes(32, "little")  # OverflowError for negatives, as in the reference


def _pack_rows_py(rows, width, p):
    parts = []
    for row in rows:
        m = len(row)
        if m >= width:
            parts.extend(_el_bytes(row[j], p) for j in range(width))
        else:
            parts.extend(_el_bytes(v, p) for v in row)
            parts.append(_ZERO32 * (width - m))
    buf = b"".join(parts)
    return np.frombuffer(buf, dtype=np.uint64).reshape(len(rows), width, 4)


def _unpack_rows_py(arr):
    batch, width = arr.shape[0], arr.shape[1]
    if batch * width == 0:
        return [[] for _ in range(batch)]
    mv = memoryview(np.ascontiguousarray(arr)).cast("B")
    frm = int.from_bytes
    out = []
    pos = 0
    for _ in range(batch):
        row = [frm(mv[pos + 32 * j: pos + 32 * j + 32], "little") for j in range(width)]
        pos += 32 * width
        out.append(row)
    return out


def _load_marshal():
    """The C helper (csrc/pymarshal.c, built by __graft_entry__.build()); it only
    converts ints <-> limbs, so a missing helper means the slower pure-Python
    conversion, not a different result."""
    import ctypes
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "libhbmpc_pymarshal.so")
    if not os.path.exists(path) or os.environ.get("HBMPC_B200_PY_MARSHAL"):
        return None
    try:
        lib = ctypes.PyDLL(path)
        lib.hbg_py_pack_rows.restype = ctypes.py_object
        lib.hbg_py_pack_rows.argtypes = [ctypes.py_object, ctypes.c_ssize_t, ctypes.py_object,
                                         ctypes.c_void_p]
        lib.hbg_py_unpack_rows.restype = ctypes.py_object
        lib.hbg_py_unpack_rows.argtypes = [ctypes.c_void_p, ctypes.c_ssize_t, ctypes.c_ssize_t]
        lib.hbg_py_pack_elements.restype = ctypes.py_object
        lib.hbg_py_pack_elements.argtypes = [ctypes.py_object, ctypes.c_ssize_t, ctypes.py_object,
                                             ctypes.py_object, ctypes.c_void_p]
        lib.hbg_py_wrap_elements.restype = ctypes.py_object
        lib.hbg_py_wrap_elements.argtypes = [ctypes.c_void_p, ctypes.c_ssize_t, ctypes.py_object,
                                             ctypes.py_object, ctypes.py_object]
        return lib
    except (OSError, AttributeError):
        return None


_marshal = _load_marshal()


def pack_rows(rows, width, p):
    """list of (ragged) int rows -> uint64[len(rows), width, 4]; short rows are
    zero padded, long rows truncated; values are reduced mod p (intToZZp,
    pyx:31-32); negative ints raise OverflowError as in the reference."""
    if _marshal is None:
        return _pack_rows_py(rows, width, p)
    out = np.empty((len(rows), width, 4), dtype=np.uint64)
    _marshal.hbg_py_pack_rows(rows, width, p, out.ctypes.data)
    return out


def pack_vec(values, p):
    return pack_rows([values], len(values), p)[0]


def unpack_rows(arr):
    """uint64[batch, width, 4] -> list of lists of int."""
    if _marshal is None:
        return _unpack_rows_py(arr)
    arr = np.ascontiguousarray(arr)
    return _marshal.hbg_py_unpack_rows(arr.ctypes.data, arr.shape[0], arr.shape[1])


def pack_elements(elements, width, p):
    """sequence of field elements (objects with an int ``value``) -> uint64[width, 4], reduced mod p
    and zero padded: ``pack_rows([[e.value for e in elements]], width, p)[0]`` in one C pass."""
    if _marshal is None:
        return _pack_rows_py([[e.value for e in elements]], width, p)[0]
    from ..field import GFElement

    out = np.empty((width, 4), dtype=np.uint64)
    _marshal.hbg_py_pack_elements(elements, width, GFElement, p, out.ctypes.data)
    return out


def wrap_elements(arr, field):
    """uint64[count, 4] canonical residues -> list of ``GFElement`` of ``field`` (what
    ``batch_reconstruct`` returns), built in C without a Python-level loop."""
    from ..field import GFElement

    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    if _marshal is None:
        return field.wrap_canonical(unpack_rows(arr[None])[0]) if arr.shape[0] else []
    return _marshal.hbg_py_wrap_elements(arr.ctypes.data, arr.shape[0], GFElement, field, field.modulus)


def _strip(a):
    a = list(a)
    while a and a[-1] == 0:
        a.pop()
    return a


def _ctx(p):
    return _native.get_context(p)


# --------------------------------------------------------------------------
# limb-array entry points
# --------------------------------------------------------------------------


def vandermonde_batch_evaluate_limbs(xs, polys, modulus):
    """xs: uint64[n,4]; polys: uint64[batch,d,4] -> uint64[batch,n,4]"""
    batch, d = polys.shape[0], polys.shape[1]
    out = np.empty((batch, len(xs), 4), dtype=np.uint64)
    _ctx(modulus).vandermonde_batch_evaluate(
        np.ascontiguousarray(xs), np.ascontiguousarray(polys), batch, d, out)
    return out


def vandermonde_batch_interpolate_limbs(xs, ys, modulus):
    """xs: uint64[k,4]; ys: uint64[batch,k,4] -> uint64[batch,k,4]"""
    batch = ys.shape[0]
    out = np.empty((batch, len(xs), 4), dtype=np.uint64)
    try:
        _ctx(modulus).vandermonde_batch_interpolate(
            np.ascontiguousarray(xs), np.ascontiguousarray(ys), batch, out)
    except _native.SingularError as e:
        raise InterpolationError("Interpolation failed") from e
    return out


def fft_batch_evaluate_limbs(polys, omega, modulus, n, k):
    """polys: uint64[batch,d,4]; omega: uint64[4] -> uint64[batch,k,4]"""
    batch, d = polys.shape[0], polys.shape[1]
    out = np.empty((batch, k, 4), dtype=np.uint64)
    _ctx(modulus).fft_batch_evaluate(
        np.ascontiguousarray(omega), n, np.ascontiguousarray(polys), batch, d, k, out)
    return out


def fft_batch_interpolate_limbs(zs, ys, omega, modulus, n):
    """zs: k ints; ys: uint64[batch,k,4] -> uint64[batch,k,4]"""
    batch = ys.shape[0]
    out = np.empty((batch, len(zs), 4), dtype=np.uint64)
    _ctx(modulus).fft_batch_interpolate(
        np.ascontiguousarray(omega), n, zs, np.ascontiguousarray(ys), batch, out)
    return out


# --------------------------------------------------------------------------
# the reference API
# --------------------------------------------------------------------------


def vandermonde_batch_evaluate(x, polynomials, modulus):
    """pyx:199-244: ``result[j][i] = sum_l polynomials[j][l] * x[i]**l``.
    Ragged rows are zero padded to the longest (pyx:217,232-233)."""
    p = _to_int(modulus)
    if not isinstance(x, (list, tuple)):
        raise ValueError("Invalid arguments")
    xs = pack_vec(list(x), p)
    d = max(len(poly) for poly in polynomials)
    polys = pack_rows(polynomials, d, p)
    return unpack_rows(vandermonde_batch_evaluate_limbs(xs, polys, p))


def vandermonde_batch_interpolate(x, data_list, modulus):
    """pyx:139-197: rows of exactly k = len(x) coefficients (not stripped);
    ``InterpolationError`` when two points coincide (pyx:168-169)."""
    p = _to_int(modulus)
    xs = pack_vec([_to_int(v) for v in x], p)
    k = max(len(row) for row in data_list)
    if k != len(xs):
        # the singularity check comes first in the reference (pyx:167-169); after
        # it NTL aborts the process on the (len(x) x len(x)) * (k x batch) product
        vandermonde_batch_interpolate_limbs(xs, np.zeros((0, len(xs), 4), np.uint64), p)
        raise ValueError("dimension mismatch (NTL would abort the process)")
    ys = pack_rows(data_list, k, p)
    return unpack_rows(vandermonde_batch_interpolate_limbs(xs, ys, p))


def lagrange_interpolate(x, y, modulus):
    """pyx:73-99: coefficients of the interpolant, trailing zeros stripped."""
    assert len(x) == len(y)
    p = _to_int(modulus)
    xs = pack_vec([_to_int(v) for v in x], p)
    ys = pack_rows([[_to_int(v) for v in y]], len(xs), p)
    if len(xs) == 0:
        _ctx(p)
        return []
    try:
        out = vandermonde_batch_interpolate_limbs(xs, ys, p)
    except InterpolationError as e:
        raise ZeroDivisionError("repeated interpolation point (NTL would abort the process)") from e
    return _strip(unpack_rows(out)[0])


def evaluate(polynomial, x, modulus):
    """pyx:101-113"""
    p = _to_int(modulus)
    xs = pack_vec([x], p)
    if len(polynomial) == 0:
        _ctx(p)
        return 0
    polys = pack_rows([polynomial], len(polynomial), p)
    return unpack_rows(vandermonde_batch_evaluate_limbs(xs, polys, p))[0][0]


def vandermonde_inverse(x, modulus):
    """pyx:115-132: the inverse Vandermonde matrix in NTL's textual form."""
    p = _to_int(modulus)
    xs = pack_vec([_to_int(v) for v in x], p)
    k = len(xs)
    ident = pack_rows([[1 if i == j else 0 for j in range(k)] for i in range(k)], k, p)
    try:
        cols = unpack_rows(vandermonde_batch_interpolate_limbs(xs, ident, p))
    except InterpolationError:
        return "[]"
    return "[" + "".join(
        "[" + " ".join(str(cols[j][i]) for j in range(k)) + "]\n" for i in range(k)) + "]"


def fft(coeffs, omega, modulus, n):
    """pyx:246-264: ``out[i] = sum_j coeffs[j] * omega**(i*j)``, i < n;
    coefficients beyond n are dropped (rsdecode_impl.h:173-175)."""
    return partial_fft(coeffs, omega, modulus, n, n)


def partial_fft(coeffs, omega, modulus, n, k):
    """pyx:266-284"""
    if not isinstance(coeffs, (list, tuple)):
        raise ValueError("Invalid arguments")
    p = _to_int(modulus)
    n, k = int(n), int(k)
    d = min(len(coeffs), n)
    polys = pack_rows([coeffs], d, p)
    return unpack_rows(fft_batch_evaluate_limbs(polys, pack_vec([omega], p)[0], p, n, k))[0]


def fft_batch_evaluate(coeffs, omega, modulus, n, k):
    """pyx:286-316: every row is read to ``len(coeffs[0])`` entries (pyx:295)."""
    p = _to_int(modulus)
    n, k = int(n), int(k)
    d = len(coeffs[0])
    for row in coeffs:
        if len(row) < d:
            raise IndexError("list index out of range")  # what pyx:302 would raise
    polys = pack_rows(coeffs, d, p)
    return unpack_rows(fft_batch_evaluate_limbs(polys, pack_vec([omega], p)[0], p, n, k))


def fft_interpolate(zs, ys, omega, modulus, n):
    """pyx:318-340"""
    return fft_batch_interpolate(zs, [ys], omega, modulus, n)[0]


def fft_batch_interpolate(zs, ys_list, omega, modulus, n):
    """pyx:342-381: per row, the k = len(zs) coefficients of the polynomial
    through ``(omega**zs[i], ys[i])``."""
    p = _to_int(modulus)
    zs = [int(z) for z in zs]
    k = len(zs)
    for row in ys_list:
        if len(row) < k:
            raise IndexError("list index out of range")
    ys = pack_rows(ys_list, k, p)
    try:
        out = fft_batch_interpolate_limbs(zs, ys, pack_vec([omega], p)[0], p, int(n))
    except _native.SingularError as e:
        raise ZeroDivisionError("repeated z (NTL would abort the process in inv(0))") from e
    return unpack_rows(out)


def gao_interpolate(x, y, k, modulus, z=None, omega=None, order=None, use_omega_powers=False):
    """pyx:389-439"""
    assert len(x) == len(y)
    from .. import robust

    return robust.gao_interpolate(x, y, k, modulus, z, omega, order, use_omega_powers)


def sqrt_mod(a, n):
    """pyx:441-444 (NTL SqrRootMod): a square root of ``a`` modulo the prime
    ``n``.  Scalar helper, no batch axis: Tonelli-Shanks on Python ints.  The
    reference test accepts either root (tests/test_ntl.py:331-341)."""
    p = _to_int(n)
    a = _to_int(a) % p
    if a == 0 or p == 2:
        return a
    from ..field import sqrt_mod_prime

    return sqrt_mod_prime(a, p)


# thread knobs: the GPU path has no thread pool; the values are stored and
# echoed because DecoderSelector branches on AvailableNTLThreads()
# (reed_solomon.py:455-459).
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
