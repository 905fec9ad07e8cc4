"""CPU-side checks of the drop-in boundary: the CUDA library builds and loads,
exports every symbol declared in include/hbmpc_b200.h (no compute calls -- there
is no GPU here), the Python shim fails loudly without a device, and the product
package never imports the oracle."""

import ctypes
import os
import re
import subprocess
import sys

import pytest
from conftest import BLS12_381_R as P
from conftest import ROOT

sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hbmpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hbg_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    path = graft.build_cuda()
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 13
    for name in names:
        assert hasattr(lib, name), f"{name} declared in hbmpc_b200.h but not exported"
    lib.hbg_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.hbg_version()


def test_binding_covers_header():
    from honeybadgermpc_b200 import _native

    assert sorted(_native.SIGNATURES) == _declared_symbols()
    _native.load_library()


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "honeybadgermpc_b200", "libhbmpc_b200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from honeybadgermpc_b200 import ntl
    from honeybadgermpc_b200._native import NativeLibraryError

    for call in (lambda: ntl.vandermonde_batch_evaluate([1, 2], [[0, 1]], P),
                 lambda: ntl.fft([0, 1], 5, 13, 4),
                 lambda: ntl.gao_interpolate([1, 2, 3], [1, 2, 3], 1, P),
                 lambda: ntl.lagrange_interpolate([1, 2], [1, 2], P)):
        with pytest.raises(NativeLibraryError):
            call()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "honeybadgermpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "cpu_ref" not in text and "hbmpc_oracle" not in text, f


def test_c_marshalling_matches_python():
    """csrc/pymarshal.c against the pure-Python conversion, incl. the reference's
    error behaviour (negative ints -> OverflowError, pyx:20-22)"""
    import random

    import numpy as np

    graft.build_marshal()
    import importlib

    import honeybadgermpc_b200.ntl as ntl

    ntl = importlib.reload(ntl)
    assert ntl._marshal is not None
    rng = random.Random(5)
    rows = [[rng.randrange(P) for _ in range(rng.randint(0, 7))] for _ in range(200)]
    rows += [[0, 1, P - 1, P, 2 * P + 5, 2 ** 300 + 7], (3, 4, 5), [True]]
    # the digit-level fast path of pack_one: every 30-bit digit and 64-bit limb boundary, the
    # 256-bit edge (values >= 2^256 and >= p take the general path and are reduced)
    rows += [[2 ** (30 * k) - 1, 2 ** (30 * k), 2 ** (30 * k) + 1] for k in range(1, 10)]
    rows += [[2 ** (64 * k) - 1, 2 ** (64 * k), 2 ** (64 * k) + 1] for k in range(1, 5)]
    rows += [[2 ** 255, 2 ** 256 - 1, 2 ** 256, 2 ** 256 + P, (1 << 256) - P, 2 ** 269, 2 ** 270]]
    for width in (0, 1, 4, 7):
        a = ntl.pack_rows(rows, width, P)
        assert np.array_equal(a, ntl._pack_rows_py(rows, width, P))
        assert ntl.unpack_rows(a) == ntl._unpack_rows_py(a)
    assert ntl.pack_rows([[27, 5]], 2, 13).tolist() == ntl._pack_rows_py([[27, 5]], 2, 13).tolist()
    for bad, exc in (([[-1]], OverflowError), ([["x"]], TypeError), ([[1.5]], TypeError)):
        with pytest.raises(exc):
            ntl.pack_rows(bad, 1, P)


def test_c_long_construction_matches_cpython():
    """pymarshal.c builds 256-bit ints from 30-bit digits (_PyLong_FromDigits, CPython >= 3.12):
    same objects as _PyLong_FromByteArray on digit-boundary values, zero, and random widths --
    value, hash, bit_length, arithmetic and str (which walk the digits) all agree"""
    import ctypes
    import random

    graft.build_marshal()
    lib = ctypes.PyDLL(os.path.join(ROOT, "honeybadgermpc_b200", "libhbmpc_pymarshal.so"))
    lib.hbg_py_long_from_le32.restype = ctypes.py_object
    lib.hbg_py_long_from_le32.argtypes = [ctypes.c_char_p, ctypes.c_int]
    rng = random.Random(5)
    vals = [0, 1, 2, 255, 256, 2 ** 256 - 1, 2 ** 255, P, P - 1]
    for k in range(1, 9):  # around every 30-bit digit boundary and every 64-bit limb boundary
        for base in (2 ** (30 * k), 2 ** (64 * min(k, 3))):
            vals += [base - 1, base, base + 1]
    vals += [rng.getrandbits(rng.randrange(1, 257)) for _ in range(5000)]
    for v in vals:
        b = v.to_bytes(32, "little")
        fast, ref = lib.hbg_py_long_from_le32(b, 0), lib.hbg_py_long_from_le32(b, 1)
        assert type(fast) is int and fast == ref == v, v
        assert hash(fast) == hash(v) and fast.bit_length() == v.bit_length() and str(fast) == str(v)
        assert fast + 1 == v + 1 and fast * fast == v * v and (fast >> 31) == (v >> 31) and -fast == -v


def test_host_marshalling_round_trip():
    from honeybadgermpc_b200.ntl import pack_rows, unpack_rows

    rows = [[0, 1, P - 1], [P, 2 * P + 5], []]
    arr = pack_rows(rows, 3, P)
    assert arr.shape == (3, 3, 4)
    assert unpack_rows(arr) == [[0, 1, P - 1], [0, 5, 0], [0, 0, 0]]
    with pytest.raises(OverflowError):
        pack_rows([[-1]], 1, P)


def test_c_element_packing_matches_python():
    """hbg_py_pack_elements: [e.value for e in shares] + pack_rows in one pass -- GFElement slots,
    foreign objects with a `value` attribute, reduction mod p, padding, truncation, errors"""
    import random

    import numpy as np

    graft.build_marshal()
    import importlib

    import honeybadgermpc_b200.ntl as ntl
    from honeybadgermpc_b200.field import GF

    ntl = importlib.reload(ntl)
    assert ntl._marshal is not None
    field = GF(P)
    rng = random.Random(3)

    class Foreign:
        def __init__(self, v):
            self.value = v

    elems = [field(rng.randrange(P)) for _ in range(300)] + [Foreign(7), Foreign(P + 3), Foreign(2 ** 300), field(0)]
    want = ntl._pack_rows_py([[e.value for e in elems]], len(elems) + 5, P)[0]
    assert np.array_equal(ntl.pack_elements(elems, len(elems) + 5, P), want)       # zero padded
    assert np.array_equal(ntl.pack_elements(tuple(elems), 10, P), want[:10])       # truncated, any sequence
    assert ntl.pack_elements([], 3, P).tolist() == [[0] * 4] * 3
    with pytest.raises(AttributeError):
        ntl.pack_elements([object()], 1, P)
    with pytest.raises(OverflowError):
        ntl.pack_elements([Foreign(-1)], 1, P)
    saved, ntl._marshal = ntl._marshal, None
    try:
        assert np.array_equal(ntl.pack_elements(elems, len(elems) + 5, P), want)
    finally:
        ntl._marshal = saved


def test_c_element_wrapping_matches_python():
    """hbg_py_wrap_elements (csrc/pymarshal.c) builds the same GFElement objects as the
    Python constructor; the pure-Python fallback of ntl.wrap_elements agrees"""
    import gc
    import random

    import numpy as np

    graft.build_marshal()
    import importlib

    import honeybadgermpc_b200.ntl as ntl
    from honeybadgermpc_b200.field import GF, GFElement

    ntl = importlib.reload(ntl)
    assert ntl._marshal is not None
    field = GF(P)
    rng = random.Random(9)
    vals = [0, 1, P - 1, 2 ** 64, 2 ** 255 % P] + [rng.randrange(P) for _ in range(500)]
    arr = ntl.pack_rows([vals], len(vals), P)[0]
    got = ntl.wrap_elements(arr, field)
    assert [type(e) for e in got] == [GFElement] * len(vals)
    assert got == [field(v) for v in vals]
    assert all(e.field is field and e.modulus == P and e.value == v for e, v in zip(got, vals))
    assert (got[3] * got[4] + 1).value == (vals[3] * vals[4] + 1) % P  # usable like any other element
    assert ntl.wrap_elements(np.zeros((0, 4), np.uint64), field) == []
    saved, ntl._marshal = ntl._marshal, None
    try:
        assert ntl.wrap_elements(arr, field) == got
    finally:
        ntl._marshal = saved
    del got
    gc.collect()
