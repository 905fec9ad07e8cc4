"""CPU unit tests of the host twins of the device field arithmetic
(honeybadgermpc_b200/csrc/fp256.cuh, host_math.hpp): the same algorithm text
the kernels run (even/odd CIOS multiplier, the BLS fast reduction rows, the
lazy dot-product accumulator), checked against Python ints."""

import ctypes
import os
import random
import sys

import numpy as np
import pytest
from conftest import BLS12_381_R as P
from conftest import ROOT

sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(graft.build_host_selftest())


def limbs(v):
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint64).copy()


def many(vs):
    return np.concatenate([limbs(v) for v in vs]) if vs else np.zeros(0, np.uint64)


def val(a):
    return int.from_bytes(a.tobytes(), "little")


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


EDGE = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, 2 ** 32 - 1, 2 ** 32, 2 ** 64 - 1, 2 ** 224 + 5]


@pytest.mark.parametrize("p", [P, 13, 53, 2 ** 61 - 1, 2 ** 127 - 1, 2 ** 254 - 127 * 2 ** 100 + 1 | 1])
def test_mulmod_addsub(lib, p):
    rng = random.Random(p % 1000)
    vals = [v % p for v in EDGE] + [rng.randrange(p) for _ in range(40)]
    out, s, d, n = (np.zeros(4, np.uint64) for _ in range(4))
    for a in vals[:14]:
        for b in vals:
            assert lib.hbt_mulmod(ptr(limbs(p)), ptr(limbs(a)), ptr(limbs(b)), ptr(out)) == 0
            assert val(out) == a * b % p
            if p == P:
                assert lib.hbt_mulmod_lowones(ptr(limbs(p)), ptr(limbs(a)), ptr(limbs(b)), ptr(out)) == 0
                assert val(out) == a * b % p
            lib.hbt_addsub(ptr(limbs(p)), ptr(limbs(a)), ptr(limbs(b)), ptr(s), ptr(d), ptr(n))
            assert (val(s), val(d), val(n)) == ((a + b) % p, (a - b) % p, (-a) % p)


def test_mulmod_unreduced_left_operand(lib):
    # mont_mul allows a < 2^256 on one side (used for un-normalised inputs)
    out = np.zeros(4, np.uint64)
    a = 2 ** 256 - 1
    lib.hbt_mulmod(ptr(limbs(P)), ptr(limbs(a % P)), ptr(limbs(P - 1)), ptr(out))
    assert val(out) == a * (P - 1) % P


@pytest.mark.parametrize("p,fold", [(P, 1), (P, 2), (13, 1), (2 ** 127 - 1, 1), (2 ** 127 - 1, 2)])
def test_lazy_dot(lib, p, fold):
    rng = random.Random(fold)
    out = np.zeros(4, np.uint64)
    for n in [0, 1, 2, 3, 6, 16, 43, 128]:
        for mode in ("rand", "max"):
            if mode == "rand":
                a = [rng.randrange(p) for _ in range(n)]
                b = [rng.randrange(p) for _ in range(n)]
            else:
                a = [p - 1] * n
                b = [p - 1] * n
            assert lib.hbt_dot(ptr(limbs(p)), n, ptr(many(a)), ptr(many(b)), fold, ptr(out)) == 0
            assert val(out) == sum(x * y for x, y in zip(a, b)) % p, (n, mode)


@pytest.mark.parametrize("p", [P, 53])
def test_vandermonde_inverse(lib, p):
    rng = random.Random(3)
    for k in [1, 2, 3, 6, 16]:
        xs = rng.sample(range(1, min(p, 10 ** 9)), k) if p < 2 ** 64 else [rng.randrange(p) for _ in range(k)]
        out = np.zeros(4 * k * k, np.uint64)
        assert lib.hbt_vandermonde_inverse(ptr(limbs(p)), k, ptr(many(xs)), ptr(out)) == 0
        inv = [[val(out[4 * (i * k + j): 4 * (i * k + j) + 4]) for j in range(k)] for i in range(k)]
        for i in range(k):
            for j in range(k):
                # (V^-1 V)[i][j] with V[l][j] = xs[l]^j
                acc = sum(inv[i][l] * pow(xs[l], j, p) for l in range(k)) % p
                assert acc == (1 if i == j else 0)
    out = np.zeros(16, np.uint64)
    assert lib.hbt_vandermonde_inverse(ptr(limbs(p)), 2, ptr(many([5, 5])), ptr(out)) == 2


def test_pow_inv(lib):
    rng = random.Random(4)
    pw, inv = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
    for p in (P, 53):
        for _ in range(5):
            a = rng.randrange(1, p)
            e = rng.randrange(2 ** 40)
            lib.hbt_pow_inv(ptr(limbs(p)), ptr(limbs(a)), ctypes.c_uint64(e), ptr(pw), ptr(inv))
            assert val(pw) == pow(a, e, p) and val(inv) == pow(a, -1, p)


def test_bad_modulus(lib):
    out = np.zeros(4, np.uint64)
    assert lib.hbt_mulmod(ptr(limbs(16)), ptr(limbs(1)), ptr(limbs(1)), ptr(out)) == 1
    assert lib.hbt_mulmod(ptr(limbs(1)), ptr(limbs(1)), ptr(limbs(1)), ptr(out)) == 1
