"""CPU unit tests of the host twins of the device field arithmetic
(honeybadgermpc_b200/csrc/fp256.cuh, host_math.hpp): the same algorithm text
the kernels run (even/odd CIOS multiplier, the BLS fast reduction rows, the
lazy dot-product accumulator), checked against Python ints."""

import ctypes
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


@pytest.mark.parametrize("p,omega", [(P, None), (17, 3), (97, 8), (2 ** 64 - 2 ** 32 + 1, None)])
def test_ntt16_row_math(lib, p, omega):
    # ntt16_half (the row math of ntt16_g4_kernel) against the DFT definition
    if omega is None:
        g = next(g for g in range(2, 50) if pow(g, (p - 1) // 2, p) != 1)
        omega = pow(g, (p - 1) // 16, p)
    assert pow(omega, 8, p) == p - 1
    rng = random.Random(16)
    out = np.zeros(64, np.uint64)
    for d in range(0, 9):
        for mode in ("rand", "max", "one"):
            c = [rng.randrange(p) if mode == "rand" else (p - 1 if mode == "max" else 1) for _ in range(d)]
            want = [sum(c[i] * pow(omega, i * k, p) for i in range(d)) % p for k in range(16)]
            for lowones in ((0, 1) if p == P else (0,)):
                for bal in (0, 1):  # with / without the even thread's hand-over to the odd thread
                    out[:] = 0xDEAD
                    rc = lib.hbt_ntt16(ptr(limbs(p)), d, ptr(many(c)), ptr(limbs(omega)), lowones, bal, ptr(out))
                    assert rc == 0
                    got = [val(out[4 * k: 4 * k + 4]) for k in range(16)]
                    assert got == want, (d, mode, lowones, bal)


@pytest.mark.parametrize("p", [P, 13, 53, 2 ** 61 - 1, 2 ** 127 - 1, 2 ** 255 - 19])
def test_lazy_dot_radix29(lib, p):
    # to_limbs29 / mac29 / norm29 / redc29 (the row math of interp_small_kernel)
    rng = random.Random(29)
    out = np.zeros(4, np.uint64)
    for n in range(0, 9):
        for mode in ("rand", "max", "mixed"):
            if mode == "rand":
                a = [rng.randrange(p) for _ in range(n)]
                b = [rng.randrange(p) for _ in range(n)]
            elif mode == "max":
                a, b = [p - 1] * n, [p - 1] * n
            else:
                a = [rng.choice([0, 1, p - 1, 2 ** 29 - 1, 2 ** 232 % p]) for _ in range(n)]
                b = [rng.choice([0, 1, p - 1, (p - 1) // 2]) for _ in range(n)]
            want = sum(x * y for x, y in zip(a, b)) % p
            for lowones in ((0, 1) if p == P else (0,)):
                for prenorm in ((0, 1) if n <= 6 else (1,)):
                    rc = lib.hbt_dot29(ptr(limbs(p)), n, ptr(many(a)), ptr(many(b)), lowones, prenorm, ptr(out))
                    assert rc == 0
                    assert val(out) == want, (n, mode, lowones, prenorm)


def test_bad_modulus(lib):
    out = np.zeros(4, np.uint64)
    assert lib.hbt_mulmod(ptr(limbs(16)), ptr(limbs(1)), ptr(limbs(1)), ptr(out)) == 1
    assert lib.hbt_mulmod(ptr(limbs(1)), ptr(limbs(1)), ptr(limbs(1)), ptr(out)) == 1
