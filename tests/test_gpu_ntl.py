"""Parity of the CUDA path (through the C-ABI and the ``honeybadgermpc.ntl``
drop-in) with the CPU oracle, the reference's known answers and the committed
golden fixtures.  Needs a B200: ``pytest -m gpu``.  Bit-exact everywhere."""

import random

import kats
import numpy as np
import pytest
from conftest import BLS12_381_R as P
from conftest import ROOTS_OF_UNITY

from oracle import hbmpc_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ntl():
    from honeybadgermpc_b200 import ntl as m

    m._ctx(P)  # raises NativeLibraryError when the .so or the device is missing
    return m


def test_library_is_loaded(ntl):
    from honeybadgermpc_b200 import _native

    lib = _native.load_library()
    assert b"sm_100a" in lib.hbg_version()
    before = ntl._ctx(P).launch_count()
    ntl.vandermonde_batch_evaluate([1, 2], [[0, 1]], P)
    assert ntl._ctx(P).launch_count() == before + 1


def test_small_kats(ntl):
    kats.check_small_kats(ntl)


def test_fft_properties(ntl):
    for path in ("auto", "matrix", "ntt"):
        ntl._ctx(P).set_fft_path(path)
        kats.check_fft_properties(ntl)
    ntl._ctx(P).set_fft_path("auto")


def test_fft_interpolate(ntl):
    kats.check_fft_interpolate(ntl)


def test_evaluate(ntl):
    kats.check_evaluate(ntl)


def test_sqrt(ntl):
    kats.check_sqrt(ntl)


def test_threads(ntl):
    kats.check_threads(ntl)


def test_errors(ntl):
    kats.check_errors(ntl)


def test_golden(ntl, golden):
    for path in ("auto", "matrix", "ntt", "ntt-smem", "ntt-split", "ntt-bal"):
        ntl._ctx(P).set_fft_path(path)
        kats.check_golden(ntl, golden)
    ntl._ctx(P).set_fft_path("auto")


def test_interpolate_at_zero(ntl, golden):
    for case in golden["interpolate_at_zero"]:
        xs = list(range(1, case["t"] + 2))
        assert ntl.vandermonde_batch_interpolate(xs, [case["shares"]], P)[0][0] == case["secret"]


@pytest.mark.parametrize("p", [P, 13, 53, 2 ** 127 - 1])
@pytest.mark.parametrize("n,d,batch", [(1, 1, 1), (4, 2, 128), (16, 6, 257), (16, 16, 33),
                                       (7, 3, 5), (64, 22, 40), (128, 43, 9), (8, 8, 300), (5, 5, 129),
                                       (9, 7, 64), (3, 1, 7)])
def test_vandermonde_vs_oracle(ntl, p, n, d, batch):
    rng = random.Random(n * 1000 + d)
    if p <= n:
        xs = [rng.randrange(p) for _ in range(n)]  # repeated points are fine for evaluation
    else:
        xs = list(range(1, n + 1))
    polys = [[rng.randrange(p) for _ in range(rng.randint(1, d))] for _ in range(batch)]
    polys[0] = polys[0] + [p - 1] * (d - len(polys[0]))
    want = orc.vandermonde_batch_evaluate(xs, polys, p)
    paths = ("auto", "global", "smem", "small", "small-r29", "no-tc") + (("tc",) if p == P else ())
    for path in paths:
        ntl._ctx(p).set_matvec_path(path)
        assert ntl.vandermonde_batch_evaluate(xs, polys, p) == want, path
        if path == "tc" and n * d <= 96:
            assert ntl._ctx(p).last_kernel() == "tc_apply_kernel"
    if p > n:
        k = d
        xk = rng.sample(xs, k)
        ys = [[rng.randrange(p) for _ in range(k)] for _ in range(batch)]
        want = orc.vandermonde_batch_interpolate(xk, ys, p)
        for path in paths:
            ntl._ctx(p).set_matvec_path(path)
            assert ntl.vandermonde_batch_interpolate(xk, ys, p) == want, path
            if path == "tc" and k * k <= 96:
                assert ntl._ctx(p).last_kernel() == "tc_apply_kernel"
    ntl._ctx(p).set_matvec_path("auto")
    if p != P:
        with pytest.raises(NotImplementedError):  # the tensor-core path is BLS12-381 only
            ntl._ctx(p).set_matvec_path("tc")


def test_worst_case_values(ntl):
    # every operand p-1: the lazy accumulator's carry bounds
    for n, d in [(16, 16), (128, 128)]:
        xs = [P - 1 - i for i in range(n)]
        polys = [[P - 1] * d, [P - 2] * d]
        assert ntl.vandermonde_batch_evaluate(xs, polys, P) == \
            orc.vandermonde_batch_evaluate(xs, polys, P)


@pytest.mark.parametrize("r,d,k,batch", [(1, 2, 2, 3), (2, 3, 4, 70), (4, 6, 16, 300), (4, 16, 11, 65),
                                         (4, 1, 16, 9), (4, 4, 16, 130), (4, 5, 7, 129), (4, 8, 16, 31),
                                         (4, 9, 16, 33), (4, 11, 1, 5), (4, 2, 16, 64), (4, 3, 5, 63),
                                         (4, 7, 16, 129), (4, 6, 6, 1), (4, 8, 3, 200), (4, 5, 16, 1000), (4, 12, 16, 200), (4, 20, 16, 77),
                                         (5, 20, 25, 64), (7, 43, 128, 21), (8, 100, 256, 5),
                                         (10, 700, 1024, 3), (11, 1500, 2048, 2), (12, 4096, 100, 1)])
def test_fft_vs_oracle(ntl, r, d, k, batch):
    rng = random.Random(r * 100 + d)
    n = 2 ** r
    omega = ROOTS_OF_UNITY[r] if r < len(ROOTS_OF_UNITY) else pow(7, (P - 1) // n, P)
    polys = [[rng.randrange(P) for _ in range(d)] for _ in range(batch)]
    want = orc.fft_batch_evaluate(polys, omega, P, n, k)
    for path in ("matrix", "ntt", "ntt-smem", "ntt-split", "ntt-bal"):
        if path == "matrix" and k * min(d, n) > 2 ** 18:
            continue
        ntl._ctx(P).set_fft_path(path)
        assert ntl.fft_batch_evaluate(polys, omega, P, n, k) == want, path
    ntl._ctx(P).set_fft_path("auto")
    assert ntl.fft_batch_evaluate(polys, omega, P, n, k) == want
    ntl._ctx(P).set_matvec_path("tc")  # the matrix form on the tensor cores, where it fits
    assert ntl.fft_batch_evaluate(polys, omega, P, n, k) == want
    ntl._ctx(P).set_fft_path("tc-split")  # ... and its radix-2 split form (butterfly epilogue)
    assert ntl.fft_batch_evaluate(polys, omega, P, n, k) == want
    if n >= 4 and 2 <= d and min(d, n) * min(k, n // 2) <= 48:
        assert ntl._ctx(P).last_kernel() == "tc_apply_kernel"
    ntl._ctx(P).set_fft_path("auto")
    ntl._ctx(P).set_matvec_path("auto")


def test_tensor_core_path(ntl):
    """tc_apply_kernel (exact u8 GEMM on tcgen05 + Barrett epilogue) against the oracle and
    against the IMAD kernels: tile-boundary batch sizes, extreme operands, every block shape
    (outputs per accumulator block 1..8, one and several blocks)."""
    ctx = ntl._ctx(P)
    rng = random.Random(0x7C)
    # the last five need the STREAMED constant operand (it does not fit shared memory):
    # cfg4's 16 x 16, cfg5's 43 x 43, a 64 x 22 encode, a wide 5 x 70, and 128 x 43
    shapes = [(1, 1), (2, 2), (3, 2), (4, 4), (5, 3), (6, 6), (7, 5), (8, 8), (16, 6), (12, 7), (9, 2),
              (24, 4), (6, 10), (2, 16), (32, 3), (16, 16), (43, 43), (64, 22), (5, 70), (128, 43)]
    try:
        for n, d in shapes:
            xs = rng.sample(range(1, 1 << 20), n)
            for batch in ((1, 127, 128, 129, 300, 1500) if n * d <= 256 else (1, 129, 700)):
                polys = [[rng.randrange(P) for _ in range(d)] for _ in range(batch)]
                polys[0] = [P - 1] * d
                polys[-1] = [0] * d
                if batch > 2:
                    polys[1] = [(P - 1) if j % 2 else 1 for j in range(d)]
                ctx.set_matvec_path("tc")
                got = ntl.vandermonde_batch_evaluate(xs, polys, P)
                assert ctx.last_kernel() == "tc_apply_kernel", (n, d)
                # the two store paths of the epilogue (staged full-line stores / 32 bytes per thread)
                # and a launch confined to a few CTAs (hbg_ctx_set_sm_limit)
                ctx.set_tc_store("staged")
                ctx.set_sm_limit(3)
                try:
                    assert got == ntl.vandermonde_batch_evaluate(xs, polys, P), (n, d, batch, "staged stores")
                finally:
                    ctx.set_tc_store("direct")
                    ctx.set_sm_limit(0)
                ctx.set_matvec_path("no-tc")
                assert got == ntl.vandermonde_batch_evaluate(xs, polys, P), (n, d, batch)
                assert ctx.last_kernel() != "tc_apply_kernel"
                if batch <= 129:
                    assert got == orc.vandermonde_batch_evaluate(xs, polys, P), (n, d, batch)
        # auto mode: large batches take the tensor-core kernel, small ones the IMAD kernels
        ctx.set_matvec_path("auto")
        xs = list(range(1, 7))
        ys = [[rng.randrange(P) for _ in range(6)] for _ in range(600)]
        a = ntl.vandermonde_batch_interpolate(xs, ys, P)
        assert ctx.last_kernel() == "tc_apply_kernel"
        b = ntl.vandermonde_batch_interpolate(xs, ys[:100], P)
        assert ctx.last_kernel() != "tc_apply_kernel" and a[:100] == b
        assert a[:40] == orc.vandermonde_batch_interpolate(xs, ys[:40], P)
        # non-canonical limbs (>= p, up to 2^256 - 1) reduce like to_ZZ_p would
        raw = np.full((256, 6, 4), 2 ** 64 - 1, dtype=np.uint64)
        raw[::2, :, 3] = np.uint64(P >> 192)
        xl = ntl.pack_vec(xs, P)
        ctx.set_matvec_path("tc")
        got = ntl.unpack_rows(ntl.vandermonde_batch_interpolate_limbs(xl, raw, P))
        want = orc.vandermonde_batch_interpolate(xs, [[v % P for v in row] for row in ntl.unpack_rows(raw)], P)
        assert got == want
        # the narrow fold of the epilogue at its limit (K = 256 bytes: d = 8; every input byte 0xFF
        # makes the column sums as large as they get) and the first shape of the wide fold (d = 9)
        for d in (8, 9):
            xs = rng.sample(range(1, 1 << 30), d)
            raw = np.full((300, d, 4), 2 ** 64 - 1, dtype=np.uint64)
            raw[1::3] = rng.getrandbits(63)
            got = ntl.unpack_rows(ntl.vandermonde_batch_interpolate_limbs(ntl.pack_vec(xs, P), raw, P))
            assert ctx.last_kernel() == "tc_apply_kernel"
            rows = [0, 1, 2, 128, 299]
            want = orc.vandermonde_batch_interpolate(
                xs, [[v % P for v in row] for row in ntl.unpack_rows(raw[rows])], P)
            assert [got[i] for i in rows] == want, d
    finally:
        ctx.set_matvec_path("auto")


@pytest.mark.parametrize("r,k,batch", [(4, 6, 70), (4, 16, 9), (7, 43, 40), (7, 128, 5), (10, 342, 6),
                                       (10, 129, 3), (12, 700, 2), (3, 1, 4), (5, 2, 3)])
def test_fnt_interpolation_path(ntl, r, k, batch):
    """The NTT-structured interpolation (fnt_decode_step1/2, rsdecode_impl.h:194-265: scale by
    1/A'(x_i), scatter, inverse NTT, MulTrunc by A) against the oracle's restatement of the same
    two steps and against the V^-1 matrix path: identical bits.  k = 342 / n = 1024 and
    k = 700 / n = 4096 never form a k x k inverse."""
    rng = random.Random(r * 1000 + k)
    n = 2 ** r
    omega = ROOTS_OF_UNITY[r] if r < len(ROOTS_OF_UNITY) else pow(7, (P - 1) // n, P)
    zs = rng.sample(range(n), k)
    ys = [[rng.randrange(P) for _ in range(k)] for _ in range(batch)]
    ys[0] = [P - 1] * k
    ctx = ntl._ctx(P)
    try:
        ctx.set_interp_path("fnt")
        got = ntl.fft_batch_interpolate(zs, ys, omega, P, n)
        assert ctx.last_kernel() == "fnt_decode_step2"
        want = orc.fft_batch_interpolate(zs, ys[: min(batch, 3)], omega, P, n)
        assert got[: len(want)] == want
        if k <= 342:
            ctx.set_interp_path("matrix")
            assert ntl.fft_batch_interpolate(zs, ys, omega, P, n) == got
        ctx.set_interp_path("auto")
        assert ntl.fft_batch_interpolate(zs, ys, omega, P, n) == got
        assert (ctx.last_kernel() == "fnt_decode_step2") == (k > 128)
        # the interpolant really passes through the points
        xs = [pow(omega, z, P) for z in zs[:4]]
        for row, y in list(zip(got, ys))[:2]:
            assert [orc.poly_eval(row, x, P) for x in xs] == y[:4]
    finally:
        ctx.set_interp_path("auto")


def test_fnt_interpolation_other_fields(ntl):
    # p = 257: every power-of-two order up to 256 exists; p = 13: order 4
    rng = random.Random(2)
    for p, w, n in ((257, 3, 256), (257, pow(3, 16, 257), 16), (13, 5, 4), (97, 8, 16)):
        for k in (1, 2, min(n, 7), n // 2):
            if k < 1:
                continue
            zs = rng.sample(range(n), k)
            ys = [[rng.randrange(p) for _ in range(k)] for _ in range(5)]
            ctx = ntl._ctx(p)
            want = orc.fft_batch_interpolate(zs, ys, w, p, n)
            try:
                ctx.set_interp_path("fnt")
                try:
                    got = ntl.fft_batch_interpolate(zs, ys, w, p, n)
                except NotImplementedError:
                    # no root of unity of order >= 2k in this field: the automatic path must
                    # fall back to the matrix silently
                    ctx.set_interp_path("auto")
                    got = ntl.fft_batch_interpolate(zs, ys, w, p, n)
                assert got == want, (p, n, k)
            finally:
                ctx.set_interp_path("auto")
    # repeated z: the reference divides by zero (inv(0) in fnt_decode_step1)
    ntl._ctx(P).set_interp_path("fnt")
    try:
        with pytest.raises(ZeroDivisionError):
            ntl.fft_batch_interpolate([1, 1], [[1, 2]], ROOTS_OF_UNITY[2], P, 4)
    finally:
        ntl._ctx(P).set_interp_path("auto")


def test_fft_small_prime(ntl):
    # p = 13: omega = 5 has order 4 (tests/test_ntl.py:57-68); p = 257: order 256
    for path in ("matrix", "ntt"):
        ntl._ctx(13).set_fft_path(path)
        assert ntl.fft([0, 1], 5, 13, 4) == [1, 5, 12, 8]
        assert ntl.fft([3, 1, 4, 1, 5, 9], 5, 13, 4) == orc.fft([3, 1, 4, 1, 5, 9], 5, 13, 4)
    rng = random.Random(1)
    c = [rng.randrange(257) for _ in range(200)]
    for path in ("matrix", "ntt"):
        ntl._ctx(257).set_fft_path(path)
        assert ntl.fft(c, 3, 257, 256) == orc.fft(c, 3, 257, 256)
    # n = 16 over small generic fields (FieldAny instantiations of the two n = 16 kernels)
    for p, w in ((17, 3), (97, 8), (2 ** 64 - 2 ** 32 + 1, pow(7, (2 ** 64 - 2 ** 32) // 16, 2 ** 64 - 2 ** 32 + 1))):
        for d in (1, 4, 5, 6, 8, 13):
            polys = [[rng.randrange(p) for _ in range(d)] for _ in range(70)] + [[p - 1] * d]
            want = orc.fft_batch_evaluate(polys, w, p, 16, 16)
            for path in ("ntt", "ntt-split", "ntt-bal", "ntt-smem", "matrix"):
                ntl._ctx(p).set_fft_path(path)
                assert ntl.fft_batch_evaluate(polys, w, p, 16, 16) == want, (p, d, path)
            ntl._ctx(p).set_fft_path("auto")
    with pytest.raises(ValueError):
        ntl.fft([1, 2], 4, 13, 4)  # 4 is not a primitive 4th root mod 13


@pytest.mark.parametrize("r,k,batch", [(3, 3, 7), (4, 6, 500), (6, 22, 30), (7, 43, 12), (7, 128, 3)])
def test_fft_interpolate_vs_oracle(ntl, r, k, batch):
    rng = random.Random(r * 10 + k)
    n = 2 ** r
    omega = ROOTS_OF_UNITY[r]
    zs = rng.sample(range(n), k)
    ys = [[rng.randrange(P) for _ in range(k)] for _ in range(batch)]
    got = ntl.fft_batch_interpolate(zs, ys, omega, P, n)
    assert got == orc.fft_batch_interpolate(zs, ys, omega, P, n)
    xs = [pow(omega, z, P) for z in zs]
    assert got == ntl.vandermonde_batch_interpolate(xs, ys, P)
    assert ntl.fft_interpolate(zs, ys[0], omega, P, n) == got[0]


def test_round_trip_full_size(ntl):
    """BASELINE config 2 at full size (n=16, t=5, 65 536 polynomials) on the
    limb-array boundary: interpolate(encode(c)) == c for two z sets, plus
    linearity of the encoder (size-independent properties)."""
    n, k, batch = 16, 6, 65536
    pt = orc.EvalPoint(P, n, True)
    rng = np.random.default_rng(0xB202)
    c = rng.integers(0, 2 ** 63, size=(batch, k, 4), dtype=np.uint64)
    c[:, :, 3] >>= np.uint64(2)  # < 2^253 < p: canonical
    omega = ntl.pack_vec([pt.omega], P)[0]
    enc = ntl.fft_batch_evaluate_limbs(c, omega, P, pt.order, n)
    for zs in ([0, 1, 2, 3, 4, 5], [1, 3, 4, 9, 12, 15]):
        ys = np.ascontiguousarray(enc[:, zs, :])
        dec = ntl.fft_batch_interpolate_limbs(zs, ys, omega, P, pt.order)
        assert np.array_equal(dec, c)
    # spot-check rows against the oracle
    rows = [0, 1, 4095, 65535]
    ints = ntl.unpack_rows(c[rows])
    assert ntl.unpack_rows(enc[rows]) == orc.fft_batch_evaluate(ints, pt.omega, P, pt.order, n)
    # the Vandermonde path computes the same map
    xs = ntl.pack_vec([pt(i) for i in range(n)], P)
    assert np.array_equal(ntl.vandermonde_batch_evaluate_limbs(xs, c, P), enc)


def test_empty_and_ragged(ntl):
    # ragged rows are zero-padded to the longest (pyx:217,232-233): d = 1 here
    assert ntl.vandermonde_batch_evaluate([1, 2, 3], [[], [1]], P) == [[0, 0, 0], [1, 1, 1]]
    assert orc.vandermonde_batch_evaluate([1, 2, 3], [[], [1]], P) == [[0, 0, 0], [1, 1, 1]]
    assert ntl.vandermonde_batch_evaluate([1, 2, 3], [[7], [1, 1]], P) == [[7, 7, 7], [2, 3, 4]]
    assert ntl.evaluate([], 5, P) == 0
    assert ntl.lagrange_interpolate([], [], P) == []
    assert ntl.fft([], ROOTS_OF_UNITY[2], P, 4) == [0, 0, 0, 0]


def test_fused_allgather_world1(ntl):
    """hbg_fft_batch_interpolate_allgather with a single rank: the block lands at
    rank*batch of the 'gathered' buffer (the multi-rank form is exercised by
    bench.py --gpus N, which asserts the same on every rank)."""
    import torch

    n, k, batch = 16, 6, 1000
    pt = orc.EvalPoint(P, n, True)
    rng = np.random.default_rng(3)
    c = rng.integers(0, 2 ** 62, size=(batch, k, 4), dtype=np.uint64)
    omega = ntl.pack_vec([pt.omega], P)[0]
    enc = ntl.fft_batch_evaluate_limbs(c, omega, P, pt.order, n)
    zs = [0, 2, 5, 7, 11, 13]
    ys = torch.from_numpy(np.ascontiguousarray(enc[:, zs, :]).view(np.int64)).cuda()
    out = torch.zeros((batch, k, 4), dtype=torch.int64, device="cuda")
    ctx = ntl._ctx(P)
    torch.cuda.synchronize()
    ctx.fft_batch_interpolate_allgather(omega, pt.order, zs, ys.data_ptr(), batch, [out.data_ptr()], 0, 0)
    ctx.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), c)


def test_config5_shard_round_trip(ntl):
    """BASELINE config 5 (n=128, t=42) on one shard-sized batch: NTT-128 encode,
    interpolate from 43 scattered points, all kernels' variants agree."""
    n, k, batch = 128, 43, 8192
    pt = orc.EvalPoint(P, n, True)
    rng = np.random.default_rng(0xB205)
    c = rng.integers(0, 2 ** 62, size=(batch, k, 4), dtype=np.uint64)
    omega = ntl.pack_vec([pt.omega], P)[0]
    enc = ntl.fft_batch_evaluate_limbs(c, omega, P, pt.order, n)
    zs = sorted(random.Random(5).sample(range(n), k))
    ys = np.ascontiguousarray(enc[:, zs, :])
    for path in ("auto", "global", "smem", "small-r29", "tc"):
        ntl._ctx(P).set_matvec_path(path)
        assert np.array_equal(ntl.fft_batch_interpolate_limbs(zs, ys, omega, P, pt.order), c), path
    ntl._ctx(P).set_matvec_path("auto")
    rows = [0, 17, batch - 1]
    assert ntl.unpack_rows(enc[rows]) == orc.fft_batch_evaluate(ntl.unpack_rows(c[rows]), pt.omega, P, pt.order, n)
    ntl._ctx(P).set_fft_path("matrix")
    assert np.array_equal(ntl.fft_batch_evaluate_limbs(c[:512], omega, P, pt.order, n), enc[:512])
    ntl._ctx(P).set_fft_path("auto")


def test_host_async_mode(ntl):
    """hbg_ctx_set_host_async: calls only enqueue; results are valid after synchronize;
    more calls in flight than staging slots still complete in order."""
    n, k, batch = 16, 6, 40000  # > 4 MB per call: the chunked pipeline
    pt = orc.EvalPoint(P, n, True)
    omega = ntl.pack_vec([pt.omega], P)[0]
    rng = np.random.default_rng(11)
    ctx = ntl._ctx(P)
    cs = [rng.integers(0, 2 ** 62, size=(batch, k, 4), dtype=np.uint64) for _ in range(6)]
    want = [ntl.fft_batch_evaluate_limbs(c, omega, P, pt.order, n) for c in cs]
    outs = [np.zeros((batch, n, 4), dtype=np.uint64) for _ in cs]
    ctx.set_host_async(True)
    try:
        for c, o in zip(cs, outs):
            ctx.fft_batch_evaluate(omega, pt.order, c, batch, k, n, o)
        ctx.synchronize()
    finally:
        ctx.set_host_async(False)
    for o, w in zip(outs, want):
        assert np.array_equal(o, w)
    # hbg_ctx_wait_pending: after enqueueing call i, every call older than the newest `keep` is
    # complete in host memory -- checked call by call while later calls are still in flight
    outs = [np.zeros((batch, n, 4), dtype=np.uint64) for _ in cs]
    ctx.set_host_async(True)
    try:
        for i, (c, o) in enumerate(zip(cs, outs)):
            ctx.fft_batch_evaluate(omega, pt.order, c, batch, k, n, o)
            ctx.wait_pending(2)
            if i >= 2:
                assert np.array_equal(outs[i - 2], want[i - 2]), i
        ctx.wait_pending(0)
        assert np.array_equal(outs[-1], want[-1]) and np.array_equal(outs[-2], want[-2])
        with pytest.raises(Exception):
            ctx.wait_pending(4)
    finally:
        ctx.set_host_async(False)


def test_c_abi_error_paths(ntl):
    """the C-ABI returns codes (never throws, never crashes) on bad arguments"""
    import ctypes

    from honeybadgermpc_b200 import _native

    lib = _native.load_library()
    ctx = ntl._ctx(P)
    h = ctx.handle
    good = ntl.pack_vec([1, 2, 3], P)
    polys = ntl.pack_rows([[1, 2]], 2, P)
    out = np.zeros((1, 3, 4), np.uint64)
    # null pointers / bad sizes / bad mem flag
    assert lib.hbg_vandermonde_batch_evaluate(h, None, 3, polys.ctypes.data, 1, 2, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    assert lib.hbg_vandermonde_batch_evaluate(h, good.ctypes.data, 3, None, 1, 2, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    assert lib.hbg_vandermonde_batch_evaluate(h, good.ctypes.data, 3, polys.ctypes.data, 1, 2, out.ctypes.data, 7) == _native.HBG_ERR_INVALID
    assert lib.hbg_vandermonde_batch_evaluate(h, good.ctypes.data, -1, polys.ctypes.data, 1, 2, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    assert b"" != lib.hbg_ctx_last_error(h)
    # non-canonical evaluation point
    bad = ntl.pack_vec([1, 2, 3], P).copy()
    bad[2] = np.frombuffer((P + 1).to_bytes(32, "little"), dtype=np.uint64)
    assert lib.hbg_vandermonde_batch_evaluate(h, bad.ctypes.data, 3, polys.ctypes.data, 1, 2, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    # repeated points -> singular
    rep = ntl.pack_vec([5, 5], P)
    assert lib.hbg_vandermonde_batch_interpolate(h, rep.ctypes.data, 2, polys.ctypes.data, 1, out.ctypes.data, 0) == _native.HBG_ERR_SINGULAR
    # fft: size not a power of two, omega of the wrong order, k_out > n
    w4 = ntl.pack_vec([ROOTS_OF_UNITY[2]], P)
    assert lib.hbg_fft_batch_evaluate(h, w4.ctypes.data, 6, polys.ctypes.data, 1, 2, 3, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    assert lib.hbg_fft_batch_evaluate(h, w4.ctypes.data, 8, polys.ctypes.data, 1, 2, 3, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    assert lib.hbg_fft_batch_evaluate(h, w4.ctypes.data, 4, polys.ctypes.data, 1, 2, 5, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    zs = np.array([0, 0], dtype=np.int32)
    assert lib.hbg_fft_batch_interpolate(h, w4.ctypes.data, 4, zs.ctypes.data, 2, polys.ctypes.data, 1, out.ctypes.data, 0) == _native.HBG_ERR_SINGULAR
    zs = np.array([0, 9], dtype=np.int32)
    assert lib.hbg_fft_batch_interpolate(h, w4.ctypes.data, 4, zs.ctypes.data, 2, polys.ctypes.data, 1, out.ctypes.data, 0) == _native.HBG_ERR_INVALID
    # contexts: even modulus / too large modulus / bad device
    hh = ctypes.c_void_p()
    for mod, want in ((16, _native.HBG_ERR_INVALID), (2 ** 255 + 95, _native.HBG_ERR_UNSUPPORTED)):
        limbs = np.frombuffer(mod.to_bytes(32, "little"), dtype=np.uint64).copy()
        assert lib.hbg_ctx_create(ctypes.byref(hh), limbs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 0) == want
    limbs = np.frombuffer(P.to_bytes(32, "little"), dtype=np.uint64).copy()
    assert lib.hbg_ctx_create(ctypes.byref(hh), limbs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 99) == _native.HBG_ERR_CUDA
    # the context still works afterwards
    assert ntl.vandermonde_batch_evaluate([1, 2, 3], [[1, 2]], P) == [[3, 5, 7]]


def test_empty_batches(ntl):
    z = np.zeros((0, 6, 4), np.uint64)
    omega = ntl.pack_vec([ROOTS_OF_UNITY[4]], P)[0]
    assert ntl.fft_batch_evaluate_limbs(z, omega, P, 16, 16).shape == (0, 16, 4)
    assert ntl.fft_batch_interpolate_limbs([0, 1, 2, 3, 4, 5], z, omega, P, 16).shape == (0, 6, 4)
    xs = ntl.pack_vec([1, 2, 3, 4, 5, 6], P)
    assert ntl.vandermonde_batch_evaluate_limbs(xs, z, P).shape == (0, 6, 4)
    assert ntl.vandermonde_batch_interpolate_limbs(xs, z, P).shape == (0, 6, 4)
