"""The host-side mirror of the reference's modules (ntl shim, codec classes,
IncrementalDecoder, batch_reconstruct, robust_reconstruct) on CPU: the protocol
tests of tests/test_gpu_protocol.py and the small known answers of tests/kats.py,
run with the kernels replaced by the oracle behind ``_native.Context``'s interface
(tests/host_backend.py).  What this covers is the Python logic around the C-ABI --
marshalling, chunking, message rounds, optimistic / robust decoding policy --
which the GPU suite exercises again with the real kernels."""

import importlib

import host_backend
import kats
import pytest
from conftest import BLS12_381_R as P

gp = importlib.import_module("test_gpu_protocol")


@pytest.fixture()
def rs(monkeypatch):
    host_backend.install(monkeypatch)
    from honeybadgermpc_b200 import ntl, reed_solomon

    ntl._ctx(P)
    return reed_solomon


@pytest.fixture()
def ntl(monkeypatch):
    host_backend.install(monkeypatch)
    from honeybadgermpc_b200 import ntl as shim

    return shim


def test_ntl_shim_kats(ntl):
    kats.check_small_kats(ntl)
    kats.check_evaluate(ntl)
    kats.check_fft_interpolate(ntl)
    kats.check_threads(ntl)
    kats.check_errors(ntl)
    kats.check_sqrt(ntl)
    kats.check_fft_properties(ntl)


def test_encoder_decoder_kats(rs):
    gp.test_encoder_decoder_kats(rs)


def test_selectors(rs, monkeypatch):
    gp.test_selectors(rs, monkeypatch)


@pytest.mark.parametrize("n,t,omega", [(4, 1, False), (7, 2, True)])
def test_codec_round_trip(rs, n, t, omega):
    gp.test_codec_round_trip_vs_oracle(rs, n, t, omega)


@pytest.mark.parametrize("omega", [False, True])
@pytest.mark.parametrize("algo", ["gao", "welch-berlekamp"])
def test_incremental_decoder(rs, omega, algo):
    gp.test_incremental_decoder(rs, omega, algo)


def test_batch_reconstruct_kats(rs):
    gp.test_batch_reconstruct_kats()


@pytest.mark.parametrize("n,t,count,omega,algo", [(4, 1, 40, False, "gao"), (7, 2, 25, True, "gao"),
                                                  (7, 2, 12, False, "welch-berlekamp")])
def test_batch_reconstruct_random(rs, n, t, count, omega, algo):
    gp.test_batch_reconstruct_random(n, t, count, omega, algo)


@pytest.mark.parametrize("omega", [False, True])
def test_robust_reconstruct_single_share(rs, omega):
    gp.test_robust_reconstruct_single_share(rs, omega)


def test_opened_polynomial_and_sqrt():
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.polynomial import OpenedPolynomial

    f = GF(P)
    poly = OpenedPolynomial([5, 0, 7, 0], f)
    assert poly == [5, 0, 7, 0] and poly.coeffs == [f(5), f(0), f(7)] and poly.degree() == 2
    assert poly(f(0)) == 5 and poly(3) == 5 + 7 * 9 and type(poly(f(2))).__name__ == "GFElement"
    assert OpenedPolynomial([], f)(f(4)) == 0
    for v in (4, 9, 1234567 ** 2):
        r = f(v).sqrt()
        assert r * r == v
    with pytest.raises(AssertionError):  # like the reference's assertion (field.py:175)
        GF(13)(2).sqrt()  # 2 is not a square mod 13
    assert (~f(3)) * 3 == 1 and f(10) // f(5) == 2 and f(P - 1).signed() == -1 and f(6).bit(1) == 1


def test_offline_callers(rs):
    """randousha / refine_triples / _write_polys compute steps on the limb path (CPU: the host
    logic with the oracle behind the context; tests/test_gpu_protocol.py runs the kernels)"""
    import offline_cases

    offline_cases.check_offline_callers(batch=5)
