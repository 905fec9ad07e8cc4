"""A stand-in for ``_native.Context`` that computes with the Python oracle on HOST numpy
buffers.  Test infrastructure only: it lets the ``-m "not gpu"`` suite run the whole
host-side mirror (ntl shim, codec classes, IncrementalDecoder, batch_reconstruct,
robust_reconstruct) on a machine without a GPU, with the kernels replaced by the oracle
behind the same method signatures and status conventions (include/hbmpc_b200.h)."""

import numpy as np

from honeybadgermpc_b200 import _native

from oracle import hbmpc_oracle as orc


def _ints(arr):
    """uint64[..., 4] -> nested lists of ints (last axis consumed)"""
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    raw = a.tobytes()
    flat = [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]
    if a.ndim == 2:
        return flat
    w = a.shape[1]
    return [flat[i * w:(i + 1) * w] for i in range(a.shape[0])]


def _store(out, rows):
    """write rows (lists of ints, possibly shorter than the row) into uint64[batch, width, 4]"""
    out[...] = 0
    for i, r in enumerate(rows):
        for j, v in enumerate(r):
            out[i, j] = np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint64)


class OracleContext:
    def __init__(self, modulus):
        self.p = int(modulus)
        self.launches = 0

    # plumbing used by the shim / tests
    def launch_count(self):
        return self.launches

    def set_fft_path(self, path):
        pass

    def set_matvec_path(self, path):
        pass

    def vandermonde_batch_evaluate(self, xs, polys, batch, d, out, mem=0):
        self.launches += 1
        if batch:
            _store(out, orc.vandermonde_batch_evaluate(_ints(xs), _ints(polys), self.p))

    def vandermonde_batch_interpolate(self, xs, ys, batch, out, mem=0):
        self.launches += 1
        x = _ints(xs)
        if len(set(x)) != len(x):
            raise _native.SingularError("singular Vandermonde matrix")
        if batch:
            _store(out, orc.vandermonde_batch_interpolate(x, _ints(ys), self.p))

    def fft_batch_evaluate(self, omega, n, polys, batch, d, k_out, out, mem=0):
        self.launches += 1
        w = _ints(np.asarray(omega).reshape(1, 4))[0]
        if n > 1 and pow(w, n // 2, self.p) != self.p - 1:
            raise ValueError("omega is not a primitive n-th root of unity")
        if batch:
            _store(out, orc.fft_batch_evaluate(_ints(polys), w, self.p, n, k_out))

    def fft_batch_interpolate(self, omega, n, zs, ys, batch, out, mem=0):
        self.launches += 1
        w = _ints(np.asarray(omega).reshape(1, 4))[0]
        zs = [int(z) for z in zs]
        if len(set(zs)) != len(zs):
            raise _native.SingularError("repeated z")
        if batch:
            _store(out, orc.fft_batch_interpolate(zs, _ints(ys), w, self.p, n))

    def interpolate_reencode(self, xs_k, xs_all, ys, batch, out, mem=0):
        self.launches += 1
        xk, xa = _ints(xs_k), _ints(xs_all)
        if len(set(xk)) != len(xk):
            raise _native.SingularError("singular Vandermonde matrix")
        if batch:
            coeffs = orc.vandermonde_batch_interpolate(xk, _ints(ys), self.p)
            ev = orc.vandermonde_batch_evaluate(xa, coeffs, self.p) if xa else [[] for _ in coeffs]
            _store(out, [c + e for c, e in zip(coeffs, ev)])

    def gao_decode_batch(self, xs, k, ys, batch, coeffs, locator, loc_stride, loc_len, status, mem=0):
        self.launches += 1
        x = _ints(xs)
        if len(set(x)) != len(x):
            raise _native.SingularError("repeated point")
        coeffs[...] = 0
        locator[...] = 0
        for b, row in enumerate(_ints(ys)):
            dec, loc = orc.gao_interpolate(x, row, k, self.p)
            if dec is None:
                status[b], loc_len[b] = 1, 0
                continue
            status[b], loc_len[b] = 0, len(loc)
            _store(coeffs[b:b + 1], [dec])
            _store(locator[b:b + 1], [loc])

    def wb_decode_batch(self, xs, k, e_max, ys, batch, coeffs, out_len, status, mem=0):
        self.launches += 1
        x = _ints(xs)
        coeffs[...] = 0
        for b, row in enumerate(_ints(ys)):
            try:
                q, e = orc.wb_solve_system(list(zip(x, row)), k, self.p, e_max)
                quo, rem = orc.poly_divrem(q, e, self.p)
                assert not rem
                status[b], out_len[b] = 0, len(quo)
                _store(coeffs[b:b + 1], [quo])
            except ValueError:
                status[b], out_len[b] = 1, 0
            except ZeroDivisionError:
                status[b], out_len[b] = 3, 0
            except Exception as exc:  # noqa: BLE001 - "No solution" of some_solution
                if str(exc) != "No solution":
                    raise
                status[b], out_len[b] = 2, 0


def install(monkeypatch):
    """route every ``_native.get_context`` of the host mirror to the oracle"""
    cache = {}

    def get_context(modulus, device=None):
        return cache.setdefault(int(modulus), OracleContext(modulus))

    monkeypatch.setattr(_native, "get_context", get_context)
    return cache
