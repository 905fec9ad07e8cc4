"""ctypes wrapper of oracle/cpu_ref.cpp (the C++ CPU restatement of the
reference's NTL path).  TEST INFRASTRUCTURE ONLY -- imported by tests/ and by
bench.py's cpu_baseline / --impl reference legs, never by the product."""

import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhbmpc_cpuref.so")


def _limbs(values, p):
    buf = b"".join((int(v) % p).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype=np.uint64).reshape(len(values), 4).copy()


def _rows(rows, p):
    w = len(rows[0])
    return _limbs([v for row in rows for v in row], p).reshape(len(rows), w, 4)


def _ints(arr):
    b = np.ascontiguousarray(arr).tobytes()
    flat = [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]
    w = arr.shape[1]
    return [flat[i:i + w] for i in range(0, len(flat), w)]


class CpuRef:
    def __init__(self):
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: run __graft_entry__.build_oracle()")
        self.lib = ctypes.CDLL(LIB)
        self.lib.cpuref_max_threads.restype = ctypes.c_int
        V = ctypes.c_void_p
        self.lib.cpuref_vandermonde_batch_evaluate.argtypes = [
            V, V, ctypes.c_int, V, ctypes.c_size_t, ctypes.c_int, V, ctypes.c_int]
        self.lib.cpuref_vandermonde_batch_interpolate.argtypes = [
            V, V, ctypes.c_int, V, ctypes.c_size_t, V, ctypes.c_int]
        self.lib.cpuref_fft_batch_evaluate.argtypes = [
            V, V, ctypes.c_int, V, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, V, ctypes.c_int]
        self.lib.cpuref_fft_batch_interpolate.argtypes = [
            V, V, ctypes.c_int, V, ctypes.c_int, V, ctypes.c_size_t, V, ctypes.c_int]

    def max_threads(self):
        return self.lib.cpuref_max_threads()

    # limb boundary -------------------------------------------------------
    def vandermonde_batch_evaluate_limbs(self, xs, polys, p, threads=0):
        pl = _limbs([p], 2 ** 256)
        x = _limbs(xs, p)
        polys = np.ascontiguousarray(polys)
        out = np.empty((polys.shape[0], len(xs), 4), np.uint64)
        rc = self.lib.cpuref_vandermonde_batch_evaluate(
            pl.ctypes.data, x.ctypes.data, len(xs), polys.ctypes.data, polys.shape[0],
            polys.shape[1], out.ctypes.data, threads)
        assert rc == 0
        return out

    def vandermonde_batch_interpolate_limbs(self, xs, ys, p, threads=0):
        pl = _limbs([p], 2 ** 256)
        x = _limbs(xs, p)
        ys = np.ascontiguousarray(ys)
        out = np.empty((ys.shape[0], len(xs), 4), np.uint64)
        rc = self.lib.cpuref_vandermonde_batch_interpolate(
            pl.ctypes.data, x.ctypes.data, len(xs), ys.ctypes.data, ys.shape[0], out.ctypes.data,
            threads)
        if rc == 2:
            raise ZeroDivisionError("singular Vandermonde matrix")
        assert rc == 0
        return out

    def fft_batch_evaluate_limbs(self, polys, omega, p, n, k, threads=0):
        pl = _limbs([p], 2 ** 256)
        w = _limbs([omega], p)
        polys = np.ascontiguousarray(polys)
        out = np.empty((polys.shape[0], k, 4), np.uint64)
        rc = self.lib.cpuref_fft_batch_evaluate(
            pl.ctypes.data, w.ctypes.data, n, polys.ctypes.data, polys.shape[0], polys.shape[1], k,
            out.ctypes.data, threads)
        assert rc == 0
        return out

    def fft_batch_interpolate_limbs(self, zs, ys, omega, p, n, threads=0):
        pl = _limbs([p], 2 ** 256)
        w = _limbs([omega], p)
        z = np.ascontiguousarray(zs, dtype=np.int32)
        ys = np.ascontiguousarray(ys)
        out = np.empty((ys.shape[0], len(zs), 4), np.uint64)
        rc = self.lib.cpuref_fft_batch_interpolate(
            pl.ctypes.data, w.ctypes.data, n, z.ctypes.data, len(zs), ys.ctypes.data, ys.shape[0],
            out.ctypes.data, threads)
        if rc == 2:
            raise ZeroDivisionError("repeated z")
        assert rc == 0
        return out

    # int-list boundary -----------------------------------------------------
    def vandermonde_batch_evaluate(self, xs, polys, p):
        return _ints(self.vandermonde_batch_evaluate_limbs(xs, _rows(polys, p), p))

    def vandermonde_batch_interpolate(self, xs, ys, p):
        return _ints(self.vandermonde_batch_interpolate_limbs(xs, _rows(ys, p), p))

    def fft_batch_evaluate(self, polys, omega, p, n, k):
        return _ints(self.fft_batch_evaluate_limbs(_rows(polys, p), omega, p, n, k))

    def fft_batch_interpolate(self, zs, ys, omega, p, n):
        return _ints(self.fft_batch_interpolate_limbs(zs, _rows(ys, p), omega, p, n))
