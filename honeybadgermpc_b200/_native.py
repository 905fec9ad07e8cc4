"""ctypes binding of ``libhbmpc_b200.so`` (C-ABI in ``include/hbmpc_b200.h``).

This is the only way the package computes anything: if the shared library is
missing, or there is no CUDA device, the calls raise -- there is no CPU
fallback (and nothing here may import ``oracle/``).

Field elements cross this boundary as rows of 4 little-endian ``uint64`` limbs
(numpy arrays of shape ``[..., 4]``) in host memory, or as raw device pointers.
"""

import ctypes
import os
import threading

import numpy as np

HBG_OK = 0
HBG_ERR_INVALID = 1
HBG_ERR_SINGULAR = 2
HBG_ERR_CUDA = 3
HBG_ERR_UNSUPPORTED = 4
HBG_ERR_NOMEM = 5

MEM_HOST = 0
MEM_DEVICE = 1

LIB_NAME = "libhbmpc_b200.so"
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, LIB_NAME)

_lib = None
_lib_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing or unusable.  Never swallowed."""


class SingularError(Exception):
    """HBG_ERR_SINGULAR: repeated evaluation points."""


_u64p = ctypes.POINTER(ctypes.c_uint64)
_i32p = ctypes.POINTER(ctypes.c_int32)

# name -> (restype, argtypes); must list every symbol of include/hbmpc_b200.h
SIGNATURES = {
    "hbg_version": (ctypes.c_char_p, []),
    "hbg_ctx_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), _u64p, ctypes.c_int]),
    "hbg_ctx_destroy": (None, [ctypes.c_void_p]),
    "hbg_ctx_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "hbg_ctx_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "hbg_ctx_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "hbg_ctx_set_host_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_ctx_launch_count": (ctypes.c_uint64, [ctypes.c_void_p]),
    "hbg_ctx_last_kernel": (ctypes.c_char_p, [ctypes.c_void_p]),
    "hbg_ctx_set_fft_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_ctx_set_matvec_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_vandermonde_batch_evaluate": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
         ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "hbg_vandermonde_batch_interpolate": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
         ctypes.c_void_p, ctypes.c_int]),
    "hbg_fft_batch_evaluate": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]),
    "hbg_fft_batch_interpolate": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]),
    "hbg_fft_batch_interpolate_allgather": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_int]),
    "hbg_allgather_block": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_size_t, ctypes.c_int, ctypes.c_int]),
    "hbg_gao_decode_batch": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
         ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_int]),
    "hbg_wb_decode_batch": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
         ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_int]),
    "hbg_ctx_set_cache_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "hbg_ctx_set_sm_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_ctx_wait_pending": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_ctx_set_tc_store": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_ctx_set_interp_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_ctx_set_wb_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "hbg_allgather_block_signal": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
         ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "hbg_allgather_block_ce": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
         ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "hbg_allgather_block_bulk": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "hbg_gather_wait": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
         ctypes.c_int]),
    "hbg_gather_fence": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
         ctypes.c_int, ctypes.c_int]),
    "hbg_gather_release": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "hbg_columns_to_rows": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "hbg_interpolate_reencode": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
         ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]),
    "hbg_compare_columns": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
         ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
}


def load_library():
    """dlopen the in-tree library and bind every declared symbol."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  honeybadgermpc_b200 has no CPU fallback.")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # missing libcudart etc.
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def _ptr(a):
    """Host numpy array or integer device pointer -> void*."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


def int_to_limbs(v):
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint64).copy()


class Context:
    """One modulus bound to one CUDA device (``hbg_ctx``)."""

    def __init__(self, modulus, device=0):
        self.lib = load_library()
        self.modulus = int(modulus)
        self.device = device
        if self.modulus < 3 or self.modulus % 2 == 0 or self.modulus >> 255:
            raise ValueError("modulus must be odd, >= 3 and below 2**255")
        handle = ctypes.c_void_p()
        limbs = int_to_limbs(self.modulus)
        rc = self.lib.hbg_ctx_create(ctypes.byref(handle), limbs.ctypes.data_as(_u64p), device)
        if rc != HBG_OK:
            raise NativeLibraryError(
                f"hbg_ctx_create failed with code {rc} (no usable CUDA device {device}?); "
                "honeybadgermpc_b200 has no CPU fallback")
        self.handle = handle

    def close(self):
        if getattr(self, "handle", None):
            self.lib.hbg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def _check(self, rc):
        if rc == HBG_OK:
            return
        msg = self.lib.hbg_ctx_last_error(self.handle).decode()
        if rc == HBG_ERR_SINGULAR:
            raise SingularError(msg)
        if rc == HBG_ERR_INVALID:
            raise ValueError(msg)
        if rc == HBG_ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        if rc == HBG_ERR_NOMEM:
            raise MemoryError(msg)
        raise NativeLibraryError(f"libhbmpc_b200 error {rc}: {msg}")

    # -- plumbing ---------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.lib.hbg_ctx_set_stream(self.handle, cuda_stream))

    def synchronize(self):
        self._check(self.lib.hbg_ctx_synchronize(self.handle))

    def set_host_async(self, on):
        self._check(self.lib.hbg_ctx_set_host_async(self.handle, 1 if on else 0))

    def launch_count(self):
        return int(self.lib.hbg_ctx_launch_count(self.handle))

    def last_kernel(self):
        return self.lib.hbg_ctx_last_kernel(self.handle).decode()

    def set_fft_path(self, path):
        self._check(self.lib.hbg_ctx_set_fft_path(self.handle, {"auto": 0, "matrix": 1, "ntt": 2, "ntt-smem": 3, "ntt-split": 4, "ntt-bal": 5, "tc-split": 6}[path]))

    def set_matvec_path(self, path):
        self._check(self.lib.hbg_ctx_set_matvec_path(
            self.handle, {"auto": 0, "global": 1, "smem": 2, "small": 3, "small-r29": 4, "tc": 5,
                          "no-tc": 6}[path]))

    # -- batch operations (limb arrays or device pointers) ------------------
    def vandermonde_batch_evaluate(self, xs, polys, batch, d, out, mem=MEM_HOST):
        n = len(xs)
        self._check(self.lib.hbg_vandermonde_batch_evaluate(
            self.handle, _ptr(xs), n, _ptr(polys), batch, d, _ptr(out), mem))

    def vandermonde_batch_interpolate(self, xs, ys, batch, out, mem=MEM_HOST):
        k = len(xs)
        self._check(self.lib.hbg_vandermonde_batch_interpolate(
            self.handle, _ptr(xs), k, _ptr(ys), batch, _ptr(out), mem))

    def fft_batch_evaluate(self, omega, n, polys, batch, d, k_out, out, mem=MEM_HOST):
        self._check(self.lib.hbg_fft_batch_evaluate(
            self.handle, _ptr(omega), n, _ptr(polys), batch, d, k_out, _ptr(out), mem))

    def fft_batch_interpolate(self, omega, n, zs, ys, batch, out, mem=MEM_HOST):
        zs = np.ascontiguousarray(zs, dtype=np.int32)
        self._check(self.lib.hbg_fft_batch_interpolate(
            self.handle, _ptr(omega), n, _ptr(zs), len(zs), _ptr(ys), batch, _ptr(out), mem))


    def fft_batch_interpolate_allgather(self, omega, n, zs, ys, batch, peer_ptrs, multicast_ptr, rank):
        """device pointers only; peer_ptrs: one pointer per rank (ints)"""
        zs = np.ascontiguousarray(zs, dtype=np.int32)
        arr = peer_ptrs if isinstance(peer_ptrs, ctypes.Array) else self.peer_array(peer_ptrs)
        self._check(self.lib.hbg_fft_batch_interpolate_allgather(
            self.handle, _ptr(omega), n, _ptr(zs), len(zs), _ptr(ys), batch, arr,
            int(multicast_ptr) if multicast_ptr else None, len(peer_ptrs), rank))

    @staticmethod
    def peer_array(peer_ptrs):
        """marshal the per-rank buffer pointers once; the result can be passed wherever
        `peer_ptrs` is expected (hot loops call with the same buffers every step)"""
        return (ctypes.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])

    def allgather_block(self, block_ptr, nbytes, peer_ptrs, multicast_ptr, offset_bytes, max_ctas=0):
        arr = peer_ptrs if isinstance(peer_ptrs, ctypes.Array) else self.peer_array(peer_ptrs)
        self._check(self.lib.hbg_allgather_block(
            self.handle, int(block_ptr), nbytes, arr, int(multicast_ptr) if multicast_ptr else None,
            offset_bytes, len(peer_ptrs), max_ctas))

    # -- device-resident IncrementalDecoder -----------------------------------
    def set_cache_limit(self, nbytes):
        self._check(self.lib.hbg_ctx_set_cache_limit(self.handle, int(nbytes)))

    def set_tc_store(self, mode):
        """"direct" (default: 32 bytes per thread) or "staged" (through shared memory, full 128-byte lines)"""
        self._check(self.lib.hbg_ctx_set_tc_store(self.handle, {"direct": 0, "staged": 1}[mode]))

    def wait_pending(self, keep):
        """host_async: wait until at most `keep` of the latest host-buffer calls are in flight"""
        self._check(self.lib.hbg_ctx_wait_pending(self.handle, int(keep)))

    def set_sm_limit(self, ctas):
        """At most `ctas` CTAs (= SMs) per tensor-core launch of this context; 0 = all."""
        self._check(self.lib.hbg_ctx_set_sm_limit(self.handle, int(ctas)))

    def columns_to_rows(self, colbuf_ptr, batch, idx, rows_ptr):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._check(self.lib.hbg_columns_to_rows(self.handle, int(colbuf_ptr), batch, _ptr(idx), len(idx),
                                                 int(rows_ptr)))

    def interpolate_reencode(self, xs_k, xs_all, ys, batch, out, mem=MEM_HOST):
        self._check(self.lib.hbg_interpolate_reencode(
            self.handle, _ptr(xs_k), len(xs_k), _ptr(xs_all), len(xs_all), _ptr(ys), batch, _ptr(out), mem))

    def compare_columns(self, rows_ptr, row_width, col_offset, colbuf_ptr, batch, idx, flags_dev_ptr,
                        flags_host=None):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._check(self.lib.hbg_compare_columns(
            self.handle, int(rows_ptr), row_width, col_offset, int(colbuf_ptr), batch, _ptr(idx), len(idx),
            int(flags_dev_ptr), _ptr(flags_host)))

    def set_wb_path(self, path):
        self._check(self.lib.hbg_ctx_set_wb_path(self.handle, {"auto": 0, "exact": 1}[path]))

    def set_interp_path(self, path):
        self._check(self.lib.hbg_ctx_set_interp_path(self.handle, {"auto": 0, "matrix": 1, "fnt": 2}[path]))

    def allgather_block_signal(self, block_ptr, nbytes, peer_arr, multicast_ptr, offset_bytes, rank,
                               max_ctas, flags_arr, n_slots, slot, parts, first_part):
        self._check(self.lib.hbg_allgather_block_signal(
            self.handle, int(block_ptr), nbytes, peer_arr, int(multicast_ptr) if multicast_ptr else None,
            offset_bytes, len(peer_arr), rank, max_ctas, flags_arr, n_slots, slot, parts,
            1 if first_part else 0))

    def allgather_block_ce(self, block_ptr, nbytes, peer_arr, offset_bytes, rank, flags_arr, n_slots, slot,
                           parts, first_part):
        self._check(self.lib.hbg_allgather_block_ce(
            self.handle, int(block_ptr), nbytes, peer_arr, offset_bytes, len(peer_arr), rank, flags_arr,
            n_slots, slot, parts, 1 if first_part else 0))

    def allgather_block_bulk(self, block_ptr, nbytes, peer_arr, offset_bytes, rank, max_ctas, flags_arr, n_slots,
                             slot, parts, first_part):
        self._check(self.lib.hbg_allgather_block_bulk(
            self.handle, int(block_ptr), nbytes, peer_arr, offset_bytes, len(peer_arr), rank, max_ctas,
            flags_arr, n_slots, slot, parts, 1 if first_part else 0))

    def gather_wait(self, flags_arr, rank, n_slots, slot, parts):
        self._check(self.lib.hbg_gather_wait(self.handle, flags_arr, len(flags_arr), rank, n_slots, slot, parts))

    def gather_fence(self, flags_arr, rank, n_slots, slot, parts, phase):
        """phase 0: wait until every rank has released the slot; phase 1: signal that this rank's
        part has landed everywhere (around a fill the caller does itself: the fused gather)"""
        self._check(self.lib.hbg_gather_fence(self.handle, flags_arr, len(flags_arr), rank, n_slots, slot, parts,
                                              phase))

    def gather_release(self, flags_arr, rank, n_slots, slot):
        self._check(self.lib.hbg_gather_release(self.handle, flags_arr, len(flags_arr), rank, n_slots, slot))

    def gao_decode_batch(self, xs, k, ys, batch, coeffs, locator, loc_stride, loc_len, status,
                         mem=MEM_HOST):
        self._check(self.lib.hbg_gao_decode_batch(
            self.handle, _ptr(xs), len(xs), k, _ptr(ys), batch, _ptr(coeffs), _ptr(locator),
            loc_stride, _ptr(loc_len), _ptr(status), mem))

    def wb_decode_batch(self, xs, k, e_max, ys, batch, coeffs, out_len, status, mem=MEM_HOST):
        self._check(self.lib.hbg_wb_decode_batch(
            self.handle, _ptr(xs), len(xs), k, e_max, _ptr(ys), batch, _ptr(coeffs),
            _ptr(out_len), _ptr(status), mem))


_contexts = {}
_ctx_lock = threading.Lock()


def get_context(modulus, device=None):
    """Process-wide context cache (the reference re-inits NTL's modulus on
    every call, pyx:107,220,...; a context is that state made explicit)."""
    if device is None:
        device = int(os.environ.get("HBMPC_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    key = (int(modulus), device)
    with _ctx_lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = Context(modulus, device)
            _contexts[key] = ctx
        return ctx
