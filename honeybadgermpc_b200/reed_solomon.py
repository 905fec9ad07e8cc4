"""Reed-Solomon codec objects of the reconstruction path on the B200 kernels.

Same public surface as the reference's ``honeybadgermpc/reed_solomon.py``
(class names, constructor arguments, ``encode/decode/robust_decode``,
``IncrementalDecoder.add/done/get_results``, the selectors' thresholds and the
three factories with ``Algorithm``), so ``batch_reconstruct`` and callers such
as ``offline_randousha`` / ``refine_randoms`` run on it unchanged.

What is different underneath:
  * every codec also has ``*_limbs`` methods working on ``uint64[batch, w, 4]``
    arrays, so a pipeline (decode -> re-encode -> compare) never goes through
    Python ints;
  * ``IncrementalDecoder`` keeps the received columns in DEVICE memory (one
    ``colbuf[n][batch]`` buffer): the optimistic decode and the re-encode are one
    fused kernel (``hbg_interpolate_reencode``), later columns are validated by
    a compare kernel, and only per-column mismatch flags and the final rows come
    back (``device=False``: the same logic on host limb arrays); on the Byzantine path it
    decodes ALL remaining rows in one batched kernel call per eviction round
    (``robust_decode_batch``) instead of one Python call per row
    (reed_solomon.py:334-365) -- same results, because rows are still accepted
    in order and re-decoded after every eviction.
"""

import logging
from abc import ABC, abstractmethod

import numpy as np
import psutil

from . import ntl, robust
from .ntl import pack_rows, pack_vec, unpack_rows


class HoneyBadgerMPCError(Exception):
    """honeybadgermpc/exceptions.py"""


class DecodeValidationError(HoneyBadgerMPCError):
    pass


def _canonical(col, p):
    """every row of uint64[rows, 4] is < p"""
    pl = [(p >> (64 * i)) & (2 ** 64 - 1) for i in range(4)]
    if pl[3] and bool((col[:, 3] < np.uint64(pl[3])).all()):
        return True  # decided by the top limb alone (all but ~2^-64 of the canonical residues of a 255-bit p)
    lt = np.zeros(col.shape[0], dtype=bool)
    eq = np.ones(col.shape[0], dtype=bool)
    for i in (3, 2, 1, 0):
        lt |= eq & (col[:, i] < np.uint64(pl[i]))
        eq &= col[:, i] == np.uint64(pl[i])
    return bool(lt.all())


def _is_batch(data):
    return type(data[0]) in (list, tuple)


class Encoder(ABC):
    """reed_solomon.py:21-45"""

    def encode(self, data):
        return self.encode_batch(data) if _is_batch(data) else self.encode_one(data)

    def encode_one(self, data):
        return self.encode_batch([data])[0]

    def encode_batch(self, data):
        width = max(len(row) for row in data)
        return unpack_rows(self.encode_batch_limbs(pack_rows(data, width, self.modulus)))

    @abstractmethod
    def encode_batch_limbs(self, polys):
        """uint64[batch, d, 4] coefficients -> uint64[batch, n, 4] evaluations"""


class Decoder(ABC):
    """reed_solomon.py:48-74"""

    def decode(self, z, encoded):
        return self.decode_batch(z, encoded) if _is_batch(encoded) else self.decode_one(z, encoded)

    def decode_one(self, z, encoded):
        return self.decode_batch(z, [encoded])[0]

    def decode_batch(self, z, encoded):
        return unpack_rows(self.decode_batch_limbs(z, pack_rows(encoded, len(z), self.modulus)))

    @abstractmethod
    def decode_batch_limbs(self, z, ys):
        """z: k party indices; uint64[batch, k, 4] -> uint64[batch, k, 4] coefficients"""


class RowFailure:
    """Outcome of one row of ``robust_decode_batch`` for which the reference's
    one-row ``robust_decode`` raises (Welch-Berlekamp "No solution" / zero
    divisor, which reed_solomon.py:205-212 does not swallow).  The batched call
    must not raise for the whole batch: the reference would only ever reach this
    row after accepting every row before it (reed_solomon.py:334-365), so the
    exception travels with the row and is raised by whoever consumes it."""

    __slots__ = ("exc",)

    def __init__(self, exc):
        self.exc = exc


class RobustDecoder(ABC):
    """reed_solomon.py:77-85"""

    def robust_decode(self, z, encoded):
        out = self.robust_decode_batch(z, [encoded])[0]
        if isinstance(out, RowFailure):
            raise out.exc
        return out

    @abstractmethod
    def robust_decode_batch(self, z, rows):
        """rows: received words on the parties ``z`` -> one entry per row:
        ``(coefficients, sorted error party indices)``, ``(None, None)``, or a
        ``RowFailure`` carrying the exception the one-row call raises"""


def _points(point, idx):
    return [point(i).value for i in idx]


class VandermondeEncoder(Encoder):
    """reed_solomon.py:88-99 -> vandermonde_batch_evaluate"""

    def __init__(self, point):
        self.n = point.n
        self.modulus = point.field.modulus
        self.x = _points(point, range(self.n))
        self._xl = pack_vec(self.x, self.modulus)

    def encode_batch_limbs(self, polys):
        return ntl.vandermonde_batch_evaluate_limbs(self._xl, polys, self.modulus)


class FFTEncoder(Encoder):
    """reed_solomon.py:102-118 -> fft / fft_batch_evaluate on the omega powers"""

    def __init__(self, point):
        assert point.use_omega_powers is True, "FFTEncoder only usable with roots of unity evaluation points"
        self.n = point.n
        self.order = point.order
        self.omega = point.omega.value
        self.modulus = point.field.modulus
        self._wl = pack_vec([self.omega], self.modulus)[0]

    def encode_one(self, data):
        return ntl.fft(data, self.omega, self.modulus, self.order)[: self.n]

    def encode_batch(self, data):
        return ntl.fft_batch_evaluate(data, self.omega, self.modulus, self.order, self.n)

    def encode_batch_limbs(self, polys):
        return ntl.fft_batch_evaluate_limbs(polys, self._wl, self.modulus, self.order, self.n)


class VandermondeDecoder(Decoder):
    """reed_solomon.py:121-133 -> vandermonde_batch_interpolate"""

    def __init__(self, point):
        self.n = point.n
        self.modulus = point.field.modulus
        self.point = point

    def decode_batch(self, z, encoded):
        return ntl.vandermonde_batch_interpolate(_points(self.point, z), encoded, self.modulus)

    def decode_batch_limbs(self, z, ys):
        xl = pack_vec(_points(self.point, z), self.modulus)
        return ntl.vandermonde_batch_interpolate_limbs(xl, ys, self.modulus)


class FFTDecoder(Decoder):
    """reed_solomon.py:136-148 -> fft_interpolate / fft_batch_interpolate"""

    def __init__(self, point):
        assert point.use_omega_powers is True, "FFTDecoder only usable with roots of unity evaluation points"
        self.n = point.n
        self.order = point.order
        self.omega = point.omega.value
        self.modulus = point.field.modulus
        self._wl = pack_vec([self.omega], self.modulus)[0]

    def decode_batch(self, z, encoded):
        return ntl.fft_batch_interpolate(z, encoded, self.omega, self.modulus, self.order)

    def decode_batch_limbs(self, z, ys):
        try:
            return ntl.fft_batch_interpolate_limbs(list(z), ys, self._wl, self.modulus, self.order)
        except ntl._native.SingularError as e:
            raise ZeroDivisionError("repeated z") from e


class GaoRobustDecoder(RobustDecoder):
    """reed_solomon.py:151-186: Gao decode, then the error parties are the roots of
    the locator among all n evaluation points."""

    def __init__(self, d, point):
        self.d = d
        self.point = point
        self.modulus = point.field.modulus
        self.use_omega_powers = point.use_omega_powers
        self._all = pack_vec(_points(point, range(point.n)), self.modulus)

    def robust_decode(self, z, encoded):
        if any(v is None for v in encoded):  # erasures inside one word: pyx:399-403
            keep = [i for i, v in enumerate(encoded) if v is not None]
            return self.robust_decode_batch([z[i] for i in keep], [[encoded[i] for i in keep]])[0]
        return self.robust_decode_batch(z, [encoded])[0]

    def robust_decode_batch(self, z, rows):
        p = self.modulus
        xl = pack_vec(_points(self.point, z), p)
        yl = pack_rows(rows, len(z), p)
        coeffs, locator, loc_len, status = robust.gao_decode_batch_limbs(xl, yl, self.d + 1, p)
        ints = unpack_rows(coeffs)
        # one launch evaluates every locator on all n points (rows with a constant
        # locator have no roots; reed_solomon.py:174-184)
        ev = ntl.vandermonde_batch_evaluate_limbs(self._all, locator, p)
        zero = ~ev.any(axis=2)
        out = []
        for i in range(len(rows)):
            if status[i] != 0:
                out.append((None, None))
            elif loc_len[i] > 1:
                out.append((ints[i], [int(j) for j in np.nonzero(zero[i])[0]]))
            else:
                out.append((ints[i], []))
        return out


class WelchBerlekampRobustDecoder(RobustDecoder):
    """reed_solomon.py:189-225 over reed_solomon_wb.py:79-151; the error parties
    are the received positions that disagree with the decoded polynomial."""

    def __init__(self, d, point):
        self.n = point.n
        self.d = d
        self.modulus = point.field.modulus
        self.point = point
        self._all = pack_vec(_points(point, range(point.n)), self.modulus)

    def robust_decode_batch(self, z, rows):
        p, n, k = self.modulus, self.n, self.d + 1
        # positions in party order, as the reference's enc_extended list (:201-204)
        order = sorted(range(len(z)), key=lambda i: z[i])
        zs = [z[i] for i in order]
        rows_sorted = [[row[i] % p for i in order] for row in rows]
        decoded = robust.wb_decode_rows(_points(self.point, zs), rows_sorted, n, n - len(zs), k, p)
        good = [i for i, c in enumerate(decoded) if isinstance(c, list)]
        out = [RowFailure(c) if isinstance(c, BaseException) else (None, None) for c in decoded]
        if good:
            width = max(1, max(len(decoded[i]) for i in good))
            ev = unpack_rows(ntl.vandermonde_batch_evaluate_limbs(
                self._all, pack_rows([decoded[i] for i in good], width, p), p))
            for slot, i in enumerate(good):
                errs = [zi for zi, v in zip(zs, rows_sorted[i]) if ev[slot][zi] != v]
                out[i] = (decoded[i], errs)
        return out


_GUESS = object()  # IncrementalDecoder._result: "the accepted optimistic guess, still as limbs"

DEVICE_MIN_BATCH = 256  # below this the host path's few numpy ops beat kernel launches


class _DeviceColumns:
    """Device side of one IncrementalDecoder: the column buffer ``colbuf[n][batch]``, the
    fused decode+re-encode output ``guess[batch][k+n]`` and the per-column mismatch flags.
    torch is the memory / stream plumbing; the work is the three C-ABI calls
    ``hbg_columns_to_rows``, ``hbg_interpolate_reencode``, ``hbg_compare_columns``."""

    _stream = None
    totals = {"decoders": 0, "c_abi_calls": 0, "h2d_bytes": 0, "d2h_bytes": 0}  # process-wide counters

    def __init__(self, ctx, n, batch, k):
        import torch

        _DeviceColumns.totals["decoders"] += 1

        self.torch, self.ctx, self.n, self.batch, self.k = torch, ctx, n, batch, k
        self.dev = torch.device("cuda", ctx.device)
        if _DeviceColumns._stream is None or _DeviceColumns._stream.device != self.dev:
            _DeviceColumns._stream = torch.cuda.Stream(device=self.dev)
        self.stream = _DeviceColumns._stream
        ctx.set_stream(self.stream.cuda_stream)
        with torch.cuda.stream(self.stream):
            self.colbuf = torch.empty((n, batch, 4), dtype=torch.int64, device=self.dev)
            self.guess = None
            self.flags = torch.zeros(n, dtype=torch.int32, device=self.dev)
        self.compared = []   # party indices whose compare kernel has been enqueued

    def put(self, idx, col):
        """H2D of one received column (uint64[batch, 4])"""
        t = self.torch
        src = np.require(col, requirements=["C", "W"]).view(np.int64)  # frombuffer views are read-only
        with t.cuda.stream(self.stream):
            self.colbuf[idx].copy_(t.from_numpy(src), non_blocking=True)
        _DeviceColumns.totals["h2d_bytes"] += col.nbytes

    def decode_reencode(self, z, xs_k, xs_all):
        t = self.torch
        with t.cuda.stream(self.stream):
            rows = t.empty((self.batch, self.k, 4), dtype=t.int64, device=self.dev)
            self.guess = t.empty((self.batch, self.k + self.n, 4), dtype=t.int64, device=self.dev)
        self.ctx.columns_to_rows(self.colbuf.data_ptr(), self.batch, z, rows.data_ptr())
        self.ctx.interpolate_reencode(xs_k, xs_all, rows.data_ptr(), self.batch, self.guess.data_ptr(),
                                      mem=ntl._native.MEM_DEVICE)
        _DeviceColumns.totals["c_abi_calls"] += 2

    def compare(self, idx):
        """enqueue the validation of column idx against the re-encoded guess (no sync)"""
        self.ctx.compare_columns(self.guess.data_ptr(), self.k + self.n, self.k, self.colbuf.data_ptr(),
                                 self.batch, [idx], self.flags[idx:idx + 1].data_ptr())
        self.compared.append(idx)
        _DeviceColumns.totals["c_abi_calls"] += 1

    def any_mismatch(self):
        """one D2H of all flags written so far"""
        if not self.compared:
            return False
        with self.torch.cuda.stream(self.stream):
            host = self.flags.cpu()
        _DeviceColumns.totals["d2h_bytes"] += host.numel() * 4
        return bool(host[self.compared].any())

    def coefficients(self):
        """uint64[batch, k, 4] on the host"""
        with self.torch.cuda.stream(self.stream):
            out = self.guess[:, : self.k, :].contiguous().cpu().numpy().view(np.uint64)
        _DeviceColumns.totals["d2h_bytes"] += out.nbytes
        return out

    def encoded_host(self):
        with self.torch.cuda.stream(self.stream):
            return self.guess[:, self.k:, :].contiguous().cpu().numpy().view(np.uint64)


class IncrementalDecoder:
    """reed_solomon.py:232-403.  Feed it one party's column at a time with
    ``add(idx, data)``; it decodes optimistically from the first degree+1
    columns, checks later columns against the re-encoded guess, falls back to
    robust decoding (evicting the parties found in error) when a column
    disagrees, and is ``done()`` once degree+1+max_errors-|confirmed errors|
    parties agree on every row."""

    def __init__(self, encoder, decoder, robust_decoder, degree, batch_size, max_errors,
                 confirmed_errors=None, validator=None, device="auto"):
        self.encoder, self.decoder, self.robust_decoder = encoder, decoder, robust_decoder
        self.degree, self.batch_size, self.max_errors = degree, batch_size, max_errors
        self.validator = validator
        self.modulus = encoder.modulus if hasattr(encoder, "modulus") else decoder.modulus
        # device-resident columns: needs the evaluation points (every codec of this module
        # carries them), a real CUDA context, and a batch worth a kernel launch
        self._dev = None
        self._point = getattr(decoder, "point", None) or getattr(encoder, "point", None) or \
            getattr(robust_decoder, "point", None)
        if device is True or (device == "auto" and batch_size >= DEVICE_MIN_BATCH):
            ctx = ntl._ctx(self.modulus)
            if isinstance(ctx, ntl._native.Context) and self._point is not None:
                n = self._point.n
                self._dev = _DeviceColumns(ctx, n, batch_size, degree + 1)
                self._xs_all = pack_vec(_points(self._point, range(n)), self.modulus)
            elif device is True:
                raise ntl._native.NativeLibraryError("device-resident IncrementalDecoder needs the CUDA library")
        self._confirmed_errors = confirmed_errors if confirmed_errors is not None else set()
        self._z = []            # party indices in arrival order
        self._cols = []         # their columns, uint64[batch, 4] each
        self._points_seen = set()
        self._optimistic = True
        self._guess = None      # uint64[batch, degree+1, 4]
        self._guess_encoded = None  # uint64[batch, n, 4]
        self._done_rows = []    # rows finished by the robust path (lists of ints)
        self._result_is_guess = False
        self._result = None

    # -- helpers ------------------------------------------------------------
    def _need(self):
        return self.degree + 1 + self.max_errors - len(self._confirmed_errors)

    def _check_input(self, data):
        if len(data) != self.batch_size:
            raise DecodeValidationError("Incorrect length of data")
        if self.validator is not None:
            for v in data:
                self.validator(v)

    def _to_column(self, data):
        """list of ints (the reference's wire format) or a packed column
        (``bytes`` of batch*32 little-endian bytes / ``uint64[batch, 4]``, the
        limb wire format of ``batch_reconstruct(..., wire="limbs")``)"""
        if isinstance(data, (bytes, bytearray, memoryview)):
            if len(data) != self.batch_size * 32:  # a faulty sender must not surface as ValueError
                raise DecodeValidationError("Incorrect length of data")
            col = np.frombuffer(data, dtype=np.uint64).reshape(-1, 4)
        elif isinstance(data, np.ndarray):
            if data.size != self.batch_size * 4:
                raise DecodeValidationError("Incorrect length of data")
            col = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
        else:
            self._check_input(data)
            return pack_rows([data], self.batch_size, self.modulus)[0] if self.batch_size else \
                np.zeros((0, 4), np.uint64)
        if col.shape[0] != self.batch_size:
            raise DecodeValidationError("Incorrect length of data")
        if not _canonical(col, self.modulus):  # a faulty sender: reduce like to_ZZ_p would
            col = pack_rows(unpack_rows(col[None]), self.batch_size, self.modulus)[0]
        if self.validator is not None:
            for v in unpack_rows(col[None])[0]:
                self.validator(v)
        return col

    def _try_guess(self, idx, col):
        """optimistic path; True while the guess is still standing"""
        if self._dev is not None:
            return self._try_guess_device(idx)
        if len(self._z) == self.degree + 1:
            ys = np.stack(self._cols, axis=1)
            self._guess = self.decoder.decode_batch_limbs(self._z, ys)
            self._guess_encoded = self.encoder.encode_batch_limbs(self._guess)
        elif not np.array_equal(col, self._guess_encoded[:, idx, :]):
            logging.critical("Optimistic decoding failed")
            self._guess = self._guess_encoded = None
            self._optimistic = False
            return False
        if len(self._z) >= self._need():
            self._result = _GUESS  # the int rows are made on demand (get_results)
            self._result_is_guess = True
        return True

    def _try_guess_device(self, idx):
        """The same decision on the device (reed_solomon.py:305-331): decode + re-encode is
        one kernel when the (degree+1)-th column arrives, every later column costs one compare
        launch and NO synchronisation; the flags are read once, when enough columns are in for
        the guess to be accepted.  A mismatch found then sends the decoder down the robust
        path at the same ``add`` at which the reference's robust path first has enough
        points (both need ``_min_points_required`` columns), so the observable behaviour
        -- done(), results, confirmed errors after every add -- is the reference's."""
        dev = self._dev
        if len(self._z) == self.degree + 1:
            xs_k = pack_vec(_points(self._point, self._z), self.modulus)
            dev.decode_reencode(self._z, xs_k, self._xs_all)
        else:
            dev.compare(idx)
        if len(self._z) < self._need():
            return True
        if dev.any_mismatch():
            logging.critical("Optimistic decoding failed")
            self._optimistic = False
            return False
        self._guess = dev.coefficients()
        self._result = _GUESS
        self._result_is_guess = True
        return True

    def _robust_rounds(self):
        """batched form of _robust_update (reed_solomon.py:334-365)"""
        while len(self._done_rows) < self.batch_size:
            start = len(self._done_rows)
            rows = unpack_rows(np.stack([c[start:] for c in self._cols], axis=1))
            decoded = self.robust_decoder.robust_decode_batch(list(self._z), rows)
            evicted = False
            for outcome in decoded:
                if isinstance(outcome, RowFailure):
                    raise outcome.exc  # only now: the rows before it were all accepted
                coeffs, errors = outcome
                if coeffs is None or len(self._z) - len(errors) < self._need():
                    return  # wait for more columns
                self._done_rows.append(coeffs)
                if errors:
                    self._confirmed_errors |= set(errors)
                    for e in errors:
                        at = self._z.index(e)
                        del self._z[at]
                        del self._cols[at]
                        self._points_seen.discard(e)
                    evicted = True
                    break  # the remaining rows are re-decoded without the evicted parties
            if not evicted:
                break
        if len(self._done_rows) == self.batch_size:
            self._result = self._done_rows

    # -- public API -----------------------------------------------------------
    def add(self, idx, data):
        if self.done() or idx in self._points_seen or idx in self._confirmed_errors:
            return
        col = self._to_column(data)
        self._points_seen.add(idx)
        self._z.append(idx)
        self._cols.append(col)  # host copy: the (rare) robust path decodes from it
        if self._dev is not None and self._optimistic:
            self._dev.put(idx, col)
        if len(self._z) <= self.degree:
            return
        if self._optimistic and self._try_guess(idx, col):
            return
        if len(self._z) >= self._need():
            self._robust_rounds()

    def done(self):
        return self._result is not None

    def get_results(self):
        if self._result is None:
            return None, None
        if self._result is _GUESS:
            self._result = unpack_rows(self._guess)
        return self._result, self._confirmed_errors

    def get_results_limbs(self):
        """the decoded rows as ``uint64[batch, degree+1, 4]`` (no Python ints)"""
        if self._result is None:
            return None, None
        if self._guess is not None and self._result_is_guess:
            return self._guess, self._confirmed_errors  # no Python ints were ever made
        return pack_rows(self._result, self.degree + 1, self.modulus), self._confirmed_errors


class EncoderSelector:
    """reed_solomon.py:406-434.  The thresholds are the reference's CPU heuristics;
    on the GPU both classes end in the same kernels, but the class a caller gets
    is part of the observable behaviour (tests/test_reed_solomon.py:186-277)."""

    LOW_VAN_THRESHOLD = 8
    HIGH_VAN_THRESHOLD = 128

    @staticmethod
    def set_optimal_thread_count(k):
        ntl.SetNumThreads(min(k, psutil.cpu_count(logical=False)))

    @staticmethod
    def select(point, k):
        assert point.use_omega_powers is True
        n = point.n
        if n < EncoderSelector.LOW_VAN_THRESHOLD:
            return VandermondeEncoder(point)
        if n >= EncoderSelector.HIGH_VAN_THRESHOLD:
            return FFTEncoder(point)
        npow2 = n if n & (n - 1) == 0 else 2 ** n.bit_length()
        far_from_pow2 = npow2 - n > npow2 // 4
        return VandermondeEncoder(point) if far_from_pow2 else FFTEncoder(point)


class DecoderSelector:
    """reed_solomon.py:437-459"""

    LOW_VAN_THRESHOLD = 8
    BATCH_SIZE_THRESH_SLOPE = 0.5

    @staticmethod
    def set_optimal_thread_count(k):
        ntl.SetNumThreads(min(k, psutil.cpu_count(logical=False)))

    @staticmethod
    def select(point, k):
        assert point.use_omega_powers is True
        n = point.n
        if n < DecoderSelector.LOW_VAN_THRESHOLD:
            return VandermondeDecoder(point)
        if k > DecoderSelector.BATCH_SIZE_THRESH_SLOPE * n * ntl.AvailableNTLThreads():
            return VandermondeDecoder(point)
        return FFTDecoder(point)


class OptimalEncoder(Encoder):
    """reed_solomon.py:462-475"""

    def __init__(self, point):
        assert point.use_omega_powers is True
        self.point = point
        self.modulus = point.field.modulus

    def _pick(self, count):
        EncoderSelector.set_optimal_thread_count(count)
        return EncoderSelector.select(self.point, count)

    def encode_one(self, data):
        return self._pick(1).encode_one(data)

    def encode_batch(self, data):
        return self._pick(len(data)).encode_batch(data)

    def encode_batch_limbs(self, polys):
        return self._pick(polys.shape[0]).encode_batch_limbs(polys)


class OptimalDecoder(Decoder):
    """reed_solomon.py:478-491"""

    def __init__(self, point):
        assert point.use_omega_powers is True
        self.point = point
        self.modulus = point.field.modulus

    def _pick(self, count):
        DecoderSelector.set_optimal_thread_count(count)
        return DecoderSelector.select(self.point, count)

    def decode_one(self, z, data):
        return self._pick(1).decode_one(z, data)

    def decode_batch(self, z, data):
        return self._pick(len(data)).decode_batch(z, data)

    def decode_batch_limbs(self, z, ys):
        return self._pick(ys.shape[0]).decode_batch_limbs(z, ys)


class Algorithm:
    VANDERMONDE = "vandermonde"
    FFT = "fft"
    GAO = "gao"
    WELCH_BERLEKAMP = "welch-berlekamp"


def _unknown(kind, names):
    return ValueError(f"Incorrect algorithm for a {kind}. Supported algorithms are {names}; "
                      "pass algorithm=None with omega-power points for automatic selection")


class EncoderFactory:
    @staticmethod
    def get(point, algorithm=None):
        if algorithm is None:
            return OptimalEncoder(point) if point.use_omega_powers else VandermondeEncoder(point)
        table = {Algorithm.VANDERMONDE: VandermondeEncoder, Algorithm.FFT: FFTEncoder}
        if algorithm not in table:
            raise _unknown("encoder", list(table))
        return table[algorithm](point)


class DecoderFactory:
    @staticmethod
    def get(point, algorithm=None):
        if algorithm is None:
            return OptimalDecoder(point) if point.use_omega_powers else VandermondeDecoder(point)
        table = {Algorithm.VANDERMONDE: VandermondeDecoder, Algorithm.FFT: FFTDecoder}
        if algorithm not in table:
            raise _unknown("decoder", list(table))
        return table[algorithm](point)


class RobustDecoderFactory:
    @staticmethod
    def get(t, point, algorithm=Algorithm.GAO):
        table = {Algorithm.GAO: GaoRobustDecoder, Algorithm.WELCH_BERLEKAMP: WelchBerlekampRobustDecoder}
        if algorithm not in table:
            raise _unknown("robust decoder", list(table))
        return table[algorithm](t, point)
