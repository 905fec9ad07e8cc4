"""n parties on n GPUs: ``batch_reconstruct`` with its two message rounds as collectives.

SURVEY.md section 8(f), row 4: a simulation mode for benchmarking the opening
protocol without the asyncio router -- every rank of a ``torch.distributed``
group IS one party (party index = rank, n = world size), its shares live in
device memory, R1 (``batch_reconstruction.py:158-167``: party i sends the
evaluation of its chunk polynomials at x_j to party j) is ONE all-to-all and R2
(``:193-196``: party j broadcasts the constant coefficient it decoded) is ONE
all-gather; the encodes / interpolations in between are the CUDA kernels of this
package on device pointers.

Algebra: party i holds shares s_{i,b} = f_b(x_i) of the
secrets S_b.  Chunk c, k = degree + 1: g_{i,c}(X) = sum_l s_{i,ck+l} X^l.  For fixed
j, c the map i -> g_{i,c}(x_j) is a polynomial of degree <= degree in x_i whose
constant term is G_c(x_j), G_c(X) = sum_l S_{ck+l} X^l.  Party j interpolates it
from the first k parties, re-encodes and compares the remaining columns like the
optimistic path of ``IncrementalDecoder`` does (reed_solomon.py:305-331); if a column
disagrees (a faulty sender) it falls back to the robust decoder on ALL n columns
(``hbg_gao_decode_batch`` on device pointers; the error parties are the roots of each
row's locator, reed_solomon.py:174-184) and accepts when every row decodes with at most t
errors.  It then publishes G_c(x_j); everybody decodes j -> G_c(x_j) the same way and
reads the secrets off the coefficients.

The three local phases are separate functions so that a single process can play
all parties on one GPU (tests/test_gpu_protocol.py) and so that the CPU test of
the message pattern (gloo, tests/test_party_sim.py) can plug in another codec.
"""

import numpy as np
import torch
import torch.distributed as dist

from . import _native
from .field import GF
from .polynomial import EvalPoint


class CudaCodec:
    """Encode / interpolate on device tensors (int64[rows, width, 4] limb arrays on a
    CUDA device) through the C-ABI with ``HBG_MEM_DEVICE`` pointers.  Plain points
    x_i = i + 1 (what ``Mpc`` uses, mpc.py:139) or omega powers."""

    def __init__(self, modulus, n, device=None, use_omega_powers=False):
        if not torch.cuda.is_available():
            raise _native.NativeLibraryError("CudaCodec needs a CUDA device (there is no CPU fallback)")
        from .ntl import pack_vec

        self.modulus, self.n = int(modulus), n
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.point = EvalPoint(GF(self.modulus), n, use_omega_powers)
        self.ctx = _native.Context(self.modulus, device=self.device.index)
        self.use_omega_powers = use_omega_powers
        if use_omega_powers:
            self._omega = pack_vec([self.point.omega.value], self.modulus)[0]
        self._xs = pack_vec([self.point(i).value for i in range(n)], self.modulus)

    def _bind(self):
        # torch's default stream has handle 0, which hbg_ctx_set_stream reads as "the context's
        # own stream": name the legacy default stream explicitly (cudaStreamLegacy = 1), or the
        # kernels would race with the torch ops around them
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream or 1)

    def encode(self, coeffs):
        """[rows, d, 4] coefficients -> [rows, n, 4] evaluations at all n points"""
        rows, d = coeffs.shape[0], coeffs.shape[1]
        out = torch.empty((rows, self.n, 4), dtype=torch.int64, device=self.device)
        if rows == 0:
            return out
        self._bind()
        coeffs = coeffs.contiguous()
        if self.use_omega_powers:
            self.ctx.fft_batch_evaluate(self._omega, self.point.order, coeffs.data_ptr(), rows, d, self.n,
                                        out.data_ptr(), _native.MEM_DEVICE)
        else:
            self.ctx.vandermonde_batch_evaluate(self._xs, coeffs.data_ptr(), rows, d, out.data_ptr(),
                                                _native.MEM_DEVICE)
        return out

    def interpolate(self, z, ys):
        """z: k party indices, [rows, k, 4] values at those parties' points -> [rows, k, 4] coefficients"""
        rows, k = ys.shape[0], ys.shape[1]
        assert len(z) == k
        out = torch.empty((rows, k, 4), dtype=torch.int64, device=self.device)
        if rows == 0:
            return out
        self._bind()
        ys = ys.contiguous()
        if self.use_omega_powers:
            self.ctx.fft_batch_interpolate(self._omega, self.point.order, np.asarray(z, dtype=np.int32),
                                           ys.data_ptr(), rows, out.data_ptr(), _native.MEM_DEVICE)
        else:
            self.ctx.vandermonde_batch_interpolate(np.ascontiguousarray(self._xs[list(z)]), ys.data_ptr(),
                                                   rows, out.data_ptr(), _native.MEM_DEVICE)
        return out


    def robust_decode(self, rows, k):
        """Gao decode of every word ``rows[C, n, 4]`` over all n parties (max (n-k)/2 errors per
        word).  Returns ``(coeffs [C, k, 4], decoded bool[C], bad bool[C, n])``: ``bad[c, i]``
        marks party i as an error position of word c (a root of its error locator)."""
        c, n = rows.shape[0], self.n
        loc_stride = n - (n + k) // 2 + 1
        mk = dict(device=self.device)
        coeffs = torch.zeros((c, k, 4), dtype=torch.int64, **mk)
        locator = torch.zeros((c, loc_stride, 4), dtype=torch.int64, **mk)
        loc_len = torch.zeros(c, dtype=torch.int32, **mk)
        status = torch.ones(c, dtype=torch.int32, **mk)
        if c == 0:
            return coeffs, status == 0, torch.zeros((0, n), dtype=torch.bool, **mk)
        self._bind()
        rows = rows.contiguous()
        self.ctx.gao_decode_batch(self._xs, k, rows.data_ptr(), c, coeffs.data_ptr(), locator.data_ptr(),
                                  loc_stride, loc_len.data_ptr(), status.data_ptr(), _native.MEM_DEVICE)
        ev = torch.empty((c, n, 4), dtype=torch.int64, **mk)
        self.ctx.vandermonde_batch_evaluate(self._xs, locator.data_ptr(), c, loc_stride, ev.data_ptr(),
                                            _native.MEM_DEVICE)
        bad = ~(ev != 0).any(dim=2) & (loc_len > 1)[:, None]
        return coeffs, status == 0, bad


class PartyState:
    """What a party keeps between the phases of one opening."""

    def __init__(self, batch, k, t=0, ok=True):
        self.batch, self.k, self.t, self.ok = batch, k, t, ok
        self.errors = set()   # parties found lying (by the robust fallback)
        self.robust = 0       # how many of the two rounds needed the robust decoder


def _decode_round(codec, state, rows, check):
    """optimistic interpolate + re-encode + compare; robust fallback on a mismatch"""
    h = codec.interpolate(list(range(state.k)), rows[:, :state.k].contiguous())
    if not check or bool(torch.equal(codec.encode(h), rows)):
        return h
    state.robust += 1
    h, decoded, bad = codec.robust_decode(rows, state.k)
    errs = set(torch.nonzero(bad.any(dim=0)).flatten().tolist())
    state.errors |= errs
    state.ok = state.ok and bool(decoded.all()) and int(bad.sum(dim=1).max()) <= state.t
    return h


def phase1_encode(codec, shares, degree):
    """shares: [B, 4] limbs of this party's shares.  Returns (state, send) with
    send[j] = [C, 4], the R1 message for party j (C = ceil(B / k) chunk polynomials)."""
    k = degree + 1
    batch = shares.shape[0]
    chunks = -(-batch // k) if batch else 0
    padded = torch.zeros((chunks * k, 4), dtype=shares.dtype, device=shares.device)
    padded[:batch] = shares
    enc = codec.encode(padded.view(chunks, k, 4))          # enc[c][j] = g_{me,c}(x_j)
    return PartyState(batch, k, t=degree), enc.transpose(0, 1).contiguous()


def phase2_decode_r1(codec, state, recv, check=True):
    """recv[i] = [C, 4]: what party i sent to this party.  Returns the R2 message [C, 4]."""
    rows = recv.transpose(0, 1).contiguous()               # [C, n, 4], a word over the parties i
    h = _decode_round(codec, state, rows, check)
    return h[:, 0].contiguous()                            # constant coefficient = G_c(x_me)


def phase3_decode_r2(codec, state, allv, check=True):
    """allv[j] = [C, 4]: party j's R2 message.  Returns the opened secrets [B, 4]."""
    rows = allv.transpose(0, 1).contiguous()               # [C, n, 4], a word over the parties j
    g = _decode_round(codec, state, rows, check)
    return g.reshape(-1, 4)[:state.batch]


def exchange_columns(send, group=None):
    """R1: block j of ``send`` goes to rank j; returns the blocks received, by sender.
    One all-to-all on NCCL; point-to-point sends on backends without it (gloo)."""
    recv = torch.empty_like(send)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, group=group)
        return recv
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ops = []
    for peer in range(world):
        if peer == rank:
            recv[peer] = send[peer]
            continue
        ops.append(dist.P2POp(dist.isend, send[peer].contiguous(), peer, group))
        ops.append(dist.P2POp(dist.irecv, recv[peer], peer, group))
    for w in dist.batch_isend_irecv(ops) if ops else []:
        w.wait()
    return recv


def batch_reconstruct_collective(shares, t, codec, degree=None, group=None, check=True, byzantine=False,
                                 info=None):
    """Open ``shares`` ([B, 4] limb tensor, this rank's = this party's shares) among the
    ranks of ``group``.  Returns ``(secrets [B, 4], ok)`` on every rank; a faulty sender makes
    the affected round fall back to the robust decoder (``ok`` stays True as long as every
    word has at most t errors).  ``byzantine=True`` turns THIS rank into a faulty party: it
    sends noise in both rounds.  ``info`` (a dict) receives ``errors`` / ``robust_rounds``."""
    degree = t if degree is None else degree
    world = dist.get_world_size(group)
    assert codec.n == world, "one party per rank"
    state, send = phase1_encode(codec, shares, degree)
    state.t = t
    if byzantine:
        send = torch.randint(0, 2 ** 60, send.shape, dtype=send.dtype, device=send.device)
    recv = exchange_columns(send, group)
    r2 = phase2_decode_r1(codec, state, recv, check)
    if byzantine:
        r2 = torch.randint(0, 2 ** 60, r2.shape, dtype=r2.dtype, device=r2.device)
    allv = torch.empty((world * r2.shape[0], 4), dtype=r2.dtype, device=r2.device)  # blocks by rank
    dist.all_gather_into_tensor(allv, r2, group=group)
    out = phase3_decode_r2(codec, state, allv.view(world, r2.shape[0], 4), check)
    if info is not None:
        info["errors"], info["robust_rounds"] = sorted(state.errors), state.robust
    return out, state.ok


def simulate_in_process(codecs, shares_by_party, t, degree=None, check=True, byzantine=(), info=None):
    """All n parties played by one process (one GPU): the same three phases, the two
    message rounds as tensor shuffles.  Parties in ``byzantine`` send noise in both rounds.
    Returns ``[(secrets, ok)]`` per party; ``info`` (a list) receives each party's
    ``(errors, robust_rounds)``."""
    degree = t if degree is None else degree
    n = len(shares_by_party)
    states, sends = zip(*[phase1_encode(codecs[i], shares_by_party[i], degree) for i in range(n)])
    sends = list(sends)
    for st in states:
        st.t = t
    for i in byzantine:
        sends[i] = torch.randint(0, 2 ** 60, sends[i].shape, dtype=sends[i].dtype, device=sends[i].device)
    r2 = [phase2_decode_r1(codecs[j], states[j], torch.stack([sends[i][j] for i in range(n)]), check)
          for j in range(n)]
    for i in byzantine:
        r2[i] = torch.randint(0, 2 ** 60, r2[i].shape, dtype=r2[i].dtype, device=r2[i].device)
    allv = torch.stack(r2)
    out = [(phase3_decode_r2(codecs[j], states[j], allv, check), states[j].ok) for j in range(n)]
    if info is not None:
        info.extend((sorted(st.errors), st.robust) for st in states)
    return out
