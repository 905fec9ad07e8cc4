"""Batched public reconstruction of secret-shared values (two rounds, R1/R2).

Same coroutine signature and message format as the reference's
``honeybadgermpc/batch_reconstruction.py:88-227`` -- it is what
``Mpc.open_share_array`` (mpc.py:203) schedules for every ``ShareArray.open()``
-- with the compute (3 encodes, 2 decodes, column validation) running on the
CUDA kernels through ``honeybadgermpc_b200.reed_solomon``.

Algebra: party i holds shares s_{i,b}; in chunks of
k = degree+1 it forms g_{i,c}(X) = sum_l s_{i,ck+l} X^l and sends g_{i,c}(x_j)
to party j (R1).  For fixed (j, c), i -> g_{i,c}(x_j) is a polynomial of degree
<= degree in x_i whose constant term is G_c(x_j), G_c(X) = sum_l S_{ck+l} X^l;
party j decodes it and broadcasts the constant terms (R2); decoding
j -> G_c(x_j) yields the secrets S_{ck+l}.
"""

import asyncio
import logging
import random
import time

from .field import GF
from .polynomial import EvalPoint
from .reed_solomon import (
    Algorithm,
    DecoderFactory,
    EncoderFactory,
    IncrementalDecoder,
    RobustDecoderFactory,
)
from .utils import chunk_data, flatten_lists, subscribe_recv, transpose_lists


async def fetch_one(awaitables):
    """Yield ``(index, result)`` in completion order (batch_reconstruction.py:25-40)."""
    index = {aw: i for i, aw in enumerate(awaitables)}
    pending = set(awaitables)
    while pending:
        done, pending = await asyncio.wait(pending, return_when=asyncio.FIRST_COMPLETED)
        for d in done:
            yield index[d], await d


async def incremental_decode(receivers, encoder, decoder, robust_decoder, batch_size, t, degree, n):
    """batch_reconstruction.py:43-61"""
    inc = IncrementalDecoder(encoder, decoder, robust_decoder, degree=degree,
                             batch_size=batch_size, max_errors=t)
    async for idx, column in fetch_one(receivers):
        inc.add(idx, column)
        if inc.done():
            return inc.get_results()[0]
    return None


def recv_each_party(recv, n):
    """One queue per sender (batch_reconstruction.py:64-85)."""
    queues = [asyncio.Queue() for _ in range(n)]

    async def pump():
        while True:
            sender, payload = await recv()
            queues[sender].put_nowait(payload)

    return asyncio.create_task(pump()), [q.get for q in queues]


async def batch_reconstruct(secret_shares, p, t, n, myid, send, recv, config=None,
                            use_omega_powers=False, debug=False, degree=None, wire="ints"):
    """Open ``len(secret_shares)`` shared values; returns them as ``GFElement``
    (or ``None`` when a round cannot be decoded).

    ``wire="ints"`` (default) sends the reference's message format -- lists of
    Python ints -- and interoperates with reference parties.  ``wire="limbs"``
    sends every R1/R2 payload as ``bytes`` of 32-byte little-endian limbs (the
    encoder's output column as is): no int <-> limb
    marshalling on the steady-state path, 32 B per element on the wire instead
    of a pickled int list.  All parties of a run must use the same format."""
    bench = logging.LoggerAdapter(logging.getLogger("benchmark_logger"), {"node_id": myid})
    if degree is None:
        degree = t
    shares = [v.value for v in secret_shares]
    if config is not None and config.induce_faults:
        logging.debug("[FAULT][BatchReconstruction] Sending random shares.")
        shares = [random.randint(0, p - 1) for _ in shares]

    subscribe_task, subscribe = subscribe_recv(recv)
    del recv
    task_r1, getters_r1 = recv_each_party(subscribe("R1"), n)
    data_r1 = [asyncio.create_task(g()) for g in getters_r1]
    task_r2, getters_r2 = recv_each_party(subscribe("R2"), n)
    data_r2 = [asyncio.create_task(g()) for g in getters_r2]
    del subscribe
    background = [task_r1, task_r2, subscribe_task, *data_r1, *data_r2]

    def cancel_all():
        for task in background:
            task.cancel()

    fp = GF(p)
    point = EvalPoint(fp, n, use_omega_powers=use_omega_powers)
    algo = Algorithm.FFT if use_omega_powers else Algorithm.VANDERMONDE
    enc = EncoderFactory.get(point, algo)
    dec = DecoderFactory.get(point, algo)
    robust_dec = RobustDecoderFactory.get(
        t, point, algorithm=Algorithm.GAO if config is None else config.decoding_algorithm)

    # round 1: every party gets one evaluation of each chunk polynomial
    chunks = chunk_data(shares, degree + 1)
    num_chunks = len(chunks)
    t0 = time.time()
    if wire == "limbs":
        from .ntl import pack_rows

        encoded = enc.encode_batch_limbs(pack_rows(chunks, degree + 1, p))  # [chunks][n][4]
        for dest in range(n):
            send(dest, ("R1", encoded[:, dest, :].tobytes()))
    else:
        for dest, column in enumerate(transpose_lists(enc.encode(chunks))):
            send(dest, ("R1", column))
    bench.info(f"[BatchReconstruct] P1 Send: {time.time() - t0}")

    t0 = time.time()
    try:
        round1 = await incremental_decode(data_r1, enc, dec, robust_dec, num_chunks, t, degree, n)
    except asyncio.CancelledError:
        cancel_all()
        raise
    if round1 is None:
        logging.error("[BatchReconstruct] P1 reconstruction failed!")
        return None
    bench.info(f"[BatchReconstruct] P1 Reconstruct: {time.time() - t0}")

    # round 2: broadcast the constant terms (= the chunk polynomials at my point)
    t0 = time.time()
    if wire == "limbs":
        from .ntl import pack_rows

        message = pack_rows([[row[0] for row in round1]], num_chunks, p)[0].tobytes()
    else:
        message = [row[0] for row in round1]
    for dest in range(n):
        send(dest, ("R2", message))
    bench.info(f"[BatchReconstruct] P2 Send: {time.time() - t0}")

    t0 = time.time()
    try:
        round2 = await incremental_decode(data_r2, enc, dec, robust_dec, num_chunks, t, degree, n)
    except asyncio.CancelledError:
        cancel_all()
        raise
    if round2 is None:
        logging.error("[BatchReconstruct] P2 reconstruction failed!")
        return None
    bench.info(f"[BatchReconstruct] P2 Reconstruct: {time.time() - t0}")

    cancel_all()
    opened = flatten_lists(round2)
    assert len(opened) >= len(shares)
    return [fp(v) for v in opened[: len(shares)]]
