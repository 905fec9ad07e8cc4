"""Robust opening of a single share (reference:
honeybadgermpc/robust_reconstruction.py:14-30): the incremental decoder with a
batch of one.  Latency bound -- it exists for API completeness of ``Mpc.open_share``."""

from .batch_reconstruction import fetch_one
from .reed_solomon import (
    Algorithm,
    DecoderFactory,
    EncoderFactory,
    IncrementalDecoder,
    RobustDecoderFactory,
)


async def robust_reconstruct(field_futures, field, n, t, point, degree):
    """Returns ``(coefficient list of the opened polynomial, error parties)``
    (the reference wraps the coefficients in its pure-Python ``Polynomial``)."""
    algo = Algorithm.FFT if point.use_omega_powers else Algorithm.VANDERMONDE
    inc = IncrementalDecoder(EncoderFactory.get(point, algo), DecoderFactory.get(point, algo),
                             RobustDecoderFactory.get(t, point, algorithm=Algorithm.GAO),
                             degree, 1, t)
    async for idx, value in fetch_one(field_futures):
        inc.add(idx, [value.value])
        if inc.done():
            rows, errors = inc.get_results()
            return rows[0], errors
    return None, None
