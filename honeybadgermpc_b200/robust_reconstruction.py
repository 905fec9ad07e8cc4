"""Robust opening of ONE shared value: what ``Mpc.open_share`` awaits
(reference: honeybadgermpc/robust_reconstruction.py:14-30).  It is the
incremental decoder with a batch of one, so it runs the same kernels as
``batch_reconstruct`` -- latency bound, kept for API completeness."""

from . import reed_solomon as rs
from .batch_reconstruction import fetch_one


def _codec_for(point, t):
    kind = rs.Algorithm.FFT if point.use_omega_powers else rs.Algorithm.VANDERMONDE
    return (rs.EncoderFactory.get(point, kind), rs.DecoderFactory.get(point, kind),
            rs.RobustDecoderFactory.get(t, point, algorithm=rs.Algorithm.GAO))


async def robust_reconstruct(field_futures, field, n, t, point, degree):
    """``field_futures[i]`` resolves to party i's share (a ``GFElement``).
    Returns ``(coefficients of the opened polynomial, error parties)``; the
    reference wraps the coefficients in its pure-Python ``Polynomial`` class,
    which is outside this path.  ``(None, None)`` if too few shares arrive."""
    state = rs.IncrementalDecoder(*_codec_for(point, t), degree, 1, t)
    async for party, share in fetch_one(field_futures):
        state.add(party, [share.value])
        if not state.done():
            continue
        rows, errors = state.get_results()
        return rows[0], errors
    return None, None
