"""Robust opening of ONE shared value: what ``Mpc.open_share`` awaits
(reference: honeybadgermpc/robust_reconstruction.py:14-30).  It is the
incremental decoder with a batch of one, so it runs the same kernels as
``batch_reconstruct`` -- latency bound, kept for API completeness."""

from . import reed_solomon as rs
from .batch_reconstruction import fetch_one
from .polynomial import OpenedPolynomial


def _codec_for(point, t):
    kind = rs.Algorithm.FFT if point.use_omega_powers else rs.Algorithm.VANDERMONDE
    return (rs.EncoderFactory.get(point, kind), rs.DecoderFactory.get(point, kind),
            rs.RobustDecoderFactory.get(t, point, algorithm=rs.Algorithm.GAO))


async def robust_reconstruct(field_futures, field, n, t, point, degree):
    """``field_futures[i]`` resolves to party i's share (a ``GFElement``).
    Returns ``(opened polynomial, error parties)``: the polynomial is the list of its
    coefficients and is callable like the reference's ``Polynomial`` object
    (``Mpc.open_share`` evaluates it at zero).  ``(None, None)`` if too few shares arrive."""
    state = rs.IncrementalDecoder(*_codec_for(point, t), degree, 1, t)
    async for party, share in fetch_one(field_futures):
        state.add(party, [share.value])
        if not state.done():
            continue
        rows, errors = state.get_results()
        return OpenedPolynomial(rows[0], field), errors
    return None, None
