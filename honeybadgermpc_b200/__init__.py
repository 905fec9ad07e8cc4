"""honeybadgermpc_b200 -- B200-native batched secret-share reconstruction for
HoneyBadgerMPC: the path ``ShareArray.open()`` -> ``batch_reconstruct`` ->
``EncoderFactory/DecoderFactory`` -> ``honeybadgermpc.ntl`` re-built on
hand-written sm_100a CUDA kernels behind a C-ABI (``include/hbmpc_b200.h``).

Modules mirror the reference's names for this path only:
``ntl``, ``polynomial`` (``EvalPoint``, ``get_omega``), ``reed_solomon``,
``batch_reconstruction``, ``robust_reconstruction``, ``field``.
"""

__version__ = "0.1.0"
