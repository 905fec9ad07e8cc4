"""Robust (error-correcting) decoders -- filled in by the CUDA Gao / WB kernels."""


def gao_interpolate(x, y, k, modulus, z, omega, order, use_omega_powers):
    raise NotImplementedError("gao_interpolate: CUDA kernel not built yet")
