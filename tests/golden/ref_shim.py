"""Make the *pure-Python* parts of the reference importable in the authoring
container (it has no NTL, gmpy2, pypairing).  Test infrastructure only.

Used by ``make_golden.py`` (to generate the committed fixtures) and by
``tests/test_reference_suite.py`` (which runs only where ``/root/reference``
exists; it does not exist on the GPU box).

What is stubbed, and why it does not weaken the fixtures:
  * ``gmpy2``      -> ``is_prime`` via sympy, ``mpz`` = int   (field.py:25,53)
  * ``pypairing``  -> seven empty classes (betterpairing.py:6 imports names only)
  * ``honeybadgermpc.ntl._hbmpc_ntl_helpers`` -> whatever module the caller
    passes (the compiled NTL extension cannot be built here).  The golden
    generator only records outputs of reference code that never calls it.
  * ``logging.config.dictConfig`` is a no-op during ``import honeybadgermpc``
    (logging.yaml wants /var/log/hbmpc/).
"""

import importlib
import logging.config
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "honeybadgermpc"))


def install(ntl_module):
    """Import the reference package with ``ntl_module`` standing in for the
    compiled NTL extension.  Returns the ``honeybadgermpc`` package."""
    if not reference_available():
        raise RuntimeError("reference tree not present")

    if "gmpy2" not in sys.modules:
        import sympy

        g = types.ModuleType("gmpy2")
        g.is_prime = lambda v: bool(sympy.isprime(int(v)))
        g.mpz = int
        sys.modules["gmpy2"] = g
    if "pypairing" not in sys.modules:
        pp = types.ModuleType("pypairing")
        for name in ("PyFq", "PyFq12", "PyFq2", "PyFqRepr", "PyFr", "PyG1", "PyG2"):
            setattr(pp, name, type(name, (), {}))
        sys.modules["pypairing"] = pp

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    if "honeybadgermpc" not in sys.modules:
        saved = logging.config.dictConfig
        saved_stdout = sys.stdout
        logging.config.dictConfig = lambda cfg: None
        sys.stdout = open(os.devnull, "w")  # honeybadgermpc/__init__.py prints
        try:
            importlib.import_module("honeybadgermpc")
        finally:
            sys.stdout.close()
            sys.stdout = saved_stdout
            logging.config.dictConfig = saved

    helpers = types.ModuleType("honeybadgermpc.ntl._hbmpc_ntl_helpers")
    names = [
        "lagrange_interpolate", "evaluate", "vandermonde_inverse", "InterpolationError",
        "vandermonde_batch_interpolate", "vandermonde_batch_evaluate", "fft", "partial_fft",
        "fft_batch_evaluate", "fft_interpolate", "fft_batch_interpolate", "SetNTLNumThreads",
        "AvailableNTLThreads", "gao_interpolate", "sqrt_mod", "SetNumThreads", "GetMaxThreads",
    ]
    for name in names:
        setattr(helpers, name, getattr(ntl_module, name))
    helpers.__all__ = names
    sys.modules["honeybadgermpc.ntl._hbmpc_ntl_helpers"] = helpers
    # (re)bind honeybadgermpc.ntl and every module that did `from ...ntl import`
    for mod in [m for m in sys.modules if m.startswith("honeybadgermpc.") and m != helpers.__name__]:
        del sys.modules[mod]
    importlib.import_module("honeybadgermpc.ntl")
    return sys.modules["honeybadgermpc"]
