"""pytest plugin: run the REFERENCE's own test files against OUR host-side mirror modules.

``honeybadgermpc.field``, ``.reed_solomon``, ``.batch_reconstruction`` and
``.robust_reconstruction`` are replaced by the modules of ``honeybadgermpc_b200`` before the
reference's tests import them; the compiled NTL extension is replaced by our ctypes shim,
and the CUDA library behind it by the oracle (tests/host_backend.py), so this runs on CPU.
Test infrastructure only (authoring container)."""

import os
import sys

os.environ["HBMPC_NTL_IMPL"] = "b200-host"
import ref_plugin  # noqa: E402,F401  (imports the reference with our shim as its NTL extension)
from ref_plugin import pytest_configure, pytest_pyfunc_call  # noqa: E402,F401

import honeybadgermpc  # noqa: E402
import honeybadgermpc_b200.batch_reconstruction as br  # noqa: E402
import honeybadgermpc_b200.field as fld  # noqa: E402
import honeybadgermpc_b200.reed_solomon as rs  # noqa: E402
import honeybadgermpc_b200.robust_reconstruction as rr  # noqa: E402

for name, mod in (("field", fld), ("reed_solomon", rs), ("batch_reconstruction", br),
                  ("robust_reconstruction", rr)):
    sys.modules["honeybadgermpc." + name] = mod
    setattr(honeybadgermpc, name, mod)
