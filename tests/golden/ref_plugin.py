"""pytest plugin used to run the REFERENCE's own test files, unchanged, from
/root/reference/tests against a stand-in for its compiled NTL extension.

    HBMPC_NTL_IMPL=oracle|b200|b200-host  python -m pytest -p ref_plugin /root/reference/tests/test_ntl.py

Test infrastructure only (authoring container; /root/reference does not exist
on the GPU box).  Replaces pytest-asyncio (absent here) with a tiny hook."""

import asyncio
import importlib
import inspect
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_shim  # noqa: E402

_impl = os.environ.get("HBMPC_NTL_IMPL", "oracle")
if _impl == "oracle":
    _mod = importlib.import_module("oracle.hbmpc_oracle")
elif _impl == "b200-host":
    # our ctypes shim (argument handling, shapes, error behaviour) with the oracle standing in for
    # the CUDA library behind `_native.Context` (tests/host_backend.py): runs without a GPU
    sys.path.insert(0, os.path.dirname(HERE))
    _hb = importlib.import_module("host_backend")

    class _Patch:
        @staticmethod
        def setattr(obj, name, value):
            setattr(obj, name, value)

    _hb.install(_Patch)
    _mod = importlib.import_module("honeybadgermpc_b200.ntl")
else:
    _mod = importlib.import_module("honeybadgermpc_b200.ntl")
ref_shim.install(_mod)


def pytest_configure(config):
    config.addinivalue_line("markers", "asyncio: run the coroutine test on a fresh event loop")


def pytest_pyfunc_call(pyfuncitem):
    fn = pyfuncitem.obj
    if inspect.iscoroutinefunction(fn):
        kwargs = {a: pyfuncitem.funcargs[a] for a in pyfuncitem._fixtureinfo.argnames}
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)
        try:
            loop.run_until_complete(fn(**kwargs))
        finally:
            loop.close()
        return True
    return None
