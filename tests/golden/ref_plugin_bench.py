"""pytest plugin: the REFERENCE's benchmark files (benchmark/test_benchmark_*.py) on top of our
mirror modules (ref_plugin_mirror), with a stand-in for the pytest-benchmark fixture, which is not
installed here: the benchmarked function runs once.  Test infrastructure only."""

import asyncio

import pytest
from ref_plugin_mirror import pytest_configure, pytest_pyfunc_call  # noqa: F401


@pytest.fixture
def benchmark():
    def run(fn, *args, setup=None, **kwargs):
        try:
            asyncio.get_event_loop()
        except RuntimeError:
            asyncio.set_event_loop(asyncio.new_event_loop())
        if setup is not None:
            setup()
        return fn(*args, **kwargs)

    return run
