"""Run the reference's OWN test files, unchanged, with the oracle -- or our shim
on top of the oracle -- standing in for the compiled NTL extension.  Only possible where /root/reference exists
(the authoring container); skipped elsewhere.  This is the strongest pin on
the oracle short of NTL itself: the reference's encoders, decoders,
IncrementalDecoder, batch_reconstruct, randousha and refinement programs all
run on top of it."""

import os
import subprocess
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_shim  # noqa: E402

FILES = [
    "tests/test_ntl.py",
    "tests/test_reed_solomon.py",
    "tests/test_reed_solomon_wb.py",
    "tests/test_polynomial.py",
    "tests/test_batch_reconstruction.py",
    "tests/test_offline_randousha.py",
    "tests/test_mpc.py",
    "tests/progs/test_random_refinement.py",
    "tests/progs/test_triple_refinement.py",
]
# Not about this path: pypairing (Rust) fixtures; and two tests that rely on
# Python 3.7 cancellation semantics (UnboundLocalError in the reference's own
# batch_reconstruction.py:183 on Python >= 3.8).
DESELECT = ["rust", "reconstruction_timeout"]


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("impl", ["oracle", "b200-host"])
def test_reference_tests_pass(impl):
    """impl = "oracle": the oracle module stands in for the NTL extension (pins the oracle).
    impl = "b200-host": OUR ctypes shim (honeybadgermpc_b200.ntl) stands in for it, with the
    oracle behind the native Context interface instead of the CUDA library -- the drop-in
    boundary of INTEGRATION.md exercised by the reference's own callers and tests, on CPU
    (argument conventions, shapes, padding / truncation, error behaviour of the shim)."""
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "pytest.ini"), "w") as fh:
            fh.write("[pytest]\n")
        env = dict(os.environ, PYTHONPATH=os.path.join(HERE, "golden"), HBMPC_NTL_IMPL=impl)
        cmd = [sys.executable, "-m", "pytest", "-c", os.path.join(tmp, "pytest.ini"),
               "--rootdir", tmp, "-p", "ref_plugin", "-p", "no:cacheprovider", "-q",
               "-k", " and ".join(f"not {d}" for d in DESELECT)]
        cmd += [os.path.join(ref_shim.REFERENCE_ROOT, f) for f in FILES]
        res = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
        tail = res.stdout[-2000:] + res.stderr[-2000:]
        assert res.returncode == 0, tail
        assert " passed" in res.stdout and "failed" not in res.stdout.splitlines()[-1], tail


# the reference's tests of the path's classes, run against OUR mirror modules (field, reed_solomon,
# batch_reconstruction, robust_reconstruction injected in place of the reference's)
# all nine (Mpc.open / ShareArray.open, randousha and the refinement programs included), plus the
# reference's tests of the field class and of the preprocessing files written through the encoders
MIRROR_FILES = FILES + ["tests/test_field.py", "tests/test_preprocessing.py",
                        "tests/progs/mixins/test_share_arithmetic.py"]


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_reference_tests_pass_on_our_mirror_modules():
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "pytest.ini"), "w") as fh:
            fh.write("[pytest]\n")
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(HERE, "golden"), HERE]))
        cmd = [sys.executable, "-m", "pytest", "-c", os.path.join(tmp, "pytest.ini"),
               "--rootdir", tmp, "-p", "ref_plugin_mirror", "-p", "no:cacheprovider", "-q",
               "--timeout", "120", "-k", " and ".join(f"not {d}" for d in DESELECT)]
        cmd += [os.path.join(ref_shim.REFERENCE_ROOT, f) for f in MIRROR_FILES]
        res = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
        tail = res.stdout[-2000:] + res.stderr[-2000:]
        assert res.returncode == 0, tail
        assert " passed" in res.stdout and "failed" not in res.stdout.splitlines()[-1], tail


BENCH_FILES = [
    "benchmark/test_benchmark_batch_opening.py",   # BASELINE.json configs[0]: n=4, t=1, ShareArray.open()
    "benchmark/test_benchmark_refinement.py",
    "benchmark/test_benchmark_preprocessing.py",
    "benchmark/test_benchmark_reed_solomon.py",     # Gao robust decode up to t = 256, n = 769
]


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
def test_reference_benchmark_logic_on_our_mirror():
    """The reference's benchmark files for this path -- TaskProgramRunner with 4 / 7 parties in one
    process opening up to 1024 random shares (BASELINE.json configs[0]), refinement, preprocessing
    files, Gao robust decoding -- run unchanged (each benchmarked function once, untimed) on top of
    our mirror modules and shim.  Deselected: the `use_fft=` variants, a keyword the reference's own
    EvalPoint does not have either."""
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "pytest.ini"), "w") as fh:
            fh.write("[pytest]\n")
        env = dict(os.environ, PYTHONPATH=os.pathsep.join(
            [os.path.join(HERE, "golden"), HERE, ref_shim.REFERENCE_ROOT]))
        cmd = [sys.executable, "-m", "pytest", "-c", os.path.join(tmp, "pytest.ini"),
               "--rootdir", tmp, "-p", "ref_plugin_bench", "-p", "no:cacheprovider", "-q",
               "--timeout", "120", "-k", "not fft"]
        cmd += [os.path.join(ref_shim.REFERENCE_ROOT, f) for f in BENCH_FILES]
        res = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
        tail = res.stdout[-2000:] + res.stderr[-2000:]
        assert res.returncode == 0, tail
        assert " passed" in res.stdout and "failed" not in res.stdout.splitlines()[-1], tail
