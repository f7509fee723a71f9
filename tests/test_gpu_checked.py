"""The checked build of the CUDA library (-DDNMF_CHECKED, `python -m dnmf_b200.build --checked`): device-side
assertions on every index the kernels form without a clamp (frame ids, list lengths, the staged-window index of the
unclamped main loops, table rows of the slice gathers, partial-block slots).  compute-sanitizer is closed on the GPU
pool this was developed on, so a cross-section of the GPU suite -- ragged sizes, out-of-bounds deformations, every
kernel family, the full-size configurations -- runs once through that library in a child process: a failed assertion
is a sticky CUDA error there and the child's tests fail."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SELECTION = [
    "tests/test_gpu_parity.py",
    "tests/test_gpu_edge.py",
    "tests/test_gpu_aux.py",
    "tests/test_gpu_extension.py",
    "tests/test_gpu_fullsize.py::test_loss_and_gradient_vs_closed_form",
    "tests/test_gpu_fullsize.py::test_trace_statistics_vs_closed_form_all_paths",
]


def _checked_lib():
    from dnmf_b200 import build as b
    path = b.OUT_CHECKED
    if not os.path.isfile(path):
        path = b.build(checked=True)
    return path


def _env():
    env = dict(os.environ)
    env["DNMF_B200_LIB"] = _checked_lib()
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    return env


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_checked_build_assertions_are_live():
    code = ("import dnmf_b200; lib = dnmf_b200.load(); assert lib.dnmf_build_info() & 1, 'not a checked build'; "
            "rc = lib.dnmf_debug_trip_assert(0); print('rc', rc, lib.dnmf_last_error()); "
            "raise SystemExit(0 if rc != 0 else 3)")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "assert" in (res.stdout + res.stderr).lower()
    # the normal build compiles the assertions out
    import dnmf_b200
    lib = dnmf_b200.load()
    if not os.environ.get("DNMF_B200_LIB"):
        assert lib.dnmf_build_info() & 1 == 0


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_gpu_suite_cross_section_through_the_checked_build():
    if os.environ.get("DNMF_B200_LIB"):
        pytest.skip("already running against an explicit library")
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + SELECTION,
                         cwd=ROOT, env=_env(), capture_output=True, text=True, timeout=1500)
    tail = "\n".join((res.stdout + res.stderr).splitlines()[-25:])
    print(tail)
    assert res.returncode == 0, tail
