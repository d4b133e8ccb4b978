"""C++20 header overlay (include/smooth_feedback_b200/qp_solver_b200.hpp): the reference's template surface over the C ABI.

CPU: it compiles against a minimal Eigen stand-in and fails LOUDLY without a device.  GPU: the transliterated reference
unit tests inside tests/cpp/test_overlay.cpp pass.
"""
import glob
import os
import subprocess
import sysconfig

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_overlay")


def _cudart_dir():
    for base in {sysconfig.get_paths()["purelib"], sysconfig.get_paths()["platlib"]}:
        hits = glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib"))
        if hits:
            return hits[0]
    return "/usr/local/cuda/lib64"


def _build():
    from smooth_feedback_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["/usr/bin/g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", "-O1", "-o", EXE,
           os.path.join(ROOT, "tests", "cpp", "test_overlay.cpp"), f"-L{libdir}", "-lsfb", f"-Wl,-rpath,{libdir}",
           f"-L{_cudart_dir()}", f"-Wl,-rpath,{_cudart_dir()}"]
    subprocess.check_call(cmd)


def test_overlay_compiles_and_fails_loudly_without_gpu():
    import torch

    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_overlay_runs_on_gpu")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_overlay_runs_on_gpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "overlay ok" in r.stdout, r.stdout + r.stderr
