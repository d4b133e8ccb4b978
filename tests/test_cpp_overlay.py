"""C++20 header overlays (include/smooth_feedback_b200/{qp_solver,ekf}.hpp): the reference's template surfaces over the C ABI,
compiled against the reference's own problem / solution type templates (Eigen / smooth / Boost stand-ins under
tests/cpp/mock_include, because none of them is installed in this image).

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
PROGRAMS = ["test_overlay", "replay_mpc_asif", "test_ekf_overlay"]


def _cudart_dir():
    for base in {sysconfig.get_paths()["purelib"], sysconfig.get_paths()["platlib"]}:
        hits = glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib"))
        if hits:
            return hits[0]
    return "/usr/local/cuda/lib64"


def _build(name="test_overlay"):
    from smooth_feedback_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = os.path.join(ROOT, "tests", "cpp", name)
    cmd = ["/usr/bin/g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", "-O1", f"-I{os.path.join(ROOT, 'include')}",
           f"-I{os.path.join(ROOT, 'tests', 'cpp', 'mock_include')}", "-o", exe,
           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), f"-L{libdir}", "-lsfb", f"-Wl,-rpath,{libdir}",
           f"-L{_cudart_dir()}", f"-Wl,-rpath,{_cudart_dir()}"]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.parametrize("name", PROGRAMS)
def test_overlay_compiles_and_fails_loudly_without_gpu(name):
    import torch

    exe = _build(name)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_overlay_runs_on_gpu")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PROGRAMS)
def test_overlay_runs_on_gpu(name):
    exe = _build(name)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " ok" in r.stdout, r.stdout + r.stderr
