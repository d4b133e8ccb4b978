"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, exports every symbol include/sfb.h
declares, and fails loudly (no fallback) without a device.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from smooth_feedback_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sfb.h")).read()
    declared = sorted(set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    L = lib.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sfb.h but not exported by libsfb.so"
    assert sorted(lib.EXPORTED_SYMBOLS) == declared


def test_params_default_matches_reference(lib):
    # QPSolverParams defaults, reference qp_solver.hpp:29-68
    p = lib.SfbQpParams()
    lib.lib().sfb_qp_params_default(C.byref(p))
    import numpy as np

    f32 = lambda v: float(np.float32(v))
    assert (p.alpha, p.rho, p.sigma) == (f32(1.6), f32(0.1), f32(1e-6))
    assert (p.eps_abs, p.eps_rel, p.eps_primal_inf, p.eps_dual_inf) == (f32(1e-3), f32(1e-3), f32(1e-4), f32(1e-4))
    assert (p.scaling, p.polish, p.polish_iter, p.stop_check_iter, p.has_max_iter, p.has_max_time) == (1, 1, 5, 25, 0, 0)
    assert p.delta == f32(1e-6)
    from smooth_feedback_b200.qp import QPSolverParams

    q = QPSolverParams().to_c()
    for name, _ in lib.SfbQpParams._fields_:
        assert getattr(p, name) == getattr(q, name), name


def test_status_enum_order(lib):
    from smooth_feedback_b200.qp import QPSolutionStatus as S

    # qp.hpp:82-92
    assert [s.name for s in S] == ["Optimal", "PolishFailed", "PrimalInfeasible", "DualInfeasible", "MaxIterations",
                                   "MaxTime", "Unknown"]
    assert [int(s) for s in S] == list(range(7))


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from smooth_feedback_b200 import Handle, SfbError

    with pytest.raises(SfbError) as e:
        Handle(0)
    assert e.value.code == 2  # SFB_ERR_NO_DEVICE


def test_only_sm100a_code_in_library(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "smooth_feedback_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, fn)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle's", "").replace("oracle/", "") or \
                    not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle", src, re.M), fn


def test_header_is_plain_c(tmp_path):
    """include/sfb.h must be consumable from C (cgo / JNI / ctypes-style bindings): compile and link a C99 translation unit
    that takes the address of every declared entry point."""
    import re
    import subprocess

    from smooth_feedback_b200 import _lib

    hdr = os.path.join(ROOT, "include", "sfb.h")
    names = sorted(set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", open(hdr).read())))
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    src = tmp_path / "use_sfb.c"
    src.write_text('#include "sfb.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t p[] = {' +
                   ", ".join(f"(fn_t){n}" for n in names) +
                   '};\n  sfb_qp_params prm; sfb_qp_params_default(&prm);\n'
                   '  printf("%d %d %g\\n", (int)(sizeof(p) / sizeof(p[0])), sfb_version(), (double)prm.rho);\n  return prm.stop_check_iter == 25 ? 0 : 1;\n}\n')
    exe = tmp_path / "use_sfb"
    libdir = os.path.dirname(_lib.LIB_PATH)
    import glob
    import sysconfig

    cudart = None
    for base in {sysconfig.get_paths()["purelib"], sysconfig.get_paths()["platlib"]}:
        hits = glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib"))
        if hits:
            cudart = hits[0]
    cudart = cudart or "/usr/local/cuda/lib64"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}",
                           "-o", str(exe), str(src), f"-L{libdir}", "-lsfb", f"-Wl,-rpath,{libdir}", f"-L{cudart}", f"-Wl,-rpath,{cudart}"])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[0] == str(len(names)), r.stdout + r.stderr


def test_struct_layouts_match_ctypes(tmp_path):
    """The ctypes mirrors of the POD parameter structs must have the C compiler's layout (size and every offset)."""
    from smooth_feedback_b200 import _lib
    from smooth_feedback_b200.asif import SfbAsifVehicleParams
    from smooth_feedback_b200.mpc import SfbMpcVehicleParams

    structs = {"sfb_qp_params": _lib.SfbQpParams, "sfb_asif_vehicle_params": SfbAsifVehicleParams,
               "sfb_mpc_vehicle_params": SfbMpcVehicleParams}
    body = []
    for cname, ct in structs.items():
        body.append(f'printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in ct._fields_:
            body.append(f'printf(" %zu", offsetof({cname}, {fname}));')
        body.append('printf("\\n");')
    src = tmp_path / "layout.c"
    src.write_text('#include "sfb.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(void) {\n' + "\n".join(body) + "\nreturn 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", f"-I{os.path.join(ROOT, 'include')}", "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line in out:
        tok = line.split()
        ct = structs[tok[0]]
        assert int(tok[1]) == C.sizeof(ct), tok[0]
        assert [int(t) for t in tok[2:]] == [getattr(ct, f).offset for f, _ in ct._fields_], tok[0]


def test_fleet_param_defaults_match_python_mirrors():
    """sfb_*_vehicle_params_default (the constants of examples/mpc_asif_vehicle.cpp) == the defaults of the python dataclasses."""
    from smooth_feedback_b200 import ASIFVehicleParams, MPCVehicleParams, _lib
    from smooth_feedback_b200.asif import SfbAsifVehicleParams
    from smooth_feedback_b200.mpc import SfbMpcVehicleParams

    def flat(s):
        out = []
        for name, _ in s._fields_:
            v = getattr(s, name)
            if isinstance(v, C.Structure):
                out += flat(v)
            elif hasattr(v, "__len__"):
                out += [float(t) for t in v]
            else:
                out.append(v)
        return out

    a = SfbAsifVehicleParams(); _lib.lib().sfb_asif_vehicle_params_default(C.byref(a))
    assert flat(a) == flat(ASIFVehicleParams().to_c())
    m = SfbMpcVehicleParams(); _lib.lib().sfb_mpc_vehicle_params_default(C.byref(m))
    assert flat(m) == flat(MPCVehicleParams().to_c())
