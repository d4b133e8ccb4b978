"""ctypes loader for libsfb.so, the C-ABI engine declared in include/sfb.h.

There is deliberately no fallback: if the CUDA library is missing or no B200 is visible, importing works
(so that CPU-only tooling can inspect the package) but every numerical entry point raises.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsfb.so")
CSRC = os.path.join(_HERE, "csrc")


class SfbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsfb error {code}: {msg}")
        self.code = code


class SfbQpParams(C.Structure):
    """sfb_qp_params == QPSolverParams (reference qp_solver.hpp:29-68), field for field."""

    _fields_ = [
        ("verbose", C.c_int32),
        ("alpha", C.c_float),
        ("rho", C.c_float),
        ("sigma", C.c_float),
        ("scaling", C.c_int32),
        ("eps_abs", C.c_float),
        ("eps_rel", C.c_float),
        ("eps_primal_inf", C.c_float),
        ("eps_dual_inf", C.c_float),
        ("has_max_iter", C.c_int32),
        ("max_iter", C.c_uint32),
        ("has_max_time", C.c_int32),
        ("max_time_ns", C.c_int64),
        ("stop_check_iter", C.c_uint32),
        ("polish", C.c_int32),
        ("polish_iter", C.c_uint32),
        ("delta", C.c_float),
    ]


# every symbol include/sfb.h declares (tests check the library exports exactly these)
EXPORTED_SYMBOLS = [
    "sfb_version", "sfb_error_string", "sfb_last_error_message", "sfb_create", "sfb_destroy", "sfb_set_stream",
    "sfb_synchronize", "sfb_set_option", "sfb_kernel_launch_count", "sfb_qp_params_default", "sfb_qp_solve_dense_batch_f64",
    "sfb_qp_solve_dense_batch_f32", "sfb_qp_dense_max_m", "sfb_qp_scale_dense_batch_f64",
    "sfb_ekf_predict_batch_f64", "sfb_ekf_update_batch_f64", "sfb_ekf_step_batch_f64",
    "sfb_qp_sparse_analyze", "sfb_qp_sparse_symbolic", "sfb_qp_sparse_cta_selfcheck", "sfb_qp_sparse_uses_onchip", "sfb_qp_sparse_pattern_destroy", "sfb_qp_sparse_pattern_info",
    "sfb_qp_solve_sparse_batch_f64", "sfb_qp_solve_sparse_batch_f32", "sfb_qp_sparse_analyze_csc", "sfb_qp_solve_sparse_batch_csc_f64",
    "sfb_asif_vehicle_params_default", "sfb_asif_fleet_create", "sfb_asif_fleet_destroy", "sfb_asif_fleet_reset_warmstart",
    "sfb_asif_fleet_set_warmstart", "sfb_asif_fleet_filter_f64", "sfb_asif_fleet_filter_f32", "sfb_asif_fleet_to_qp_f64",
    "sfb_mpc_vehicle_params_default", "sfb_mpc_fleet_create", "sfb_mpc_fleet_destroy", "sfb_mpc_fleet_reset_warmstart",
    "sfb_mpc_fleet_dims", "sfb_mpc_fleet_pattern", "sfb_mpc_fleet_to_qp_f64", "sfb_mpc_fleet_step_f64", "sfb_mpc_fleet_step_f32",
    "sfb_mpc_fleet_nodes", "sfb_mpc_fleet_trajectories_f64", "sfb_mpc_fleet_trajectories_f32",
    "sfb_comm_unique_id", "sfb_comm_create", "sfb_comm_destroy", "sfb_allgather_results", "sfb_comm_wait",
]


def build(verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU). Returns the library path."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC], stdout=out)
    return LIB_PATH


def _preload_cudart() -> None:
    # libsfb links libcudart.so.12 dynamically so that it shares one runtime with torch
    try:
        C.CDLL("libcudart.so.12", mode=C.RTLD_GLOBAL)
        return
    except OSError:
        pass
    import sysconfig

    for base in {sysconfig.get_paths()["purelib"], sysconfig.get_paths()["platlib"]}:
        for p in glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib", "libcudart.so.12*")):
            C.CDLL(p, mode=C.RTLD_GLOBAL)
            return
    for p in glob.glob("/usr/local/cuda/lib64/libcudart.so.12*"):
        C.CDLL(p, mode=C.RTLD_GLOBAL)
        return


OPT_DUAL_INF_DX_GUARD = 1
OPT_FORCE_POLISH_SCRATCH = 2
OPT_POLISH_FORM = 3

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SfbError(-1, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    _preload_cudart()
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    L.sfb_version.restype = C.c_int
    L.sfb_error_string.argtypes = [C.c_int]
    L.sfb_error_string.restype = C.c_char_p
    L.sfb_last_error_message.argtypes = [vp]
    L.sfb_last_error_message.restype = C.c_char_p
    L.sfb_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.sfb_destroy.argtypes = [vp]
    L.sfb_set_stream.argtypes = [vp, vp]
    L.sfb_synchronize.argtypes = [vp]
    L.sfb_set_option.argtypes = [vp, i32, i32]
    L.sfb_kernel_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.sfb_qp_params_default.argtypes = [C.POINTER(SfbQpParams)]
    L.sfb_qp_params_default.restype = None
    qp_sig = [vp, C.POINTER(SfbQpParams), i64, i32, i32] + [vp] * 14
    L.sfb_qp_solve_dense_batch_f64.argtypes = qp_sig
    L.sfb_qp_solve_dense_batch_f32.argtypes = qp_sig
    L.sfb_qp_dense_max_m.argtypes = [vp, i32, i32]
    L.sfb_qp_scale_dense_batch_f64.argtypes = [vp, i64, i32, i32] + [vp] * 6
    L.sfb_ekf_predict_batch_f64.argtypes = [vp, i64, i32, i32, vp, vp, vp, C.c_double, C.c_double, vp]
    L.sfb_ekf_update_batch_f64.argtypes = [vp, i64, i32, i32] + [vp] * 6
    L.sfb_ekf_step_batch_f64.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp, C.c_double, C.c_double] + [vp] * 5
    L.sfb_qp_sparse_analyze.argtypes = [vp, i32, i32, vp, vp, vp, vp, C.POINTER(vp)]
    L.sfb_qp_sparse_symbolic.argtypes = [i32, i32, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(i64), vp, vp]
    L.sfb_qp_sparse_cta_selfcheck.argtypes = [i32, i32, vp, vp, vp, vp, i32, vp, C.POINTER(C.c_double)]
    L.sfb_qp_sparse_uses_onchip.argtypes = [vp, vp, i32, vp]
    L.sfb_qp_sparse_pattern_destroy.argtypes = [vp]
    L.sfb_qp_sparse_pattern_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), vp]
    sp_sig = [vp, vp, C.POINTER(SfbQpParams), i64] + [vp] * 14
    L.sfb_qp_solve_sparse_batch_f64.argtypes = sp_sig
    L.sfb_qp_solve_sparse_batch_f32.argtypes = sp_sig
    L.sfb_qp_sparse_analyze_csc.argtypes = [vp, i32, i32, vp, vp, vp, vp, C.POINTER(vp)]
    L.sfb_qp_solve_sparse_batch_csc_f64.argtypes = sp_sig
    L.sfb_asif_vehicle_params_default.argtypes = [vp]
    L.sfb_asif_vehicle_params_default.restype = None
    L.sfb_asif_fleet_create.argtypes = [vp, vp, i64, i32, C.POINTER(vp)]
    L.sfb_asif_fleet_destroy.argtypes = [vp]
    L.sfb_asif_fleet_reset_warmstart.argtypes = [vp]
    L.sfb_asif_fleet_set_warmstart.argtypes = [vp, i32]
    L.sfb_asif_fleet_filter_f64.argtypes = [vp] * 6
    L.sfb_asif_fleet_filter_f32.argtypes = [vp] * 6
    L.sfb_asif_fleet_to_qp_f64.argtypes = [vp] * 8
    L.sfb_mpc_vehicle_params_default.argtypes = [vp]
    L.sfb_mpc_vehicle_params_default.restype = None
    L.sfb_mpc_fleet_create.argtypes = [vp, vp, i64, i32, C.POINTER(vp)]
    L.sfb_mpc_fleet_destroy.argtypes = [vp]
    L.sfb_mpc_fleet_reset_warmstart.argtypes = [vp]
    L.sfb_mpc_fleet_dims.argtypes = [vp] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(i64)]
    L.sfb_mpc_fleet_pattern.argtypes = [vp] * 5
    L.sfb_mpc_fleet_to_qp_f64.argtypes = [vp] * 8
    L.sfb_mpc_fleet_step_f64.argtypes = [vp] * 8
    L.sfb_mpc_fleet_step_f32.argtypes = [vp] * 8
    L.sfb_mpc_fleet_nodes.argtypes = [vp, C.POINTER(C.c_int), vp]
    L.sfb_mpc_fleet_trajectories_f64.argtypes = [vp] * 4
    L.sfb_mpc_fleet_trajectories_f32.argtypes = [vp] * 4
    L.sfb_comm_unique_id.argtypes = [vp]
    L.sfb_comm_create.argtypes = [vp, i32, i32, vp, C.POINTER(vp)]
    L.sfb_comm_destroy.argtypes = [vp]
    L.sfb_allgather_results.argtypes = [vp, i32, vp, vp, vp]
    L.sfb_comm_wait.argtypes = [vp, i32]
    for name in EXPORTED_SYMBOLS:
        getattr(L, name)
    _lib = L
    return L


class Handle:
    """Owns one sfb_handle_t (one per host thread / stream, like a QPSolver object)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        L = lib()
        rc = L.sfb_create(int(device), C.c_void_p(stream or 0), C.byref(self._h))
        if rc != 0:
            raise SfbError(rc, L.sfb_last_error_message(None).decode())
        self.device = int(device)

    def check(self, rc: int) -> None:
        if rc != 0:
            raise SfbError(rc, lib().sfb_last_error_message(self._h).decode() or lib().sfb_error_string(rc).decode())

    @property
    def raw(self) -> C.c_void_p:
        return self._h

    def set_stream(self, stream: int | None) -> None:
        self.check(lib().sfb_set_stream(self._h, C.c_void_p(stream or 0)))

    def set_option(self, option: int, value: int) -> None:
        self.check(lib().sfb_set_option(self._h, int(option), int(value)))

    def synchronize(self) -> None:
        self.check(lib().sfb_synchronize(self._h))

    def launch_count(self) -> int:
        v = C.c_uint64()
        self.check(lib().sfb_kernel_launch_count(self._h, C.byref(v)))
        return int(v.value)

    def close(self) -> None:
        if self._h:
            lib().sfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
