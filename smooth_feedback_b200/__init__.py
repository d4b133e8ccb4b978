"""B200-native batched QP / EKF engine behind smooth_feedback's QPSolver / solve_qp / EKF surfaces.

The numerics live in lib/libsfb.so (hand-written sm_100a CUDA, C ABI in include/sfb.h); this package is the
host-side mirror of the reference interface.  No CPU fallback exists.
"""
from ._lib import Handle, SfbError, build  # noqa: F401
from .qp import (  # noqa: F401
    QPBatchResult, QPSolution, QPSolutionStatus, QPSolver, QPSolverParams, QuadraticProgram, solve_dense_batch,
    solve_qp, to_colmajor,
)
from .qp_sparse import (QuadraticProgramSparse, SparsePattern, solve_sparse_batch, sparse_onchip_selfcheck,  # noqa: F401
                        sparse_symbolic)
from .asif import ASIFVehicleFleet, ASIFVehicleParams  # noqa: F401
from .comm import Communicator  # noqa: F401
from .mpc import MPCVehicleFleet, MPCVehicleParams  # noqa: F401
from .ekf import ekf_predict_batch, ekf_step_batch, ekf_step_batch_host, ekf_update_batch  # noqa: F401

__version__ = "0.2.0"
