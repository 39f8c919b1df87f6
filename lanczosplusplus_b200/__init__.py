"""B200-native Lanczos engine for the hot path of g1257/LanczosPlusPlus (x += H y and the Lanczos recurrence).

The compute path is liblpp_b200.so (hand-written sm_100a CUDA behind the C-ABI of include/lpp_b200.h); this package is
the thin host-side mirror of the reference's interface.  There is no CPU fallback.
"""
from . import geometry  # noqa: F401
from ._lib import LppError, build, lib  # noqa: F401
from .engine import (FEAS, HEISENBERG, HUBBARD, TJ, KERNEL_AUTO, KERNEL_GENERIC, KERNEL_STORED, KERNEL_TABLE,  # noqa: F401
                     KERNEL_TILED, OP_C, OP_CDAGGER, OP_N, OP_SMINUS, OP_SPLUS, OP_SZ, ContinuedFraction, Engine, InternalProductCuda,
                     LanczosSolver, ParametersForSolver, comm_unique_id, kernel_from_solver_options, tridiag_eig)
