"""tests/sanitize_small.py -- one small invocation of every kernel that is AUTO for some BASELINE config, for
`compute-sanitizer --tool memcheck|racecheck python tests/sanitize_small.py` (summaries under profiles/).

  Hubbard 4x3, 6 up 6 down (dim 853 776): k_sweep_down_lean + k_sweep_up_packed, k_axpy_norm, device-resident Lanczos scalars,
                                           stored CRS (k_crs_count/fill, k_spmv_crs), k_apply_op, k_spmv_table, k_spmv_generic
  FeAs 2x2 two orbitals (k_sweep_twospin_tab + the packed up sweep with two hop magnitudes)
  Heisenberg 16-ring (k_spmv_heis)
  Hubbard 14-chain 7 up 2 down (blocked up sweep of config 5: k_sweep_up_pipe)
Every result is checked against the oracle, so a run under the sanitizer is also a parity run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lanczosplusplus_b200 as lpp  # noqa: E402
from lanczosplusplus_b200 import geometry as geo  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests import cases  # noqa: E402

lpp.build()
orc.build()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def run(name, case, kernels, lanczos=True):
    o = cases.make_oracle(orc, case, fast_rank=1)
    e = cases.make_engine(lpp, case)
    n = e.rows()
    y = geo.splitmix64_vector(n, 42)
    x0 = geo.splitmix64_vector(n, 7)
    xr = x0.copy()
    o.matvec(xr, y, faithful=False)
    for k in kernels:
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=k)
        err = relerr(x, xr)
        assert err < 1e-13, (name, k, err)
    if lanczos:
        init = geo.splitmix64_vector(n, 1234)
        a, b, _ = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=12, eps=0.0)).decomposition(init)
        a0, b0 = o.decomposition(init, steps=12, eps=0.0)
        assert relerr(a, a0) < 1e-10 and relerr(b[:-1], b0[:-1]) < 1e-10, name
        en, _, _, _ = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=60, eps=1e-10)).computeOneState(init, want_vector=True)
    print("ok", name, "dim", n, flush=True)
    return e, o


e, o = run("hubbard 4x3", cases.hubbard_square(4, 3, 6, 6), (lpp.KERNEL_AUTO, lpp.KERNEL_TILED, lpp.KERNEL_TABLE, lpp.KERNEL_GENERIC, lpp.KERNEL_STORED))
dst = e.sector(5, 6)
e.apply_op(dst, lpp.OP_C, 1, 0, 1.0)
a, b, _ = lpp.LanczosSolver(dst, lpp.ParametersForSolver(steps=8, eps=0.0)).decomposition(None, use_modified=True)
print("ok apply_op + continued-fraction decomposition", len(a), flush=True)
dst.close()
e.close()
run("feas 2x2", cases.SMALL_CASES["feas_2x2"], (lpp.KERNEL_AUTO, lpp.KERNEL_GENERIC))[0].close()
run("feas chain 4 (inter-orbital hoppings)", cases.feas_chain(4, 2, 2, inter_orbital=0.5), (lpp.KERNEL_AUTO,))[0].close()
run("heisenberg 16", cases.heisenberg_ring(16, 8), (lpp.KERNEL_AUTO, lpp.KERNEL_GENERIC))[0].close()
run("hubbard 14-chain 7 up 2 down", cases.hubbard_chain(14, 7, 2), (lpp.KERNEL_AUTO,))[0].close()
run("tj 3x3", cases.TJ_CASES["tj_3x3"], (lpp.KERNEL_AUTO, lpp.KERNEL_STORED), lanczos=False)[0].close()
print("ALL OK")
