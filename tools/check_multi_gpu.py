"""torchrun --nproc-per-node N tools/check_multi_gpu.py : the row-sharded Krylov loop (NCCL gather + all-reduce inside the
engine) must reproduce the single-GPU tridiagonal coefficients and energy on the same seeded initial vector.
LPP_PIPELINE=<chunks> in the environment runs the two-layout handles through the pipelined recurrence (lanczos_pipelined);
LPP_CHECK_FULL=0 skips config 3 at its named size."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lanczosplusplus_b200 as lpp  # noqa: E402
from lanczosplusplus_b200 import distributed as D, geometry as geo  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, kw in (("hubbard 18-chain 9 up 2 down (blocked up sweep)", dict(model=lpp.HUBBARD, nsite=18, nup=9, ndown=2, hop=geo.chain(18, -1.0), U=np.full(18, 4.0), V=np.zeros(18))),
                 ("hubbard 4x3", dict(model=lpp.HUBBARD, nsite=12, nup=6, ndown=6, hop=geo.square(4, 3, -1.0), U=np.full(12, 4.0), V=np.zeros(12))),
                 ("feas 2x3", dict(model=lpp.FEAS, nsite=6, nup=4, ndown=4, orbitals=2,
                                   hop=geo.with_orbitals(geo.square(2, 3, -1.0, False, False), 2, 1.0, 0.5),
                                   U=np.array([4.0, 3.0, -0.8, -0.4]), V=np.zeros(24), D=np.array([0.0]))),
                 ("heisenberg 20", dict(model=lpp.HEISENBERG, nsite=20, nup=10, hop=geo.chain(20, 1.0, True), jzz=geo.chain(20, 1.0, True)))):
    for kernel in (lpp.KERNEL_AUTO, lpp.KERNEL_GENERIC):
        sharded = lpp.InternalProductCuda(device=local, rank=rank, nranks=world, kernel=kernel, **kw)
        D.attach(sharded, dist)
        p = lpp.ParametersForSolver(steps=60, eps=0.0, seed=1234)
        a, b, _ = lpp.LanczosSolver(sharded, p).decomposition(None)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        a, b, _ = lpp.LanczosSolver(sharded, p).decomposition(None)
        torch.cuda.synchronize()
        ms_iter = 1e3 * (time.perf_counter() - t0) / len(a)
        single = lpp.InternalProductCuda(device=local, kernel=kernel, **kw)
        a1, b1, _ = lpp.LanczosSolver(single, p).decomposition(None)
        err = max(np.abs(a[:25] - a1[:25]).max(), np.abs(b[:25] - b1[:25]).max())
        e = lpp.tridiag_eig(a, b)[0]
        e1 = lpp.tridiag_eig(a1, b1)[0]
        good = err < 1e-10 and abs(e - e1) < 1e-9
        # the converged ground state (the convergence test runs on batches of device-resident steps): same energy, same step count
        pg = lpp.ParametersForSolver(steps=300, eps=1e-12, seed=1234)
        eg, _, ag, _ = lpp.LanczosSolver(sharded, pg).computeOneState(None, want_vector=False)
        eg1, _, ag1, _ = lpp.LanczosSolver(single, pg).computeOneState(None, want_vector=False)
        good = good and abs(eg - eg1) < 1e-9 and abs(len(ag) - len(ag1)) <= 1
        ok = ok and good
        if rank == 0:
            print("%-14s kernel %d ranks %d: max|d(a,b)| %.2e  E %.12f vs %.12f  %.3f ms/iteration  %s" % (name, kernel, world, err, e, e1, ms_iter, "OK" if good else "FAIL"), flush=True)
        sharded.close()
        single.close()
# config 3 at its named size (dim 165 636 900): first coefficients against the oracle pin and the committed 1-GPU values
import json  # noqa: E402
import bench  # noqa: E402
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("LPP_CHECK_FULL", "1") != "0":
    case, _ = bench.workload("c3")
    big = lpp.InternalProductCuda(case["model"], case["nsite"], case["nup"], case["ndown"], case["orbitals"], hop=case["hop"], U=case["U"],
                                  V=case["V"], device=local, rank=rank, nranks=world)
    D.attach(big, dist)
    a, b, _ = lpp.LanczosSolver(big, lpp.ParametersForSolver(steps=20, eps=0.0, seed=1234)).decomposition(None)
    chk = bench.checks_vs_fixtures("c3", np.asarray(a), np.asarray(b))
    pin = chk.get("oracle_pin", {})
    good = pin.get("alpha0_rel_err", 1.0) <= 1e-10 and pin.get("beta0_rel_err", 1.0) <= 1e-10
    if "vs_1gpu" in chk:
        good = good and chk["vs_1gpu"]["max_rel_diff"] <= 1e-10
    ok = ok and good
    if rank == 0:
        print("config 3 full size, ranks %d: %s  %s" % (world, json.dumps(chk), "OK" if good else "FAIL"), flush=True)
    big.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0 and ok:
    print("ALL OK", flush=True)
sys.exit(0 if ok else 1)
