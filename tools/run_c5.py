"""Config 5 of BASELINE.json end to end, row-sharded over the ranks of one node:
HubbardOneOrbital 1D chain (default 18 sites, 9 up 9 down, dim 2 363 904 400), ground state, then the continued-fraction
local Green's function of `c_i` / `c_i^dagger` (spin up) at site `--site` (Engine.h:133-206 type loop).

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c5.py [--sites 18] [--steps 200]

Checks that need no reference run: the sum rule  <c_i c_i^dagger> + <c_i^dagger c_i> = 1.  The reference applies the
operator at isite and accumulates it again at jsite (Engine.h:509-531), so for isite == jsite the modified state is
2 c_i |gs> and the two continued-fraction weights add up to 4; and -- at sizes a single GPU of the node can hold quickly (--compare) -- agreement with the
unsharded engine.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lanczosplusplus_b200 as lpp  # noqa: E402
from lanczosplusplus_b200 import distributed as D, geometry as geo  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=18)
ap.add_argument("--site", type=int, default=8)
ap.add_argument("--steps", type=int, default=200)         # SpectralSteps
ap.add_argument("--gs-steps", type=int, default=300)
ap.add_argument("--compare", action="store_true")
ap.add_argument("--ndown", type=int, default=-1)          # default: half filling; fewer down electrons keep 18-site runs small
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = args.sites
ndown = n // 2 if args.ndown < 0 else args.ndown
kw = dict(model=lpp.HUBBARD, nsite=n, nup=n // 2, ndown=ndown, hop=geo.chain(n, -1.0, False), U=np.full(n, 4.0), V=np.zeros(n))
io = {"LanczosSteps": args.gs_steps, "LanczosEps": 1e-12, "SpectralSteps": args.steps, "SpectralEps": 0.0}

t0 = time.perf_counter()
eng = lpp.InternalProductCuda(device=local, rank=rank, nranks=world, **kw)
if world > 1:
    D.attach(eng, dist)
t1 = time.perf_counter()
en = lpp.Engine(eng, io)                                    # ground state (+ eigenvector by the replayed recurrence)
torch.cuda.synchronize()
t2 = time.perf_counter()
cfs = en.spectralFunction(lpp.OP_C, args.site, args.site, spin=0)
torch.cuda.synchronize()
t3 = time.perf_counter()
omega = np.linspace(-6.0, 6.0, 25)
out = {"config": "HubbardOneOrbital %d-site open chain t=-1 U=4, %d up %d down" % (n, n // 2, ndown), "rows": eng.rows(),
       "ranks": world, "energy": en.energy, "gs_lanczos_steps": len(en.a), "setup_s": t1 - t0, "ground_state_s": t2 - t1,
       "gs_s_per_iteration": (t2 - t1) / (2 * max(len(en.a), 1)),   # decomposition + replay for the eigenvector
       "cf_s": t3 - t2, "cf_s_per_iteration": (t3 - t2) / (args.steps * max(len(cfs), 1)), "cf": []}
wsum = 0.0
for typ, cf in cfs:
    g = cf(omega, 0.1)
    out["cf"].append({"type": typ, "weight": cf.weight, "steps": int(cf.a.size), "a0": float(cf.a[0]), "b0": float(cf.b[0]),
                      "minus_im_g_over_pi_max": float((-g.imag / np.pi).max())})
    wsum += abs(cf.weight)
out["sum_rule_weights_over_4"] = wsum / 4.0                     # <c c^dagger> + <c^dagger c> = 1
ok = abs(wsum / 4.0 - 1.0) < 1e-8 and all(0.0 <= abs(c["weight"]) <= 4.0 for c in out["cf"])
if args.compare:
    single = lpp.InternalProductCuda(device=local, **kw)
    en1 = lpp.Engine(single, io)
    cfs1 = en1.spectralFunction(lpp.OP_C, args.site, args.site, spin=0)
    de = abs(en1.energy - en.energy)
    dw = max(abs(c1.weight - c.weight) for (_, c1), (_, c) in zip(cfs1, cfs))
    dg = max(np.abs(c1(omega, 0.1) - c(omega, 0.1)).max() for (_, c1), (_, c) in zip(cfs1, cfs))
    out["vs_single_gpu"] = {"d_energy": de, "d_weight": dw, "d_spectrum": float(dg)}
    ok = ok and de < 1e-9 and dw < 1e-9 and dg < 1e-6
    single.close()
out["ok"] = bool(ok)
if rank == 0:
    print(json.dumps(out), flush=True)
eng.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
