import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lanczosplusplus_b200 as lpp
from lanczosplusplus_b200 import distributed as D, geometry as geo
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
which = sys.argv[1]
if which == "chain":
    kw = dict(model=lpp.HUBBARD, nsite=18, nup=9, ndown=2, hop=geo.chain(18, -1.0), U=np.full(18, 4.0), V=np.zeros(18))
else:
    kw = dict(model=lpp.HUBBARD, nsite=12, nup=6, ndown=6, hop=geo.square(4, 3, -1.0), U=np.full(12, 4.0), V=np.zeros(12))
e = lpp.InternalProductCuda(device=local, rank=rank, nranks=world, **kw)
D.attach(e, dist)
try:
    for steps in (3, 40):
        a, b, _ = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=steps, eps=0.0, seed=1234)).decomposition(None)
        print("rank", rank, which, "steps", steps, "a0", a[0], "ok", flush=True)
except Exception as ex:
    print("rank", rank, which, "FAILED", ex, flush=True)
dist.barrier()
dist.destroy_process_group()
