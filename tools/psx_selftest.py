"""torchrun --nproc-per-node N tools/psx_selftest.py : the peer-memory scalar all-reduce of the sharded Krylov loop
(k_psx_allreduce: NVLink stores + sequence flags) against the expected sums, many rounds back to back."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lanczosplusplus_b200 as lpp  # noqa: E402
from lanczosplusplus_b200 import _lib, distributed as D, geometry as geo  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
e = lpp.InternalProductCuda(lpp.HUBBARD, 12, 6, 6, hop=geo.square(4, 3, -1.0), U=np.full(12, 4.0), V=np.zeros(12), device=local, rank=rank,
                            nranks=world)
D.attach(e, dist)
ok = True
for it in range(200):
    n = it % 5
    v = np.array([rank + 1.0 + 0.25 * k + it for k in range(4)])
    rc = _lib.lib().lpp_allreduce_selftest(e.h, v.ctypes.data, n)
    if rc != 0:
        print("rank", rank, "round", it, "status", rc, _lib.lib().lpp_last_error().decode(), flush=True)
        ok = False
        break
    want = [sum(r + 1.0 + 0.25 * k + it for r in range(world)) for k in range(n)]
    if not np.allclose(v[:n], want, rtol=0, atol=1e-12):
        print("rank", rank, "round", it, "got", v[:n], "want", want, flush=True)
        ok = False
        break
dist.barrier()
if rank == 0:
    print("PSX OK" if ok else "PSX FAIL", flush=True)
e.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
