"""world_size-2 gloo test (CPU): the host-side logic of the multi-GPU path -- shard ranges, id broadcast -- and the
sharded algorithm itself (gather y, local rows of x += H y, all-reduce of the Lanczos scalars), with the oracle standing in
for the CUDA kernels as the per-rank row evaluator.  Result must equal the single-process recurrence."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lanczosplusplus_b200 as lpp
    from lanczosplusplus_b200 import distributed as D, geometry as geo
    from oracle import oracle as orc
    lpp.build()
    case = cases.hubbard_chain(6, 3, 2)
    o = cases.make_oracle(orc, case, fast_rank=1)
    n1, n2 = len(o.basis(0)), len(o.basis(1))
    r0, nloc = D.local_rows(n1, n2, rank, world)
    # every rank gets the same id bytes
    uid = D.broadcast_unique_id(dist, lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    # sharded Lanczos: y local, gathered to full each step; scalars all-reduced
    rows = o.rows()
    y = geo.splitmix64_vector(nloc, 1234, offset=r0)
    x = np.zeros(nloc)

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def gather(v):
        parts = [None] * world
        dist.all_gather_object(parts, v)
        return np.concatenate(parts)

    y /= np.sqrt(allsum(y @ y))
    a, b = [], []
    for j in range(12):
        yfull = gather(y)
        assert yfull.size == rows
        o.matvec_range(x, yfull, r0, r0 + nloc, faithful=False)
        aj = allsum(y @ x)
        x -= aj * y
        bj = np.sqrt(allsum(x @ x))
        y, x = x / bj, -bj * y
        a.append(aj)
        b.append(bj)
    if rank == 0:
        out.put((a, b))
    dist.destroy_process_group()


def test_sharded_recurrence_matches_single_process(oracle, lpp):
    from lanczosplusplus_b200 import geometry as geo
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    a, b = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = cases.hubbard_chain(6, 3, 2)
    o = cases.make_oracle(oracle, case, fast_rank=1)
    a0, b0 = o.decomposition(geo.splitmix64_vector(o.rows(), 1234), steps=12, eps=0.0)
    assert np.abs(np.array(a) - a0).max() < 1e-11 and np.abs(np.array(b) - b0).max() < 1e-11


def test_local_rows_cover_basis(lpp):
    from lanczosplusplus_b200 import distributed as D
    for n1, n2, w in ((12870, 12870, 8), (8008, 8008, 4), (20, 15, 2), (3, 2, 4)):
        nxt = 0
        for r in range(w):
            f, c = D.local_rows(n1, n2, r, w)
            assert f == nxt and c % n1 == 0
            nxt += c
        assert nxt == n1 * n2
