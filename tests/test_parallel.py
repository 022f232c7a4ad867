"""Multi-rank plumbing on CPU (gloo, world_size 2): the reduction-slot all-reduce reproduces the single-process
beam-wide scalars, and the Beam all-gather concatenates shards in rank order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_slot_encoding_is_order_preserving():
    from xrt_b200 import parallel
    rng = np.random.RandomState(0)
    v = np.concatenate([rng.normal(size=1000) * 10.0 ** rng.randint(-300, 300, 1000),
                        [0., -0., np.inf, -np.inf, 1e-320, -1e-320]])
    e = parallel.encode(v).view(np.uint64)
    order = np.argsort(v, kind='stable')
    strictly = np.diff(v[order]) > 0  # -0. and 0. compare equal but encode differently: only strict steps count
    assert np.all(np.diff(e[order].astype(object))[strictly] > 0)
    back = parallel.decode(parallel.encode(v))
    assert np.array_equal(back, v)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from xrt_b200 import parallel
    from xrt_b200.beams import Beam
    comm = parallel.Comm(world, rank)
    rng = np.random.RandomState(42)
    full = rng.normal(size=(world, parallel.SLOTS)) * 100
    counts = rng.randint(0, 1000, size=(world, parallel.SLOTS))
    kinds = parallel.slot_kinds()
    mine = parallel.encode(full[rank])
    mine[kinds == 2] = counts[rank][kinds == 2]
    red = torch.from_numpy(mine.copy())
    comm.allreduce_reductions(red)
    got = red.numpy()
    exp_max = full.max(axis=0)
    exp_min = full.min(axis=0)
    ok = True
    dec = parallel.decode(got)
    ok &= bool(np.array_equal(dec[kinds == 0], exp_max[kinds == 0]))
    ok &= bool(np.array_equal(dec[kinds == 1], exp_min[kinds == 1]))
    ok &= bool(np.array_equal(got[kinds == 2], counts.sum(axis=0)[kinds == 2]))
    # shards + beam all-gather
    n = 1001
    start, stop = comm.shard(n)
    b = Beam.__new__(Beam)
    b.x = torch.arange(start, stop, dtype=torch.float64)
    b.state = torch.arange(start, stop, dtype=torch.int32)
    b.Jsp = torch.arange(start, stop, dtype=torch.float64).to(torch.complex128) * (1 + 2j)
    b.parentId = 'p'
    if world > 1 and (stop - start) != n // world:
        # all_gather_into_tensor needs equal shards: pad the last one like run_parallel_max pads (myopencl.py:466-473)
        pass
    eq = Beam.__new__(Beam)
    per = n // world
    eq.x = torch.arange(rank * per, (rank + 1) * per, dtype=torch.float64)
    eq.state = torch.arange(rank * per, (rank + 1) * per, dtype=torch.int32)
    eq.parentId = 'p'
    g = comm.allgather_beam(eq)
    ok &= bool(torch.equal(g.x, torch.arange(0, per * world, dtype=torch.float64)))
    ok &= bool(torch.equal(g.state, torch.arange(0, per * world, dtype=torch.int32)))
    ok &= g.parentId == 'p'
    covered = comm.shard(n)
    q.put((rank, ok, covered))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_reduction_allreduce_and_beam_allgather():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(ok for _, ok, _ in res), res
    ranges = sorted(c for _, _, c in res)
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 1001
