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
    # shards + beam all-gather: the real shard() of 1001 rays over 2 ranks is ragged (500 + 501)
    n = 1001
    start, stop = comm.shard(n)
    b = Beam.__new__(Beam)
    b.x = torch.arange(start, stop, dtype=torch.float64)
    b.state = torch.arange(start, stop, dtype=torch.int32)
    b.Jsp = torch.arange(start, stop, dtype=torch.float64).to(torch.complex128) * (1 + 2j)
    b.parentId = 'p'
    g = comm.allgather_beam(b)
    ok &= bool(torch.equal(g.x, torch.arange(0, n, dtype=torch.float64)))
    ok &= bool(torch.equal(g.state, torch.arange(0, n, dtype=torch.int32)))
    ok &= bool(torch.equal(g.Jsp, torch.arange(0, n, dtype=torch.float64).to(torch.complex128) * (1 + 2j)))
    ok &= g.parentId == 'p' and len(g.x) == n
    # equal shards take the un-padded path
    per = n // world
    eq = Beam.__new__(Beam)
    eq.x = torch.arange(rank * per, (rank + 1) * per, dtype=torch.float64)
    eq.state = torch.arange(rank * per, (rank + 1) * per, dtype=torch.int32)
    eq.parentId = 'p'
    g = comm.allgather_beam(eq)
    ok &= bool(torch.equal(g.x, torch.arange(0, per * world, dtype=torch.float64)))
    ok &= bool(torch.equal(g.state, torch.arange(0, per * world, dtype=torch.int32)))
    # ragged rows (the packed Kirchhoff records travel like this) and host arrays (run_parallel_max)
    rows = torch.arange(start, stop, dtype=torch.float64)[:, None] * torch.ones(1, 16, dtype=torch.float64)
    gr = comm.allgather_rows(rows)
    ok &= bool(torch.equal(gr[:, 3], torch.arange(0, n, dtype=torch.float64))) and tuple(gr.shape) == (n, 16)
    hc = (np.arange(start, stop) * (1 - 1j)).astype(np.complex128)
    gh = comm.allgather_host(hc, device='cpu')
    ok &= bool(np.array_equal(gh, np.arange(n) * (1 - 1j))) and gh.dtype == np.complex128
    # XRT_CUDA.run_parallel_max slices like myopencl.py:478-491 and returns whole arrays (fake kernel: no GPU here)
    from xrt_b200.cuda_backend import XRT_CUDA
    cl = XRT_CUDA.__new__(XRT_CUDA)
    cl.lastTargetOpenCL = 'auto'
    cl.comm = comm
    seen = {}

    def _k_fake(scalarArgs, slicedRO, nonSlicedRO, slicedRW, nonSlicedRW, dimension):
        seen['dim'] = dimension
        seen['ro0'] = slicedRO[0][0]
        slicedRW[0][:] = slicedRO[0] * 2 + nonSlicedRO[0][0]
        slicedRW[1][:] = (slicedRO[1][0::2] + 1j * slicedRO[1][1::2])
        return tuple(slicedRW)
    cl._k_fake = _k_fake
    xin = np.arange(n, dtype=np.float64)
    pairs = np.arange(2 * n, dtype=np.float64)  # secondDim = 2 (myopencl.py:501-503)
    o1 = np.zeros(n)
    o2 = np.zeros(n, dtype=np.complex128)
    r1, r2 = cl.run_parallel_max('fake', [np.int32(3)], [xin, pairs], [np.array([10.])], [o1, o2], None, n)
    ok &= r1 is o1 and r2 is o2 and seen['dim'] == stop - start and seen['ro0'] == start
    ok &= bool(np.array_equal(o1, xin * 2 + 10)) and bool(np.array_equal(o2, pairs[0::2] + 1j * pairs[1::2]))
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
