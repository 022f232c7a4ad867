"""torchrun --nproc-per-node 2 tools/gpu_sharded_check.py: the sharded product paths on real GPUs against the
single-GPU results -- waves.diffract(comm=, samplesSharded=True) with both transports (torch.distributed, C-ABI NCCL),
Comm.allgather_beam (ragged), Comm.allreduce_reductions.  Prints one JSON line on rank 0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from xrt_b200 import _lib, parallel, waves, synthetic  # noqa: E402
from xrt_b200.beamline import BeamLine  # noqa: E402
from xrt_b200.beams import Beam  # noqa: E402
from xrt_b200.planes import Screen, RectangularAperture  # noqa: E402


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    lr = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    _lib.init(lr)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    res = {}
    bl = BeamLine()
    slit = RectangularAperture(bl, 'slit', center=(0, 1e4, 0), opening=[-0.05, 0.05, -0.05, 0.05])
    scr = Screen(bl, 'scr', center=(0, 2e4, 0))
    n = 20011
    g = synthetic.geometric_beam(3, n, dx=0.02, dz=0.02, dxprime=1e-6, dzprime=1e-6, energies=(9000., 9000.), kill=0.1,
                                 center=(0., 0., 0.))
    # a local beam ON the slit plane (y = 0 locally): what slit.propagate returns
    g['y'] = np.zeros(n)
    full = Beam.from_numpy(g)
    xm = np.linspace(-0.3, 0.3, 96)
    zm = np.linspace(-0.3, 0.3, 64)
    # single-GPU reference: the whole wave from all samples
    wave0 = scr.prepare_wave(slit, xm, zm)
    lb0 = Beam(copyFrom=full)
    lb0.area = 0
    waves.diffract(lb0, wave0)
    ref = {k: getattr(wave0, k).clone() for k in ('Es', 'Ep', 'Jss', 'a', 'b', 'c')}
    for capi in (False, True):
        comm = parallel.Comm(world, rank, use_capi=capi)
        zlo, zhi = comm.shard(len(zm))
        s0, s1 = comm.shard(n)
        lb = Beam.__new__(Beam)
        for k, t in full.tensors().items():
            setattr(lb, k, t[s0:s1].contiguous())
        lb.parentId = None
        lb.area = 0
        wave = scr.prepare_wave(slit, xm, zm[zlo:zhi])
        waves.diffract(lb, wave, comm=comm, samplesSharded=True)
        sl = slice(zlo * len(xm), zhi * len(xm))
        errs = {}
        for k, v in ref.items():
            scale = float(v.abs().max())
            errs[k] = float((getattr(wave, k) - v[sl]).abs().max()) / scale
        tag = 'capi' if capi else 'torch'
        res['diffract_' + tag] = errs
        # ragged beam all-gather
        gb = comm.allgather_beam(lb)
        ok = all(bool(torch.equal(getattr(gb, k), t)) for k, t in full.tensors().items())
        res['allgather_beam_' + tag] = ok and len(gb) == n
        # slot all-reduce against the numpy combination
        rng = np.random.RandomState(5)
        vals = rng.normal(size=(world, parallel.SLOTS)) * 100
        cnts = rng.randint(0, 1000, size=(world, parallel.SLOTS))
        kinds = parallel.slot_kinds()
        mine = parallel.encode(vals[rank])
        mine[kinds == 2] = cnts[rank][kinds == 2]
        red = torch.from_numpy(mine.copy()).cuda()
        comm.allreduce_reductions(red)
        got = red.cpu().numpy()
        dec = parallel.decode(got)
        res['allreduce_slots_' + tag] = bool(np.array_equal(dec[kinds == 0], vals.max(0)[kinds == 0]) and
                                             np.array_equal(dec[kinds == 1], vals.min(0)[kinds == 1]) and
                                             np.array_equal(got[kinds == 2], cnts.sum(0)[kinds == 2]))
        torch.cuda.synchronize()
        comm.close()
    # every rank must agree
    flags = torch.tensor([int(all(v if isinstance(v, bool) else max(v.values()) < 1e-9 for v in res.values()))],
                         device='cuda')
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    sys.stdout.flush()
    os.dup2(saved, 1)
    if rank == 0:
        res['all_ranks_ok'] = bool(flags.item())
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
