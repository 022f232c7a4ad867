# -*- coding: utf-8 -*-
"""Multi-GPU plumbing: one process per GPU (torchrun), rays / receiving points sharded by contiguous index
ranges -- the slicing of XRT_CL.run_parallel_max (xrt/backends/raycing/myopencl.py:478-491) -- and the two
exchange steps the path really has:

  * the beam-wide scalars numpy takes with whole-array min/max (dominant direction, global t bracket,
    secant-vs-Brent choice, first good ray, footprint, Imax): an all-reduce of the 64 reduction slots the
    reflect pre-pass fills (order-preserving uint64 encodings: min/max commute with the encoding);
  * optionally the all-gather of the Beam SoA at a Screen.

torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport; no data-path collective is needed
anywhere else because rays are independent."""
import numpy as np
import torch
import torch.distributed as dist

SLOTS = 64
_SIGN = -(1 << 63)


def slot_kinds():
    """0 = max, 1 = min, 2 = sum for each reduction slot (layout: xrt_b200/csrc/reflect.cu)."""
    kinds = np.zeros(SLOTS, dtype=np.int64)
    kinds[3] = 1                      # first good ray key
    for k in range(4, 28):
        if ((k - 4) & 3) == 0:
            kinds[k] = 1              # min t1 of each (axis, sign) variant
    kinds[28:31] = 1                  # footprint min
    kinds[31:34] = 0                  # footprint max
    kinds[34:] = 2                    # counters
    kinds[40] = 1                     # multiple_reflect, derivative stage: min t1 ...
    kinds[41:44] = 0                  # ... max t2, max |dz1|, max |dz2|
    return kinds


def encode(d):
    """double -> order-preserving uint64 (as int64 bit pattern); numpy mirror of enc() in reflect.cu."""
    u = np.asarray(d, dtype=np.float64).view(np.uint64)
    neg = (u >> np.uint64(63)).astype(bool)
    out = np.where(neg, ~u, u | np.uint64(1 << 63))
    return out.view(np.int64)


def decode(s):
    u = np.asarray(s, dtype=np.int64).view(np.uint64)
    pos = (u >> np.uint64(63)).astype(bool)
    out = np.where(pos, u & np.uint64((1 << 63) - 1), ~u)
    return out.view(np.float64)


def combine_slots(stacked, sums=True, own=None):
    """stacked: (world, 64) int64 tensor of slot buffers -> (64,) combined buffer.  With sums=False the
    counter slots keep this rank's values `own` (they were already summed by an earlier all-reduce)."""
    kinds = torch.from_numpy(slot_kinds()).to(stacked.device)
    s = stacked ^ _SIGN  # unsigned order -> signed order
    mx = s.max(dim=0).values ^ _SIGN
    mn = s.min(dim=0).values ^ _SIGN
    sm = stacked.sum(dim=0) if sums or own is None else own
    return torch.where(kinds == 0, mx, torch.where(kinds == 1, mn, sm))


class Comm(object):
    def __init__(self, world=1, rank=0):
        self.world = world
        self.rank = rank

    def shard(self, n):
        """Contiguous index range of this rank for n items (equal chunks, last rank takes the remainder)."""
        per = n // self.world
        start = self.rank * per
        stop = n if self.rank == self.world - 1 else start + per
        return start, stop

    def allreduce_reductions(self, red, sums=True):
        """In-place all-reduce of a reduction-slot buffer (int64[64]) over all ranks."""
        if self.world == 1:
            return red
        gathered = torch.empty(self.world * SLOTS, dtype=torch.int64, device=red.device)
        dist.all_gather_into_tensor(gathered, red)
        red.copy_(combine_slots(gathered.view(self.world, SLOTS), sums, red.clone()))
        return red

    def allreduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def allgather_beam(self, beam):
        """The single all-gather of the Beam SoA at a Screen named in the north star: every rank ends up with
        the rays of all ranks, in rank order."""
        if self.world == 1:
            return beam
        from .beams import Beam
        out = Beam.__new__(Beam)
        for k, v in beam.__dict__.items():
            if isinstance(v, torch.Tensor):
                g = torch.empty((self.world * v.numel(),), dtype=v.dtype, device=v.device)
                dist.all_gather_into_tensor(g, v.contiguous())
                setattr(out, k, g)
            else:
                setattr(out, k, v)
        return out
