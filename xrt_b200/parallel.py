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
    """The communicator of one rank.  Transport: torch.distributed (NCCL on GPUs, gloo in the CPU tests); with
    `use_capi=True` the collectives of the ray / wave path go through the library's own C-ABI NCCL entry points
    (include/xrt_b200.h: xrt_allreduce_slots, xrt_allgather_beam, xrt_allgather_bytes) on a communicator made by
    xrt_nccl_comm_init -- what a host that is not Python uses."""

    def __init__(self, world=1, rank=0, use_capi=False):
        self.world = world
        self.rank = rank
        self._nccl = None
        self._scratch = None
        if use_capi and world > 1:
            self.init_capi()

    # ------------------------------------------------------------------ C-ABI NCCL communicator
    def init_capi(self):
        """Create an NCCL communicator through the C ABI; the 128-byte unique id travels over torch.distributed."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        idbuf = (ctypes.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(lib.xrt_nccl_unique_id(ctypes.addressof(idbuf)), 'xrt_nccl_unique_id')
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device='cuda')
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
        comm = ctypes.c_void_p()
        _lib.check(lib.xrt_nccl_comm_init(self.world, self.rank, ctypes.addressof(idbuf), ctypes.byref(comm)),
                   'xrt_nccl_comm_init')
        self._nccl = comm
        self._scratch = torch.empty(self.world * SLOTS, dtype=torch.int64, device='cuda')
        return self

    def close(self):
        if self._nccl is not None:
            from . import _lib
            _lib.load().xrt_nccl_comm_destroy(self._nccl)
            self._nccl = None

    @staticmethod
    def _stream():
        import ctypes
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def shard(self, n):
        """Contiguous index range of this rank for n items (equal chunks, last rank takes the remainder)."""
        per = n // self.world
        start = self.rank * per
        stop = n if self.rank == self.world - 1 else start + per
        return start, stop

    # ------------------------------------------------------------------ beam-wide scalars
    def allreduce_reductions(self, red, sums=True):
        """In-place all-reduce of a reduction-slot buffer (int64[64]) over all ranks."""
        if self.world == 1:
            return red
        if self._nccl is not None and red.is_cuda:
            from . import _lib
            _lib.check(_lib.load().xrt_allreduce_slots(self._nccl, self.world, red.data_ptr(), int(bool(sums)),
                                                       self._scratch.data_ptr(), self._stream()),
                       'xrt_allreduce_slots')
            return red
        gathered = torch.empty(self.world * SLOTS, dtype=torch.int64, device=red.device)
        dist.all_gather_into_tensor(gathered, red)
        red.copy_(combine_slots(gathered.view(self.world, SLOTS), sums, red.clone()))
        return red

    def allreduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def allreduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def broadcast(self, t, src=0):
        if self.world > 1:
            dist.broadcast(t, src)
        return t

    # ------------------------------------------------------------------ ragged all-gathers
    def lengths(self, n, device):
        """Shard lengths of all ranks (rays per rank differ: shard() gives the last rank the remainder, rejection
        sampling and filters change them further)."""
        if self.world == 1:
            return [int(n)]
        t = torch.zeros(self.world, dtype=torch.int64, device=device)
        mine = torch.tensor([int(n)], dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(t, mine)
        return [int(v) for v in t.tolist()]

    def allgather_rows(self, t, counts=None):
        """Concatenation over ranks, in rank order, of 2-D (rows, C) or 1-D tensors with differing row counts: ONE
        collective on a buffer padded to the longest shard (torch transport) or one grouped NCCL operation (C ABI)."""
        if self.world == 1:
            return t
        t = t.contiguous()
        if counts is None:
            counts = self.lengths(t.shape[0], t.device)
        row = t[0:1].numel() * t.element_size() if t.shape[0] else int(np.prod(t.shape[1:])) * t.element_size()
        total = sum(counts)
        out = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        if self._nccl is not None and t.is_cuda:
            import ctypes
            from . import _lib
            cnt = (ctypes.c_uint64 * self.world)(*[c * row for c in counts])
            _lib.check(_lib.load().xrt_allgather_bytes(self._nccl, self.world, self.rank, ctypes.addressof(cnt),
                                                       t.data_ptr(), out.data_ptr(), self._stream()),
                       'xrt_allgather_bytes')
            return out
        mx = max(counts)
        if min(counts) == mx:
            dist.all_gather_into_tensor(out, t)
            return out
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        buf = torch.empty((self.world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad)
        o = 0
        for r, c in enumerate(counts):
            out[o:o + c] = buf[r * mx:r * mx + c]
            o += c
        return out

    def allgather_beam(self, beam):
        """The single all-gather of the Beam SoA at a Screen named in the north star: every rank ends up with the
        rays of all ranks, in rank order.  Shards may differ in length.  One collective: the C-ABI path groups the
        arrays in one NCCL operation (xrt_allgather_beam); the torch path packs them into one byte buffer."""
        if self.world == 1:
            return beam
        from .beams import Beam
        names = [k for k, v in beam.__dict__.items() if isinstance(v, torch.Tensor)]
        n = len(beam)
        counts = self.lengths(n, beam.device)
        total = sum(counts)
        out = Beam.__new__(Beam)
        for k, v in beam.__dict__.items():
            if not isinstance(v, torch.Tensor):
                setattr(out, k, v)
        if self._nccl is not None and beam.x.is_cuda:
            import ctypes
            from . import _lib
            for k in names:
                v = getattr(beam, k)
                setattr(out, k, torch.empty((total,), dtype=v.dtype, device=v.device))
            cin, cout = beam.c_struct(), out.c_struct()
            cnt = (ctypes.c_uint64 * self.world)(*counts)
            _lib.check(_lib.load().xrt_allgather_beam(self._nccl, self.world, self.rank, ctypes.addressof(cnt),
                                                      ctypes.byref(cin), ctypes.byref(cout), self._stream()),
                       'xrt_allgather_beam')
            return out
        # torch transport: one byte buffer [array 0 | array 1 | ...], each array padded to the longest shard
        mx = max(counts)
        sizes = [getattr(beam, k).element_size() for k in names]
        offs = np.concatenate([[0], np.cumsum([s * mx for s in sizes])]).astype(np.int64)
        send = torch.zeros(int(offs[-1]), dtype=torch.uint8, device=beam.device)
        for k, s, o in zip(names, sizes, offs[:-1]):
            v = getattr(beam, k).contiguous()
            send[int(o):int(o) + s * n] = v.view(torch.uint8) if not v.is_complex() else \
                torch.view_as_real(v).contiguous().view(torch.uint8).reshape(-1)
        recv = torch.empty(self.world * int(offs[-1]), dtype=torch.uint8, device=beam.device)
        dist.all_gather_into_tensor(recv, send)
        recv = recv.view(self.world, int(offs[-1]))
        for k, s, o in zip(names, sizes, offs[:-1]):
            v = getattr(beam, k)
            parts = [recv[r, int(o):int(o) + s * counts[r]] for r in range(self.world)]
            flat = torch.cat(parts).contiguous()
            if v.is_complex():
                g = torch.view_as_complex(flat.view(torch.float64).view(-1, 2))
            else:
                g = flat.view(v.dtype)
            setattr(out, k, g)
        return out

    def allgather_host(self, arr, device=None):
        """numpy array slices of all ranks -> the whole array on every rank (first axis; ragged), through device
        memory when the transport is NCCL.  Used by XRT_CUDA.run_parallel_max."""
        if self.world == 1:
            return arr
        a = np.ascontiguousarray(arr)
        cplx = np.iscomplexobj(a)
        t = torch.from_numpy(a.view(np.float64).reshape(a.shape[0], -1) if cplx else a.reshape(a.shape[0], -1))
        if device is None:
            device = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
        g = self.allgather_rows(t.to(device)).cpu().numpy()
        if cplx:
            g = g.view(np.complex128)
        return g.reshape((-1,) + a.shape[1:])
