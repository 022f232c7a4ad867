"""GPU-side: wrap-around sums of the bit patterns of every array Undulator.shine returns (fused uniformRayDensity
kernel and the three-kernel rejection path), for comparing two builds of the library bit for bit:

    [XRT_B200_LIB=variants/TAG/libxrt_b200.so] python tools/gpu_shine_hash.py out.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from xrt_b200 import _lib  # noqa: E402
from xrt_b200.beamline import BeamLine  # noqa: E402
from xrt_b200.beams import Beam  # noqa: E402
from xrt_b200.sources import Undulator  # noqa: E402


def bits(t):
    if t.dtype == torch.complex128:
        t = torch.view_as_real(t)
    if t.dtype == torch.float64:
        return t.contiguous().view(torch.int64)
    return t.to(torch.int64)


def main():
    _lib.init(0)
    torch.cuda.set_device(0)
    n = 1000000
    out = {}
    for tag, kw in (('uniform', {}), ('uniform_espread', dict(eEspread=8e-4)), ('rejection', dict(uniformRayDensity=False))):
        k = bench.undulator_kwargs(n)
        k.update(kw)
        src = Undulator(BeamLine(), 'u', eMin=bench.E1 * 0.995, eMax=bench.E1 * 1.005, seed=4321, **k)
        src.reset()
        if k.get('uniformRayDensity', True):
            beam = Beam(n, withAmplitudes=True, uninitialized=True)
            src.shine_uniform_fused(beam, index_offset=0)
        else:
            src.nrays = 200000
            beam = src.shine()
        torch.cuda.synchronize()
        out[tag] = {f: int(bits(v).sum().item()) for f, v in beam.tensors().items()}
        out[tag]['n'] = len(beam)
    with open(sys.argv[1], 'w') as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(json.dumps(out)[:300])


if __name__ == '__main__':
    main()
