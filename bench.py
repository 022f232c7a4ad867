#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Benchmark of the xrt raycing hot path on B200 (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nrays R] [--no-kirchhoff]

Workload (BASELINE.json configs[1], SURVEY.md 8d "Config 2", sub-mode 2u): Undulator (uniformRayDensity) ->
ToroidMirror (Au) -> Screen, 1e8 rays per GPU, fp64.  One step = one pass of shine + reflect + expose over a
fresh batch of rays.  `value` = rays/s with everything resident in HBM; `e2e` = the same pipeline called
through the public API with the resulting screen Beam copied to pinned host memory inside the timed region.
The Kirchhoff integral of config 4 (N = 1e6 samples -> 1024 x 1024 screen) is timed in the same run and
reported under "kirchhoff" with its own roofline.

Multi-GPU (torchrun, one rank per GPU): rays are sharded (weak scaling, `--nrays` per rank); the beam-wide
scalars of reflect() and Imax are all-reduced (min/max over 64 slots); `gather_beam` reports the same step with
the all-gather of the screen Beam named in the north star as a second timed line; the Kirchhoff sum shards
its receiving points AND its samples (packed records all-gathered); `config5` is BASELINE configs[4].
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'Mrays/s through 3-OE beamline; Kirchhoff GFLOP/s vs roofline at 1/2/4/8 B200'
WORKLOAD = 'config2u: Undulator(uniformRayDensity, 64 nodes) -> ToroidMirror(Au) -> Screen'
FLOP_PER_SAMPLE_NODE = 63   # 60 + sincos (2) + reciprocal (1)
FLOP_PER_PAIR = 57


def undulator_kwargs(nrays):
    """tests/speed/2_synchrotronSources_speed.py:136-155 base parameters, first harmonic +-0.5 %."""
    return dict(nrays=nrays, eE=3., eI=0.5, period=30., n=40, K=1.45, eSigmaX=48.65, eSigmaZ=6.197,
                eEpsilonX=0.263, eEpsilonZ=0.008, xPrimeMax=0.1, zPrimeMax=0.1, xPrimeMaxAutoReduce=False,
                zPrimeMaxAutoReduce=False, gNodes=32, gIntervals=2, uniformRayDensity=True)


E1 = 1388.8614066632863  # first harmonic of that undulator [eV] (Undulator.report_E1)
MIRROR = dict(center=(0, 25000, 0), pitch=3e-3, R=25000. * 2 / np.sin(3e-3) / 2, r=25000. * np.sin(3e-3),
              limPhysX=[-10, 10], limPhysY=[-500, 500])
SCREEN_Y = 50000.


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        reasons = []
        for i, name in enumerate(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')):
            if any(len(s) > 3 + i and s[3 + i].lower().startswith('active') for s in self.samples):
                reasons.append(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.samples)}


def measured_traffic(kernel_prefix, nrays):
    """dram bytes per launch of a ray kernel, from the committed ncu --set full capture (bytes/ray x rays)."""
    path = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if not os.path.exists(path):
        path = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
    try:
        with open(path) as f:
            t = json.load(f)
        for k, v in t['bytes_per_ray'].items():
            if k.startswith(kernel_prefix):
                return v * nrays
    except Exception:
        pass
    return None


def _finite(o):
    """NaN/inf -> None so that the line is strict JSON."""
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    if isinstance(o, float) and not np.isfinite(o):
        return None
    return o


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return {'hbm_gbs': 6650.0}, 'fallback'


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_pipeline(nrays, seed):
    """The reference's numpy path for the same workload, as restated in oracle/ (one process, one thread)."""
    from oracle import shine as osh, reflect as orf, planes as op, materials as om, undulator as ou
    gamma = ou.gamma_from_eE(3.)
    dx, dz = 48.65e-3, 6.197e-3
    dxp, dzp = 0.263e-6 / dx, 0.008e-6 / dz
    src = osh.UndulatorSpec(L0=30., Np=40., Kx=0., Ky=1.45, eI=0.5, gamma=gamma, phase=0., quadm=32, gIntervals=2,
                            E_min=E1 * 0.995, E_max=E1 * 1.005, Theta_min=-1e-4 - dxp, Theta_max=1e-4 + dxp,
                            Psi_min=-1e-4 - dzp, Psi_max=1e-4 + dzp, dx=dx, dz=dz, dxprime=dxp, dzprime=dzp,
                            uniformRayDensity=True)
    oe = orf.OESpec(center=list(MIRROR['center']), pitch=MIRROR['pitch'], surface='toroid',
                    surfParams=dict(R=MIRROR['R'], r=MIRROR['r']), surfPhysX=MIRROR['limPhysX'],
                    surfPhysY=MIRROR['limPhysY'], material=om.Material('Au', rho=19.3), lostNum=-1)
    np.random.seed(seed)
    t0 = time.perf_counter()
    beam = osh.shine(src, nrays)
    gb, lb = orf.reflect(oe, beam)
    op.expose(gb, [0., SCREEN_Y, 25000. * np.tan(2 * 3e-3)], op.xyz_from_xz(0., 1.), -2001)
    return time.perf_counter() - t0


def _cpu_worker(args):
    return cpu_pipeline(*args)


def cpu_baseline(sample_rays=100000):
    cpu_pipeline(2000, 1)  # warm-up (imports, table loads)
    t = cpu_pipeline(sample_rays, 2)
    return {'value': sample_rays / t / 1e6, 'unit': 'Mrays/s', 'cores': 1, 'kind': 'port',
            'sample': '%d rays of the same workload through oracle/ (numpy restatement of the reference path), '
                      '1 process, 1 thread' % sample_rays}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is Python and
    cannot travel to the GPU box) on all host cores, xrt's own scaling model = independent processes
    (xrt/multipro.py:376-387)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1))
    per_proc = 50000
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        for w in range(max(1, min(args.warmup, 1))):
            pool.map(_cpu_worker, [(2000, 100 + i) for i in range(cores)])
        times = []
        for k in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(per_proc, 1000 * k + i) for i in range(cores)])
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = args.steps * cores * per_proc / total / 1e6
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'Mrays/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'rays_per_step': cores * per_proc},
            'cpu_baseline': {'value': value, 'unit': 'Mrays/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d processes x %d rays per step of the same workload through oracle/'
                                       % (cores, per_proc)},
            'e2e': {'value': value, 'unit': 'Mrays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def count_own_launches(fn):
    """How many kernels of libxrt_b200.so one call of fn() launches, MEASURED with the CUPTI-backed torch profiler
    (every kernel of the process is recorded, torch's own included): kernels whose names carry the library's
    namespace / kernel names are ours.  Returns (ours, others) or None when the profiler is unavailable."""
    try:
        import torch
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        ours = others = 0
        for e in prof.events():
            if getattr(e, 'device_type', None) is None or 'cuda' not in str(e.device_type).lower():
                continue
            name = e.name
            if name.startswith('Memcpy') or name.startswith('Memset'):
                continue
            if 'xrtb::' in name or name.startswith('xrtb') or '_kernel' in name and 'at::' not in name:
                ours += 1
            else:
                others += 1
        return ours, others
    except Exception:
        return None


def config5_geometry():
    """BASELINE.json configs[4] (SURVEY 8d 'Config 5'): Undulator tuned to its 7th harmonic -> DCM Si(111) ->
    vertical + horizontal elliptical-cylinder KB pair (template tests/speed/3_Softi_CXIw2D_speed.py:232-243) -> Screen;
    the same geometry as tests/test_pipeline.py, where the chain is parity-tested against the numpy oracle."""
    from xrt_b200.materials import CrystalSi
    si = CrystalSi(hkl=(1, 1, 1))
    E0 = 7 * E1
    Ea = np.array([E0])
    theta = float(si.get_Bragg_angle(Ea)[0] - si.get_dtheta(Ea)[0])
    return dict(E0=E0, theta=theta, si=si, dcm_y=20000., vfm_y=30000., hfm_y=30600., scr_y=31800., pitch=3.5e-3,
                dcm_offset=2 * 10. * np.cos(theta))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from xrt_b200 import _lib, kernels, synthetic
    from xrt_b200.beamline import BeamLine
    from xrt_b200.beams import Beam
    from xrt_b200.sources import Undulator
    from xrt_b200.oes import ToroidMirror
    from xrt_b200.materials import Material
    from xrt_b200.planes import Screen
    from xrt_b200 import parallel

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: xrt_b200 has no CPU path')
    torch.cuda.set_device(local_rank)
    try:  # run next to the GPU: pinned result buffers then come from the NUMA node the GPU hangs on
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    _lib.init(local_rank)
    if world > 1:
        # NCCL writes its version banner (NCCL_DEBUG=VERSION / WARN) and any warnings to stdout, which carries exactly
        # one JSON line here: communicator set-up runs with fd 1 pointing at stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            import datetime
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank),
                                    timeout=datetime.timedelta(seconds=300))  # a hang fails fast instead of after 10 min
            dist.barrier()
            torch.cuda.synchronize()
            # the exchange steps of the ray / wave path go through the library's own C-ABI NCCL entry points
            comm = parallel.Comm(world, rank, use_capi=not args.torch_collectives)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    else:
        comm = parallel.Comm(world, rank)

    nrays = args.nrays
    bl = BeamLine()
    src = Undulator(bl, 'u', eMin=E1 * 0.995, eMax=E1 * 1.005, seed=1234, **undulator_kwargs(nrays))
    mirror = ToroidMirror(bl, 'tm', material=Material('Au', rho=19.3), **MIRROR)
    screen = Screen(bl, 'scr', center=(0, SCREEN_Y, 25000. * np.tan(2 * 3e-3)))
    src.reset()
    nodes = len(src.tg)
    ar = comm.allreduce_reductions if world > 1 else None

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    copy_stream = torch.cuda.Stream()
    host_out = None
    e2e_count = [0]

    def step(e2e=False, record=None, fused_chain=True):
        """One pass of the hot path over a fresh batch: Undulator.shine -> ToroidMirror.reflect -> Screen.expose.
        fused_chain: the reflect kernel continues with the screen (OE.reflect(exposeOn=)), the global beam between
        them never reaches HBM; the local beam of the mirror is materialised in both variants."""
        marks = [ev() for _ in range(4)] if record is not None else None
        off = rank * nrays
        beam = Beam(nrays, withAmplitudes=True, uninitialized=True)
        if marks:
            marks[0].record()
        src.shine_uniform_fused(beam, index_offset=off)
        if marks:
            marks[1].record()
        if fused_chain:
            blo, lb = mirror.reflect(beam, needLocal=True, checkTables=False, index_offset=off, allreduce=ar,
                                     exposeOn=screen)
            if marks:
                marks[2].record()
                marks[3].record()
        else:
            gb, lb = mirror.reflect(beam, needLocal=True, checkTables=False, index_offset=off, allreduce=ar)
            if marks:
                marks[2].record()
            blo = screen.expose(gb)
            if marks:
                marks[3].record()
        if marks:
            record.append(marks)
        if e2e:
            # D2H of the result on a copy stream: the copy of step k overlaps the kernels of step k+1
            buf = host_out[e2e_count[0] % len(host_out)]
            e2e_count[0] += 1
            done = torch.cuda.Event()
            done.record()
            copy_stream.wait_event(done)
            with torch.cuda.stream(copy_stream):
                for k, t in blo.tensors().items():
                    buf[k].copy_(t[:buf[k].numel()], non_blocking=True)
                    t.record_stream(copy_stream)
        return blo

    def barrier():
        copy_stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(steps):
            fn()
        torch.cuda.current_stream().wait_stream(copy_stream)  # the last D2H copies belong to the timed region
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda', dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    fused_chain = not args.unfused_chain
    # ---- device-resident throughput
    for _ in range(args.warmup):
        step(fused_chain=fused_chain)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stage_events = []
    ms_total = timed(lambda: step(record=stage_events, fused_chain=fused_chain), args.steps)
    torch.cuda.synchronize()
    stage_ms = {n: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in stage_events]))
                for i, n in enumerate(('shine', 'reflect', 'expose'))}
    ms_step = ms_total / args.steps
    value = world * nrays / (ms_step * 1e-3) / 1e6
    # the other chain variant, for the comparison the fusion is judged by (few steps, untimed region of the contract)
    other_steps = max(1, min(args.steps, 3))
    step(fused_chain=not fused_chain)
    ms_other = timed(lambda: step(fused_chain=not fused_chain), other_steps) / other_steps
    # (every rank runs the profiled step: it contains the slot all-reduces)
    launches = count_own_launches(lambda: step(fused_chain=fused_chain))

    # ---- end to end (i): the whole result Beam lands in pinned host memory inside the timed region
    blo = step(fused_chain=fused_chain)
    bytes_per_ray = blo.bytes_per_ray()
    n_host = len(blo)
    host_out = []
    e2e_note = None
    nbuf = 2 if world <= 2 else 1  # 2 x 13.2 GB pinned per rank; one when many ranks share a host (211 GB at 8)
    for _ in range(nbuf):
        try:
            host_out.append({k: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                             for k, t in blo.tensors().items()})
        except RuntimeError:
            if not host_out:
                host_out.append({k: torch.empty(t.shape, dtype=t.dtype) for k, t in blo.tensors().items()})
                e2e_note = 'pinned allocation failed: pageable host buffers'
            break
    del blo
    for _ in range(max(1, args.warmup - 1)):
        step(e2e=True, fused_chain=fused_chain)
    ms_e2e = timed(lambda: step(e2e=True, fused_chain=fused_chain), args.steps) / args.steps
    e2e_value = world * nrays / (ms_e2e * 1e-3) / 1e6
    h2d = 6 * nodes * 8 + 8 * 64  # node tables + scalar structs per step
    d2h = n_host * bytes_per_ray
    # concurrent D2H bandwidth of the box with every rank copying at once: the limiter of (i) named with a number
    probe = host_out[0]['Es'] if 'Es' in host_out[0] else host_out[0]['x']
    dprobe = torch.empty(probe.shape, dtype=probe.dtype, device='cuda')
    pbytes = probe.numel() * probe.element_size()

    def d2h_probe():
        probe.copy_(dprobe, non_blocking=True)
    d2h_probe()
    ms_probe = timed(d2h_probe, 3) / 3
    d2h_gbs = {'per_gpu_GBs': pbytes / ms_probe / 1e6, 'aggregate_GBs': world * pbytes / ms_probe / 1e6,
               'bytes_per_copy': pbytes, 'ranks_copying_at_once': world}
    del dprobe, probe
    host_out = None

    # ---- end to end (ii): what raycing.get_output reads (x, z, E, Jss, Jpp of the good rays), compacted on the device
    fields = ('x', 'z', 'E', 'Jss', 'Jpp')
    pinned = {k: torch.empty(nrays, dtype=torch.float64, pin_memory=True) for k in fields}
    sel_count = [0]

    def step_compact():
        b = step(fused_chain=fused_chain)
        out = b.to_host(fields=fields, good_only=True, rayFlag=(1,), out=pinned)
        sel_count[0] = len(out['x'])
    step_compact()
    ms_comp = timed(step_compact, args.steps) / args.steps
    e2e_compact = {'value': world * nrays / (ms_comp * 1e-3) / 1e6, 'unit': 'Mrays/s', 'ms_per_step': ms_comp,
                   'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': sel_count[0] * 8 * len(fields) + 8,
                   'selected_rays': sel_count[0],
                   'note': 'Beam.to_host(fields=(x, z, E, Jss, Jpp), good_only=True): xrt_compact_rays_f64 packs the '
                           'fields raycing.get_output reads (raycing/__init__.py:170-300) of the good rays, in ray '
                           'order, on the device; only they cross PCIe (blocking call: no overlap with the next step)'}
    # ... and the same with the transfer of step k overlapping the kernels of step k + 1 (Beam.to_host_begin /
    # handle.finish(): compaction and an 8-byte count on the compute stream, the fields on a copy stream, two pinned sets)
    try:
        second = {k: torch.empty(nrays, dtype=torch.float64, pin_memory=True) for k in fields}
    except RuntimeError:  # no room for a second pinned set on this host (every rank of the box pins its own)
        second = None
    ok_all = second is not None
    if world > 1:  # the timed loop has a barrier: every rank takes the same branch
        flag = torch.tensor([1 if ok_all else 0], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all = bool(flag.item())
    if ok_all:
        pinned2 = [pinned, second]
        pend = [None]
        kc = [0]

        def step_compact_async():
            b = step(fused_chain=fused_chain)
            h = b.to_host_begin(fields=fields, rayFlag=(1,), out=pinned2[kc[0] % 2])
            kc[0] += 1
            if pend[0] is not None:
                sel_count[0] = len(pend[0].finish()['x'])  # D2H of the previous step while this step's kernels run
            pend[0] = h
        step_compact_async()
        ms_compa = timed(step_compact_async, args.steps) / args.steps
        if pend[0] is not None:
            pend[0].finish()
        e2e_compact['pipelined'] = {'value': world * nrays / (ms_compa * 1e-3) / 1e6, 'unit': 'Mrays/s',
                                    'ms_per_step': ms_compa,
                                    'note': 'Beam.to_host_begin() / handle.finish(): the compacted fields of step k cross '
                                            'PCIe on a copy stream while the kernels of step k + 1 run (two pinned '
                                            'result sets)'}
        del pinned2
    else:
        e2e_compact['pipelined'] = None
    del pinned, second

    # ---- end to end (iii) the way xrt's runner consumes a repeat: only the screen image (get_output + do_hist2d,
    # xrt/multipro.py:119-171) leaves the GPU -- device histogram, 256 x 256 doubles D2H per step
    from xrt_b200 import histogram
    blo = step(fused_chain=fused_chain)
    gd = blo.state == 1
    lims = torch.stack((blo.x[gd].min(), blo.x[gd].max(), blo.z[gd].min(), blo.z[gd].max(), blo.E[gd].min(),
                        blo.E[gd].max()))
    if world > 1:  # common plot limits on every rank
        lo = lims[0::2].clone()
        hi = lims[1::2].clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        lims[0::2], lims[1::2] = lo, hi
    lims = [float(v) for v in lims.tolist()]
    del blo, gd
    ph = histogram.PlotHistograms(histogram.Axis('x', lims[0:2], 256), histogram.Axis('z', lims[2:4], 256),
                                  histogram.Axis('E', lims[4:6], 256), fluxKind='total', rayFlag=(1,))
    hbytes = sum(t.numel() * t.element_size() for t in (ph._h2, ph._hx, ph._hy, ph._hc, ph._counters))
    host_img = torch.empty(hbytes // 8, dtype=torch.float64, pin_memory=True)

    def step_hist():
        b = step(fused_chain=fused_chain)
        for t in (ph._h2, ph._hx, ph._hy, ph._hc, ph._counters):
            t.zero_()
        ph.accumulate(b)
        if world > 1:
            ph.allreduce()
        o = 0
        for t in (ph._h2, ph._hx, ph._hy, ph._hc):
            host_img[o:o + t.numel()].copy_(t.reshape(-1), non_blocking=True)
            o += t.numel()

    step_hist()
    ms_hist = timed(step_hist, args.steps) / args.steps
    e2e_hist = {'value': world * nrays / (ms_hist * 1e-3) / 1e6, 'unit': 'Mrays/s', 'ms_per_step': ms_hist,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': hbytes,
                'note': 'same step + one XYCPlot on the device (xrt_b200.histogram.PlotHistograms: 256 x 256 image and '
                        'the x, z, E 1-D histograms, each with flux + RGB channels == get_output + do_hist1d/2d), '
                        'all-reduced over the ranks; only the histograms cross PCIe, as in xrt where histograms, not '
                        'rays, are accumulated over repeats (xrt/runner.py:526-530, xrt/multipro.py:296-340)'}
    if rank == 0:
        sampler.stop_flag.set()

    # ---- the north star's collective: one all-gather of the screen Beam over all ranks (a second timed line)
    gather_rep = None
    if world > 1:
        ng = int(min(nrays, 48e9 / (132 * world)))  # every rank receives world x ng rays: keep the result <= 48 GB
        ng -= ng % 512
        src_g = Undulator(BeamLine(), 'ug', eMin=E1 * 0.995, eMax=E1 * 1.005, seed=1234, **undulator_kwargs(ng))
        src_g.reset()

        def step_gather(gather=True):
            beam = Beam(ng, withAmplitudes=True, uninitialized=True)
            src_g.shine_uniform_fused(beam, index_offset=rank * ng)
            b, _ = mirror.reflect(beam, needLocal=True, checkTables=False, index_offset=rank * ng, allreduce=ar,
                                  exposeOn=screen)
            del beam
            return comm.allgather_beam(b) if gather else b
        g = step_gather()
        gathered_rays = len(g)
        del g
        gsteps = max(1, min(args.steps, 3))
        ms_g = timed(step_gather, gsteps) / gsteps
        step_gather(False)
        ms_ng = timed(lambda: step_gather(False), gsteps) / gsteps
        gbytes = 132 * ng * (world - 1)
        gather_rep = {'workload': 'same step at %d rays per GPU + ONE all-gather of the screen Beam (132 B/ray, 14 arrays '
                                  'in one NCCL group, %s): every rank ends with all %d rays'
                                  % (ng, 'xrt_allgather_beam of the C ABI' if not args.torch_collectives
                                     else 'torch.distributed on one packed buffer', gathered_rays),
                      'rays_per_gpu': ng, 'ms_per_step': ms_g, 'ms_per_step_without_gather': ms_ng,
                      'value': world * ng / (ms_g * 1e-3) / 1e6, 'unit': 'Mrays/s',
                      'value_without_gather': world * ng / (ms_ng * 1e-3) / 1e6,
                      'allgather_ms': ms_g - ms_ng, 'received_bytes_per_gpu': gbytes,
                      'allgather_GBs_per_gpu': gbytes / max(ms_g - ms_ng, 1e-6) / 1e6}
        del src_g

    # ---- per-kernel times of the materialising reflect() (pre-pass A, pre-pass B, main) in a few extra untimed steps
    def kernel_times(**kw):
        mirror.profileEvents = []
        for _ in range(4):
            beam = Beam(nrays, withAmplitudes=True, uninitialized=True)
            src.shine_uniform_fused(beam, index_offset=rank * nrays)
            out = mirror.reflect(beam, needLocal=True, checkTables=False, index_offset=rank * nrays, allreduce=ar, **kw)
            del out, beam
        torch.cuda.synchronize()
        evs = mirror.profileEvents[1:]
        mirror.profileEvents = None
        return [float(np.mean([e[a].elapsed_time(e[b]) for e in evs])) for a, b in ((0, 1), (4, 2), (5, 3))]
    refl_ms = kernel_times()
    fused_ms = kernel_times(exposeOn=screen)

    # ---- rooflines.  Denominators: HBM from MEASURED_PEAKS.json; FP64 has no entry there, so a DFMA stream is
    # measured in this run (xrt_b200_fp64_probe).
    peaks, peak_kind = measured_peaks()
    dfma_tflops, _ = kernels.fp64_probe(0, 8192)
    und_ms = stage_ms['shine']
    und_tflops = FLOP_PER_SAMPLE_NODE * nrays * nodes / (und_ms * 1e-3) / 1e12
    pair_path = bool(src.gridSymmetric and src.Kx == 0)
    # DFMA-pipe instructions per (sample, node) in SASS: 38 per node pair when every sample has the same gamma (the
    # gamma-only node quantities come from the node tile), 44 with an electron energy spread
    und_instr = (19.0 if not (getattr(src, 'eEspread', 0) or 0) > 0 else 22.0) if pair_path else 57.0
    roofline_und = {'kernel': 'shine_uniform_kernel<2,planar,%s> (Undulator.shine: sample + field integral + Beam '
                              'assembly)' % ('pairs' if pair_path else 'nodes'),
                    'bound': 'fp64', 'achieved': und_tflops, 'peak': dfma_tflops, 'unit': 'TFLOP/s',
                    'frac': und_tflops / dfma_tflops, 'traffic': measured_traffic('shine_uniform_kernel', nrays),
                    'issue_slot_frac': 2 * und_instr * nrays * nodes / (und_ms * 1e-3) / 1e12 / dfma_tflops,
                    'note': 'achieved counts the 63 flop per (sample, node) of the reference algorithm (SURVEY 8d); '
                            'the kernel evaluates mirror PAIRS of nodes for a planar undulator on a symmetric grid '
                            '(38 DFMA-pipe instructions per pair: 49 in r01; the sin of the pair phase through a '
                            '1024-entry shared-memory table: 44; beta_x, beta_z, beta\'_z of a node from the node tile '
                            'when gamma is the same for every sample: 38), so issue_slot_frac is the pipe utilisation',
                    'ms': und_ms, 'share_of_step': und_ms / ms_step}
    main_ms = refl_ms[2]
    main_bytes = 132 + 132 + 140
    main_gbs = main_bytes * nrays / (main_ms * 1e-3) / 1e9
    tile = os.environ.get('XRTB_REFLECT_TILE', '') != '0'
    main_name = 'reflect_tile_kernel' if tile else 'reflect_main_kernel'
    roofline_main = {'kernel': '%s<toroid, mirror> (OE.reflect, materialising gb and lb)' % main_name, 'bound': 'hbm',
                     'achieved': main_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': main_gbs / peaks['hbm_gbs'],
                     'traffic': measured_traffic(main_name + '<toroid, mirror> (gb + lb)', nrays),
                     'peak_source': '%s HBM copy bandwidth (MEASURED_PEAKS.json)' % peak_kind,
                     'note': '%d algorithmic bytes per ray (132 in, 132 gb, 140 lb); timed in extra steps through '
                             'OE.reflect() without exposeOn (the step itself runs the fused reflect + expose variant: '
                             'stages.reflect_expose_fused_main).  r02 tile structure: the 512 rays of a block live in '
                             'shared memory (bulk-async loads, results in place, bulk-async stores), rays are ranked '
                             'by class so that the root search and the Fresnel amplitudes run on dense warps (41 %% of '
                             'the rays miss the mirror), divisions / roots without the range test of the IEEE '
                             'sequences.  Bound by the dependent FP64 chain of one ray (secant search, complex Fresnel '
                             'arithmetic in numpy rounding order) at 1024 rays in flight per SM (shared memory: 210 B '
                             'per ray), not by HBM' % main_bytes,
                     'ms': main_ms, 'share_of_step': main_ms / ms_step}
    roofline = roofline_main if main_ms >= und_ms else roofline_und
    roofline_other = roofline_und if main_ms >= und_ms else roofline_main
    hbm = {}
    fused_bytes = 132 + 132 + 140  # in, screen beam, lb
    stage_list = [('reflect_prepass_dir', 28, refl_ms[0]), ('reflect_prepass_bracket', 52, refl_ms[1]),
                  ('reflect_main', main_bytes, refl_ms[2]), ('reflect_expose_fused_main', fused_bytes, fused_ms[2])]
    if not fused_chain:
        stage_list.append(('expose', 264, stage_ms['expose']))
    for name, bpr, ms in stage_list:
        gbs = bpr * nrays / (ms * 1e-3) / 1e9
        hbm[name] = {'ms': ms, 'bytes_per_ray': bpr, 'achieved_gbs': gbs,
                     'frac_of_%s_hbm' % peak_kind: gbs / peaks['hbm_gbs']}
    hbm['shine_fused (sample + undulator + fill)'] = {'ms': und_ms}
    hbm['reflect_total_in_step'] = {'ms': stage_ms['reflect']}
    chain = {'fused (OE.reflect(exposeOn=screen): xrt_reflect_expose_f64)': ms_step if fused_chain else ms_other,
             'unfused (OE.reflect + Screen.expose: the global beam goes through HBM)':
                 ms_other if fused_chain else ms_step}

    # ---- sub-mode 2r: rejection sampling instead of uniformRayDensity, source only
    rej = None
    if not args.no_rejection:
        nacc = min(nrays, 10000000)
        kw = undulator_kwargs(nacc)
        kw['uniformRayDensity'] = False
        src_r = Undulator(BeamLine(), 'ur', eMin=E1 * 0.995, eMax=E1 * 1.005, seed=4321, **kw)
        src_r.reset()
        src_r.shine()  # warm-up; also settles Imax like a converged run
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        b = src_r.shine()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rej = {'workload': 'Undulator.shine, rejection sampling, %d accepted rays' % len(b), 'ms': ms,
               'accepted_Mrays_per_s': len(b) / ms / 1e3, 'acceptance': len(b) / float(b.seeded),
               'sample_node_per_s': float(b.seeded) * nodes / (ms * 1e-3)}
        del b, src_r

    # ---- config 3: seeded GeometricSource -> DCM Si(111) -> Screen (crystal amplitude path), same ray count
    dcm_rep = None
    if not args.no_dcm:
        from xrt_b200.sources import GeometricSource
        from xrt_b200.oes import DCM
        from xrt_b200.materials import CrystalSi
        bl3 = BeamLine()
        gs = GeometricSource(bl3, 'gs', nrays=nrays, dx=0.2, dz=0.02, dxprime=5e-5, dzprime=2e-5, distE='flat',
                             energies=(8995., 9005.), polarization='h', seed=4242)
        si = CrystalSi(hkl=(1, 1, 1))
        E0 = np.array([9000.])
        theta = float(si.get_Bragg_angle(E0)[0] - si.get_dtheta(E0)[0])
        dcm = DCM(bl3, 'dcm', center=(0, 2e4, 0), bragg=theta, material=si, material2=si, cryst2perpTransl=10.,
                  limPhysX=[-50, 50], limPhysY=[-100, 100], limPhysX2=[-50, 50], limPhysY2=[-100, 200])
        scr3 = Screen(bl3, 'scr3', center=(0, 3e4, 2 * 10. * np.cos(theta)))

        def step3(fused=True):
            b = gs.shine(withAmplitudes=True, index_offset=rank * nrays)
            out = dcm.double_reflect(b, needLocal=True, checkTables=False, index_offset=rank * nrays,
                                     allreduce=ar, exposeOn=scr3 if fused else None)
            return out[0] if fused else scr3.expose(out[0])
        out3 = step3()
        good3 = float((out3.state == 1).sum()) / len(out3)
        del out3
        n3 = max(1, min(args.steps, 3))
        ms3 = timed(step3, n3) / n3
        step3(False)
        ms3u = timed(lambda: step3(False), n3) / n3
        dcm_rep = {'workload': 'config3: GeometricSource -> DCM Si(111) (two crystals, thick Bragg amplitudes) -> '
                               'Screen (fused into the second crystal\'s kernel), %d rays per GPU' % nrays,
                   'ms_per_step': ms3, 'ms_per_step_unfused': ms3u,
                   'value': world * nrays / (ms3 * 1e-3) / 1e6, 'unit': 'Mrays/s', 'good_fraction': good3}
        del gs, dcm, scr3

    # ---- Kirchhoff sum, config 4: the M receiving points are sharded over the ranks; the N samples are sharded too
    # (every rank owns N / world of them, as if they were its ray shard) and their packed records all-gathered
    kir = None
    if not args.no_kirchhoff:
        N, M = args.kirchhoff_n, args.kirchhoff_m
        Mloc = M // world
        case = synthetic.kirchhoff_case(7, N, M, mesh=(int(round(M ** 0.5)) ** 2 == M))
        sl = slice(rank * Mloc, (rank + 1) * Mloc)
        d = {k: torch.from_numpy(np.ascontiguousarray(case[k][sl])).cuda() for k in ('xM', 'yM', 'zM')}
        nlo, nhi = comm.shard(N)
        for k in ('nl', 'k', 'Es', 'Ep'):
            d[k] = torch.from_numpy(np.ascontiguousarray(case[k][nlo:nhi])).cuda()
        d['pos'] = torch.from_numpy(np.ascontiguousarray(case['pos'].T[nlo:nhi])).cuda()
        d['nrm'] = torch.from_numpy(np.ascontiguousarray(case['nrm'].T[nlo:nhi])).cuda()
        del case
        out = torch.empty((5, Mloc), dtype=torch.complex128, device='cuda')
        kev = []
        # k r <= 4.6e6 /mm x 1e4 mm = 4.6e10 rad here: far below the 3.4e12 limit of the table-based sin / cos
        ktab = float(d['k'].abs().max()) * 2e4 < kernels.PHASE_LIMIT_TABLE_SINCOS

        def kstep():
            m = [ev() for _ in range(3)]
            m[0].record()
            rec = kernels.kirchhoff_pack(d['nl'], d['Es'], d['Ep'], d['k'], d['pos'], d['nrm'])
            rec = comm.allgather_rows(rec)  # world == 1: returns rec
            m[1].record()
            kernels.kirchhoff_packed(d['xM'], d['yM'], d['zM'], rec, out=out, R=args.kirchhoff_r,
                                     uniform_normal=True, table_sincos=ktab)  # slit: n = (0, 1, 0)
            m[2].record()
            kev.append(m)
        kstep()
        ksteps = max(1, min(args.steps, 2))
        kms = timed(kstep, ksteps) / ksteps
        torch.cuda.synchronize()
        gather_ms = float(np.mean([m[0].elapsed_time(m[1]) for m in kev[1:]]))
        kernel_ms = float(np.mean([m[1].elapsed_time(m[2]) for m in kev[1:]]))
        pairs = float(N) * M / (kms * 1e-3)
        ktf = FLOP_PER_PAIR * pairs / 1e12
        ktf_kernel = FLOP_PER_PAIR * float(N) * M / (kernel_ms * 1e-3) / 1e12
        kir = {'workload': 'config4: %d samples -> %d receiving points, fp64, strict numpy rounding; receiving points '
                           'sharded over %d GPU(s), samples sharded too: packed 128-byte records all-gathered (one '
                           'collective), then summed on every rank' % (N, M, world),
               'ms': kms, 'pack_and_allgather_ms': gather_ms, 'kernel_ms': kernel_ms,
               'allgather_bytes_received_per_gpu': 128 * (N - (nhi - nlo)),
               'pairs_per_s': pairs, 'value': ktf * 1e3, 'unit': 'GFLOP/s (57 flop/pair), all GPUs',
               'roofline': {'bound': 'fp64', 'achieved': ktf_kernel / world, 'peak': dfma_tflops,
                            'unit': 'TFLOP/s per GPU', 'frac': ktf_kernel / world / dfma_tflops, 'traffic': None,
                            'issue_slot_frac': 57 * 2 * float(N) * M / (kernel_ms * 1e-3) / world / 1e12 / dfma_tflops,
                            'note': 'kirchhoff_kernel<R=4, strict, uniform normal, table sincos>: 57 DFMA-pipe '
                                    'instructions per pair in SASS (65 in r01, 63 with a 64-entry table: sin / cos through '
                                    'a 4096-entry pi/2048 table in shared memory and two-term polynomials, k folded into '
                                    'the packed Es + Ep); frac counts the 57 algorithmic flops of SURVEY 8d, issue_slot_frac the '
                                    'issued ones; DMMA shares the FP64 pipe (profiles/r02_kirchhoff_variants.json)'}}
        del d, out

    # ---- config 5: Undulator (7th harmonic) -> DCM Si(111) -> KB pair -> Screen as rays, then the wave step from the
    # last KB mirror to a 1024 x 1024 detector with the samples handed over across ranks
    c5 = None
    if not args.no_config5:
        from xrt_b200 import oes, waves
        g5 = config5_geometry()
        n5 = int(args.config5_nrays if args.config5_nrays else nrays * 1.25)
        bl5 = BeamLine()
        src5 = Undulator(bl5, 'u5', nrays=n5, eE=3., eI=0.5, period=30., n=40, K=1.45, eSigmaX=48.65, eSigmaZ=6.197,
                         eEpsilonX=0.263, eEpsilonZ=0.008, xPrimeMax=0.03, zPrimeMax=0.03, xPrimeMaxAutoReduce=False,
                         zPrimeMaxAutoReduce=False, eMin=g5['E0'] - 1., eMax=g5['E0'] + 1., gNodes=64, gIntervals=2,
                         uniformRayDensity=True, seed=555)
        src5.reset()
        au = Material('Au', rho=19.3)
        P = g5['pitch']
        dcm5 = oes.DCM(bl5, 'dcm', center=(0, g5['dcm_y'], 0), bragg=g5['theta'], material=g5['si'],
                       material2=g5['si'], cryst2perpTransl=10., limPhysX=[-20, 20], limPhysY=[-60, 60],
                       limPhysX2=[-20, 20], limPhysY2=[-60, 120])
        z1 = g5['dcm_offset']
        vfm = oes.EllipticalMirrorParam(bl5, 'vfm', center=(0, g5['vfm_y'], z1), pitch=P, p=g5['vfm_y'],
                                        q=g5['scr_y'] - g5['vfm_y'], isCylindrical=True, material=au,
                                        limPhysX=[-5, 5], limPhysY=[-150, 150])
        dyh = g5['hfm_y'] - g5['vfm_y']
        hfm = oes.EllipticalMirrorParam(bl5, 'hfm', center=(0, g5['vfm_y'] + dyh * np.cos(2 * P),
                                                          z1 + dyh * np.sin(2 * P)),
                                        pitch=P, positionRoll=np.pi / 2, p=g5['hfm_y'], q=g5['scr_y'] - g5['hfm_y'],
                                        isCylindrical=True, material=au, limPhysX=[-5, 5], limPhysY=[-150, 150],
                                        extraPitch=2 * P)
        dys = g5['scr_y'] - g5['hfm_y']
        scr5 = Screen(bl5, 'scr5', center=(0, hfm.center[1] + dys * np.cos(2 * P),
                                           hfm.center[2] + dys * np.sin(2 * P)))
        marks5 = []

        def step5(keep_local=False):
            m = [ev() for _ in range(5)]
            off5 = rank * n5
            beam = Beam(n5, withAmplitudes=True, uninitialized=True)
            m[0].record()
            src5.shine_uniform_fused(beam, index_offset=off5)
            m[1].record()
            d1 = dcm5.double_reflect(beam, needLocal=False, checkTables=False, index_offset=off5, allreduce=ar)[0]
            del beam
            m[2].record()
            d2 = vfm.reflect(d1, needLocal=False, checkTables=False, index_offset=off5, allreduce=ar)[0]
            del d1
            m[3].record()
            d3, lb3 = hfm.reflect(d2, needLocal=keep_local, checkTables=False, index_offset=off5, allreduce=ar,
                                  exposeOn=scr5)
            m[4].record()
            marks5.append(m)
            return d3, (lb3 if keep_local else None)
        # a different workload size: hand the cached blocks of the previous sections back first (the 1.0 / 2.0 GB arrays
        # of this one do not fit the 0.8 / 1.6 GB blocks cached above; the caching allocator then splits and re-mallocs
        # inside the timed steps and the GPU idles behind cudaMalloc / cudaFree -- seen as a 3x longer DCM stage)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        d3, lb3 = step5(keep_local=True)
        good5 = float((d3.state == 1).sum()) / len(d3)
        del d3
        step5()
        n5s = max(1, min(args.steps, 3))
        ast0 = torch.cuda.memory_stats()
        ms5 = timed(step5, n5s) / n5s
        torch.cuda.synchronize()
        ast1 = torch.cuda.memory_stats()
        alloc5 = {k: int(ast1.get(k, 0) - ast0.get(k, 0)) for k in ('num_alloc_retries', 'num_device_alloc',
                                                                     'num_device_free')}
        alloc5['reserved_GB'] = ast1.get('reserved_bytes.all.current', 0) / 1e9
        st5 = {nm: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in marks5[2:]]))
               for i, nm in enumerate(('shine (64 nodes)', 'dcm.double_reflect (2 crystals)', 'vfm.reflect',
                                       'hfm.reflect + screen.expose (fused)'))}
        # wave step: N5 good rays of the last mirror's local beam (this rank's share) diffract to the detector mesh
        N5 = args.config5_wave_n
        side = int(round(args.config5_wave_m ** 0.5))
        idx = torch.nonzero(lb3.state == 1).flatten()[:N5 // world]
        lbs = Beam(len(idx), withAmplitudes=True, device=lb3.device, uninitialized=True)
        for k, t in lb3.tensors().items():
            if hasattr(lbs, k):
                setattr(lbs, k, t[idx].contiguous())
        for k in ('s', 'phi', 'r', 'theta'):
            if hasattr(lb3, k):
                setattr(lbs, k, getattr(lb3, k)[idx].contiguous())
        del lb3
        zlo, zhi = comm.shard(side)
        xmesh = np.linspace(-0.05, 0.05, side)
        zmesh = np.linspace(-0.05, 0.05, side)[zlo:zhi]
        wev = []

        def wave5():
            m = [ev(), ev()]
            wave = scr5.prepare_wave(hfm, xmesh, zmesh)
            lbs.area = 0
            m[0].record()
            waves.diffract(lbs, wave, comm=comm, samplesSharded=world > 1)
            m[1].record()
            wev.append(m)
            return wave
        w = wave5()
        Mw = side * side
        Nw = int(comm.allreduce_sum(torch.tensor([len(lbs)], dtype=torch.int64, device='cuda')).item())
        del w
        msw = timed(wave5, 1)
        pairs5 = float(Nw) * Mw / (msw * 1e-3)
        c5 = {'workload': 'config5: Undulator(7th harmonic, 64 nodes) -> DCM Si(111) -> vertical + horizontal '
                          'elliptical KB -> Screen, %d rays per GPU (x %d GPUs); then wave mode: %d good rays of the '
                          'last mirror (each rank\'s share, records all-gathered) -> Kirchhoff sum to a %d x %d '
                          'detector sharded by rows' % (n5, world, Nw, side, side),
              'rays': {'ms_per_step': ms5, 'value': world * n5 / (ms5 * 1e-3) / 1e6, 'unit': 'Mrays/s',
                       'good_fraction_on_screen': good5, 'stages_ms': st5, 'allocator_in_timed_steps': alloc5},
              'wave': {'ms': msw, 'samples': Nw, 'receiving_points': Mw, 'pairs_per_s': pairs5,
                       'value': FLOP_PER_PAIR * pairs5 / 1e9, 'unit': 'GFLOP/s (57 flop/pair), all GPUs',
                       'frac_of_fp64_peak_per_gpu': FLOP_PER_PAIR * pairs5 / 1e12 / world / dfma_tflops,
                       'note': 'waves.diffract(comm=, samplesSharded=True) through the general-normal kernel (the KB '
                               'surface normals differ per sample); includes prepare-side packing, the all-gather '
                               'and the post-processing of diffract'}}
        del lbs

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'Mrays/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'rays_per_gpu': nrays, 'nodes': nodes, 'with_amplitudes': True,
                           'chain': 'reflect + expose fused in one kernel (lb materialised)' if fused_chain else
                                    'three separate calls (gb through HBM)',
                           'l2': 'inputs larger than L2 (>= 13 GB of Beam arrays per step)',
                           'parallelism': 'rays sharded over %d GPU(s); slot all-reduce %s' % (
                               world, 'through the C-ABI NCCL entry points' if world > 1 and not args.torch_collectives
                               else ('torch.distributed' if world > 1 else 'n/a'))},
                'e2e': {'value': e2e_value, 'unit': 'Mrays/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e,
                        'note': 'the WHOLE screen Beam (132 B x every ray) to pinned host memory on a copy stream; '
                                'PCIe / host-memory bound: see d2h_bandwidth; e2e_compact and e2e_histogram move what a '
                                'consumer reads' + ('; ' + e2e_note if e2e_note else '')},
                'd2h_bandwidth': d2h_gbs, 'e2e_compact': e2e_compact, 'e2e_histogram': e2e_hist,
                'gather_beam': gather_rep,
                'gpu_launches': (launches[0] * args.steps) if launches else args.steps * (7 if fused_chain else 8),
                'gpu_launches_source': ('torch profiler (CUPTI) on one step: %d kernels of libxrt_b200.so, %d others'
                                        % launches) if launches else 'static count of the call path',
                'clocks': sampler.summary(),
                'roofline': roofline, 'roofline_second_kernel': roofline_other, 'stages': hbm, 'chain_ms_per_step': chain,
                'shine_rejection': rej, 'dcm': dcm_rep, 'kirchhoff': kir, 'config5': c5,
                'cpu_baseline': cpu_baseline() if world == 1 and not args.no_cpu_baseline else None}
        print(json.dumps(_finite(line)))
    if world > 1:
        comm.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--nrays', type=int, default=100000000, help='rays per GPU per step')
    ap.add_argument('--unfused-chain', action='store_true',
                    help='time reflect and expose as separate calls (the global beam goes through HBM)')
    ap.add_argument('--torch-collectives', action='store_true',
                    help='exchange steps through torch.distributed instead of the C-ABI NCCL entry points')
    ap.add_argument('--no-config5', action='store_true', help='skip the config-5 (DCM + KB + wave) sub-report')
    ap.add_argument('--config5-nrays', type=int, default=0, help='rays per GPU in config 5 (default 1.25 x --nrays)')
    ap.add_argument('--config5-wave-n', type=int, default=1000000)
    ap.add_argument('--config5-wave-m', type=int, default=1024 * 1024)
    ap.add_argument('--no-kirchhoff', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-rejection', action='store_true')
    ap.add_argument('--no-dcm', action='store_true', help='skip the config-3 (DCM) sub-report')
    ap.add_argument('--kirchhoff-n', type=int, default=1000000)
    ap.add_argument('--kirchhoff-m', type=int, default=1024 * 1024)
    ap.add_argument('--kirchhoff-r', type=int, default=4)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = max(args.warmup, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
