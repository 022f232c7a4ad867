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
scalars of reflect() and Imax are all-reduced (min/max over 64 slots); `--gather beam` adds the all-gather of
the screen Beam named in the north star.
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from xrt_b200 import _lib, kernels, synthetic
    from xrt_b200.beamline import BeamLine
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
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    comm = parallel.Comm(world, rank)

    nrays = args.nrays
    bl = BeamLine()
    src = Undulator(bl, 'u', eMin=E1 * 0.995, eMax=E1 * 1.005, seed=1234, **undulator_kwargs(nrays))
    mirror = ToroidMirror(bl, 'tm', material=Material('Au', rho=19.3), **MIRROR)
    screen = Screen(bl, 'scr', center=(0, SCREEN_Y, 25000. * np.tan(2 * 3e-3)))
    src.reset()
    nodes = len(src.tg)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    stage_names = ('sample', 'undulator', 'fill', 'reflect', 'expose')
    stage_events = []
    host_out = None
    e2e_count = [0]
    copy_stream = torch.cuda.Stream()

    def step(e2e=False, record=None, fused=not args.unfused_source):
        marks = [ev() for _ in range(6)] if record is not None else None
        off = rank * nrays
        from xrt_b200.beams import Beam
        beam = Beam(nrays, withAmplitudes=True, uninitialized=True)
        if fused:
            # Undulator.shine(): one fused kernel (sample + field integral + Beam assembly)
            if marks:
                marks[0].record()
                marks[1].record()
            src.shine_uniform_fused(beam, index_offset=off)
            if marks:
                marks[2].record()
                marks[3].record()
        else:
            if marks:
                marks[0].record()
            s = src.draw_samples(nrays, index_offset=off)
            if marks:
                marks[1].record()
            I, mJs, mJp = src.build_I_map_device(s['E'], s['theta'], s['psi'])
            if marks:
                marks[2].record()
            src.fill_beam(s, mJs, mJp, beam)
            if marks:
                marks[3].record()
        gb, lb = mirror.reflect(beam, needLocal=True, checkTables=False, index_offset=off,
                                allreduce=comm.allreduce_reductions if world > 1 else None)
        if marks:
            marks[4].record()
        blo = screen.expose(gb)
        if marks:
            marks[5].record()
            record.append(marks)
        if args.gather == 'beam' and world > 1:
            blo = comm.allgather_beam(blo)
        if e2e:
            # D2H of the result on a copy stream: the copy of step k overlaps the kernels of step k+1; two pinned
            # host buffers alternate so a consumer owns one result while the next one lands
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

    # ---- device-resident throughput
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(lambda: step(record=stage_events), args.steps)
    torch.cuda.synchronize()
    stage_ms = {n: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in stage_events]))
                for i, n in enumerate(stage_names)}
    ms_step = ms_total / args.steps
    value = world * nrays / (ms_step * 1e-3) / 1e6

    # ---- end to end: result Beam lands in pinned host memory inside the timed region
    blo = step()
    bytes_per_ray = blo.bytes_per_ray()
    n_host = len(blo)
    host_out = []
    e2e_note = None
    for _ in range(2):  # two pinned result buffers (2 x 13.2 GB at 1e8 rays); one if the host cannot pin that much
        try:
            host_out.append({k: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                             for k, t in blo.tensors().items()})
        except RuntimeError:
            if not host_out:
                # not even one pinned Beam: pageable host memory keeps the run alive (slower copies, said in the line)
                host_out.append({k: torch.empty(t.shape, dtype=t.dtype) for k, t in blo.tensors().items()})
                e2e_note = 'pinned allocation failed: pageable host buffers'
            break
    del blo
    for _ in range(max(1, args.warmup - 1)):
        step(e2e=True)
    ms_e2e = timed(lambda: step(e2e=True), args.steps) / args.steps  # barrier() drains the copy stream
    if rank == 0:
        sampler.stop_flag.set()
    e2e_value = world * nrays / (ms_e2e * 1e-3) / 1e6
    h2d = 6 * nodes * 8 + 8 * 64  # node tables + scalar structs per step
    d2h = n_host * bytes_per_ray
    host_out = None

    # ---- end to end the way xrt's runner consumes a repeat: only the screen image (get_output + do_hist2d,
    # xrt/multipro.py:119-171) leaves the GPU -- device histogram, 256 x 256 doubles D2H per step
    from xrt_b200 import histogram
    blo = step()
    gd = blo.state == 1
    xl = [float(blo.x[gd].min()), float(blo.x[gd].max())]
    zl = [float(blo.z[gd].min()), float(blo.z[gd].max())]
    blo_E_min, blo_E_max = blo.E[gd].min(), blo.E[gd].max()
    del blo, gd
    el = [float(blo_E_min), float(blo_E_max)]
    ph = histogram.PlotHistograms(histogram.Axis('x', xl, 256), histogram.Axis('z', zl, 256),
                                  histogram.Axis('E', el, 256), fluxKind='total', rayFlag=(1,))
    hbytes = sum(t.numel() * t.element_size() for t in (ph._h2, ph._hx, ph._hy, ph._hc, ph._counters))
    host_img = torch.empty(hbytes // 8, dtype=torch.float64, pin_memory=True)

    def step_hist():
        b = step()
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
                        'the x, z, E 1-D histograms, each with flux + RGB channels == get_output + do_hist1d/2d); '
                        'only the histograms cross PCIe, as in xrt where histograms, not rays, are accumulated over '
                        'repeats (xrt/runner.py:526-530, xrt/multipro.py:296-340)'}

    # ---- per-kernel times of reflect() (pre-pass A, pre-pass B, main) in a few extra untimed steps
    mirror.profileEvents = []
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    refl_ms = [float(np.mean([e[a].elapsed_time(e[b]) for e in mirror.profileEvents]))
               for a, b in ((0, 1), (4, 2), (5, 3))]
    mirror.profileEvents = None

    # ---- rooflines.  Denominators: HBM from MEASURED_PEAKS.json; FP64 has no entry there, so a DFMA stream is
    # measured in this run (xrt_b200_fp64_probe).
    peaks, peak_kind = measured_peaks()
    dfma_tflops, _ = kernels.fp64_probe(0, 8192)
    und_ms = stage_ms['undulator']
    und_tflops = FLOP_PER_SAMPLE_NODE * nrays * nodes / (und_ms * 1e-3) / 1e12
    pair_path = bool(src.gridSymmetric and src.Kx == 0)
    und_instr = 24.5 if pair_path else 57.0  # DFMA-pipe instructions per (sample, node) in SASS
    roofline_und = {'kernel': ('undulator_kernel<2,1,planar,%s> (Undulator.build_I_map)' if args.unfused_source else
                               'shine_uniform_kernel<2,planar,%s> (Undulator.shine: sample + field integral + Beam '
                               'assembly)') % ('pairs' if pair_path else 'nodes'),
                    'bound': 'fp64', 'achieved': und_tflops, 'peak': dfma_tflops, 'unit': 'TFLOP/s',
                    'frac': und_tflops / dfma_tflops, 'traffic': measured_traffic('undulator_kernel' if args.unfused_source else 'shine_uniform_kernel', nrays),
                    'issue_slot_frac': 2 * und_instr * nrays * nodes / (und_ms * 1e-3) / 1e12 / dfma_tflops,
                    'note': 'achieved counts the 63 flop per (sample, node) of the reference algorithm (SURVEY 8d); '
                            'the kernel evaluates mirror PAIRS of nodes for a planar undulator on a symmetric grid '
                            '(49 DFMA-pipe instructions per pair), so issue_slot_frac is the pipe utilisation',
                    'ms': und_ms, 'share_of_step': und_ms / ms_step}
    main_ms = refl_ms[2]
    main_bytes = 132 + 132 + 140
    main_gbs = main_bytes * nrays / (main_ms * 1e-3) / 1e9
    roofline_main = {'kernel': 'reflect_main_kernel<toroid, mirror> (OE.reflect)', 'bound': 'hbm',
                     'achieved': main_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': main_gbs / peaks['hbm_gbs'],
                     'traffic': measured_traffic('reflect_main_kernel', nrays),
                     'peak_source': '%s HBM copy bandwidth (MEASURED_PEAKS.json)' % peak_kind,
                     'note': '%d algorithmic bytes per ray (132 in, 132 gb, 140 lb); currently limited by the issue rate '
                             'of the IEEE div/sqrt chains numpy parity requires (~3600 instructions per ray), not by HBM'
                             % main_bytes,
                     'ms': main_ms, 'share_of_step': main_ms / ms_step}
    # the dominant kernel of the step goes first
    roofline = roofline_main if main_ms >= und_ms else roofline_und
    roofline_other = roofline_und if main_ms >= und_ms else roofline_main
    hbm = {}
    stage_list = [('reflect_prepass_dir', 28, refl_ms[0]), ('reflect_prepass_bracket', 52, refl_ms[1]),
                  ('reflect_main', main_bytes, refl_ms[2]), ('expose', 264, stage_ms['expose'])]
    if args.unfused_source:
        stage_list += [('fill', 96 + 132, stage_ms['fill']), ('sample', 64, stage_ms['sample'])]
    for name, bpr, ms in stage_list:
        gbs = bpr * nrays / (ms * 1e-3) / 1e9
        hbm[name] = {'ms': ms, 'bytes_per_ray': bpr, 'achieved_gbs': gbs,
                     'frac_of_%s_hbm' % peak_kind: gbs / peaks['hbm_gbs']}
    hbm['undulator' if args.unfused_source else 'shine_fused (sample + undulator + fill)'] = {'ms': und_ms}
    hbm['reflect_total_in_step'] = {'ms': stage_ms['reflect']}

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
        ar = comm.allreduce_reductions if world > 1 else None

        def step3():
            b = gs.shine(withAmplitudes=True, index_offset=rank * nrays)
            gb, lo1, lo2 = dcm.double_reflect(b, needLocal=True, checkTables=False, index_offset=rank * nrays,
                                              allreduce=ar)
            return scr3.expose(gb)
        out3 = step3()
        good3 = float((out3.state == 1).sum()) / len(out3)
        del out3
        step3()
        ms3 = timed(step3, max(1, min(args.steps, 3))) / max(1, min(args.steps, 3))
        dcm_rep = {'workload': 'config3: GeometricSource -> DCM Si(111) (two crystals, thick Bragg amplitudes) -> '
                               'Screen, %d rays per GPU' % nrays, 'ms_per_step': ms3,
                   'value': world * nrays / (ms3 * 1e-3) / 1e6, 'unit': 'Mrays/s', 'good_fraction': good3}
        del gs, dcm, scr3

    # ---- Kirchhoff sum, config 4
    kir = None
    if not args.no_kirchhoff:
        N, M = args.kirchhoff_n, args.kirchhoff_m
        Mloc = M // world
        case = synthetic.kirchhoff_case(7, N, M, mesh=(int(round(M ** 0.5)) ** 2 == M))
        sl = slice(rank * Mloc, (rank + 1) * Mloc)
        d = {k: torch.from_numpy(np.ascontiguousarray(case[k][sl])).cuda() for k in ('xM', 'yM', 'zM')}
        for k in ('nl', 'k', 'Es', 'Ep'):
            d[k] = torch.from_numpy(np.ascontiguousarray(case[k])).cuda()
        d['pos'] = torch.from_numpy(np.ascontiguousarray(case['pos'].T)).cuda()
        d['nrm'] = torch.from_numpy(np.ascontiguousarray(case['nrm'].T)).cuda()
        out = torch.empty((5, Mloc), dtype=torch.complex128, device='cuda')

        def kstep():
            kernels.kirchhoff(d['xM'], d['yM'], d['zM'], d['nl'], d['Es'], d['Ep'], d['k'], d['pos'], d['nrm'],
                              out=out, R=args.kirchhoff_r, uniform_normal=True)  # slit: n = (0, 1, 0)
        kstep()
        ksteps = max(1, min(args.steps, 2))
        kms = timed(kstep, ksteps) / ksteps
        pairs = float(N) * M / (kms * 1e-3)
        ktf = FLOP_PER_PAIR * pairs / 1e12
        kir = {'workload': 'config4: %d samples -> %d receiving points (sharded over %d GPU(s)), fp64, strict '
                           'numpy rounding' % (N, M, world),
               'ms': kms, 'pairs_per_s': pairs, 'value': ktf * 1e3, 'unit': 'GFLOP/s (57 flop/pair), all GPUs',
               'roofline': {'bound': 'fp64', 'achieved': ktf / world, 'peak': dfma_tflops, 'unit': 'TFLOP/s per GPU',
                            'frac': ktf / world / dfma_tflops, 'traffic': None,
                            'issue_slot_frac': 65 * 2 * pairs / world / 1e12 / dfma_tflops,
                            'note': '65 DFMA-pipe instructions per pair in SASS (uniform-normal variant); frac counts '
                                    'the 57 algorithmic flops, issue_slot_frac the issued ones'}}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': 'Mrays/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'rays_per_gpu': nrays, 'nodes': nodes, 'with_amplitudes': True,
                           'l2': 'inputs larger than L2 (>= 13 GB of Beam arrays per step)',
                           'parallelism': 'rays sharded over %d GPU(s), gather=%s' % (world, args.gather)},
                'e2e': {'value': e2e_value, 'unit': 'Mrays/s', 'h2d_bytes_per_step': h2d,
                        'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e,
                        **({'note': e2e_note} if e2e_note else {})},
                'e2e_histogram': e2e_hist,
                'gpu_launches': args.steps * (9 if args.unfused_source else 7),  # source (1 fused or 3), 2 slot inits,
                # pre-pass A/B, reflect main, expose
                'clocks': sampler.summary(),
                'roofline': roofline, 'roofline_second_kernel': roofline_other, 'stages': hbm, 'shine_rejection': rej, 'dcm': dcm_rep, 'kirchhoff': kir,
                'cpu_baseline': cpu_baseline() if world == 1 and not args.no_cpu_baseline else None}
        print(json.dumps(_finite(line)))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--nrays', type=int, default=100000000, help='rays per GPU per step')
    ap.add_argument('--gather', default='none', choices=['none', 'beam'])
    ap.add_argument('--no-kirchhoff', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-rejection', action='store_true')
    ap.add_argument('--no-dcm', action='store_true', help='skip the config-3 (DCM) sub-report')
    ap.add_argument('--unfused-source', action='store_true',
                    help='time Undulator.shine as three kernels (sample, field integral, fill) instead of the fused one')
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
