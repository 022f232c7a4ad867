#!/usr/bin/env python
"""Summarise `ncu --set full` reports into the JSON files kept under profiles/ (B200_PROFILING.md recipe):

    python tools/ncu_summary.py gpurun_out/prof_rays.ncu-rep gpurun_out/prof_kirchhoff.ncu-rep > profiles/rNN_ncu_full_summary.json

One entry per captured launch with the metrics DESIGN.md quotes: duration, DRAM bytes, pipe / issue utilisation,
occupancy, registers, stall reasons per issue."""
import csv
import io
import json
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.avg.per_second']
STALL = 'smsp__average_warps_issue_stalled_'


def short(name):
    return name.split('(')[0].replace('xrtb::', '').replace('void ', '')


def main(paths):
    out = {}
    for path in paths:
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        head, units = rows[0], rows[1]
        for row in rows[2:]:
            d = dict(zip(head, row))
            u = dict(zip(head, units))
            entry = {}
            for k in head:
                if k in KEEP or (k.startswith(STALL) and k.endswith('_per_issue_active.ratio')):
                    if d[k] != '':
                        entry[k] = ('%s %s' % (d[k], u[k])).strip()
            key = short(d['Kernel Name'])
            while key in out:
                key += "'"
            out[key] = entry
    return out


def _bytes(v):
    num, unit = v.split()
    return float(num) * {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]


def traffic(summary, nrays, note):
    """profiles/rNN_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch (and per ray for the
    ray kernels captured at `nrays` rays per launch)."""
    per_launch = {}
    for k, v in summary.items():
        if k.endswith("'") or 'dram__bytes_read.sum' not in v:
            continue
        per_launch[k] = _bytes(v['dram__bytes_read.sum']) + _bytes(v['dram__bytes_write.sum'])
    return {'note': note, 'rays_per_launch': float(nrays), 'bytes_per_launch': per_launch,
            'bytes_per_ray': {k: b / nrays for k, b in per_launch.items() if 'kirchhoff' not in k}}


if __name__ == '__main__':
    args = sys.argv[1:]
    if args and args[0] == '--traffic':
        nrays = int(float(args[1]))
        s = main(args[2:])
        json.dump(traffic(s, nrays, 'dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full '
                          'capture (tools/gpu_profile.sh): ray kernels at %g rays per launch' % nrays), sys.stdout,
                  indent=1)
    else:
        json.dump(main(args), sys.stdout, indent=1)
