#!/usr/bin/env python
"""Round-2 experiment: the tensor-path kernel at constant bytes (N * nx = 2e8 doubles = 1.6 GB)
over the channel count -- how much of the gap to the roofline at 200 channels is per-tile cost
(epilogue every 13 chunks) and how much the half-empty last chunk.

    python tools/r2_nx_sweep.py [--out gpurun_out/r2_nx_sweep.json]
"""
import argparse
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402


def timed(ds, reps):
    for _ in range(40):
        ds.launch_clike(0.01, -0.5)
    best = 1e30
    for _ in range(3):
        ds.timer_start()
        for _ in range(reps):
            ds.launch_clike(0.01, -0.5)
        best = min(best, ds.timer_stop() / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'r2_nx_sweep.json'))
    ap.add_argument('--cells', type=float, default=2e8)
    ap.add_argument('--nx', default='96,192,200,208,400,800,1000')
    ap.add_argument('--K', default='8,16')
    ap.add_argument('--tunings', default='0,0,0,0', help='semicolon-separated lanes,unroll,ktile,rows')
    ap.add_argument('--horns', action='store_true', help='gensimple_horns-style data instead of noise')
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    lib = _lib.load()
    res = []
    for nx in [int(v) for v in args.nx.split(',')]:
        n = int(args.cells // nx)
        x = numpy.linspace(400., 800., nx)
        rg = numpy.random.default_rng(nx)
        if args.horns:
            x, y, _ = synth.horns(n, nx=nx, legacy=False, seed=1000)
        else:
            y = rg.standard_normal((nx, n))
            y *= synth.NOISE_LEVEL
        ds = ResidentDataset(x, y)
        del y
        ds.set_mask(None)
        for K in [int(v) for v in args.K.split(',')]:
            ds.stage_params(synth.parameter_points(K, seed=7))
            for tun in args.tunings.split(';'):
                ds.set_tuning(*[int(v) for v in tun.split(',')])
                ms = timed(ds, 30)
                b = n * nx * 8 + K * nx * 8 + K * n * 8 + n
                row = {'nx': nx, 'n': n, 'K': K, 'tuning': tun, 'ms': round(ms, 5),
                       'frac': round(b / (ms * 1e-3) / 1e9 / peak, 4),
                       'kernel': lib.mdns_last_kernel().decode()}
                res.append(row)
                print(row, flush=True)
            ds.set_tuning(0, 0, 0, 0)
        ds.close()
    with open(args.out, 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
