#!/usr/bin/env python
"""Device-side sweep of the likelihood kernels (experiments; results go to gpurun_out/).

    python tools/sweep.py [--ndata 1000000] [--nx 200] [--out gpurun_out/sweep.json]
Times `launch_clike` with CUDA events for several candidate counts K, masks and kernel
variants (lanes per data set, fragments in flight, candidates per pass)."""
import argparse
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402


def time_launch(ds, steps, fn):
    for _ in range(3):
        fn()
    ds.sync()
    ds.timer_start()
    for _ in range(steps):
        fn()
    return ds.timer_stop() / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ndata', type=int, default=1000000)
    ap.add_argument('--nx', type=int, default=200)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'sweep.json'))
    ap.add_argument('--quick', action='store_true')
    args = ap.parse_args()
    peak = 6529.7
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    x, y, _ = synth.horns(args.ndata, nx=args.nx, legacy=False, seed=1000)
    ds = ResidentDataset(x, y)
    masks = synth.masks(args.ndata)
    rows = []

    def record(tag, K, mask_name, tuning, ms):
        n_act = int(masks[mask_name].sum())
        b = n_act * args.nx * 8 + K * args.nx * 8 + K * n_act * 8 + args.ndata
        gbs = b / (ms * 1e-3) / 1e9
        row = dict(tag=tag, K=K, mask=mask_name, tuning=tuning, ms=ms, gbs=gbs, frac=gbs / peak,
                   evals_per_s=K * n_act / (ms * 1e-3))
        rows.append(row)
        print('%-10s K=%-3d mask=%-6s tuning=%-10s %8.4f ms  %7.1f GB/s (%.3f)  %.3e evals/s'
              % (tag, K, mask_name, tuning, ms, gbs, gbs / peak, row['evals_per_s']), flush=True)

    # 1. kernel variants at K=1 and K=8, full mask
    variants = ['0,0,0,0', '8,4,8,2', '8,2,4,4',
                '1,32,8,3', '1,16,8,4', '1,16,16,4', '1,116,4,3', '1,116,8,3', '1,116,8,4',
                '1,116,16,3', '1,116,16,4', '1,116,32,3', '1,132,8,3', '1,132,16,3']
    if args.quick:
        variants = ['0,0,0,0']
    for K in (4, 8, 16):
        pts = synth.parameter_points(K)
        ds.stage_params(pts)
        ds.set_mask(None)
        for t in variants:
            ds.set_tuning(*[int(v) for v in t.split(',')])
            record('variant', K, 'all', t, time_launch(ds, args.steps, lambda: ds.launch_clike(0.01, -0.5)))
    ds.set_tuning(0, 0, 0, 0)
    # 2. K sweep, full mask (ktile variants for K >= 2)
    for K in (1, 2, 4, 8, 16, 32, 64, 400):
        pts = synth.parameter_points(K)
        ds.stage_params(pts)
        ds.set_mask(None)
        steps = max(3, args.steps // max(1, K // 8))
        record('ksweep', K, 'all', 'auto', time_launch(ds, steps, lambda: ds.launch_clike(0.01, -0.5)))
    # 3. masks at K=1 and 8
    for K in (1, 8):
        pts = synth.parameter_points(K)
        ds.stage_params(pts)
        for name in ('all', 'half', 'sparse', 'prefix'):
            ds.set_mask(masks[name])
            record('mask', K, name, 'auto', time_launch(ds, args.steps, lambda: ds.launch_clike(0.01, -0.5)))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
