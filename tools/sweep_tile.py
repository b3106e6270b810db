#!/usr/bin/env python
"""Device-side sweep of the candidate-batch tile kernel variants (experiments).

    python tools/sweep_tile.py [--ndata 1000000] [--nx 200]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ndata', type=int, default=1000000)
    ap.add_argument('--nx', type=int, default=200)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'sweep_tile.json'))
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    x, y, _ = synth.horns(args.ndata, nx=args.nx, legacy=False, seed=1000)
    ds = ResidentDataset(x, y)
    lib = _lib.load()
    rows = []
    variants = {
        8: ['0,0,0,0', '3,0,8,4', '3,0,8,13', '3,0,8,14'],
        16: ['0,0,0,0', '3,0,16,3', '3,0,16,12', '3,0,16,13'],
        32: ['0,0,0,0', '3,0,32,3', '3,0,32,12', '3,0,32,13', '3,0,32,14'],
        64: ['0,0,0,0', '3,0,32,3', '3,0,32,13'],
        400: ['0,0,0,0', '3,0,32,3', '3,0,32,13'],
    }
    for K, vs in variants.items():
        ds.stage_params(synth.parameter_points(K, seed=7))
        ds.set_mask(None)
        for v in vs:
            ds.set_tuning(*[int(t) for t in v.split(',')])
            for _ in range(3):
                ds.launch_clike(0.01, -0.5)
            ds.sync()
            reps = max(3, min(args.steps, 2000 // K))
            ds.timer_start()
            for _ in range(reps):
                ds.launch_clike(0.01, -0.5)
            ms = ds.timer_stop() / reps
            b = args.ndata * args.nx * 8 + K * args.nx * 8 + K * args.ndata * 8 + args.ndata
            row = dict(K=K, tuning=v, ms=ms, evals_per_s=K * args.ndata / (ms * 1e-3),
                       hbm_frac=b / (ms * 1e-3) / 1e9 / peak, kernel=lib.mdns_last_kernel().decode())
            rows.append(row)
            print('K=%-3d tuning=%-11s %8.4f ms %.3e evals/s hbm %.3f %s'
                  % (K, v, ms, row['evals_per_s'], row['hbm_frac'], row['kernel']), flush=True)
    with open(args.out, 'w') as f:
        json.dump(rows, f, indent=1)


if __name__ == '__main__':
    main()
