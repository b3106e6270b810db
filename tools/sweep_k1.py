#!/usr/bin/env python
"""Sweep of the streaming kernels (K = 1, 2, 4) over lanes / fragments in flight for a given
shape (experiments).   python tools/sweep_k1.py --ndata 125000 --nx 1000"""
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
    ap.add_argument('--ndata', type=int, default=125000)
    ap.add_argument('--nx', type=int, default=1000)
    ap.add_argument('--steps', type=int, default=50)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    x, y, _ = synth.horns(args.ndata, nx=args.nx, legacy=False, seed=5)
    ds = ResidentDataset(x, y)
    lib = _lib.load()
    for K, variants in ((1, ['0,0,0,0', '8,4,1,1', '8,8,1,1', '8,13,1,1', '8,16,1,1', '32,2,1,1', '32,4,1,1',
                              '32,8,1,1', '32,16,1,1']),
                        (2, ['0,0,0,0', '8,8,2,1', '8,16,2,1', '32,4,2,1', '32,8,2,1', '32,16,2,1',
                              '8,2,4,4', '8,4,4,2', '3,0,8,2']),
                        (4, ['0,0,0,0', '8,2,4,4', '8,4,4,2', '8,1,4,4', '3,0,8,2', '3,0,8,3'])):
        ds.stage_params(synth.parameter_points(K, seed=7))
        ds.set_mask(None)
        for v in variants:
            ds.set_tuning(*[int(t) for t in v.split(',')])
            for _ in range(3):
                ds.launch_clike(0.01, -0.5)
            ds.sync()
            ds.timer_start()
            for _ in range(args.steps):
                ds.launch_clike(0.01, -0.5)
            ms = ds.timer_stop() / args.steps
            b = args.ndata * args.nx * 8 + K * args.nx * 8 + K * args.ndata * 8 + args.ndata
            print('K=%d tuning=%-10s %8.4f ms  hbm %.3f  %s' % (K, v, ms, b / (ms * 1e-3) / 1e9 / peak,
                                                              lib.mdns_last_kernel().decode()), flush=True)


if __name__ == '__main__':
    main()
