#!/usr/bin/env python
"""Masked candidate batches: gather4-fed tensor path vs the lanes-across-channels block kernel.
    python tools/sweep_masked.py [--ndata 1000000] [--nx 200]"""
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
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    x, y, _ = synth.horns(args.ndata, nx=args.nx, legacy=False, seed=1000)
    ds = ResidentDataset(x, y)
    lib = _lib.load()
    masks = synth.masks(args.ndata)
    for K in (8, 16):
        ds.stage_params(synth.parameter_points(K, seed=7))
        for mname in ('half', 'prefix'):
            n_act = ds.set_mask(masks[mname])
            for label, setup in (('auto', lambda: (ds.set_expanded(True), ds.set_tuning(0, 0, 0, 0))),
                                 ('direct', lambda: (ds.set_expanded(False), ds.set_tuning(0, 0, 0, 0))),
                                 ('gather 8w', lambda: (ds.set_expanded(True), ds.set_tuning(3, 0, min(K, 32) if K >= 8 else 8, 3))),
                                 ('gather 16w', lambda: (ds.set_expanded(True), ds.set_tuning(3, 0, min(K, 32) if K >= 8 else 8, 13))),
                                 ('gather 2st', lambda: (ds.set_expanded(True), ds.set_tuning(3, 0, min(K, 32) if K >= 8 else 8, 2))),
                                 ('gather 16w2', lambda: (ds.set_expanded(True), ds.set_tuning(3, 0, min(K, 32) if K >= 8 else 8, 12))),
                                 ('slab 2', lambda: (ds.set_expanded(True), ds.set_tuning(6, 0, 16 if K > 8 else 8, 2))),
                                 ('slab 3', lambda: (ds.set_expanded(True), ds.set_tuning(6, 0, 16 if K > 8 else 8, 3)))):
                setup()
                for _ in range(3):
                    ds.launch_clike(0.01, -0.5)
                ds.sync()
                ds.timer_start()
                for _ in range(args.steps):
                    ds.launch_clike(0.01, -0.5)
                ms = ds.timer_stop() / args.steps
                b = n_act * args.nx * 8 + K * args.nx * 8 + K * n_act * 8 + args.ndata
                print('K=%-3d mask=%-6s %-10s %8.4f ms  %.3e evals/s  hbm %.3f  %s'
                      % (K, mname, label, ms, K * n_act / (ms * 1e-3), b / (ms * 1e-3) / 1e9 / peak,
                         lib.mdns_last_kernel().decode()), flush=True)
    ds.set_expanded(True)
    ds.set_tuning(0, 0, 0, 0)


if __name__ == '__main__':
    main()
