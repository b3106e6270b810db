#!/usr/bin/env python
"""Small launches (N = 1e4 ... 3e5 data sets x 200 channels), L2 flushed before every timed launch:
stream-K kernel against the per-warp slab kernel with 2 / 3 slots.   python tools/r2_small_n.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402

peak = bench.hbm_peak()[0]
lib = _lib.load()
res = []
for n in (10000, 30000, 100000, 300000):
    x, y, _ = synth.horns(n, nx=200, legacy=False, seed=1000)
    ds = ResidentDataset(x, y)
    ds.set_mask(None)
    for K in (16, 8):
        ds.stage_params(synth.parameter_points(K, seed=7))
        for tun in ('0,0,0,0', '3,0,%d,3' % K, '6,0,%d,2' % K, '6,0,%d,3' % K, '8,0,0,0', '32,0,0,0'):
            ds.set_tuning(*[int(v) for v in tun.split(',')])
            t = bench.device_time(ds, 40, flush=True)
            tb = bench.device_time(ds, 40, flush=False)
            b = bench.algorithmic_bytes(n, n, 200, K)
            row = {'n': n, 'K': K, 'tuning': tun, 'ms_flushed': round(t, 5), 'ms_back_to_back': round(tb, 5),
                   'frac_flushed': round(b / (t * 1e-3) / 1e9 / peak, 3), 'kernel': lib.mdns_last_kernel().decode()}
            res.append(row)
            print(row, flush=True)
        ds.set_tuning(0, 0, 0, 0)
    ds.close()
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'r2_small_n.json'), 'w'), indent=1)
