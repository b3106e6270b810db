#!/usr/bin/env python
"""Stream-K kernel against the slab kernel at the bench shape, three rounds in one process: 300
launches back to back and 40 launches with the L2 flushed before each -- how the two drift as the
board warms up under its power cap.   python tools/r2_warmup_drift.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402

peak = bench.hbm_peak()[0]
lib = _lib.load()
n, K = 1000000, 16
x, y, _ = synth.horns(n, nx=200, legacy=False, seed=1000)
ds = ResidentDataset(x, y)
ds.set_mask(None)
ds.stage_params(synth.parameter_points(K, seed=7))
for rep in range(3):
    for tun in ('3,0,16,3', '6,0,16,2'):
        ds.set_tuning(*[int(v) for v in tun.split(',')])
        tb = bench.device_time(ds, 300, flush=False)
        t = bench.device_time(ds, 40, flush=True)
        b = bench.algorithmic_bytes(n, n, 200, K)
        print(rep, tun, 'flushed', round(t, 5), 'back-to-back x300', round(tb, 5),
              round(b / (tb * 1e-3) / 1e9 / peak, 3), lib.mdns_last_kernel().decode(), flush=True)
