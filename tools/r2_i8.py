#!/usr/bin/env python
"""The tcgen05 experiment (clike_i8_kernel, tuning lanes = 5) against the FP64 tensor path:
accuracy on horns data and on the cancellation fixture, and device time at K = 64 / 400.

    python tools/r2_i8.py N [K ...]
"""
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1])
Ks = [int(v) for v in sys.argv[2:]] or [64]
peak = bench.hbm_peak()[0]
x, y, _ = synth.horns(n, legacy=False, seed=1000)
ds = ResidentDataset(x, y)
ds.set_mask(None)
res = []
for K in Ks:
    pts = synth.parameter_points(K, seed=7)
    ds.stage_params(pts)
    row = {'n': n, 'K': K}
    outs = {}
    for name, tun in (('dmma', (0, 0, 0, 0)), ('i8', (5, 0, 0, 0))):
        ds.set_tuning(*tun)
        o = numpy.empty((K, n))
        ds.launch_clike(0.01, 1.0)
        ds.fetch(o)
        outs[name] = o
        for _ in range(3):
            ds.launch_clike(0.01, 1.0)
        ds.sync()
        reps = 10 if K <= 64 else 3
        ds.timer_start()
        for _ in range(reps):
            ds.launch_clike(0.01, 1.0)
        ms = ds.timer_stop() / reps
        row[name] = {'ms': ms, 'evals_per_s': K * n / (ms * 1e-3), 'kernel': ds._lib.mdns_last_kernel().decode(),
                     'hbm_frac_fp64_bytes': bench.algorithmic_bytes(n, n, 200, K) / (ms * 1e-3) / 1e9 / peak}
    rel = numpy.abs(outs['i8'] - outs['dmma']) / numpy.abs(outs['dmma'])
    row['max_rel_i8_vs_dmma'] = float(rel.max())
    row['expanded_stats'] = ds.expanded_stats()
    # against the FP64 direct form (numpy) on a sample of rows
    rows = numpy.linspace(0, n - 1, 200).astype(int)
    for k in (0, K // 2, K - 1):
        p = pts[k]
        m = p[0] * numpy.exp(-0.5 * ((p[1] - x) / p[2]) ** 2)
        want = (((m[:, None] - y[:, rows]) / 0.01) ** 2).sum(axis=0)
        row.setdefault('max_rel_i8_vs_numpy', 0.0)
        row['max_rel_i8_vs_numpy'] = max(row['max_rel_i8_vs_numpy'],
                                         float(numpy.max(numpy.abs(outs['i8'][k][rows] - want) / want)))
    res.append(row)
    print(row, flush=True)
ds.set_tuning(0, 0, 0, 0)
out = os.path.join(ROOT, 'gpurun_out', 'r2_i8_%d.json' % n)
json.dump(res, open(out, 'w'), indent=1)
