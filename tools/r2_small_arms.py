#!/usr/bin/env python
"""Small launches, all active and masked (2.5e3 ... 5e4 active data sets): the one-launch direct kernel
(set_tuning 7) against the stream-K (3) and slab (6) tensor-path kernels (gathered for masks) and
what the automatic choice took before the one-launch kernel existed (MDNS_SMALL_EVALS=0),
device-timed with the L2 flushed and back to back.   python tools/r2_small_arms.py"""
import json
import os
import sys

os.environ['MDNS_SMALL_EVALS'] = '0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402

lib = _lib.load()
res = []
for n, mname in ((5000, 'all'), (10000, 'all'), (20000, 'all'), (30000, 'all'), (5000, 'half'), (10000, 'half'),
                 (20000, 'half'), (30000, 'half'), (50000, 'half')):
    x, y, _ = synth.horns(n, nx=200, legacy=False, seed=1000)
    ds = ResidentDataset(x, y)
    n_act = ds.set_mask(None if mname == 'all' else synth.masks(n, seed=11)['half'])
    for K in (4, 8, 16, 32):
        ds.stage_params(synth.parameter_points(K, seed=7))
        kt = 16 if K > 8 else 8
        row = {'n': n, 'mask': mname, 'n_act': n_act, 'K': K}
        for arm, tun in (('before', (0, 0, 0, 0)), ('small', (7, 0, 0, 0)), ('streamk', (3, 0, kt, 3)),
                         ('slab', (6, 0, kt, 2))):
            ds.set_tuning(*tun)
            t = bench.device_time(ds, 30, flush=True)
            tb = bench.device_time(ds, 30, flush=False)
            row[arm] = {'ms_flushed': round(t, 5), 'ms_back_to_back': round(tb, 5),
                        'kernel': lib.mdns_last_kernel().decode()}
        ds.set_tuning(0, 0, 0, 0)
        res.append(row)
        print(row, flush=True)
    ds.close()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'r2_small_arms.json'), 'w'), indent=1)
