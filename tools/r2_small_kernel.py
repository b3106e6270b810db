#!/usr/bin/env python
"""Small batches: the one-launch direct-form kernel (clike_small_kernel, set_tuning(7)) against the
automatic choice without it (MDNS_SMALL_EVALS=0: model kernel + tensor path / lanes-across-channels
kernel + fix-up), device-timed with the L2 flushed and back to back, plus one accept pass end to
end (begin_draw once, draw_counts per pass, wall clock).   python tools/r2_small_kernel.py"""
import json
import os
import sys
import time

os.environ['MDNS_SMALL_EVALS'] = '0'          # automatic choice = the tree before this kernel
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy  # noqa: E402
import bench  # noqa: E402
from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402

lib = _lib.load()
res = []
for n in (10000, 30000, 100000):
    x, y, _ = synth.horns(n, nx=200, legacy=False, seed=1000)
    ds = ResidentDataset(x, y)
    for mname in ('all', 'half'):
        mask = None if mname == 'all' else synth.masks(n, seed=11)['half']
        n_act = ds.set_mask(mask)
        for K in (2, 4, 8, 16, 32):
            ds.stage_params(synth.parameter_points(K, seed=7))
            row = {'n': n, 'mask': mname, 'n_act': n_act, 'K': K}
            for arm, tun in (('before', (0, 0, 0, 0)), ('small', (7, 0, 0, 0))):
                ds.set_tuning(*tun)
                t = bench.device_time(ds, 40, flush=True)
                tb = bench.device_time(ds, 40, flush=False)
                row[arm] = {'ms_flushed': round(t, 5), 'ms_back_to_back': round(tb, 5),
                            'kernel': lib.mdns_last_kernel().decode()}
            if mname == 'all' and K == 16:
                # one speculative pass end to end: K points in, K counts out
                Lmins = numpy.full(n_act, 1e300)
                pts = synth.parameter_points(K, seed=7)
                for arm, tun in (('before', (0, 0, 0, 0)), ('small', (7, 0, 0, 0))):
                    ds.set_tuning(*tun)
                    ds.begin_draw(mask, Lmins)
                    for _ in range(20):
                        ds.draw_counts(pts, synth.NOISE_LEVEL)
                    t0 = time.perf_counter()
                    for _ in range(200):
                        ds.draw_counts(pts, synth.NOISE_LEVEL)
                    row[arm]['accept_pass_e2e_ms'] = round((time.perf_counter() - t0) / 200 * 1e3, 5)
            ds.set_tuning(0, 0, 0, 0)
            res.append(row)
            print(row, flush=True)
    ds.close()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'r2_small_kernel.json'), 'w'), indent=1)
