#!/usr/bin/env python
"""MUSE-type likelihood, device-timed: direct kernels against the expanded form on the tensor path,
on the cube of the reference shape and on a 40 000 x 3600 cube.   python tools/r2_muse.py"""
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402
from oracle import ref  # noqa: E402
import bench  # noqa: E402

peak = bench.hbm_peak()[0]
res = []
for ndata, nspec in ((4223, 3600), (40000, 3600)):
    y, v, t = synth.muse(ndata=ndata, nspec=nspec)
    ds = ResidentDataset(None, y, variance=v)
    for mname, mask in (('all', numpy.ones(ndata, dtype=bool)),
                        ('p70', numpy.random.RandomState(2).uniform(size=ndata) < 0.7)):
        n_act = int(mask.sum())
        ds.set_mask(mask)
        for K in (1, 4, 16, 64):
            ypreds = numpy.array([synth.muse_template(nspec, phase=0.05 + 0.1 * k) for k in range(K)])
            ds.stage_spectra(ypreds)
            row = {'ndata': ndata, 'mask': mname, 'K': K}
            for name, en in (('direct', False), ('expanded', True)):
                if name == 'direct' and K == 64:
                    continue
                ds.set_expanded(en)
                for _ in range(3):
                    ds.launch_muse()
                ds.sync()
                reps = 30 if ndata < 10000 else 10
                ds.timer_start()
                for _ in range(reps):
                    ds.launch_muse()
                ms = ds.timer_stop() / reps
                b = bench.muse_bytes(n_act, ndata, nspec, K)
                row[name] = {'ms': ms, 'frac': b / (ms * 1e-3) / 1e9 / peak,
                             'kernel': ds._lib.mdns_last_kernel().decode()}
            L = numpy.zeros((K, ndata))
            ds.muse_loglike(ypreds, mask, L)
            want = ref.cmuselike(y, v, ypreds[K - 1], mask)
            row['rel_err_vs_reference'] = float(numpy.max(numpy.abs(L[K - 1][mask] - want[mask]) / numpy.abs(want[mask])))
            row['expanded_stats'] = ds.expanded_stats()
            res.append(row)
            print(row, flush=True)
    ds.close()
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'r2_muse.json'), 'w'), indent=1)
