#!/usr/bin/env python
"""A few launches of the likelihood step at a given shape (for ncu): n nx K [tuning]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402

n, nx, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x, y, _ = synth.horns(n, nx=nx, legacy=False, seed=1000) if nx != 1000 else synth.realistic_fast(n, nx=nx) + (None,)
os.environ['MDNS_NO_GRAPH'] = '1'
ds = ResidentDataset(x, y)
ds.set_mask(None)
ds.stage_params(synth.parameter_points(K, seed=7))
if len(sys.argv) > 4:
    ds.set_tuning(*[int(v) for v in sys.argv[4].split(',')])
for _ in range(4):
    ds.launch_clike(0.01, -0.5)
    ds.sync()
