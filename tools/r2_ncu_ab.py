#!/usr/bin/env python
"""One launch each of the stream-K and the whole-tile tensor-path kernel (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402

n, nx, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x, y, _ = synth.horns(n, nx=nx, legacy=False, seed=1000)
os.environ['MDNS_NO_GRAPH'] = '1'
ds = ResidentDataset(x, y)
ds.set_mask(None)
ds.stage_params(synth.parameter_points(K, seed=7))
for tun in ((0, 0, 0, 0), (3, 1, 0, 0), (0, 0, 0, 0), (3, 1, 0, 0)):
    ds.set_tuning(*tun)
    ds.launch_clike(0.01, -0.5)
    ds.sync()
