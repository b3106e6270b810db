#!/usr/bin/env python
"""Live-point table reductions, warmed up: LiveTable.prepare (min / argmin / max per data set, results
to the host) and stage_thresholds (the same kernel, minima straight into the data set's threshold
buffer).   python tools/r2_livetable.py"""
import numpy, time, sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massivedatans_b200 import synth, _lib
from massivedatans_b200.likelihood import ResidentDataset
from massivedatans_b200.livepoints import LiveTable
N, nlive = 200000, 400
x, y, _ = synth.horns(N, nx=16, legacy=False, seed=1)
ds = ResidentDataset(x, y)
L = numpy.random.RandomState(1).normal(size=(nlive, N)) * 100
t = LiveTable(ds, nlive); t.upload(L)
for _ in range(5): r = t.prepare()
t0 = time.perf_counter()
for _ in range(50): r = t.prepare()
dt = (time.perf_counter() - t0) / 50
print('prepare ms', dt * 1e3, 'GB/s', L.nbytes / dt / 1e9, all(numpy.array_equal(a, b) for a, b in zip(r, (L.min(axis=0), L.argmin(axis=0), L.max(axis=0)))))
for _ in range(3): t.stage_thresholds()
t0 = time.perf_counter()
for _ in range(50): t.stage_thresholds()
dt = (time.perf_counter() - t0) / 50
print('stage_thresholds ms', dt * 1e3, 'GB/s', L.nbytes / dt / 1e9)
