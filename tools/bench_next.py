#!/usr/bin/env python
"""Timing of the "next" rows of the hot path beside their host forms:
the live-point table reductions (prepare / Lmins_higher / replace) and the subset partition.

    python tools/bench_next.py [--ndata 200000] [--nlive 400] [--out gpurun_out/next.json]
Host forms: the numpy expressions of multi_nested_sampler.py:134-137,438-447 and the oracle's
union-find restatement of :204-260 (the reference's own Python loop took 306 s of a 348 s run
at 200 data sets)."""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402
from massivedatans_b200.livepoints import LiveTable  # noqa: E402
from oracle import port  # noqa: E402


def wall(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ndata', type=int, default=200000)
    ap.add_argument('--nlive', type=int, default=400)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'next.json'))
    args = ap.parse_args()
    N, nlive = args.ndata, args.nlive
    rs = numpy.random.RandomState(1)
    x, y, _ = synth.horns(N, nx=16, legacy=False, seed=1)       # the data itself is not used here
    ds = ResidentDataset(x, y)
    L = rs.normal(size=(nlive, N)) * 100
    t = LiveTable(ds, nlive)
    t_up, _ = wall(lambda: t.upload(L), reps=1)
    res = {'ndata': N, 'nlive': nlive, 'table_bytes': L.nbytes, 'upload_ms': 1e3 * t_up}
    # prepare
    g, got = wall(t.prepare)
    c, want = wall(lambda: (L.min(axis=0), L.argmin(axis=0), L.max(axis=0)))
    assert all(numpy.array_equal(a, b) for a, b in zip(got, want))
    res['prepare'] = {'gpu_ms': 1e3 * g, 'numpy_ms': 1e3 * c, 'gpu_gbs': L.nbytes / g / 1e9}
    # Lmins_higher for a tenth of the data sets, shelves of 0..5 entries
    idx = numpy.sort(rs.choice(N, size=N // 10, replace=False))
    shelves = [rs.normal(size=int(k)) * 100 for k in rs.randint(0, 6, size=len(idx))]
    g, got = wall(lambda: t.lmins_higher(idx, shelves))

    def host_rank():
        out = numpy.empty(len(idx))
        for j, d in enumerate(idx):
            n = len(shelves[j])
            out[j] = numpy.partition(numpy.concatenate((L[:, d], shelves[j])), n)[n]
        return out
    c, want = wall(host_rank, reps=1)
    assert numpy.array_equal(got, want)
    res['lmins_higher'] = {'data_sets': len(idx), 'gpu_ms': 1e3 * g, 'numpy_ms': 1e3 * c}
    # replace
    rows = want = None
    lo, at, hi = t.prepare()
    vals = rs.normal(size=N)
    g, _ = wall(lambda: t.replace(at, vals))
    res['replace'] = {'gpu_ms': 1e3 * g}
    # subset partition: groups of ~40 data sets sharing pools of live points
    ngroups = max(1, N // 40)
    group = rs.randint(0, ngroups, size=N)
    pool = 3 * nlive
    P = (group * pool)[None, :] + rs.randint(0, pool, size=(nlive, N))
    P = numpy.ascontiguousarray(P, dtype=numpy.int64)
    npoints = ngroups * pool
    t_up, _ = wall(lambda: t.upload_points(P), reps=1)
    g, got = wall(lambda: t.subsets(None, npoints))
    allm = numpy.ones(N, dtype=bool)
    c, want = wall(lambda: port.subsets_labels(P, allm, npoints), reps=1)
    assert numpy.array_equal(got, want)
    res['subsets'] = {'groups': int((got == numpy.arange(N)).sum()), 'rounds': t.last_rounds,
                      'gpu_ms': 1e3 * g, 'oracle_c_union_find_ms': 1e3 * c,
                      'upload_points_ms': 1e3 * t_up, 'edges': int(P.size)}
    print(json.dumps(res, indent=1))
    with open(args.out, 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
