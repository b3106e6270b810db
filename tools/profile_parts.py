#!/usr/bin/env python
"""Fixed-size invocations of the neighbour and MUSE kernels for ncu captures and timing.

    python tools/profile_parts.py [--members 50000] [--candidates 100000] [--reps 5]
Prints wall-clock times per call (C-ABI one-shot calls, host arrays in and out)."""
import argparse
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.clustering import neighbors  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--members', type=int, default=50000)
    ap.add_argument('--candidates', type=int, default=100000)
    ap.add_argument('--ndim', type=int, default=3)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--muse-ndata', type=int, default=40000)
    args = ap.parse_args()
    lib = _lib.load()
    _lib.require_device()
    n, m, d = args.members, args.candidates, args.ndim
    xx, yy = synth.members_and_candidates(n, m, d)
    numpy.random.seed(1)
    r = neighbors.bootstrapped_maxdistance(xx, 10)        # warm-up + radius
    t0 = time.perf_counter()
    for _ in range(args.reps):
        numpy.random.seed(1)
        r = neighbors.bootstrapped_maxdistance(xx, 10)
    t_boot = (time.perf_counter() - t0) / args.reps
    print('bootstrapped_maxdistance n=%d d=%d B=10: %.3f ms (incl. numpy chosen matrix), r=%.5f, '
          '%.3e pair tests/s' % (n, d, 1e3 * t_boot, r, 10.0 * n * n * 0.63 * 0.37 / t_boot))
    c = neighbors.count_within_distance_of(xx, r, yy)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        c = neighbors.count_within_distance_of(xx, r, yy)
    t_cnt = (time.perf_counter() - t0) / args.reps
    print('count_within_distance_of n=%d m=%d: %.3f ms, %.3e pair tests/s, mean count %.2f'
          % (n, m, 1e3 * t_cnt, n * float(m) / t_cnt, c.mean()))
    a = neighbors.any_within_distance_of(xx, r, yy)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        a = neighbors.any_within_distance_of(xx, r, yy)
    t_any = (time.perf_counter() - t0) / args.reps
    print('any_within_distance_of: %.3f ms, inside fraction %.3f' % (1e3 * t_any, a.mean()))
    t0 = time.perf_counter()
    for _ in range(args.reps):
        q = neighbors.most_distant_nearest_neighbor(xx)
    print('most_distant_nearest_neighbor: %.3f ms (%.5f)'
          % (1e3 * (time.perf_counter() - t0) / args.reps, q))
    # device-side candidate generation (ball draws + neighbour count + thinning in one kernel)
    from massivedatans_b200.clustering.radfriendsregion import RadFriendsRegion
    region = RadFriendsRegion(members=xx, maxdistance=r)
    pts = region.generate_device(m, seed=1)
    t0 = time.perf_counter()
    for i in range(args.reps):
        pts = region.generate_device(m, seed=1, first_proposal=(i + 1) * m)
    t_gen = (time.perf_counter() - t0) / args.reps
    print('generate_device n=%d proposals=%d: %.3f ms (accepted points to the host included), '
          '%.3e proposals/s, %.3e pair tests/s, accepted fraction %.3f'
          % (n, m, 1e3 * t_gen, m / t_gen, n * float(m) / t_gen, len(pts) / float(m)))
    t0 = time.perf_counter()
    numpy.random.seed(2)
    host = [region._ball_round(1000, d) for _ in range(max(1, m // 1000))]
    t_host = time.perf_counter() - t0
    print('host-RNG ball rounds (numpy draws + device neighbour count), same number of proposals: '
          '%.3f ms, accepted fraction %.3f' % (1e3 * t_host, sum(len(h) for h in host) / float(m)))
    # MUSE-type likelihood, cube wider than L2
    y, v, t = synth.muse(ndata=args.muse_ndata, nspec=3600)
    ds = ResidentDataset(None, y, variance=v)
    for K in (1, 4):
        ypreds = numpy.array([synth.muse_template(3600, phase=0.1 * k) for k in range(K)])
        ds.stage_spectra(ypreds)
        ds.set_mask(None)
        for _ in range(3):
            ds.launch_muse()
        ds.sync()
        ds.timer_start()
        for _ in range(args.reps * 4):
            ds.launch_muse()
        ms = ds.timer_stop() / (args.reps * 4)
        b = args.muse_ndata * 3600 * 16.0
        print('muse K=%d ndata=%d: %.4f ms, %.1f GB/s algorithmic, kernel %s'
              % (K, args.muse_ndata, ms, b / ms / 1e6, lib.mdns_last_kernel().decode()))


if __name__ == '__main__':
    main()
