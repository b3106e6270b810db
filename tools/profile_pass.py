#!/usr/bin/env python
"""Where the time of one speculative likelihood pass goes on the host (cProfile): the focussed
regime of tools/bench_constrainer.py -- 10^5 data sets resident, one active, K candidates.

    python tools/profile_pass.py [--K 16] [--passes 3000]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import make_multi_loglikelihood  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ndata', type=int, default=100000)
    ap.add_argument('--K', type=int, default=16)
    ap.add_argument('--passes', type=int, default=3000)
    args = ap.parse_args()
    x, y, _ = synth.horns(args.ndata, legacy=False, seed=3)
    like = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
    mask = numpy.zeros(args.ndata, dtype=bool)
    mask[1234] = True
    Lmins = numpy.array([1e300])                     # nothing is accepted: every pass is a full one
    pts = synth.parameter_points(args.K, seed=1)
    pts[:, 2] = numpy.log10(pts[:, 2])
    xs = [p for p in pts]

    def one_pass():
        like.speculate(xs, Lmins)
        like(xs[0], mask)
        return like.last_draw

    for _ in range(50):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(args.passes):
        one_pass()
    print('%.1f us per pass (K = %d, one active data set of %d)'
          % (1e6 * (time.perf_counter() - t0) / args.passes, args.K, args.ndata))
    t0 = time.perf_counter()
    for _ in range(args.passes):
        like(xs[0], mask)
    print('%.1f us per plain K = 1 call' % (1e6 * (time.perf_counter() - t0) / args.passes))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(args.passes):
        one_pass()
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(18)


if __name__ == '__main__':
    main()
