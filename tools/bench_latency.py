#!/usr/bin/env python
"""Per-call latency of the reference-shaped callable multi_loglikelihood(params, data_mask)
(sample.py:101-108) at the sizes of BASELINE configs[0] / [1] / [2], one candidate per call,
beside the reference's clike.so on this host.

    python tools/bench_latency.py [--out gpurun_out/latency.json]"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import make_multi_loglikelihood  # noqa: E402
from oracle import ref  # noqa: E402


def wall(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'latency.json'))
    args = ap.parse_args()
    rows = []
    for N in (100, 10000, 100000, 1000000):
        x, y, _ = (synth.horns(N, legacy=False, seed=N) if N != 10000 else
                   synth.nothing(N, legacy=False) + (None,))
        f = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
        p = (0.05, 650.0, 0.7)
        for mname, m in synth.masks(N, seed=1).items():
            n_act = int(m.sum())
            if n_act == 0:
                continue
            reps = 2000 if N <= 10000 else (300 if N <= 100000 else 50)
            g = wall(lambda: f(p, m), reps)
            out = numpy.zeros(n_act)

            def cpu():
                out[:] = 0
                ref.clike(x, y, p[0], p[1], 10 ** p[2], synth.NOISE_LEVEL, m, Lout=out)
            c = wall(cpu, max(2, min(200, int(2e7 / (N * 1.0)))))
            row = {'ndata': N, 'mask': mname, 'n_act': n_act, 'gpu_call_us': 1e6 * g,
                   'reference_call_us': 1e6 * c, 'speedup': c / g}
            rows.append(row)
            print(row, flush=True)
    with open(args.out, 'w') as fo:
        json.dump(rows, fo, indent=1)


if __name__ == '__main__':
    main()
