#!/usr/bin/env python
"""Measure the other rows of the hot path beside the reference CPU code:
MUSE-type likelihood (cmuselike.c) and the RadFriends neighbour kernels (cneighbors.c).

    python tools/bench_parts.py [--out gpurun_out/parts.json] [--quick]

GPU numbers: device time via CUDA events for the likelihood (staged interface), wall clock
around the synchronous C-ABI calls for the neighbour functions (they include the H2D of the
candidates and the D2H of the counts: that is the call the reference makes).
CPU numbers: oracle/_ref (the unmodified reference C) on this host, serial and OpenMP.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.clustering import neighbors  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402
from oracle import ref  # noqa: E402


def wall(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'parts.json'))
    ap.add_argument('--quick', action='store_true')
    args = ap.parse_args()
    lib = _lib.load()
    peak = 6529.7
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    res = {'host_cores': os.cpu_count(), 'hbm_peak_gbs': peak, 'muse': [], 'neighbors': []}

    # ---------------- MUSE-type likelihood: cube of the reference shape -----------------
    for ndata, nspec in ([(4223, 3600)] if args.quick else [(4223, 3600), (40000, 3600)]):
        y, v, t = synth.muse(ndata=ndata, nspec=nspec)
        ds = ResidentDataset(None, y, variance=v)
        masks = {'all': numpy.ones(ndata, dtype=bool),
                 'p70': numpy.random.RandomState(2).uniform(size=ndata) < 0.7}
        for mname, mask in masks.items():
            n_act = int(mask.sum())
            for K in (1, 4):
                ypreds = numpy.array([synth.muse_template(nspec, phase=0.1 * k) for k in range(K)])
                ds.stage_spectra(ypreds)
                ds.set_mask(mask)
                for _ in range(3):
                    ds.launch_muse()
                ds.sync()
                reps = 20
                ds.timer_start()
                for _ in range(reps):
                    ds.launch_muse()
                ms = ds.timer_stop() / reps
                bytes_alg = n_act * nspec * 16 + K * nspec * 8 + K * n_act * 8 + ndata
                L = numpy.zeros((K, ndata))
                e2e = wall(lambda: ds.muse_loglike(ypreds, mask, L), 5)
                row = {'ndata': ndata, 'nspec': nspec, 'mask': mname, 'K': K, 'gpu_ms': ms,
                       'gpu_evals_per_s': K * n_act / (ms * 1e-3),
                       'gpu_gbs': bytes_alg / (ms * 1e-3) / 1e9,
                       'hbm_frac': bytes_alg / (ms * 1e-3) / 1e9 / peak,
                       'e2e_ms': e2e * 1e3, 'kernel': lib.mdns_last_kernel().decode()}
                if K == 1 and ndata <= 5000:
                    Lr = numpy.zeros(ndata)
                    cs = wall(lambda: ref.cmuselike(y, v, ypreds[0], mask, Lout=Lr), 2)
                    cp = wall(lambda: ref.cmuselike(y, v, ypreds[0], mask, Lout=Lr, parallel=True), 3)
                    row.update(cpu_serial_ms=cs * 1e3, cpu_openmp_ms=cp * 1e3,
                               cpu_serial_evals_per_s=n_act / cs, cpu_openmp_evals_per_s=n_act / cp)
                res['muse'].append(row)
                print(row, flush=True)
        del ds

    # ---------------- neighbour kernels ---------------------------------------------------
    cases = [(400, 1000, 3), (400, 10000, 3), (5000, 10000, 3), (5000, 1000, 5)]
    if not args.quick:
        cases += [(100000, 10000, 3)]
    for n, m, ndim in cases:
        xx, yy = synth.members_and_candidates(n, m, ndim, seed=n + m)
        chosen = synth.bootstrap_chosen(n, 10, numpy.random.RandomState(1))
        r = lib.mdns_bootstrapped_maxdistance(xx.ctypes.data, n, ndim, chosen.ctypes.data, 10)
        reps = 20 if n <= 5000 else 3
        row = {'members': n, 'candidates': m, 'ndim': ndim, 'radius': r}
        row['gpu_count_ms'] = 1e3 * wall(lambda: neighbors.count_within_distance_of(xx, r, yy), reps)
        row['gpu_any_ms'] = 1e3 * wall(lambda: neighbors.any_within_distance_of(xx, r, yy), reps)
        row['gpu_bootstrap_ms'] = 1e3 * wall(
            lambda: lib.mdns_bootstrapped_maxdistance(xx.ctypes.data, n, ndim, chosen.ctypes.data, 10), reps)
        row['gpu_mdnn_ms'] = 1e3 * wall(lambda: neighbors.most_distant_nearest_neighbor(xx), reps)
        row['gpu_is_within_ms'] = 1e3 * wall(lambda: neighbors.is_within_distance_of(xx, r, yy[0]), reps)
        row['gpu_pair_tests_per_s'] = n * m / (row['gpu_count_ms'] * 1e-3)
        creps = 3 if n * m <= 5e7 else 1
        if n <= 5000:
            row['cpu_count_ms'] = 1e3 * wall(lambda: ref.count_within_distance_of(xx, r, yy), creps)
            row['cpu_any_ms'] = 1e3 * wall(lambda: ref.any_within_distance_of(xx, r, yy), creps)
            row['cpu_bootstrap_ms'] = 1e3 * wall(lambda: ref.bootstrapped_maxdistance_chosen(xx, chosen), creps)
            row['cpu_bootstrap_openmp_ms'] = 1e3 * wall(
                lambda: ref.bootstrapped_maxdistance_chosen(xx, chosen, parallel=True), creps)
            row['cpu_mdnn_ms'] = 1e3 * wall(lambda: ref.most_distant_nearest_neighbor(xx), creps)
            row['cpu_pair_tests_per_s'] = n * m / (row['cpu_count_ms'] * 1e-3)
            assert numpy.array_equal(neighbors.count_within_distance_of(xx, r, yy),
                                     ref.count_within_distance_of(xx, r, yy))
            assert r == ref.bootstrapped_maxdistance_chosen(xx, chosen)
        res['neighbors'].append(row)
        print(row, flush=True)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
