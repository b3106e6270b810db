#!/usr/bin/env python
"""MUSE likelihood call end to end with the model on the device:
`multi_loglikelihood_clike(params, data_mask)` of musefuse.py:520-535 = model() + cmuselike.

Cube of the reference shape (3600 channels x 4223 data sets), template grids of the BC03
high-resolution shape (7 metallicities x 111 ages x 6900 wavelengths, synthetic content).
Device: the staged model kernels alone (CUDA events), the whole callable per call (K = 1) and
per parameter point in batches.  CPU: the reference's model() as restated bit for bit by
oracle.np.muse_model (numpy, one thread) + the reference's own cmuselike.so (serial and OpenMP).

    python tools/bench_muse_model.py [--out gpurun_out/muse_model.json]
"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import synth  # noqa: E402
from massivedatans_b200.likelihood import calzetti, make_muse_loglikelihood_device  # noqa: E402
from oracle import np as onp  # noqa: E402
from oracle import ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ndata', type=int, default=synth.MUSE_NDATA)
    ap.add_argument('--nwave', type=int, default=6900)
    ap.add_argument('--reps', type=int, default=200)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'muse_model.json'))
    args = ap.parse_args()
    nspec = synth.MUSE_NSPEC
    y, v, _ = synth.muse(ndata=args.ndata, nspec=nspec)
    Zs, ages, wl_A, grids = synth.muse_grids(nwave=args.nwave)
    mw, wl = wl_A / 10., synth.muse_wavelength(nspec) / 10.
    f = make_muse_loglikelihood_device(y, v, grids, Zs, ages, mw, wl, jitter=0)
    ds, model = f.dataset, f.model
    mask = numpy.ones(args.ndata, dtype=bool)
    res = {'ndata': args.ndata, 'nspec': nspec, 'grids': list(grids.shape),
           'grid_bytes': int(grids.nbytes)}
    # (1) model kernels alone
    for K in (1, 16, 64):
        p = synth.muse_parameter_points(K, seed=K)
        p[:, 1] = 10 ** p[:, 1]
        for _ in range(5):
            model.stage(p)
        ds.sync()
        ds.timer_start()
        for _ in range(args.reps):
            model._lib.mdns_muse_model_stage(model._h, p.ctypes.data, K, None)
        ms = ds.timer_stop() / args.reps
        res['model_stage_K%d' % K] = {'ms': ms, 'us_per_point': 1e3 * ms / K,
                                      'template_sum_l2_gbs': K * (grids.shape[1] - 1) * args.nwave * 8 / ms / 1e6}
    # (2) the callable, K = 1 per call
    pts = synth.muse_parameter_points(64, seed=9)
    for q in pts[:5]:
        f(q, mask)
    t0 = time.perf_counter()
    for i in range(args.reps):
        f(pts[i % 64], mask)
    res['callable_ms_per_call'] = 1e3 * (time.perf_counter() - t0) / args.reps
    for K in (16, 64):
        f.batch(pts[:K], mask)
        t0 = time.perf_counter()
        n = max(3, args.reps // K)
        for _ in range(n):
            f.batch(pts[:K], mask)
        res['batch_K%d_ms_per_point' % K] = 1e3 * (time.perf_counter() - t0) / n / K
    # (3) CPU: model restatement + reference cmuselike.so
    cz = calzetti(mw)
    n = 20
    t0 = time.perf_counter()
    for i in range(n):
        q = pts[i]
        spec = onp.muse_model(Zs, ages, mw, cz, grids, wl, q[0], 10 ** q[1], q[2], q[3], q[4])
    t_model = (time.perf_counter() - t0) / n
    Lout = numpy.zeros(args.ndata)
    t0 = time.perf_counter()
    for i in range(5):
        ref.cmuselike(y, v, spec, mask, Lout=Lout)
    t_like = (time.perf_counter() - t0) / 5
    cpu = {'model_numpy_ms': 1e3 * t_model, 'cmuselike_serial_ms': 1e3 * t_like,
           'call_serial_ms': 1e3 * (t_model + t_like), 'host_cores': os.cpu_count()}
    try:
        os.environ['OMP_NUM_THREADS'] = str(os.cpu_count())
        ref.cmuselike(y, v, spec, mask, Lout=Lout, parallel=True)
        t0 = time.perf_counter()
        for i in range(10):
            ref.cmuselike(y, v, spec, mask, Lout=Lout, parallel=True)
        t_par = (time.perf_counter() - t0) / 10
        cpu['cmuselike_openmp_ms'] = 1e3 * t_par
        cpu['call_openmp_ms'] = 1e3 * (t_model + t_par)
    except Exception as e:      # the OpenMP build is optional
        cpu['cmuselike_openmp_ms'] = None
        cpu['note'] = str(e)
    res['cpu'] = cpu
    # parity of the last point, for the record
    got = f(pts[n - 1], mask)
    want = ref.cmuselike(y, v, spec, mask)[mask]
    res['rel_err_vs_cpu'] = float(numpy.max(numpy.abs(got - want) / numpy.abs(want)))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
