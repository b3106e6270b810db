#!/usr/bin/env python
"""Constrained draws end to end at BASELINE configs[2] scale (gensimple_horns, 100 000 data sets,
400 live points, one B200): wall time of ``draw_constrained`` -- region build (bootstrapped
radius), candidate generation (neighbour kernels), likelihood passes, accept test -- for

* the device constrainer one candidate per pass (the reference's calling pattern),
* the device constrainer with speculative batches (--batch, default 16; fixed width and the
  adaptive default that starts every draw with one candidate), and
* the CPU arm: the SAME host logic (massivedatans_b200/hiermetriclearn.py is a draw-by-draw
  mirror of the reference class, tests/test_constrainer.py) on the reference's own unmodified C
  libraries (oracle/_ref clike.so + cneighbors.so, serial builds as sample.py / neighbors.py load
  them without OMP_NUM_THREADS), for the first --cpu-draws draws of the same seeded run.

All three make identical draws (checked); only the time differs.  The live-point table
bookkeeping between the draws is the sampler's (host numpy here for every arm) and is not timed.

    python tools/bench_constrainer.py [--ndata 100000] [--nlive 400] [--draws 200] [--out ...]
"""
import argparse
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.clustering.radfriendsregion import RadFriendsRegion  # noqa: E402
from massivedatans_b200.hiermetriclearn import MetricLearningFriendsConstrainer  # noqa: E402
from massivedatans_b200.likelihood import make_multi_loglikelihood  # noqa: E402
from harness_constrainer import NDIM, group_mask, priortransform  # noqa: E402

CONSTRAINER = dict(metriclearner='truncatedscaling', force_shrink=True, rebuild_every=1000,
                   metric_rebuild_every=20)         # sample.py:133-137


class Timed(object):
    """Wraps a callable, accumulating wall time and calls."""

    def __init__(self, fn):
        self.fn, self.t, self.n = fn, 0.0, 0

    def __call__(self, *a, **k):
        t0 = time.perf_counter()
        try:
            return self.fn(*a, **k)
        finally:
            self.t += time.perf_counter() - t0
            self.n += 1


def run(constrainer, like, init_L, pile0, ndata, nlive, ndraws, seed, groups='cycle', rank=0):
    """The loop of tests/harness_constrainer.py::run_draws with vectorised bookkeeping and a
    timer around draw_constrained."""
    numpy.random.seed(seed)
    rs = numpy.random.RandomState(seed + 1000)
    pile = list(pile0)
    live_L = init_L.copy()
    live_p = numpy.repeat(numpy.arange(nlive)[:, None], ndata, axis=1)
    worst = live_L.argmin(axis=0)
    cols = numpy.arange(ndata)
    us, tries, per_draw = [], [], []
    for it in range(ndraws):
        if groups == 'single':
            mask = numpy.zeros(ndata, dtype=bool)
            mask[rs.randint(ndata)] = True
        else:
            mask = group_mask(it, ndata, rs)
        members = numpy.unique(live_p[:, mask])
        live_u = numpy.array([pile[i] for i in members])
        if rank == 0:
            Lmins = live_L[worst, cols][mask]
        else:
            Lmins = numpy.sort(live_L[:, mask], axis=0)[rank]
        t0 = time.perf_counter()
        u, x, L, n = constrainer.draw_constrained(
            Lmins=Lmins, priortransform=priortransform,
            loglikelihood=lambda p: like(p, mask), live_pointsu=live_u, ndim=NDIM)
        per_draw.append(time.perf_counter() - t0)
        L = numpy.asarray(L)
        idx = numpy.where(mask)[0][L > Lmins]
        pile.append(numpy.array(u))
        live_p[worst[idx], idx] = len(pile) - 1
        live_L[worst[idx], idx] = L[L > Lmins]
        worst[idx] = live_L[:, idx].argmin(axis=0)
        us.append(u)
        tries.append(n)
    return numpy.array(us), numpy.array(tries), numpy.array(per_draw)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ndata', type=int, default=100000)
    ap.add_argument('--nlive', type=int, default=400)
    ap.add_argument('--draws', type=int, default=200)
    ap.add_argument('--cpu-draws', type=int, default=40)
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--groups', choices=('cycle', 'single'), default='cycle',
                    help='cycle: all, all, random half, single data set; single: focussed draws '
                         'for one data set at a time (the late-stage regime of the sampler)')
    ap.add_argument('--rank', type=int, default=0,
                    help='which live likelihood is the threshold: 0 = minimum (nested sampling), '
                         '-1 = maximum (long rejection chains)')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'constrainer.json'))
    args = ap.parse_args()
    N, nlive = args.ndata, args.nlive
    x, y, _ = synth.horns(N, legacy=False, seed=3)
    like = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
    # initial population (multi_nested_sampler.py:91-103): one batched launch
    rs = numpy.random.RandomState(args.seed + 7)
    pile0 = [rs.uniform(size=NDIM) for _ in range(nlive)]
    t0 = time.perf_counter()
    init_L = numpy.array(like.batch([priortransform(u) for u in pile0], numpy.ones(N, dtype=bool)))
    t_init = time.perf_counter() - t0
    res = {'ndata': N, 'nlive': nlive, 'nx': int(y.shape[0]), 'draws': args.draws,
           'constrainer': CONSTRAINER, 'threshold_rank': args.rank,
           'groups': 'all, all, random half, single data set (cyclic)' if args.groups == 'cycle'
           else 'one data set per draw',
           'initial_population_gpu_s': t_init}
    lib = _lib.load()
    arms, pers = {}, {}
    for name, batch, adaptive, devprop in (('gpu_one_by_one', 1, False, 0),
                                           ('gpu_speculative', args.batch, False, 0),
                                           ('gpu_speculative_adaptive', args.batch, True, 0),
                                           # statistical mode: proposals from the fused device
                                           # generator (not numpy's stream -> different draws)
                                           ('gpu_speculative_device_proposals', 4 * args.batch, True, 8192)):
        tl = Timed(like)
        c = MetricLearningFriendsConstrainer(batch_size=batch, speculator=like if batch > 1 else None,
                                             adaptive=adaptive, device_proposals=devprop, **CONSTRAINER)
        l0 = lib.mdns_launch_count()
        us, tries, per = run(c, tl, init_L, pile0, N, nlive, args.draws, args.seed, args.groups,
                             args.rank)
        arms[name] = (us, tries)
        pers[name] = per
        res[name] = {'batch': batch, 'draw_s_total': float(per.sum()), 'ms_per_draw': 1e3 * float(per.mean()),
                     'ms_per_draw_median': 1e3 * float(numpy.median(per)),
                     'tries': int(tries.sum()), 'max_tries': int(tries.max()),
                     'likelihood_passes': int(c.nbatches), 'candidates_scored': int(c.nscored),
                     'likelihood_s': tl.t, 'kernel_launches': int(lib.mdns_launch_count() - l0)}
    assert numpy.array_equal(arms['gpu_one_by_one'][0], arms['gpu_speculative'][0])
    assert numpy.array_equal(arms['gpu_one_by_one'][1], arms['gpu_speculative'][1])
    assert numpy.array_equal(arms['gpu_one_by_one'][0], arms['gpu_speculative_adaptive'][0])
    assert numpy.array_equal(arms['gpu_one_by_one'][1], arms['gpu_speculative_adaptive'][1])
    res['identical_draws_gpu_arms'] = True
    if args.cpu_draws > 0:
        from oracle import ref

        class RefMembers(object):
            def __init__(self, members, device=0):
                self.set(members)

            def set(self, members):
                self.xx = numpy.ascontiguousarray(members, dtype=numpy.float64)
                self.n, self.ndim = self.xx.shape

            def counts(self, r, us, countmax):
                out = numpy.zeros(len(us))
                ref.count_within_distance_of_raw(self.xx, r, numpy.ascontiguousarray(us), out, countmax)
                return out

            def is_within(self, r, u):
                return bool(ref.is_within_distance_of(self.xx, r, numpy.ascontiguousarray(u)))

            def bootstrapped_maxdistance(self, nbootstraps):
                chosen = numpy.zeros((self.n, nbootstraps))
                for b in range(nbootstraps):
                    chosen[numpy.random.choice(numpy.arange(self.n), size=self.n, replace=True), b] = 1.
                return ref.bootstrapped_maxdistance_chosen(self.xx, chosen)

        class RefRegion(RadFriendsRegion):
            members_class = RefMembers

        def ref_like(params, mask):
            A, mu, log_sig = params
            Lout = numpy.zeros(int(mask.sum()))
            ref.clike(x, y, A, mu, 10 ** log_sig, synth.NOISE_LEVEL, numpy.ascontiguousarray(mask), Lout=Lout)
            return -0.5 * Lout

        tl = Timed(ref_like)
        c = MetricLearningFriendsConstrainer(batch_size=1, region_class=RefRegion, **CONSTRAINER)
        us, tries, per = run(c, tl, init_L, pile0, N, nlive, args.cpu_draws, args.seed, args.groups,
                             args.rank)
        n = args.cpu_draws
        same = numpy.array_equal(us, arms['gpu_one_by_one'][0][:n]) and \
            numpy.array_equal(tries, arms['gpu_one_by_one'][1][:n])
        res['cpu_reference_libs'] = {'draws': n, 'draw_s_total': float(per.sum()),
                                     'ms_per_draw': 1e3 * float(per.mean()),
                                     'ms_per_draw_median': 1e3 * float(numpy.median(per)),
                                     'tries': int(tries.sum()), 'likelihood_s': tl.t,
                                     'identical_draws_to_gpu': bool(same), 'threads': 1}
        for name in pers:
            # the same first n draws on the device arms
            res[name]['ms_per_draw_first_%d' % n] = 1e3 * float(pers[name][:n].mean())
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
