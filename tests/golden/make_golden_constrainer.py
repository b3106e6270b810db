#!/usr/bin/env python
"""Generate tests/golden/constrainer.npz: a fixed-seed sequence of constrained draws made by the
REFERENCE's own hiermetriclearn.MetricLearningFriendsConstrainer (+ clustering/radfriendsregion.py,
clustering/sdml.py, clustering/neighbors.py on the reference cneighbors.so) with the reference
clike.so as the likelihood (oracle/_ref).  Build container only; nothing of the reference is
copied (modules are imported from a temporary symlink farm, see make_golden_sampler.py, whose
stubs for the absent matplotlib etc. and Python-3 placeholder for `prev_maxdistance` are reused).

The driving loop is tests/harness_constrainer.py::run_draws; the device mirror
(massivedatans_b200/hiermetriclearn.py) must reproduce every draw: u bit for bit, the number of
tries exactly, logL within 1e-9.

    python tests/golden/make_golden_constrainer.py
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, HERE)

CASES = {
    # tag: (ndata, nlive, niter, data seed, run seed, constructor arguments[, threshold rank])
    'a': (16, 40, 600, 21, 3, dict(metriclearner='truncatedscaling', force_shrink=True,
                                   rebuild_every=30, metric_rebuild_every=5)),
    'b': (6, 60, 100, 22, 4, dict(metriclearner='simplescaling', force_shrink=False,
                                  rebuild_every=1000, metric_rebuild_every=20)),
    'c': (10, 30, 400, 23, 5, dict(metriclearner='none', force_shrink=True,
                                  rebuild_every=8, metric_rebuild_every=3)),
    # thresholds at the BEST live point of every data set: hundreds of tries per draw, the
    # `ntoaccept > 200` metric rebuild inside the loop (hiermetriclearn.py:206-211)
    'd': (4, 30, 24, 24, 6, dict(metriclearner='truncatedscaling', force_shrink=True,
                                 rebuild_every=1000, metric_rebuild_every=20), -1),
}


def main():
    from make_golden_sampler import _NoPrevious, link_farm, stub_modules
    from harness_constrainer import run_draws
    from massivedatans_b200 import synth
    from oracle import ref
    os.environ.pop('OMP_NUM_THREADS', None)
    stub_modules()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        link_farm(tmp)
        sys.path.insert(0, tmp)
        from hiermetriclearn import MetricLearningFriendsConstrainer
        import clustering.neighbors as nb
        assert nb.bootstrapped_maxdistance is not None, 'reference cneighbors.so did not load'
        for tag, case in CASES.items():
            ndata, nlive, niter, seed_data, seed_run, kw = case[:6]
            rank = case[6] if len(case) > 6 else 0
            x, y, _ = synth.horns(ndata, seed=seed_data)
            ncalls = [0]

            def multi_loglikelihood(params, data_mask):     # sample.py:101-108 on clike.so
                A, mu, log_sig = params
                Lout = numpy.zeros(int(data_mask.sum()))
                ref.clike(x, y, A, mu, 10 ** log_sig, synth.NOISE_LEVEL,
                          numpy.ascontiguousarray(data_mask), Lout=Lout)
                ncalls[0] += 1
                return -0.5 * Lout

            c = MetricLearningFriendsConstrainer(verbose=False, **kw)
            c.prev_maxdistance = _NoPrevious()
            sink = io.StringIO()
            with contextlib.redirect_stdout(sink):
                res = run_draws(c, multi_loglikelihood, ndata, nlive, niter, seed_run, rank=rank)
            for k, v in res.items():
                out['%s_%s' % (tag, k)] = v
            out[tag + '_cfg'] = numpy.array([ndata, nlive, niter, seed_data, seed_run])
            out[tag + '_rank'] = rank
            out[tag + '_ncalls'] = ncalls[0]
            out[tag + '_maxdistance'] = c.region.maxdistance
            print(tag, 'draws', niter, 'tries', res['ntoaccept'].sum(), 'max tries',
                  res['ntoaccept'].max(), 'likelihood calls', ncalls[0],
                  'final radius', c.region.maxdistance,
                  'rebuilds at call / in loop / metric in loop',
                  sink.getvalue().count('rebuild triggered at call'),
                  sink.getvalue().count('RadFriends rebuild triggered'),
                  sink.getvalue().count('RadFriends metric rebuild triggered'))
        sys.path.remove(tmp)
    path = os.path.join(HERE, 'constrainer.npz')
    numpy.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    main()
