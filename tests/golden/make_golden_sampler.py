#!/usr/bin/env python
"""Generate tests/golden/sampler_run.npz: evidences of a fixed-seed run of the REFERENCE's own
sampler stack (multi_nested_integrator.py + multi_nested_sampler.py + cachedconstrainer.py +
hiermetriclearn.py + clustering/*.py) on top of the REFERENCE's own C (oracle/_ref/*.so).

Only runs where /root/reference exists (this container); the fixture travels.  The script
restates what sample.py:27-31,44-58,101-108,131-197 does, because sample.py itself is a script
that needs h5py and a data file.  Nothing of the reference is copied: its modules are imported
from a symlink farm under a temporary directory (the C libraries must sit beside
clustering/neighbors.py, neighbors.py:97-98, and /root/reference is read-only).

Stubs for packages that are absent here: progressbar (no-op), igraph
(unused with USE_GRAPH=0), nestle and matplotlib (imported, never called on this path).
Python-3 guard: hiermetriclearn.py:53 compares a float with
`prev_maxdistance = None`; Python 2 evaluated that, Python 3 raises.  The instances get a
`prev_maxdistance` placeholder whose reflected comparison answers False, i.e. the
`force_shrink` branch is skipped on the first region, which is the guard the survey describes.

    python tests/golden/make_golden_sampler.py
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

NDATA = 12
NLIVE = 100
SEED_DATA = 12
SEED_RUN = 1            # sample.py:162


def stub_modules():
    pb = types.ModuleType('progressbar')

    class _W(object):
        def __init__(self, *a, **k):
            pass

    class ProgressBar(object):
        def __init__(self, widgets=None, maxval=None, **k):
            self.maxval = maxval
            self.currval = 0
            self.finished = False
            self.seconds_elapsed = 0.0

        def start(self):
            return self

        def update(self, i=None):
            if i is not None:
                self.currval = i

        def finish(self):
            self.finished = True

    for name in ('Counter', 'Bar', 'Percentage', 'Timer', 'ETA', 'SimpleProgress'):
        setattr(pb, name, type(name, (_W,), {}))
    pb.ProgressBar = ProgressBar
    sys.modules['progressbar'] = pb
    sys.modules['igraph'] = types.ModuleType('igraph')
    nestle = types.ModuleType('nestle')
    for name in ('bounding_ellipsoid', 'bounding_ellipsoids', 'sample_ellipsoids'):
        setattr(nestle, name, None)
    sys.modules['nestle'] = nestle
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = plt


def link_farm(tmp):
    for name in os.listdir(REF):
        if name.endswith('.py'):
            os.symlink(os.path.join(REF, name), os.path.join(tmp, name))
    os.mkdir(os.path.join(tmp, 'clustering'))
    for name in os.listdir(os.path.join(REF, 'clustering')):
        if name.endswith('.py'):
            os.symlink(os.path.join(REF, 'clustering', name), os.path.join(tmp, 'clustering', name))
    refdir = os.path.join(ROOT, 'oracle', '_ref')
    os.symlink(os.path.join(refdir, 'cneighbors.so'), os.path.join(tmp, 'clustering', 'cneighbors.so'))


class _NoPrevious(object):
    """prev_maxdistance placeholder: `maxdistance > placeholder` is False."""

    def __lt__(self, other):
        return False

    def __gt__(self, other):
        return False


def main():
    from massivedatans_b200 import synth
    from oracle import ref
    os.environ['USE_GRAPH'] = '0'
    os.environ.pop('OMP_NUM_THREADS', None)
    stub_modules()
    x, y, truth = synth.horns(NDATA, seed=SEED_DATA)
    nx, ndata = y.shape
    noise_level = synth.NOISE_LEVEL

    def priortransform(cube):               # sample.py:52-58
        cube = cube.copy()
        cube[0] = 10 ** (cube[0] * 2 - 2)
        cube[1] = cube[1] * 400 + 400
        cube[2] = cube[2] * 2
        return cube

    def multi_loglikelihood(params, data_mask):     # sample.py:101-108 on the reference clike.so
        A, mu, log_sig = params
        Lout = numpy.zeros(int(data_mask.sum()))
        ref.clike(x, y, A, mu, 10 ** log_sig, noise_level, numpy.ascontiguousarray(data_mask), Lout=Lout)
        return -0.5 * Lout

    with tempfile.TemporaryDirectory() as tmp:
        link_farm(tmp)
        sys.path.insert(0, tmp)
        import cachedconstrainer
        from cachedconstrainer import (CachedConstrainer, MetricLearningFriendsConstrainer,
                                       generate_individual_constrainer)
        from multi_nested_integrator import multi_nested_integrator
        from multi_nested_sampler import MultiNestedSampler
        import clustering.neighbors as nb
        assert nb.bootstrapped_maxdistance is not None, 'reference cneighbors.so did not load'

        def fresh():                        # sample.py:133-137
            c = MetricLearningFriendsConstrainer(metriclearner='truncatedscaling', force_shrink=True,
                                                 rebuild_every=1000, metric_rebuild_every=20,
                                                 verbose=False)
            c.prev_maxdistance = _NoPrevious()
            return c

        cachedconstrainer.generate_fresh_constrainer = fresh
        superset = fresh()
        cc = CachedConstrainer()
        _, _, individual_draw_constrained = generate_individual_constrainer()
        numpy.random.seed(SEED_RUN)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            sampler = MultiNestedSampler(nlive_points=NLIVE, priortransform=priortransform,
                                         multi_loglikelihood=multi_loglikelihood, ndim=3,
                                         ndata=ndata,
                                         superset_draw_constrained=superset.draw_constrained,
                                         individual_draw_constrained=individual_draw_constrained,
                                         draw_constrained=cc.get, nsuperset_draws=10,
                                         use_graph=False)
            superset.sampler = sampler
            cc.sampler = sampler
            results = multi_nested_integrator(tolerance=0.5, multi_sampler=sampler,
                                              min_samples=0, max_samples=0)
        sys.path.remove(tmp)
    logZ = numpy.asarray(results['logZ'], dtype=float)
    logZerr = numpy.asarray(results['logZerr'], dtype=float)
    # analytic no-signal value for orientation (plotevidences.py:17)
    null = (-0.5 * (y / noise_level) ** 2).sum(axis=0)
    # posterior moments per data set from the weighted samples, with the weights
    # plotposterior.py:21-27 uses (logwidth + L, normalised); x = (A, mu, log10 sig)
    u, xs, L, w, mask = [numpy.asarray(a, dtype=float) for a in zip(*results['weights'])]
    lw = w + L                                      # [nsamples, ndata]
    lw[~numpy.isfinite(lw)] = -numpy.inf
    lw -= lw.max(axis=0)
    p = numpy.exp(lw)
    p /= p.sum(axis=0)
    feat = xs.copy()                                # [nsamples, ndata, 3]
    feat[:, :, 0] = numpy.log10(numpy.where(feat[:, :, 0] > 0, feat[:, :, 0], 1.0))
    post_mean = (p[:, :, None] * feat).sum(axis=0)
    post_std = numpy.sqrt((p[:, :, None] * (feat - post_mean) ** 2).sum(axis=0))
    post_ess = 1.0 / (p ** 2).sum(axis=0)
    out = os.path.join(HERE, 'sampler_run.npz')
    numpy.savez(out, post_mean=post_mean, post_std=post_std, post_ess=post_ess, ndata=NDATA, nlive=NLIVE, seed_data=SEED_DATA, seed_run=SEED_RUN,
                logZ=logZ, logZerr=logZerr, ndraws=int(sampler.ndraws),
                niterations=int(results['niterations']), null_logZ=null,
                information=numpy.asarray(results['information'], dtype=float))
    print('wrote', out)
    for d in range(ndata):
        print('data set %2d  logZ %10.3f +- %.3f   (null %10.3f, line height %.4f)'
              % (d, logZ[d], logZerr[d], null[d], truth['height_narrow'][d]))
    print('ndraws', sampler.ndraws, 'niterations', results['niterations'])
    print('posterior mu mean', post_mean[:, 1].round(2), 'std', post_std[:, 1].round(2), 'ess', post_ess.round(0))


if __name__ == '__main__':
    main()
