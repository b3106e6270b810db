"""MLFriends constrainer on the device path -- mirror of the reference's hiermetriclearn.py:30-211.

``MetricLearningFriendsConstrainer`` keeps the reference's constructor, ``draw_constrained``
signature and return value ``(u, x, L, ntoaccept)`` (hiermetriclearn.py:173-196), so
multi_nested_sampler.py:462-472 and cachedconstrainer.py:19-114 call it unchanged.  With the same
numpy seed it reproduces the reference draw by draw: the proposals come from the same
``numpy.random`` calls in the same order (region: clustering/radfriendsregion.py here, a bit-exact
mirror; metric: clustering/sdml.py), the neighbour tests are bit-exact device kernels.  The one
qualification: logL agrees with clike.c to rounding (1e-13 in the direct form, within the enforced
1e-10 in the expanded forms that speculative batches of >= 3 candidates take), so an accept test
``L > Lmins`` can come out differently only where |L - Lmins| is below that -- a tie that no
seeded run in tests/golden/constrainer.npz (1124 draws, 22 000 tries) contains.

What is new is *speculation*.  The reference scores one candidate per likelihood call and stops
at the first with ``numpy.any(L > Lmins)`` (hiermetriclearn.py:181-196).  The candidates of one
proposal round exist before any of them is scored (hiermetriclearn.py:112-123 yields them out of
an array), so the next ``batch_size`` of them can be scored in ONE pass over the resident data
(K x n_act evaluations for one read of the matrix) and the accept test run on the device.  The
result is the same as the one-by-one loop: the first accepted candidate in order, with
``ntoaccept`` counting the candidates up to it; candidates behind it stay queued for the next
draw exactly as they stay inside the reference's generator; speculation never crosses the end of
a proposal round, so no random number is drawn earlier than the reference would draw it.

The sampler hides the data-set mask in a lambda (multi_nested_sampler.py:465), so the batch goes
through the likelihood callable itself: ``speculator.speculate(xs, Lmins)`` announces the batch,
the next ``loglikelihood(xs[0])`` call carries the mask and runs it, ``speculator.last_draw``
holds ``(k, L_k, counts)``.  ``massivedatans_b200.likelihood.make_multi_loglikelihood`` returns
such a callable.  Without a speculator (or with ``batch_size=1``) every candidate is scored by
its own call, as in the reference.  ``device_proposals=m`` (off by default) switches the proposal
rounds to the fused device generator, m proposals per round: statistically equivalent draws, not
the reference's random stream.  ``adaptive`` (default) starts every draw with one candidate
and doubles the pass width after each fully rejected pass, so an easy draw costs what it costs
in the reference and a long rejection chain of n candidates takes about log2(n) passes.

Python-3 note: hiermetriclearn.py:53 compares ``maxdistance`` with
``prev_maxdistance = None``; here "no previous radius" skips the ``force_shrink`` branch.
"""
import numpy

from .clustering.radfriendsregion import RadFriendsRegion
from .clustering.sdml import IdentityMetric, SimpleScaling, TruncatedScaling

UNIT_CUBE_DRAWS = 10000          # hiermetriclearn.py:106
UNIT_CUBE_CHANCE = 0.1           # hiermetriclearn.py:126


class MetricLearningFriendsConstrainer(object):
    def __init__(self, metriclearner, rebuild_every=50, metric_rebuild_every=50, verbose=False,
                 keep_phantom_points=False, optimize_phantom_points=False, force_shrink=False,
                 batch_size=16, speculator=None, adaptive=True, device_proposals=0,
                 region_class=RadFriendsRegion):
        if metriclearner not in ('none', 'simplescaling', 'truncatedscaling'):
            raise ValueError('unknown metriclearner %r' % (metriclearner,))
        self.iter_since_metric_rebuild = 0
        self.ndraws_since_rebuild = 0
        self.region = None
        self.rebuild_every = int(rebuild_every)
        self.metric_rebuild_every = int(metric_rebuild_every)
        self.verbose = verbose
        self.force_shrink = force_shrink
        self.metriclearner = metriclearner
        self.metric = IdentityMetric()
        self.clusters = None
        self.direct_draws_efficient = True
        self.last_cluster_points = None
        self.prev_maxdistance = None
        self.generator = None
        self.batch_size = max(1, int(batch_size))
        self.speculator = speculator
        self.adaptive = bool(adaptive)
        self.device_proposals = int(device_proposals)    # proposals per device round, 0 = host RNG
        self.region_class = region_class
        self._queue = []             # candidates of the current proposal round, not yet scored
        self.nbatches = 0            # likelihood passes issued
        self.nscored = 0             # candidates scored (>= the reference's count: speculation)

    def _say(self, *args):
        if self.verbose:
            print(*args)

    # -- region ------------------------------------------------------------------------------
    def _learn_metric(self, u):
        """hiermetriclearn.py:64-83 -> (metric, metric_updated)"""
        centred = u - numpy.mean(u, axis=0)
        if self.metriclearner == 'none':
            return self.metric, False
        if self.metriclearner == 'simplescaling':
            metric = SimpleScaling()
            metric.fit(centred)
            return metric, True
        metric = TruncatedScaling()
        metric.fit(centred)
        changed = self.metric == IdentityMetric() or not numpy.all(self.metric.scale == metric.scale)
        return metric, changed

    def _shrunk(self, region, w):
        """force_shrink: never let the radius grow while the metric stands
        (hiermetriclearn.py:52-54, 87-89)."""
        if self.force_shrink and self.prev_maxdistance is not None \
                and region.maxdistance > self.prev_maxdistance:
            return self.region_class(members=w, maxdistance=self.prev_maxdistance)
        return region

    def cluster(self, u, ndim, keepMetric=False):
        w = self.metric.transform(u)
        if keepMetric:
            self.region = self._shrunk(self.region_class(members=w), w)
            self.prev_maxdistance = self.region.maxdistance
            return
        metric, metric_updated = self._learn_metric(u)
        self.metric = metric
        region = self.region_class(members=self.metric.transform(u))
        if not metric_updated:
            # like the reference (hiermetriclearn.py:89) the shrunk region gets the members
            # under the PREVIOUS metric: same scale, but its own (rounding-sized) mean
            region = self._shrunk(region, w)
        self.region = region
        self.prev_maxdistance = self.region.maxdistance

    def are_inside_cluster(self, points):
        return self.region.are_inside(self.metric.transform(points))

    def is_inside(self, point):
        if not ((point >= 0).all() and (point <= 1).all()):
            return False
        return self.region.is_inside(self.metric.transform(point))

    # -- proposals ---------------------------------------------------------------------------
    def _rounds(self, ndim):
        """The proposal rounds of hiermetriclearn.py:104-137, one array of unit-cube candidates
        per round: ``(us[k, ndim], proposals spent since the previous non-empty round)``.  The
        reference yields the rows of each array one at a time, the first carrying the count."""
        spent = 0
        if self.device_proposals > 0:
            # statistical mode: the ball draws, the neighbour count and
            # the 1/count thinning happen in one kernel (RadFriendsRegion.generate_device); only
            # the accepted points come back.  Same distribution (uniform in the region), not
            # numpy's random stream: one numpy draw per region seeds the device stream, so a run
            # is still reproducible from the numpy seed, but it is not the reference's run.
            seed = int(numpy.random.randint(0, 2 ** 31 - 1))
            first = 0
            while True:
                ws = self.region.generate_device(self.device_proposals, seed, first_proposal=first)
                first += self.device_proposals
                spent += self.device_proposals
                us = self.metric.untransform(ws)
                inside = numpy.logical_and(us < 1, us > 0).all(axis=1)
                if inside.any():
                    yield us[inside, :], spent
                    spent = 0
        while True:
            if ndim < 40:
                for ws, n in self.region.generate(UNIT_CUBE_DRAWS):
                    us = self.metric.untransform(ws)
                    spent += n
                    inside = numpy.logical_and(us < 1, us > 0).all(axis=1)
                    if inside.any():
                        yield us[inside, :], spent
                        spent = 0
            if numpy.random.uniform() < UNIT_CUBE_CHANCE:
                spent += UNIT_CUBE_DRAWS
                us = numpy.random.uniform(size=(UNIT_CUBE_DRAWS, ndim))
                inside = self.region.are_inside(self.metric.transform(us))
                if inside.any():
                    yield us[inside, :], spent
                    spent = 0

    def generate(self, ndim):
        """Candidate by candidate, as hiermetriclearn.py:104-137: yields ``(u, ntotal)``."""
        for us, spent in self._rounds(ndim):
            for u in us:
                yield u, spent
                spent = 0

    def _peek(self, k):
        """Up to k queued candidates; a new proposal round is drawn only when none is left,
        which is when the reference's generator would resume."""
        if not self._queue:
            us, spent = next(self.generator)
            self._queue = [(u, spent if i == 0 else 0) for i, u in enumerate(us)]
        return self._queue[:k]

    def rebuild(self, u, ndim, keepMetric=False):
        if self.last_cluster_points is not None and len(self.last_cluster_points) == len(u) \
                and numpy.all(self.last_cluster_points == u):
            return                       # hiermetriclearn.py:140-144: same points, same region
        self.cluster(u=u, ndim=ndim, keepMetric=keepMetric)
        self.last_cluster_points = u
        self._say('maxdistance:', self.region.maxdistance)
        self.generator = self._rounds(ndim)
        self._queue = []

    def _draw_constrained_prepare(self, Lmins, priortransform, loglikelihood, live_pointsu, ndim,
                                  **kwargs):
        rebuild = self.ndraws_since_rebuild > self.rebuild_every or self.region is None
        rebuild_metric = self.iter_since_metric_rebuild > self.metric_rebuild_every
        if rebuild:
            self.rebuild(numpy.asarray(live_pointsu), ndim, keepMetric=not rebuild_metric)
            self.ndraws_since_rebuild = 0
            if rebuild_metric:
                self.iter_since_metric_rebuild = 0
        else:
            rebuild_metric = False
        assert self.generator is not None
        return rebuild, rebuild_metric

    # -- the draw ----------------------------------------------------------------------------
    def _score(self, batch, xs, Lmins, loglikelihood):
        """Index of the first accepted candidate of the batch (or -1) and its logL vector."""
        self.nbatches += 1
        self.nscored += len(batch)
        if self.speculator is not None and len(batch) > 1:
            self.speculator.speculate(xs, Lmins)
            loglikelihood(xs[0])
            if self.speculator.last_draw is None:
                raise RuntimeError('the speculator did not see the announced batch: '
                                   '`loglikelihood` must call the callable given as `speculator`')
            k, L, _ = self.speculator.last_draw
            return k, L
        for k, x in enumerate(xs):
            L = loglikelihood(x)
            if numpy.any(L > Lmins):
                return k, L
        return -1, None

    def draw_constrained(self, Lmins, priortransform, loglikelihood, live_pointsu, ndim, **kwargs):
        ntoaccept = 0
        self.iter_since_metric_rebuild += 1
        rebuild, rebuild_metric = self._draw_constrained_prepare(
            Lmins, priortransform, loglikelihood, live_pointsu, ndim, **kwargs)
        speculative = self.speculator is not None and self.batch_size > 1
        # adaptive: most draws are accepted at the first try, so the first pass scores one
        # candidate; every fully rejected pass doubles the next one up to batch_size
        width = 1 if self.adaptive else self.batch_size
        while True:
            batch = self._peek(min(width, self.batch_size) if speculative else 1)
            width *= 2
            cube = numpy.array([u for u, _ in batch])
            assert (cube >= 0).all() and (cube <= 1).all(), cube       # hiermetriclearn.py:182
            xs = [priortransform(u) for u, _ in batch]
            k, L = self._score(batch, xs, Lmins, loglikelihood)
            # replay the reference's per-candidate bookkeeping (hiermetriclearn.py:181-211) up
            # to the accepted candidate; later candidates stay queued
            last = k if k >= 0 else len(batch) - 1
            for i in range(last + 1):
                ntoaccept += 1
                self.ndraws_since_rebuild += 1
                if batch[i][1] > 100000:
                    self.direct_draws_efficient = False
                if i == k:
                    del self._queue[:i + 1]
                    return batch[i][0], xs[i], L, ntoaccept
                # a rebuild inside the loop (hiermetriclearn.py:200-211): candidates 0..i are
                # spent; `rebuild` empties the queue when it really builds a new region, and
                # leaves the rest queued when the live points are the ones the region was built
                # from (the reference then resumes the same generator)
                if not rebuild and self.ndraws_since_rebuild > self.rebuild_every:
                    rebuild = True
                    self._say('RadFriends rebuild triggered after %d draws' % self.ndraws_since_rebuild)
                    del self._queue[:i + 1]
                    self.rebuild(numpy.asarray(live_pointsu), ndim, keepMetric=True)
                    self.ndraws_since_rebuild = 0
                    break
                if not rebuild_metric and ntoaccept > 200:
                    rebuild_metric = True
                    self._say('RadFriends metric rebuild triggered after %d draws'
                              % self.ndraws_since_rebuild)
                    del self._queue[:i + 1]
                    self.rebuild(numpy.asarray(live_pointsu), ndim, keepMetric=False)
                    self.iter_since_metric_rebuild = 0
                    break
            else:
                del self._queue[:last + 1]
