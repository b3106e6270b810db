"""MLFriends constrainer mirror (massivedatans_b200/hiermetriclearn.py) against a fixed-seed run
of the REFERENCE's own class on the reference's C libraries (tests/golden/constrainer.npz, made
by tests/golden/make_golden_constrainer.py): every draw must come out the same -- the accepted
point bit for bit, the number of tries exactly, logL within 1e-9 relative -- one candidate at a
time and with speculative batches (which must not change anything but the number of passes).

CPU tests answer the neighbour queries and the likelihood with the oracle (host logic only);
GPU tests run the product path: region members resident on the device, the batch scored by the
candidate-batch kernels with the accept test on the device.
"""
import os

import numpy
import pytest

from harness_constrainer import run_draws
from massivedatans_b200 import synth
from massivedatans_b200.clustering.radfriendsregion import RadFriendsRegion
from massivedatans_b200.hiermetriclearn import MetricLearningFriendsConstrainer

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'constrainer.npz')
CONFIG = {      # constructor arguments of tests/golden/make_golden_constrainer.py
    'a': dict(metriclearner='truncatedscaling', force_shrink=True, rebuild_every=30,
              metric_rebuild_every=5),
    'b': dict(metriclearner='simplescaling', force_shrink=False, rebuild_every=1000,
              metric_rebuild_every=20),
    'c': dict(metriclearner='none', force_shrink=True, rebuild_every=8, metric_rebuild_every=3),
    # thresholds at the best live point: up to 2926 tries per draw, in-loop metric rebuilds
    'd': dict(metriclearner='truncatedscaling', force_shrink=True, rebuild_every=1000,
              metric_rebuild_every=20),
}


@pytest.fixture(scope='module')
def fixture():
    return numpy.load(GOLDEN)


def check(res, g, tag):
    assert numpy.array_equal(res['ntoaccept'], g[tag + '_ntoaccept'])
    assert numpy.array_equal(res['u'], g[tag + '_u'])
    assert numpy.array_equal(res['naccepted'], g[tag + '_naccepted'])
    want = g[tag + '_L']
    drawn = ~numpy.isnan(want)
    assert numpy.array_equal(drawn, ~numpy.isnan(res['L']))
    assert numpy.max(numpy.abs(res['L'][drawn] - want[drawn]) / numpy.abs(want[drawn])) < 1e-9


# ------------------------------------------------------------------ CPU: host logic on the oracle
class OracleMembers(object):
    """The interface of clustering.radfriendsregion.ResidentMembers answered by the CPU oracle."""

    def __init__(self, members, device=0):
        from oracle import port
        self.port = port
        self.set(members)

    def set(self, members):
        self.xx = numpy.ascontiguousarray(members, dtype=numpy.float64)
        self.n, self.ndim = self.xx.shape

    def counts(self, maxdistance, us, countmax):
        out = numpy.zeros(len(us))
        self.port.count_within_distance_of_raw(self.xx, maxdistance,
                                               numpy.ascontiguousarray(us, dtype=numpy.float64),
                                               out, countmax)
        return out

    def is_within(self, maxdistance, u):
        return bool(self.port.is_within_distance_of(self.xx, maxdistance,
                                                    numpy.ascontiguousarray(u, dtype=numpy.float64)))

    def bootstrapped_maxdistance(self, nbootstraps):
        chosen = numpy.zeros((self.n, nbootstraps))          # clustering/neighbors.py:172-174
        for b in range(nbootstraps):
            chosen[numpy.random.choice(numpy.arange(self.n), size=self.n, replace=True), b] = 1.
        return self.port.bootstrapped_maxdistance_chosen(self.xx, chosen)


class OracleRegion(RadFriendsRegion):
    members_class = OracleMembers


class OracleLikelihood(object):
    """sample.py:101-108 on the oracle, with the speculate / last_draw hook of
    massivedatans_b200.likelihood.make_multi_loglikelihood restated on the host."""

    def __init__(self, x, y):
        from oracle import port
        self.port, self.x, self.y = port, x, y
        self.pending = None
        self.last_draw = None
        self.ncalls = 0

    def one(self, params, data_mask):
        A, mu, log_sig = params
        return -0.5 * self.port.clike(self.x, self.y, A, mu, 10 ** log_sig, synth.NOISE_LEVEL,
                                      numpy.ascontiguousarray(data_mask))

    def speculate(self, params_list, Lmins):
        self.pending = ([numpy.array(p) for p in params_list], numpy.array(Lmins))

    def __call__(self, params, data_mask):
        self.ncalls += 1
        if self.pending is None:
            return self.one(params, data_mask)
        plist, Lmins = self.pending
        self.pending = None
        assert numpy.array_equal(plist[0], params)
        Ls = [self.one(p, data_mask) for p in plist]
        counts = numpy.array([(L > Lmins).sum() for L in Ls])
        hits = numpy.where(counts > 0)[0]
        k = int(hits[0]) if len(hits) else -1
        self.last_draw = (k, Ls[k] if k >= 0 else None, counts)
        return Ls[0] if k == 0 else None


@pytest.mark.parametrize('tag,batch,adaptive', [('a', 1, False), ('a', 16, False), ('a', 16, True),
                                                ('b', 4, False), ('c', 1, True), ('c', 7, False),
                                                ('c', 8, True), ('d', 1, False), ('d', 16, True),
                                                ('d', 5, False)])
def test_constrainer_reproduces_reference_draws_on_the_oracle(fixture, tag, batch, adaptive):
    ndata, nlive, niter, seed_data, seed_run = fixture[tag + '_cfg']
    x, y, _ = synth.horns(int(ndata), seed=int(seed_data))
    like = OracleLikelihood(x, y)
    c = MetricLearningFriendsConstrainer(batch_size=batch, speculator=like if batch > 1 else None,
                                         adaptive=adaptive, region_class=OracleRegion, **CONFIG[tag])
    res = run_draws(c, like, int(ndata), int(nlive), int(niter), int(seed_run),
                    rank=int(fixture[tag + '_rank']))
    check(res, fixture, tag)
    assert c.region.maxdistance == float(fixture[tag + '_maxdistance'])
    if batch == 1:
        assert like.ncalls == int(fixture[tag + '_ncalls'])      # one pass per candidate
    else:
        # fewer passes over the data than candidates tried
        assert like.ncalls - int(nlive) == c.nbatches < fixture[tag + '_ntoaccept'].sum()


def test_constrainer_rejects_unknown_metric():
    with pytest.raises(ValueError):
        MetricLearningFriendsConstrainer(metriclearner='mahalanobis')


# ------------------------------------------------------------------------- GPU: the product path
@pytest.mark.gpu
@pytest.mark.parametrize('tag,batch,adaptive', [('a', 1, False), ('a', 16, False), ('a', 16, True),
                                                ('b', 8, False), ('c', 16, True), ('d', 16, True),
                                                ('d', 16, False)])
def test_constrainer_reproduces_reference_draws_on_the_device(fixture, tag, batch, adaptive):
    from massivedatans_b200 import _lib
    from massivedatans_b200.likelihood import make_multi_loglikelihood
    ndata, nlive, niter, seed_data, seed_run = fixture[tag + '_cfg']
    x, y, _ = synth.horns(int(ndata), seed=int(seed_data))
    like = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
    c = MetricLearningFriendsConstrainer(batch_size=batch, speculator=like if batch > 1 else None,
                                         adaptive=adaptive, **CONFIG[tag])
    before = _lib.load().mdns_launch_count()
    res = run_draws(c, like, int(ndata), int(nlive), int(niter), int(seed_run),
                    rank=int(fixture[tag + '_rank']))
    assert _lib.load().mdns_launch_count() > before
    check(res, fixture, tag)
    assert c.region.maxdistance == float(fixture[tag + '_maxdistance'])
    tries = int(fixture[tag + '_ntoaccept'].sum())
    if batch == 1:
        assert c.nbatches == tries
    else:
        assert c.nbatches < tries <= c.nscored


@pytest.mark.gpu
def test_speculation_at_scale_matches_one_by_one():
    # 20 000 data sets: the speculative run must make the same draws as the one-candidate run
    # of the same seed (no reference fixture at this size; the one-by-one form is the pinned one)
    from massivedatans_b200.likelihood import make_multi_loglikelihood
    ndata, nlive, niter = 20000, 60, 40
    x, y, _ = synth.horns(ndata, seed=31)
    like = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
    runs = []
    for batch in (1, 16):
        c = MetricLearningFriendsConstrainer(batch_size=batch, speculator=like if batch > 1 else None,
                                             adaptive=False, **CONFIG['a'])
        runs.append((run_draws(c, like, ndata, nlive, niter, 9), c))
    (r1, c1), (r16, c16) = runs
    assert numpy.array_equal(r1['u'], r16['u'])
    assert numpy.array_equal(r1['ntoaccept'], r16['ntoaccept'])
    assert numpy.array_equal(r1['naccepted'], r16['naccepted'])
    assert numpy.allclose(r1['L'], r16['L'], rtol=1e-10, atol=0, equal_nan=True)
    assert c16.nbatches <= c1.nbatches


@pytest.mark.gpu
def test_device_proposals_are_a_statistically_equivalent_generator(fixture):
    # proposals from the fused device generator (not numpy's stream): every draw is still a valid
    # constrained draw, and the run needs about as many tries as the reference's run
    from massivedatans_b200.likelihood import make_multi_loglikelihood
    tag = 'a'
    ndata, nlive, niter, seed_data, seed_run = fixture[tag + '_cfg']
    x, y, _ = synth.horns(int(ndata), seed=int(seed_data))
    like = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
    c = MetricLearningFriendsConstrainer(batch_size=16, speculator=like, device_proposals=4096,
                                         **CONFIG[tag])
    res = run_draws(c, like, int(ndata), int(nlive), int(niter), int(seed_run))
    assert ((res['u'] > 0) & (res['u'] < 1)).all()
    assert (res['naccepted'] >= 1).all()                   # run_draws asserts L > Lmins somewhere
    assert not numpy.array_equal(res['u'], fixture[tag + '_u'])        # a different stream
    want = fixture[tag + '_ntoaccept']
    # same efficiency.  The number of tries is heavy-tailed (up to ~100 for the single-data-set
    # groups) and runs diverge after the first draw; six host-RNG runs of this configuration
    # with seeds 3..8 (oracle backend) gave mean tries 2.85 .. 4.44, first-try share 0.34 .. 0.46
    # and share within four tries 0.745 .. 0.862 (the fixture's run: 3.95, 0.34, 0.745)
    tries = res['ntoaccept']
    assert 2.3 < tries.mean() < 5.5, tries.mean()
    assert 0.28 < (tries <= 1).mean() < 0.54, (tries <= 1).mean()
    assert 0.68 < (tries <= 4).mean() < 0.92, (tries <= 4).mean()
    assert want.mean() > 0
    # and the same typical likelihood gain per draw (the accepted points sit in the same shells)
    gain = numpy.nanmedian(res['L'], axis=1)
    gain_ref = numpy.nanmedian(fixture[tag + '_L'], axis=1)
    assert abs(numpy.median(gain[-100:]) - numpy.median(gain_ref[-100:])) < \
        0.5 * numpy.std(gain_ref[-100:]) + 1.0
    # reproducible from the numpy seed
    c2 = MetricLearningFriendsConstrainer(batch_size=16, speculator=like, device_proposals=4096,
                                          **CONFIG[tag])
    res2 = run_draws(c2, like, int(ndata), int(nlive), int(niter), int(seed_run))
    assert numpy.array_equal(res['u'], res2['u'])
