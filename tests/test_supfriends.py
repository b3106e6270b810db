"""Per-axis "SupFriends" distance (clustering/neighbors.py:22-73) against
tests/golden/supfriends.npz, which the reference's own functions produced
(tests/golden/make_golden_supfriends.py).  CPU: the oracle's numpy restatement, bit for bit,
including the position of the numpy random stream afterwards.  GPU: the mirror, whose pairwise
parts run on the device, bit for bit as well."""
import os

import numpy
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'supfriends.npz')
TAGS = ['a', 'b', 'c', 'd', 'e']


@pytest.fixture(scope='module')
def fixture():
    return numpy.load(GOLDEN)


@pytest.mark.parametrize('tag', TAGS)
def test_oracle_reproduces_reference(fixture, tag):
    from oracle import np as onp
    u = fixture[tag + '_u']
    run_seed, nboot = fixture[tag + '_cfg']
    assert numpy.array_equal(onp.initial_maxdistance_guess(u), fixture[tag + '_initial'])
    numpy.random.seed(int(run_seed))
    assert numpy.array_equal(onp.find_maxdistance(u, nbootstraps=int(nboot)), fixture[tag + '_maxdistance'])
    numpy.random.seed(int(run_seed) + 100)
    got = onp.update_maxdistance(u, fixture[tag + '_initial'] * 0.25)
    assert numpy.array_equal(got, fixture[tag + '_round_small'])
    assert numpy.random.uniform() == float(fixture[tag + '_after'])


@pytest.mark.gpu
@pytest.mark.parametrize('tag', TAGS)
def test_device_mirror_reproduces_reference(fixture, tag):
    from massivedatans_b200 import _lib
    from massivedatans_b200.clustering import neighbors
    u = fixture[tag + '_u']
    run_seed, nboot = fixture[tag + '_cfg']
    before = _lib.load().mdns_launch_count()
    assert numpy.array_equal(neighbors.initial_maxdistance_guess(u), fixture[tag + '_initial'])
    numpy.random.seed(int(run_seed))
    got = neighbors.find_maxdistance(u, nbootstraps=int(nboot))
    assert numpy.array_equal(got, fixture[tag + '_maxdistance'])
    assert _lib.load().mdns_launch_count() >= before + 2 + int(nboot)
    numpy.random.seed(int(run_seed) + 100)
    got = neighbors.update_maxdistance(u, 0, fixture[tag + '_initial'] * 0.25)
    assert numpy.array_equal(got, fixture[tag + '_round_small'])
    assert numpy.random.uniform() == float(fixture[tag + '_after'])


@pytest.mark.gpu
def test_nearest_index_and_coverage_against_numpy():
    import scipy.spatial
    from massivedatans_b200.clustering.radfriendsregion import ResidentMembers
    rs = numpy.random.RandomState(8)
    for n, d in ((2, 1), (33, 2), (700, 4), (1500, 9)):
        u = rs.uniform(size=(n, d))
        m = ResidentMembers(u)
        dist = scipy.spatial.distance.cdist(u, u)
        numpy.fill_diagonal(dist, numpy.inf)
        assert numpy.array_equal(m.nearest_index(), dist.argmin(axis=1))
        md = rs.uniform(0.02, 0.3, size=d)
        query = rs.permutation(n)[:max(1, n // 3)]
        ref = rs.permutation(n)[:max(1, n // 2)]
        want = numpy.array([numpy.all(numpy.abs(u[i] - u[ref]) < md, axis=1).any() for i in query])
        assert numpy.array_equal(m.axis_covered(md, query, ref), want)
    assert not m.axis_covered(md, query, []).any()       # nobody to be covered by
    with pytest.raises(Exception):
        m.axis_covered(md, [n], ref)                     # index out of range
