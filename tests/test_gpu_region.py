"""RadFriendsRegion mirror vs tests/golden/region.npz -- arrays produced by the REFERENCE's own
clustering/radfriendsregion.py on the reference's cneighbors.so (tests/golden/make_golden_region.py).
Same seeds, same numpy.random call order, bit-exact neighbour decisions: every array must match
bit for bit (radius, bounding box, membership, counts, and every batch of generated candidates)."""
import numpy
import pytest

from massivedatans_b200.clustering.radfriendsregion import RadFriendsRegion

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_region_reproduces_reference_class(golden, tag):
    g = golden('region')
    members, cand = g[tag + '_members'], g[tag + '_cand']
    numpy.random.seed(int(g[tag + '_seed']))
    region = RadFriendsRegion(members=members, nbootstraps=10)
    assert region.maxdistance == float(g[tag + '_maxdistance'])
    assert numpy.array_equal(region.lo, g[tag + '_lo']) and numpy.array_equal(region.hi, g[tag + '_hi'])
    inside = region.are_inside(cand)
    assert inside.dtype == numpy.bool_ and numpy.array_equal(inside, g[tag + '_inside'])
    nnear = region.count_nearby_members(cand)
    assert nnear.dtype.kind == 'i' and numpy.array_equal(nnear, g[tag + '_nnear'])
    assert numpy.array_equal([region.is_inside(c) for c in cand[:40]], g[tag + '_is_inside'])
    spent = []
    for i, (us, ntotal) in enumerate(region.generate(nmax=8000)):
        assert numpy.array_equal(us, g['%s_gen%d' % (tag, i)]), i
        spent.append(ntotal)
    assert numpy.array_equal(spent, g[tag + '_ntotal'])
    region.add_members(cand[:25] * 0.5)
    assert numpy.array_equal(region.lo, g[tag + '_lo2'])
    assert numpy.array_equal(region.are_inside(cand), g[tag + '_inside2'])


def test_region_given_radius_and_far_point():
    rs = numpy.random.RandomState(1)
    members = rs.uniform(size=(50, 4))
    region = RadFriendsRegion(members=members, maxdistance=0.05)
    assert region.maxdistance == 0.05
    assert region.is_inside(members[3]) and not region.is_inside(members[3] + 10.0)
    assert region.are_inside(members).all()
    assert (region.count_nearby_members(members) >= 1).all()
    with pytest.raises(NotImplementedError):
        RadFriendsRegion(members=members, metric='chebyshev')


@pytest.mark.parametrize('ndim,n', [(3, 400), (2, 60), (5, 300)])
def test_device_generation_is_uniform_in_the_region(ndim, n):
    # statistical parity with the host generator (radfriendsregion.py:156-178): both draw
    # uniformly from the union of balls
    rs = numpy.random.RandomState(ndim)
    members = rs.uniform(size=(n, ndim))
    region = RadFriendsRegion(members=members, maxdistance=0.18 if ndim == 3 else 0.25)
    m = 200000
    dev = region.generate_device(m, seed=7)
    numpy.random.seed(5)
    host = numpy.vstack([region._ball_round(1000, ndim) for _ in range(m // 1000)])
    # (1) every generated point lies in the region
    assert region.are_inside(dev).all()
    # (2) same acceptance rate (binomial error)
    p_dev, p_host = len(dev) / m, len(host) / m
    sigma = numpy.sqrt(p_host * (1 - p_host) / m * 2)
    assert abs(p_dev - p_host) < 5 * sigma, (p_dev, p_host)
    # (3) same first and second moments
    se = host.std(axis=0) / numpy.sqrt(len(host)) * numpy.sqrt(2)
    assert (numpy.abs(dev.mean(axis=0) - host.mean(axis=0)) < 5 * se).all()
    assert numpy.allclose(numpy.cov(dev.T), numpy.cov(host.T), rtol=0.05, atol=2e-3)
    # (4) uniform density: the number of members near a generated point is distributed alike
    nd = numpy.bincount(region.count_nearby_members(dev[:50000]), minlength=40)[:40]
    nh = numpy.bincount(region.count_nearby_members(host[:50000]), minlength=40)[:40]
    fd, fh = nd / nd.sum(), nh / nh.sum()
    err = numpy.sqrt((fd + fh) / min(nd.sum(), nh.sum())) + 1e-4
    assert (numpy.abs(fd - fh) < 6 * err).all()
    # (5) the stream is counter-based: same seed -> same points, split launches concatenate
    again = region.generate_device(m, seed=7)
    assert numpy.array_equal(dev, again)
    a = region.generate_device(50000, seed=7, first_proposal=0)
    b = region.generate_device(m - 50000, seed=7, first_proposal=50000)
    assert numpy.array_equal(numpy.vstack([a, b]), dev)
    assert not numpy.array_equal(region.generate_device(1000, seed=8)[:5], dev[:5])
