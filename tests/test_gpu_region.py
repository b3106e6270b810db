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
