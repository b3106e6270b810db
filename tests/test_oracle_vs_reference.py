"""Where the compiled reference is present (oracle/_ref, built from /root/reference by
oracle/Makefile) check the plain-C restatement against it on fresh random inputs,
bit for bit.  oracle/_ref travels to the GPU box, so this also runs there."""
import numpy
import pytest

from massivedatans_b200 import synth
from oracle import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason='oracle/_ref not built')


@pytest.mark.parametrize('N,nx', [(1, 200), (33, 200), (500, 57), (128, 1000)])
def test_clike(oracle_port, N, nx):
    x, y, _ = synth.horns(N, nx=nx, seed=N + nx)
    rs = numpy.random.RandomState(N)
    for p in synth.parameter_points(4, seed=N):
        for mask in (numpy.ones(N, dtype=bool), rs.uniform(size=N) < 0.4):
            a = ref.clike(x, y, p[0], p[1], p[2], 0.01, mask)
            b = oracle_port.clike(x, y, p[0], p[1], p[2], 0.01, mask)
            assert numpy.array_equal(a, b)


def test_clike_accumulates(oracle_port):
    x, y, _ = synth.horns(20)
    mask = numpy.ones(20, dtype=bool)
    a = ref.clike(x, y, 0.3, 600., 4., 0.01, mask, Lout=numpy.full(20, 5.0))
    b = oracle_port.clike(x, y, 0.3, 600., 4., 0.01, mask, Lout=numpy.full(20, 5.0))
    assert numpy.array_equal(a, b)


@pytest.mark.parametrize('ndata,nspec', [(5, 11), (40, 360), (17, 3600)])
def test_cmuselike(oracle_port, ndata, nspec):
    y, v, t = synth.muse(ndata=ndata, nspec=nspec, seed=ndata)
    rs = numpy.random.RandomState(ndata)
    mask = rs.uniform(size=ndata) < 0.6
    assert numpy.array_equal(ref.cmuselike(y, v, t, mask), oracle_port.cmuselike(y, v, t, mask))
    assert numpy.array_equal(ref.cmuselike(y, v, t, mask, parallel=True),
                             oracle_port.cmuselike(y, v, t, mask))


@pytest.mark.parametrize('n,m,ndim', [(2, 3, 1), (50, 200, 2), (400, 500, 3), (300, 100, 7)])
def test_neighbors(oracle_port, n, m, ndim):
    xx, yy = synth.members_and_candidates(n, m, ndim, seed=n)
    rs = numpy.random.RandomState(n)
    chosen = synth.bootstrap_chosen(n, 10, rs)
    r = ref.bootstrapped_maxdistance_chosen(xx, chosen)
    assert r == oracle_port.bootstrapped_maxdistance_chosen(xx, chosen)
    assert r == ref.bootstrapped_maxdistance_chosen(xx, chosen, parallel=True)
    assert ref.most_distant_nearest_neighbor(xx) == oracle_port.most_distant_nearest_neighbor(xx)
    for rr in (r, 0.5 * r, 3 * r, 0.0):
        assert numpy.array_equal(ref.count_within_distance_of(xx, rr, yy),
                                 oracle_port.count_within_distance_of(xx, rr, yy))
        for cm in (1, 2, 5):
            a = ref.count_within_distance_of_raw(xx, rr, yy, numpy.zeros(m), cm)
            b = oracle_port.count_within_distance_of_raw(xx, rr, yy, numpy.zeros(m), cm)
            assert numpy.array_equal(a, b)
        for j in range(min(m, 20)):
            assert ref.is_within_distance_of(xx, rr, yy[j].copy()) == \
                oracle_port.is_within_distance_of(xx, rr, yy[j].copy())
