"""GPU parity tests of the RadFriends neighbour kernels: outputs must be IDENTICAL
(bit-exact doubles, equal counts / booleans) to the CPU oracle, which is pinned
bit-exact to cneighbors.c."""
import ctypes
import os

import numpy
import pytest

from massivedatans_b200 import _lib, synth
from massivedatans_b200.clustering import neighbors

pytestmark = pytest.mark.gpu


def test_golden_vectors(golden):
    g = golden('neighbors')
    for ndim in (2, 3, 5):
        k = 'd%d_' % ndim
        xx, yy = synth.members_and_candidates(400, 1000, ndim, seed=ndim)
        r = float(g[k + 'r'])
        numpy.random.seed(ndim)
        assert neighbors.bootstrapped_maxdistance(xx, 10) == r
        assert neighbors.most_distant_nearest_neighbor(xx) == float(g[k + 'mdnn'])
        c = neighbors.count_within_distance_of(xx, r, yy)
        assert c.dtype.kind == 'i' and numpy.array_equal(c, g[k + 'counts'])
        a = neighbors.any_within_distance_of(xx, r, yy)
        assert a.dtype == numpy.bool_ and numpy.array_equal(a, g[k + 'any'])
        w = [neighbors.is_within_distance_of(xx, r, yy[j]) for j in range(50)]
        assert numpy.array_equal(numpy.array(w), g[k + 'within'])
        cm3 = numpy.zeros(len(yy))
        lib = _lib.load()
        assert lib.mdns_count_within_distance_of(xx.ctypes.data, 400, ndim, r, yy.ctypes.data,
                                                 len(yy), cm3.ctypes.data, 3) == 0
        assert numpy.array_equal(cm3, g[k + 'counts_cm3'])


def test_reference_selftest_seeds(golden):
    # clustering/neighbors.py:240-250 -- seeds 0..99 on uniform(size=(200,2)), seed 1
    g = golden('neighbors')
    u = g['selftest_u']
    for i, want in enumerate(g['selftest_maxdistance']):
        numpy.random.seed(i)
        a = neighbors.bootstrapped_maxdistance(u, 10)
        numpy.random.seed(i)
        b = neighbors.find_rdistance(u, nbootstraps=10, metric='euclidean', verbose=False)
        assert a == want and b == want


def test_quirk_sample0_skipped(golden):
    g = golden('neighbors')
    x = numpy.ascontiguousarray(g['quirk_x'])
    c = numpy.ascontiguousarray(g['quirk_chosen'])
    r = _lib.load().mdns_bootstrapped_maxdistance(x.ctypes.data, len(x), 2, c.ctypes.data, 1)
    assert r == float(g['quirk_r'])


@pytest.mark.parametrize('n,m,ndim', [(1, 1, 1), (2, 3, 1), (31, 33, 2), (400, 1000, 3),
                                      (5000, 10000, 3), (777, 12000, 4), (1000, 500, 8),
                                      (300, 100, 11), (40000, 3000, 3)])
def test_count_any_within_vs_oracle(oracle_port, n, m, ndim):
    xx, yy = synth.members_and_candidates(n, m, ndim, seed=n + m)
    r0 = 0.5 * n ** (-1.0 / ndim)
    for r in (r0, 3 * r0, 0.0, 10.0):
        assert numpy.array_equal(neighbors.count_within_distance_of(xx, r, yy),
                                 oracle_port.count_within_distance_of(xx, r, yy))
        assert numpy.array_equal(neighbors.any_within_distance_of(xx, r, yy),
                                 oracle_port.any_within_distance_of(xx, r, yy))
    for j in range(min(m, 10)):
        assert neighbors.is_within_distance_of(xx, r0, yy[j]) == \
            oracle_port.is_within_distance_of(xx, r0, yy[j].copy())


@pytest.mark.parametrize('n,m,ndim', [(4099, 4100, 1), (4099, 4100, 2), (4099, 4101, 3), (5001, 3400, 4),
                                      (4099, 4100, 5), (4099, 4100, 6), (4099, 4100, 7), (4099, 4103, 8),
                                      (257, 70001, 3)])
def test_full_counts_of_large_problems_take_the_tiled_kernel(oracle_port, n, m, ndim):
    # lanes over candidates, compare on the bit patterns (count_tile_kernel): same counts, bit for bit
    xx, yy = synth.members_and_candidates(n, m, ndim, seed=n + m + ndim)
    yy[::97] = xx[numpy.arange(0, m, 97) % n]           # candidates ON members: distance exactly 0
    r0 = 0.5 * n ** (-1.0 / ndim)
    for r in (r0, 2.5 * r0, 0.0, 10.0, numpy.nextafter(r0, 1)):
        got = neighbors.count_within_distance_of(xx, r, yy)
        assert _lib.load().mdns_last_kernel() == b'count_tile_kernel'
        assert numpy.array_equal(got, oracle_port.count_within_distance_of(xx, r, yy)), r
    # `any` queries keep the early-exit kernel
    neighbors.any_within_distance_of(xx, r0, yy)
    assert _lib.load().mdns_last_kernel() == b'count_within_kernel'


def test_count_ties_exact_on_lattice_tiled(oracle_port):
    g = numpy.arange(20, dtype=float)
    xx = numpy.ascontiguousarray(numpy.array(numpy.meshgrid(g, g, g)).reshape(3, -1).T)      # 8000 members
    yy = numpy.ascontiguousarray(numpy.tile(xx[::3], (1, 1)) + numpy.array([0.5, 0.0, 0.0]))  # 2667 candidates
    assert len(xx) * len(yy) >= 1 << 24
    for r in (0.5, 1.5, numpy.sqrt(2.25 + 1), 2.5, numpy.nextafter(2.5, 3)):
        got = neighbors.count_within_distance_of(xx, r, yy)
        assert _lib.load().mdns_last_kernel() == b'count_tile_kernel'
        assert numpy.array_equal(got, oracle_port.count_within_distance_of(xx, r, yy))


def test_count_ties_exact_on_lattice(oracle_port):
    # integer lattice: many distances coincide exactly with the radius -> strict '<'
    g = numpy.arange(12, dtype=float)
    xx = numpy.ascontiguousarray(numpy.array(numpy.meshgrid(g, g, g)).reshape(3, -1).T)
    yy = xx[::7].copy() + numpy.array([0.5, 0.0, 0.0])
    for r in (0.5, 1.5, numpy.sqrt(2.25 + 1), 2.5, numpy.nextafter(2.5, 3)):
        assert numpy.array_equal(neighbors.count_within_distance_of(xx, r, yy),
                                 oracle_port.count_within_distance_of(xx, r, yy))


def test_count_semantics_of_out_and_countmax(oracle_port):
    lib = _lib.load()
    xx, yy = synth.members_and_candidates(300, 200, 3, seed=1)
    r = 0.2
    for countmax in (0, 1, 2, 5):
        for start in (0.0, 2.0, -3.0, 0.5):
            a = numpy.full(len(yy), start)
            b = a.copy()
            assert lib.mdns_count_within_distance_of(xx.ctypes.data, 300, 3, r, yy.ctypes.data,
                                                     len(yy), a.ctypes.data, countmax) == 0
            oracle_port.count_within_distance_of_raw(xx, r, yy, b, countmax)
            assert numpy.array_equal(a, b), (countmax, start)


@pytest.mark.parametrize('n,ndim,nboot', [(2, 1, 1), (5, 2, 3), (200, 2, 10), (400, 3, 10),
                                          (2500, 3, 10), (5180, 3, 10), (1000, 5, 20),
                                          (300, 9, 15), (9000, 4, 15)])
def test_bootstrap_and_mdnn_vs_oracle(oracle_port, n, ndim, nboot):
    xx, _ = synth.members_and_candidates(n, 1, ndim, seed=n)
    lib = _lib.load()
    rs = numpy.random.RandomState(n)
    chosen = synth.bootstrap_chosen(n, nboot, rs)
    got = lib.mdns_bootstrapped_maxdistance(xx.ctypes.data, n, ndim, chosen.ctypes.data, nboot)
    assert got == oracle_port.bootstrapped_maxdistance_chosen(xx, chosen)
    assert neighbors.most_distant_nearest_neighbor(xx) == \
        oracle_port.most_distant_nearest_neighbor(xx)
    assert neighbors.nearest_rdistance_guess(xx) == oracle_port.most_distant_nearest_neighbor(xx)


def test_bootstrap_degenerate_rounds(oracle_port):
    lib = _lib.load()
    xx, _ = synth.members_and_candidates(50, 1, 3, seed=2)
    for chosen in (numpy.ones((50, 2)), numpy.zeros((50, 2)),
                   numpy.column_stack([numpy.ones(50), numpy.zeros(50)])):
        chosen = numpy.ascontiguousarray(chosen)
        got = lib.mdns_bootstrapped_maxdistance(xx.ctypes.data, 50, 3, chosen.ctypes.data, 2)
        assert got == oracle_port.bootstrapped_maxdistance_chosen(xx, chosen)
    dup = numpy.zeros((6, 2))          # identical points: distance exactly 0
    ch = numpy.array([[1.], [0.], [1.], [0.], [0.], [1.]])
    assert lib.mdns_bootstrapped_maxdistance(dup.ctypes.data, 6, 2, ch.ctypes.data, 1) == 0.0
    assert neighbors.most_distant_nearest_neighbor(dup) == 0.0


def test_dropin_cneighbors_with_reference_argtypes(oracle_port):
    from numpy.ctypeslib import ndpointer
    lib = ctypes.CDLL(os.path.join(_lib.DROPIN_DIR, 'cneighbors.so'))
    f2 = ndpointer(dtype=numpy.float64, ndim=2, flags='C_CONTIGUOUS')
    f1 = ndpointer(dtype=numpy.float64, ndim=1, flags='C_CONTIGUOUS')
    lib.most_distant_nearest_neighbor.argtypes = [f2, ctypes.c_int, ctypes.c_int]
    lib.most_distant_nearest_neighbor.restype = ctypes.c_double
    lib.is_within_distance_of.argtypes = [f2, ctypes.c_int, ctypes.c_int, ctypes.c_double, f1]
    lib.is_within_distance_of.restype = ctypes.c_int
    lib.count_within_distance_of.argtypes = [f2, ctypes.c_int, ctypes.c_int, ctypes.c_double, f2,
                                             ctypes.c_int, f1, ctypes.c_int]
    lib.bootstrapped_maxdistance.argtypes = [f2, ctypes.c_int, ctypes.c_int, f2, ctypes.c_int]
    lib.bootstrapped_maxdistance.restype = ctypes.c_double
    xx, yy = synth.members_and_candidates(500, 300, 3, seed=8)
    chosen = synth.bootstrap_chosen(500, 10, numpy.random.RandomState(3))
    r = lib.bootstrapped_maxdistance(xx, 500, 3, chosen, 10)
    assert r == oracle_port.bootstrapped_maxdistance_chosen(xx, chosen)
    assert lib.most_distant_nearest_neighbor(xx, 500, 3) == \
        oracle_port.most_distant_nearest_neighbor(xx)
    counts = numpy.zeros(300)
    lib.count_within_distance_of(xx, 500, 3, r, yy, 300, counts, 0)
    assert numpy.array_equal(counts.astype(int), oracle_port.count_within_distance_of(xx, r, yy))
    assert lib.is_within_distance_of(xx, 500, 3, r, yy[0].copy()) == \
        int(oracle_port.is_within_distance_of(xx, r, yy[0].copy()))
    _lib.load().mdns_legacy_reset()
