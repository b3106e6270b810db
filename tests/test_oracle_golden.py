"""Pin the CPU oracle (oracle/mdns_oracle.c + numpy twin) against the committed golden
vectors, which were produced by the unmodified reference C (tests/golden/make_golden.py).
Bit-exact for the C restatement; 1e-12 relative for the numpy twin (different summation
order)."""
import numpy

from conftest import rel_err
from massivedatans_b200 import synth
from oracle import np as onp


def test_clike_port_bit_exact(golden, oracle_port):
    g = golden('clike')
    N = int(g['N'])
    x, y, _ = synth.horns(N)
    for name in ('all', 'half', 'sparse', 'prefix'):
        m = g['mask_' + name]
        for p, want in zip(g['params'], g['Lout_' + name]):
            got = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
            assert got.shape == (int(m.sum()),)
            assert numpy.array_equal(got, want), name


def test_clike_spectrum_matches_line_model(golden, oracle_port):
    g = golden('clike')
    N = int(g['N'])
    x, y, _ = synth.horns(N)
    m = g['mask_half']
    for p, want in zip(g['params'], g['Lout_half']):
        ypred = onp.line_model(x, p[0], p[1], p[2])
        got = oracle_port.clike_spectrum(ypred, y, synth.NOISE_LEVEL, m)
        assert rel_err(got, want) < 1e-13


def test_clike_numpy_twin(golden):
    g = golden('clike')
    N2 = int(g['nothing_N'])
    x, y = synth.nothing(N2)
    allm = numpy.ones(N2, dtype=bool)
    for p, want in zip(g['params'][:3], g['nothing_Lout']):
        assert rel_err(onp.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm), want) < 1e-12


def test_clike_null_evidence_identity(golden):
    # plotevidences.py:17 -- with no line (A -> 0) logL = sum -0.5 (y/0.01)^2
    g = golden('clike')
    N2 = int(g['nothing_N'])
    x, y = synth.nothing(N2)
    want = ((y / synth.NOISE_LEVEL) ** 2).sum(axis=0)
    got = onp.clike(x, y, 0.0, 500., 1., synth.NOISE_LEVEL, numpy.ones(N2, dtype=bool))
    assert rel_err(got, want) < 1e-13


def test_cmuselike_port_bit_exact(golden, oracle_port):
    g = golden('cmuselike')
    y, v, _ = synth.muse(ndata=int(g['ndata']), nspec=int(g['nspec']))
    mask = g['mask']
    for ph, want, want_all in zip(g['phases'], g['Lout'], g['Lall']):
        yp = synth.muse_template(int(g['nspec']), phase=float(ph))
        got = oracle_port.cmuselike(y, v, yp, mask)
        assert numpy.array_equal(got, want)          # incl. zeros outside the mask
        assert numpy.array_equal(oracle_port.cmuselike(y, v, yp, numpy.ones(len(mask), dtype=bool)),
                                 want_all)
        twin = onp.cmuselike(y, v, yp, mask)
        assert rel_err(twin[mask], want[mask]) < 1e-12
    got = oracle_port.cmuselike(g['small_y'], g['small_v'], g['small_ypred'], g['small_mask'])
    assert numpy.array_equal(got, g['small_Lout'])


def test_cmuselike_leaves_unmasked_untouched(golden, oracle_port):
    g = golden('cmuselike')
    L = numpy.full(7, 123.0)
    oracle_port.cmuselike(g['small_y'], g['small_v'], g['small_ypred'], g['small_mask'], Lout=L)
    assert (L[~g['small_mask']] == 123.0).all()
    assert numpy.array_equal(L[g['small_mask']], g['small_Lout'][g['small_mask']])


def test_neighbors_selftest_values(golden, oracle_port):
    # clustering/neighbors.py:240-250: seeds 0..99 on uniform(size=(200,2)) seed 1
    g = golden('neighbors')
    u = g['selftest_u']
    for i, want in enumerate(g['selftest_maxdistance']):
        numpy.random.seed(i)
        chosen = synth.bootstrap_chosen(200, 10)
        assert oracle_port.bootstrapped_maxdistance_chosen(u, chosen) == want
        if i < 5:
            assert abs(onp.bootstrapped_maxdistance(u, chosen) - want) < 1e-14
    # values recorded in SURVEY.md section 8c for seeds 97/98/99
    assert abs(g['selftest_maxdistance'][97] - 0.114310016) < 1e-8
    assert abs(g['selftest_maxdistance'][98] - 0.184130467) < 1e-8
    assert abs(g['selftest_maxdistance'][99] - 0.133890081) < 1e-8


def test_neighbors_port_bit_exact(golden, oracle_port):
    g = golden('neighbors')
    for ndim in (2, 3, 5):
        k = 'd%d_' % ndim
        xx, yy = synth.members_and_candidates(400, 1000, ndim, seed=ndim)
        r = float(g[k + 'r'])
        assert oracle_port.bootstrapped_maxdistance_chosen(xx, g[k + 'chosen']) == r
        assert oracle_port.most_distant_nearest_neighbor(xx) == float(g[k + 'mdnn'])
        assert numpy.array_equal(oracle_port.count_within_distance_of(xx, r, yy), g[k + 'counts'])
        assert numpy.array_equal(oracle_port.any_within_distance_of(xx, r, yy), g[k + 'any'])
        cm3 = oracle_port.count_within_distance_of_raw(xx, r, yy, numpy.zeros(len(yy)), 3)
        assert numpy.array_equal(cm3, g[k + 'counts_cm3'])
        assert numpy.array_equal(cm3, numpy.minimum(g[k + 'counts'], 3))
        w = [oracle_port.is_within_distance_of(xx, r, yy[j].copy()) for j in range(50)]
        assert numpy.array_equal(numpy.array(w), g[k + 'within'])
        # the identity the reference states at neighbors.py:143-146,155-158
        assert numpy.array_equal(onp.count_within_distance_of(xx, r, yy), g[k + 'counts'])
        assert numpy.array_equal(onp.any_within_distance_of(xx, r, yy), g[k + 'any'])
        assert abs(onp.most_distant_nearest_neighbor(xx) - float(g[k + 'mdnn'])) < 1e-14


def test_neighbors_quirk_sample0_skipped(golden, oracle_port):
    g = golden('neighbors')
    got = oracle_port.bootstrapped_maxdistance_chosen(g['quirk_x'], g['quirk_chosen'])
    assert got == float(g['quirk_r'])
    assert got < 1.0          # the outlier at index 0 (distance ~14) is ignored


def test_live_table_oracle_matches_numpy_twin(oracle_port):
    # multi_nested_sampler.py:38-47,134-137,531: C restatement vs the numpy expressions
    from oracle import np_oracle
    rs = numpy.random.RandomState(11)
    L = rs.normal(size=(60, 41)) * 10
    L[9, 4] = L[3, 4] = L[:, 4].min() - 1          # tie: first occurrence
    got = oracle_port.live_colstats(L)
    want = np_oracle.live_colstats(L)
    for g, w in zip(got, want):
        assert numpy.array_equal(g, w)
    for n in (0, 1, 2, 7, 25):
        shelf = rs.normal(size=n) * 10
        if n > 1:
            shelf[0] = L[5, 0]
        assert oracle_port.find_nsmallest(n, numpy.ascontiguousarray(L[:, 0]), shelf) == \
            np_oracle.find_nsmallest(n, L[:, 0], shelf)


def test_subsets_oracle_matches_scipy_components(oracle_port):
    # multi_nested_sampler.py:204-355: union-find restatement vs scipy connected components
    from oracle import np_oracle
    rs = numpy.random.RandomState(3)
    nlive, ndata, ngroups = 10, 200, 9
    group = rs.randint(0, ngroups, size=ndata)
    P = numpy.empty((nlive, ndata), dtype=numpy.int64)
    for d in range(ndata):
        P[:, d] = group[d] * 40 + rs.choice(40, size=nlive, replace=False)
    for m in (numpy.ones(ndata, dtype=bool), rs.uniform(size=ndata) < 0.3):
        got = oracle_port.subsets_labels(P, m, ngroups * 40)
        assert numpy.array_equal(got, np_oracle.subsets_labels(P, m, ngroups * 40))
        assert (got[~m] == -1).all() and (got[m] <= numpy.nonzero(m)[0]).all()
