"""GPU parity tests of the live-point table against the oracle
restatements of multi_nested_sampler.py:38-47,134-137,438-447,520-524,531.  Everything is a
selection: results must be bit-identical."""
import numpy
import pytest

from massivedatans_b200 import _lib, synth
from massivedatans_b200.likelihood import ResidentDataset
from massivedatans_b200.livepoints import LiveTable

pytestmark = pytest.mark.gpu


def _table(nlive, ndata, seed, devices=None):
    x, y, _ = synth.horns(ndata, legacy=False, seed=seed)
    ds = ResidentDataset(x, y, devices=devices)
    rs = numpy.random.RandomState(seed)
    L = rs.normal(size=(nlive, ndata)) * 100
    # ties: the minimum of some columns occurs twice (numpy.argmin keeps the first)
    for d in range(0, ndata if nlive > 1 else 0, 7):
        i, j = sorted(rs.choice(nlive, size=2, replace=False))
        L[j, d] = L[i, d] = L[:, d].min() - 1.0
    t = LiveTable(ds, nlive)
    t.upload(L)
    return x, y, ds, t, L, rs


@pytest.mark.parametrize('nlive,ndata', [(1, 1), (5, 3), (400, 257), (100, 40001), (33, 100000)])
def test_prepare_matches_numpy_and_oracle(oracle_port, nlive, ndata):
    _, _, ds, t, L, _ = _table(nlive, ndata, seed=nlive + ndata)
    lo, at, hi = t.prepare()
    assert numpy.array_equal(lo, L.min(axis=0))
    assert numpy.array_equal(at, L.argmin(axis=0)) and at.dtype == numpy.int64
    assert numpy.array_equal(hi, L.max(axis=0))
    if nlive * ndata <= 200000:
        wlo, wat, whi = oracle_port.live_colstats(L)
        assert numpy.array_equal(lo, wlo) and numpy.array_equal(at, wat) and numpy.array_equal(hi, whi)
    assert numpy.array_equal(t.download(), L)


def test_replace_and_roundtrip():
    nlive, ndata = 50, 3001
    _, _, ds, t, L, rs = _table(nlive, ndata, seed=4)
    lo, at, hi = t.prepare()
    rows = at.copy()
    rows[::3] = -1                      # data sets that are not advanced this time
    vals = rs.normal(size=ndata)
    t.replace(rows, vals)
    want = L.copy()
    adv = rows >= 0
    want[rows[adv], numpy.nonzero(adv)[0]] = vals[adv]
    assert numpy.array_equal(t.download(), want)
    lo2, at2, _ = t.prepare()
    assert numpy.array_equal(lo2, want.min(axis=0)) and numpy.array_equal(at2, want.argmin(axis=0))


@pytest.mark.parametrize('nlive,ndata,maxshelf', [(400, 500, 6), (100, 2000, 40), (7, 64, 3),
                                                  (30, 300, 90)])
def test_lmins_higher_matches_find_nsmallest(oracle_port, nlive, ndata, maxshelf):
    _, _, ds, t, L, rs = _table(nlive, ndata, seed=9)
    idx = numpy.sort(rs.choice(ndata, size=max(1, ndata // 3), replace=False))
    shelves = []
    for d in idx:
        n = int(rs.randint(0, maxshelf + 1))
        s = rs.normal(size=n) * 100
        if n > 2:
            s[0] = L[rs.randint(nlive), d]          # duplicates across the two arrays
            s[1] = s[2]                               # and inside the shelf
        shelves.append(s)
    got = t.lmins_higher(idx, shelves)
    for j, d in enumerate(idx):
        n = len(shelves[j])
        want = oracle_port.find_nsmallest(n, numpy.ascontiguousarray(L[:, d]), shelves[j])
        assert got[j] == want, (j, d, n)
        assert want == numpy.partition(numpy.concatenate((L[:, d], shelves[j])), n)[n]
    # empty shelves give the column minimum (Lmins_higher starts as a copy of Lmins)
    lo, _, _ = t.prepare()
    assert numpy.array_equal(t.lmins_higher(idx, [[] for _ in idx]), lo[idx])
    with pytest.raises(_lib.MdnsError):
        t.lmins_higher(idx[::-1], shelves[::-1])     # indices must be increasing


def test_initial_population_fills_the_table(oracle_port):
    # multi_nested_sampler.py:91-111: nlive full-mask likelihood calls -> live_pointsL
    nlive, ndata = 40, 33000
    x, y, _ = synth.horns(ndata, legacy=False, seed=2)
    ds = ResidentDataset(x, y)
    t = LiveTable(ds, nlive)
    pts = synth.parameter_points(nlive, seed=6)
    want = numpy.empty((nlive, ndata))
    for r0 in range(0, nlive, 16):
        ds.stage_params(pts[r0:r0 + 16])
        ds.set_mask(None)
        ds.launch_clike(synth.NOISE_LEVEL, -0.5)
        t.fill_from_launch(r0)
        want[r0:r0 + 16] = ds.loglike_batch(pts[r0:r0 + 16], None, synth.NOISE_LEVEL)
    got = t.download()
    assert numpy.array_equal(got, want)
    allm = numpy.ones(ndata, dtype=bool)
    for k in (0, 17, nlive - 1):
        p = pts[k]
        ref = -0.5 * oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert numpy.max(numpy.abs(got[k] - ref) / numpy.abs(ref)) < 1e-10
    # a masked launch cannot fill rows of the table
    ds.set_mask(synth.masks(ndata)['half'])
    ds.stage_params(pts[:2])
    ds.launch_clike(synth.NOISE_LEVEL, -0.5)
    with pytest.raises(_lib.MdnsError):
        t.fill_from_launch(0)


def test_sharded_table_matches_single_device():
    if _lib.load().mdns_device_count() < 2:
        pytest.skip('needs two devices')
    _, _, ds, t, L, rs = _table(60, 5003, seed=3, devices=[0, 1])
    lo, at, hi = t.prepare()
    assert numpy.array_equal(lo, L.min(axis=0)) and numpy.array_equal(at, L.argmin(axis=0))
    idx = numpy.arange(0, 5003, 5)
    shelves = [rs.normal(size=int(rs.randint(0, 5))) * 100 for _ in idx]
    got = t.lmins_higher(idx, shelves)
    for j, d in enumerate(idx):
        n = len(shelves[j])
        assert got[j] == numpy.partition(numpy.concatenate((L[:, d], shelves[j])), n)[n]


def _grouped_points(nlive, ndata, ngroups, rs, chain=False):
    """live_pointsp with a known group structure: data sets of a group draw their live points
    from the group's own pool (plus, with `chain`, one point shared with the next data set only,
    so that labels must travel along a chain)."""
    group = rs.randint(0, ngroups, size=ndata)
    pool = 3 * nlive
    P = numpy.empty((nlive, ndata), dtype=numpy.int64)
    for d in range(ndata):
        P[:, d] = group[d] * pool + rs.choice(pool, size=nlive, replace=False)
    npoints = ngroups * pool
    if chain:
        # every data set owns private points; neighbours d, d+1 share exactly one
        P = numpy.arange(nlive * ndata, dtype=numpy.int64).reshape((nlive, ndata))
        for d in range(ndata - 1):
            if (d + 1) % 97:                 # break the chain now and then
                P[0, d + 1] = P[1, d]
        npoints = nlive * ndata
    return P, npoints


@pytest.mark.parametrize('nlive,ndata,ngroups,chain', [(20, 300, 7, False), (400, 2000, 50, False),
                                                       (5, 5000, 1, True), (50, 20000, 900, False),
                                                       (3, 1, 1, False)])
def test_subsets_match_oracle_components(oracle_port, nlive, ndata, ngroups, chain):
    from oracle import np_oracle
    from massivedatans_b200.livepoints import generate_subsets
    rs = numpy.random.RandomState(ndata + nlive)
    P, npoints = _grouped_points(nlive, ndata, ngroups, rs, chain)
    x, y, _ = synth.horns(ndata, legacy=False, seed=1)
    t = LiveTable(ResidentDataset(x, y), nlive)
    t.upload_points(P)
    masks = {'all': numpy.ones(ndata, dtype=bool), 'half': rs.uniform(size=ndata) < 0.5}
    for name, m in masks.items():
        if not m.any():
            continue
        got = t.subsets(m if name != 'all' else None, npoints)
        want = oracle_port.subsets_labels(P, m, npoints)
        assert numpy.array_equal(got, want), name
        if ndata <= 5000:
            assert numpy.array_equal(got, np_oracle.subsets_labels(P, m, npoints))
        groups = list(generate_subsets(m, P, got))
        assert len(groups) == len(numpy.unique(want[want >= 0]))
        seen = numpy.zeros(ndata, dtype=bool)
        firsts = []
        for gm, gp in groups:
            assert not (seen & gm).any()
            seen |= gm
            firsts.append(numpy.nonzero(gm)[0][0])
            assert numpy.array_equal(gp, numpy.unique(P[:, gm]))
        assert numpy.array_equal(seen, m) and firsts == sorted(firsts)
    # replacing points merges groups: give data set 0 a point of the last data set
    if ndata > 1:
        rows = numpy.full(ndata, -1, dtype=numpy.int64)
        ids = numpy.zeros(ndata, dtype=numpy.int64)
        rows[0] = 0
        ids[0] = P[-1, ndata - 1]
        t.replace_points(rows, ids)
        P2 = P.copy()
        P2[0, 0] = ids[0]
        allm = numpy.ones(ndata, dtype=bool)
        assert numpy.array_equal(t.subsets(None, npoints), oracle_port.subsets_labels(P2, allm, npoints))


def test_thresholds_staged_from_the_table_match_the_host_path(oracle_port):
    # live table -> thresholds -> device-side accept test, no host round trip of Lmins
    nlive, ndata = 20, 40000
    x, y, _ = synth.horns(ndata, legacy=False, seed=12)
    ds = ResidentDataset(x, y)
    t = LiveTable(ds, nlive)
    init = synth.parameter_points(nlive, seed=3)
    for r0 in range(0, nlive, 10):
        ds.stage_params(init[r0:r0 + 10])
        ds.set_mask(None)
        ds.launch_clike(synth.NOISE_LEVEL, -0.5)
        t.fill_from_launch(r0)
    Lmins, _, _ = t.prepare()
    cands = synth.parameter_points(12, seed=99)
    ds.begin_draw(None, Lmins)                       # host path
    k1, L1, c1 = ds.draw_batch(cands, synth.NOISE_LEVEL)
    t.stage_thresholds()                             # device path
    k2, L2, c2 = ds.draw_batch(cands, synth.NOISE_LEVEL)
    assert k1 == k2 and numpy.array_equal(c1, c2)
    if k1 >= 0:
        assert numpy.array_equal(L1, L2)
    # the counts are what numpy gives on the downloaded table
    full = ds.loglike_batch(cands, None, synth.NOISE_LEVEL)
    assert numpy.array_equal(c2, (full > t.download().min(axis=0)).sum(axis=1))
    ds.set_mask(synth.masks(ndata)['half'])
    with pytest.raises(_lib.MdnsError):
        _lib.check(_lib.load().mdns_livetable_stage_thresholds(t._h, ds._h), 'stage_thresholds')
