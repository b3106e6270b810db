#!/usr/bin/env python
"""Small invocation of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tools/sanitize_smoke.py
Checks results against the oracle as it goes (sizes chosen to hit ragged tiles)."""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from massivedatans_b200 import _lib, synth  # noqa: E402
from massivedatans_b200.clustering import neighbors  # noqa: E402
from massivedatans_b200.likelihood import ResidentDataset  # noqa: E402
from massivedatans_b200.livepoints import LiveTable  # noqa: E402
from oracle import port  # noqa: E402


def rel(a, b):
    return numpy.max(numpy.abs(a - b) / numpy.maximum(numpy.abs(b), 1e-300))


def main():
    lib = _lib.load()
    N, nx = 777, 203
    x, y, _ = synth.horns(N, nx=nx, seed=1)
    ds = ResidentDataset(x, y)
    allm = numpy.ones(N, dtype=bool)
    masks = synth.masks(N)
    seen = set()
    for tuning, K in (((0, 0, 0, 0), 1), ((0, 0, 0, 0), 3), ((8, 2, 4, 4), 5), ((8, 4, 8, 2), 9),
                      ((1, 116, 8, 3), 9), ((1, 32, 16, 3), 17), ((2, 2, 8, 3), 9), ((2, 4, 16, 2), 17),
                      ((3, 0, 8, 4), 9), ((3, 0, 16, 3), 17), ((3, 0, 32, 2), 33), ((3, 0, 8, 14), 9),
                      ((3, 0, 16, 13), 17), ((6, 0, 16, 2), 17), ((6, 0, 16, 3), 9), ((6, 0, 8, 2), 9)):
        ds.set_tuning(*tuning)
        pts = synth.parameter_points(K, seed=K)
        for mname in ('all', 'half'):
            m = masks[mname]
            got = ds.loglike_batch(pts, None if mname == 'all' else m, synth.NOISE_LEVEL, scale=1.0)
            seen.add(lib.mdns_last_kernel().decode())
            for k in (0, K - 1):
                want = port.clike(x, y, pts[k][0], pts[k][1], pts[k][2], synth.NOISE_LEVEL,
                                  allm if mname == 'all' else m)
                assert rel(got[k], want) < 1e-10, (tuning, K, mname, k)
    # the slab kernel in row-pair mode (pitch 200 = 8 mod 16, odd count) and on an aligned pitch
    for Ns, nxs in ((4097, 200), (1500, 192)):
        xs_, ys_, _ = synth.horns(Ns, nx=nxs, legacy=False, seed=2)
        dss = ResidentDataset(xs_, ys_)
        dss.set_tuning(6, 0, 16, 2)
        ptss = synth.parameter_points(16, seed=3)
        for mname in ('all', 'half'):
            ms_ = synth.masks(Ns)[mname]
            gots = dss.loglike_batch(ptss, None if mname == 'all' else ms_, synth.NOISE_LEVEL, scale=1.0)
            seen.add(lib.mdns_last_kernel().decode())
            want = port.clike(xs_, ys_, ptss[15][0], ptss[15][1], ptss[15][2], synth.NOISE_LEVEL, ms_)
            assert rel(gots[15], want) < 1e-10, (Ns, nxs, mname)
        # one launch per speculative pass over a handful of data sets
        few = numpy.zeros(Ns, dtype=bool)
        few[[3, 77, Ns - 1]] = True
        dss.set_tuning(0, 0, 0, 0)
        Lf = dss.loglike_batch(ptss, few, synth.NOISE_LEVEL).copy()
        th = numpy.sort(Lf, axis=0)[12]                # three candidates per data set get through
        kf, Lk, cf = dss.draw_pass(few, th, ptss, synth.NOISE_LEVEL)
        seen.add(lib.mdns_last_kernel().decode())
        wc = (Lf > th).sum(axis=1)
        assert numpy.array_equal(cf, wc) and kf == int(numpy.nonzero(wc)[0][0]) and numpy.array_equal(Lk, Lf[kf])
        dss.close()
    ds.set_tuning(0, 0, 0, 0)
    pts = synth.parameter_points(6, seed=2)
    Ls = numpy.array([-0.5 * port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, masks['half']) for p in pts])
    k, L, counts = ds.first_accepted(pts, masks['half'], Ls[:3].max(axis=0) + 1e-9, synth.NOISE_LEVEL)
    assert k >= 3 and rel(L, Ls[k]) < 1e-10
    # MUSE-type
    ym, vm, t = synth.muse(ndata=37, nspec=3600)
    dm = ResidentDataset(None, ym, variance=vm)
    for tuning in ((0, 0, 0, 0), (32, 8, 0, 0), (256, 0, 2, 0)):
        dm.set_tuning(*tuning)
        Lm = numpy.zeros((1, 37))
        dm.muse_loglike(t, numpy.ones(37, dtype=bool), Lm)
        seen.add(lib.mdns_last_kernel().decode())
        assert rel(Lm[0], port.cmuselike(ym, vm, t, numpy.ones(37, dtype=bool))) < 1e-10
    # neighbours
    xx, yy = synth.members_and_candidates(333, 1111, 3)
    chosen = synth.bootstrap_chosen(333, 10, numpy.random.RandomState(1))
    r = lib.mdns_bootstrapped_maxdistance(xx.ctypes.data, 333, 3, chosen.ctypes.data, 10)
    assert r == port.bootstrapped_maxdistance_chosen(xx, chosen)
    assert numpy.array_equal(neighbors.count_within_distance_of(xx, r, yy), port.count_within_distance_of(xx, r, yy))
    assert numpy.array_equal(neighbors.any_within_distance_of(xx, r, yy), port.any_within_distance_of(xx, r, yy))
    assert neighbors.most_distant_nearest_neighbor(xx) == port.most_distant_nearest_neighbor(xx)
    assert neighbors.is_within_distance_of(xx, r, yy[0]) == port.is_within_distance_of(xx, r, yy[0])
    # full counts of a large problem: lanes over candidates, member ranges, atomics
    xb, yb = synth.members_and_candidates(4099, 4101, 3)
    assert numpy.array_equal(neighbors.count_within_distance_of(xb, 0.05, yb), port.count_within_distance_of(xb, 0.05, yb))
    seen.add(lib.mdns_last_kernel().decode())
    # live table
    rs = numpy.random.RandomState(3)
    Lt = rs.normal(size=(23, N))
    tab = LiveTable(ds, 23)
    tab.upload(Lt)
    lo, at, hi = tab.prepare()
    assert numpy.array_equal(lo, Lt.min(axis=0)) and numpy.array_equal(at, Lt.argmin(axis=0))
    idx = numpy.arange(0, N, 3)
    shelves = [rs.normal(size=int(q)) for q in rs.randint(0, 40, size=len(idx))]
    got = tab.lmins_higher(idx, shelves)
    for j, d in enumerate(idx):
        n = len(shelves[j])
        assert got[j] == numpy.partition(numpy.concatenate((Lt[:, d], shelves[j])), n)[n]
    tab.replace(at, rs.normal(size=N))
    P = rs.randint(0, 500, size=(23, N)).astype(numpy.int64)
    tab.upload_points(P)
    assert numpy.array_equal(tab.subsets(None, 500), port.subsets_labels(P, allm, 500))
    print('sanitize smoke ok; kernels:', sorted(seen), 'launches', lib.mdns_launch_count())


if __name__ == '__main__':
    main()
