"""GPU parity tests of the batched likelihood: CUDA path (through the C ABI) vs the CPU
oracle on identical seeded inputs and vs the committed golden vectors.

Tolerance: the north star's 1e-9 relative for FP64 logL; what we actually assert is
1e-12 (the kernel sums in a different order and uses the device exp, so bit equality
is not expected, but anything beyond a few hundred ulps would be a bug)."""
import ctypes

import numpy
import pytest

# the FP64 tensor-path kernels: stream-K (round 2) and whole-tile (round 1, kept for the shapes
# only it is instantiated for)
TENSOR = (b'rows_dmma_kernel', b'clike_dmma_kernel', b'slab_dmma_kernel')
TENSOR_GATHER = (b'rows_dmma_kernel(gather)', b'clike_dmma_kernel(gather)', b'slab_dmma_kernel(gather)')

from conftest import rel_err
from massivedatans_b200 import _lib, synth
from massivedatans_b200.likelihood import (ResidentDataset, make_multi_loglikelihood,
                                           make_muse_loglikelihood)

pytestmark = pytest.mark.gpu
TOL = 1e-12          # asserted; the contract is 1e-9 relative
TOL_XP = 1e-10       # expanded form Syy - 2 Sym + Smm: the bound the kernel itself enforces


@pytest.fixture(scope='module')
def horns257():
    x, y, _ = synth.horns(257)
    return x, y, ResidentDataset(x, y)


def test_clike_golden_vectors(golden, horns257):
    g = golden('clike')
    x, y, ds = horns257
    for name in ('all', 'half', 'sparse', 'prefix'):
        m = g['mask_' + name]
        got = ds.loglike_batch(g['params'], m, synth.NOISE_LEVEL, scale=1.0)
        assert got.shape == (len(g['params']), int(m.sum()))
        assert rel_err(got, g['Lout_' + name]) < TOL, name
    # one candidate at a time through the reference-shaped callable (sample.py:101)
    f = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
    for p, want in zip(g['params'], g['Lout_half']):
        L = f((p[0], p[1], numpy.log10(p[2])), g['mask_half'])
        assert L.shape == want.shape and L.dtype == numpy.float64
        assert rel_err(L, -0.5 * want) < TOL


def test_clike_nothing_golden(golden):
    g = golden('clike')
    N = int(g['nothing_N'])
    x, y = synth.nothing(N)
    ds = ResidentDataset(x, y)
    got = ds.loglike_batch(g['params'][:3], None, synth.NOISE_LEVEL, scale=1.0)
    assert rel_err(got, g['nothing_Lout']) < TOL


@pytest.mark.parametrize('N,nx', [(1, 200), (2, 3), (33, 57), (700, 200), (5000, 1000),
                                  (20000, 200), (40000, 31)])
def test_clike_vs_oracle_shapes(oracle_port, N, nx):
    x, y, _ = synth.horns(N, nx=nx, seed=N + nx)
    ds = ResidentDataset(x, y)
    pts = synth.parameter_points(3, seed=N)
    for name, m in synth.masks(N, seed=N).items():
        got = ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0)
        assert got.shape == (3, int(m.sum()))
        # all-active batches of >= 3 candidates over >= 8192 data sets take the expanded form
        tol = TOL_XP if (m.all() and N >= 8192) else TOL
        for k, p in enumerate(pts):
            want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
            assert rel_err(got[k], want) < tol, (name, k)


@pytest.mark.parametrize('K', [1, 2, 3, 5, 8, 9, 17, 64])
def test_clike_batch_sizes(oracle_port, K):
    N = 3000
    x, y, _ = synth.horns(N)
    ds = ResidentDataset(x, y)
    pts = synth.parameter_points(K, seed=K)
    m = synth.masks(N)['half']
    got = ds.loglike_batch(pts, m, synth.NOISE_LEVEL)
    for k in (0, K // 2, K - 1):
        want = -0.5 * oracle_port.clike(x, y, pts[k][0], pts[k][1], pts[k][2],
                                        synth.NOISE_LEVEL, m)
        assert rel_err(got[k], want) < TOL


@pytest.mark.parametrize('lanes,unroll,ktile,rows', [
    (8, 1, 1, 1), (8, 2, 2, 1), (8, 4, 4, 1), (8, 8, 8, 1), (8, 13, 1, 1), (8, 16, 8, 1),
    (32, 1, 1, 1), (32, 4, 2, 1), (32, 8, 8, 1), (32, 13, 4, 1), (32, 16, 1, 1),
    # register-blocked kernel: rows data sets per lane group
    (8, 2, 8, 2), (8, 4, 8, 2), (8, 1, 8, 4), (8, 2, 8, 4), (8, 2, 4, 2), (8, 4, 4, 2),
    (8, 1, 4, 4), (8, 2, 4, 4),
    # lane-per-data-set tile kernel: (1, channels per stage, candidates per pass, ring stages)
    # (all-active rows only: with a mask these fall back to the lanes-across-channels kernels)
    (1, 32, 4, 3), (1, 32, 8, 3), (1, 32, 16, 3), (1, 16, 8, 4), (1, 16, 16, 6), (1, 32, 32, 3),
    (1, 16, 32, 4), (1, 16, 8, 6), (1, 0, 0, 0),
    # 256-row tiles: unroll = 100 + channels per stage
    (1, 116, 4, 3), (1, 116, 8, 3), (1, 116, 16, 4), (1, 116, 32, 3), (1, 132, 8, 3),
    (1, 132, 16, 3)])
def test_clike_kernel_variants(oracle_port, lanes, unroll, ktile, rows):
    N = 1111
    x, y, _ = synth.horns(N, nx=203, seed=3)      # odd channel count: padded fragment
    ds = ResidentDataset(x, y)
    ds.set_tuning(lanes, unroll, ktile, rows)
    pts = synth.parameter_points(5, seed=1)
    m = synth.masks(N)['half']
    got = ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0)
    full = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    allm = numpy.ones(N, dtype=bool)
    for k, p in enumerate(pts):
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
        assert rel_err(got[k], want) < TOL
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(full[k], want) < TOL


@pytest.mark.parametrize('N,nx,K', [(128, 16, 4), (129, 18, 9), (5000, 200, 17), (1000, 33, 40),
                                    (300, 1000, 5), (77, 1599, 4)])
def test_clike_tile_kernel_shapes(oracle_port, N, nx, K):
    x, y, _ = synth.horns(N, nx=nx, seed=N)
    ds = ResidentDataset(x, y)
    ds.set_tuning(1, 0, 0, 0)
    pts = synth.parameter_points(K, seed=N)
    got = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    allm = numpy.ones(N, dtype=bool)
    for k in (0, K // 2, K - 1):
        p = pts[k]
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL


@pytest.mark.parametrize('N,nx,K,lane_rows,ktile,stages', [
    (128, 16, 4, 2, 8, 3), (129, 18, 9, 2, 8, 2), (5000, 200, 17, 2, 16, 3),
    (1000, 33, 40, 2, 32, 3), (300, 1000, 5, 4, 8, 2), (3000, 203, 33, 4, 16, 3),
    (40000, 200, 16, 0, 0, 0)])
def test_clike_expanded_form_vs_oracle(oracle_port, N, nx, K, lane_rows, ktile, stages):
    # explicit selection of the expanded tile kernel (lanes = 2) on ordinary data: nothing
    # cancels, nothing is recomputed, results within the enforced bound
    x, y, _ = synth.horns(N, nx=nx, seed=N + 1)
    ds = ResidentDataset(x, y)
    ds.set_tuning(2, lane_rows, ktile, stages)
    pts = synth.parameter_points(K, seed=N)
    got = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    assert _lib.load().mdns_last_kernel() == b'clike_xtile_kernel'
    allm = numpy.ones(N, dtype=bool)
    for k in sorted(set((0, 1, K // 2, K - 2, K - 1))):
        p = pts[k]
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL_XP
    assert ds.expanded_stats() == (True, 0)
    # direct and expanded forms agree far inside the tolerance on this data
    ds.set_expanded(False)
    ds.set_tuning(1, 0, 0, 0)
    direct = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    assert _lib.load().mdns_last_kernel() == b'clike_tile_kernel'
    assert rel_err(got, direct) < TOL_XP


@pytest.mark.parametrize('lane_rows,ktile,stages', [(2, 8, 3), (2, 16, 3), (2, 32, 2), (2, 8, 2),
                                                    (4, 8, 3), (4, 16, 2), (2, 32, 3), (2, 16, 2)])
@pytest.mark.parametrize('N,nx,K', [(700, 203, 9), (70000, 200, 37)])
def test_clike_expanded_blocked_variants(oracle_port, lane_rows, ktile, stages, N, nx, K):
    # expanded form with 2 / 4 data sets per consumer lane
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N)
    ds = ResidentDataset(x, y)
    ds.set_tuning(2, lane_rows, ktile, stages)
    pts = synth.parameter_points(K, seed=N + 1)
    got = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    assert _lib.load().mdns_last_kernel() == b'clike_xtile_kernel'
    allm = numpy.ones(N, dtype=bool)
    for k in sorted(set((0, 7, 8, K // 2, K - 1))):
        p = pts[k]
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL_XP
    assert ds.expanded_stats() == (True, 0)


@pytest.mark.parametrize('ktile,stages', [(8, 2), (16, 2), (16, 3), (32, 2), (32, 3), (8, 4),
                                          (8, 14), (16, 13), (32, 12), (16, 12)])
@pytest.mark.parametrize('N,nx,K', [(128, 16, 4), (700, 203, 9), (300, 1000, 5), (70000, 200, 37),
                                    (2049, 57, 64)])
def test_clike_expanded_tensor_path_variants(oracle_port, ktile, stages, N, nx, K):
    # expanded form with the cross term as FP64 tensor-core tiles (lanes = 3)
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N + 2)
    ds = ResidentDataset(x, y)
    ds.set_tuning(3, 0, ktile, stages)
    pts = synth.parameter_points(K, seed=N + 3)
    got = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    assert _lib.load().mdns_last_kernel() in TENSOR
    allm = numpy.ones(N, dtype=bool)
    for k in sorted(set((0, 1, 7, 8, K // 2, K - 2, K - 1)) & set(range(K))):
        p = pts[k]
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL_XP, k
    assert ds.expanded_stats() == (True, 0)
    ds.set_tuning(2, 2, 8, 2)
    fma_form = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    assert rel_err(got, fma_form) < TOL_XP


@pytest.mark.parametrize('ktile,nslot', [(16, 3), (16, 2), (8, 3), (8, 2)])
@pytest.mark.parametrize('N,nx,K', [
    (70001, 200, 16),     # pitch 200 = 8 mod 16: rows read in line-aligned PAIRS, odd count
    (64, 200, 16),        # exactly one slab of pairs
    (33, 200, 3),         # one slab, half empty; 3 of 8/16 candidates valid
    (1, 200, 16),
    (4097, 192, 20),      # aligned pitch: single rows; two passes of 16, three of 8
    (5000, 199, 9),       # odd channel count: pitch 200, channel 199 is padding
    (3000, 57, 17),       # pitch 58: misaligned single rows, last box 10 channels
    (3000, 40, 16),       # pitch 40: pairs with S = 2 (five boxes per pair)
    (2000, 52, 8),        # pitch 52: last box 4 channels (k-steps of pure padding skipped)
    (300000, 72, 16),     # many slabs per warp: the slab counter hands out > 2 per warp
])
def test_clike_slab_tensor_kernel(oracle_port, ktile, nslot, N, nx, K):
    # per-warp slabs, candidate batch resident in shared memory (lanes = 6): short spectra
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N + 12)
    ds = ResidentDataset(x, y)
    ds.set_tuning(6, 0, ktile, nslot)
    pts = synth.parameter_points(K, seed=N + 13)
    lib = _lib.load()
    allm = numpy.ones(N, dtype=bool)
    got = None
    # (spectra shorter than the ring is deep fall back to the automatic choice)
    pitch = nx + (nx & 1)
    slab = pitch >= 16 * nslot
    for rep in range(3):          # the slab counter must come back at zero after every launch
        g = numpy.array(ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0))
        assert (lib.mdns_last_kernel() == b'slab_dmma_kernel') == slab
        assert got is None or numpy.array_equal(g, got)
        got = g
    for k in sorted(set((0, 1, 7, 8, K // 2, K - 2, K - 1)) & set(range(K))):
        p = pts[k]
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL_XP, k
    assert ds.expanded_stats() == (True, 0)
    # the stream-K kernel sums the same channels in the same order
    ds.set_tuning(3, 0, 16 if ktile == 16 else 8, 3)
    ref = numpy.array(ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0))
    assert lib.mdns_last_kernel() in TENSOR
    assert rel_err(got, ref) < 1e-13
    # fused accept test: counts and first accepted candidate against numpy on the same matrix
    ds.set_tuning(6, 0, ktile, nslot)
    L = -0.5 * got
    srt = numpy.sort(L, axis=0)
    Lmins = srt[-1] + 1.0 + numpy.abs(srt[-1])
    pick = numpy.arange(N) % 7 == 3
    if K > 1 and pick.any():
        Lmins[pick] = 0.5 * (srt[-1][pick] + srt[-2][pick])
    k, Lk, counts = ds.first_accepted(pts, None, Lmins, synth.NOISE_LEVEL)
    want_counts = (L > Lmins).sum(axis=1)
    assert numpy.array_equal(counts, want_counts)
    if want_counts.any():
        want_k = int(numpy.nonzero(want_counts)[0][0])
        assert k == want_k and rel_err(Lk, L[want_k]) < 1e-13
    else:
        assert k == -1


@pytest.mark.parametrize('ktile,nslot', [(16, 2), (16, 3), (8, 2)])
@pytest.mark.parametrize('N,nx,K', [(700, 203, 9), (70001, 200, 16), (2049, 57, 20), (150000, 72, 16)])
def test_clike_slab_kernel_masked_gather(oracle_port, ktile, nslot, N, nx, K):
    # masked batches on the per-warp slab kernel: each warp fetches its listed rows itself with
    # eight gather4 copies per box
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N + 21)
    ds = ResidentDataset(x, y)
    pts = synth.parameter_points(K, seed=N + 22)
    lib = _lib.load()
    for name, m in synth.masks(N, seed=N).items():
        if name == 'all' or not m.any():
            continue
        ds.set_tuning(6, 0, ktile, nslot)
        got = numpy.array(ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0))
        assert lib.mdns_last_kernel() == b'slab_dmma_kernel(gather)'
        assert got.shape == (K, int(m.sum()))
        again = numpy.array(ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0))
        assert numpy.array_equal(got, again)
        for k in sorted(set((0, 7, 8, K // 2, K - 1))):
            p = pts[k]
            want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
            assert rel_err(got[k], want) < TOL_XP, (name, k)
        # fused accept test on the compacted rows
        L = -0.5 * got
        srt = numpy.sort(L, axis=0)
        Lmins = srt[-1] + 1.0 + numpy.abs(srt[-1])
        pick = numpy.arange(L.shape[1]) % 5 == 2
        if pick.any():
            Lmins[pick] = 0.5 * (srt[-1][pick] + srt[-2][pick])
        k, Lk, counts = ds.first_accepted(pts, m, Lmins, synth.NOISE_LEVEL)
        want_counts = (L > Lmins).sum(axis=1)
        assert numpy.array_equal(counts, want_counts), name
        if want_counts.any():
            assert k == int(numpy.nonzero(want_counts)[0][0])
    assert ds.expanded_stats() == (True, 0)


@pytest.mark.parametrize('ktile,stages', [(8, 2), (8, 14), (16, 3), (16, 13), (32, 3), (32, 12)])
@pytest.mark.parametrize('N,nx,K', [(700, 203, 9), (70000, 200, 37), (2049, 57, 20)])
def test_clike_expanded_tensor_path_masked_gather(oracle_port, ktile, stages, N, nx, K):
    # masked batches on the tensor path: rows fetched with gather4 copies of the active list
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N + 5)
    ds = ResidentDataset(x, y)
    ds.set_tuning(3, 0, ktile, stages)
    pts = synth.parameter_points(K, seed=N + 6)
    for name, m in synth.masks(N, seed=N).items():
        if name == 'all' or not m.any():
            continue
        got = ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0)
        assert _lib.load().mdns_last_kernel() in TENSOR_GATHER
        assert got.shape == (K, int(m.sum()))
        for k in sorted(set((0, 7, 8, K // 2, K - 1))):
            p = pts[k]
            want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
            assert rel_err(got[k], want) < TOL_XP, (name, k)
    assert ds.expanded_stats() == (True, 0)


@pytest.mark.parametrize('N,nx,K', [(1, 200, 2), (33, 57, 3), (700, 203, 9), (10000, 200, 16),
                                    (10000, 200, 5), (20000, 200, 17), (3000, 1000, 35), (40000, 31, 8)])
def test_clike_small_batches_in_one_launch(oracle_port, N, nx, K):
    # small batches of parameter points: spectra built by every CTA, direct form, one launch
    # (clike_small_kernel) -- automatic up to a few 1e5 evaluations, lanes = 7 at any size
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N + 31)
    ds = ResidentDataset(x, y)
    lib = _lib.load()
    pts = synth.parameter_points(K, seed=N + 32)
    for name, m in synth.masks(N, seed=N).items():
        if not m.any():
            continue
        ds.set_tuning(7, 0, 0, 0)
        got = numpy.array(ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0))
        assert lib.mdns_last_kernel() == b'clike_small_kernel'
        assert got.shape == (K, int(m.sum()))
        for k in sorted({0, 1, K // 2, K - 1}):
            p = pts[k]
            want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
            assert rel_err(got[k], want) < TOL, (name, k)
        # the same lanes, fragment order and butterfly as the lanes-across-channels kernel
        ds.set_tuning(8, 0, 0, 1)
        same = numpy.array(ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0))
        assert lib.mdns_last_kernel() == b'clike_rows_kernel'
        assert numpy.array_equal(got, same), name
        # automatic choice + fused accept test
        ds.set_tuning(0, 0, 0, 0)
        auto = numpy.array(ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0))
        n_act = int(m.sum())
        if K >= 5 and n_act > 64 and n_act * ((K + 7) // 8 * 8) <= 100000:
            assert lib.mdns_last_kernel() == b'clike_small_kernel', name
            assert numpy.array_equal(auto, got)
        else:
            assert rel_err(auto, got) < TOL_XP
        ds.set_tuning(7, 0, 0, 0)
        L = -0.5 * got
        srt = numpy.sort(L, axis=0)
        Lmins = srt[-1] + 1.0 + numpy.abs(srt[-1])
        pick = numpy.arange(L.shape[1]) % 7 == 3
        if pick.any():
            Lmins[pick] = 0.5 * (srt[-1][pick] + srt[-2][pick])
        k, Lk, counts = ds.first_accepted(pts, m, Lmins, synth.NOISE_LEVEL)
        want_counts = (L > Lmins).sum(axis=1)
        assert numpy.array_equal(counts, want_counts)
        if want_counts.any():
            want_k = int(numpy.nonzero(want_counts)[0][0])
            assert k == want_k and numpy.array_equal(Lk, L[want_k])
        else:
            assert k == -1
    ds.close()


def test_clike_masked_batches_automatic_choice(oracle_port):
    # masked batches: lanes-across-channels kernels up to 4 candidates, gather-fed tensor path
    # from 5 on
    N = 80000
    x, y, _ = synth.horns(N, legacy=False, seed=8)
    ds = ResidentDataset(x, y)
    lib = _lib.load()
    m = synth.masks(N)['half']
    pts = synth.parameter_points(20, seed=9)
    got = ds.loglike_batch(pts, m, synth.NOISE_LEVEL)
    assert lib.mdns_last_kernel() in TENSOR_GATHER
    for k in (0, 15, 16, 19):
        p = pts[k]
        want = -0.5 * oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
        assert rel_err(got[k], want) < TOL_XP
    mid = ds.loglike_batch(pts[:8], m, synth.NOISE_LEVEL)
    assert lib.mdns_last_kernel() in TENSOR_GATHER
    assert rel_err(mid, got[:8]) < TOL_XP
    small = ds.loglike_batch(pts[:4], m, synth.NOISE_LEVEL)
    assert lib.mdns_last_kernel() in (b'clike_block_kernel', b'clike_rows_kernel')
    assert rel_err(small, got[:4]) < TOL_XP
    # first-accept on a masked batch of 20 goes through the same kernel
    Ls = numpy.array(got)
    Lmins = Ls[:18].max(axis=0) + 1e-9 * numpy.abs(Ls[:18].max(axis=0))
    k, L, counts = ds.first_accepted(pts, m, Lmins, synth.NOISE_LEVEL)
    want_counts = (Ls > Lmins).sum(axis=1)
    assert numpy.array_equal(counts, want_counts)
    assert k == (int(numpy.argmax(want_counts > 0)) if (want_counts > 0).any() else -1)


def test_clike_expanded_form_automatic_choice(oracle_port):
    # all-active batches of >= 3 candidates take the expanded form on their own, masked ones from
    # 5 candidates over >= 4096 active rows; smaller batches stay on the direct kernels
    N = 50000
    x, y, _ = synth.horns(N, legacy=False, seed=21)
    ds = ResidentDataset(x, y)
    lib = _lib.load()
    pts = synth.parameter_points(35, seed=2)
    got = numpy.array(ds.loglike_batch(pts, None, synth.NOISE_LEVEL))
    assert lib.mdns_last_kernel() in TENSOR
    allm = numpy.ones(N, dtype=bool)
    for k in (0, 7, 8, 31, 32, 34):
        p = pts[k]
        want = -0.5 * oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL_XP
    ds.loglike_batch(pts[:2], None, synth.NOISE_LEVEL)
    assert lib.mdns_last_kernel() not in (b'clike_xtile_kernel',) + TENSOR
    # (masked batches of >= 5 candidates over >= 4096 active rows: the gathered tensor path)
    half = synth.masks(N)['half']
    gm = ds.loglike_batch(pts, half, synth.NOISE_LEVEL)
    assert lib.mdns_last_kernel() in TENSOR_GATHER
    assert rel_err(gm, got[:, half]) < TOL_XP
    ds.set_expanded(False)
    got = ds.loglike_batch(pts[:9], None, synth.NOISE_LEVEL)
    assert lib.mdns_last_kernel() == b'clike_tile_kernel'
    for k in (0, 8):
        p = pts[k]
        want = -0.5 * oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL


@pytest.mark.parametrize('tuning,kernel', [((2, 2, 8, 3), b'clike_xtile_kernel'),
                                           ((3, 0, 8, 2), b'clike_dmma_kernel'),
                                           ((3, 0, 8, 3), b'rows_dmma_kernel'),
                                           ((3, 0, 8, 13), b'rows_dmma_kernel'),
                                           ((6, 0, 8, 3), b'slab_dmma_kernel'),
                                           ((6, 0, 16, 2), b'slab_dmma_kernel')])
def test_clike_expanded_form_cancellation_guard(oracle_port, tuning, kernel):
    # data that the candidate fits to ~1e-7 of its amplitude: Syy, Sym and Smm agree to 14
    # digits and their combination would be rounding noise.  Those (data set, candidate) pairs
    # must be caught by the guard and recomputed in the direct form.
    N, nx, K = 4096, 200, 8
    x = numpy.linspace(400, 800, nx)
    rs = numpy.random.RandomState(5)
    pts = synth.parameter_points(K, seed=3)
    pts[:, 0] = 50.0                      # strong lines
    pts[:, 2] = rs.uniform(20, 60, size=K)
    y = rs.normal(0, 1e-2, size=(nx, N))
    fitted = numpy.arange(N) % 3 == 0     # every third data set is candidate (i % K) + tiny noise
    # model spectra computed once on the host and handed to both sides (the device exp and
    # libm's differ in the last bit, which is the whole residual of such a fit)
    spectra = numpy.array([p[0] * numpy.exp(-0.5 * ((p[1] - x) / p[2]) ** 2) for p in pts])
    for i in numpy.nonzero(fitted)[0]:
        y[:, i] = spectra[i % K] + rs.normal(0, 1e-6, size=nx)
    ds = ResidentDataset(x, y)
    ds.set_tuning(*tuning)
    got = ds.loglike_spectra(spectra, None, synth.NOISE_LEVEL, scale=1.0)
    assert _lib.load().mdns_last_kernel() == kernel
    allm = numpy.ones(N, dtype=bool)
    for k in range(K):
        want = oracle_port.clike_spectrum(spectra[k], y, synth.NOISE_LEVEL, allm)
        assert rel_err(got[k], want) < TOL_XP, k
    enabled, redo = ds.expanded_stats()
    assert redo >= fitted.sum() // K        # at least the perfectly fitted pairs
    # more than 2 % of the rows needed the direct form: the data set switches itself back
    assert not enabled
    ds.set_tuning(0, 0, 0, 0)
    again = ds.loglike_spectra(spectra, None, synth.NOISE_LEVEL, scale=1.0)
    assert _lib.load().mdns_last_kernel() not in (b'clike_xtile_kernel', b'slab_dmma_kernel') + TENSOR
    assert rel_err(again, got) < TOL_XP


@pytest.mark.parametrize('K', [1, 5])
def test_clike_very_long_spectra(oracle_port, K):
    # a model row of 30001 channels (240 KB) does not fit in shared memory: every batch size and
    # mask goes to the tensor-path kernel, which streams the model in 16-channel slices
    N, nx = 300, 30001
    x, y, _ = synth.horns(N, nx=nx, seed=6)
    ds = ResidentDataset(x, y)
    pts = synth.parameter_points(K, seed=2)
    for name in ('all', 'half'):
        m = synth.masks(N)[name]
        got = ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0)
        assert _lib.load().mdns_last_kernel() in TENSOR + TENSOR_GATHER
        for k in range(K):
            p = pts[k]
            want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
            assert rel_err(got[k], want) < TOL_XP, (name, k)


def test_clike_spectra_entry_point(oracle_port):
    N = 999
    x, y, _ = synth.horns(N)
    ds = ResidentDataset(None, y)          # no grid: spectra only
    rs = numpy.random.RandomState(4)
    spectra = rs.normal(0, 0.05, size=(4, 200))
    m = synth.masks(N)['sparse']
    got = ds.loglike_spectra(spectra, m, 0.03, scale=1.0)
    for k in range(4):
        want = oracle_port.clike_spectrum(spectra[k], y, 0.03, m)
        assert rel_err(got[k], want) < TOL
    with pytest.raises(_lib.MdnsError):
        ds.loglike_batch(numpy.ones((1, 3)), m, 0.03)      # needs the x grid


def test_clike_empty_mask_and_reuse(horns257):
    x, y, ds = horns257
    none = numpy.zeros(257, dtype=bool)
    got = ds.loglike_batch(synth.parameter_points(2), none, synth.NOISE_LEVEL)
    assert got.shape == (2, 0)
    one = none.copy()
    one[256] = True
    got = ds.loglike_batch(synth.parameter_points(2), one, synth.NOISE_LEVEL)
    assert got.shape == (2, 1) and numpy.isfinite(got).all()


def test_clike_mask_cache_and_changes(oracle_port):
    # the shim skips the upload/compaction of a mask it already holds; a changed mask of the same
    # size, the all-true mask and None must all take effect
    N = 3001
    x, y, _ = synth.horns(N, seed=2)
    ds = ResidentDataset(x, y)
    p = synth.parameter_points(2, seed=1)
    rs = numpy.random.RandomState(0)
    m1 = rs.uniform(size=N) < 0.4
    m2 = m1.copy()
    m2[5] = not m2[5]
    seq = [m1, m1, m2, m2, numpy.ones(N, dtype=bool), m1, None, m2, m1.copy()]
    for m in seq:
        got = ds.loglike_batch(p, m, synth.NOISE_LEVEL, scale=1.0)
        mm = numpy.ones(N, dtype=bool) if m is None else m
        assert got.shape == (2, int(mm.sum()))
        for k in range(2):
            want = oracle_port.clike(x, y, p[k][0], p[k][1], p[k][2], synth.NOISE_LEVEL, mm)
            assert rel_err(got[k], want) < TOL
    # staged draw: thresholds survive a repeated identical mask, not a different one
    Lm = numpy.full(int(m1.sum()), -1e300)
    ds.begin_draw(m1, Lm)
    ds.set_mask(m1)
    k, L, counts = ds.draw_batch(p, synth.NOISE_LEVEL)
    assert k == 0 and (counts == int(m1.sum())).all()
    ds.set_mask(m2)
    with pytest.raises(_lib.MdnsError):
        ds.draw_batch(p, synth.NOISE_LEVEL)


def test_clike_linearity_property_large():
    # size-independent property at a size the oracle would not finish quickly:
    # chi2(A=0) = sum (y/noise)^2 (plotevidences.py:17), checked with numpy column sums
    N = 200000
    x, y = synth.nothing(N, legacy=False)
    ds = ResidentDataset(x, y)
    got = ds.loglike_batch(numpy.array([[0.0, 500.0, 1.0]]), None, synth.NOISE_LEVEL, scale=1.0)[0]
    want = ((y / synth.NOISE_LEVEL) ** 2).sum(axis=0)
    assert rel_err(got, want) < TOL
    # masked evaluation equals the compaction of the full evaluation (bit for bit:
    # the per-data-set arithmetic does not depend on which rows are active)
    m = synth.masks(N)['half']
    p = synth.parameter_points(2)
    full = ds.loglike_batch(p, None, synth.NOISE_LEVEL).copy()
    part = ds.loglike_batch(p, m, synth.NOISE_LEVEL)
    assert numpy.array_equal(part, full[:, m])


@pytest.mark.parametrize('N', [500, 70000])
def test_first_accepted_matches_one_at_a_time_loop(oracle_port, N):
    # hiermetriclearn.py:181-196: candidates are tried in order until numpy.any(L > Lmins)
    x, y, _ = synth.horns(N, legacy=False, seed=9)
    ds = ResidentDataset(x, y)
    m = synth.masks(N)['half']
    n_act = int(m.sum())
    pts = synth.parameter_points(12, seed=4)
    Ls = numpy.array([-0.5 * oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m)
                      for p in pts])
    best = Ls.max(axis=0)
    # thresholds chosen so that candidates 0..4 are rejected everywhere and 5 is accepted
    Lmins = Ls[:5].max(axis=0) + 1e-6 * numpy.abs(Ls[:5].max(axis=0))
    want_counts = (Ls > Lmins).sum(axis=1)
    want_first = int(numpy.argmax(want_counts > 0)) if (want_counts > 0).any() else -1
    k, L, counts = ds.first_accepted(pts, m, Lmins, synth.NOISE_LEVEL)
    assert k == want_first and k >= 5
    assert numpy.array_equal(counts, want_counts)
    assert L.shape == (n_act,) and rel_err(L, Ls[k]) < TOL
    # nothing accepts above the best value of every data set
    k, L, counts = ds.first_accepted(pts, m, best + 1.0, synth.NOISE_LEVEL)
    assert k == -1 and L is None and (counts == 0).all()
    # everything accepts below the worst
    k, L, counts = ds.first_accepted(pts, m, Ls.min(axis=0) - 1.0, synth.NOISE_LEVEL)
    assert k == 0 and (counts == n_act).all() and rel_err(L, Ls[0]) < TOL
    # staged form: mask and thresholds once per draw, then batches of candidates
    assert ds.begin_draw(m, Lmins) == n_act
    k, L, counts = ds.draw_batch(pts[:4], synth.NOISE_LEVEL)
    assert k == -1 and L is None and numpy.array_equal(counts, want_counts[:4])
    k, L, counts = ds.draw_batch(pts[4:], synth.NOISE_LEVEL)
    assert k == want_first - 4 and numpy.array_equal(counts, want_counts[4:])
    assert rel_err(L, Ls[want_first]) < TOL
    # sparse form: only the accepting data sets of the first accepted candidate come back
    ds.begin_draw(m, Lmins)
    k, j, Lj, counts = ds.draw_batch_sparse(pts, synth.NOISE_LEVEL)
    assert k == want_first and numpy.array_equal(counts, want_counts)
    want_j = numpy.nonzero(Ls[want_first] > Lmins)[0]
    assert j.dtype == numpy.int32 and numpy.array_equal(j, want_j)
    assert rel_err(Lj, Ls[want_first][want_j]) < TOL
    # ... with a shorter mask afterwards (stale flags beyond the new length must not count)
    m_short = m.copy()
    m_short[N // 3:] = False
    n_short = int(m_short.sum())
    Ls_short = Ls[:, :n_short]
    ds.begin_draw(m_short, Lmins[:n_short])
    k, j, Lj, counts = ds.draw_batch_sparse(pts, synth.NOISE_LEVEL)
    wc = (Ls_short > Lmins[:n_short]).sum(axis=1)
    assert numpy.array_equal(counts, wc)
    if (wc > 0).any():
        kk = int(numpy.argmax(wc > 0))
        assert k == kk and numpy.array_equal(j, numpy.nonzero(Ls_short[kk] > Lmins[:n_short])[0])
    else:
        assert k == -1 and j is None
    ds.begin_draw(m, numpy.full(n_act, 1e300))
    assert ds.draw_batch_sparse(pts, synth.NOISE_LEVEL)[0] == -1
    # two-step form (one process per GPU): counts, then any candidate of the same launch
    from massivedatans_b200 import sharding
    ds.begin_draw(m, Lmins)
    c = ds.draw_counts(pts, synth.NOISE_LEVEL)
    assert numpy.array_equal(c, want_counts)
    kk, tot = sharding.global_first_accepted(c)          # no process group: the local decision
    assert kk == want_first and numpy.array_equal(tot, want_counts)
    for cand in (0, want_first, len(pts) - 1):
        assert rel_err(ds.fetch_candidate(cand), Ls[cand]) < TOL
    with pytest.raises(_lib.MdnsError):
        ds.fetch_candidate(len(pts))
    ds.set_mask(None)
    with pytest.raises(_lib.MdnsError):
        ds.draw_batch(pts, synth.NOISE_LEVEL)       # thresholds do not survive a new mask


def test_draw_pass_is_begin_draw_plus_draw_batch(oracle_port):
    # one native call per speculative pass; repeated masks / short thresholds are recognised and
    # not uploaded again -- the answers must not depend on what was recognised
    N = 3000
    x, y, _ = synth.horns(N, legacy=False, seed=10)
    ds = ResidentDataset(x, y)
    pts = synth.parameter_points(7, seed=6)
    masks = synth.masks(N)
    masks['single'] = numpy.arange(N) == 1234
    for name in ('half', 'single', 'sparse', 'all', 'none', 'single'):
        m = masks[name]
        n_act = int(m.sum())
        if n_act == 0:
            k, L, counts = ds.draw_pass(m, numpy.zeros(0), pts, synth.NOISE_LEVEL)
            assert k == -1 and L is None and (counts == 0).all()
            continue
        Ls = ds.loglike_batch(pts, m, synth.NOISE_LEVEL).copy()
        first3 = Ls[:3].max(axis=0)
        variants = [first3 + 1e-6 * numpy.abs(first3),      # candidates 0..2 rejected everywhere
                    Ls.min(axis=0) - 1.0,                    # everything accepts
                    Ls.max(axis=0) + 1.0]                    # nothing accepts
        for Lmins in variants + variants[:1]:                # ... and back to the first again
            want_counts = (Ls > Lmins).sum(axis=1)
            want_first = int(numpy.argmax(want_counts > 0)) if (want_counts > 0).any() else -1
            for repeat in range(2):                          # second time: nothing is uploaded
                k, L, counts = ds.draw_pass(m, Lmins.copy(), pts, synth.NOISE_LEVEL)
                assert k == want_first, (name, repeat)
                assert numpy.array_equal(counts, want_counts)
                if k >= 0:
                    assert numpy.array_equal(L, Ls[k])
                else:
                    assert L is None
            ds.begin_draw(m, Lmins)
            k2, L2, c2 = ds.draw_batch(pts, synth.NOISE_LEVEL)
            assert k2 == want_first and numpy.array_equal(c2, want_counts)
        # same thresholds, other candidates
        k, L, counts = ds.draw_pass(m, variants[0], pts[3:], synth.NOISE_LEVEL)
        assert numpy.array_equal(counts, (Ls[3:] > variants[0]).sum(axis=1))
    with pytest.raises(ValueError):
        ds.draw_pass(masks['half'], numpy.zeros(3), pts, synth.NOISE_LEVEL)


def test_legacy_like_symbol_accumulates(oracle_port):
    # the reference's own argtypes (sample.py:85-96) on the drop-in veneer
    import os
    from numpy.ctypeslib import ndpointer
    lib = ctypes.CDLL(os.path.join(_lib.DROPIN_DIR, 'clike.so'))
    lib.like.argtypes = [
        ndpointer(dtype=numpy.float64, ndim=1, flags='C_CONTIGUOUS'),
        ndpointer(dtype=numpy.float64, ndim=2, flags='C_CONTIGUOUS'),
        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
        ctypes.c_double,
        ndpointer(dtype=numpy.bool_, ndim=1, flags='C_CONTIGUOUS'),
        ndpointer(dtype=numpy.float64, ndim=1, flags='C_CONTIGUOUS')]
    N = 300
    x, y, _ = synth.horns(N)
    m = synth.masks(N)['half']
    for p in synth.parameter_points(3):
        Lout = numpy.zeros(m.sum())
        ret = lib.like(x, y, N, 200, p[0], p[1], p[2], 0.01, m, Lout)
        assert ret == 0
        want = oracle_port.clike(x, y, p[0], p[1], p[2], 0.01, m)
        assert rel_err(Lout, want) < TOL
    # accumulation (clike.c:72) and change detection of a re-used buffer
    Lout = numpy.full(m.sum(), 7.0)
    lib.like(x, y, N, 200, 0.5, 600., 3., 0.01, m, Lout)
    want = oracle_port.clike(x, y, 0.5, 600., 3., 0.01, m, Lout=numpy.full(m.sum(), 7.0))
    assert rel_err(Lout, want) < TOL
    y[:] = y[::-1].copy()
    Lout = numpy.zeros(m.sum())
    lib.like(x, y, N, 200, 0.5, 600., 3., 0.01, m, Lout)
    assert rel_err(Lout, oracle_port.clike(x, y, 0.5, 600., 3., 0.01, m)) < TOL
    _lib.load().mdns_legacy_reset()


@pytest.mark.parametrize('N,nx,K', [(3000, 200, 64), (70000, 200, 37), (2049, 57, 64), (500, 1000, 130),
                                    (129, 16, 5)])
def test_clike_tcgen05_int8_split_operand_path(oracle_port, N, nx, K):
    # the north star's tcgen05 experiment (tuning lanes = 5): cross term on the INT8 tensor path from
    # 7-bit digit planes of the FP64 operands (exact integer products in TMEM, FP64 recombination).
    # Stated tolerance: the guard keeps a result only if the a-priori bound of the dropped digits is
    # below xp_tol (1e-10) relative to chi2; measured error on horns data ~1e-14.
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N + 2)
    ds = ResidentDataset(x, y)
    ds.set_tuning(5, 0, 0, 0)
    pts = synth.parameter_points(K, seed=N + 3)
    got = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    ds.stage_params(pts)
    ds.set_mask(None)
    ds.launch_clike(synth.NOISE_LEVEL, 1.0)
    assert _lib.load().mdns_last_kernel() == b'clike_i8_kernel'
    again = numpy.empty((K, N))
    ds.fetch(again)
    allm = numpy.ones(N, dtype=bool)
    for k in sorted(set(k for k in (0, 7, 8, K // 2, K - 1) if k < K)):
        p = pts[k]
        want = oracle_port.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, allm)
        assert rel_err(again[k], want) < TOL_XP
    # a masked batch is not this path's business: it falls back to the automatic choice
    m = synth.masks(N)['half']
    got = ds.loglike_batch(pts[:3], m, synth.NOISE_LEVEL, scale=1.0)
    assert _lib.load().mdns_last_kernel() != b'clike_i8_kernel'
    assert rel_err(got[1], oracle_port.clike(x, y, pts[1][0], pts[1][1], pts[1][2], synth.NOISE_LEVEL, m)) < TOL_XP


def test_clike_tcgen05_int8_path_cancellation_guard(oracle_port):
    # data that a candidate fits to 1e-7 of its amplitude: the dropped digits would be the whole
    # residual; the guard must flag those data sets and the direct form must recompute them
    N, nx, K = 4096, 200, 8
    x = numpy.linspace(400, 800, nx)
    rs = numpy.random.RandomState(5)
    pts = synth.parameter_points(K, seed=3)
    pts[:, 0] = 50.0
    pts[:, 2] = rs.uniform(20, 60, size=K)
    y = rs.normal(0, 1e-2, size=(nx, N))
    fitted = numpy.arange(N) % 3 == 0
    spectra = numpy.array([p[0] * numpy.exp(-0.5 * ((p[1] - x) / p[2]) ** 2) for p in pts])
    for i in numpy.nonzero(fitted)[0]:
        y[:, i] = spectra[i % K] + rs.normal(0, 1e-6, size=nx)
    ds = ResidentDataset(x, y)
    ds.set_tuning(5, 0, 0, 0)
    ds.set_mask(None)
    ds.stage_spectra(spectra)
    ds.launch_clike(synth.NOISE_LEVEL, 1.0)
    assert _lib.load().mdns_last_kernel() == b'clike_i8_kernel'
    got = numpy.empty((K, N))
    ds.fetch(got)
    want = (((spectra[:, :, None] - y[None, :, :]) / synth.NOISE_LEVEL) ** 2).sum(axis=1)
    assert rel_err(got, want) < 1e-9
    assert rel_err(got[:, ~fitted], want[:, ~fitted]) < TOL_XP


@pytest.mark.parametrize('N,nx', [(5000, 200), (70001, 57)])
def test_dataset_from_npy_file_is_the_same_resident_data(tmp_path, N, nx):
    # the loader half of sample.py:27-31: the matrix goes file -> pinned block -> device, no host copy
    x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N)
    path = str(tmp_path / 'y.npy')
    numpy.save(path, y)
    a = ResidentDataset(x, y)
    b = ResidentDataset.from_npy(x, path)
    assert (b.nx, b.ndata) == (nx, N)
    pts = synth.parameter_points(5, seed=1)
    for mask in (None, synth.masks(N)['half']):
        assert numpy.array_equal(a.loglike_batch(pts, mask, synth.NOISE_LEVEL),
                                 b.loglike_batch(pts, mask, synth.NOISE_LEVEL))
    # MUSE-type: data + variance files, two shards' worth of columns read separately
    ym, vm, t = synth.muse(ndata=300, nspec=360)
    numpy.save(str(tmp_path / 'ym.npy'), ym)
    numpy.save(str(tmp_path / 'vm.npy'), vm)
    c = ResidentDataset(None, ym, variance=vm)
    d = ResidentDataset.from_npy(None, str(tmp_path / 'ym.npy'), str(tmp_path / 'vm.npy'))
    allm = numpy.ones(300, dtype=bool)
    La, Lb = numpy.zeros((1, 300)), numpy.zeros((1, 300))
    c.muse_loglike(t, allm, La)
    d.muse_loglike(t, allm, Lb)
    assert numpy.array_equal(La, Lb)
    numpy.save(str(tmp_path / 'bad.npy'), y.astype(numpy.float32))
    with pytest.raises(ValueError):
        ResidentDataset.from_npy(x, str(tmp_path / 'bad.npy'))
    h = ctypes.c_void_p()                                      # and the C entry point says so itself
    assert _lib.load().mdns_dataset_create_from_npy(None, str(tmp_path / 'bad.npy').encode(), None, None, 0,
                                                    ctypes.byref(h)) != 0
    assert b'float64' in _lib.load().mdns_last_error()


def test_legacy_like_sees_any_in_place_edit(oracle_port):
    # round 1 re-validated a cached matrix with 256 strided probes: an edit between the probes
    # was served stale.  The reference re-reads yy on every call (clike.c:72); the drop-in now
    # hashes every byte by default, and keeps at most two matrices resident.
    import os
    lib = ctypes.CDLL(os.path.join(_lib.DROPIN_DIR, 'clike.so'))
    lib.like.restype = ctypes.c_int
    core = _lib.load()
    core.mdns_legacy_reset()
    assert core.mdns_legacy_trust(-1) == 0            # default: full hash
    N, nx = 3000, 200
    x, y, _ = synth.horns(N)
    m = numpy.ones(N, dtype=bool)

    def like(yy):
        Lout = numpy.zeros(N)
        assert lib.like(x.ctypes.data_as(ctypes.c_void_p), yy.ctypes.data_as(ctypes.c_void_p), N, nx,
                        ctypes.c_double(0.5), ctypes.c_double(600.), ctypes.c_double(3.),
                        ctypes.c_double(0.01), m.ctypes.data_as(ctypes.c_void_p),
                        Lout.ctypes.data_as(ctypes.c_void_p)) == 0
        return Lout

    first = like(y)
    assert rel_err(first, oracle_port.clike(x, y, 0.5, 600., 3., 0.01, m)) < TOL
    step = (N * nx) // 256
    cell = 5 * step + 1234                            # between two of the old probes
    assert cell % step != 0 and cell != N * nx - 1
    y.reshape(-1)[cell] += 0.25
    second = like(y)
    want = oracle_port.clike(x, y, 0.5, 600., 3., 0.01, m)
    assert rel_err(second, want) < TOL and not numpy.array_equal(first, second)
    # the opt-out: the caller vouches for immutability, an edit between the probes now goes unnoticed
    assert core.mdns_legacy_trust(1) == 0
    assert numpy.array_equal(like(y), second)         # (validation mode changed: uploaded afresh)
    saved = y.reshape(-1)[cell]
    y.reshape(-1)[cell] = saved - 0.25
    stale = like(y)
    assert numpy.array_equal(stale, second)
    assert core.mdns_legacy_trust(0) == 1
    assert rel_err(like(y), oracle_port.clike(x, y, 0.5, 600., 3., 0.01, m)) < TOL
    assert not numpy.array_equal(like(y), second)
    # three matrices in turn: two stay resident, every answer is right
    others = [synth.horns(N, seed=s)[1] for s in (5, 6)]
    for _ in range(2):
        for yy in [y] + others:
            assert rel_err(like(yy), oracle_port.clike(x, yy, 0.5, 600., 3., 0.01, m)) < TOL
    core.mdns_legacy_reset()


# ------------------------------------------------------------------ MUSE ----
def test_muse_golden(golden):
    g = golden('cmuselike')
    ndata, nspec = int(g['ndata']), int(g['nspec'])
    y, v, _ = synth.muse(ndata=ndata, nspec=nspec)
    ds = ResidentDataset(None, y, variance=v)
    ypreds = numpy.array([synth.muse_template(nspec, phase=float(ph)) for ph in g['phases']])
    mask = g['mask']
    L = numpy.zeros((3, ndata))
    ds.muse_loglike(ypreds, mask, L)                      # batch of 3: expanded form, tensor path
    assert _lib.load().mdns_last_kernel().startswith(b'rows_dmma_kernel(raw')
    assert rel_err(L[:, mask], g['Lout'][:, mask]) < TOL_XP
    assert (L[:, ~mask] == 0).all()                       # cmuselike.c:49 -- untouched
    ds.muse_loglike(ypreds, None, L)
    assert rel_err(L, g['Lall']) < TOL_XP
    ds.set_expanded(False)                                # the direct two-pass kernels
    L = numpy.zeros((3, ndata))
    ds.muse_loglike(ypreds, mask, L)
    assert rel_err(L[:, mask], g['Lout'][:, mask]) < TOL
    assert (L[:, ~mask] == 0).all()
    ds.muse_loglike(ypreds, None, L)
    assert rel_err(L, g['Lall']) < TOL
    small = ResidentDataset(None, g['small_y'], variance=g['small_v'])
    Ls = numpy.full((1, 7), 5.0)
    small.muse_loglike(g['small_ypred'], g['small_mask'], Ls)
    sm = g['small_mask']
    assert rel_err(Ls[0, sm], g['small_Lout'][sm]) < TOL
    assert (Ls[0, ~sm] == 5.0).all()


@pytest.mark.parametrize('ndata,nspec', [(1, 2), (50, 37), (3000, 360), (20000, 64), (300, 3600),
                                         (4223, 3600), (6000, 1030), (2500, 8190)])
def test_muse_vs_oracle(oracle_port, ndata, nspec):
    y, v, t = synth.muse(ndata=ndata, nspec=nspec, seed=ndata)
    ds = ResidentDataset(None, y, variance=v)
    rs = numpy.random.RandomState(ndata)
    for mask in (numpy.ones(ndata, dtype=bool), rs.uniform(size=ndata) < 0.7):
        for yp in (t, synth.muse_template(nspec, phase=0.7)):
            L = numpy.zeros((1, ndata))
            ds.muse_loglike(yp, mask, L)
            want = oracle_port.cmuselike(y, v, yp, mask)
            assert rel_err(L[0][mask], want[mask]) < TOL


@pytest.mark.parametrize('lanes,unroll,ktile', [(256, 0, 1), (256, 0, 2), (256, 0, 4), (256, 0, 0),
                                                (8, 2, 0), (8, 16, 0), (32, 4, 0), (32, 16, 0),
                                                (0, 0, 0)])
@pytest.mark.parametrize('ndata,nspec', [(3, 2), (50, 37), (700, 360), (90, 3600), (20, 5001)])
def test_muse_kernel_variants_and_batches(oracle_port, lanes, unroll, ktile, ndata, nspec):
    y, v, t = synth.muse(ndata=ndata, nspec=nspec, seed=ndata + 1)
    ds = ResidentDataset(None, y, variance=v)
    ds.set_tuning(lanes, unroll, ktile, 0)
    K = 5
    ypreds = numpy.array([synth.muse_template(nspec, phase=0.3 * k) for k in range(K)])
    rs = numpy.random.RandomState(ndata)
    for mask in (numpy.ones(ndata, dtype=bool), rs.uniform(size=ndata) < 0.5):
        L = numpy.full((K, ndata), 3.0)
        ds.muse_loglike(ypreds, mask, L)
        xp = _lib.load().mdns_last_kernel().startswith(b'rows_dmma_kernel(raw')
        # automatic choice only: expanded form from K = 3 (spectra of two channels cancel so badly that
        # the feedback switches it off again after the first pass)
        assert not xp or lanes == 0
        assert xp or lanes != 0 or nspec <= 2
        for k in range(K):
            want = oracle_port.cmuselike(y, v, ypreds[k], mask)
            assert rel_err(L[k][mask], want[mask]) < (TOL_XP if xp else TOL)
            assert (L[k][~mask] == 3.0).all()


@pytest.mark.parametrize('groups', [1, 2])
@pytest.mark.parametrize('K', [1, 3])
def test_muse_block_kernel_many_rows_per_cta(oracle_port, groups, K):
    # more rows than resident CTAs: the ring is refilled many times; with two groups of threads
    # per CTA the data sets of a CTA alternate between the groups (per-group mbarriers)
    ndata, nspec = 5000, 2048
    y, v, t = synth.muse(ndata=ndata, nspec=nspec, seed=21)
    ds = ResidentDataset(None, y, variance=v)
    ds.set_tuning(256, 0, 0, groups)
    ypreds = numpy.array([synth.muse_template(nspec, phase=0.3 * k) for k in range(K)])
    mask = numpy.random.RandomState(4).uniform(size=ndata) < 0.8
    L = numpy.zeros((K, ndata))
    for _ in range(3):                       # repeated launches: no state is carried over
        ds.muse_loglike(ypreds, mask, L)
    assert _lib.load().mdns_last_kernel() == b'muse_block_kernel'
    for k in range(K):
        want = oracle_port.cmuselike(y, v, ypreds[k], mask)
        assert rel_err(L[k][mask], want[mask]) < TOL


@pytest.mark.parametrize('ndata,nspec,K', [(4223, 3600, 4), (4223, 3600, 16), (700, 360, 5), (50, 37, 3),
                                           (9000, 1030, 37), (300, 8190, 9)])
def test_muse_expanded_form_vs_oracle(oracle_port, ndata, nspec, K):
    # cmuselike in expanded form: S1 = (y/v).m and S2 = (1/v).m^2 as two raw contractions on the
    # FP64 tensor path (rows_dmma_kernel, stream-K over the channels), chi = Swyy - 2 s S1 + s^2 S2
    y, v, t = synth.muse(ndata=ndata, nspec=nspec, seed=ndata + 7)
    ds = ResidentDataset(None, y, variance=v)
    ypreds = numpy.array([synth.muse_template(nspec, phase=0.05 + 0.11 * k) for k in range(K)])
    rs = numpy.random.RandomState(ndata)
    lib = _lib.load()
    for mask in (numpy.ones(ndata, dtype=bool), rs.uniform(size=ndata) < 0.6):
        L = numpy.full((K, ndata), 3.0)
        for _ in range(2):                      # repeated launches: tickets and lists start clean
            ds.muse_loglike(ypreds, mask, L)
        assert lib.mdns_last_kernel().startswith(b'rows_dmma_kernel(raw')
        for k in sorted(set((0, 1, K // 2, K - 1))):
            want = oracle_port.cmuselike(y, v, ypreds[k], mask)
            assert rel_err(L[k][mask], want[mask]) < 1e-9      # the contract; the guard enforces 1e-10
            assert (L[k][~mask] == 3.0).all()
    assert ds.expanded_stats()[0]


def test_muse_expanded_form_cancellation_guard(oracle_port):
    # a candidate that IS the template the cube was built from: chi is the noise alone, a few
    # thousandths (and less) of Swyy.  The guard must send what it cannot vouch for to the direct
    # fix-up, and data that needs it for more than 2 % of its rows switches the path off.
    ndata, nspec, K = 3000, 3600, 4
    y, v, t = synth.muse(ndata=ndata, nspec=nspec, seed=11)
    y[:, ::7] = (t.reshape((-1, 1)) * numpy.linspace(100.0, 2000.0, len(y[0, ::7]))
                 + numpy.random.RandomState(2).normal(size=(nspec, len(y[0, ::7]))) * numpy.sqrt(v[:, ::7]))
    ds = ResidentDataset(None, y, variance=v)
    ypreds = numpy.array([t] + [synth.muse_template(nspec, phase=0.2 * k) for k in range(1, K)])
    mask = numpy.ones(ndata, dtype=bool)
    L = numpy.zeros((K, ndata))
    ds.muse_loglike(ypreds, mask, L)
    assert _lib.load().mdns_last_kernel().startswith(b'rows_dmma_kernel(raw')
    for k in range(K):
        want = oracle_port.cmuselike(y, v, ypreds[k], mask)
        assert rel_err(L[k], want) < 1e-9
    enabled, redo = ds.expanded_stats()
    assert redo >= ndata // 7 - 1            # the high signal-to-noise spectra were recomputed
    assert not enabled                       # ... more than 2 % of the rows: direct kernels from now on
    ds.muse_loglike(ypreds, mask, L)
    assert _lib.load().mdns_last_kernel() == b'muse_block_kernel'
    for k in range(K):
        want = oracle_port.cmuselike(y, v, ypreds[k], mask)
        assert rel_err(L[k], want) < TOL


def test_muse_callable_matches_reference_wrapper(oracle_port):
    ndata, nspec = 200, 360
    y, v, t = synth.muse(ndata=ndata, nspec=nspec)

    def model(scale, phase):
        return scale * synth.muse_template(nspec, phase=phase)

    f = make_muse_loglikelihood(y, v, model)
    mask = numpy.random.RandomState(1).uniform(size=ndata) < 0.5
    numpy.random.seed(5)
    got = f((2.0, 0.1), mask)
    numpy.random.seed(5)
    jitter = numpy.random.normal(0, 1e-5, size=mask.sum())    # musefuse.py:535
    want = oracle_port.cmuselike(y, v, model(2.0, 0.1), mask)[mask] + jitter
    assert numpy.allclose(got, want, rtol=1e-12, atol=0)
    assert (f((0.0, 0.0), mask) == -1e100).all()             # musefuse.py:528-530
    # legacy symbol with the reference argtypes (musefuse.py:509-517)
    import os
    from numpy.ctypeslib import ndpointer
    lib = ctypes.CDLL(os.path.join(_lib.DROPIN_DIR, 'cmuselike.so'))
    lib.like.argtypes = [
        ndpointer(dtype=numpy.float64, ndim=2, flags='C_CONTIGUOUS'),
        ndpointer(dtype=numpy.float64, ndim=2, flags='C_CONTIGUOUS'),
        ndpointer(dtype=numpy.float64, ndim=1, flags='C_CONTIGUOUS'),
        ndpointer(dtype=numpy.bool_, ndim=1, flags='C_CONTIGUOUS'),
        ctypes.c_int, ctypes.c_int,
        ndpointer(dtype=numpy.float64, ndim=1, flags='C_CONTIGUOUS')]
    Lout = numpy.zeros(ndata)
    assert lib.like(y, v, t, mask, ndata, nspec, Lout) == 0
    want = oracle_port.cmuselike(y, v, t, mask)
    assert rel_err(Lout[mask], want[mask]) < TOL and (Lout[~mask] == 0).all()
    _lib.load().mdns_legacy_reset()


# ------------------------------------------------------------ multi-GPU ----
def _ndev():
    return _lib.load().mdns_device_count()


@pytest.mark.skipif(_ndev() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('ndev', [2, 4, 8])
def test_sharded_dataset_matches_single_device(oracle_port, ndev):
    if _ndev() < ndev:
        pytest.skip('needs %d GPUs' % ndev)
    N = 40003                       # not divisible: ragged shards
    x, y, _ = synth.horns(N, seed=5)
    single = ResidentDataset(x, y, devices=[0])
    multi = ResidentDataset(x, y, devices=list(range(ndev)))
    pts = synth.parameter_points(9, seed=2)
    for name, m in synth.masks(N, seed=3).items():
        a = single.loglike_batch(pts, m, synth.NOISE_LEVEL).copy()
        b = multi.loglike_batch(pts, m, synth.NOISE_LEVEL)
        assert a.shape == b.shape == (9, int(m.sum()))
        assert rel_err(b, a) < TOL, name
    m = synth.masks(N, seed=3)['half']
    got = multi.loglike_batch(pts[:2], m, synth.NOISE_LEVEL, scale=1.0)
    for k in range(2):
        want = oracle_port.clike(x, y, pts[k][0], pts[k][1], pts[k][2], synth.NOISE_LEVEL, m)
        assert rel_err(got[k], want) < TOL
    # MUSE-type, un-compacted output across shards
    ym, vm, t = synth.muse(ndata=1001, nspec=360)
    dm = ResidentDataset(None, ym, variance=vm, devices=list(range(ndev)))
    mask = numpy.random.RandomState(1).uniform(size=1001) < 0.6
    L = numpy.zeros((1, 1001))
    dm.muse_loglike(t, mask, L)
    want = oracle_port.cmuselike(ym, vm, t, mask)
    assert rel_err(L[0][mask], want[mask]) < TOL and (L[0][~mask] == 0).all()


def test_clike_full_size_properties():
    # BASELINE full size (1e6 data sets x 200 channels, the bench configuration) through
    # size-independent properties, on the automatically chosen kernels:
    N = 1000000
    x, y = synth.nothing(N, legacy=False)
    ds = ResidentDataset(x, y)
    lib = _lib.load()
    pts = synth.parameter_points(16, seed=3)
    pts[5, 0] = 0.0                                    # a candidate without a line
    full = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0).copy()
    assert lib.mdns_last_kernel() in TENSOR
    # (1) chi2(A=0) = sum (y/noise)^2 (plotevidences.py:17), numpy column sums in blocks
    want = numpy.empty(N)
    for lo in range(0, N, 100000):
        want[lo:lo + 100000] = ((y[:, lo:lo + 100000] / synth.NOISE_LEVEL) ** 2).sum(axis=0)
    assert rel_err(full[5], want) < TOL_XP
    # (2) batch == one candidate at a time on the streaming kernel (direct form)
    for k in (0, 5, 15):
        one = ds.loglike_batch(pts[k:k + 1], None, synth.NOISE_LEVEL, scale=1.0)[0]
        assert lib.mdns_last_kernel() == b'clike_rows_kernel'
        assert rel_err(full[k], one) < TOL_XP
    # (3) a masked evaluation is the compaction of the full one (direct-form kernels on both
    # sides: bit for bit)
    ds.set_expanded(False)
    direct = ds.loglike_batch(pts[:8], None, synth.NOISE_LEVEL, scale=1.0).copy()
    assert rel_err(direct, full[:8]) < TOL_XP
    m = synth.masks(N)['half']
    part = ds.loglike_batch(pts[:3], m, synth.NOISE_LEVEL, scale=1.0)
    one = ds.loglike_batch(pts[:1], None, synth.NOISE_LEVEL, scale=1.0)[0]
    assert numpy.array_equal(part[0], one[m])
    # (4) idempotence: the same launch twice gives the same bits
    ds.set_expanded(True)
    again = ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)
    assert numpy.array_equal(again, full)
