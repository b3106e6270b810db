"""One parity test per BASELINE.json configuration (the bench line is configs[2]/[3]'s shape;
"the other configs are parity-test cases").  Every test runs the CUDA path through the public
mirrors on the configuration's data shape and compares with the CPU oracle (oracle.port: the C
restatement that is bit-identical to the reference's own .so files) on the same inputs --
logL within 1e-9 relative (the kernel tests in test_gpu_likelihood.py assert 1e-12 / 1e-10 per
kernel family), neighbour outputs exact.

  configs[0]  gensimple_horns.py 10000 -> sample.py data_widths_10000.hdf5 100   (README run)
  configs[1]  gennothing.py 10000 -> sample.py data_nothing_10000.hdf5 10000     (all active)
  configs[2]  gensimple_horns 100000 data sets, 400 live points, single B200
  configs[3]  gen_realistic.py spectra, 1e6 data sets over 8 GPUs -> one GPU's shard:
              125 000 data sets x 1000 channels (oracle on a scattered sample of the rows)
  configs[4]  MUSE cube of the reference shape, 3600 channels x 4223 data sets
"""
import numpy
import pytest

from massivedatans_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-9           # the contract (BASELINE north_star)


def rel(got, want):
    return numpy.max(numpy.abs(got - want) / numpy.abs(want))


def sample_points(K, seed):
    """K parameter vectors (A, mu, log_sig) as the sampler proposes them (sample.py:52-58)."""
    p = synth.parameter_points(K, seed=seed)        # (A, mu, sig)
    p[:, 2] = numpy.log10(p[:, 2])
    return p


def check_likelihood(x, y, oracle_port, masks, Ks, rows=None):
    from massivedatans_b200.likelihood import make_multi_loglikelihood
    f = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)
    for name, mask in masks.items():
        for K in Ks:
            pts = sample_points(K, seed=K)
            got = f.batch(pts, mask)
            assert got.shape == (K, int(mask.sum()))
            one = f(pts[K - 1], mask)                 # the reference-shaped call
            if not mask.any():
                assert one.shape == (0,)
                continue
            assert rel(one, got[K - 1]) < TOL
            for k in sorted({0, K - 1}):
                A, mu, log_sig = pts[k]
                if rows is None:
                    want = -0.5 * oracle_port.clike(x, y, A, mu, 10 ** log_sig, synth.NOISE_LEVEL, mask)
                    assert rel(got[k], want) < TOL, (name, K, k)
                else:
                    # oracle on a sample of the active rows only (the full pass takes minutes)
                    sel = numpy.zeros(len(mask), dtype=bool)
                    sel[rows] = True
                    sel &= mask
                    want = -0.5 * oracle_port.clike(x, y, A, mu, 10 ** log_sig, synth.NOISE_LEVEL, sel)
                    pos = numpy.cumsum(mask)[sel] - 1     # slots of the sampled rows in `got`
                    assert rel(got[k][pos], want) < TOL, (name, K, k)
    return f


def check_region(members, oracle_port, nboot=10):
    from massivedatans_b200.clustering import neighbors
    numpy.random.seed(4)
    r = neighbors.bootstrapped_maxdistance(members, nboot)
    numpy.random.seed(4)
    chosen = synth.bootstrap_chosen(len(members), nboot)
    assert r == oracle_port.bootstrapped_maxdistance_chosen(members, chosen)
    cand = numpy.random.RandomState(9).uniform(members.min(), members.max(), size=(1000, members.shape[1]))
    assert numpy.array_equal(neighbors.count_within_distance_of(members, r, cand),
                             oracle_port.count_within_distance_of(members, r, cand))
    assert numpy.array_equal(neighbors.any_within_distance_of(members, r, cand),
                             oracle_port.any_within_distance_of(members, r, cand))
    assert neighbors.is_within_distance_of(members, r, cand[0]) == \
        bool(oracle_port.is_within_distance_of(members, r, cand[0]))
    assert neighbors.most_distant_nearest_neighbor(members) == \
        oracle_port.most_distant_nearest_neighbor(members)


def test_config0_readme_run(oracle_port):
    # sample.py:27-31 reads the 10 000-column file and keeps the first `ndata` = 100 columns
    x, y, _ = synth.horns(10000)
    y = numpy.ascontiguousarray(y[:, :100])
    masks = synth.masks(100)
    masks['single'] = numpy.arange(100) == 37
    check_likelihood(x, y, oracle_port, masks, Ks=(1, 8))
    # 400 live points in the 3-d unit cube (sample.py:165)
    check_region(numpy.random.RandomState(1).uniform(size=(400, 3)), oracle_port)


def test_config1_nothing_all_active(oracle_port):
    N = 10000
    x, y = synth.nothing(N)
    allm = numpy.ones(N, dtype=bool)
    f = check_likelihood(x, y, oracle_port, {'all': allm}, Ks=(1, 16))
    # a vanishing line reproduces the analytic no-signal value (plotevidences.py:17)
    got = f((1e-300, 600.0, 1.0), allm)
    assert rel(got, (-0.5 * (y / synth.NOISE_LEVEL) ** 2).sum(axis=0)) < TOL


def test_config2_horns_100000_with_400_live_points(oracle_port):
    N = 100000
    x, y, _ = synth.horns(N, legacy=False, seed=7)
    masks = {'all': numpy.ones(N, dtype=bool), 'half': synth.masks(N)['half']}
    check_likelihood(x, y, oracle_port, masks, Ks=(1, 16))
    # region over the union of the live points of many data sets (400 per data set; the union
    # observed in the reference run reaches ~5000 members)
    members = numpy.random.RandomState(2).uniform(size=(5000, 3)) * [1.0, 0.25, 0.5]
    check_region(members, oracle_port)


def test_bench_configuration_sampled_rows_vs_oracle(oracle_port):
    # the configuration bench.py times by default, on the kernel it times: 1e6 horns data sets x
    # 200 channels (bench.make_inputs: seed 1000), 16 candidates (seed 7), all active ->
    # slab_dmma_kernel; every candidate compared with the oracle on 2000 sampled data sets
    from massivedatans_b200 import _lib
    from massivedatans_b200.likelihood import ResidentDataset
    N, K = 1000000, 16
    x, y, _ = synth.horns(N, nx=200, legacy=False, seed=1000)
    pts = synth.parameter_points(K, seed=7)
    ds = ResidentDataset(x, y)
    ds.set_mask(None)
    ds.stage_params(pts)
    for _ in range(2):                       # the second launch replays the captured graph
        ds.launch_clike(synth.NOISE_LEVEL, -0.5)
    assert _lib.load().mdns_last_kernel() == b'slab_dmma_kernel'
    got = numpy.empty((K, N))
    ds.fetch(got)
    sel = numpy.zeros(N, dtype=bool)
    sel[numpy.random.RandomState(5).permutation(N)[:2000]] = True
    sel[[0, 1, N - 2, N - 1]] = True
    for k in range(K):
        A, mu, sig = pts[k]
        want = -0.5 * oracle_port.clike(x, y, A, mu, sig, synth.NOISE_LEVEL, sel)
        assert rel(got[k][sel], want) < TOL, k
    ds.close()


def test_config3_realistic_shard_of_one_gpu(oracle_port):
    N, nx = 125000, 1000
    x, y, _ = synth.realistic(N, nx=nx)
    rows = numpy.random.RandomState(3).permutation(N)[:1500]
    masks = {'all': numpy.ones(N, dtype=bool), 'tenth': synth.masks(N)['half'] & (numpy.arange(N) % 5 == 0)}
    check_likelihood(x, y, oracle_port, masks, Ks=(1, 16), rows=rows)


def test_config4_muse_cube_of_the_reference_shape(oracle_port):
    from massivedatans_b200.likelihood import make_muse_loglikelihood
    y, v, template = synth.muse()
    assert y.shape == (3600, 4223)
    models = {0: template, 1: synth.muse_template(phase=0.7)}
    f = make_muse_loglikelihood(y, v, lambda which: models[which], jitter=0)
    for mask in (numpy.ones(4223, dtype=bool), synth.masks(4223)['half']):
        for which in models:
            got = f((which,), mask)
            want = oracle_port.cmuselike(y, v, models[which], mask)[mask]
            assert rel(got, want) < TOL
