"""MUSE stellar-population model (musefuse.py:222-284) on the device.

Pin: tests/golden/muse_model.npz holds spectra computed by the reference's OWN `model()`, `ages`
table and Calzetti block (executed from the reference's source text by
tests/golden/make_golden_muse_model.py) on the seeded synthetic template grids of
massivedatans_b200/synth.py.  CPU tests check the oracle restatement (oracle.np.muse_model,
bit for bit) and the package's host-side Calzetti table against it; GPU tests check the device
model against the fixture and the oracle (1e-12 relative: the only differences are exp/pow of
the device libm vs glibc) and the fused model + cmuselike path against the CPU oracle (1e-9).
"""
import os

import numpy
import pytest

from massivedatans_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'muse_model.npz')


@pytest.fixture(scope='module')
def fixture():
    g = numpy.load(GOLDEN)
    Zs, _, wl_A, grids = synth.muse_grids(nwave=int(g['nwave']))
    return dict(params=g['params'], spectra=g['spectra'], ages=g['ages'], calzetti=g['calzetti'],
                Zs=Zs, model_wavelength=wl_A / 10., grids=grids,
                wavelength=synth.muse_wavelength(int(g['nspec'])) / 10., nspec=int(g['nspec']))


def rel(got, want):
    return numpy.max(numpy.abs(got - want) / numpy.maximum(numpy.abs(want), 1e-300))


def oracle_spectra(f, params, wavelength=None):
    from oracle import np as onp
    wl = f['wavelength'] if wavelength is None else wavelength
    return numpy.array([onp.muse_model(f['Zs'], f['ages'], f['model_wavelength'], f['calzetti'],
                                       f['grids'], wl, p[0], 10 ** p[1], p[2], p[3], p[4])
                        for p in params])


# ------------------------------------------------------------------------------------- CPU
def test_oracle_model_reproduces_the_reference_function(fixture):
    from oracle import np as onp
    assert numpy.array_equal(onp.calzetti(fixture['model_wavelength']), fixture['calzetti'])
    got = oracle_spectra(fixture, fixture['params'])
    assert numpy.array_equal(got, fixture['spectra'])
    assert (fixture['spectra'] >= 0).all() and fixture['spectra'].any(axis=1).all()


def test_host_calzetti_table_matches_the_reference(fixture):
    from massivedatans_b200.likelihood import calzetti
    assert numpy.array_equal(calzetti(fixture['model_wavelength']), fixture['calzetti'])


# ------------------------------------------------------------------------------------- GPU
def device_model(f, ndata, nspec=None, wavelength=None, grids=None):
    from massivedatans_b200.likelihood import DeviceMuseModel, ResidentDataset
    nspec = f['nspec'] if nspec is None else nspec
    y, v, _ = synth.muse(ndata=ndata, nspec=nspec)
    ds = ResidentDataset(None, y, variance=v)
    wl = f['wavelength'] if wavelength is None else wavelength
    m = DeviceMuseModel(ds, f['grids'] if grids is None else grids, f['Zs'], f['ages'],
                        f['model_wavelength'], wl)
    return ds, m, y, v


def as_model_args(params):
    q = numpy.array(params, dtype=float)
    q[:, 1] = 10 ** q[:, 1]
    return q


@pytest.mark.gpu
def test_device_model_matches_the_reference_spectra(fixture):
    ds, m, _, _ = device_model(fixture, ndata=64)
    nonzero = m.stage(as_model_args(fixture['params']))
    got = m.spectra()
    assert nonzero.all()
    assert got.shape == fixture['spectra'].shape
    assert rel(got, fixture['spectra']) < 1e-12
    # one point through the reference's call signature
    p = as_model_args(fixture['params'][5:6])[0]
    assert rel(m(*p), fixture['spectra'][5]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('K', [1, 7, 40])
def test_device_model_full_cube_shape(fixture, K):
    wl = synth.muse_wavelength(synth.MUSE_NSPEC) / 10.
    ds, m, _, _ = device_model(fixture, ndata=96, nspec=synth.MUSE_NSPEC, wavelength=wl)
    params = synth.muse_parameter_points(K, seed=100 + K)
    m.stage(as_model_args(params))
    assert rel(m.spectra(), oracle_spectra(fixture, params, wl)) < 1e-12


@pytest.mark.gpu
def test_model_and_likelihood_fused_on_the_device(fixture):
    from oracle import port
    ndata = 300
    ds, m, y, v = device_model(fixture, ndata=ndata)
    params = synth.muse_parameter_points(9, seed=77)
    mask = synth.masks(ndata)['half']
    Lout = numpy.full((len(params), ndata), 123.0)
    nonzero = m.loglike_batch(as_model_args(params), mask, Lout)
    assert nonzero.all()
    for k, spec in enumerate(oracle_spectra(fixture, params)):
        want = port.cmuselike(y, v, spec, mask)
        assert rel(Lout[k][mask], want[mask]) < 1e-9
        assert (Lout[k][~mask] == 123.0).all()          # cmuselike.c:49 leaves them alone


@pytest.mark.gpu
def test_callable_with_guard_jitter_and_batch(fixture):
    from oracle import port
    from massivedatans_b200.likelihood import make_muse_loglikelihood_device
    ndata = 200
    y, v, _ = synth.muse(ndata=ndata, nspec=fixture['nspec'])
    grids = fixture['grids'].copy()
    grids[0] = 0.0                                       # a metallicity bin with no stars
    f = make_muse_loglikelihood_device(y, v, grids, fixture['Zs'], fixture['ages'],
                                       fixture['model_wavelength'], fixture['wavelength'])
    mask = synth.masks(ndata)['half']
    params = synth.muse_parameter_points(6, seed=5)
    params[2, 0] = fixture['Zs'][0] + 0.01               # falls into the empty bin
    fx = dict(fixture, grids=grids)
    specs = oracle_spectra(fx, params)
    empty = [k for k in range(len(params)) if not specs[k].any()]      # by chance others too
    assert 2 in empty and len(empty) < len(params)
    numpy.random.seed(3)
    got = [f(p, mask) for p in params]
    numpy.random.seed(3)
    for k, p in enumerate(params):
        if k in empty:
            assert (got[k] == -1e100).all() and got[k].shape == (mask.sum(),)   # musefuse.py:528-530
            continue
        want = port.cmuselike(y, v, specs[k], mask)[mask] + numpy.random.normal(0, 1e-5, size=mask.sum())
        assert numpy.max(numpy.abs(got[k] - want) / numpy.abs(want)) < 1e-9
    L = f.batch(params, mask)
    assert L.shape == (6, mask.sum())
    for k in range(len(params)):
        if k in empty:
            assert (L[k] == -1e100).all()
            continue
        want = port.cmuselike(y, v, specs[k], mask)[mask]
        assert rel(L[k], want) < 1e-9


@pytest.mark.gpu
def test_model_on_a_sharded_cube(fixture):
    # two shards (on one device if the box has a single GPU): every shard builds the spectra
    # for its own pass
    from oracle import port
    from massivedatans_b200 import _lib
    from massivedatans_b200.likelihood import make_muse_loglikelihood_device
    ndata = 257
    y, v, _ = synth.muse(ndata=ndata, nspec=fixture['nspec'])
    devices = [0, 1] if _lib.load().mdns_device_count() >= 2 else [0, 0]
    f = make_muse_loglikelihood_device(y, v, fixture['grids'], fixture['Zs'], fixture['ages'],
                                       fixture['model_wavelength'], fixture['wavelength'],
                                       jitter=0, devices=devices)
    mask = synth.masks(ndata)['half']
    params = synth.muse_parameter_points(5, seed=12)
    L = f.batch(params, mask)
    for k, spec in enumerate(oracle_spectra(fixture, params)):
        assert rel(L[k], port.cmuselike(y, v, spec, mask)[mask]) < 1e-9
    assert rel(f(params[3], mask), L[3]) < 1e-14


@pytest.mark.gpu
def test_model_argument_errors(fixture):
    from massivedatans_b200 import _lib
    from massivedatans_b200.likelihood import DeviceMuseModel
    ds, m, _, _ = device_model(fixture, ndata=32)
    bad = as_model_args(fixture['params'][:2])
    bad[1, 0] = fixture['Zs'][0] - 1.0                   # the reference raises IndexError here
    with pytest.raises(_lib.MdnsError):
        m.stage(bad)
    with pytest.raises(ValueError):
        DeviceMuseModel(ds, fixture['grids'], fixture['Zs'], fixture['ages'],
                        fixture['model_wavelength'], fixture['wavelength'][:-1])
    with pytest.raises(_lib.MdnsError):
        DeviceMuseModel(ds, fixture['grids'], fixture['Zs'], fixture['ages'],
                        fixture['model_wavelength'][::-1], fixture['wavelength'])
