"""The REFERENCE's own sampler stack (multi_nested_integrator.py, multi_nested_sampler.py,
cachedconstrainer.py, hiermetriclearn.py, clustering/*.py -- imported as bytecode compiled from
the unmodified files: oracle/compile_pyref.py, oracle/pyref_loader.py) running on the product path on a GPU:

* its `clustering/neighbors.py` loads `cneighbors.so` from its own directory (neighbors.py:97-99):
  the file placed there is the zero-edit drop-in veneer massivedatans_b200/dropin/cneighbors.so,
  i.e. every RadFriends query of the run goes through libmdns_b200.so;
* its `multi_loglikelihood(params, data_mask)` callable (sample.py:101-108) is the one
  massivedatans_b200.likelihood.make_multi_loglikelihood returns.

Fixed seed, same data and settings as tests/golden/make_golden_sampler.py, which ran the same
stack on the reference's own C libraries: evidences must agree within nested-sampling noise
(north star), and -- the neighbour answers being bit-exact and logL agreeing to ~1e-13 -- the run
retraces the golden one draw by draw unless an accept test falls on a tie."""
import contextlib
import io
import os
import shutil
import sys

import numpy
import pytest

from conftest import ROOT
from massivedatans_b200 import _lib, synth
from massivedatans_b200.likelihood import make_multi_loglikelihood

pytestmark = pytest.mark.gpu

PYREF = os.path.join(ROOT, 'oracle', '_ref', 'pyref')


@pytest.mark.timeout(1500)
def test_reference_sampler_stack_runs_on_the_shim(golden):
    if not os.path.exists(os.path.join(PYREF, 'multi_nested_sampler.bc')):
        pytest.skip('reference bytecode not built (make -C oracle where /root/reference exists)')
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden_sampler as mg          # its stubs and settings; nothing of it touches /root/reference here
    g = golden('sampler_run')
    lib = _lib.load()                         # RTLD_GLOBAL: the veneer resolves libmdns_b200.so by soname
    _lib.require_device()
    shutil.copy(os.path.join(_lib.DROPIN_DIR, 'cneighbors.so'), os.path.join(PYREF, 'clustering', 'cneighbors.so'))
    os.environ['USE_GRAPH'] = '0'
    os.environ.pop('OMP_NUM_THREADS', None)
    mg.stub_modules()
    x, y, truth = synth.horns(int(g['ndata']), seed=int(g['seed_data']))
    multi_loglikelihood = make_multi_loglikelihood(x, y, synth.NOISE_LEVEL)

    def priortransform(cube):               # sample.py:52-58
        cube = cube.copy()
        cube[0] = 10 ** (cube[0] * 2 - 2)
        cube[1] = cube[1] * 400 + 400
        cube[2] = cube[2] * 2
        return cube

    for name in [m for m in sys.modules if m == 'clustering' or m.startswith('clustering.')]:
        del sys.modules[name]               # the reference's package, not ours, under that name
    from oracle import pyref_loader
    finder = pyref_loader.install(PYREF)
    try:
        import cachedconstrainer
        from cachedconstrainer import (CachedConstrainer, MetricLearningFriendsConstrainer,
                                       generate_individual_constrainer)
        from multi_nested_integrator import multi_nested_integrator
        from multi_nested_sampler import MultiNestedSampler
        import clustering.neighbors as nb
        assert os.path.dirname(os.path.abspath(nb.__file__)) == os.path.join(PYREF, 'clustering')
        assert nb.bootstrapped_maxdistance is not None, 'the drop-in cneighbors.so did not load'
        launches0 = lib.mdns_launch_count()

        def fresh():                        # sample.py:133-137
            c = MetricLearningFriendsConstrainer(metriclearner='truncatedscaling', force_shrink=True,
                                                 rebuild_every=1000, metric_rebuild_every=20,
                                                 verbose=False)
            c.prev_maxdistance = mg._NoPrevious()
            return c

        cachedconstrainer.generate_fresh_constrainer = fresh
        superset = fresh()
        cc = CachedConstrainer()
        _, _, individual_draw_constrained = generate_individual_constrainer()
        numpy.random.seed(int(g['seed_run']))
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
            sampler = MultiNestedSampler(nlive_points=int(g['nlive']), priortransform=priortransform,
                                         multi_loglikelihood=multi_loglikelihood, ndim=3,
                                         ndata=y.shape[1],
                                         superset_draw_constrained=superset.draw_constrained,
                                         individual_draw_constrained=individual_draw_constrained,
                                         draw_constrained=cc.get, nsuperset_draws=10,
                                         use_graph=False)
            superset.sampler = sampler
            cc.sampler = sampler
            results = multi_nested_integrator(tolerance=0.5, multi_sampler=sampler,
                                              min_samples=0, max_samples=0)
    finally:
        pyref_loader.uninstall(finder)
        for name in [m for m in sys.modules if m == 'clustering' or m.startswith('clustering.')]:
            del sys.modules[name]
    logZ = numpy.asarray(results['logZ'], dtype=float)
    logZerr = numpy.asarray(results['logZerr'], dtype=float)
    launches = lib.mdns_launch_count() - launches0
    assert launches > int(g['ndraws'])          # every likelihood call and every neighbour query ran on the GPU
    # evidences within nested-sampling noise of the reference run on the reference's C
    z = (logZ - g['logZ']) / numpy.sqrt(logZerr ** 2 + g['logZerr'] ** 2)
    assert numpy.abs(z).max() < 4.0, z
    retraced = int(sampler.ndraws) == int(g['ndraws']) and int(results['niterations']) == int(g['niterations'])
    print('reference stack on the shim: ndraws %d (golden %d), max |z| %.3g, retraced draw by draw: %s, '
          '%d kernel launches' % (sampler.ndraws, int(g['ndraws']), numpy.abs(z).max(), retraced, launches))
    if retraced:
        assert numpy.allclose(logZ, g['logZ'], rtol=0, atol=1e-6)
