import numpy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from massivedatans_b200 import synth, _lib
from massivedatans_b200.likelihood import ResidentDataset
N, nx, K = 700, 203, 9
x, y, _ = synth.horns(N, nx=nx, legacy=False, seed=N + 21)
pts = synth.parameter_points(K, seed=N + 22)
m = synth.masks(N, seed=N)['half']
spectra = numpy.array([p[0] * numpy.exp(-0.5 * ((p[1] - x) / p[2]) ** 2) for p in pts])
want = (((spectra[:, :, None] - y[None, :, :][:, :, m]) / synth.NOISE_LEVEL) ** 2).sum(axis=1)
for tun in ((6, 0, 16, 2), (3, 0, 16, 3), (6, 0, 16, 2)):
    ds = ResidentDataset(x, y)
    ds.set_tuning(*tun)
    for rep in range(2):
        got = numpy.array(ds.loglike_batch(pts, m, synth.NOISE_LEVEL, scale=1.0))
        rel = numpy.abs(got - want) / numpy.abs(want)
        bad = numpy.nonzero(rel > 1e-9)
        print(tun, rep, _lib.load().mdns_last_kernel(), 'max rel', rel.max(), 'bad', list(zip(bad[0][:10], bad[1][:10])), ds.expanded_stats())
        if len(bad[0]):
            k, r = bad[0][0], bad[1][0]
            print('   got', got[k, r], 'want', want[k, r], 'diff', got[k, r] - want[k, r], 'Smm/noise2', (spectra[k] ** 2).sum() / synth.NOISE_LEVEL ** 2, 'Syy/noise2', (y[:, numpy.nonzero(m)[0][r]] ** 2).sum() / synth.NOISE_LEVEL ** 2)
