"""Generate tests/golden/*.npz from the UNMODIFIED reference C (oracle/_ref/*.so).

Run in the dev container (needs /root/reference to have been compiled by
`make -C oracle`):   python tests/golden/make_golden.py
The fixtures are small, committed, and travel to the GPU box, where
/root/reference does not exist.  Inputs come from massivedatans_b200.synth
(seeded restatements of the reference generators) so the files only need to
store parameters + reference outputs, plus the raw inputs for the small cases.
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from massivedatans_b200 import synth  # noqa: E402
from oracle import ref  # noqa: E402

# parameter triples printed (commented out) at sample.py:73-75 -- (A, mu, log_sig)
SAMPLE_PY_TRIPLES = numpy.array([
    [0.88091237, 444.44207558, 2.77671952],
    [1.65758829e-01, 4.45518543e+02, 3.25894638e+00],
    [0.95572931, 443.99407818, 2.95764509],
])


def golden_clike():
    out = {}
    N = 257                       # ragged: not a multiple of any tile size
    x, y, _ = synth.horns(N)
    msk = synth.masks(N)
    pts = numpy.vstack([synth.parameter_points(5),
                        numpy.column_stack([SAMPLE_PY_TRIPLES[:, 0], SAMPLE_PY_TRIPLES[:, 1],
                                            10 ** SAMPLE_PY_TRIPLES[:, 2]])])
    out['N'] = N
    out['params'] = pts
    for name in ('all', 'half', 'sparse', 'prefix'):
        m = msk[name]
        out['mask_' + name] = m
        out['Lout_' + name] = numpy.array(
            [ref.clike(x, y, p[0], p[1], p[2], synth.NOISE_LEVEL, m) for p in pts])
    # the no-signal data set of gennothing.py (BASELINE config 1), full mask, 3 points
    N2 = 1000
    x2, y2 = synth.nothing(N2)
    out['nothing_N'] = N2
    out['nothing_Lout'] = numpy.array(
        [ref.clike(x2, y2, p[0], p[1], p[2], synth.NOISE_LEVEL, numpy.ones(N2, dtype=bool))
         for p in pts[:3]])
    numpy.savez_compressed(os.path.join(HERE, 'clike.npz'), **out)


def golden_cmuselike():
    ndata, nspec = 61, 3600
    y, v, template = synth.muse(ndata=ndata, nspec=nspec)
    rs = numpy.random.RandomState(2)
    mask = rs.uniform(size=ndata) < 0.7
    ypreds = numpy.array([synth.muse_template(nspec, phase=ph) for ph in (0.0, 0.3, 1.1)])
    Lout = numpy.array([ref.cmuselike(y, v, yp, mask) for yp in ypreds])
    Lall = numpy.array([ref.cmuselike(y, v, yp, numpy.ones(ndata, dtype=bool)) for yp in ypreds])
    # small odd-sized case with raw inputs stored
    ys, vs, ts = synth.muse(ndata=7, nspec=37, seed=9)
    ms = numpy.array([1, 0, 1, 1, 0, 0, 1], dtype=bool)
    numpy.savez_compressed(os.path.join(HERE, 'cmuselike.npz'), ndata=ndata, nspec=nspec,
                           mask=mask, phases=numpy.array([0.0, 0.3, 1.1]), Lout=Lout, Lall=Lall,
                           small_y=ys, small_v=vs, small_ypred=ts, small_mask=ms,
                           small_Lout=ref.cmuselike(ys, vs, ts, ms))


def golden_neighbors():
    out = {}
    # the reference's own self-test inputs: clustering/neighbors.py:240-247
    numpy.random.seed(1)
    u = numpy.random.uniform(size=(200, 2))
    vals = []
    for i in range(100):
        numpy.random.seed(i)
        chosen = synth.bootstrap_chosen(200, 10)
        vals.append(ref.bootstrapped_maxdistance_chosen(u, chosen))
    out['selftest_u'] = u
    out['selftest_maxdistance'] = numpy.array(vals)
    for ndim in (2, 3, 5):
        xx, yy = synth.members_and_candidates(400, 1000, ndim, seed=ndim)
        numpy.random.seed(ndim)
        chosen = synth.bootstrap_chosen(400, 10)
        r = ref.bootstrapped_maxdistance_chosen(xx, chosen)
        k = 'd%d_' % ndim
        out[k + 'r'] = r
        out[k + 'chosen'] = chosen
        out[k + 'mdnn'] = ref.most_distant_nearest_neighbor(xx)
        out[k + 'counts'] = ref.count_within_distance_of(xx, r, yy)
        out[k + 'any'] = ref.any_within_distance_of(xx, r, yy)
        out[k + 'counts_cm3'] = ref.count_within_distance_of_raw(xx, r, yy, numpy.zeros(len(yy)), 3)
        out[k + 'within'] = numpy.array([ref.is_within_distance_of(xx, r, yy[j].copy())
                                         for j in range(50)])
    # quirk A3 (cneighbors.c:162): sample 0 never counts as an un-chosen point
    q = numpy.array([[10., 10.], [0., 0.], [0.1, 0.], [0., 0.2], [0.3, 0.3]])
    qc = numpy.array([[0.], [1.], [1.], [0.], [1.]])
    out['quirk_x'] = q
    out['quirk_chosen'] = qc
    out['quirk_r'] = ref.bootstrapped_maxdistance_chosen(q, qc)
    numpy.savez_compressed(os.path.join(HERE, 'neighbors.npz'), **out)


if __name__ == '__main__':
    golden_clike()
    golden_cmuselike()
    golden_neighbors()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')
