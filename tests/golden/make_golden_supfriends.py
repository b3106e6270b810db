#!/usr/bin/env python
"""Generate tests/golden/supfriends.npz: per-axis "SupFriends" distances computed by the
REFERENCE's own clustering/neighbors.py (`initial_maxdistance_guess`, `update_maxdistance`,
`find_maxdistance`, neighbors.py:22-73) on seeded point sets.  Build container only; the module is
imported from a temporary symlink farm (nothing copied), with the reference cneighbors.so beside
it as neighbors.py:97-98 expects.

    python tests/golden/make_golden_supfriends.py
"""
import os
import sys
import tempfile

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

CASES = {
    # tag: (n, ndim, point seed, run seed, nbootstraps, shape of the cloud)
    'a': (200, 2, 1, 11, 15, 'uniform'),
    'b': (1000, 3, 2, 12, 15, 'uniform'),
    'c': (400, 5, 3, 13, 10, 'gauss'),
    'd': (60, 3, 4, 14, 20, 'clusters'),
    'e': (2500, 3, 5, 15, 5, 'clusters'),
}


def points(n, ndim, seed, shape):
    rs = numpy.random.RandomState(seed)
    if shape == 'uniform':
        return rs.uniform(size=(n, ndim))
    if shape == 'gauss':
        return rs.normal(0.5, [0.1 * (k + 1) for k in range(ndim)], size=(n, ndim))
    centres = rs.uniform(size=(4, ndim))
    return centres[rs.randint(0, 4, size=n)] + rs.normal(0, 0.02, size=(n, ndim))


def main():
    os.environ.pop('OMP_NUM_THREADS', None)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, 'clustering'))
        for name in os.listdir(os.path.join(REF, 'clustering')):
            if name.endswith('.py'):
                os.symlink(os.path.join(REF, 'clustering', name), os.path.join(tmp, 'clustering', name))
        os.symlink(os.path.join(ROOT, 'oracle', '_ref', 'cneighbors.so'),
                   os.path.join(tmp, 'clustering', 'cneighbors.so'))
        sys.path.insert(0, tmp)
        import clustering.neighbors as nb
        for tag, (n, ndim, seed, run_seed, nboot, shape) in CASES.items():
            u = points(n, ndim, seed, shape)
            out[tag + '_u'] = u
            out[tag + '_cfg'] = numpy.array([run_seed, nboot])
            out[tag + '_initial'] = nb.initial_maxdistance_guess(u)
            numpy.random.seed(run_seed)
            out[tag + '_maxdistance'] = nb.find_maxdistance(u, nbootstraps=nboot)
            # one round from a deliberately small start: many uncovered points in one round
            numpy.random.seed(run_seed + 100)
            out[tag + '_round_small'] = nb.update_maxdistance(u, 0, out[tag + '_initial'] * 0.25)
            out[tag + '_after'] = numpy.random.uniform()       # the stream position afterwards
            print(tag, n, ndim, 'initial', out[tag + '_initial'], '->', out[tag + '_maxdistance'])
        sys.path.remove(tmp)
    path = os.path.join(HERE, 'supfriends.npz')
    numpy.savez_compressed(path, **out)
    print('wrote', path)


if __name__ == '__main__':
    main()
