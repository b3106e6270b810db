#!/usr/bin/env python
"""Generate tests/golden/region.npz from the REFERENCE's own clustering/radfriendsregion.py on
the reference's cneighbors.so (build container only; the reference modules are imported from a
temporary symlink farm, nothing is copied).  Fixed seeds; the GPU mirror must reproduce every
array bit for bit."""
import os
import sys
import tempfile

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)


def main():
    from massivedatans_b200 import synth
    os.environ.pop('OMP_NUM_THREADS', None)
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, 'clustering'))
        for name in os.listdir(os.path.join(REF, 'clustering')):
            if name.endswith('.py'):
                os.symlink(os.path.join(REF, 'clustering', name), os.path.join(tmp, 'clustering', name))
        os.symlink(os.path.join(ROOT, 'oracle', '_ref', 'cneighbors.so'),
                   os.path.join(tmp, 'clustering', 'cneighbors.so'))
        sys.path.insert(0, tmp)
        from clustering.radfriendsregion import RadFriendsRegion
        import clustering.neighbors as nb
        assert nb.bootstrapped_maxdistance is not None
        out = {}
        for tag, n, ndim, seed in (('a', 400, 3, 11), ('b', 150, 5, 12)):
            members, cand = synth.members_and_candidates(n, 500, ndim, seed=seed)
            numpy.random.seed(seed)
            region = RadFriendsRegion(members=members, nbootstraps=10)
            out[tag + '_members'] = members
            out[tag + '_seed'] = seed
            out[tag + '_maxdistance'] = region.maxdistance
            out[tag + '_lo'] = region.lo
            out[tag + '_hi'] = region.hi
            out[tag + '_cand'] = cand
            out[tag + '_inside'] = region.are_inside(cand)
            out[tag + '_nnear'] = region.count_nearby_members(cand)
            out[tag + '_is_inside'] = numpy.array([region.is_inside(c) for c in cand[:40]])
            ys = []
            for i, (us, ntotal) in enumerate(region.generate(nmax=8000)):
                out['%s_gen%d' % (tag, i)] = us
                ys.append(ntotal)
            out[tag + '_ntotal'] = numpy.array(ys)
            extra = cand[:25] * 0.5
            region.add_members(extra)
            out[tag + '_lo2'] = region.lo
            out[tag + '_inside2'] = region.are_inside(cand)
        sys.path.remove(tmp)
    path = os.path.join(HERE, 'region.npz')
    numpy.savez_compressed(path, **out)
    print('wrote', path, 'yields:', len(out['a_ntotal']), len(out['b_ntotal']),
          'r =', out['a_maxdistance'], out['b_maxdistance'])


if __name__ == '__main__':
    main()
