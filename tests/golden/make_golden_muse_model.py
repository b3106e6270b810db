#!/usr/bin/env python
"""Generate tests/golden/muse_model.npz: model spectra computed by the REFERENCE's own
`model()` (musefuse.py:222-284), its own `ages` table (:191) and its own Calzetti block
(:205-217).

musefuse.py is a script that reads a FITS cube, a region file and BC03 template files at import
(all absent), so it cannot be imported.  Instead the script is parsed and exactly those
statements are executed -- unmodified, compiled from the reference's own source text, nothing
copied into this repository -- in a namespace whose inputs (template grids, wavelength grids)
are the seeded synthetic ones of massivedatans_b200/synth.py (`muse_grids`, `muse_wavelength`).
Build container only; the fixture travels.  It pins oracle.np.muse_model / calzetti, which in
turn check the device model (tests/test_muse_model.py).

    python tests/golden/make_golden_muse_model.py
"""
import ast
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/musefuse.py'
sys.path.insert(0, ROOT)

NSPEC = 900
NWAVE = 2400
NPOINTS = 24


def reference_pieces():
    """The statements of musefuse.py that define `ages`, `Zs`, the wavelength unit change, the
    Calzetti curve and `model`, in file order."""
    src = open(REF).read()
    tree = ast.parse(src, REF)
    wanted = []
    for node in tree.body:
        seg = ast.get_source_segment(src, node) or ''
        if isinstance(node, ast.FunctionDef) and node.name == 'model':
            wanted.append(node)
        elif isinstance(node, ast.Assign):
            names = [t.id for t in node.targets if isinstance(t, ast.Name)]
            subs = [t.value.id for t in node.targets
                    if isinstance(t, ast.Subscript) and isinstance(t.value, ast.Name)]
            if names and names[0] in ('ages', 'Zs', 'calzetti_result'):
                wanted.append(node)
            elif names and names[0] == 'mask' and 'model_wavelength' in seg:
                wanted.append(node)
            elif names and names[0] in ('wavelength', 'model_wavelength') and '/ 10.' in seg:
                wanted.append(node)
            elif subs and subs[0] == 'calzetti_result':
                wanted.append(node)
    return ast.Module(body=wanted, type_ignores=[]), [ast.get_source_segment(src, n)[:60] for n in wanted]


def main():
    from massivedatans_b200 import synth
    Zs, ages_syn, model_wl_A, grids = synth.muse_grids(nwave=NWAVE)
    module, heads = reference_pieces()
    for h in heads:
        print('  exec:', h.replace('\n', ' '))
    ns = {'numpy': numpy, 'grid': list(grids), 'wavelength': synth.muse_wavelength(NSPEC),
          'model_wavelength': model_wl_A.copy(), 'nspec': NSPEC}
    exec(compile(module, REF, 'exec'), ns)
    assert len(ns['ages']) == grids.shape[1], (len(ns['ages']), grids.shape)
    assert numpy.array_equal(ns['Zs'], Zs)
    pts = synth.muse_parameter_points(NPOINTS)
    pts[0, 3] = 0.0                      # z = 0: data grid inside the template grid
    pts[1, 3] = 4.0                      # far blue of the template grid: clamped interpolation
    pts[2, 0] = Zs[3]                    # exactly on a metallicity node
    pts[3, 2] = 1e-4                     # star formation just started: one age bin
    spectra = numpy.array([ns['model'](Z, 10 ** logtau, age, z, ebv) for Z, logtau, age, z, ebv in pts])
    assert spectra.shape == (NPOINTS, NSPEC)
    out = os.path.join(HERE, 'muse_model.npz')
    numpy.savez_compressed(out, params=pts, spectra=spectra, ages=ns['ages'],
                           calzetti=ns['calzetti_result'], nspec=NSPEC, nwave=NWAVE)
    print('wrote', out, 'spectra range', spectra.min(), spectra.max(),
          'all-zero spectra:', int((~spectra.any(axis=1)).sum()))


if __name__ == '__main__':
    main()
