"""ctypes bindings of the UNMODIFIED reference C built into oracle/_ref/ (test infrastructure).

The argtypes are the reference's own: sample.py:85-96 (clike), musefuse.py:509-517
(cmuselike), clustering/neighbors.py:100-167 (cneighbors).  Python-level helpers
mirror the reference wrappers (sample.py:101-108, neighbors.py:107-177).
"""
import os
from ctypes import c_double, c_int, cdll

import numpy
from numpy.ctypeslib import ndpointer

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')

_f1 = ndpointer(dtype=numpy.float64, ndim=1, flags='C_CONTIGUOUS')
_f2 = ndpointer(dtype=numpy.float64, ndim=2, flags='C_CONTIGUOUS')
_b1 = ndpointer(dtype=numpy.bool_, ndim=1, flags='C_CONTIGUOUS')

_cache = {}


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n + '.so'))
               for n in ('clike', 'cmuselike', 'cneighbors'))


def _load(name):
    if name not in _cache:
        path = os.path.join(REF_DIR, name + '.so')
        if not os.path.exists(path):
            raise RuntimeError('reference library %s missing: run `make -C oracle` where '
                               '/root/reference exists' % path)
        lib = cdll.LoadLibrary(path)
        base = name.replace('-parallel', '')
        if base == 'clike':
            lib.like.argtypes = [_f1, _f2, c_int, c_int, c_double, c_double, c_double,
                                 c_double, _b1, _f1]
        elif base == 'cmuselike':
            lib.like.argtypes = [_f2, _f2, _f1, _b1, c_int, c_int, _f1]
        else:
            lib.most_distant_nearest_neighbor.argtypes = [_f2, c_int, c_int]
            lib.most_distant_nearest_neighbor.restype = c_double
            lib.is_within_distance_of.argtypes = [_f2, c_int, c_int, c_double, _f1]
            lib.is_within_distance_of.restype = c_int
            lib.count_within_distance_of.argtypes = [_f2, c_int, c_int, c_double, _f2,
                                                     c_int, _f1, c_int]
            lib.bootstrapped_maxdistance.argtypes = [_f2, c_int, c_int, _f2, c_int]
            lib.bootstrapped_maxdistance.restype = c_double
        _cache[name] = lib
    return _cache[name]


def clike(x, y, A, mu, sig, noise, data_mask, Lout=None):
    """Raw clike.c `like`: returns the accumulated Lout (no -0.5)."""
    nx, ndata = y.shape
    if Lout is None:
        Lout = numpy.zeros(int(data_mask.sum()))
    _load('clike').like(x, y, ndata, nx, A, mu, sig, noise, data_mask, Lout)
    return Lout


def cmuselike(y, v, ypred, data_mask, Lout=None, parallel=False):
    """Raw cmuselike.c `like`: writes -0.5*chi2 into the masked entries of Lout[ndata]."""
    nx, ndata = y.shape
    if Lout is None:
        Lout = numpy.zeros(ndata)
    _load('cmuselike-parallel' if parallel else 'cmuselike').like(
        y, v, ypred, data_mask, ndata, nx, Lout)
    return Lout


def _nb(parallel):
    return _load('cneighbors-parallel' if parallel else 'cneighbors')


def most_distant_nearest_neighbor(xx, parallel=False):
    n, d = xx.shape
    return _nb(parallel).most_distant_nearest_neighbor(xx, n, d)


def is_within_distance_of(xx, maxdistance, y):
    n, d = xx.shape
    return _nb(False).is_within_distance_of(xx, n, d, maxdistance, y) == 1


def count_within_distance_of_raw(xx, maxdistance, yy, counts, countmax):
    n, d = xx.shape
    _nb(False).count_within_distance_of(xx, n, d, maxdistance, yy, len(yy), counts, countmax)
    return counts


def count_within_distance_of(xx, maxdistance, yy):
    return count_within_distance_of_raw(xx, maxdistance, yy, numpy.zeros(len(yy)), 0).astype(int)


def any_within_distance_of(xx, maxdistance, yy):
    return count_within_distance_of_raw(xx, maxdistance, yy, numpy.zeros(len(yy)), 1) > 0


def bootstrapped_maxdistance_chosen(xx, chosen, parallel=False):
    n, d = xx.shape
    return _nb(parallel).bootstrapped_maxdistance(xx, n, d, chosen, chosen.shape[1])
