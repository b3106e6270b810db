"""ctypes bindings of oracle/mdns_oracle.c (liboracle.so) -- test infrastructure."""
import os
import subprocess
from ctypes import c_double, c_int, c_longlong as ctypes_longlong, cdll

import numpy
from numpy.ctypeslib import ndpointer

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'liboracle.so')

_f1 = ndpointer(dtype=numpy.float64, ndim=1, flags='C_CONTIGUOUS')
_f2 = ndpointer(dtype=numpy.float64, ndim=2, flags='C_CONTIGUOUS')
_b1 = ndpointer(dtype=numpy.bool_, ndim=1, flags='C_CONTIGUOUS')
_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = cdll.LoadLibrary(LIB_PATH)
        L.oracle_clike.argtypes = [_f1, _f2, c_int, c_int, c_double, c_double, c_double,
                                   c_double, _b1, _f1]
        L.oracle_clike_spectrum.argtypes = [_f1, _f2, c_int, c_int, c_double, _b1, _f1]
        L.oracle_cmuselike.argtypes = [_f2, _f2, _f1, _b1, c_int, c_int, _f1]
        L.oracle_most_distant_nearest_neighbor.argtypes = [_f2, c_int, c_int]
        L.oracle_most_distant_nearest_neighbor.restype = c_double
        L.oracle_is_within_distance_of.argtypes = [_f2, c_int, c_int, c_double, _f1]
        L.oracle_count_within_distance_of.argtypes = [_f2, c_int, c_int, c_double, _f2,
                                                      c_int, _f1, c_int]
        L.oracle_bootstrapped_maxdistance.argtypes = [_f2, c_int, c_int, _f2, c_int]
        L.oracle_bootstrapped_maxdistance.restype = c_double
        L.oracle_live_colstats.argtypes = [_f2, c_int, c_int, _f1,
                                           ndpointer(dtype=numpy.int64, ndim=1, flags='C_CONTIGUOUS'),
                                           _f1]
        L.oracle_live_colstats.restype = None
        L.oracle_find_nsmallest.argtypes = [c_int, _f1, c_int, _f1, c_int]
        L.oracle_find_nsmallest.restype = c_double
        L.oracle_subsets_labels.argtypes = [
            ndpointer(dtype=numpy.int64, ndim=2, flags='C_CONTIGUOUS'), c_int, c_int, _b1,
            ctypes_longlong, ndpointer(dtype=numpy.int32, ndim=1, flags='C_CONTIGUOUS')]
        L.oracle_subsets_labels.restype = None
        _lib = L
    return _lib


def clike(x, y, A, mu, sig, noise, data_mask, Lout=None):
    nx, ndata = y.shape
    if Lout is None:
        Lout = numpy.zeros(int(data_mask.sum()))
    lib().oracle_clike(x, y, ndata, nx, A, mu, sig, noise, data_mask, Lout)
    return Lout


def clike_spectrum(ypred, y, noise, data_mask, Lout=None):
    nx, ndata = y.shape
    if Lout is None:
        Lout = numpy.zeros(int(data_mask.sum()))
    lib().oracle_clike_spectrum(ypred, y, ndata, nx, noise, data_mask, Lout)
    return Lout


def cmuselike(y, v, ypred, data_mask, Lout=None):
    nx, ndata = y.shape
    if Lout is None:
        Lout = numpy.zeros(ndata)
    lib().oracle_cmuselike(y, v, ypred, data_mask, ndata, nx, Lout)
    return Lout


def most_distant_nearest_neighbor(xx):
    n, d = xx.shape
    return lib().oracle_most_distant_nearest_neighbor(xx, n, d)


def is_within_distance_of(xx, maxdistance, y):
    n, d = xx.shape
    return lib().oracle_is_within_distance_of(xx, n, d, maxdistance, y) == 1


def count_within_distance_of_raw(xx, maxdistance, yy, counts, countmax):
    n, d = xx.shape
    lib().oracle_count_within_distance_of(xx, n, d, maxdistance, yy, len(yy), counts, countmax)
    return counts


def count_within_distance_of(xx, maxdistance, yy):
    return count_within_distance_of_raw(xx, maxdistance, yy, numpy.zeros(len(yy)), 0).astype(int)


def any_within_distance_of(xx, maxdistance, yy):
    return count_within_distance_of_raw(xx, maxdistance, yy, numpy.zeros(len(yy)), 1) > 0


def bootstrapped_maxdistance_chosen(xx, chosen):
    n, d = xx.shape
    return lib().oracle_bootstrapped_maxdistance(xx, n, d, chosen, chosen.shape[1])


def live_colstats(live_pointsL):
    """(Lmins, Lmini, Lmax) of live_pointsL[nlive, ndata] -- multi_nested_sampler.py:134-137, :531."""
    L = numpy.ascontiguousarray(live_pointsL, dtype=numpy.float64)
    nlive, ndata = L.shape
    lo, hi = numpy.empty(ndata), numpy.empty(ndata)
    at = numpy.empty(ndata, dtype=numpy.int64)
    lib().oracle_live_colstats(L, nlive, ndata, lo, at, hi)
    return lo, at, hi


def find_nsmallest(n, arr1, arr2):
    """multi_nested_sampler.py:38-47."""
    a1 = numpy.ascontiguousarray(arr1, dtype=numpy.float64)
    a2 = numpy.ascontiguousarray(arr2, dtype=numpy.float64)
    return lib().oracle_find_nsmallest(n, a1, len(a1), a2, len(a2))


def subsets_labels(live_pointsp, data_mask, npoints):
    """labels[d] = smallest data-set index of d's group -- multi_nested_sampler.py:204-355."""
    P = numpy.ascontiguousarray(live_pointsp, dtype=numpy.int64)
    nlive, ndata = P.shape
    m = numpy.ascontiguousarray(data_mask, dtype=numpy.bool_)
    labels = numpy.empty(ndata, dtype=numpy.int32)
    lib().oracle_subsets_labels(P, nlive, ndata, m, int(npoints), labels)
    return labels
