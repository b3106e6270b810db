"""Neighbourhood helper functions -- host mirror of the reference's clustering/neighbors.py
with the native part running on the B200 (libmdns_b200.so, bit-exact with cneighbors.c).

Same function names, argument order and return dtypes as the reference module:

  most_distant_nearest_neighbor(xx)            neighbors.py:107-110 -> float
  is_within_distance_of(xx, maxdistance, y)    neighbors.py:121-124 -> bool
  count_within_distance_of(xx, maxdistance, yy) neighbors.py:137-147 -> int array
  any_within_distance_of(xx, maxdistance, yy)  neighbors.py:149-159 -> bool array
  bootstrapped_maxdistance(xx, nbootstraps)    neighbors.py:170-177 -> float
  nearest_rdistance_guess(u, metric)           neighbors.py:185-187
  find_rdistance(u, verbose, nbootstraps, metric) neighbors.py:229-231
  initial_maxdistance_guess / update_maxdistance / find_maxdistance
                                                neighbors.py:22-73 (per-axis SupFriends)

The bootstrap selection matrix is drawn with numpy.random on the host exactly as the
reference does (neighbors.py:172-174), so seeded runs consume the same RNG stream.
Only the euclidean metric is implemented natively (as in the reference); there is no
scipy fallback -- a missing library or GPU raises.
"""
import numpy

from .. import _lib


def _f64(a, ndim):
    a = numpy.ascontiguousarray(a, dtype=numpy.float64)
    if a.ndim != ndim:
        raise ValueError('expected a %d-d float64 array' % ndim)
    return a


def most_distant_nearest_neighbor(xx):
    xx = _f64(xx, 2)
    i, m = xx.shape
    r = _lib.load().mdns_most_distant_nearest_neighbor(xx.ctypes.data, i, m)
    if r != r:
        raise _lib.MdnsError('most_distant_nearest_neighbor: ' + _lib.last_error())
    return r


def is_within_distance_of(xx, maxdistance, y):
    xx = _f64(xx, 2)
    y = _f64(y, 1)
    i, m = xx.shape
    r = _lib.load().mdns_is_within_distance_of(xx.ctypes.data, i, m, maxdistance, y.ctypes.data)
    if r < 0:
        raise _lib.MdnsError('is_within_distance_of: ' + _lib.last_error())
    return r == 1


def _count(xx, maxdistance, yy, countmax):
    xx = _f64(xx, 2)
    yy = _f64(yy, 2)
    i, m = xx.shape
    j = len(yy)
    counts = numpy.zeros(j)
    r = _lib.load().mdns_count_within_distance_of(xx.ctypes.data, i, m, maxdistance,
                                                  yy.ctypes.data, j, counts.ctypes.data, countmax)
    if r != 0:
        raise _lib.MdnsError('count_within_distance_of: ' + _lib.last_error())
    return counts


def count_within_distance_of(xx, maxdistance, yy):
    return _count(xx, maxdistance, yy, 0).astype(int)


def any_within_distance_of(xx, maxdistance, yy):
    return _count(xx, maxdistance, yy, 1) > 0


def bootstrapped_maxdistance(xx, nbootstraps):
    xx = _f64(xx, 2)
    nsamples, ndim = xx.shape
    chosen = numpy.zeros((nsamples, nbootstraps))
    for b in range(nbootstraps):
        chosen[numpy.random.choice(numpy.arange(nsamples), size=nsamples, replace=True), b] = 1.
    r = _lib.load().mdns_bootstrapped_maxdistance(xx.ctypes.data, nsamples, ndim,
                                                  chosen.ctypes.data, nbootstraps)
    if r != r:
        raise _lib.MdnsError('bootstrapped_maxdistance: ' + _lib.last_error())
    return r


def nearest_rdistance_guess(u, metric='euclidean'):
    if metric != 'euclidean':
        raise NotImplementedError('only the euclidean metric runs natively (neighbors.py:186)')
    return most_distant_nearest_neighbor(u)


def find_rdistance(u, verbose=False, nbootstraps=15, metric='euclidean'):
    if metric != 'euclidean':
        raise NotImplementedError('only the euclidean metric runs natively (neighbors.py:230)')
    return bootstrapped_maxdistance(u, nbootstraps)


# --- per-axis "SupFriends" distance (neighbors.py:22-73): host logic, same RNG use ---------

def initial_maxdistance_guess(u):
    """neighbors.py:22-29: per-axis |delta| to each point's nearest neighbour, maximised."""
    u = _f64(u, 2)
    n = len(u)
    d2 = ((u[:, None, :] - u[None, :, :]) ** 2).sum(axis=2)
    numpy.fill_diagonal(d2, numpy.inf)
    nearest = d2.argmin(axis=1)
    return numpy.abs(u[nearest, :] - u[numpy.arange(n), :]).max(axis=0)


def update_maxdistance(u, ibootstrap, maxdistance, verbose=False):
    """neighbors.py:31-62: one bootstrap round of the per-axis box half-widths."""
    n, ndim = u.shape
    choice = list(set(numpy.random.choice(numpy.arange(n), size=n)))
    notchosen = set(range(n)) - set(choice)
    for i in notchosen:
        dists = numpy.abs(u[i, :] - u[choice, :])
        close = numpy.all(dists < maxdistance.reshape((1, -1)), axis=1)
        if not close.any():
            suggest = numpy.where(maxdistance > dists, dists, maxdistance)
            increase = numpy.log(suggest).sum(axis=1) - numpy.log(maxdistance).sum()
            nearest = numpy.argmin(increase)
            if verbose:
                print(ibootstrap, 'nearest:', u[i], u[nearest], increase[nearest])
            maxdistance = numpy.where(dists[nearest] > maxdistance, dists[nearest], maxdistance)
            if verbose:
                print(ibootstrap, 'extending:', maxdistance)
    return maxdistance


def find_maxdistance(u, verbose=False, nbootstraps=15):
    """neighbors.py:64-73."""
    u = _f64(u, 2)
    maxdistance = initial_maxdistance_guess(u)
    for ibootstrap in range(nbootstraps):
        maxdistance = update_maxdistance(u, ibootstrap, maxdistance, verbose=verbose)
    return maxdistance
