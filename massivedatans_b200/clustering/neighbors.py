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


# --- per-axis "SupFriends" distance (neighbors.py:22-73) ------------------------------------
# The O(n^2) parts run on the device over members uploaded once per call: the nearest other
# member of every member (neighbors.py:24-25) and, per bootstrap round, which un-chosen members
# already have a chosen member inside the per-axis box (neighbors.py:40-43).  The box only ever
# grows within a round, so a member the device reports covered stays covered; the host walks
# the uncovered ones in the reference's order, re-checks them against the grown box and applies
# the reference's growth step (neighbors.py:44-58).  numpy.random is consumed exactly as in the
# reference (one `choice` per round).

def _resident(u):
    from .radfriendsregion import ResidentMembers
    return ResidentMembers(u)


def initial_maxdistance_guess(u, _members=None):
    """neighbors.py:22-29: per axis, the largest |delta| between a point and its nearest
    neighbour."""
    u = _f64(u, 2)
    members = _members if _members is not None else _resident(u)
    nearest = members.nearest_index()
    return numpy.abs(u[nearest, :] - u).max(axis=0)


def update_maxdistance(u, ibootstrap, maxdistance, verbose=False, _members=None):
    """neighbors.py:31-62: one bootstrap round -- every point left out of the resample must have
    a resampled point inside the box; where none has, the box grows towards the point whose
    clipped offsets have the smallest log-volume (the reference's choice, neighbors.py:47-55)."""
    u = _f64(u, 2)
    n = len(u)
    members = _members if _members is not None else _resident(u)
    choice = list(set(numpy.random.choice(numpy.arange(n), size=n)))
    left_out = list(set(range(n)) - set(choice))       # the reference's iteration order
    covered = members.axis_covered(maxdistance, left_out, choice)
    for i, done in zip(left_out, covered):
        if done:
            continue
        offsets = numpy.abs(u[i, :] - u[choice, :])
        if numpy.all(offsets < maxdistance.reshape((1, -1)), axis=1).any():
            continue                                   # an earlier growth of this round got it
        clipped = numpy.where(maxdistance > offsets, offsets, maxdistance)
        cost = numpy.log(clipped).sum(axis=1) - numpy.log(maxdistance).sum()
        towards = numpy.argmin(cost)
        if verbose:
            print(ibootstrap, 'nearest:', u[i], u[towards], cost[towards])
        maxdistance = numpy.where(offsets[towards] > maxdistance, offsets[towards], maxdistance)
        if verbose:
            print(ibootstrap, 'extending:', maxdistance)
    return maxdistance


def find_maxdistance(u, verbose=False, nbootstraps=15):
    """neighbors.py:64-73."""
    u = _f64(u, 2)
    members = _resident(u)
    maxdistance = initial_maxdistance_guess(u, _members=members)
    for ibootstrap in range(nbootstraps):
        maxdistance = update_maxdistance(u, ibootstrap, maxdistance, verbose=verbose,
                                         _members=members)
    return maxdistance
