"""RadFriends region on resident members -- host mirror of the reference's
clustering/radfriendsregion.py:58-182 (class ``RadFriendsRegion``).

Same constructor, attributes (``members``, ``maxdistance``, ``lo``, ``hi``) and methods
(``are_inside``, ``is_inside``, ``count_nearby_members``, ``add_members``, ``generate``), and the
same use of ``numpy.random`` -- the draws are made on the host in the reference's order and
shapes, so a seeded run proposes exactly the same candidates (the neighbour decisions behind
them are bit-exact with cneighbors.c).  What changes: the members are uploaded to the device
once per region (``mdns_region_set_members``) instead of once per neighbour call.

SURVEY.md section 8(f) rank 3 asks for candidate generation fused with the neighbour test on
the device; that changes the RNG stream (statistical parity only), so ``generate`` keeps the
host RNG and moves only the member set, and the fused device form is the separate
``generate_device`` (counter-based Philox stream, ``mdns_region_generate``).
"""
import ctypes
import weakref

import numpy

from .. import _lib


class ResidentMembers(object):
    """Member points of one region resident on a device (``mdns_region_*``)."""

    def __init__(self, members, device=0):
        lib = _lib.load()
        _lib.require_device()
        handle = ctypes.c_void_p()
        _lib.check(lib.mdns_region_create(int(device), ctypes.byref(handle)), 'mdns_region_create')
        self._lib = lib
        self._h = handle
        self._finalizer = weakref.finalize(self, lib.mdns_region_destroy, handle)
        self.set(members)

    def set(self, members):
        xx = numpy.ascontiguousarray(members, dtype=numpy.float64)
        if xx.ndim != 2:
            raise ValueError('members must be [n, ndim]')
        self.n, self.ndim = xx.shape
        _lib.check(self._lib.mdns_region_set_members(self._h, xx.ctypes.data, self.n, self.ndim),
                   'mdns_region_set_members')

    def counts(self, maxdistance, us, countmax):
        yy = numpy.ascontiguousarray(us, dtype=numpy.float64)
        if yy.ndim != 2 or yy.shape[1] != self.ndim:
            raise ValueError('candidates must be [m, ndim]')
        out = numpy.zeros(len(yy))
        if len(yy):
            _lib.check(self._lib.mdns_region_count_within(self._h, maxdistance, yy.ctypes.data,
                                                          len(yy), out.ctypes.data, countmax),
                       'mdns_region_count_within')
        return out

    def is_within(self, maxdistance, u):
        y = numpy.ascontiguousarray(u, dtype=numpy.float64)
        res = ctypes.c_int()
        _lib.check(self._lib.mdns_region_is_within(self._h, maxdistance, y.ctypes.data,
                                                   ctypes.byref(res)), 'mdns_region_is_within')
        return res.value == 1

    def nearest_index(self):
        """Index of every member's nearest other member (clustering/neighbors.py:24-25)."""
        out = numpy.empty(self.n, dtype=numpy.int32)
        _lib.check(self._lib.mdns_region_nearest_index(self._h, out.ctypes.data),
                   'mdns_region_nearest_index')
        return out

    def axis_covered(self, maxdistance, query, ref):
        """For every listed member: is some member of `ref` inside the per-axis box
        `maxdistance` around it (clustering/neighbors.py:40-43)."""
        md = numpy.ascontiguousarray(maxdistance, dtype=numpy.float64)
        if md.shape != (self.ndim,):
            raise ValueError('maxdistance must have one entry per axis')
        q = numpy.ascontiguousarray(query, dtype=numpy.int32)
        r = numpy.ascontiguousarray(ref, dtype=numpy.int32)
        out = numpy.zeros(len(q), dtype=numpy.uint8)
        _lib.check(self._lib.mdns_region_axis_covered(self._h, md.ctypes.data, q.ctypes.data, len(q),
                                                      r.ctypes.data, len(r), out.ctypes.data),
                   'mdns_region_axis_covered')
        return out != 0

    def generate(self, maxdistance, nproposals, seed, first_proposal=0):
        """``nproposals`` ball draws fused with the neighbour count on the device; returns the
        accepted points [k, ndim] (uniform in the union of balls), in proposal order."""
        out = numpy.empty((int(nproposals), self.ndim))
        n = ctypes.c_int(0)
        _lib.check(self._lib.mdns_region_generate(self._h, maxdistance, int(seed), int(first_proposal),
                                                  int(nproposals), out.ctypes.data, int(nproposals),
                                                  ctypes.byref(n)), 'mdns_region_generate')
        return out[:n.value]

    def bootstrapped_maxdistance(self, nbootstraps):
        # selection matrix drawn on the host exactly as clustering/neighbors.py:172-174
        chosen = numpy.zeros((self.n, nbootstraps))
        for b in range(nbootstraps):
            chosen[numpy.random.choice(numpy.arange(self.n), size=self.n, replace=True), b] = 1.
        r = ctypes.c_double()
        _lib.check(self._lib.mdns_region_bootstrapped_maxdistance(self._h, chosen.ctypes.data,
                                                                  nbootstraps, ctypes.byref(r)),
                   'mdns_region_bootstrapped_maxdistance')
        return r.value


class RadFriendsRegion(object):
    """Union of balls of radius ``maxdistance`` around ``members`` (radfriendsregion.py:58-70)."""

    PROPOSALS = 1000          # points per proposal round (radfriendsregion.py:124)
    members_class = ResidentMembers   # who answers the neighbour queries (tests: the CPU oracle)

    def __init__(self, members, maxdistance=None, metric='euclidean', nbootstraps=10,
                 verbose=False, device=0):
        if metric != 'euclidean':
            raise NotImplementedError('only the euclidean metric runs natively '
                                      '(radfriendsregion.py:61)')
        self.members = members
        self._resident = self.members_class(members, device=device)
        if maxdistance is None:
            maxdistance = self._resident.bootstrapped_maxdistance(nbootstraps)
        self.maxdistance = maxdistance
        self.metric = metric
        self.verbose = verbose
        self._update_box()

    def _update_box(self):
        self.lo = numpy.min(self.members, axis=0) - self.maxdistance
        self.hi = numpy.max(self.members, axis=0) + self.maxdistance

    def add_members(self, us):
        self.members = numpy.vstack((self.members, us))
        self._resident.set(self.members)
        self._update_box()

    def count_nearby_members(self, us):
        return self._resident.counts(self.maxdistance, us, 0).astype(int)

    def are_inside(self, us):
        return self._resident.counts(self.maxdistance, us, 1) > 0

    def is_inside(self, u):
        u = numpy.asarray(u)
        if not ((u >= self.lo).all() and (u <= self.hi).all()):
            return False
        return self._resident.is_within(self.maxdistance, u)

    # -- proposals ---------------------------------------------------------------------------
    def _box_round(self, n, ndim):
        us = numpy.random.uniform(self.lo, self.hi, size=(n, ndim))
        return us[self.are_inside(us), :]

    def _ball_round(self, n, ndim, members=None, maxdistance=None):
        # centres and radius from what the generator captured when it was created
        # (radfriendsregion.py:118-119); the neighbour count below uses the live member set (:167)
        members = self.members if members is None else members
        maxdistance = self.maxdistance if maxdistance is None else maxdistance
        centres = members[numpy.random.randint(0, len(members), n), :]
        direction = numpy.random.normal(0, 1, size=(n, ndim))
        direction = direction / ((direction ** 2).sum(axis=1) ** 0.5).reshape((-1, 1))
        radius = maxdistance * numpy.random.uniform(0, 1, size=(n, 1)) ** (1. / ndim)
        us = centres + direction * radius
        nnear = self.count_nearby_members(us)
        coin = numpy.random.uniform(size=len(us))
        with numpy.errstate(divide='ignore'):
            keep = coin < 1. / nnear
        return us[keep, :]

    def generate_device(self, nproposals, seed, first_proposal=0):
        """Device-side candidate generation: ball draws, neighbour count
        and 1/count thinning in one kernel; the random numbers are Philox(seed, proposal index),
        NOT numpy's, so this is a statistically equivalent alternative to ``generate`` -- use the
        latter where a seeded run must reproduce the reference draw by draw."""
        return self._resident.generate(self.maxdistance, nproposals, seed, first_proposal)

    def generate(self, nmax=0):
        """Yield ``(accepted points [k, ndim], proposals spent since the last yield)`` like
        radfriendsregion.py:117-182: alternating rounds of box draws (kept if inside the region)
        and ball draws (thinned by 1/number of members nearby)."""
        n = self.PROPOSALS
        members, maxdistance = self.members, self.maxdistance     # captured once, as the reference does
        ndim = numpy.shape(members)[1]
        spent_total = 0
        spent = 0
        while nmax == 0 or spent_total < nmax:
            spent += n
            spent_total += n
            us = self._box_round(n, ndim)
            if len(us):
                yield us, spent
                spent = 0
            spent += n
            spent_total += n
            us = self._ball_round(n, ndim, members, maxdistance)
            if len(us):
                yield us, spent
                spent = 0
