"""The sampler's live-point likelihood table resident next to the data.

Host-side mirror of what multi_nested_sampler.py does with ``live_pointsL[nlive, ndata]``:

* ``LiveTable.prepare()``        -- ``prepare`` :134-137 (``Lmins``, ``Lmini``) and ``Lmax`` :531
* ``LiveTable.lmins_higher(...)`` -- the ``Lmins_higher`` loop :438-447 (``find_nsmallest`` :44-47)
* ``LiveTable.replace(...)``     -- the replacement of the dead points :520-524
* ``LiveTable.fill_from_launch`` -- the initial population :91-111, K rows per batched launch
* ``LiveTable.subsets(...)`` / ``generate_subsets(...)`` -- ``generate_subsets_graph`` /
  ``generate_subsets_nograph`` :204-355: groups of data sets that share live points

All results are selections and bit-identical to the numpy expressions they replace.
"""
import ctypes
import weakref

import numpy

from . import _lib
from .likelihood import _addr, _pool


class LiveTable(object):
    """``live_pointsL[nlive, ndata]`` on the devices of `dataset` (a ResidentDataset)."""

    def __init__(self, dataset, nlive):
        lib = _lib.load()
        handle = ctypes.c_void_p()
        _lib.check(lib.mdns_livetable_create(dataset._h, int(nlive), ctypes.byref(handle)),
                   'mdns_livetable_create')
        self._lib = lib
        self._h = handle
        self._dataset = dataset          # keeps the data set (and its devices' context) alive
        self.nlive = int(nlive)
        self.ndata = dataset.ndata
        self._finalizer = weakref.finalize(self, lib.mdns_livetable_destroy, handle)

    def close(self):
        self._finalizer()

    def upload(self, live_pointsL):
        L = numpy.ascontiguousarray(live_pointsL, dtype=numpy.float64)
        if L.shape != (self.nlive, self.ndata):
            raise ValueError('live_pointsL must be [nlive, ndata]')
        _lib.check(self._lib.mdns_livetable_upload(self._h, _addr(L)), 'mdns_livetable_upload')

    def download(self):
        L = numpy.empty((self.nlive, self.ndata))
        _lib.check(self._lib.mdns_livetable_download(self._h, _addr(L)), 'mdns_livetable_download')
        return L

    def fill_from_launch(self, row0):
        """Rows [row0, row0+K) := the K logL vectors of the data set's last all-active launch."""
        _lib.check(self._lib.mdns_livetable_fill_from_launch(self._h, self._dataset._h, int(row0)),
                   'mdns_livetable_fill_from_launch')

    def prepare(self):
        """(Lmins, Lmini, Lmax): min, argmin and max over the live points, per data set."""
        # pinned result vectors: three pageable downloads were most of the call (round 1: 1.24 ms
        # for a 640 MB table whose reduction streams in 0.1-0.2 ms)
        lo = _pool.empty(self.ndata)
        hi = _pool.empty(self.ndata)
        at = _pool.empty(self.ndata).view(numpy.int64)
        _lib.check(self._lib.mdns_livetable_colstats(self._h, _addr(lo), _addr(at), _addr(hi)),
                   'mdns_livetable_colstats')
        return lo, at, hi

    def stage_thresholds(self):
        """The column minima become the accept thresholds of the data set (all data sets active),
        device to device: afterwards ``dataset.draw_batch(params, noise)`` tests candidates
        against ``Lmins = live_pointsL.min(axis=0)`` without the vector ever visiting the host."""
        self._dataset.set_mask(None)
        _lib.check(self._lib.mdns_livetable_stage_thresholds(self._h, self._dataset._h),
                   'mdns_livetable_stage_thresholds')
        self._dataset._draw_n_act = self.ndata
        self._dataset._draw_mask = None

    def replace(self, rows, values):
        """live_pointsL[rows[d], d] = values[d] for every data set d with rows[d] >= 0."""
        rows = numpy.ascontiguousarray(rows, dtype=numpy.int64)
        values = numpy.ascontiguousarray(values, dtype=numpy.float64)
        if rows.shape != (self.ndata,) or values.shape != (self.ndata,):
            raise ValueError('rows and values must have one entry per data set')
        _lib.check(self._lib.mdns_livetable_replace(self._h, _addr(rows), _addr(values)),
                   'mdns_livetable_replace')

    def lmins_higher(self, joint_indices, shelf_L):
        """``Lmins_higher`` of multi_nested_sampler.py:438-447.

        joint_indices : increasing data-set indices; shelf_L : for each of them the likelihoods
        queued on its shelf (sequence of sequences).  Entry j is the element of rank
        ``len(shelf_L[j])`` of ``live_pointsL[:, d]`` joined with ``shelf_L[j]`` -- for an empty
        shelf that is the column minimum, as in the reference."""
        idx = numpy.ascontiguousarray(joint_indices, dtype=numpy.int32)
        if len(shelf_L) != len(idx):
            raise ValueError('one shelf per listed data set')
        off = numpy.zeros(len(idx) + 1, dtype=numpy.int64)
        off[1:] = numpy.cumsum([len(s) for s in shelf_L])
        vals = (numpy.concatenate([numpy.asarray(s, dtype=numpy.float64) for s in shelf_L])
                if off[-1] > 0 else numpy.zeros(1))
        vals = numpy.ascontiguousarray(vals, dtype=numpy.float64)
        out = numpy.empty(len(idx))
        _lib.check(self._lib.mdns_livetable_lmins_higher(self._h, _addr(idx), len(idx), _addr(off),
                                                         _addr(vals), _addr(out)),
                   'mdns_livetable_lmins_higher')
        return out

    # -- live-point indices and the subset partition ---------------------------------------
    def upload_points(self, live_pointsp):
        """``live_pointsp[nlive, ndata]``: indices into the point pile (:111)."""
        P = numpy.ascontiguousarray(live_pointsp, dtype=numpy.int64)
        if P.shape != (self.nlive, self.ndata):
            raise ValueError('live_pointsp must be [nlive, ndata]')
        _lib.check(self._lib.mdns_livetable_upload_points(self._h, _addr(P)),
                   'mdns_livetable_upload_points')

    def replace_points(self, rows, point_ids):
        """live_pointsp[rows[d], d] = point_ids[d] for every data set d with rows[d] >= 0 (:523)."""
        rows = numpy.ascontiguousarray(rows, dtype=numpy.int64)
        ids = numpy.ascontiguousarray(point_ids, dtype=numpy.int64)
        if rows.shape != (self.ndata,) or ids.shape != (self.ndata,):
            raise ValueError('rows and point_ids must have one entry per data set')
        _lib.check(self._lib.mdns_livetable_replace_points(self._h, _addr(rows), _addr(ids)),
                   'mdns_livetable_replace_points')

    def subsets(self, data_mask, npoints):
        """labels[d] = smallest data-set index of the group of d (-1 outside the mask)."""
        m = None
        if data_mask is not None:
            m = numpy.ascontiguousarray(data_mask, dtype=numpy.bool_)
            if m.shape != (self.ndata,):
                raise ValueError('data_mask must have ndata entries')
        labels = numpy.empty(self.ndata, dtype=numpy.int32)
        ncomp = ctypes.c_int()
        rounds = ctypes.c_int()
        _lib.check(self._lib.mdns_livetable_subsets(self._h, _addr(m), int(npoints), _addr(labels),
                                                    ctypes.byref(ncomp), ctypes.byref(rounds)),
                   'mdns_livetable_subsets')
        self.last_rounds = rounds.value
        return labels


def generate_subsets(data_mask, live_pointsp, labels):
    """Yield ``(member_data_mask, member_live_pointsp)`` per group like
    ``generate_subsets_nograph`` (multi_nested_sampler.py:237-260), in the same group order
    (increasing first member).  The live points of a group come sorted (``numpy.unique``); the
    reference lists the same set in discovery order."""
    ndata = len(data_mask)
    order = numpy.argsort(labels, kind='stable')
    order = order[labels[order] >= 0]
    bounds = numpy.nonzero(numpy.diff(labels[order]))[0] + 1
    for members in numpy.split(order, bounds):
        if len(members) == 0:
            continue
        member_data_mask = numpy.zeros(ndata, dtype=bool)
        member_data_mask[members] = True
        yield member_data_mask, numpy.unique(live_pointsp[:, members])
