"""Host-side mirror of the reference's likelihood callables on top of the resident-data shim.

Reference surface kept (same names, argument meaning and return values):

* ``multi_loglikelihood(params, data_mask) -> L[n_act]``  -- sample.py:101-108
  (``params = (A, mu, log_sig)`` after ``priortransform``; ``sig = 10**log_sig``;
  result is ``-0.5 * Lout`` compacted over the masked-in data sets).
* ``multi_loglikelihood_clike(params, data_mask)``         -- musefuse.py:520-535
  (``model(*params)`` on the host, all-zero guard ``-1e100``, result
  ``Lout[data_mask] + normal(0, 1e-5)`` with the jitter drawn from numpy.random
  on the host so that seeded runs consume the same RNG stream).

New, additive: ``ResidentDataset.loglike_batch`` scores K candidates in one pass.
"""
import ctypes
import math
import os
import weakref

import numpy

from . import _lib


def _addr(a):
    # (a.ctypes.data builds a ctypes helper object on every call: 1.4 us against 0.3 us)
    return a.__array_interface__['data'][0] if a is not None else None


class _PinnedPool(object):
    """Result vectors live in pinned host memory so the D2H copy needs no bounce
    buffer; blocks are recycled once the numpy array that wraps them is collected."""

    def __init__(self, max_cached_bytes=1 << 30):
        self.free = {}
        self.cached = 0
        self.max_cached = max_cached_bytes

    def empty(self, n):
        lib = _lib.load()
        nbytes = max(int(n) * 8, 8)
        cap = 1 << (nbytes - 1).bit_length()
        lst = self.free.get(cap)
        if lst:
            ptr = lst.pop()
            self.cached -= cap
        else:
            ptr = lib.mdns_host_alloc(cap)
            if not ptr:
                raise _lib.MdnsError('pinned allocation failed: ' + _lib.last_error())
        buf = (ctypes.c_double * (cap // 8)).from_address(ptr)
        arr = numpy.frombuffer(buf, dtype=numpy.float64, count=int(n))
        weakref.finalize(buf, self._release, ptr, cap)
        return arr

    def _release(self, ptr, cap):
        if self.cached + cap <= self.max_cached:
            self.free.setdefault(cap, []).append(ptr)
            self.cached += cap
        else:
            _lib.load().mdns_host_free(ptr)


_pool = _PinnedPool()


class ResidentDataset(object):
    """Data (and optionally per-element variances) resident in HBM, sharded over `devices`.

    x : [nx] wavelength grid (sample.py:30) or None
    y : [nx, ndata] C-contiguous float64, data-set index fastest (sample.py:31)
    variance : [nx, ndata] or None (musefuse.py:205-206 `noise_level`)
    """

    def __init__(self, x, y, variance=None, devices=None):
        lib = _lib.load()
        _lib.require_device()
        y = numpy.ascontiguousarray(y, dtype=numpy.float64)
        if y.ndim != 2:
            raise ValueError('y must be [nx, ndata]')
        self.nx, self.ndata = y.shape
        if x is not None:
            x = numpy.ascontiguousarray(x, dtype=numpy.float64)
            if x.shape != (self.nx,):
                raise ValueError('x must have nx entries')
        if variance is not None:
            variance = numpy.ascontiguousarray(variance, dtype=numpy.float64)
            if variance.shape != y.shape:
                raise ValueError('variance must match y')
        devs = None
        ndev = 0
        if devices is not None:
            devs = numpy.ascontiguousarray(devices, dtype=numpy.int32)
            ndev = len(devs)
        handle = ctypes.c_void_p()
        _lib.check(lib.mdns_dataset_create(_addr(x), _addr(y), _addr(variance), self.ndata,
                                           self.nx, _addr(devs), ndev, ctypes.byref(handle)),
                   'mdns_dataset_create')
        self._lib = lib
        self._h = handle
        self.has_variance = variance is not None
        self._draw_n_act = None         # active data sets of the draw in progress (begin_draw)
        self._draw_mask = None          # and a copy of its mask (None = all)
        self._comm = False
        self._finalizer = weakref.finalize(self, lib.mdns_dataset_destroy, handle)

    @classmethod
    def from_npy(cls, x, y_path, variance_path=None, devices=None):
        """Data set made resident straight from .npy files holding the [nx, ndata] float64
        matrices (numpy.save): the matrix never exists in host memory -- column blocks stream
        through a 64 MB pinned buffer (sample.py:27-31 reads all of `y` into RAM instead)."""
        lib = _lib.load()
        _lib.require_device()
        hdr = numpy.load(y_path, mmap_mode='r')
        if hdr.ndim != 2 or hdr.dtype != numpy.float64:
            raise ValueError('%s must hold a 2-d float64 array [nx, ndata]' % y_path)
        nx, ndata = hdr.shape
        del hdr
        if x is not None:
            x = numpy.ascontiguousarray(x, dtype=numpy.float64)
            if x.shape != (nx,):
                raise ValueError('x must have nx entries')
        devs, ndev = None, 0
        if devices is not None:
            devs = numpy.ascontiguousarray(devices, dtype=numpy.int32)
            ndev = len(devs)
        handle = ctypes.c_void_p()
        _lib.check(lib.mdns_dataset_create_from_npy(
            _addr(x), os.fsencode(y_path), os.fsencode(variance_path) if variance_path else None,
            _addr(devs), ndev, ctypes.byref(handle)), 'mdns_dataset_create_from_npy')
        self = cls.__new__(cls)
        self.nx, self.ndata = int(nx), int(ndata)
        self._lib = lib
        self._h = handle
        self.has_variance = variance_path is not None
        self._draw_n_act = None
        self._draw_mask = None
        self._comm = False
        self._finalizer = weakref.finalize(self, lib.mdns_dataset_destroy, handle)
        return self

    def close(self):
        self._finalizer()

    # -- one process per GPU -------------------------------------------------
    def comm_unique_id(self):
        """128 bytes identifying a new communicator (rank 0 creates, the others receive it:
        sharding.exchange_unique_id)."""
        buf = ctypes.create_string_buffer(128)
        _lib.check(self._lib.mdns_comm_unique_id(buf), 'mdns_comm_unique_id')
        return buf.raw

    def init_comm(self, unique_id, nranks, rank):
        """Attach the communicator of the exchange step (collective: every rank calls it).  From
        then on draw_batch / draw_batch_sparse / draw_pass / first_accepted return the GLOBAL
        decision: accept counts summed over the ranks by ncclAllReduce on the device."""
        uid = ctypes.create_string_buffer(bytes(unique_id), 128)
        _lib.check(self._lib.mdns_comm_init(self._h, uid, int(nranks), int(rank)), 'mdns_comm_init')
        self._comm = True

    def comm_allreduce(self, values, op='sum'):
        """Sum / max of a few host doubles over the ranks (doubles as a barrier)."""
        v = numpy.ascontiguousarray(values, dtype=numpy.float64).copy()
        _lib.check(self._lib.mdns_comm_allreduce(self._h, _addr(v), v.size, 1 if op == 'max' else 0),
                   'mdns_comm_allreduce')
        return v

    def allgather_candidate(self, k, nranks):
        """logL vectors of candidate k of the last launch of every rank -> (L_all, n_per_rank)."""
        nper = numpy.zeros(int(nranks), dtype=numpy.int32)
        _lib.check(self._lib.mdns_comm_allgather_candidate(self._h, int(k), None, 0, _addr(nper)),
                   'mdns_comm_allgather_candidate')
        out = _pool.empty(int(nper.sum()))
        _lib.check(self._lib.mdns_comm_allgather_candidate(self._h, int(k), _addr(out), out.size,
                                                           _addr(nper)),
                   'mdns_comm_allgather_candidate')
        return out, nper

    def set_draw_chunks(self, nchunks=0):
        _lib.check(self._lib.mdns_set_draw_chunks(self._h, int(nchunks)), 'mdns_set_draw_chunks')

    # -- staged interface ----------------------------------------------------
    def set_mask(self, data_mask):
        n = ctypes.c_int()
        if data_mask is not None:
            data_mask = self._mask(data_mask)
        _lib.check(self._lib.mdns_set_mask(self._h, _addr(data_mask), ctypes.byref(n)),
                   'mdns_set_mask')
        # the thresholds of a draw in progress are aligned with ITS mask (the shim keeps them when
        # the same mask comes again, mdns_set_mask); any other mask ends the draw
        if self._draw_n_act is not None and not self._same_as_draw_mask(data_mask):
            self._draw_n_act = None
        return n.value

    def _same_as_draw_mask(self, data_mask):
        dm = self._draw_mask
        if data_mask is None or dm is None:
            return data_mask is None and dm is None
        return data_mask is dm or numpy.array_equal(data_mask, dm)

    def _draw_state(self):
        if self._draw_n_act is None:
            raise _lib.MdnsError('no constrained draw in progress: call begin_draw(data_mask, Lmins) '
                               '(or draw_pass) before draw_batch / draw_counts / fetch_candidate')
        return self._draw_n_act

    def _mask(self, data_mask):
        m = numpy.ascontiguousarray(data_mask)
        if m.dtype != numpy.bool_:
            m = m.astype(numpy.bool_)
        if m.shape != (self.ndata,):
            raise ValueError('data_mask must have ndata entries')
        return m

    def stage_params(self, params):
        p = numpy.ascontiguousarray(params, dtype=numpy.float64).reshape((-1, 3))
        _lib.check(self._lib.mdns_stage_params(self._h, _addr(p), len(p)), 'mdns_stage_params')
        return len(p)

    def stage_spectra(self, ypred):
        s = numpy.ascontiguousarray(ypred, dtype=numpy.float64).reshape((-1, self.nx))
        _lib.check(self._lib.mdns_stage_spectra(self._h, _addr(s), len(s)), 'mdns_stage_spectra')
        return len(s)

    def launch_clike(self, noise, scale=-0.5):
        _lib.check(self._lib.mdns_clike_launch(self._h, noise, scale), 'mdns_clike_launch')

    def launch_muse(self):
        _lib.check(self._lib.mdns_muse_launch(self._h), 'mdns_muse_launch')

    def fetch(self, out):
        _lib.check(self._lib.mdns_fetch(self._h, _addr(out), out.size), 'mdns_fetch')
        return out

    def sync(self):
        _lib.check(self._lib.mdns_sync(self._h), 'mdns_sync')

    def timer_start(self):
        _lib.check(self._lib.mdns_timer_start(self._h), 'mdns_timer_start')

    def timer_stop(self):
        ms = ctypes.c_float()
        _lib.check(self._lib.mdns_timer_stop(self._h, ctypes.byref(ms)), 'mdns_timer_stop')
        return ms.value

    def flush_l2(self):
        _lib.check(self._lib.mdns_flush_l2(self._h), 'mdns_flush_l2')

    def set_tuning(self, lanes=0, unroll=0, ktile=0, rows=0):
        _lib.check(self._lib.mdns_set_tuning(self._h, lanes, unroll, ktile, rows),
                   'mdns_set_tuning')

    def set_expanded(self, enable=True, rel_tol=0.0):
        """Allow / forbid the expanded form Syy - 2 Sym + Smm of the candidate-batch kernel
        (K >= 3, all data sets active); rel_tol > 0 sets the error bound it enforces."""
        _lib.check(self._lib.mdns_set_expanded(self._h, 1 if enable else 0, rel_tol),
                   'mdns_set_expanded')

    def expanded_stats(self):
        """(enabled, rows recomputed in the direct form so far)"""
        en = ctypes.c_int()
        redo = ctypes.c_int64()
        _lib.check(self._lib.mdns_expanded_stats(self._h, ctypes.byref(en), ctypes.byref(redo)),
                   'mdns_expanded_stats')
        return bool(en.value), int(redo.value)

    # -- one-call forms ------------------------------------------------------
    def loglike_batch(self, params, data_mask, noise, scale=-0.5, out=None):
        """K parameter points (A, mu, sig) x all active data sets -> L[K, n_act]."""
        K = self.stage_params(params)
        n_act = self.set_mask(data_mask)
        if out is None:
            out = _pool.empty(K * n_act)
        if n_act > 0:
            _lib.check(self._lib.mdns_clike_launch_fetch(self._h, noise, scale, _addr(out), out.size),
                       'mdns_clike_launch_fetch')
        return out[:K * n_act].reshape((K, n_act))

    def first_accepted(self, params, data_mask, Lmins, noise, scale=-0.5):
        """Speculative batch of the constrained draw (hiermetriclearn.py:181-196).

        Scores the K parameter points in one pass and returns ``(k, L, counts)``: the index of
        the first candidate for which ``numpy.any(L > Lmins)`` holds (what the reference's
        one-candidate-at-a-time loop would have stopped at), its logL vector ``L[n_act]``, and
        the number of accepting data sets of every candidate.  ``(-1, None, counts)`` if none.
        """
        K = self.stage_params(params)
        n_act = self.set_mask(data_mask)
        Lmins = numpy.ascontiguousarray(Lmins, dtype=numpy.float64)
        if Lmins.shape != (n_act,):
            raise ValueError('Lmins must have one entry per active data set')
        counts = numpy.zeros(K, dtype=numpy.int32)
        if n_act == 0 and not self._comm:
            return -1, None, counts
        out = _pool.empty(n_act)
        first = ctypes.c_int(-1)
        _lib.check(self._lib.mdns_clike_first_accept(self._h, noise, scale, _addr(Lmins),
                                                     _addr(counts), ctypes.byref(first),
                                                     _addr(out), out.size),
                   'mdns_clike_first_accept')
        if first.value < 0:
            return -1, None, counts
        return first.value, out, counts

    def begin_draw(self, data_mask, Lmins):
        """Stage what stays constant during one constrained draw (hiermetriclearn.py:173-211):
        the data-set mask and the accept thresholds of the active data sets.  Returns n_act."""
        n_act = self.set_mask(data_mask)
        Lmins = numpy.ascontiguousarray(Lmins, dtype=numpy.float64)
        if Lmins.shape != (n_act,):
            raise ValueError('Lmins must have one entry per active data set')
        if n_act == 0:
            Lmins = numpy.zeros(1)      # nothing to compare with; keeps the call sequence uniform
        _lib.check(self._lib.mdns_set_thresholds(self._h, _addr(Lmins)), 'mdns_set_thresholds')
        self._draw_n_act = n_act
        self._draw_mask = None if data_mask is None else numpy.array(data_mask, dtype=bool, copy=True)
        return n_act

    def draw_batch(self, params, noise, scale=-0.5):
        """Score the next K candidates of the draw started with ``begin_draw``; returns
        ``(k, L, counts)`` like ``first_accepted``.  Per call only the K parameter points go to
        the device and K counts plus at most one logL vector come back."""
        n_act = self._draw_state()
        K = self.stage_params(params)
        counts = numpy.zeros(K, dtype=numpy.int32)
        if n_act == 0 and not self._comm:
            return -1, None, counts
        out = _pool.empty(n_act)
        first = ctypes.c_int(-1)
        _lib.check(self._lib.mdns_clike_first_accept(self._h, noise, scale, None, _addr(counts),
                                                     ctypes.byref(first), _addr(out), out.size),
                   'mdns_clike_first_accept')
        if first.value < 0:
            return -1, None, counts
        return first.value, out, counts

    def draw_pass(self, data_mask, Lmins, params, noise, scale=-0.5):
        """``begin_draw`` + ``draw_batch`` in one native call (one speculative pass of a
        constrained draw): a repeated mask and repeated short thresholds are recognised and not
        uploaded again.  Returns ``(k, L, counts)`` like ``draw_batch``."""
        p = numpy.ascontiguousarray(params, dtype=numpy.float64).reshape((-1, 3))
        K = len(p)
        counts = numpy.zeros(K, dtype=numpy.int32)
        Lmins = numpy.ascontiguousarray(Lmins, dtype=numpy.float64)
        if Lmins.ndim != 1:
            raise ValueError('Lmins must have one entry per active data set')
        if data_mask is None:
            m, n_act = None, self.ndata
        else:
            m = self._mask(data_mask)
            # the same mask object as last pass (the constrained draw keeps calling with its
            # joint_data_mask): the shim compares the bytes anyway, only the count is reused
            if m is self._draw_mask and self._draw_n_act is not None:
                n_act = self._draw_n_act
            else:
                n_act = int(numpy.count_nonzero(m))
        if Lmins.shape != (n_act,):
            if m is not None:
                n_act = int(numpy.count_nonzero(m))
            if Lmins.shape != (n_act,):
                raise ValueError('Lmins must have one entry per active data set')
        if n_act == 0 and not self._comm:
            self.set_mask(m)
            self._draw_n_act = 0
            self._draw_mask = None if m is None else m.copy()
            return -1, None, counts
        # (a few values are copied by the host anyway: no pinned block for them)
        out = _pool.empty(n_act) if n_act > 512 else numpy.empty(n_act)
        first = ctypes.c_int(-1)
        seen = ctypes.c_int(-1)
        _lib.check(self._lib.mdns_clike_draw_pass(self._h, _addr(m), _addr(Lmins), _addr(p), K, noise,
                                                  scale, _addr(counts), ctypes.byref(first),
                                                  _addr(out), out.size, ctypes.byref(seen)),
                   'mdns_clike_draw_pass')
        if seen.value != n_act:
            # the mask was edited in place since its entries were last counted
            self._draw_n_act = None
            raise ValueError('data_mask has %d active data sets, Lmins %d entries' % (seen.value, n_act))
        self._draw_n_act = n_act
        self._draw_mask = m          # (by reference: the shim itself compares the bytes on every pass)
        if first.value < 0:
            return -1, None, counts
        return first.value, out, counts

    def draw_counts(self, params, noise, scale=-0.5):
        """Counts only: score the next K candidates of the draw started with ``begin_draw`` and
        return the accept counts -- of THIS process's data sets, or summed over the ranks when a
        communicator is attached (``init_comm``).  The logL vectors stay on the device
        (``fetch_candidate(k)`` downloads one; ``LiveTable`` consumes them in place).  Without a
        communicator a caller may run the exchange itself: ``sharding.global_first_accepted``."""
        n_act = self._draw_state()
        K = self.stage_params(params)
        counts = numpy.zeros(K, dtype=numpy.int32)
        if n_act > 0 or self._comm:
            _lib.check(self._lib.mdns_clike_accept_counts(self._h, noise, scale, None, _addr(counts)),
                       'mdns_clike_accept_counts')
        return counts

    def fetch_candidate(self, k):
        """Second step: the logL vector of candidate k of the launch behind ``draw_counts``."""
        out = _pool.empty(self._draw_state())
        if self._draw_n_act > 0:
            _lib.check(self._lib.mdns_fetch_candidate(self._h, int(k), _addr(out), out.size),
                       'mdns_fetch_candidate')
        return out

    def draw_batch_sparse(self, params, noise, scale=-0.5):
        """Like ``draw_batch`` but returns only what the sampler consumes
        (multi_nested_sampler.py:482-485): ``(k, j, Lj, counts)`` with ``j`` the positions (in the
        compacted active order, increasing) of the data sets candidate ``k`` is accepted for and
        ``Lj`` their logL; ``(-1, None, None, counts)`` if no candidate is accepted."""
        n_act = self._draw_state()
        K = self.stage_params(params)
        counts = numpy.zeros(K, dtype=numpy.int32)
        if n_act == 0 and not self._comm:
            return -1, None, None, counts
        idx = numpy.empty(max(n_act, 1), dtype=numpy.int32)
        val = _pool.empty(max(n_act, 1))
        first = ctypes.c_int(-1)
        n = ctypes.c_int(0)
        _lib.check(self._lib.mdns_clike_first_accept_sparse(
            self._h, noise, scale, None, _addr(counts), ctypes.byref(first), _addr(idx), _addr(val),
            n_act, ctypes.byref(n)), 'mdns_clike_first_accept_sparse')
        if first.value < 0:
            return -1, None, None, counts
        return first.value, idx[:n.value], val[:n.value], counts

    def loglike_spectra(self, ypred, data_mask, noise, scale=-0.5, out=None):
        """K model spectra x all active data sets -> L[K, n_act] (scalar-noise chi-square)."""
        K = self.stage_spectra(ypred)
        n_act = self.set_mask(data_mask)
        if out is None:
            out = _pool.empty(K * n_act)
        if n_act > 0:
            _lib.check(self._lib.mdns_clike_launch_fetch(self._h, noise, scale, _addr(out), out.size),
                       'mdns_clike_launch_fetch')
        return out[:K * n_act].reshape((K, n_act))

    def muse_loglike(self, ypred, data_mask, Lout):
        """cmuselike.c semantics: writes -0.5*chi2 into the masked entries of Lout[K, ndata]."""
        K = self.stage_spectra(ypred)
        n_act = self.set_mask(data_mask)
        if n_act > 0:
            self.launch_muse()
            _lib.check(self._lib.mdns_fetch(self._h, _addr(Lout), K * self.ndata), 'mdns_fetch')
        return Lout


def make_multi_loglikelihood(x, y, noise_level=0.01, devices=None):
    """Build the callable of sample.py:101-108 over a resident copy of (x, y).

    Returns ``multi_loglikelihood(params, data_mask)`` with ``params = (A, mu, log_sig)``;
    the attribute ``.dataset`` exposes the ResidentDataset, ``.batch(params_list, data_mask)``
    evaluates several parameter vectors in one pass, and ``.speculate`` / ``.last_draw`` are the
    hook massivedatans_b200.hiermetriclearn uses to batch a constrained draw through the
    sampler's lambda.
    """
    ds = ResidentDataset(x, y, devices=devices)
    p = numpy.empty((1, 3))

    pending = []

    def _points(params_list):
        q = numpy.array(params_list, dtype=numpy.float64).reshape((-1, 3))
        # scalar power per point, as sample.py:103 computes it: numpy's vectorised pow rounds
        # differently from the scalar one in ~5 % of the cases; math.pow is the same libm call
        # as the numpy scalar power
        q[:, 2] = [math.pow(10.0, v) for v in q[:, 2].tolist()]
        return q

    def multi_loglikelihood(params, data_mask):
        A, mu, log_sig_kms = params
        if pending:
            # a batch announced by speculate(): the mask only arrives with this call
            first, q, Lmins = pending.pop()
            if first[0] == A and first[1] == mu and first[2] == log_sig_kms:
                k, L, counts = ds.draw_pass(data_mask, Lmins, q, noise_level)
                multi_loglikelihood.last_draw = (k, L, counts)
                return L if k == 0 else None
            # some other caller came in between: the announcement is void
            multi_loglikelihood.last_draw = None
        p[0, 0] = A
        p[0, 1] = mu
        p[0, 2] = 10 ** log_sig_kms
        return ds.loglike_batch(p, data_mask, noise_level)[0]

    def batch(params_list, data_mask):
        return ds.loglike_batch(_points(params_list), data_mask, noise_level)

    def speculate(params_list, Lmins):
        """Announce the next candidates of a constrained draw (hiermetriclearn.py:181-196) and
        the thresholds they are tested against.  The NEXT ``multi_loglikelihood(params_list[0],
        data_mask)`` call scores all of them in one pass with the accept test on the device and
        leaves ``(k, L_k, counts)`` in ``.last_draw``: k = first candidate with
        ``any(L > Lmins)`` (-1: none), L_k = its logL vector, counts = accepting data sets per
        candidate.  That call returns L_0 only if k == 0 (no other logL vector leaves the
        device), else None."""
        del pending[:]
        multi_loglikelihood.last_draw = None
        first = numpy.array(params_list[0], dtype=numpy.float64)
        pending.append((first, _points(params_list), numpy.array(Lmins, dtype=numpy.float64)))

    multi_loglikelihood.dataset = ds
    multi_loglikelihood.batch = batch
    multi_loglikelihood.speculate = speculate
    multi_loglikelihood.last_draw = None
    return multi_loglikelihood


def make_muse_loglikelihood(y, noise_level, model, jitter=1e-5, devices=None):
    """Build ``multi_loglikelihood_clike`` of musefuse.py:520-535.

    `model(*params)` returns the model spectrum ypred[nspec] on the host
    (musefuse.py:222-284).  `noise_level` is the per-element variance matrix.
    """
    ds = ResidentDataset(None, y, variance=noise_level, devices=devices)
    Lout = numpy.zeros(ds.ndata)       # persistent, as the global at musefuse.py:519

    def multi_loglikelihood_clike(params, data_mask):
        ypred = model(*params)
        if not numpy.any(ypred):
            # musefuse.py:528-530 -- give low probability to solutions with no stars
            return numpy.ones(data_mask.sum()) * -1e100
        ds.muse_loglike(ypred, data_mask, Lout)
        res = Lout[data_mask]
        if jitter:
            res = res + numpy.random.normal(0, jitter, size=data_mask.sum())
        return res

    multi_loglikelihood_clike.dataset = ds
    multi_loglikelihood_clike.Lout = Lout
    return multi_loglikelihood_clike


def calzetti(model_wavelength_nm):
    """Calzetti attenuation curve on the template grid in nm (musefuse.py:208-217): host-side
    table built once per model, like the reference's module-level ``calzetti_result``."""
    w = numpy.asarray(model_wavelength_nm, dtype=numpy.float64)
    k = numpy.zeros_like(w)
    blue = w < 630
    k[blue] = 2.659 * (-2.156 + 1.509e3 / w[blue] - 0.198e6 / w[blue] ** 2
                       + 0.011e9 / w[blue] ** 3) + 4.05
    red = w >= 630
    k[red] = 2.659 * (-1.857 + 1.040e3 / w[red]) + 4.05
    return k


class DeviceMuseModel(object):
    """``model(Z, SFtau, sfage, z, EBV)`` of musefuse.py:222-284 evaluated on the devices of a
    ResidentDataset for batches of parameter points; the spectra are staged where
    ``mdns_muse_launch`` reads them and never visit the host.

    grids : [nZ, nages, nwave] template grids (musefuse.py:176-185, one file per metallicity)
    Zs, ages, model_wavelength : their axes (musefuse.py:188,191,181/207); wavelengths in the unit
        of ``wavelength``, the data grid (musefuse.py:82/206)
    """

    def __init__(self, dataset, grids, Zs, ages, model_wavelength, wavelength, calzetti_curve=None,
                 norm_index=2050):
        lib = _lib.load()
        grids = numpy.ascontiguousarray(grids, dtype=numpy.float64)
        if grids.ndim != 3:
            raise ValueError('grids must be [nZ, nages, nwave]')
        nZ, nages, nwave = grids.shape
        Zs = numpy.ascontiguousarray(Zs, dtype=numpy.float64)
        ages = numpy.ascontiguousarray(ages, dtype=numpy.float64)
        mw = numpy.ascontiguousarray(model_wavelength, dtype=numpy.float64)
        wl = numpy.ascontiguousarray(wavelength, dtype=numpy.float64)
        if Zs.shape != (nZ,) or ages.shape != (nages,) or mw.shape != (nwave,):
            raise ValueError('Zs, ages, model_wavelength must match the axes of grids')
        if wl.shape != (dataset.nx,):
            raise ValueError('wavelength must have one entry per channel of the data set')
        if calzetti_curve is None:
            calzetti_curve = calzetti(mw)
        cz = numpy.ascontiguousarray(calzetti_curve, dtype=numpy.float64)
        if cz.shape != (nwave,):
            raise ValueError('calzetti_curve must have nwave entries')
        handle = ctypes.c_void_p()
        _lib.check(lib.mdns_muse_model_create(dataset._h, _addr(grids), nZ, _addr(Zs), nages,
                                              _addr(ages), nwave, _addr(mw), _addr(cz), _addr(wl),
                                              dataset.nx, int(norm_index), ctypes.byref(handle)),
                   'mdns_muse_model_create')
        self._lib = lib
        self._h = handle
        self._dataset = dataset
        self.nx = dataset.nx
        self._last_K = 0
        self._finalizer = weakref.finalize(self, lib.mdns_muse_model_destroy, handle)

    def close(self):
        self._finalizer()

    def stage(self, params):
        """params[K, 5] = (Z, SFtau, sfage, z, EBV); returns nonzero[K] (numpy.any per spectrum)."""
        p = numpy.ascontiguousarray(params, dtype=numpy.float64).reshape((-1, 5))
        self._last_K = len(p)
        nonzero = numpy.zeros(len(p), dtype=numpy.int32)
        _lib.check(self._lib.mdns_muse_model_stage(self._h, _addr(p), len(p), _addr(nonzero)),
                   'mdns_muse_model_stage')
        return nonzero != 0

    def spectra(self):
        """The spectra of the last ``stage`` call, [K, nx] (for checks; costs a download)."""
        if self._last_K <= 0:
            raise RuntimeError('no spectra staged yet: call stage(params) first')
        out = numpy.empty((self._last_K, self.nx))
        _lib.check(self._lib.mdns_muse_model_spectra(self._h, _addr(out)), 'mdns_muse_model_spectra')
        return out

    def __call__(self, Z, SFtau, sfage, z, EBV):
        """One model spectrum on the host, the reference's call signature (musefuse.py:222)."""
        self.stage([[Z, SFtau, sfage, z, EBV]])
        return self.spectra()[0]

    def loglike_batch(self, params, data_mask, Lout):
        """K parameter points -> ``Lout[K, ndata]`` as cmuselike.c writes it (masked entries),
        model and likelihood on the device; returns nonzero[K]."""
        p = numpy.ascontiguousarray(params, dtype=numpy.float64).reshape((-1, 5))
        nonzero = self.stage(p)
        ds = self._dataset
        if ds.set_mask(data_mask) > 0:
            ds.launch_muse()
            _lib.check(self._lib.mdns_fetch(ds._h, _addr(Lout), len(p) * ds.ndata), 'mdns_fetch')
        return nonzero


def make_muse_loglikelihood_device(y, noise_level, grids, Zs, ages, model_wavelength, wavelength,
                                   jitter=1e-5, devices=None, norm_index=2050):
    """``multi_loglikelihood_clike`` of musefuse.py:520-535 with the model on the device too.

    Returns ``multi_loglikelihood_clike(params, data_mask)``, params = (Z, logSFtau, SFage, z, EBV);
    ``.batch(params_list, data_mask)`` scores several parameter vectors in one pass and returns
    ``L[K, n_act]`` WITHOUT the jitter term (one draw per call in the reference; the caller adds
    it where it wants the reference's stream); ``.model`` is the DeviceMuseModel.
    """
    ds = ResidentDataset(None, y, variance=noise_level, devices=devices)
    model = DeviceMuseModel(ds, grids, Zs, ages, model_wavelength, wavelength, norm_index=norm_index)
    Lout = numpy.zeros(ds.ndata)       # persistent, as the global at musefuse.py:519

    def _points(params_list):
        q = numpy.array(params_list, dtype=numpy.float64).reshape((-1, 5))
        # SFtau = 10**logSFtau, a scalar power as at musefuse.py:524
        q[:, 1] = [math.pow(10.0, v) for v in q[:, 1].tolist()]
        return q

    def multi_loglikelihood_clike(params, data_mask):
        nonzero = model.loglike_batch(_points([params]), data_mask, Lout.reshape((1, -1)))
        if not nonzero[0]:
            # musefuse.py:528-530 -- give low probability to solutions with no stars
            return numpy.ones(data_mask.sum()) * -1e100
        res = Lout[data_mask]
        if jitter:
            res = res + numpy.random.normal(0, jitter, size=data_mask.sum())
        return res

    def batch(params_list, data_mask):
        q = _points(params_list)
        full = numpy.zeros((len(q), ds.ndata))
        nonzero = model.loglike_batch(q, data_mask, full)
        res = full[:, numpy.asarray(data_mask, dtype=bool)]
        res[~nonzero, :] = -1e100
        return res

    multi_loglikelihood_clike.dataset = ds
    multi_loglikelihood_clike.model = model
    multi_loglikelihood_clike.batch = batch
    return multi_loglikelihood_clike
