"""Test harness: a compact collaborative nested sampler on top of the hot-path API.

This is NOT a rebuild of the reference's sampler (multi_nested_sampler.py, hiermetriclearn.py,
cachedconstrainer.py stay callers that keep their signatures).  It is the
smallest correct joint sampler that exercises the two replaced pieces the way those callers
do, so that a fixed-seed end-to-end run can be checked on the GPU box, where the reference
sources do not exist:

* one RadFriends region (union of balls of radius `bootstrapped_maxdistance` around the
  metric-scaled live points, radfriendsregion.py:58-70,117-182) shared by all data sets,
  sampled with ball draws thinned by 1/`count_within_distance_of`
  (radfriendsregion.py:156-178);
* every candidate is scored against all data sets that still need a point in ONE masked
  batched likelihood call (multi_nested_sampler.py:373-388,462-472) and accepted for each data
  set whose threshold it exceeds (hiermetriclearn.py:193, multi_nested_sampler.py:476-489);
  surplus accepted points wait on per-data-set shelves (multi_nested_sampler.py:117,134-143);
* per data set the usual nested-sampling evidence sum with the live-point remainder
  (multi_nested_integrator.py:26-60,98-157, simplified to the standard estimator).

A uniform draw from a region that contains the likelihood contours of several data sets is a
valid constrained draw for each of them (the reference's core idea, README.rst:5-7), so the
evidences are unbiased; sharing one region only costs efficiency, which does not matter at the
sizes used in tests.

Backends provide
    loglike(points[K,3] (A, mu, log10 sig), mask[ndata]) -> L[K, n_act]
    count_within_distance_of(members, r, candidates) -> int[len(candidates)]
    bootstrapped_maxdistance(members, nbootstraps) -> float      (draws from numpy.random)
All randomness comes from the global numpy.random stream (as in the reference), so two
backends that agree on every comparison reproduce each other's run draw by draw.
"""
import numpy


def priortransform(u):
    """sample.py:52-58 for an array of unit-cube points [K, 3] -> (A, mu, log10 sig)."""
    x = numpy.array(u, dtype=float, copy=True)
    x[:, 0] = 10 ** (x[:, 0] * 2 - 2)
    x[:, 1] = x[:, 1] * 400 + 400
    x[:, 2] = x[:, 2] * 2
    return x


class OracleBackend(object):
    """CPU oracle (oracle/port.py) behind the harness interface."""

    def __init__(self, x, y, noise):
        from oracle import port
        self.port = port
        self.x, self.y, self.noise = x, y, noise

    def loglike(self, points, mask):
        mask = numpy.ascontiguousarray(mask)
        out = numpy.empty((len(points), int(mask.sum())))
        for k, (A, mu, log_sig) in enumerate(points):
            out[k] = -0.5 * self.port.clike(self.x, self.y, A, mu, 10 ** log_sig, self.noise, mask)
        return out

    def count_within_distance_of(self, members, r, candidates):
        return self.port.count_within_distance_of(numpy.ascontiguousarray(members), r,
                                                  numpy.ascontiguousarray(candidates))

    def bootstrapped_maxdistance(self, members, nbootstraps):
        members = numpy.ascontiguousarray(members)
        n = len(members)
        chosen = numpy.zeros((n, nbootstraps))          # clustering/neighbors.py:172-174
        for b in range(nbootstraps):
            chosen[numpy.random.choice(numpy.arange(n), size=n, replace=True), b] = 1.
        return self.port.bootstrapped_maxdistance_chosen(members, chosen)


class GpuBackend(object):
    """The product path: resident data set + neighbour kernels through the public mirrors."""

    def __init__(self, x, y, noise, devices=None):
        from massivedatans_b200.clustering import neighbors
        from massivedatans_b200.likelihood import make_multi_loglikelihood
        self.f = make_multi_loglikelihood(x, y, noise, devices=devices)
        self.neighbors = neighbors

    def loglike(self, points, mask):
        if len(points) == 1:
            return self.f(tuple(points[0]), mask).reshape((1, -1)).copy()
        return numpy.array(self.f.batch([tuple(p) for p in points], mask))

    def count_within_distance_of(self, members, r, candidates):
        return self.neighbors.count_within_distance_of(members, r, candidates)

    def bootstrapped_maxdistance(self, members, nbootstraps):
        return self.neighbors.bootstrapped_maxdistance(members, nbootstraps)


class Region(object):
    """RadFriends region in a per-axis scaled space (radfriendsregion.py:58-70)."""

    def __init__(self, backend, members_u, nbootstraps=10):
        self.backend = backend
        self.scale = members_u.std(axis=0)
        self.scale[self.scale == 0] = 1.0
        self.members = numpy.ascontiguousarray(members_u / self.scale)
        self.r = backend.bootstrapped_maxdistance(self.members, nbootstraps)
        self.ndim = members_u.shape[1]

    def draw(self, n):
        """n ball draws (radfriendsregion.py:156-178); returns the accepted unit-cube points."""
        centre = self.members[numpy.random.randint(len(self.members), size=n)]
        direction = numpy.random.normal(size=(n, self.ndim))
        direction /= numpy.sqrt((direction ** 2).sum(axis=1)).reshape((-1, 1))
        radius = self.r * numpy.random.uniform(size=n) ** (1.0 / self.ndim)
        w = centre + direction * radius.reshape((-1, 1))
        u = w * self.scale
        coin = numpy.random.uniform(size=n)
        inside = numpy.logical_and(u > 0, u < 1).all(axis=1)
        if not inside.any():
            return u[:0]
        w_in = numpy.ascontiguousarray(w[inside])
        nnear = self.backend.count_within_distance_of(self.members, self.r, w_in)
        keep = coin[inside] * nnear < 1.0          # accept with probability 1/nnear (nnear >= 1)
        return u[inside][keep]


def run(backend, ndata, nlive=100, niter=600, batch=8, rebuild_every=25, seed=1, ndim=3,
        proposals=400):
    """Joint nested sampling run; returns dict(logZ, logZerr, H, ndraws, nbatches, trace)."""
    numpy.random.seed(seed)
    allmask = numpy.ones(ndata, dtype=bool)
    pile_u = [numpy.random.uniform(size=ndim) for _ in range(nlive)]
    live_idx = numpy.tile(numpy.arange(nlive).reshape((-1, 1)), (1, ndata))
    live_L = numpy.empty((nlive, ndata))
    u0 = numpy.array(pile_u)
    for i0 in range(0, nlive, 64):                       # initial population: batched, all active
        live_L[i0:i0 + 64] = backend.loglike(priortransform(u0[i0:i0 + 64]), allmask)
    ndraws = nlive
    nbatches = 0
    shelves = [[] for _ in range(ndata)]
    logZ = numpy.full(ndata, -numpy.inf)
    H = numpy.zeros(ndata)
    region = None
    pending = numpy.zeros((0, ndim))
    trace = []
    dead_u, dead_logw = [], []
    for it in range(niter):
        worst = live_L.argmin(axis=0)
        Lmin = live_L[worst, numpy.arange(ndata)]
        for d in range(ndata):
            shelves[d] = [e for e in shelves[d] if e[1] > Lmin[d]]
        need = numpy.array([len(s) == 0 for s in shelves])
        if region is None or it % rebuild_every == 0:
            members = numpy.array([pile_u[i] for i in numpy.unique(live_idx)])
            region = Region(backend, members)
            pending = numpy.zeros((0, ndim))
        # the first batch of an iteration comes from the region of all data sets (superset
        # draw, multi_nested_sampler.py:373-376); what is still missing afterwards is drawn
        # from a region around the live points of just those data sets (focussed draw,
        # multi_nested_sampler.py:377-388), rebuilt whenever their number has halved
        current, current_pending, focus_size = region, pending, None
        first = True
        while need.any():
            if not first and (focus_size is None or need.sum() * 2 <= focus_size):
                focus_size = int(need.sum())
                members = numpy.array([pile_u[i] for i in numpy.unique(live_idx[:, need])])
                current = Region(backend, members)
                current_pending = numpy.zeros((0, ndim))
            while len(current_pending) < batch:
                current_pending = numpy.vstack([current_pending, current.draw(proposals)])
            cand_u, current_pending = current_pending[:batch], current_pending[batch:]
            if first:
                pending = current_pending
            first = False
            L = backend.loglike(priortransform(cand_u), need)
            nbatches += 1
            need_ids = numpy.nonzero(need)[0]
            acc = L > Lmin[need_ids]
            # candidates are consumed in order until every data set in `need` has a point
            # (hiermetriclearn.py:181-196 stops at the first accepted candidate; the rest of a
            # batch is still a set of valid uniform draws and is queued like the reference's
            # surplus points, multi_nested_sampler.py:476-489)
            ndraws += len(cand_u)
            for k in numpy.nonzero(acc.any(axis=1))[0]:
                pile_u.append(cand_u[k])
                for j in numpy.nonzero(acc[k])[0]:
                    shelves[need_ids[j]].append((len(pile_u) - 1, L[k, j]))
            need = numpy.array([len(s) == 0 for s in shelves])
        # advance every data set by one dead point (multi_nested_sampler.py:494-534)
        logw = numpy.log(numpy.exp(-it / float(nlive)) - numpy.exp(-(it + 1.0) / nlive))
        wi = logw + Lmin
        logZnew = numpy.logaddexp(logZ, wi)
        with numpy.errstate(invalid='ignore'):
            Hnew = (numpy.exp(wi - logZnew) * Lmin
                    + numpy.where(numpy.isfinite(logZ), numpy.exp(logZ - logZnew) * (H + logZ), 0.0)
                    - logZnew)
        H, logZ = Hnew, logZnew
        # weighted posterior samples: the dead point of every data set with logwidth + L
        # (multi_nested_integrator.py:118, plotposterior.py:21)
        dead_u.append(numpy.array([pile_u[i] for i in live_idx[worst, numpy.arange(ndata)]]))
        dead_logw.append(wi)
        for d in range(ndata):
            idx, Ld = shelves[d].pop(0)
            live_idx[worst[d], d] = idx
            live_L[worst[d], d] = Ld
        if it % 50 == 0:
            trace.append((it, float(Lmin[0]), ndraws))
    # remainder: the live points share the remaining volume exp(-niter/nlive)
    logw = -niter / float(nlive) - numpy.log(nlive)
    Lmax = live_L.max(axis=0)
    rest = logw + Lmax + numpy.log(numpy.exp(live_L - Lmax).sum(axis=0))
    logZ = numpy.logaddexp(logZ, rest)
    H = numpy.maximum(H, 0.0)
    # posterior moments per data set in (log10 A, mu, log10 sig), dead points + live remainder
    # (multi_nested_integrator.py:163-170 appends the live points as the tail)
    live_u = numpy.array([[pile_u[i] for i in live_idx[:, d]] for d in range(ndata)])  # [ndata, nlive, ndim]
    samples_u = numpy.concatenate([numpy.array(dead_u), live_u.transpose((1, 0, 2))], axis=0)
    samples_lw = numpy.concatenate([numpy.array(dead_logw), logw + live_L], axis=0)     # [ns, ndata]
    feat = priortransform(samples_u.reshape((-1, ndim))).reshape(samples_u.shape)
    feat[:, :, 0] = numpy.log10(feat[:, :, 0])
    p = numpy.exp(samples_lw - samples_lw.max(axis=0))
    p /= p.sum(axis=0)
    post_mean = (p[:, :, None] * feat).sum(axis=0)
    post_std = numpy.sqrt((p[:, :, None] * (feat - post_mean) ** 2).sum(axis=0))
    post_ess = 1.0 / (p ** 2).sum(axis=0)
    return dict(post_mean=post_mean, post_std=post_std, post_ess=post_ess, logZ=logZ, logZerr=numpy.sqrt(H / nlive), H=H, ndraws=ndraws, nbatches=nbatches,
                trace=trace, remainder_fraction=numpy.exp(rest - logZ))
