"""Test harness: a fixed-seed sequence of constrained draws through a constrainer object.

Drives ``constrainer.draw_constrained`` the way multi_nested_sampler.py:415-489,494-534 does --
joint draws for a group of data sets with thresholds from a per-data-set live-point table, the
accepted point replacing the worst live point of every data set it is accepted for -- without
the shelves, graph and integrator of the sampler (callers, not rebuilt).  The same function runs

* the REFERENCE's hiermetriclearn.MetricLearningFriendsConstrainer on the reference's C
  libraries (tests/golden/make_golden_constrainer.py, build container only) -> the fixture, and
* massivedatans_b200.hiermetriclearn.MetricLearningFriendsConstrainer (GPU tests; CPU tests
  with the neighbour queries and the likelihood answered by the oracle),

and because every random number comes from the global numpy stream, the two must agree draw by
draw: same candidates, same number of tries, same accepted point.
"""
import numpy

NDIM = 3


def priortransform(cube):
    """sample.py:52-58"""
    cube = cube.copy()
    cube[0] = 10 ** (cube[0] * 2 - 2)
    cube[1] = cube[1] * 400 + 400
    cube[2] = cube[2] * 2
    return cube


def group_mask(it, ndata, rs):
    """Which data sets draw jointly at iteration `it`: all / a random half / a single one."""
    mask = numpy.zeros(ndata, dtype=bool)
    kind = it % 4
    if kind in (0, 1):
        mask[:] = True
    elif kind == 2:
        mask[rs.permutation(ndata)[:max(1, ndata // 2)]] = True
    else:
        mask[rs.randint(ndata)] = True
    return mask


def run_draws(constrainer, multi_loglikelihood, ndata, nlive, niter, seed, rank=0):
    """Returns dict(u[niter, 3], ntoaccept[niter], L[niter, ndata] (nan where not drawn for),
    naccepted[niter]); `multi_loglikelihood(params, data_mask)` as sample.py:101-108.
    `rank`: which live likelihood of every data set is the threshold -- 0 = the minimum (nested
    sampling), -1 = the maximum (long rejection chains: exercises the in-loop rebuild triggers of
    hiermetriclearn.py:200-211)."""
    numpy.random.seed(seed)
    rs = numpy.random.RandomState(seed + 1000)      # group choice: not part of the shared stream
    everyone = numpy.ones(ndata, dtype=bool)
    pile = []
    live_L = numpy.empty((nlive, ndata))
    for i in range(nlive):                          # multi_nested_sampler.py:91-103
        u = numpy.random.uniform(0, 1, size=NDIM)
        pile.append(u)
        live_L[i] = multi_loglikelihood(priortransform(u), everyone)
    live_p = numpy.repeat(numpy.arange(nlive)[:, None], ndata, axis=1)
    out = dict(u=numpy.empty((niter, NDIM)), ntoaccept=numpy.zeros(niter, dtype=int),
               L=numpy.full((niter, ndata), numpy.nan), naccepted=numpy.zeros(niter, dtype=int))
    for it in range(niter):
        mask = group_mask(it, ndata, rs)
        members = numpy.unique(live_p[:, mask])     # multi_nested_sampler.py:139-143
        live_u = numpy.array([pile[i] for i in members])
        worst = live_L.argmin(axis=0)
        Lmins = (live_L.min(axis=0) if rank == 0 else numpy.sort(live_L, axis=0)[rank])[mask]
        u, x, L, ntoaccept = constrainer.draw_constrained(
            Lmins=Lmins, priortransform=priortransform,
            loglikelihood=lambda p: multi_loglikelihood(p, mask),
            live_pointsu=live_u, ndim=NDIM, iter=it, nlive_points=nlive, max_draws=100)
        L = numpy.asarray(L)
        accepted = L > Lmins
        assert accepted.any()
        pile.append(numpy.array(u))
        for j, ok, Lj in zip(numpy.where(mask)[0], accepted, L):
            if ok:                                   # multi_nested_sampler.py:520-524
                live_p[worst[j], j] = len(pile) - 1
                live_L[worst[j], j] = Lj
        out['u'][it] = u
        out['ntoaccept'][it] = ntoaccept
        out['L'][it, mask] = L
        out['naccepted'][it] = accepted.sum()
    return out
