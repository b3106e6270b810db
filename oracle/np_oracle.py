"""numpy/scipy restatement of the hot-path formulas (test infrastructure).

Not bit-identical to the C (different summation order); agrees to ~1e-15
relative for the likelihoods and exactly for the neighbour counts except at
floating-point ties of the distance threshold.
"""
import numpy
import scipy.spatial


def line_model(x, A, mu, sig):
    """clike.c:65 / sample.py:68 -- Gaussian line on the wavelength grid."""
    return A * numpy.exp(-0.5 * ((mu - x) / sig) ** 2)


def clike(x, y, A, mu, sig, noise, data_mask):
    """sample.py:64-71 without the -0.5 (what clike.c accumulates in Lout)."""
    ypred = line_model(x, A, mu, sig)
    return (((ypred.reshape((-1, 1)) - y[:, data_mask]) / noise) ** 2).sum(axis=0)


def clike_spectrum(ypred, y, noise, data_mask):
    return (((ypred.reshape((-1, 1)) - y[:, data_mask]) / noise) ** 2).sum(axis=0)


def cmuselike(y, v, ypred, data_mask):
    """musefuse.py:474-480 (vectorised twin of cmuselike.c:48-64), without the
    random jitter term; returns the full-length vector with NaN outside the mask."""
    yd = y[:, data_mask]
    vd = v[:, data_mask]
    yp = ypred.reshape((-1, 1))
    s = numpy.sum(yd * yp / vd, axis=0) / (numpy.sum(yp ** 2 / vd, axis=0) + 1e-10)
    chi2 = numpy.sum((yd - s.reshape((1, -1)) * yp) ** 2 / vd, axis=0)
    out = numpy.full(y.shape[1], numpy.nan)
    out[data_mask] = -0.5 * chi2
    return out


def count_within_distance_of(members, maxdistance, us):
    """clustering/neighbors.py:79-81."""
    dists = scipy.spatial.distance.cdist(members, us, metric='euclidean')
    return (dists < maxdistance).sum(axis=0)


def any_within_distance_of(members, maxdistance, us):
    """clustering/neighbors.py:83-85."""
    dists = scipy.spatial.distance.cdist(members, us, metric='euclidean')
    return (dists < maxdistance).any(axis=0)


def most_distant_nearest_neighbor(u):
    """clustering/neighbors.py:188-193 (scipy branch of nearest_rdistance_guess)."""
    distances = scipy.spatial.distance.cdist(u, u, metric='euclidean')
    numpy.fill_diagonal(distances, 1e300)
    return numpy.max(numpy.min(distances, axis=1))


def bootstrapped_maxdistance(u, chosen):
    """Textbook restatement of cneighbors.c:140-176 with the i>=1 quirk (:162)."""
    n, nboot = chosen.shape
    d = scipy.spatial.distance.cdist(u, u, metric='sqeuclidean')
    best = None
    for b in range(nboot):
        sel = chosen[:, b] != 0
        unsel = ~sel
        unsel[0] = False
        if unsel.any():
            if sel.any():
                nearest = d[numpy.ix_(unsel, sel)].min(axis=1)
            else:
                nearest = numpy.full(unsel.sum(), 1e300)
            furthest = numpy.sqrt(nearest).max()
        else:
            furthest = 0.0
        best = furthest if best is None or furthest > best else best
    return best


def live_colstats(live_pointsL):
    """multi_nested_sampler.py:134-137 and :531."""
    return live_pointsL.min(axis=0), live_pointsL.argmin(axis=0), live_pointsL.max(axis=0)


def find_nsmallest(n, arr1, arr2):
    """multi_nested_sampler.py:44-47 (the numpy.partition version)."""
    arr = numpy.concatenate((arr1, arr2))
    return numpy.partition(arr, n)[n]


def subsets_labels(live_pointsp, data_mask, npoints):
    """Groups of multi_nested_sampler.py:204-355 as labels (smallest member index, -1 outside
    the mask), via scipy's connected components of the data set <-> live point graph."""
    import scipy.sparse
    import scipy.sparse.csgraph
    nlive, ndata = live_pointsp.shape
    sel = numpy.nonzero(data_mask)[0]
    rows = numpy.repeat(sel, nlive)
    cols = ndata + live_pointsp[:, sel].T.reshape(-1)
    g = scipy.sparse.coo_matrix((numpy.ones(len(rows)), (rows, cols)),
                                shape=(ndata + npoints, ndata + npoints))
    _, comp = scipy.sparse.csgraph.connected_components(g, directed=False)
    labels = numpy.full(ndata, -1, dtype=numpy.int32)
    first = {}
    for d in sel:
        first.setdefault(comp[d], d)
        labels[d] = first[comp[d]]
    return labels


def calzetti(model_wavelength_nm):
    """musefuse.py:208-217 -- Calzetti attenuation curve k(lambda) on the template grid (nm)."""
    w = numpy.asarray(model_wavelength_nm, dtype=float)
    k = numpy.zeros_like(w)
    blue = w < 630
    k[blue] = 2.659 * (-2.156 + 1.509e3 / w[blue] - 0.198e6 / w[blue] ** 2
                       + 0.011e9 / w[blue] ** 3) + 4.05
    red = w >= 630
    k[red] = 2.659 * (-1.857 + 1.040e3 / w[red]) + 4.05
    return k


def muse_model(Zs, ages, model_wavelength_nm, calz, grids, wavelength_nm, Z, SFtau, sfage, z, EBV,
               norm_index=2050):
    """musefuse.py:222-284 -- the stellar-population model spectrum on the data grid:
    metallicity bin, delayed-exponential star-formation history over the template ages,
    SFH-weighted template sum, normalisation at one template channel, Calzetti extinction at
    rest frame, linear interpolation onto the redshifted data wavelengths."""
    iZ = numpy.where(Zs <= Z)[-1][-1]
    templates = grids[iZ]
    t = sfage * 1.e9 - ages
    t[t <= 0] = 0
    SFtau = float(SFtau)
    sfh = t / SFtau ** 2 * numpy.exp(-t / SFtau)
    sfh /= sfh.max()
    dage = ages[1:] - ages[:-1]
    spec = numpy.sum(templates[:-1] * sfh[:-1].reshape((-1, 1)) * dage.reshape((-1, 1)), axis=0)
    spec /= 1e-10 + spec[norm_index]
    spec = spec * 10 ** (-2.5 * calz * EBV)
    return numpy.interp(x=wavelength_nm / (1 + z), xp=model_wavelength_nm, fp=spec)


def initial_maxdistance_guess(u):
    """clustering/neighbors.py:22-29 -- per axis, the largest |delta| between a point and its
    nearest neighbour (second entry of the sorted cdist row)."""
    dist = scipy.spatial.distance.cdist(u, u)
    nearest = numpy.array([dist[i, :].argsort()[1] for i in range(len(u))])
    return numpy.abs(u[nearest, :] - u).max(axis=0)


def update_maxdistance(u, maxdistance):
    """clustering/neighbors.py:31-62 -- one bootstrap round of the per-axis box (draws one
    numpy.random.choice like the reference)."""
    n = len(u)
    choice = list(set(numpy.random.choice(numpy.arange(n), size=n)))
    for i in set(range(n)) - set(choice):
        offsets = numpy.abs(u[i, :] - u[choice, :])
        if numpy.all(offsets < maxdistance.reshape((1, -1)), axis=1).any():
            continue
        clipped = numpy.where(maxdistance > offsets, offsets, maxdistance)
        cost = numpy.log(clipped).sum(axis=1) - numpy.log(maxdistance).sum()
        towards = numpy.argmin(cost)
        maxdistance = numpy.where(offsets[towards] > maxdistance, offsets[towards], maxdistance)
    return maxdistance


def find_maxdistance(u, nbootstraps=15):
    """clustering/neighbors.py:64-73."""
    maxdistance = initial_maxdistance_guess(u)
    for _ in range(nbootstraps):
        maxdistance = update_maxdistance(u, maxdistance)
    return maxdistance
