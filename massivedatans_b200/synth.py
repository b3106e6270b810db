"""Seeded synthetic inputs restating the reference's data generators.

These are host-side fixtures for tests and bench.py (the data is then made
resident on the device by the shim); nothing here is on the hot path.

* horns      -- gensimple_horns.py:15-39  (x: 200 channels, one narrow line, noise 0.01)
* nothing    -- gennothing.py:7-12        (pure noise)
* realistic  -- gen_realistic.py:16-50    (1000 channels, broad+narrow line); the
                reference always simulates 10 000 spectra and truncates
                (gen_realistic.py:18,52-53); here N is a parameter.
* muse       -- a cube of the reference shape: nspec=3600 (musefuse.py:35),
                ndata=4223 (pres/massivens4.lyx:2229), variance bands inflated by
                1e10 as at musefuse.py:128-132.

`legacy=True` reproduces the reference's numpy.random (RandomState) stream
bit-for-bit for the same N; `legacy=False` uses the faster PCG64 generator for
the 1e6-data-set measurement inputs.  Matrices are returned channel-major
(y[nx, ndata], C-contiguous, data-set index fastest) exactly as sample.py:31
hands them to the native code.
"""
import numpy

NOISE_LEVEL = 0.01            # sample.py:45, gensimple_horns.py:28


def _gauss(x, A, mu, sig):
    # gensimple_horns.py:8-13 -- rows = data sets, columns = channels
    return A.reshape((-1, 1)) * numpy.exp(
        -0.5 * ((mu.reshape((-1, 1)) - x.reshape((1, -1))) / sig.reshape((-1, 1))) ** 2)


def horns(N, nx=200, legacy=True, seed=None):
    """gensimple_horns.py: returns x[nx], y[nx, N], truth dict."""
    x = numpy.linspace(400, 800, nx)
    seed = N if seed is None else seed
    if legacy:
        rs = numpy.random.RandomState(seed)
        z = numpy.arctan(rs.uniform(-numpy.pi, numpy.pi, size=N)) * 0.1
        height = 0.02 / rs.power(3, size=N)
        # gensimple_horns.py:37-38 draws len(x) normals per data set, in data-set order;
        # one (N, nx) draw consumes the RandomState stream identically (nx even).
        noise = rs.normal(0, NOISE_LEVEL, size=(N, nx))
    else:
        rg = numpy.random.default_rng(seed)
        z = numpy.arctan(rg.uniform(-numpy.pi, numpy.pi, size=N)) * 0.1
        height = 0.02 / rg.power(3, size=N)
        noise = rg.standard_normal(size=(N, nx))
        noise *= NOISE_LEVEL
    mean = 656 * (1 + z)
    width = 5.0 * numpy.ones(N)
    step = 65536
    y = numpy.empty((nx, N))
    for lo in range(0, N, step):      # chunked to bound temporaries at N = 1e6
        hi = min(N, lo + step)
        ym = _gauss(x, height[lo:hi], mean[lo:hi], width[lo:hi])
        ym += noise[lo:hi]
        y[:, lo:hi] = ym.T
    return x, y, dict(z=z, mean_narrow=mean, width_narrow=width, height_narrow=height)


def nothing(N, nx=200, legacy=True, seed=None):
    """gennothing.py:7-12."""
    x = numpy.linspace(400, 800, nx)
    seed = N if seed is None else seed
    if legacy:
        y = numpy.random.RandomState(seed).normal(0, NOISE_LEVEL, size=(nx, N))
    else:
        y = numpy.random.default_rng(seed).standard_normal(size=(nx, N))
        y *= NOISE_LEVEL
    return x, y


def realistic(N, nx=1000, seed=1):
    """gen_realistic.py:16-50 with the 10 000 cap lifted (same distributions)."""
    x = numpy.linspace(400, 800, nx)
    rs = numpy.random.RandomState(seed)
    z = rs.beta(2, 30, size=N) * 2
    rest_wave = 440
    width_broad = 10 ** rs.normal(3, 0.2, size=N) * rest_wave / 300000
    width_narrow = 10 ** rs.normal(1, 0.2, size=N) * rest_wave / 300000
    signal_level = 1. / (rs.power(1, size=N) * 100 + 2)
    is_type1 = rs.uniform(size=N) < 0.5
    height_broad = numpy.where(is_type1, 10 ** rs.normal(0, 0.2, size=N),
                               10 ** rs.normal(-2, 0.2, size=N)) * signal_level
    height_narrow = signal_level
    mu = rest_wave * numpy.ones(N)
    y = numpy.empty((nx, N))
    step = 16384
    rg = numpy.random.default_rng(seed)
    for lo in range(0, N, step):
        hi = min(N, lo + step)
        xz = x.reshape((1, -1)) / (1. + z[lo:hi].reshape((-1, 1)))
        ym = height_broad[lo:hi, None] * numpy.exp(
            -0.5 * ((mu[lo:hi, None] - xz) / width_broad[lo:hi, None]) ** 2)
        ym += height_narrow[lo:hi, None] * numpy.exp(
            -0.5 * ((mu[lo:hi, None] - xz) / width_narrow[lo:hi, None]) ** 2)
        ym += rg.standard_normal(size=ym.shape) * NOISE_LEVEL
        y[:, lo:hi] = ym.T
    return x, y, dict(z=z)


def realistic_fast(N, nx=1000, seed=1, threads=8, chunk=8192):
    """The same spectra as `realistic` (same line parameters from the same RandomState) built
    chunk by chunk on several host threads, directly in the channel-major orientation, with one
    counter-based noise stream per chunk: 1e6 x 1000 (8 GB) in seconds instead of minutes.
    Returns (x, y); the result does not depend on the number of threads."""
    from concurrent.futures import ThreadPoolExecutor
    x = numpy.linspace(400, 800, nx)
    rs = numpy.random.RandomState(seed)
    z = rs.beta(2, 30, size=N) * 2
    rest_wave = 440
    width_broad = 10 ** rs.normal(3, 0.2, size=N) * rest_wave / 300000
    width_narrow = 10 ** rs.normal(1, 0.2, size=N) * rest_wave / 300000
    signal_level = 1. / (rs.power(1, size=N) * 100 + 2)
    is_type1 = rs.uniform(size=N) < 0.5
    height_broad = numpy.where(is_type1, 10 ** rs.normal(0, 0.2, size=N),
                               10 ** rs.normal(-2, 0.2, size=N)) * signal_level
    height_narrow = signal_level
    y = numpy.empty((nx, N))
    xc = x.reshape((-1, 1))

    def work(lo):
        hi = min(N, lo + chunk)
        d = rest_wave - xc / (1. + z[lo:hi].reshape((1, -1)))          # mu - x/(1+z), [nx, chunk]
        ym = numpy.exp(-0.5 * (d / width_broad[lo:hi]) ** 2)
        ym *= height_broad[lo:hi]
        t = numpy.exp(-0.5 * (d / width_narrow[lo:hi]) ** 2)
        t *= height_narrow[lo:hi]
        ym += t
        rg = numpy.random.Generator(numpy.random.Philox(key=seed, counter=[0, 0, 0, lo]))
        t = rg.standard_normal(size=ym.shape)
        t *= NOISE_LEVEL
        ym += t
        y[:, lo:hi] = ym

    with ThreadPoolExecutor(max(1, int(threads))) as pool:
        list(pool.map(work, range(0, N, chunk)))
    return x, y


MUSE_NSPEC = 3600
MUSE_NDATA = 4223
MUSE_BANDS = ((1600, 1670), (1730, 1780), (1950, 2000), (2250, 2700), (2800, 3000))


def muse_template(nspec=MUSE_NSPEC, phase=0.0):
    """Non-negative smooth model spectrum of the cube's length (stand-in for
    musefuse.py:222-284, which needs external BC03 grids)."""
    t = numpy.linspace(0, 1, nspec)
    return (1.0 + 0.5 * numpy.sin(7 * t + phase) + 0.3 * numpy.exp(-0.5 * ((t - 0.4) / 0.01) ** 2)
            + 0.8 * t)


def muse(ndata=MUSE_NDATA, nspec=MUSE_NSPEC, seed=3):
    """MUSE-shaped cube: y[nspec, ndata], variance v[nspec, ndata] (>0, bands += 1e10)."""
    rg = numpy.random.default_rng(seed)
    v = rg.uniform(0.5, 2.0, size=(nspec, ndata))
    template = muse_template(nspec)
    scale = rg.uniform(0.5, 20.0, size=ndata)
    y = template.reshape((-1, 1)) * scale.reshape((1, -1))
    y += rg.standard_normal(size=y.shape) * numpy.sqrt(v)
    for lo, hi in MUSE_BANDS:
        if lo < nspec:
            v[lo:min(hi, nspec), :] += 1e10
    return y, v, template


MUSE_ZS = numpy.log10([0.0001, 0.0004, 0.004, 0.008, 0.02, 0.05, 0.1])    # musefuse.py:188


def muse_grids(nZ=7, nages=111, nwave=2400, seed=4):
    """Synthetic stand-in for the stellar-population template grids musefuse.py:176-185 reads
    from external BC03 text files (absent from the reference tree): per metallicity one
    non-negative array [nages, nwave], a continuum that reddens with age plus absorption
    features.  Returns (Zs[nZ], ages[nages] in yr (0 first, increasing), model_wavelength[nwave]
    in Angstrom as in the files (musefuse.py:181, divided by 10 at :207), grids[nZ, nages, nwave]).
    """
    rg = numpy.random.default_rng(seed)
    ages = numpy.concatenate(([0.0], numpy.logspace(5, numpy.log10(2e10), nages - 1)))
    wl = numpy.linspace(2000.0, 11000.0, nwave)
    t = (wl - wl[0]) / (wl[-1] - wl[0])
    grids = numpy.empty((nZ, nages, nwave))
    lines = rg.uniform(0.05, 0.95, size=12)
    for iZ in range(nZ):
        for a in range(nages):
            red = a / float(nages)
            cont = (1.2 - red) * numpy.exp(-3.0 * t * (1.0 - red)) + red * t ** (0.5 + 0.1 * iZ)
            absorb = 1.0
            for c in lines:
                absorb = absorb - (0.1 + 0.05 * iZ) * red * numpy.exp(-0.5 * ((t - c) / 0.004) ** 2)
            grids[iZ, a] = cont * absorb * (1.0 + 0.02 * rg.standard_normal(nwave)).clip(0.5) \
                * 10.0 ** (-3.0 * red)
    return MUSE_ZS[:nZ].copy(), ages, wl, grids


def muse_wavelength(nspec=MUSE_NSPEC):
    """Data wavelength grid in Angstrom (MUSE: 4750-9350 A; musefuse.py:82)."""
    return 4750.0 + (9350.0 - 4750.0) / nspec * numpy.arange(nspec)


def muse_parameter_points(K, seed=8, nZ=7):
    """(Z, logSFtau, SFage, z, EBV) inside the prior of musefuse.py:329-346."""
    rg = numpy.random.default_rng(seed)
    p = numpy.empty((K, 5))
    p[:, 0] = rg.uniform(MUSE_ZS[0], MUSE_ZS[nZ - 1] + 0.3, size=K)
    p[:, 1] = rg.uniform(6.0, numpy.log10(4e9), size=K)
    p[:, 2] = rg.uniform(0.05, 13.0, size=K)
    p[:, 3] = rg.uniform(0.0, 0.9, size=K)
    p[:, 4] = rg.uniform(0.0, 2.0, size=K)
    return p


def priortransform(cube):
    """sample.py:52-58 -- unit cube -> (A, mu, log_sig)."""
    cube = numpy.array(cube, dtype=float, copy=True)
    cube[..., 0] = 10 ** (cube[..., 0] * 2 - 2)
    cube[..., 1] = cube[..., 1] * 400 + 400
    cube[..., 2] = cube[..., 2] * 2
    return cube


def parameter_points(K, seed=7):
    """K prior draws -> rows (A, mu, sig) with sig = 10**log_sig (sample.py:102-103)."""
    u = numpy.random.RandomState(seed).uniform(size=(K, 3))
    p = priortransform(u)
    p[:, 2] = 10 ** p[:, 2]
    return p


def masks(N, seed=11):
    """The mask shapes the sampler produces (multi_nested_sampler.py:373-388,148-162)."""
    rs = numpy.random.RandomState(seed)
    return {
        'all': numpy.ones(N, dtype=bool),
        'half': rs.uniform(size=N) < 0.5,
        'sparse': rs.uniform(size=N) < 0.01,
        'prefix': numpy.arange(N) < max(1, N // 3),
        'none': numpy.zeros(N, dtype=bool),
    }


def members_and_candidates(n, m, ndim, seed=5):
    """Live-point union in a unit box and uniform candidates (radfriendsregion.py:139)."""
    rs = numpy.random.RandomState(seed)
    return rs.uniform(size=(n, ndim)), rs.uniform(size=(m, ndim))


def bootstrap_chosen(n, nbootstraps, rs=numpy.random):
    """The 0/1 float64 matrix drawn at clustering/neighbors.py:172-174."""
    chosen = numpy.zeros((n, nbootstraps))
    for b in range(nbootstraps):
        chosen[rs.choice(numpy.arange(n), size=n, replace=True), b] = 1.
    return chosen
