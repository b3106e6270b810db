/*
 * mdns_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
 *
 * Plain-C CPU restatement of the massivedatans hot path, used as the parity
 * checker for the CUDA kernels in massivedatans_b200/csrc and as the "port"
 * CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Every function states the reference file:line whose arithmetic it follows
 * (paths relative to the upstream JohannesBuchner/massivedatans tree).  The
 * summation order, the un-fused IEEE-754 double operations and the quirks of
 * the reference are kept so that results are bit-identical to the reference
 * shared objects built by oracle/Makefile into oracle/_ref/ -- this is pinned
 * by tests/test_oracle_vs_reference.py (runs where /root/reference exists) and
 * by the committed fixtures in tests/golden/ (generated from oracle/_ref by
 * tests/golden/make_golden.py).
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no FMA contraction, the
 * reference is compiled for baseline x86-64 which has no FMA either).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* Column `i` of a channel-major matrix: element (channel j, data set i) lives
 * at m[i + j*ndata]  (clike.c:72, cmuselike.c:53). */
#define AT(m, i, j, ndata) ((m)[(size_t)(i) + (size_t)(j) * (size_t)(ndata)])

/*
 * Toy Gaussian-line likelihood.  Follows clike.c:64-76 (serial branch):
 *   ypred_j = A * exp(-0.5 * ((mu - x_j)/sig)^2)              clike.c:65
 *   Lout[k] += ((ypred_j - yy[i + j*ndata]) / noise)^2        clike.c:72
 * k = rank of data set i among the masked-in data sets         clike.c:67-74
 * Lout is accumulated into (caller zeroes it, sample.py:104).
 */
int oracle_clike(const double *x, const double *yy, int ndata, int nx,
                 double A, double mu, double sig, double noise,
                 const uint8_t *mask, double *Lout)
{
	for (int j = 0; j < nx; j++) {
		const double t = (mu - x[j]) / sig;
		const double ypred = A * exp(-0.5 * (t * t));
		int k = 0;
		for (int i = 0; i < ndata; i++) {
			if (!mask[i])
				continue;
			const double r = (ypred - AT(yy, i, j, ndata)) / noise;
			Lout[k] += r * r;
			k++;
		}
	}
	return 0;
}

/*
 * Same chi-square sum for a caller-provided model spectrum ypred[nx]
 * (what clike.c:72 does once clike.c:65 has been evaluated); used to check
 * the batched "K spectra" entry point.
 */
int oracle_clike_spectrum(const double *ypred, const double *yy, int ndata,
                          int nx, double noise, const uint8_t *mask,
                          double *Lout)
{
	for (int j = 0; j < nx; j++) {
		int k = 0;
		for (int i = 0; i < ndata; i++) {
			if (!mask[i])
				continue;
			const double r = (ypred[j] - AT(yy, i, j, ndata)) / noise;
			Lout[k] += r * r;
			k++;
		}
	}
	return 0;
}

/*
 * MUSE scaled chi-square likelihood.  Follows cmuselike.c:48-64:
 *   s1 = sum_j y*m/v ; s2 = 1e-10 + sum_j m^2/v             cmuselike.c:51-56
 *   s = s1/s2                                                cmuselike.c:57
 *   chi = sum_j (y - s*m)^2 / v                              cmuselike.c:58-61
 *   Lout[i] = -0.5*chi  (un-compacted, masked entries only)  cmuselike.c:62
 */
int oracle_cmuselike(const double *yy, const double *vv, const double *ypred,
                     const uint8_t *mask, int ndata, int nx, double *Lout)
{
	for (int i = 0; i < ndata; i++) {
		if (!mask[i])
			continue;
		double s1 = 0.0, s2 = 1e-10;
		for (int j = 0; j < nx; j++) {
			const double v = AT(vv, i, j, ndata);
			s1 += AT(yy, i, j, ndata) * ypred[j] / v;
			s2 += (ypred[j] * ypred[j]) / v;
		}
		const double s = s1 / s2;
		double chi = 0.0;
		for (int j = 0; j < nx; j++) {
			const double r = AT(yy, i, j, ndata) - s * ypred[j];
			chi += (r * r) / AT(vv, i, j, ndata);
		}
		Lout[i] = -0.5 * chi;
	}
	return 0;
}

/* Squared euclidean distance, k-sequential, starting from 0
 * (cneighbors.c:54-57, :84-87, :104-107, :153-156). */
static double sqdist(const double *a, const double *b, int ndim)
{
	double d = 0.0;
	for (int k = 0; k < ndim; k++) {
		const double t = a[k] - b[k];
		d += t * t;
	}
	return d;
}

/*
 * Jackknife radius: max_i sqrt(min_{j != i} |x_i - x_j|^2).
 * Follows cneighbors.c:50-73 (nearest init 1e300 :51, strict '<' :58,
 * max starts from sample 0 :67-72).
 */
double oracle_most_distant_nearest_neighbor(const double *xx, int n, int ndim)
{
	double furthest = 0.0;
	for (int i = 0; i < n; i++) {
		double nearest = 1e300;
		for (int j = 0; j < n; j++) {
			if (j == i)
				continue;
			const double d = sqdist(xx + (size_t)i * ndim, xx + (size_t)j * ndim, ndim);
			if (d < nearest)
				nearest = d;
		}
		const double r = sqrt(nearest);
		if (i == 0 || r > furthest)
			furthest = r;
	}
	return furthest;
}

/* Any member within maxdistance of the single point y?  cneighbors.c:83-91 */
int oracle_is_within_distance_of(const double *xx, int n, int ndim,
                                 double maxdistance, const double *y)
{
	for (int i = 0; i < n; i++)
		if (sqrt(sqdist(xx + (size_t)i * ndim, y, ndim)) < maxdistance)
			return 1;
	return 0;
}

/*
 * Members within maxdistance of each candidate.  Follows cneighbors.c:103-117:
 * out[j] (double, caller-zeroed) is incremented per hit; when countmax > 0 the
 * member scan for candidate j stops as soon as out[j] >= countmax (:112).
 */
int oracle_count_within_distance_of(const double *xx, int n, int ndim,
                                    double maxdistance, const double *yy,
                                    int m, double *out, int countmax)
{
	for (int j = 0; j < m; j++) {
		for (int i = 0; i < n; i++) {
			const double d = sqdist(xx + (size_t)i * ndim, yy + (size_t)j * ndim, ndim);
			if (sqrt(d) < maxdistance) {
				out[j] += 1.0;
				if (countmax > 0 && out[j] >= countmax)
					break;
			}
		}
	}
	return 0;
}

/*
 * Bootstrapped max nearest-neighbour distance.  Follows cneighbors.c:140-176:
 * per round b, every un-chosen sample i gets nearest = min over chosen j of
 * |x_i-x_j|^2 (init 1e300 :148); the round's value is the max of
 * sqrt(nearest) over un-chosen i >= 1 -- sample 0 is skipped (:162) --
 * starting from 0 (:142); the result is the max over rounds (:172-176).
 * chosen is a double 0/1 matrix [n][nboot], round index fastest (:146).
 */
double oracle_bootstrapped_maxdistance(const double *xx, int n, int ndim,
                                       const double *chosen, int nboot)
{
	double best = 0.0;
	for (int b = 0; b < nboot; b++) {
		double furthest = 0.0;
		for (int i = 1; i < n; i++) {
			if (chosen[(size_t)i * nboot + b] != 0)
				continue;
			double nearest = 1e300;
			for (int j = 0; j < n; j++) {
				if (chosen[(size_t)j * nboot + b] == 0)
					continue;
				const double d = sqdist(xx + (size_t)i * ndim, xx + (size_t)j * ndim, ndim);
				if (d < nearest)
					nearest = d;
			}
			const double r = sqrt(nearest);
			if (r > furthest)
				furthest = r;
		}
		if (b == 0 || furthest > best)
			best = furthest;
	}
	return best;
}

/* ---- live-point table (multi_nested_sampler.py) --------------------------- */
#include <stdlib.h>

/*
 * prepare(), multi_nested_sampler.py:134-137, and Lmax, :531, on
 * live_pointsL[nlive][ndata] (C order): per data set the minimum over the live
 * points, the row of its first occurrence (numpy.argmin) and the maximum.
 */
void oracle_live_colstats(const double *L, int nlive, int ndata, double *Lmins,
                          long long *Lmini, double *Lmax)
{
	for (int d = 0; d < ndata; d++) {
		double lo = L[d], hi = L[d];
		long long at = 0;
		for (int i = 1; i < nlive; i++) {
			const double v = L[(size_t)i * ndata + d];
			if (v < lo) {
				lo = v;
				at = i;
			}
			if (v > hi)
				hi = v;
		}
		Lmins[d] = lo;
		Lmini[d] = at;
		Lmax[d] = hi;
	}
}

static int cmp_double(const void *a, const void *b)
{
	const double x = *(const double *)a, y = *(const double *)b;
	return (x > y) - (x < y);
}

/*
 * find_nsmallest, multi_nested_sampler.py:38-42 (the "old version": join the
 * two arrays, sort everything, return element n; the new version :44-47 uses
 * numpy.partition and returns the same element).
 */
double oracle_find_nsmallest(int n, const double *arr1, int n1, const double *arr2, int n2)
{
	double *arr = (double *)malloc(sizeof(double) * (size_t)(n1 + n2));
	for (int i = 0; i < n1; i++)
		arr[i] = arr1[i];
	for (int i = 0; i < n2; i++)
		arr[n1 + i] = arr2[i];
	qsort(arr, (size_t)(n1 + n2), sizeof(double), cmp_double);
	const double r = arr[n];
	free(arr);
	return r;
}

/*
 * Subset partition, multi_nested_sampler.py:204-260 (generate_subsets_nograph;
 * the igraph variant :262-355 yields the same groups): data sets that share live
 * points, transitively.  Restated as union-find over the point -> first holder
 * map; labels[d] = smallest data-set index of d's group (the reference yields the
 * groups in that order, `firstmember` :239), -1 outside the mask.
 */
static int uf_find(int *parent, int a)
{
	while (parent[a] != a) {
		parent[a] = parent[parent[a]];
		a = parent[a];
	}
	return a;
}

void oracle_subsets_labels(const long long *live_pointsp, int nlive, int ndata,
                           const unsigned char *mask, long long npoints, int *labels)
{
	int *parent = (int *)malloc(sizeof(int) * (size_t)ndata);
	int *holder = (int *)malloc(sizeof(int) * (size_t)npoints);
	for (int d = 0; d < ndata; d++)
		parent[d] = d;
	for (long long p = 0; p < npoints; p++)
		holder[p] = -1;
	for (int d = 0; d < ndata; d++) {
		if (mask && !mask[d])
			continue;
		for (int i = 0; i < nlive; i++) {
			const long long p = live_pointsp[(size_t)i * ndata + d];
			if (holder[p] < 0) {
				holder[p] = d;
			} else {
				int a = uf_find(parent, holder[p]), b = uf_find(parent, d);
				if (a < b)
					parent[b] = a;
				else
					parent[a] = b;      /* the smaller index stays the root */
			}
		}
	}
	for (int d = 0; d < ndata; d++)
		labels[d] = (mask && !mask[d]) ? -1 : uf_find(parent, d);
	free(parent);
	free(holder);
}
