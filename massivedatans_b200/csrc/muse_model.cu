// muse_model.cu -- the MUSE stellar-population model spectrum on the device.
//
// Reference: musefuse.py:222-284 `model(Z, SFtau, sfage, z, EBV)` (numpy on the host, ~ms per
// call; the cmuselike pass it feeds takes 0.044 ms on this GPU, so a host model would be the
// bottleneck of every MUSE likelihood call).  For a batch of K parameter points:
//   1. metallicity bin iZ = last node with Zs <= Z                        (:223, on the host)
//   2. star-formation history over the template ages: t = max(sfage*1e9 - age, 0),
//      sfh = t/tau^2 * exp(-t/tau), normalised to its maximum                   (:233-239)
//   3. template[w] = sum over ages a < nages-1 of grid[iZ][a][w] * sfh[a] * (age[a+1]-age[a])
//      in age order (numpy.sum over the leading axis adds the rows one after the other) (:250-251)
//   4. template /= 1e-10 + template[norm_index]                                      (:255)
//   5. template *= 10**(-2.5 * calzetti * EBV)                                      (:258)
//   6. linear interpolation onto wavelength/(1+z), clamped at the ends (numpy.interp) (:279)
// Three launches per batch: weights (2), template sum (3), and 4-6 fused in the resampling kernel.
// and the K spectra land in the data set's staged model buffer (what mdns_stage_spectra fills
// from the host), so mdns_muse_launch follows without any spectrum crossing PCIe.
//
// All arithmetic FP64 in the reference's operation order (no FMA contraction: -fmad=false);
// differences to numpy come only from exp/pow (device libm vs glibc, <= 2 ulp).
//
// Memory: the grids (nZ x nages x nwave doubles, 43 MB for the BC03 high-resolution files) stay
// resident on every device of the data set and are L2-resident after the first batch; step 3 is
// the only pass that reads them: K x (nages-1) x nwave x 8 bytes from L2.
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

using namespace mdns;

extern "C" int mdns_internal_shard_view(mdns_dataset *ds, int shard, int *device, int *i0, int *n,
                                        int *n_act, int *K, const double **d_out, void **stream);
extern "C" int mdns_internal_shard_count(const mdns_dataset *ds);
extern "C" int mdns_internal_model_buffer(mdns_dataset *ds, int shard, int K, double **d_model,
                                          long long *pitch, int *nx);
extern "C" int mdns_internal_spectra_staged(mdns_dataset *ds, int K);

namespace {

constexpr int MM_THREADS = 128;

// params[k] = (Z, SFtau, sfage, z, EBV, metallicity bin); sfh[k][a] for a < nages (normalised to
// the maximum); also clears the candidate's "spectrum is not all zero" flag
constexpr int MM_P = 6;
__global__ void __launch_bounds__(MM_THREADS) muse_sfh_kernel(const double *__restrict__ params,
                                                              const double *__restrict__ ages, int nages,
                                                              double *__restrict__ sfh,
                                                              int *__restrict__ nonzero)
{
	__shared__ double red[MM_THREADS / 32];
	const int k = blockIdx.x;
	if (threadIdx.x == 0) nonzero[k] = 0;
	const double tau = params[k * MM_P + 1], sfage = params[k * MM_P + 2];
	const double tau2 = tau * tau;
	const double start = sfage * 1.e9;
	double best = -1.0;       // sfh >= 0; NaN entries are tracked separately (numpy's max returns NaN)
	bool any_nan = false;
	for (int a = threadIdx.x; a < nages; a += MM_THREADS) {
		double t = start - ages[a];
		if (t <= 0) t = 0;
		const double v = t / tau2 * exp(-t / tau);
		sfh[(size_t)k * nages + a] = v;
		if (v != v) any_nan = true;
		best = v > best ? v : best;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const double other = __shfl_xor_sync(0xffffffffu, best, o);
		best = other > best ? other : best;
	}
	any_nan = __syncthreads_or(any_nan);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
	__syncthreads();
	double m = red[0];
#pragma unroll
	for (int w = 1; w < MM_THREADS / 32; ++w) m = red[w] > m ? red[w] : m;
	if (any_nan) m = nan("");      // numpy's max propagates NaN
	for (int a = threadIdx.x; a < nages; a += MM_THREADS) sfh[(size_t)k * nages + a] /= m;
}

// tmpl[k][w] = sum_a (grid[iZ_k][a][w] * sfh[k][a]) * dage[a], a = 0 .. nages-2, in order
__global__ void __launch_bounds__(MM_THREADS) muse_template_kernel(
    const double *__restrict__ grids, const double *__restrict__ params, const double *__restrict__ sfh,
    const double *__restrict__ dage, int nages, int nwave, double *__restrict__ tmpl)
{
	extern __shared__ double sm[];        // sfh[k][0..nages-2], dage[0..nages-2]
	const int k = blockIdx.y;
	const int na = nages - 1;
	for (int a = threadIdx.x; a < na; a += MM_THREADS) {
		sm[a] = sfh[(size_t)k * nages + a];
		sm[na + a] = dage[a];
	}
	__syncthreads();
	const int w = blockIdx.x * MM_THREADS + threadIdx.x;
	if (w >= nwave) return;
	const double *g = grids + (size_t)(int)params[k * MM_P + 5] * nages * nwave + w;
	double acc = 0.0;
	// four rows in flight per iteration; the additions stay in age order
	int a = 0;
	for (; a + 4 <= na; a += 4) {
		const double g0 = g[(size_t)a * nwave], g1 = g[(size_t)(a + 1) * nwave];
		const double g2 = g[(size_t)(a + 2) * nwave], g3 = g[(size_t)(a + 3) * nwave];
		acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(g0, sm[a]), sm[na + a]));
		acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(g1, sm[a + 1]), sm[na + a + 1]));
		acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(g2, sm[a + 2]), sm[na + a + 2]));
		acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(g3, sm[a + 3]), sm[na + a + 3]));
	}
	for (; a < na; ++a)
		acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(g[(size_t)a * nwave], sm[a]), sm[na + a]));
	tmpl[(size_t)k * nwave + w] = acc;
}

// steps 4-6: normalisation, extinction and numpy.interp(x = wavelength/(1+z), xp = model_wavelength,
// fp = template) -> model[k][c].  The normalised, extincted template is only needed at the two
// nodes around every output channel, so it is evaluated there (two pow per output channel, same
// operations as a pass over the whole template: v/norm * 10**((-2.5*calz)*EBV)).
__device__ __forceinline__ double muse_node(const double *__restrict__ raw, const double *__restrict__ calz,
                                            int j, double norm, double ebv)
{
	return (raw[j] / norm) * pow(10.0, __dmul_rn(__dmul_rn(-2.5, calz[j]), ebv));
}

__global__ void __launch_bounds__(MM_THREADS) muse_resample_kernel(
    const double *__restrict__ params, const double *__restrict__ wavelength, int nx,
    const double *__restrict__ xp, const double *__restrict__ calz, int nwave, int norm_index,
    const double *__restrict__ tmpl, double *__restrict__ model, long long mpitch,
    int *__restrict__ nonzero)
{
	const int k = blockIdx.y;
	const int c = blockIdx.x * MM_THREADS + threadIdx.x;
	if (c >= nx) return;
	const double x = wavelength[c] / (1 + params[k * MM_P + 3]);
	const double ebv = params[k * MM_P + 4];
	const double *raw = tmpl + (size_t)k * nwave;
	const double norm = 1e-10 + raw[norm_index];
	double r;
	if (x > xp[nwave - 1]) {
		r = muse_node(raw, calz, nwave - 1, norm, ebv);
	} else if (x < xp[0]) {
		r = muse_node(raw, calz, 0, norm, ebv);
	} else if (x != x) {
		r = x;
	} else {
		// largest j with xp[j] <= x
		int lo = 0, hi = nwave - 1;
		while (hi - lo > 1) {
			const int mid = (lo + hi) >> 1;
			if (xp[mid] <= x) lo = mid; else hi = mid;
		}
		int j = lo;
		if (xp[hi] <= x) j = hi;
		const double fj = muse_node(raw, calz, j, norm, ebv);
		if (j == nwave - 1 || xp[j] == x) {
			r = fj;
		} else {
			const double fj1 = muse_node(raw, calz, j + 1, norm, ebv);
			const double slope = (fj1 - fj) / (xp[j + 1] - xp[j]);
			r = __dadd_rn(__dmul_rn(slope, x - xp[j]), fj);
			if (r != r) {
				// numpy retries from the right node, then gives the common value
				r = __dadd_rn(__dmul_rn(slope, x - xp[j + 1]), fj1);
				if (r != r && fj == fj1) r = fj;
			}
		}
	}
	model[(size_t)k * mpitch + c] = r;
	if (r != 0.0) nonzero[k] = 1;      // numpy.any(ypred), musefuse.py:528 (NaN counts as true)
}

struct MmDevice {
	int device = 0;
	double *grids = nullptr, *ages = nullptr, *dage = nullptr, *xp = nullptr, *calz = nullptr,
	       *wavelength = nullptr;
	double *params = nullptr, *sfh = nullptr, *tmpl = nullptr;
	int *nonzero = nullptr;
	int cap_K = 0;
};

}  // namespace

struct mdns_muse_model {
	mdns_dataset *ds = nullptr;
	int nZ = 0, nages = 0, nwave = 0, nx = 0, norm_index = 0;
	std::vector<double> Zs;
	std::vector<MmDevice> devs;      // one per shard of the data set
	int K = 0;
};

static void mm_free(MmDevice &d)
{
	cudaSetDevice(d.device);
	cudaFree(d.grids);
	cudaFree(d.ages);
	cudaFree(d.dage);
	cudaFree(d.xp);
	cudaFree(d.calz);
	cudaFree(d.wavelength);
	cudaFree(d.params);
	cudaFree(d.sfh);
	cudaFree(d.tmpl);
	cudaFree(d.nonzero);
}

template <typename T>
static int mm_upload(T **dst, const T *src, size_t n)
{
	MDNS_CUDA(cudaMalloc((void **)dst, n * sizeof(T)));
	MDNS_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
	return MDNS_OK;
}

static int mm_reserve(mdns_muse_model *m, MmDevice &d, int K)
{
	if (K <= d.cap_K) return MDNS_OK;
	cudaFree(d.params);
	cudaFree(d.sfh);
	cudaFree(d.tmpl);
	cudaFree(d.nonzero);
	d.params = d.sfh = d.tmpl = nullptr;
	d.nonzero = nullptr;
	d.cap_K = 0;
	const int cap = K + K / 2 + 4;
	MDNS_CUDA(cudaMalloc((void **)&d.params, (size_t)cap * MM_P * sizeof(double)));
	MDNS_CUDA(cudaMalloc((void **)&d.sfh, (size_t)cap * m->nages * sizeof(double)));
	MDNS_CUDA(cudaMalloc((void **)&d.tmpl, (size_t)cap * m->nwave * sizeof(double)));
	MDNS_CUDA(cudaMalloc((void **)&d.nonzero, (size_t)cap * sizeof(int)));
	d.cap_K = cap;
	return MDNS_OK;
}

extern "C" {

int mdns_muse_model_create(mdns_dataset *ds, const double *grids, int nZ, const double *Zs, int nages,
                           const double *ages, int nwave, const double *model_wavelength,
                           const double *calzetti, const double *wavelength, int nx, int norm_index,
                           mdns_muse_model **out)
{
	if (!ds || !grids || !Zs || !ages || !model_wavelength || !calzetti || !wavelength || !out ||
	    nZ <= 0 || nages < 2 || nwave < 2 || nx <= 0 || norm_index < 0 || norm_index >= nwave) {
		set_error("mdns_muse_model_create: need ds, grids[nZ][nages][nwave], Zs, ages (>= 2), "
		          "model_wavelength (>= 2), calzetti, wavelength[nx], 0 <= norm_index < nwave");
		return MDNS_EINVAL;
	}
	for (int w = 1; w < nwave; ++w)
		if (!(model_wavelength[w] > model_wavelength[w - 1])) {
			set_error("model_wavelength must increase strictly (numpy.interp's xp), entry %d does not", w);
			return MDNS_EINVAL;
		}
	double *d_model = nullptr;
	int ds_nx = 0;
	int rc = mdns_internal_model_buffer(ds, 0, 1, &d_model, nullptr, &ds_nx);
	if (rc != MDNS_OK) return rc;
	if (ds_nx != nx) {
		set_error("wavelength grid has %d channels, the data set %d", nx, ds_nx);
		return MDNS_EINVAL;
	}
	mdns_muse_model *m = new mdns_muse_model;
	m->ds = ds;
	m->nZ = nZ;
	m->nages = nages;
	m->nwave = nwave;
	m->nx = nx;
	m->norm_index = norm_index;
	m->Zs.assign(Zs, Zs + nZ);
	std::vector<double> dage(nages - 1);
	for (int a = 0; a + 1 < nages; ++a) dage[a] = ages[a + 1] - ages[a];
	const int nshards = mdns_internal_shard_count(ds);
	m->devs.resize(nshards);
	for (int s = 0; s < nshards; ++s) {
		MmDevice &d = m->devs[s];
		mdns_internal_shard_view(ds, s, &d.device, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
		cudaError_t e = cudaSetDevice(d.device);
		rc = e == cudaSuccess ? MDNS_OK : MDNS_ECUDA;
		if (rc == MDNS_OK) rc = mm_upload(&d.grids, grids, (size_t)nZ * nages * nwave);
		if (rc == MDNS_OK) rc = mm_upload(&d.ages, ages, (size_t)nages);
		if (rc == MDNS_OK) rc = mm_upload(&d.dage, dage.data(), dage.size());
		if (rc == MDNS_OK) rc = mm_upload(&d.xp, model_wavelength, (size_t)nwave);
		if (rc == MDNS_OK) rc = mm_upload(&d.calz, calzetti, (size_t)nwave);
		if (rc == MDNS_OK) rc = mm_upload(&d.wavelength, wavelength, (size_t)nx);
		if (rc != MDNS_OK) {
			for (auto &dd : m->devs) mm_free(dd);
			delete m;
			return rc;
		}
	}
	*out = m;
	return MDNS_OK;
}

int mdns_muse_model_destroy(mdns_muse_model *m)
{
	if (!m) return MDNS_OK;
	for (auto &d : m->devs) mm_free(d);
	delete m;
	return MDNS_OK;
}

int mdns_muse_model_stage(mdns_muse_model *m, const double *params, int K, int *nonzero)
{
	if (!m || !params || K <= 0) {
		set_error("mdns_muse_model_stage: need model, params[K][5], K > 0");
		return MDNS_EINVAL;
	}
	// one upload per shard: the five model arguments and the metallicity bin of every point
	std::vector<double> packed((size_t)K * MM_P);
	for (int k = 0; k < K; ++k) {
		const double Z = params[k * 5];
		int j = -1;
		for (int i = 0; i < m->nZ; ++i)
			if (m->Zs[i] <= Z) j = i;
		if (j < 0) {
			// the reference raises IndexError here (numpy.where(Zs <= Z)[-1][-1], musefuse.py:223)
			set_error("parameter point %d: metallicity %g lies below the first grid node %g", k, Z,
			          m->Zs[0]);
			return MDNS_EINVAL;
		}
		for (int q = 0; q < 5; ++q) packed[(size_t)k * MM_P + q] = params[k * 5 + q];
		packed[(size_t)k * MM_P + 5] = (double)j;
	}
	const size_t smem = (size_t)2 * (m->nages - 1) * sizeof(double);
	if (smem > 48 * 1024) {
		set_error("%d template ages do not fit the shared-memory table", m->nages);
		return MDNS_EINVAL;
	}
	for (size_t s = 0; s < m->devs.size(); ++s) {
		MmDevice &d = m->devs[s];
		double *d_model = nullptr;
		long long pitch = 0;
		int rc = mdns_internal_model_buffer(m->ds, (int)s, K, &d_model, &pitch, nullptr);
		if (rc != MDNS_OK) return rc;
		void *vst = nullptr;
		mdns_internal_shard_view(m->ds, (int)s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &vst);
		cudaStream_t st = (cudaStream_t)vst;
		rc = mm_reserve(m, d, K);
		if (rc != MDNS_OK) return rc;
		MDNS_CUDA(cudaMemcpyAsync(d.params, packed.data(), packed.size() * sizeof(double),
		                          cudaMemcpyHostToDevice, st));
		muse_sfh_kernel<<<K, MM_THREADS, 0, st>>>(d.params, d.ages, m->nages, d.sfh, d.nonzero);
		MDNS_LAUNCHED_HELPER("muse_sfh_kernel");
		muse_template_kernel<<<dim3(ceil_div(m->nwave, MM_THREADS), K), MM_THREADS, smem, st>>>(
		    d.grids, d.params, d.sfh, d.dage, m->nages, m->nwave, d.tmpl);
		MDNS_LAUNCHED("muse_template_kernel");
		muse_resample_kernel<<<dim3(ceil_div(m->nx, MM_THREADS), K), MM_THREADS, 0, st>>>(
		    d.params, d.wavelength, m->nx, d.xp, d.calz, m->nwave, m->norm_index, d.tmpl, d_model,
		    pitch, d.nonzero);
		MDNS_LAUNCHED_HELPER("muse_resample_kernel");
		if (s == 0 && nonzero) {
			MDNS_CUDA(cudaMemcpyAsync(nonzero, d.nonzero, (size_t)K * sizeof(int),
			                          cudaMemcpyDeviceToHost, st));
			MDNS_CUDA(cudaStreamSynchronize(st));
		}
	}
	m->K = K;
	return mdns_internal_spectra_staged(m->ds, K);
}

int mdns_muse_model_spectra(mdns_muse_model *m, double *ypred_out)
{
	if (!m || !ypred_out || m->K <= 0) {
		set_error("mdns_muse_model_spectra: stage a batch first");
		return MDNS_ESTATE;
	}
	double *d_model = nullptr;
	long long pitch = 0;
	int rc = mdns_internal_model_buffer(m->ds, 0, m->K, &d_model, &pitch, nullptr);
	if (rc != MDNS_OK) return rc;
	void *vst = nullptr;
	mdns_internal_shard_view(m->ds, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &vst);
	MDNS_CUDA(cudaMemcpy2DAsync(ypred_out, (size_t)m->nx * sizeof(double), d_model,
	                            (size_t)pitch * sizeof(double), (size_t)m->nx * sizeof(double), m->K,
	                            cudaMemcpyDeviceToHost, (cudaStream_t)vst));
	MDNS_CUDA(cudaStreamSynchronize((cudaStream_t)vst));
	return MDNS_OK;
}

}  // extern "C"
