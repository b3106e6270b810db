// livetable.cu -- the sampler's live-point likelihood table resident next to the data
// (the accept / shelf epilogue of the hot path).
//
// Reference: `live_pointsL[nlive_points, ndata]` of multi_nested_sampler.py:111 (3.2 GB at
// nlive = 400, ndata = 1e6, kept in host RAM and walked with numpy every iteration):
//   * prepare()  multi_nested_sampler.py:134-137  Lmins = min(axis 0), Lmini = argmin(axis 0)
//                and :531 Lmax = max(axis 0);
//   * Lmins_higher  :438-447 via find_nsmallest :44-47: the element of rank n of
//                live_pointsL[:, d] joined with the n likelihoods on data set d's shelf;
//   * advance    :520-524 live_pointsL[Lmini[d], d] = Lj (one replacement per data set);
//   * initial population :91-111: nlive full-mask likelihood calls fill the table.
//
// Layout: row-major like the numpy array (row = live point, data set index fastest), sharded
// over the same devices and data-set ranges as the data set it was created from, so the
// batched likelihood can write its result rows straight into the table.  Every operation is a
// selection (comparisons only): results are bit-identical to numpy's.
#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace mdns {

// ---- kernels -----------------------------------------------------------------------------
// one thread per data set; rows are walked with UNROLL loads in flight (coalesced across d)
__global__ void __launch_bounds__(256) lt_colstats_kernel(const double *__restrict__ T, int nlive,
                                                          int n, double *__restrict__ cmin,
                                                          long long *__restrict__ cargmin,
                                                          double *__restrict__ cmax)
{
	const int d = blockIdx.x * 256 + threadIdx.x;
	if (d >= n) return;
	double lo = T[d], hi = T[d];
	int at = 0;
	int i = 1;
	// (eight rows in flight per thread: with one thread per data set a 2e5-column table has
	// 2e5 threads on the machine, and four loads each left HBM half idle)
	for (; i + 8 <= nlive; i += 8) {
		double v[8];
#pragma unroll
		for (int u = 0; u < 8; ++u) v[u] = __ldcs(T + (size_t)(i + u) * n + d);
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			if (v[u] < lo) {        // strict: first occurrence wins, like numpy.argmin
				lo = v[u];
				at = i + u;
			}
			if (v[u] > hi) hi = v[u];
		}
	}
	for (; i < nlive; ++i) {
		const double v = T[(size_t)i * n + d];
		if (v < lo) {
			lo = v;
			at = i;
		}
		if (v > hi) hi = v;
	}
	if (cmin) cmin[d] = lo;
	if (cargmin) cargmin[d] = at;
	if (cmax) cmax[d] = hi;
}

__global__ void __launch_bounds__(256) lt_replace_kernel(double *__restrict__ T, int nlive, int n,
                                                         const long long *__restrict__ rows,
                                                         const double *__restrict__ values)
{
	const int d = blockIdx.x * 256 + threadIdx.x;
	if (d >= n) return;
	const long long r = rows[d];
	if (r >= 0 && r < nlive) T[(size_t)r * n + d] = values[d];
}

// find_nsmallest (multi_nested_sampler.py:44-47) for a list of data sets: one thread each.
// Single pass with a sorted buffer of the m = n_shelf + 1 smallest values seen so far while
// m <= NB; longer shelves take the threshold walk below (exact, duplicates included).
template <int NB>
__global__ void __launch_bounds__(128) lt_rank_kernel(const double *__restrict__ T, int nlive, int n,
                                                      const int *__restrict__ cols, int ncols,
                                                      const long long *__restrict__ off,
                                                      const double *__restrict__ shelf,
                                                      double *__restrict__ out)
{
	const int j = blockIdx.x * 128 + threadIdx.x;
	if (j >= ncols) return;
	const int d = cols[j];
	const long long s0 = off[j], s1 = off[j + 1];
	const int ns = (int)(s1 - s0);
	const int m = ns + 1;
	const int total = nlive + ns;
	auto value = [&](int i) { return i < nlive ? T[(size_t)i * n + d] : shelf[s0 + (i - nlive)]; };
	if (m <= NB) {
		double buf[NB];
		int cnt = 0;
		for (int i = 0; i < total; ++i) {
			const double v = value(i);
			if (cnt < m) {
				int p = cnt++;
				while (p > 0 && buf[p - 1] > v) {
					buf[p] = buf[p - 1];
					--p;
				}
				buf[p] = v;
			} else if (v < buf[m - 1]) {
				int p = m - 1;
				while (p > 0 && buf[p - 1] > v) {
					buf[p] = buf[p - 1];
					--p;
				}
				buf[p] = v;
			}
		}
		out[j] = buf[m - 1];
		return;
	}
	// rank walk: repeatedly take the smallest value above the current one
	double cur = 0.0;
	bool have = false;
	int remaining = ns;       // rank still to skip
	for (;;) {
		double best = 0.0;
		int count = 0;
		for (int i = 0; i < total; ++i) {
			const double v = value(i);
			if (have && !(v > cur)) continue;
			if (count == 0 || v < best) {
				best = v;
				count = 1;
			} else if (v == best) {
				++count;
			}
		}
		if (count == 0 || remaining < count) {
			out[j] = count ? best : cur;
			return;
		}
		remaining -= count;
		cur = best;
		have = true;
	}
}

// ---- subset partition: connected components of the data set <-> live point graph ----------
// generate_subsets_graph / generate_subsets_nograph (multi_nested_sampler.py:204-355): two
// data sets belong to the same group when they share a live point (directly or through a chain
// of data sets).  Label propagation with the smallest data-set index as the label: every
// round each live point learns the smallest label among the data sets that hold it
// (atomicMin), every data set takes the smallest label among its points, and pointer jumping
// flattens label chains; rounds repeat until nothing changes.  The final label of a group is
// its smallest member index, which is also the order in which the reference yields the groups
// (`firstmember`, :239).
__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t *__restrict__ mask, int n,
                                                      int *__restrict__ label)
{
	const int d = blockIdx.x * 256 + threadIdx.x;
	if (d < n) label[d] = (!mask || mask[d]) ? d : -1;
}

__global__ void __launch_bounds__(256) cc_push_kernel(const int *__restrict__ P, int nlive, int n,
                                                      const int *__restrict__ label,
                                                      int *__restrict__ pointmin)
{
	const int d = blockIdx.x * 256 + threadIdx.x;
	if (d >= n) return;
	const int l = label[d];
	if (l < 0) return;
	for (int i = 0; i < nlive; ++i) {
		int *slot = pointmin + P[(size_t)i * n + d];
		if (*slot > l) atomicMin(slot, l);     // plain read first: most points are settled
	}
}

__global__ void __launch_bounds__(256) cc_pull_kernel(const int *__restrict__ P, int nlive, int n,
                                                      int *__restrict__ label,
                                                      const int *__restrict__ pointmin,
                                                      int *__restrict__ changed)
{
	const int d = blockIdx.x * 256 + threadIdx.x;
	if (d >= n) return;
	const int l = label[d];
	if (l < 0) return;
	int m = l;
	for (int i = 0; i < nlive; ++i) m = min(m, pointmin[P[(size_t)i * n + d]]);
	if (m < l) {
		label[d] = m;
		*changed = 1;
	}
}

__global__ void __launch_bounds__(256) cc_jump_kernel(int n, int *__restrict__ label)
{
	const int d = blockIdx.x * 256 + threadIdx.x;
	if (d >= n) return;
	int l = label[d];
	if (l < 0) return;
	// labels only ever decrease and label[x] <= x, so the chain ends at a root
	int r = label[l];
	while (r != l) {
		l = r;
		r = label[l];
	}
	label[d] = l;
}

__global__ void __launch_bounds__(256) lt_replace_points_kernel(int *__restrict__ P, int nlive, int n,
                                                                const long long *__restrict__ rows,
                                                                const long long *__restrict__ ids)
{
	const int d = blockIdx.x * 256 + threadIdx.x;
	if (d >= n) return;
	const long long r = rows[d];
	if (r >= 0 && r < nlive) P[(size_t)r * n + d] = (int)ids[d];
}

__global__ void __launch_bounds__(256) lt_narrow_points_kernel(const long long *__restrict__ in,
                                                               size_t count, int *__restrict__ out,
                                                               int *__restrict__ bad)
{
	const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= count) return;
	const long long v = in[i];
	if (v < 0 || v > 0x7ffffffeLL) *bad = 1;
	out[i] = (int)v;
}

}  // namespace mdns

using namespace mdns;

// internal view of a data set's shards (capi.cu)
extern "C" int mdns_internal_shard_view(mdns_dataset *ds, int shard, int *device, int *i0, int *n,
                                        int *n_act, int *K, const double **d_out, void **stream);
extern "C" int mdns_internal_shard_count(const mdns_dataset *ds);
extern "C" int mdns_internal_threshold_buffer(mdns_dataset *ds, int shard, double **d_lmins);

struct LtShard {
	int device = 0, i0 = 0, n = 0;
	cudaStream_t stream = nullptr;
	double *T = nullptr;
	double *d_min = nullptr, *d_max = nullptr, *d_vals = nullptr;
	long long *d_arg = nullptr, *d_rows = nullptr;
	int *d_cols = nullptr;
	long long *d_off = nullptr;
	double *d_shelf = nullptr, *d_rank = nullptr;
	size_t cols_cap = 0, off_cap = 0, rank_cap = 0, shelf_cap = 0;
};

struct mdns_livetable {
	int nlive = 0, ndata = 0;
	std::vector<LtShard> shards;
	// live_pointsp[nlive][ndata] (multi_nested_sampler.py:111) as int32, whole on the first
	// shard's device: the groups of the subset partition span the shards
	int *P = nullptr;
	int *d_label = nullptr, *d_pointmin = nullptr, *d_flag = nullptr;
	uint8_t *d_cmask = nullptr;
	long long *d_prow = nullptr, *d_pid = nullptr;
	size_t pointmin_cap = 0;
};

static void lt_free(mdns_livetable *t)
{
	if (!t->shards.empty()) {
		cudaSetDevice(t->shards[0].device);
		if (t->shards[0].stream) cudaStreamSynchronize(t->shards[0].stream);
		cudaFree(t->P);
		cudaFree(t->d_label);
		cudaFree(t->d_pointmin);
		cudaFree(t->d_flag);
		cudaFree(t->d_cmask);
		cudaFree(t->d_prow);
		cudaFree(t->d_pid);
	}
	for (auto &s : t->shards) {
		cudaSetDevice(s.device);
		if (s.stream) cudaStreamSynchronize(s.stream);
		cudaFree(s.T);
		cudaFree(s.d_min);
		cudaFree(s.d_max);
		cudaFree(s.d_vals);
		cudaFree(s.d_arg);
		cudaFree(s.d_rows);
		cudaFree(s.d_cols);
		cudaFree(s.d_off);
		cudaFree(s.d_shelf);
		cudaFree(s.d_rank);
		if (s.stream) cudaStreamDestroy(s.stream);
	}
	delete t;
}

template <typename T>
static int lt_grow(T **p, size_t *cap, size_t want)
{
	if (want <= *cap) return MDNS_OK;
	if (*p) MDNS_CUDA(cudaFree(*p));
	*p = nullptr;
	*cap = 0;
	const size_t n = want + want / 2 + 16;
	MDNS_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
	*cap = n;
	return MDNS_OK;
}

static int lt_sync(mdns_livetable *t)
{
	for (auto &s : t->shards) {
		MDNS_CUDA(cudaSetDevice(s.device));
		MDNS_CUDA(cudaStreamSynchronize(s.stream));
	}
	return MDNS_OK;
}

extern "C" {

int mdns_livetable_create(mdns_dataset *ds, int nlive, mdns_livetable **out)
{
	if (!ds || !out || nlive <= 0) {
		set_error("mdns_livetable_create: need a data set, out and nlive > 0");
		return MDNS_EINVAL;
	}
	mdns_livetable *t = new mdns_livetable();
	t->nlive = nlive;
	const int nshards = mdns_internal_shard_count(ds);
	t->shards.resize(nshards);
	for (int k = 0; k < nshards; ++k) {
		LtShard &s = t->shards[k];
		mdns_internal_shard_view(ds, k, &s.device, &s.i0, &s.n, nullptr, nullptr, nullptr, nullptr);
		t->ndata += s.n;
		cudaError_t e = cudaSetDevice(s.device);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaMalloc((void **)&s.T, (size_t)nlive * s.n * sizeof(double));
		if (e == cudaSuccess) e = cudaMalloc((void **)&s.d_min, (size_t)s.n * sizeof(double));
		if (e == cudaSuccess) e = cudaMalloc((void **)&s.d_max, (size_t)s.n * sizeof(double));
		if (e == cudaSuccess) e = cudaMalloc((void **)&s.d_vals, (size_t)s.n * sizeof(double));
		if (e == cudaSuccess) e = cudaMalloc((void **)&s.d_arg, (size_t)s.n * sizeof(long long));
		if (e == cudaSuccess) e = cudaMalloc((void **)&s.d_rows, (size_t)s.n * sizeof(long long));
		if (e != cudaSuccess) {
			set_error("live table (%d x %d doubles) on device %d failed: %s", nlive, s.n, s.device,
			          cudaGetErrorString(e));
			lt_free(t);
			return e == cudaErrorMemoryAllocation ? MDNS_ENOMEM : MDNS_ECUDA;
		}
	}
	*out = t;
	return MDNS_OK;
}

int mdns_livetable_destroy(mdns_livetable *t)
{
	if (t) lt_free(t);
	return MDNS_OK;
}

int mdns_livetable_upload(mdns_livetable *t, const double *L)
{
	if (!t || !L) {
		set_error("mdns_livetable_upload: need the table and L");
		return MDNS_EINVAL;
	}
	for (auto &s : t->shards) {
		MDNS_CUDA(cudaSetDevice(s.device));
		MDNS_CUDA(cudaMemcpy2DAsync(s.T, (size_t)s.n * sizeof(double), L + s.i0,
		                            (size_t)t->ndata * sizeof(double), (size_t)s.n * sizeof(double),
		                            t->nlive, cudaMemcpyHostToDevice, s.stream));
	}
	return lt_sync(t);
}

int mdns_livetable_download(mdns_livetable *t, double *L)
{
	if (!t || !L) {
		set_error("mdns_livetable_download: need the table and L");
		return MDNS_EINVAL;
	}
	for (auto &s : t->shards) {
		MDNS_CUDA(cudaSetDevice(s.device));
		MDNS_CUDA(cudaMemcpy2DAsync(L + s.i0, (size_t)t->ndata * sizeof(double), s.T,
		                            (size_t)s.n * sizeof(double), (size_t)s.n * sizeof(double),
		                            t->nlive, cudaMemcpyDeviceToHost, s.stream));
	}
	return lt_sync(t);
}

int mdns_livetable_fill_from_launch(mdns_livetable *t, mdns_dataset *ds, int row0)
{
	if (!t || !ds) {
		set_error("mdns_livetable_fill_from_launch: need the table and the data set");
		return MDNS_EINVAL;
	}
	for (size_t k = 0; k < t->shards.size(); ++k) {
		LtShard &s = t->shards[k];
		int n_act = 0, K = 0, n = 0;
		const double *d_out = nullptr;
		void *stream = nullptr;
		if (mdns_internal_shard_view(ds, (int)k, nullptr, nullptr, &n, &n_act, &K, &d_out, &stream) !=
		    MDNS_OK)
			return MDNS_EINVAL;
		if (n != s.n || n_act != n || !d_out) {
			set_error("fill_from_launch needs a launch with every data set active on the table's own "
			          "data set (shard %zu: %d of %d active)", k, n_act, n);
			return MDNS_ESTATE;
		}
		if (row0 < 0 || row0 + K > t->nlive) {
			set_error("rows [%d, %d) outside the table of %d live points", row0, row0 + K, t->nlive);
			return MDNS_EINVAL;
		}
		MDNS_CUDA(cudaSetDevice(s.device));
		// on the data set's stream: ordered after the launch that produced the rows
		MDNS_CUDA(cudaMemcpyAsync(s.T + (size_t)row0 * s.n, d_out, (size_t)K * s.n * sizeof(double),
		                          cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
		MDNS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
	}
	return MDNS_OK;
}

int mdns_livetable_colstats(mdns_livetable *t, double *Lmins, int64_t *Lmini, double *Lmax)
{
	if (!t) {
		set_error("null live table");
		return MDNS_EINVAL;
	}
	for (auto &s : t->shards) {
		MDNS_CUDA(cudaSetDevice(s.device));
		lt_colstats_kernel<<<ceil_div(s.n, 256), 256, 0, s.stream>>>(s.T, t->nlive, s.n, s.d_min, s.d_arg,
		                                                            s.d_max);
		MDNS_LAUNCHED("lt_colstats_kernel");
		if (Lmins)
			MDNS_CUDA(cudaMemcpyAsync(Lmins + s.i0, s.d_min, (size_t)s.n * sizeof(double),
			                          cudaMemcpyDeviceToHost, s.stream));
		if (Lmini)
			MDNS_CUDA(cudaMemcpyAsync(Lmini + s.i0, s.d_arg, (size_t)s.n * sizeof(long long),
			                          cudaMemcpyDeviceToHost, s.stream));
		if (Lmax)
			MDNS_CUDA(cudaMemcpyAsync(Lmax + s.i0, s.d_max, (size_t)s.n * sizeof(double),
			                          cudaMemcpyDeviceToHost, s.stream));
	}
	return lt_sync(t);
}

// The current minimum of every data set's live points becomes its accept threshold on the data
// set's devices (the `Lmins` of a superset draw, multi_nested_sampler.py:134-137 -> :462-472 ->
// hiermetriclearn.py:193) without a round trip through the host: live table -> thresholds ->
// mdns_clike_first_accept(..., Lmins = NULL, ...).  Needs the all-active mask.
int mdns_livetable_stage_thresholds(mdns_livetable *t, mdns_dataset *ds)
{
	if (!t || !ds) {
		set_error("mdns_livetable_stage_thresholds: need the table and its data set");
		return MDNS_EINVAL;
	}
	if (mdns_internal_shard_count(ds) != (int)t->shards.size()) {
		set_error("the live table belongs to a data set with another sharding");
		return MDNS_EINVAL;
	}
	for (size_t k = 0; k < t->shards.size(); ++k) {
		LtShard &s = t->shards[k];
		int n = 0;
		void *stream = nullptr;
		mdns_internal_shard_view(ds, (int)k, nullptr, nullptr, &n, nullptr, nullptr, nullptr, &stream);
		if (n != s.n) {
			set_error("the live table belongs to a data set with another sharding");
			return MDNS_EINVAL;
		}
		double *d_lmins = nullptr;
		int rc = mdns_internal_threshold_buffer(ds, (int)k, &d_lmins);
		if (rc != MDNS_OK) return rc;
		MDNS_CUDA(cudaSetDevice(s.device));
		lt_colstats_kernel<<<ceil_div(s.n, 256), 256, 0, s.stream>>>(s.T, t->nlive, s.n, d_lmins, nullptr,
		                                                            nullptr);
		MDNS_LAUNCHED("lt_colstats_kernel");
		// the data set's stream consumes the thresholds: order it behind this kernel
		MDNS_CUDA(cudaStreamSynchronize(s.stream));
	}
	return MDNS_OK;
}

int mdns_livetable_replace(mdns_livetable *t, const int64_t *rows, const double *values)
{
	if (!t || !rows || !values) {
		set_error("mdns_livetable_replace: need the table, rows and values");
		return MDNS_EINVAL;
	}
	for (auto &s : t->shards) {
		MDNS_CUDA(cudaSetDevice(s.device));
		MDNS_CUDA(cudaMemcpyAsync(s.d_rows, rows + s.i0, (size_t)s.n * sizeof(long long),
		                          cudaMemcpyHostToDevice, s.stream));
		MDNS_CUDA(cudaMemcpyAsync(s.d_vals, values + s.i0, (size_t)s.n * sizeof(double),
		                          cudaMemcpyHostToDevice, s.stream));
		lt_replace_kernel<<<ceil_div(s.n, 256), 256, 0, s.stream>>>(s.T, t->nlive, s.n, s.d_rows, s.d_vals);
		MDNS_LAUNCHED("lt_replace_kernel");
	}
	return lt_sync(t);
}

int mdns_livetable_lmins_higher(mdns_livetable *t, const int *indices, int nidx,
                                const int64_t *shelf_offsets, const double *shelf_values,
                                double *out)
{
	if (!t || (nidx > 0 && (!indices || !shelf_offsets || !out))) {
		set_error("mdns_livetable_lmins_higher: need the table, indices, shelf_offsets and out");
		return MDNS_EINVAL;
	}
	if (nidx <= 0) return MDNS_OK;
	for (int j = 0; j < nidx; ++j) {
		if (indices[j] < 0 || indices[j] >= t->ndata || shelf_offsets[j + 1] < shelf_offsets[j]) {
			set_error("lmins_higher: entry %d (data set %d) is malformed", j, indices[j]);
			return MDNS_EINVAL;
		}
		if (j > 0 && indices[j] <= indices[j - 1]) {
			set_error("lmins_higher: data-set indices must be increasing");
			return MDNS_EINVAL;
		}
	}
	// the listed data sets are increasing, so each shard owns one contiguous run of the list
	int j0 = 0;
	for (auto &s : t->shards) {
		int j1 = j0;
		while (j1 < nidx && indices[j1] < s.i0 + s.n) ++j1;
		const int cnt = j1 - j0;
		if (cnt > 0) {
			MDNS_CUDA(cudaSetDevice(s.device));
			std::vector<int> cols(cnt);
			std::vector<long long> off(cnt + 1);
			const long long base = shelf_offsets[j0];
			long long longest = 0;
			for (int j = 0; j < cnt; ++j) {
				cols[j] = indices[j0 + j] - s.i0;
				off[j] = shelf_offsets[j0 + j] - base;
				longest = std::max<long long>(longest, shelf_offsets[j0 + j + 1] - shelf_offsets[j0 + j]);
			}
			off[cnt] = shelf_offsets[j1] - base;
			const size_t nshelf = (size_t)off[cnt];
			int rc = lt_grow(&s.d_cols, &s.cols_cap, (size_t)cnt + 1);
			if (rc == MDNS_OK) rc = lt_grow(&s.d_off, &s.off_cap, (size_t)cnt + 1);
			if (rc == MDNS_OK) rc = lt_grow(&s.d_rank, &s.rank_cap, (size_t)cnt + 1);
			if (rc == MDNS_OK) rc = lt_grow(&s.d_shelf, &s.shelf_cap, nshelf + 1);
			if (rc != MDNS_OK) return rc;
			MDNS_CUDA(cudaMemcpyAsync(s.d_cols, cols.data(), (size_t)cnt * sizeof(int),
			                          cudaMemcpyHostToDevice, s.stream));
			MDNS_CUDA(cudaMemcpyAsync(s.d_off, off.data(), (size_t)(cnt + 1) * sizeof(long long),
			                          cudaMemcpyHostToDevice, s.stream));
			if (nshelf)
				MDNS_CUDA(cudaMemcpyAsync(s.d_shelf, shelf_values + base, nshelf * sizeof(double),
				                          cudaMemcpyHostToDevice, s.stream));
			if (longest + 1 <= 8)
				lt_rank_kernel<8><<<ceil_div(cnt, 128), 128, 0, s.stream>>>(s.T, t->nlive, s.n, s.d_cols, cnt,
				                                                           s.d_off, s.d_shelf, s.d_rank);
			else
				lt_rank_kernel<32><<<ceil_div(cnt, 128), 128, 0, s.stream>>>(s.T, t->nlive, s.n, s.d_cols, cnt,
				                                                            s.d_off, s.d_shelf, s.d_rank);
			MDNS_LAUNCHED("lt_rank_kernel");
			MDNS_CUDA(cudaMemcpyAsync(out + j0, s.d_rank, (size_t)cnt * sizeof(double),
			                          cudaMemcpyDeviceToHost, s.stream));
			// cols/off live on the host stack of this call: finish before they go away
			MDNS_CUDA(cudaStreamSynchronize(s.stream));
		}
		j0 = j1;
	}
	return MDNS_OK;
}

int mdns_livetable_upload_points(mdns_livetable *t, const int64_t *live_pointsp)
{
	if (!t || !live_pointsp) {
		set_error("mdns_livetable_upload_points: need the table and live_pointsp");
		return MDNS_EINVAL;
	}
	LtShard &s = t->shards[0];
	MDNS_CUDA(cudaSetDevice(s.device));
	const size_t count = (size_t)t->nlive * t->ndata;
	if (!t->P) {
		MDNS_CUDA(cudaMalloc((void **)&t->P, count * sizeof(int)));
		MDNS_CUDA(cudaMalloc((void **)&t->d_label, (size_t)t->ndata * sizeof(int)));
		MDNS_CUDA(cudaMalloc((void **)&t->d_flag, 2 * sizeof(int)));
		MDNS_CUDA(cudaMalloc((void **)&t->d_cmask, (size_t)t->ndata));
		MDNS_CUDA(cudaMalloc((void **)&t->d_prow, (size_t)t->ndata * sizeof(long long)));
		MDNS_CUDA(cudaMalloc((void **)&t->d_pid, (size_t)t->ndata * sizeof(long long)));
	}
	// stage the int64 table in chunks and narrow it to int32 on the device
	const size_t chunk = (size_t)1 << 24;
	long long *stage = nullptr;
	MDNS_CUDA(cudaMalloc((void **)&stage, std::min(chunk, count) * sizeof(long long)));
	cudaError_t e = cudaMemsetAsync(t->d_flag, 0, 2 * sizeof(int), s.stream);
	for (size_t o = 0; o < count && e == cudaSuccess; o += chunk) {
		const size_t c = std::min(chunk, count - o);
		e = cudaMemcpyAsync(stage, live_pointsp + o, c * sizeof(long long), cudaMemcpyHostToDevice,
		                    s.stream);
		if (e == cudaSuccess) {
			lt_narrow_points_kernel<<<(unsigned)((c + 255) / 256), 256, 0, s.stream>>>(stage, c, t->P + o,
			                                                                      t->d_flag);
			e = cudaGetLastError();
		}
	}
	int bad = 0;
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(&bad, t->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s.stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
	cudaFree(stage);
	if (e != cudaSuccess) {
		set_error("upload of live_pointsp failed: %s", cudaGetErrorString(e));
		return MDNS_ECUDA;
	}
	if (bad) {
		set_error("live_pointsp holds ids outside [0, 2^31-2]");
		return MDNS_EINVAL;
	}
	g_launches.fetch_add(1);
	return MDNS_OK;
}

int mdns_livetable_replace_points(mdns_livetable *t, const int64_t *rows, const int64_t *ids)
{
	if (!t || !rows || !ids || !t->P) {
		set_error("mdns_livetable_replace_points: need the table with uploaded points, rows and ids");
		return MDNS_EINVAL;
	}
	LtShard &s = t->shards[0];
	MDNS_CUDA(cudaSetDevice(s.device));
	MDNS_CUDA(cudaMemcpyAsync(t->d_prow, rows, (size_t)t->ndata * sizeof(long long),
	                          cudaMemcpyHostToDevice, s.stream));
	MDNS_CUDA(cudaMemcpyAsync(t->d_pid, ids, (size_t)t->ndata * sizeof(long long),
	                          cudaMemcpyHostToDevice, s.stream));
	lt_replace_points_kernel<<<ceil_div(t->ndata, 256), 256, 0, s.stream>>>(t->P, t->nlive, t->ndata,
	                                                                       t->d_prow, t->d_pid);
	MDNS_LAUNCHED("lt_replace_points_kernel");
	MDNS_CUDA(cudaStreamSynchronize(s.stream));
	return MDNS_OK;
}

int mdns_livetable_subsets(mdns_livetable *t, const uint8_t *data_mask, int64_t npoints,
                           int32_t *labels, int *ncomponents, int *nrounds)
{
	if (!t || !labels || !t->P || npoints <= 0 || npoints > 0x7ffffffeLL) {
		set_error("mdns_livetable_subsets: need the table with uploaded points, labels and the "
		          "size of the point pile");
		return MDNS_EINVAL;
	}
	LtShard &s = t->shards[0];
	const int n = t->ndata;
	MDNS_CUDA(cudaSetDevice(s.device));
	if ((size_t)npoints > t->pointmin_cap) {
		if (t->d_pointmin) MDNS_CUDA(cudaFree(t->d_pointmin));
		t->d_pointmin = nullptr;
		t->pointmin_cap = 0;
		const size_t cap = (size_t)npoints + (size_t)npoints / 2 + 1024;
		MDNS_CUDA(cudaMalloc((void **)&t->d_pointmin, cap * sizeof(int)));
		t->pointmin_cap = cap;
	}
	if (data_mask)
		MDNS_CUDA(cudaMemcpyAsync(t->d_cmask, data_mask, (size_t)n, cudaMemcpyHostToDevice, s.stream));
	const int blocks = ceil_div(n, 256);
	cc_init_kernel<<<blocks, 256, 0, s.stream>>>(data_mask ? t->d_cmask : nullptr, n, t->d_label);
	MDNS_LAUNCHED("cc_init_kernel");
	int rounds = 0;
	for (;;) {
		++rounds;
		// 0x7f7f7f7f > any label
		MDNS_CUDA(cudaMemsetAsync(t->d_pointmin, 0x7f, (size_t)npoints * sizeof(int), s.stream));
		MDNS_CUDA(cudaMemsetAsync(t->d_flag, 0, sizeof(int), s.stream));
		cc_push_kernel<<<blocks, 256, 0, s.stream>>>(t->P, t->nlive, n, t->d_label, t->d_pointmin);
		MDNS_LAUNCHED("cc_push_kernel");
		cc_pull_kernel<<<blocks, 256, 0, s.stream>>>(t->P, t->nlive, n, t->d_label, t->d_pointmin,
		                                            t->d_flag);
		MDNS_LAUNCHED("cc_pull_kernel");
		cc_jump_kernel<<<blocks, 256, 0, s.stream>>>(n, t->d_label);
		MDNS_LAUNCHED("cc_jump_kernel");
		int changed = 0;
		MDNS_CUDA(cudaMemcpyAsync(&changed, t->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
		MDNS_CUDA(cudaStreamSynchronize(s.stream));
		if (!changed) break;
		if (rounds > n + 2) {
			set_error("subset partition did not converge");
			return MDNS_ECUDA;
		}
	}
	MDNS_CUDA(cudaMemcpyAsync(labels, t->d_label, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost,
	                          s.stream));
	MDNS_CUDA(cudaStreamSynchronize(s.stream));
	if (ncomponents) {
		int c = 0;
		for (int d = 0; d < n; ++d) c += labels[d] == d;
		*ncomponents = c;
	}
	if (nrounds) *nrounds = rounds;
	return MDNS_OK;
}

}  // extern "C"
