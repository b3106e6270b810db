// likelihood_kernels.cu -- batched Gaussian chi-square likelihood kernels (sm_100a).
//
// Reference behaviour being replaced: clike.c:64-76 (scalar noise, compacted output) and
// cmuselike.c:48-64 (per-element variance, free amplitude).  Design:
//   * data-set-major resident rows -> every active data set is a contiguous, 128-byte
//     aligned run in HBM whatever the mask looks like;
//   * L lanes of a warp (8 or 32) share one data set: 128-bit coalesced streaming loads,
//     FP64 accumulation, butterfly warp-shuffle reduction per data set;
//   * the K model spectra of the current candidate batch are staged into shared memory
//     with one bulk-TMA copy (cp.async.bulk + mbarrier) per CTA;
//   * stable mask compaction (prefix sum) gives the output slot k of clike.c:67-74.
#include "kernels.cuh"

namespace mdns {

// ------------------------------------------------------------------ upload ---
__global__ void __launch_bounds__(256) transpose_rows_kernel(const double *__restrict__ in,
                                                             size_t ld_in, int nx, int nb,
                                                             double *__restrict__ out,
                                                             size_t pitch, int recip)
{
	__shared__ double tile[32][33];
	const int c0 = blockIdx.x * 32;  // data sets
	const int j0 = blockIdx.y * 32;  // channels
	for (int dy = threadIdx.y; dy < 32; dy += 8) {
		const int j = j0 + dy, c = c0 + threadIdx.x;
		double v = 0.0;
		if (j < nx && c < nb) {
			v = in[(size_t)j * ld_in + c];
			if (recip) v = __ddiv_rn(1.0, v);
		}
		tile[dy][threadIdx.x] = v;
	}
	__syncthreads();
	for (int dy = threadIdx.y; dy < 32; dy += 8) {
		const int c = c0 + dy, j = j0 + threadIdx.x;
		if (c < nb && j < nx) out[(size_t)c * pitch + j] = tile[threadIdx.x][dy];
	}
}

int launch_transpose_rows(const double *in, size_t ld_in, int nx, int nb, double *out,
                          size_t pitch, int recip, cudaStream_t st)
{
	dim3 grid(ceil_div(nb, 32), ceil_div(nx, 32));
	transpose_rows_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ld_in, nx, nb, out, pitch, recip);
	MDNS_LAUNCHED("transpose_rows_kernel");
	return MDNS_OK;
}

// ------------------------------------------------------------ model batch ---
// One CTA per candidate: the model spectrum, its sum of squares (Smm of the expanded form) and,
// by CTA 0, the reset of the expanded kernels' per-pass list counters -- one launch per batch.
__global__ void __launch_bounds__(256) line_model_kernel(const double *__restrict__ x, int nx,
                                                         const double *__restrict__ params, int K,
                                                         double *__restrict__ model, int mpitch,
                                                         double *__restrict__ smm,
                                                         int *__restrict__ counters, int ncounters)
{
	__shared__ double warp_sums[8];
	// the likelihood kernel that follows may start its prologue and its first row copies now;
	// it reads the spectra, Smm and the counters only after its pdl_wait()
	pdl_trigger();
	const int k = blockIdx.x;
	double A = 0.0, mu = 0.0, sig = 1.0;
	if (k < K) {
		A = params[3 * k];
		mu = params[3 * k + 1];
		sig = params[3 * k + 2];
	}
	double s = 0.0;
	for (int j = threadIdx.x; j < mpitch; j += 256) {
		double v = 0.0;
		if (k < K && j < nx) {
			// clike.c:65, un-fused like the reference build
			const double t = __ddiv_rn(__dsub_rn(mu, x[j]), sig);
			v = __dmul_rn(A, exp(__dmul_rn(-0.5, __dmul_rn(t, t))));
		}
		model[(size_t)k * mpitch + j] = v;
		s = fma(v, v, s);
	}
	if (smm) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
		if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
		__syncthreads();
		if (threadIdx.x == 0) {
			double tot = 0.0;
#pragma unroll
			for (int w = 0; w < 8; ++w) tot += warp_sums[w];
			smm[k] = tot;
		}
	}
	if (counters && k == 0)
		for (int i = threadIdx.x; i < ncounters; i += 256) counters[1 + i] = 0;
}

int launch_line_model(const double *x, int nx, const double *params, int K, int Kpad,
                      double *model, int mpitch, double *smm, int *counters, int ncounters,
                      cudaStream_t st)
{
	line_model_kernel<<<Kpad, 256, 0, st>>>(x, nx, params, K, model, mpitch, smm, counters,
	                                        ncounters);
	MDNS_LAUNCHED("line_model_kernel");
	return MDNS_OK;
}

__global__ void pad_spectra_kernel(const double *__restrict__ src, int nx, int K, int Kpad,
                                   double *__restrict__ model, int mpitch)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int k = blockIdx.y;
	if (j >= mpitch || k >= Kpad) return;
	model[(size_t)k * mpitch + j] = (k < K && j < nx) ? src[(size_t)k * nx + j] : 0.0;
}

int launch_pad_spectra(const double *src, int nx, int K, int Kpad, double *model, int mpitch,
                       cudaStream_t st)
{
	dim3 grid(ceil_div(mpitch, 128), Kpad);
	pad_spectra_kernel<<<grid, 128, 0, st>>>(src, nx, K, Kpad, model, mpitch);
	MDNS_LAUNCHED("pad_spectra_kernel");
	return MDNS_OK;
}

// ---------------------------------------------------------- row sum of squares ---
// Syy / Smm of the expanded form (clike_tile_kernel<XP>): one warp per row, 128-bit loads,
// FP64 accumulation, butterfly reduction.  The zero pad of an odd channel count adds 0.
__global__ void __launch_bounds__(256) row_sumsq_kernel(const double *__restrict__ rows,
                                                        long long n_rows, long long pitch,
                                                        int nfrag, double *__restrict__ out)
{
	const int lane = threadIdx.x & 31;
	const long long warps = (long long)gridDim.x * 8;
	for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += warps) {
		const double2 *p = reinterpret_cast<const double2 *>(rows + r * pitch);
		double s0 = 0.0, s1 = 0.0;
		for (int f = lane; f < nfrag; f += 32) {
			const double2 y = __ldg(p + f);
			s0 = fma(y.x, y.x, s0);
			s1 = fma(y.y, y.y, s1);
		}
		double s = s0 + s1;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
		if (lane == 0) out[r] = s;
	}
}

int launch_row_sumsq(const double *rows, long long n_rows, long long pitch, int nx, double *out,
                     cudaStream_t st)
{
	if (n_rows <= 0) return MDNS_OK;
	long long blocks = (n_rows + 7) / 8;
	if (blocks > 148 * 16) blocks = 148 * 16;
	row_sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, n_rows, pitch, (nx + 1) >> 1, out);
	MDNS_LAUNCHED("row_sumsq_kernel");
	return MDNS_OK;
}

// -------------------------------------------------------- mask compaction ---
constexpr int CM_THREADS = 256;
constexpr int CM_TILE = CM_THREADS * 16;  // mask bytes per block

__device__ __forceinline__ int nonzero_bytes(uint32_t w)
{
	w |= w >> 4;
	w |= w >> 2;
	w |= w >> 1;
	return __popc(w & 0x01010101u);
}

// the mask buffer is allocated zero-padded to a multiple of 16 bytes
__device__ __forceinline__ uint4 mask_word16(const uint8_t *mask, int n, long long base)
{
	if (base >= n) return make_uint4(0, 0, 0, 0);
	return *reinterpret_cast<const uint4 *>(mask + base);
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
	__shared__ int warp_sums[CM_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int y = __shfl_up_sync(0xffffffffu, x, o);
		if (lane >= o) x += y;
	}
	if (lane == 31) warp_sums[warp] = x;
	__syncthreads();
	int before = 0, all = 0;
#pragma unroll
	for (int w = 0; w < CM_THREADS / 32; ++w) {
		const int s = warp_sums[w];
		if (w < warp) before += s;
		all += s;
	}
	*total = all;
	return before + x - v;
}

__global__ void __launch_bounds__(CM_THREADS) mask_count_kernel(const uint8_t *__restrict__ mask,
                                                                int n, int *__restrict__ block_counts)
{
	const long long base = ((long long)blockIdx.x * CM_THREADS + threadIdx.x) * 16;
	const uint4 w = mask_word16(mask, n, base);
	const int c = nonzero_bytes(w.x) + nonzero_bytes(w.y) + nonzero_bytes(w.z) + nonzero_bytes(w.w);
	int total;
	block_exclusive_scan(c, &total);
	if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_counts_kernel(int *__restrict__ counts, int nb,
                                                           int *__restrict__ total_out)
{
	__shared__ int warp_sums[32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int carry = 0;
	for (int base = 0; base < nb; base += 1024) {
		const int i = base + threadIdx.x;
		const int v = i < nb ? counts[i] : 0;
		int x = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) warp_sums[warp] = x;
		__syncthreads();
		int before = 0, all = 0;
		for (int w = 0; w < 32; ++w) {
			const int s = warp_sums[w];
			if (w < warp) before += s;
			all += s;
		}
		if (i < nb) counts[i] = carry + before + x - v;
		carry += all;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(CM_THREADS) mask_scatter_kernel(const uint8_t *__restrict__ mask,
                                                                  int n,
                                                                  const int *__restrict__ block_offsets,
                                                                  int *__restrict__ active)
{
	const long long base = ((long long)blockIdx.x * CM_THREADS + threadIdx.x) * 16;
	const uint4 w = mask_word16(mask, n, base);
	const int c = nonzero_bytes(w.x) + nonzero_bytes(w.y) + nonzero_bytes(w.z) + nonzero_bytes(w.w);
	int total;
	int pos = block_offsets[blockIdx.x] + block_exclusive_scan(c, &total);
	if (c == 0) return;
	const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
	for (int q = 0; q < 4; ++q) {
#pragma unroll
		for (int b = 0; b < 4; ++b) {
			if ((words[q] >> (8 * b)) & 0xffu) active[pos++] = (int)(base + 4 * q + b);
		}
	}
}

int launch_compact_mask(const uint8_t *mask, int n, int *scratch, int *active, int *n_active_dev,
                        cudaStream_t st)
{
	const int nb = ceil_div(n, CM_TILE);
	mask_count_kernel<<<nb, CM_THREADS, 0, st>>>(mask, n, scratch);
	MDNS_LAUNCHED("mask_count_kernel");
	scan_counts_kernel<<<1, 1024, 0, st>>>(scratch, nb, n_active_dev);
	MDNS_LAUNCHED("scan_counts_kernel");
	mask_scatter_kernel<<<nb, CM_THREADS, 0, st>>>(mask, n, scratch, active);
	MDNS_LAUNCHED("mask_scatter_kernel");
	return MDNS_OK;
}

// ------------------------------------------------------------ clike kernel ---
constexpr int LK_THREADS = 256;

// Stage KT model spectra (contiguous rows of the padded model buffer) into shared
// memory with one bulk-TMA copy; every thread then waits on the mbarrier.
template <int KT>
__device__ __forceinline__ void stage_model_tma(double2 *sm, uint64_t *bar, const double *model,
                                                int mpitch, int k0)
{
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		mbar_fence_init();
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint32_t bytes = (uint32_t)KT * (uint32_t)mpitch * 8u;
		mbar_expect_tx(bar, bytes);
		tma_load_1d(sm, model + (size_t)k0 * mpitch, bytes, bar);
	}
}

// One CTA = 8 warps; each warp serves 32/L data sets at a time, L lanes per data set.
// Per data set and lane: U 128-bit fragments are requested back to back (memory-level
// parallelism), then consumed against the KT staged model spectra.
// IM (KT = 1 only): the single candidate comes by value and every CTA builds the model spectrum
// in shared memory itself (clike.c:65, the same un-fused operations as line_model_kernel, so
// the spectrum is bit-identical): one launch per call instead of upload + model kernel + this.
template <int L, int U, int KT, bool IM>
__global__ void __launch_bounds__(LK_THREADS) clike_rows_kernel(const LikeArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double2 *sm = reinterpret_cast<double2 *>(smem_raw);
	__shared__ uint64_t bar;
	constexpr int G = 32 / L;                     // data sets per warp
	constexpr int RPC = (LK_THREADS / 32) * G;    // data sets per CTA step
	const int k0 = blockIdx.y * KT;
	const int mfp = a.mpitch >> 1;                // fragments per model row
	const int nfrag = (a.nx + 1) >> 1;
	if (IM) {
		double *smd = reinterpret_cast<double *>(smem_raw);
		for (int j = threadIdx.x; j < a.mpitch; j += LK_THREADS) {
			double v = 0.0;
			if (j < a.nx) {
				const double t = __ddiv_rn(__dsub_rn(a.line_mu, a.x[j]), a.line_sig);
				v = __dmul_rn(a.line_A, exp(__dmul_rn(-0.5, __dmul_rn(t, t))));
			}
			smd[j] = v;
		}
		__syncthreads();
	} else {
		stage_model_tma<KT>(sm, &bar, a.model, a.mpitch, k0);
	}

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g = lane / L, gl = lane % L;
	const int nchunks = (nfrag + L * U - 1) / (L * U);
	const double inv = a.scale / a.noise2;
	if (!IM) mbar_wait(&bar, 0);

	// fused accept test (hiermetriclearn.py:193): accepting data sets per candidate, counted per
	// warp with a ballot and flushed with one atomic per (warp, candidate) at the end
	int acnt[KT];
#pragma unroll
	for (int k = 0; k < KT; ++k) acnt[k] = 0;

	for (long long rb = (long long)blockIdx.x * RPC + warp * G; rb < a.n_rows;
	     rb += (long long)gridDim.x * RPC) {
		const long long r = rb + g;
		const bool valid = r < a.n_rows;
		double acc0[KT], acc1[KT];
#pragma unroll
		for (int k = 0; k < KT; ++k) acc0[k] = acc1[k] = 0.0;
		const double lm = (valid && a.lmins) ? __ldg(a.lmins + r) : 0.0;
		if (valid) {
			const long long row = a.active ? (long long)a.active[r] : r;
			const double2 *p = reinterpret_cast<const double2 *>(a.Y + row * a.pitch);
			for (int c = 0; c < nchunks; ++c) {
				const int f = c * (L * U) + gl;
				double2 y[U];
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int fi = f + u * L;
					if (fi < nfrag) y[u] = ldg_stream(p + fi);
				}
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int fi = f + u * L;
					if (fi < nfrag) {
#pragma unroll
						for (int k = 0; k < KT; ++k) {
							const double2 m = sm[k * mfp + fi];
							const double d0 = m.x - y[u].x;
							const double d1 = m.y - y[u].y;
							acc0[k] = fma(d0, d0, acc0[k]);
							acc1[k] = fma(d1, d1, acc1[k]);
						}
					}
				}
			}
		}
#pragma unroll
		for (int k = 0; k < KT; ++k) {
			double s = acc0[k] + acc1[k];
#pragma unroll
			for (int o = L / 2; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
			const bool mine = valid && gl == 0 && k0 + k < a.K;
			const double val = s * inv;
			if (mine && a.out) a.out[(long long)(k0 + k) * a.out_stride + r] = val;
			if (a.counts) acnt[k] += __popc(__ballot_sync(0xffffffffu, mine && val > lm));
		}
	}
	if (a.counts && lane == 0) {
#pragma unroll
		for (int k = 0; k < KT; ++k)
			if (acnt[k]) atomicAdd(a.counts + k0 + k, acnt[k]);
	}
}

// ---- one speculative pass over a HANDFUL of active data sets, one launch, no copies ----------
// The focussed regime of the constrained draw (hiermetriclearn.py:181-196 with a joint mask of one
// or a few data sets, hundreds of rejected candidates per accepted one): the work is a few
// microseconds, the cost was the choreography -- parameter upload, model kernel, likelihood
// kernel, decision kernel, download, synchronisation (36 us of native time per pass).  Here the K
// candidates come BY VALUE with the launch, one CTA builds their spectra in shared memory
// (clike.c:65, un-fused like line_model_kernel), scores the rows exactly as
// clike_rows_kernel<32, ...> does (same fragment order per lane, same butterfly: the logL are
// bit-identical to what the general path returns for such masks), applies `L > Lmins`, picks the
// first accepted candidate and writes {counts, first, its logL vector} straight into pinned host
// memory, the sequence number last.  The host polls that word: no copy node, no stream wait.
constexpr int DS_MAX_K = 32, DS_MAX_ROWS = 64, DS_KT = 8;
struct DrawSmallParams {
	double p[DS_MAX_K * 3];
};
// host block layout (ints): [0] sequence, [1] first, [2..2+DS_MAX_K) counts; doubles from byte 256
constexpr int DS_HOST_VALUES_OFFSET = 256;

__global__ void __launch_bounds__(LK_THREADS) draw_small_kernel(const LikeArgs a, const DrawSmallParams prm,
                                                                const int seq, int *__restrict__ host_block)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *smd = reinterpret_cast<double *>(smem_raw);                  // [K][mpitch] spectra
	const double2 *sm = reinterpret_cast<const double2 *>(smem_raw);
	__shared__ double s_val[DS_MAX_K * DS_MAX_ROWS];
	__shared__ int s_cnt[DS_MAX_K];
	__shared__ int s_first;
	const int K = a.K, n = a.n_rows;
	const int mfp = a.mpitch >> 1;
	const int nfrag = (a.nx + 1) >> 1;
	// (four independent elements per thread at a time: see clike_small_kernel)
	for (int base = threadIdx.x; base < K * a.mpitch; base += 4 * LK_THREADS) {
		double v[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const int idx = base + u * LK_THREADS;
			const int k = idx / a.mpitch, j = idx - k * a.mpitch;
			const bool on = idx < K * a.mpitch && j < a.nx;
			const int kk = on ? k : 0;
			const double t = __ddiv_rn(__dsub_rn(prm.p[3 * kk + 1], a.x[on ? j : 0]), prm.p[3 * kk + 2]);
			const double e = __dmul_rn(prm.p[3 * kk], exp(__dmul_rn(-0.5, __dmul_rn(t, t))));
			v[u] = on ? e : 0.0;
		}
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (base + u * LK_THREADS < K * a.mpitch) smd[base + u * LK_THREADS] = v[u];
	}
	if (threadIdx.x < DS_MAX_K) s_cnt[threadIdx.x] = 0;
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const double inv = a.scale / a.noise2;
	for (int r = warp; r < n; r += LK_THREADS / 32) {
		const long long row = a.active ? (long long)a.active[r] : r;
		const double2 *p = reinterpret_cast<const double2 *>(a.Y + row * a.pitch);
		const double lm = __ldg(a.lmins + r);
		for (int k0 = 0; k0 < K; k0 += DS_KT) {
			double acc0[DS_KT], acc1[DS_KT];
#pragma unroll
			for (int k = 0; k < DS_KT; ++k) acc0[k] = acc1[k] = 0.0;
			for (int fi = lane; fi < nfrag; fi += 32) {
				const double2 y = ldg_stream(p + fi);
#pragma unroll
				for (int k = 0; k < DS_KT; ++k) {
					if (k0 + k < K) {
						const double2 m = sm[(k0 + k) * mfp + fi];
						const double d0 = m.x - y.x;
						const double d1 = m.y - y.y;
						acc0[k] = fma(d0, d0, acc0[k]);
						acc1[k] = fma(d1, d1, acc1[k]);
					}
				}
			}
#pragma unroll
			for (int k = 0; k < DS_KT; ++k) {
				double sum = acc0[k] + acc1[k];
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) sum += shfl_xor_f64(sum, o);
				if (lane == 0 && k0 + k < K) {
					const double val = sum * inv;
					s_val[(k0 + k) * DS_MAX_ROWS + r] = val;
					if (a.out) a.out[(long long)(k0 + k) * a.out_stride + r] = val;
					if (val > lm) atomicAdd(&s_cnt[k0 + k], 1);
				}
			}
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		int first = -1;
		for (int k = 0; k < K && first < 0; ++k)
			if (s_cnt[k] > 0) first = k;
		s_first = first;
		host_block[1] = first;
	}
	if (threadIdx.x < K) host_block[2 + threadIdx.x] = s_cnt[threadIdx.x];
	__syncthreads();
	const int first = s_first;
	if (first >= 0) {
		double *hv = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(host_block) + DS_HOST_VALUES_OFFSET);
		for (int r = threadIdx.x; r < n; r += LK_THREADS) hv[r] = s_val[first * DS_MAX_ROWS + r];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		*reinterpret_cast<volatile int *>(host_block) = seq;
	}
}

bool draw_small_fits(const LikeArgs &a)
{
	return a.K >= 1 && a.K <= DS_MAX_K && a.n_rows >= 1 && a.n_rows <= DS_MAX_ROWS && a.x && a.lmins &&
	       !a.W && (size_t)a.K * a.mpitch * 8 <= 160 * 1024;
}

size_t draw_small_host_bytes() { return DS_HOST_VALUES_OFFSET + DS_MAX_ROWS * sizeof(double); }

// params: K rows (A, mu, sig) on the HOST; host_block: pinned, mapped; seq: any value the block
// does not hold yet
int launch_draw_small(const LikeArgs &a, const double *params, int seq, int *host_block, cudaStream_t st)
{
	if (!draw_small_fits(a)) {
		set_error("draw_small_kernel: at most %d candidates x %d active data sets", DS_MAX_K, DS_MAX_ROWS);
		return MDNS_EINVAL;
	}
	DrawSmallParams prm;
	memcpy(prm.p, params, (size_t)a.K * 3 * sizeof(double));
	const size_t smem = (size_t)a.K * a.mpitch * 8;
	static size_t smem_set = 0;
	if (smem > smem_set) {
		MDNS_CUDA(cudaFuncSetAttribute(draw_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                               (int)(160 * 1024)));
		smem_set = 160 * 1024;
	}
	draw_small_kernel<<<1, LK_THREADS, smem, st>>>(a, prm, seq, host_block);
	MDNS_LAUNCHED("draw_small_kernel");
	return MDNS_OK;
}

// ---- small candidate batches in ONE launch ---------------------------------------------------
// Up to a few 1e5 model x data-set evaluations (1e4 data sets x 16 candidates, the size of the
// reference's own runs: BASELINE configs[0] / [1]) are a few microseconds of arithmetic; the
// three-launch step of the tensor path (model kernel, contraction, fix-up) costs 20 us of launch
// latencies there.  Here every CTA builds the KT spectra of its candidate slice in shared memory
// from the staged parameter points (clike.c:65, un-fused like line_model_kernel: bit-identical
// spectra) and scores its rows in the DIRECT form: 8 lanes per data set, the fragment order per
// lane and the butterfly of clike_rows_kernel<8, ...> (so the two agree to the last bit), every
// lane group walking R data sets at once so that one 128-bit read of the model serves R of them
// (the first version, R = 1 with 16 candidates per slice, ran at 30 % of the FP64 peak: four
// shared-memory wavefronts per four FP64 instructions), warps strided over the groups of 4 R
// rows so that the FP64 work spreads evenly over the SMs.  No row sums, no guard, no fix-up list;
// the accept test is fused in like everywhere else.
template <int KT, int R>
__global__ void __launch_bounds__(LK_THREADS, (KT == 4 && R == 4) ? 1 : 2) clike_small_kernel(const LikeArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *smd = reinterpret_cast<double *>(smem_raw);                  // [KT][mpitch] spectra
	const double2 *sm = reinterpret_cast<const double2 *>(smem_raw);
	constexpr int L = 8, G = 32 / L, U = R >= 4 ? 1 : 2;
	const int k0 = blockIdx.y * KT;
	const int mfp = a.mpitch >> 1;
	const int nfrag = (a.nx + 1) >> 1;
	// (four elements per thread at a time: a division and an exp are ~70 dependent instructions,
	// and one after the other the 6 elements of a thread were 4 us on every CTA's critical path)
	for (int base = threadIdx.x; base < KT * a.mpitch; base += 4 * LK_THREADS) {
		double v[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) {
			const int idx = base + u * LK_THREADS;
			const int k = idx / a.mpitch, j = idx - k * a.mpitch;
			const bool on = idx < KT * a.mpitch && k0 + k < a.K && j < a.nx;
			const double *p = a.params + 3 * (on ? k0 + k : 0);
			const double t = __ddiv_rn(__dsub_rn(p[1], a.x[on ? j : 0]), p[2]);
			const double e = __dmul_rn(p[0], exp(__dmul_rn(-0.5, __dmul_rn(t, t))));
			v[u] = on ? e : 0.0;
		}
#pragma unroll
		for (int u = 0; u < 4; ++u)
			if (base + u * LK_THREADS < KT * a.mpitch) smd[base + u * LK_THREADS] = v[u];
	}
	__syncthreads();

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g = lane / L, gl = lane % L;
	const double inv = a.scale / a.noise2;
	int acnt[KT];
#pragma unroll
	for (int k = 0; k < KT; ++k) acnt[k] = 0;
	const long long ngroups = ((long long)a.n_rows + G * R - 1) / (G * R);
	const long long wstride = (long long)gridDim.x * (LK_THREADS / 32);
	for (long long q = (long long)warp * gridDim.x + blockIdx.x; q < ngroups; q += wstride) {
		// rows (q R + j) G + g, j < R: every 128-bit load of the warp covers four consecutive rows.
		// Rows behind the end are computed on the last row and dropped.
		const double2 *p[R];
		double acc0[R][KT], acc1[R][KT];
#pragma unroll
		for (int j = 0; j < R; ++j) {
			long long r = (q * R + j) * G + g;
			if (r >= a.n_rows) r = a.n_rows - 1;
			const long long row = a.active ? (long long)a.active[r] : r;
			p[j] = reinterpret_cast<const double2 *>(a.Y + row * a.pitch);
#pragma unroll
			for (int k = 0; k < KT; ++k) acc0[j][k] = acc1[j][k] = 0.0;
		}
		for (int f = gl; f < nfrag; f += L * U) {
			double2 y[R][U];
#pragma unroll
			for (int j = 0; j < R; ++j)
#pragma unroll
				for (int u = 0; u < U; ++u)
					if (f + u * L < nfrag) y[j][u] = ldg_stream(p[j] + f + u * L);
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int fi = f + u * L;
				if (fi < nfrag) {
#pragma unroll
					for (int k = 0; k < KT; ++k) {
						const double2 m = sm[k * mfp + fi];
#pragma unroll
						for (int j = 0; j < R; ++j) {
							const double d0 = m.x - y[j][u].x;
							const double d1 = m.y - y[j][u].y;
							acc0[j][k] = fma(d0, d0, acc0[j][k]);
							acc1[j][k] = fma(d1, d1, acc1[j][k]);
						}
					}
				}
			}
		}
#pragma unroll
		for (int j = 0; j < R; ++j) {
			const long long r = (q * R + j) * G + g;
			const bool valid = r < a.n_rows;
			const double lm = (valid && a.lmins) ? __ldg(a.lmins + r) : 0.0;
#pragma unroll
			for (int k = 0; k < KT; ++k) {
				double s = acc0[j][k] + acc1[j][k];
#pragma unroll
				for (int o = L / 2; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
				const bool mine = valid && gl == 0 && k0 + k < a.K;
				const double val = s * inv;
				if (mine && a.out) a.out[(long long)(k0 + k) * a.out_stride + r] = val;
				if (a.counts) acnt[k] += __popc(__ballot_sync(0xffffffffu, mine && val > lm));
			}
		}
	}
	if (a.counts && lane == 0) {
#pragma unroll
		for (int k = 0; k < KT; ++k)
			if (acnt[k]) atomicAdd(a.counts + k0 + k, acnt[k]);
	}
}

// candidates per slice (every slice reads the rows again and builds its own spectra)
int clike_small_ktile(int K) { return K <= 2 ? 2 : K <= 4 ? 4 : 8; }

bool clike_small_fits(int K, int mpitch)
{
	return K >= 1 && (size_t)clike_small_ktile(K) * mpitch * 8 <= 200 * 1024;
}

template <int KT, int R>
static int launch_clike_small_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	const size_t smem = (size_t)KT * a.mpitch * 8;
	auto kern = clike_small_kernel<KT, R>;
	if (smem > 48 * 1024)
		MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LK_THREADS, smem));
	if (occ < 1) {
		set_error("clike_small_kernel does not fit: %zu bytes of shared memory", smem);
		return MDNS_EINVAL;
	}
	if (occ > 2) occ = 2;
	// one group of 4 R rows per warp at least; every CTA builds the spectra, so no more CTAs
	// than there is row work for -- and all candidate slices together in ONE wave (the first
	// version let every slice fill the device: 15026 rows x 2 slices ran as 470 CTAs on 296 slots)
	const int slices = ceil_div(a.K, KT);
	long long gx = ceil_div(ceil_div(a.n_rows, 4 * R), LK_THREADS / 32);
	long long resident = (long long)sm_count * occ / slices;
	if (resident < 1) resident = 1;
	if (gx > resident) gx = resident;
	if (gx < 1) gx = 1;
	kern<<<dim3((unsigned)gx, slices), LK_THREADS, smem, st>>>(a);
	MDNS_LAUNCHED("clike_small_kernel");
	return MDNS_OK;
}

static int launch_clike_small(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	// rows per lane group: as many as still leave every warp of the device a group of its own
	// (5000 rows in groups of 16 are 40 CTAs: R = 4 lost to the lanes-across-channels kernel there)
	const long long warps = (long long)sm_count * (LK_THREADS / 32);
	const int rmax = a.n_rows >= 16 * warps ? 4 : a.n_rows >= 8 * warps ? 2 : 1;
	switch (clike_small_ktile(a.K)) {
	case 2:
		return rmax == 4   ? launch_clike_small_inst<2, 4>(a, sm_count, st)
		       : rmax == 2 ? launch_clike_small_inst<2, 2>(a, sm_count, st)
		                   : launch_clike_small_inst<2, 1>(a, sm_count, st);
	case 4:
		return rmax == 4   ? launch_clike_small_inst<4, 4>(a, sm_count, st)
		       : rmax == 2 ? launch_clike_small_inst<4, 2>(a, sm_count, st)
		                   : launch_clike_small_inst<4, 1>(a, sm_count, st);
	default:
		return rmax >= 2 ? launch_clike_small_inst<8, 2>(a, sm_count, st)
		                 : launch_clike_small_inst<8, 1>(a, sm_count, st);
	}
}

// ---- register-blocked variant for candidate batches (K >= 4) -----------------
// With KT candidates per pass the streaming kernel above becomes bound by shared-memory
// bandwidth: every (element, candidate) pair needs 8 bytes of model from shared memory
// and the crossbar delivers 128 lane-bytes/clk/SM = 16 pairs/clk, half of what the FP64
// pipe (64 lanes, 2 ops per pair) can retire.  Here each group of L lanes walks R data
// sets at once, so a model fragment fetched from shared memory is used R times.

// Sum KT per-lane partials over the L lanes of a group with a transposing butterfly:
// each step halves the number of live values, so 8 values over 8 lanes cost 7 shuffles
// instead of 24.  On return lane `gl` holds the group total of candidate `kidx` in v[0].
template <int N, int O>
struct GroupReduce {
	template <int KT>
	static __device__ __forceinline__ void run(double (&v)[KT], int gl, int &kidx)
	{
		if constexpr (O >= 1) {
			if constexpr (N > 1) {
				constexpr int H = N / 2;
				const bool upper = (gl & O) != 0;
#pragma unroll
				for (int j = 0; j < H; ++j) {
					const double keep = upper ? v[j + H] : v[j];
					const double send = upper ? v[j] : v[j + H];
					v[j] = keep + shfl_xor_f64(send, O);
				}
				if (upper) kidx += H;
				GroupReduce<H, O / 2>::run(v, gl, kidx);
			} else {
				v[0] += shfl_xor_f64(v[0], O);
				GroupReduce<1, O / 2>::run(v, gl, kidx);
			}
		}
	}
};

template <int L, int U, int R, int KT>
__global__ void __launch_bounds__(LK_THREADS, 2) clike_block_kernel(const LikeArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double2 *sm = reinterpret_cast<double2 *>(smem_raw);
	__shared__ uint64_t bar;
	constexpr int G = 32 / L;                         // groups per warp
	constexpr int RPC = (LK_THREADS / 32) * G * R;    // data sets per CTA step
	// lanes of a group that own a distinct candidate total after the reduction
	constexpr int OWNERS = KT < L ? KT : L;
	const int k0 = blockIdx.y * KT;
	const int mfp = a.mpitch >> 1;
	const int nfrag = (a.nx + 1) >> 1;
	__shared__ int s_cnt[KT];
	if (threadIdx.x < KT) s_cnt[threadIdx.x] = 0;
	stage_model_tma<KT>(sm, &bar, a.model, a.mpitch, k0);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g = lane / L, gl = lane % L;
	const int nchunks = (nfrag + L * U - 1) / (L * U);
	const double inv = a.scale / a.noise2;
	mbar_wait(&bar, 0);
	int acnt = 0, my_k = 0;      // fused accept test: this lane's candidate and its count

	for (long long rb = (long long)blockIdx.x * RPC + (warp * G + g) * R; rb - g * R < a.n_rows;
	     rb += (long long)gridDim.x * RPC) {
		const double2 *p[R];
#pragma unroll
		for (int r = 0; r < R; ++r) {
			// rows past the end re-read the last row (results discarded) so that the
			// whole warp stays convergent for the shuffles
			long long rr = rb + r < a.n_rows ? rb + r : (long long)a.n_rows - 1;
			const long long row = a.active ? (long long)a.active[rr] : rr;
			p[r] = reinterpret_cast<const double2 *>(a.Y + row * a.pitch);
		}
		double acc[R][KT];
#pragma unroll
		for (int r = 0; r < R; ++r)
#pragma unroll
			for (int k = 0; k < KT; ++k) acc[r][k] = 0.0;
		for (int c = 0; c < nchunks; ++c) {
			const int f = c * (L * U) + gl;
			double2 y[R][U];
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int fi = f + u * L;
				if (fi < nfrag) {
#pragma unroll
					for (int r = 0; r < R; ++r) y[r][u] = ldg_stream(p[r] + fi);
				}
			}
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int fi = f + u * L;
				if (fi < nfrag) {
#pragma unroll
					for (int k = 0; k < KT; ++k) {
						const double2 m = sm[k * mfp + fi];
#pragma unroll
						for (int r = 0; r < R; ++r) {
							const double d0 = m.x - y[r][u].x;
							const double d1 = m.y - y[r][u].y;
							acc[r][k] = fma(d0, d0, acc[r][k]);
							acc[r][k] = fma(d1, d1, acc[r][k]);
						}
					}
				}
			}
		}
#pragma unroll
		for (int r = 0; r < R; ++r) {
			int kidx = 0;
			GroupReduce<KT, L / 2>::run(acc[r], gl, kidx);
			// after the transposing steps the lanes whose low bits are zero own a total
			const bool owner = (gl & (L / OWNERS - 1)) == 0;
			if (owner && rb + r < a.n_rows && k0 + kidx < a.K) {
				const double val = acc[r][0] * inv;
				if (a.out) a.out[(long long)(k0 + kidx) * a.out_stride + rb + r] = val;
				if (a.lmins && val > __ldg(a.lmins + rb + r)) ++acnt;
				my_k = kidx;
			}
		}
	}
	if (a.counts) {
		if (acnt) atomicAdd(&s_cnt[my_k], acnt);
		__syncthreads();
		if (threadIdx.x < KT && s_cnt[threadIdx.x])
			atomicAdd(a.counts + k0 + threadIdx.x, s_cnt[threadIdx.x]);
	}
}

template <int L, int U, int R, int KT>
static int launch_clike_block_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	constexpr int RPC = (LK_THREADS / 32) * (32 / L) * R;
	const size_t smem = (size_t)KT * a.mpitch * 8;
	auto kern = clike_block_kernel<L, U, R, KT>;
	if (smem > 48 * 1024)
		MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                               (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LK_THREADS, smem));
	if (occ < 1) {
		set_error("clike block kernel does not fit: %zu bytes of shared memory", smem);
		return MDNS_EINVAL;
	}
	const int ktiles = ceil_div(a.K, KT);
	long long gx = ceil_div(a.n_rows, RPC);
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	if (gx < 1) gx = 1;
	kern<<<dim3((unsigned)gx, ktiles), LK_THREADS, smem, st>>>(a);
	MDNS_LAUNCHED("clike_block_kernel");
	return MDNS_OK;
}

// supported (rows, unroll) shapes of the block kernel, lanes fixed at 8
static int launch_clike_block(const LikeArgs &a, int rows, int u, int kt, int sm_count,
                              cudaStream_t st)
{
#define MDNS_BLK(RR, UU)                                                                  \
	if (rows == RR && u == UU)                                                        \
		return kt == 4 ? launch_clike_block_inst<8, UU, RR, 4>(a, sm_count, st)   \
		               : launch_clike_block_inst<8, UU, RR, 8>(a, sm_count, st)
	MDNS_BLK(2, 2);
	MDNS_BLK(2, 4);
	MDNS_BLK(4, 1);
	MDNS_BLK(4, 2);
#undef MDNS_BLK
	set_error("unsupported block-kernel shape rows=%d unroll=%d", rows, u);
	return MDNS_EINVAL;
}

template <int L, int U, int KT>
static int launch_clike_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	constexpr int RPC = (LK_THREADS / 32) * (32 / L);
	const size_t smem = (size_t)KT * a.mpitch * 8;
	const bool im = KT == 1 && a.inline_model && a.K == 1;
	auto kern = im ? clike_rows_kernel<L, U, 1, true> : clike_rows_kernel<L, U, KT, false>;
	if (smem > 48 * 1024)   // per device, so set it on every launch that needs it
		MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                               (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LK_THREADS, smem));
	if (occ < 1) {
		set_error("clike kernel does not fit: %zu bytes of shared memory", smem);
		return MDNS_EINVAL;
	}
	const int ktiles = ceil_div(a.K, KT);
	long long gx = ceil_div(a.n_rows, RPC);
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	if (gx < 1) gx = 1;
	kern<<<dim3((unsigned)gx, ktiles), LK_THREADS, smem, st>>>(a);
	MDNS_LAUNCHED("clike_rows_kernel");
	return MDNS_OK;
}

template <int L, int U>
static int launch_clike_k(const LikeArgs &a, int kt, int sm_count, cudaStream_t st)
{
	switch (kt) {
	case 1: return launch_clike_inst<L, U, 1>(a, sm_count, st);
	case 2: return launch_clike_inst<L, U, 2>(a, sm_count, st);
	case 4: return launch_clike_inst<L, U, 4>(a, sm_count, st);
	default: return launch_clike_inst<L, U, 8>(a, sm_count, st);
	}
}

template <int L>
static int launch_clike_u(const LikeArgs &a, int u, int kt, int sm_count, cudaStream_t st)
{
	switch (u) {
	case 1: return launch_clike_k<L, 1>(a, kt, sm_count, st);
	case 2: return launch_clike_k<L, 2>(a, kt, sm_count, st);
	case 4: return launch_clike_k<L, 4>(a, kt, sm_count, st);
	case 8: return launch_clike_k<L, 8>(a, kt, sm_count, st);
	case 13: return launch_clike_k<L, 13>(a, kt, sm_count, st);
	default: return launch_clike_k<L, 16>(a, kt, sm_count, st);
	}
}

// Choose the number of fragments in flight per lane: the largest supported value that wastes
// at most ~7 % of the issue slots (a first version minimised the waste alone and picked U = 1
// for 1000-channel rows: 0.74 of the HBM roofline instead of 0.99).
static int pick_unroll(int nfrag, int L)
{
	static const int cand[] = {16, 13, 8, 4, 2, 1};
	const int per_lane = ceil_div(nfrag, L);
	for (int u : cand) {
		const long long cost = (long long)ceil_div(per_lane, u) * u;
		if (cost * 100 <= (long long)per_lane * 107) return u;
	}
	return 1;
}

static int pick_ktile(int K, int mpitch, int requested)
{
	int kt = requested > 0 ? requested : (K >= 8 ? 8 : K >= 4 ? 4 : K >= 2 ? 2 : 1);
	if (kt != 1 && kt != 2 && kt != 4 && kt != 8) kt = kt > 8 ? 8 : kt > 4 ? 4 : kt > 2 ? 2 : 1;
	// keep the staged model tile within ~96 KB so that two CTAs fit per SM
	while (kt > 1 && (size_t)kt * mpitch * 8 > 96 * 1024) kt >>= 1;
	return kt;
}

// the stream-K tensor-path kernel when its workspace is there, else round 1's whole-tile kernel
static int launch_dmma_auto(const LikeArgs &a, int kt, int stages, int sm_count, cudaStream_t st,
                            int *accept_fused)
{
	if (kt == 16 && stages == 13) stages = 3;     // 8 consumer warps x 32 data sets (see rows_dmma_fits)
	if (rows_dmma_fits(a, kt, stages)) {
		if (accept_fused) *accept_fused = 1;
		return launch_rows_dmma(a, kt, stages, false, 1, sm_count, st);
	}
	return launch_clike_dmma(a, kt, stages, sm_count, st);
}

constexpr int MASKED_TENSOR_MIN_ROWS = 4096;   // masked batches of >= 5 candidates: tensor path from here

int launch_clike(const LikeArgs &a, const Tuning &t, int sm_count, cudaStream_t st, int *accept_fused)
{
	if (accept_fused) *accept_fused = 0;
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	if (a.params) {
		// small batch of parameter points: model, likelihood and accept test in one launch (the
		// caller decided, and launched no model kernel: capi.cu inline_batch)
		if (accept_fused) *accept_fused = 1;
		return launch_clike_small(a, sm_count, st);
	}
	const int nfrag = (a.nx + 1) >> 1;
	int L = t.lanes;
	const bool tile_ok = a.tmap && !a.active && 4LL * a.mpitch <= tile_constant_capacity();
	if (L == 3 && (!a.active || a.tmap_gather)) {
		// expanded form, cross term as FP64 tensor-core tiles
		int kt = t.ktile;
		if (kt != 8 && kt != 16 && kt != 32) kt = a.K >= 32 ? 32 : a.K >= 16 ? 16 : 8;
		int stages = ((t.rows >= 2 && t.rows <= 4) || (t.rows >= 12 && t.rows <= 14)) ? t.rows : 3;
		while (kt > 8 && !dmma_fits(a, kt, stages)) kt >>= 1;
		if (!dmma_fits(a, kt, stages)) stages = 2;
		// unroll = 1 selects round 1's whole-tile kernel (A/B measurements)
		if (t.unroll == 1) return launch_clike_dmma(a, kt, stages, sm_count, st);
		return launch_dmma_auto(a, kt, stages, sm_count, st, accept_fused);
	}
	if (L == 6 && (!a.active || a.tmap_gather)) {
		// expanded form, per-warp slabs with the batch resident in shared memory (short spectra)
		const int kt = t.ktile == 8 ? 8 : 16;
		const int nslot = t.rows == 2 ? 2 : 3;
		if (slab_dmma_fits(a, kt, nslot)) {
			if (accept_fused) *accept_fused = 1;
			return launch_slab_dmma(a, kt, nslot, sm_count, st);
		}
		Tuning d;
		d.allow_expanded = t.allow_expanded;
		return launch_clike(a, d, sm_count, st, accept_fused);
	}
	if (L == 2 && !a.active) {
		// expanded form, register-blocked over data sets (all-active rows only)
		int kt = t.ktile;
		if (kt != 8 && kt != 16 && kt != 32) kt = a.K >= 16 ? 16 : 8;
		int stages = t.rows == 2 ? 2 : 3;
		while (kt > 8 && !xtile_fits(a, kt, stages)) kt >>= 1;
		if (!xtile_fits(a, kt, stages)) stages = 2;
		const int lane_rows = t.unroll == 4 && kt <= 16 ? 4 : 2;
		return launch_clike_xtile(a, kt, lane_rows, stages, sm_count, st);
	}
	if (L == 1 && tile_ok) {
		// lane-per-data-set tile kernel (tensor-TMA ring), all-active rows only
		int kt = t.ktile;
		if (kt != 4 && kt != 8 && kt != 16 && kt != 32) kt = a.K >= 16 ? 16 : a.K >= 8 ? 8 : 4;
		while (kt > 4 && (long long)kt * a.mpitch > tile_constant_capacity()) kt >>= 1;
		// unroll: 16/32 channels per stage with 128-row tiles, 116/132 with 256-row tiles
		const int tile_rows = t.unroll >= 100 ? 256 : 128;
		const int cw = t.unroll >= 100 ? t.unroll - 100 : t.unroll;
		const int nbox = cw == 32 ? 2 : 1;
		int stages = (t.rows == 3 || t.rows == 4 || t.rows == 6) ? t.rows : 3;
		if (tile_rows == 128 && nbox == 1 && stages == 3) stages = 4;
		return launch_clike_tile(a, tile_rows == 256 ? a.tmap256 : a.tmap, kt, nbox, stages,
		                         tile_rows, sm_count, st);
	}
	if (L == 0 && t.unroll == 0 && t.ktile == 0 && t.rows == 0 && a.active && a.tmap_gather &&
	    t.allow_expanded && a.K >= XP_MIN_K_MASKED && a.n_rows >= MASKED_TENSOR_MIN_ROWS) {
		// masked candidate batches: the tensor path fed by gather4 copies of the listed rows
		// (two producer warps).  Measured, 5e5 active of 1e6 data sets: K=8 0.152 ms (block kernel
		// 0.180), K=16 0.173 ms (0.358), K=32 0.43 ms (0.71); up to 4 candidates the block kernel
		// wins (K=4 0.141 ms).
		// per-warp slabs, every warp gathering its own listed rows (slab_dmma_kernel.cu); same box,
		// 5e5 active of 1e6 x 200 (tools/sweep_masked.py): K=16 0.159 ms (0.84 of the roofline)
		// against 0.181 (0.73) for the stream-K gather, K=8 0.150 (0.85) against 0.152 for round 1's
		// whole-tile gather and 0.241 (0.53) for the stream-K 16-warp gather shape
		// From 4096 active rows (it was 32768): the gathered slab kernel is a 16-18 us step at 16
		// candidates whatever the size below 5e4 rows, the lanes-across-channels kernels that
		// ran there took 22 / 35 / 48 / 56 us at 5053 / 10070 / 15026 / 24963 rows (K = 32: 37 ... 104
		// against 31-33; tools/r2_small_arms.py, profiles/r02_small_arms.json); smaller batches
		// go to clike_small_kernel before they get here (capi.cu inline_batch)
		if (a.K > 8 && slab_dmma_fits(a, 16, 2)) {
			if (accept_fused) *accept_fused = 1;
			return launch_slab_dmma(a, 16, 2, sm_count, st);
		}
		// (passes of 16 whatever K: 0.158 ms each against 0.43 ms for a gathered pass of 32)
		if (a.K >= 32 && dmma_fits(a, 32, 3)) return launch_dmma_auto(a, 32, 3, sm_count, st, accept_fused);
		if (a.K >= 16 && dmma_fits(a, 16, 13)) return launch_dmma_auto(a, 16, 13, sm_count, st, accept_fused);
		if (slab_dmma_fits(a, 8, 2)) {
			if (accept_fused) *accept_fused = 1;
			return launch_slab_dmma(a, 8, 2, sm_count, st);
		}
		if (dmma_fits(a, 8, 2)) return launch_clike_dmma(a, 8, 2, sm_count, st);
		if (dmma_fits(a, 8, 13)) return launch_dmma_auto(a, 8, 13, sm_count, st, accept_fused);
	}
	if (L == 1 || L == 2 || L == 3 || L == 6) {
		// tile kernel requested but not applicable (masked rows): automatic choice
		Tuning d;
		d.allow_expanded = t.allow_expanded;
		return launch_clike(a, d, sm_count, st, accept_fused);
	}
	// (from 8192 data sets: L2 flushed, 200 channels, tools/r2_small_n.py -- 1e4: K=16 0.027 ms
	// against 0.038 for the lanes-across-channels kernel, K=8 equal; 3e4: 0.033 / 0.030 against
	// 0.068 / 0.041; the threshold used to be 32768)
	if (L == 0 && t.unroll == 0 && t.ktile == 0 && t.rows == 0 && !a.active && a.K >= XP_MIN_K &&
	    a.n_rows >= 8192) {
		// automatic choice for all-active candidate batches: expanded form with the cross term
		// on the FP64 tensor path when allowed (measured at N=1e6, C=200: K=8 0.277 ms, K=16
		// 0.313 ms, K=32 0.46 ms; FMA form 0.29 / 0.35 / 0.69; direct form 0.30 / 0.56 / 1.1) ...
		if (t.allow_expanded) {
			// K=32: 8 consumer warps x 32 data sets, 3 stages (0.46 ms); K=16: 16 warps x 16 data
			// sets, 3 stages (0.295 ms; 8 warps 0.300); K=8: 16 warps, 4 stages (0.263 ms; 8 warps
			// 0.277).  Stage counts + 10 select the 16-warp shape.  From K = 3 on a pass of 8
			// (padded) candidates on the tensor path beats the lanes-across-channels kernels
			// (K=4: 0.266 ms vs 0.289 ms block kernel).
			// passes of 16 on the slab kernel (0.285 ms each at 1e6 x 200) against passes of 32 on
			// the stream-K kernel (0.45-0.47 ms each, FP64-tensor-bound): K = 17..32 is one pass of
			// 32 (0.45-0.52 ms; two slab passes 0.54-0.60), 33..48 three passes of 16
			const int passes16 = (a.K + 15) / 16, passes32 = (a.K + 31) / 32;
			const bool prefer16 = passes16 * 5 < passes32 * 8;
			if (a.K > 16 && !prefer16 && dmma_fits(a, 32, 3)) return launch_dmma_auto(a, 32, 3, sm_count, st, accept_fused);
			// 9..16 candidates on short spectra: per-warp slabs, batch resident in shared memory
			// (slab_dmma_kernel.cu).  Same box, 1e6 x 200, K=16: 0.281-0.284 ms (0.93-0.94 of the
			// roofline) against 0.305-0.314 for the stream-K kernel; 192 channels 0.262 vs 0.282;
			// K=24 (two passes of 16) 0.554 vs 0.596.  Two slots per warp beat three (0.288); up to
			// 8 candidates the stream-K kernel stays ahead (0.258 vs 0.268 ms)
			// Small launches: every warp starts with a fixed slab, interleaved over the SMs, so the
			// kernel is at least level with the stream-K kernel from 1e4 data sets (L2 flushed,
			// tools/r2_warmup_drift.py and r2_small_n.py: 1e4 0.027 / 0.027 ms, 1e5 0.050 / 0.057, 3e5 0.117 / 0.121,
			// 1e6 0.285 / 0.322); taken from four slabs per SM
			if (a.K > 8 && slab_dmma_fits(a, 16, 2) && slab_dmma_slabs(a) >= 4LL * sm_count) {
				if (accept_fused) *accept_fused = 1;
				return launch_slab_dmma(a, 16, 2, sm_count, st);
			}
			if (a.K >= 32 && dmma_fits(a, 32, 3)) return launch_dmma_auto(a, 32, 3, sm_count, st, accept_fused);
			if (a.K >= 16 && dmma_fits(a, 16, 13)) return launch_dmma_auto(a, 16, 13, sm_count, st, accept_fused);
			if (dmma_fits(a, 8, 14)) return launch_dmma_auto(a, 8, 14, sm_count, st, accept_fused);
			const int xkt = a.K >= 16 && xtile_fits(a, 16, 2) ? 16 : 8;
			if (xtile_fits(a, xkt, 2)) return launch_clike_xtile(a, xkt, 2, 2, sm_count, st);
		}
		// ... else the direct-form tile kernel (measured at N=1e6, C=200: K=8 0.29 ms vs
		// 0.35 ms block kernel; K=16 0.51 ms vs 0.71 ms)
		if (a.K >= 8 && tile_ok && 8LL * a.mpitch <= tile_constant_capacity()) {
			const int kt = (a.K >= 16 && 16LL * a.mpitch <= tile_constant_capacity()) ? 16 : 8;
			return launch_clike_tile(a, a.tmap256, kt, 1, 3, 256, sm_count, st);
		}
	}
	// 8 lanes per data set for short rows (more data sets in flight per warp), a whole warp for
	// long ones (measured at 1000 channels: 0.150 ms vs 0.155 ms for 1 GB)
	if (L != 8 && L != 32) L = (a.n_rows >= 16384 && nfrag >= 8 && nfrag < 256) ? 8 : 32;
	int U = t.unroll;
	if (U != 1 && U != 2 && U != 4 && U != 8 && U != 13 && U != 16) U = pick_unroll(nfrag, L);
	const int kt = pick_ktile(a.K, a.mpitch, t.ktile);
	if ((size_t)kt * a.mpitch * 8 > 200 * 1024) {
		// spectra too long for a whole model row in shared memory: the tensor-path kernel
		// streams the model in 16-channel slices beside the data, whatever the channel count
		if (t.allow_expanded && dmma_fits(a, 8, 3)) return launch_dmma_auto(a, 8, 3, sm_count, st, accept_fused);
		set_error("model spectrum of %d channels does not fit in shared memory", a.nx);
		return MDNS_EINVAL;
	}
	// candidate batches: register-blocked kernel (R data sets per lane group)
	int rows = t.rows;
	// measured at N=1e6, C=200: KT=4 -> R=4,U=2 (0.27 ms); KT=8 -> R=2,U=4 (0.35 ms)
	if (rows == 0) rows = (kt >= 4 && a.n_rows >= 65536 && nfrag >= 16) ? (kt == 8 ? 2 : 4) : 1;
	if (rows > 1 && kt >= 4) {
		int bu = t.unroll;
		if (rows == 4 && bu != 1 && bu != 2) bu = 2;
		if (rows == 2 && bu != 2 && bu != 4) bu = 4;
		if (accept_fused) *accept_fused = 1;
		return launch_clike_block(a, rows, bu, kt, sm_count, st);
	}
	if (accept_fused) *accept_fused = 1;
	return L == 8 ? launch_clike_u<8>(a, U, kt, sm_count, st)
	              : launch_clike_u<32>(a, U, kt, sm_count, st);
}

// ------------------------------------------------------------ accept test ---
// hiermetriclearn.py:193 `numpy.any(L > Lmins)` on the device: counts[k] = number of active
// data sets whose logL of candidate k exceeds its threshold.  grid = (row tiles, candidates);
// the logL matrix was just written and is largely still L2-resident.
constexpr int AC_ROWS = 256 * 8;   // rows per CTA

__global__ void __launch_bounds__(256) accept_count_kernel(const double *__restrict__ L,
                                                           long long stride, int n,
                                                           const double *__restrict__ lmins,
                                                           int *__restrict__ counts)
{
	__shared__ int warp_counts[8];
	const double *row = L + (long long)blockIdx.y * stride;
	const int base = blockIdx.x * AC_ROWS;
	int c = 0;
#pragma unroll
	for (int u = 0; u < 8; ++u) {
		const int i = base + u * 256 + threadIdx.x;
		if (i < n) c += row[i] > __ldg(lmins + i) ? 1 : 0;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) warp_counts[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		int t = 0;
		for (int w = 0; w < 8; ++w) t += warp_counts[w];
		if (t) atomicAdd(counts + blockIdx.y, t);
	}
}

int launch_accept_count(const double *L, long long stride, int n, int K, const double *lmins,
                        int *counts, cudaStream_t st, bool zero)
{
	if (K <= 0) return MDNS_OK;
	if (zero) MDNS_CUDA(cudaMemsetAsync(counts, 0, (size_t)K * sizeof(int), st));
	if (n <= 0) return MDNS_OK;
	accept_count_kernel<<<dim3(ceil_div(n, AC_ROWS), K), 256, 0, st>>>(L, stride, n, lmins, counts);
	MDNS_LAUNCHED("accept_count_kernel");
	return MDNS_OK;
}

// ---- the decision on the device: no host round trip between the counts and the fetch ----
// sel[0] = first candidate with a non-zero count (the one the reference's one-at-a-time loop stops
// at, hiermetriclearn.py:181-196) or -1; sel[1] = its count; sel[2] = rows recomputed in the direct
// form so far (host feedback of the expanded form); sel[3] = spare; sel[4 + k] = counts[k].
__global__ void select_first_kernel(const int *__restrict__ counts, int K,
                                    const int *__restrict__ redo, int *__restrict__ sel)
{
	__shared__ int s_first;
	if (threadIdx.x == 0) s_first = 0x7fffffff;
	__syncthreads();
	for (int k = threadIdx.x; k < K; k += blockDim.x) {
		const int c = counts[k];
		sel[SEL_COUNTS + k] = c;
		if (c > 0) atomicMin(&s_first, k);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		const int f = s_first == 0x7fffffff ? -1 : s_first;
		sel[0] = f;
		sel[1] = f >= 0 ? counts[f] : 0;
		sel[2] = redo ? redo[0] : 0;
		sel[3] = 0;
	}
}

int launch_select_first(const int *counts, int K, const int *redo, int *sel, cudaStream_t st)
{
	select_first_kernel<<<1, 128, 0, st>>>(counts, K, redo, sel);
	MDNS_LAUNCHED_HELPER("select_first_kernel");
	return MDNS_OK;
}

// out[r] = L[sel[0]][r0 + r], r < n: the logL vector of the selected candidate (nothing if none)
__global__ void __launch_bounds__(256) gather_selected_kernel(const double *__restrict__ L,
                                                              long long stride, int r0, int n,
                                                              const int *__restrict__ sel,
                                                              double *__restrict__ out)
{
	const int f = sel[0];
	if (f < 0) return;
	const double *row = L + (long long)f * stride + r0;
	for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) out[r0 + i] = row[i];
}

int launch_gather_selected(const double *L, long long stride, int r0, int n, const int *sel,
                           double *out, cudaStream_t st)
{
	if (n <= 0) return MDNS_OK;
	int blocks = ceil_div(n, 256 * 4);
	if (blocks > 148 * 8) blocks = 148 * 8;
	gather_selected_kernel<<<blocks, 256, 0, st>>>(L, stride, r0, n, sel, out);
	MDNS_LAUNCHED_HELPER("gather_selected_kernel");
	return MDNS_OK;
}

// sparse form: flags[r] = L[sel[0]][r] > lmins[r] (all zero if no candidate was selected); the
// tail up to the next multiple of 16 bytes is cleared for the compaction
__global__ void __launch_bounds__(256) selected_flags_kernel(const double *__restrict__ L,
                                                             long long stride, int n,
                                                             const int *__restrict__ sel,
                                                             const double *__restrict__ lmins,
                                                             uint8_t *__restrict__ flags)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	const int npad = (n + 15) / 16 * 16;
	const int f = sel[0];
	if (i < npad) flags[i] = (f >= 0 && i < n && L[(long long)f * stride + i] > lmins[i]) ? 1 : 0;
}

int launch_selected_flags(const double *L, long long stride, int n, const int *sel,
                          const double *lmins, uint8_t *flags, cudaStream_t st)
{
	if (n <= 0) return MDNS_OK;
	selected_flags_kernel<<<ceil_div((n + 15) / 16 * 16, 256), 256, 0, st>>>(L, stride, n, sel, lmins, flags);
	MDNS_LAUNCHED_HELPER("selected_flags_kernel");
	return MDNS_OK;
}

// out[i] = L[sel[0]][idx[i]], i < *n_dev
__global__ void __launch_bounds__(256) gather_selected_values_kernel(const double *__restrict__ L,
                                                                     long long stride,
                                                                     const int *__restrict__ sel,
                                                                     const int *__restrict__ idx,
                                                                     const int *__restrict__ n_dev,
                                                                     double *__restrict__ out)
{
	const int f = sel[0];
	if (f < 0) return;
	const int n = *n_dev;
	const double *row = L + (long long)f * stride;
	for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) out[i] = row[idx[i]];
}

int launch_gather_selected_values(const double *L, long long stride, const int *sel, const int *idx,
                                  const int *n_dev, int n_max, double *out, cudaStream_t st)
{
	if (n_max <= 0) return MDNS_OK;
	int blocks = ceil_div(n_max, 256 * 4);
	if (blocks > 148 * 4) blocks = 148 * 4;
	gather_selected_values_kernel<<<blocks, 256, 0, st>>>(L, stride, sel, idx, n_dev, out);
	MDNS_LAUNCHED_HELPER("gather_selected_values_kernel");
	return MDNS_OK;
}

// flags[r] = L[r] > lmins[r] (one candidate); the flag buffer is zero-padded for the compaction
__global__ void __launch_bounds__(256) accept_flags_kernel(const double *__restrict__ L, int n,
                                                           const double *__restrict__ lmins,
                                                           uint8_t *__restrict__ flags)
{
	// also clears the tail up to the next multiple of 16 bytes (the compaction reads 16-byte
	// words and the number of active data sets changes from mask to mask)
	const int i = blockIdx.x * 256 + threadIdx.x;
	const int npad = (n + 15) / 16 * 16;
	if (i < npad) flags[i] = (i < n && L[i] > lmins[i]) ? 1 : 0;
}

int launch_accept_flags(const double *L, int n, const double *lmins, uint8_t *flags, cudaStream_t st)
{
	if (n <= 0) return MDNS_OK;
	accept_flags_kernel<<<ceil_div((n + 15) / 16 * 16, 256), 256, 0, st>>>(L, n, lmins, flags);
	MDNS_LAUNCHED_HELPER("accept_flags_kernel");
	return MDNS_OK;
}

__global__ void __launch_bounds__(256) gather_values_kernel(const double *__restrict__ L,
                                                            const int *__restrict__ idx, int n,
                                                            double *__restrict__ out)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i < n) out[i] = L[idx[i]];
}

int launch_gather_values(const double *L, const int *idx, int n, double *out, cudaStream_t st)
{
	if (n <= 0) return MDNS_OK;
	gather_values_kernel<<<ceil_div(n, 256), 256, 0, st>>>(L, idx, n, out);
	MDNS_LAUNCHED_HELPER("gather_values_kernel");
	return MDNS_OK;
}

// ------------------------------------------------------------- muse kernel ---
// cmuselike.c:48-64 with resident inverse variance w = 1/v:
//   s1 = sum y*m*w ; s2 = 1e-10 + sum m*m*w ; s = s1/s2 ; chi = sum (y - s*m)^2 * w
// Generic variant: the row is streamed twice (second pass served by L1/L2).
template <int L, int U>
__global__ void __launch_bounds__(LK_THREADS) muse_rows_kernel(const LikeArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double2 *sm = reinterpret_cast<double2 *>(smem_raw);
	__shared__ uint64_t bar;
	constexpr int G = 32 / L;
	constexpr int RPC = (LK_THREADS / 32) * G;
	const int k = blockIdx.y;
	const int nfrag = (a.nx + 1) >> 1;
	stage_model_tma<1>(sm, &bar, a.model, a.mpitch, k);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int g = lane / L, gl = lane % L;
	const int nchunks = (nfrag + L * U - 1) / (L * U);
	mbar_wait(&bar, 0);

	for (long long rb = (long long)blockIdx.x * RPC + warp * G; rb < a.n_rows;
	     rb += (long long)gridDim.x * RPC) {
		const long long r = rb + g;
		const bool valid = r < a.n_rows;
		long long row = 0;
		const double2 *py = nullptr, *pw = nullptr;
		double s1a = 0.0, s1b = 0.0, s2a = 0.0, s2b = 0.0;
		double2 y[U], w[U];   // kept for pass 2 when the row fits one chunk (single fetch)
		if (valid) {
			row = a.active ? (long long)a.active[r] : r;
			py = reinterpret_cast<const double2 *>(a.Y + row * a.pitch);
			pw = reinterpret_cast<const double2 *>(a.W + row * a.pitch);
			for (int c = 0; c < nchunks; ++c) {
				const int f = c * (L * U) + gl;
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int fi = f + u * L;
					if (fi < nfrag) {
						y[u] = __ldg(py + fi);
						w[u] = __ldg(pw + fi);
					}
				}
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int fi = f + u * L;
					if (fi < nfrag) {
						const double2 m = sm[fi];
						const double t0 = m.x * w[u].x, t1 = m.y * w[u].y;
						s1a = fma(y[u].x, t0, s1a);
						s1b = fma(y[u].y, t1, s1b);
						s2a = fma(m.x, t0, s2a);
						s2b = fma(m.y, t1, s2b);
					}
				}
			}
		}
		double s1 = s1a + s1b, s2 = s2a + s2b;
#pragma unroll
		for (int o = L / 2; o > 0; o >>= 1) {
			s1 += shfl_xor_f64(s1, o);
			s2 += shfl_xor_f64(s2, o);
		}
		const double s = s1 / (s2 + 1e-10);
		double ca = 0.0, cb = 0.0;
		if (valid) {
			for (int c = 0; c < nchunks; ++c) {
				const int f = c * (L * U) + gl;
				if (nchunks > 1) {   // longer rows are streamed again (served by L1/L2)
#pragma unroll
					for (int u = 0; u < U; ++u) {
						const int fi = f + u * L;
						if (fi < nfrag) {
							y[u] = __ldg(py + fi);
							w[u] = __ldg(pw + fi);
						}
					}
				}
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int fi = f + u * L;
					if (fi < nfrag) {
						const double2 m = sm[fi];
						const double r0 = fma(-s, m.x, y[u].x);
						const double r1 = fma(-s, m.y, y[u].y);
						ca = fma(r0 * r0, w[u].x, ca);
						cb = fma(r1 * r1, w[u].y, cb);
					}
				}
			}
		}
		double chi = ca + cb;
#pragma unroll
		for (int o = L / 2; o > 0; o >>= 1) chi += shfl_xor_f64(chi, o);
		if (valid && gl == 0) a.out[(long long)k * a.out_stride + row] = -0.5 * chi;
	}
}

template <int L, int U>
static int launch_muse_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	constexpr int RPC = (LK_THREADS / 32) * (32 / L);
	const size_t smem = (size_t)a.mpitch * 8;
	auto kern = muse_rows_kernel<L, U>;
	if (smem > 48 * 1024)
		MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                               (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, LK_THREADS, smem));
	if (occ < 1) {
		set_error("muse kernel does not fit: %zu bytes of shared memory", smem);
		return MDNS_EINVAL;
	}
	long long gx = ceil_div(a.n_rows, RPC);
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	if (gx < 1) gx = 1;
	kern<<<dim3((unsigned)gx, a.K), LK_THREADS, smem, st>>>(a);
	MDNS_LAUNCHED("muse_rows_kernel");
	return MDNS_OK;
}

int launch_muse(const LikeArgs &a, const Tuning &t, int sm_count, cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	const int nfrag = (a.nx + 1) >> 1;
	if ((size_t)a.mpitch * 8 > 200 * 1024) {
		set_error("model spectrum of %d channels does not fit in shared memory", a.nx);
		return MDNS_EINVAL;
	}
	int L = t.lanes;
	// long rows: one CTA per data set, rows staged once through a bulk-TMA ring
	if ((L == 0 || L == 256) && muse_block_fits(a) && (L == 256 || nfrag >= 256))
		return launch_muse_block(a, t.ktile, t.rows, sm_count, st);
	if (L != 8 && L != 32) L = (a.n_rows >= 16384 && nfrag >= 8 && nfrag <= 8 * 16) ? 8 : 32;
	const int per_lane = ceil_div(nfrag, L);
	int U = t.unroll;
	// prefer a single chunk (fragments stay in registers for pass 2) up to 16 per lane
	if (U != 2 && U != 4 && U != 8 && U != 16)
		U = per_lane > 8 ? 16 : per_lane > 4 ? 8 : per_lane > 2 ? 4 : 2;
	if (L == 8) {
		switch (U) {
		case 2: return launch_muse_inst<8, 2>(a, sm_count, st);
		case 4: return launch_muse_inst<8, 4>(a, sm_count, st);
		case 8: return launch_muse_inst<8, 8>(a, sm_count, st);
		default: return launch_muse_inst<8, 16>(a, sm_count, st);
		}
	}
	switch (U) {
	case 2: return launch_muse_inst<32, 2>(a, sm_count, st);
	case 4: return launch_muse_inst<32, 4>(a, sm_count, st);
	case 8: return launch_muse_inst<32, 8>(a, sm_count, st);
	default: return launch_muse_inst<32, 16>(a, sm_count, st);
	}
}

}  // namespace mdns
