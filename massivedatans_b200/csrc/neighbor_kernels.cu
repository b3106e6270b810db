// neighbor_kernels.cu -- RadFriends neighbour tests (sm_100a), bit-exact with cneighbors.c.
//
// Exactness contract: the squared distance is accumulated in channel order with separately
// rounded IEEE double sub/mul/add (cneighbors.c:104-107; the reference is built for baseline
// x86-64, i.e. without FMA contraction), and `sqrt(d) < r` (cneighbors.c:88,109) is evaluated
// as `d < T` with T = mdns_sqrt_threshold(r), which is the same predicate because the
// correctly rounded sqrt is monotone.  min/max over distances commute with sqrt for the
// same reason, so one sqrt is taken at the very end on the host.
//
// Mapping: one warp per candidate (or per 4 candidates), lanes over members held in SoA
// layout -> coalesced member loads, `__ballot_sync` gives the per-candidate hit count of 32
// members at once and the early exit of cneighbors.c:112.
#include "kernels.cuh"

namespace mdns {

constexpr int NB_THREADS = 256;
constexpr int MAX_TEMPLATE_DIM = 8;
constexpr unsigned long long BITS_1E300 = 0x7E37E43C8800759CULL;  // 1e300 (cneighbors.c:51,148)

template <int D>
__device__ __forceinline__ double sqdist_reg(const double (&a)[D], const double (&b)[D])
{
	double t = __dsub_rn(a[0], b[0]);
	double d = __dmul_rn(t, t);   // 0 + t*t == t*t exactly
#pragma unroll
	for (int k = 1; k < D; ++k) {
		t = __dsub_rn(a[k], b[k]);
		d = __dadd_rn(d, __dmul_rn(t, t));
	}
	return d;
}

// ------------------------------------------------------------------ count ---
// Small calls (the sampler's 400 members x 1000 candidates, two per proposal round) are pure
// latency: the candidates are read straight from pinned host memory, the counts written straight
// into it, and the last CTA to finish (ticket) raises a flag the host polls -- no copy calls, no
// stream wait.  `done` = {ticket on the device, flag in pinned memory}, null for ordinary launches.
struct CountDone {
	int *ticket;
	int *host_flag;
	int seq;
};

__device__ __forceinline__ void count_done(const CountDone &done)
{
	if (!done.ticket) return;
	__threadfence_system();      // this thread's counts, for the host
	__syncthreads();
	if (threadIdx.x == 0) {
		const int t = atomicAdd(done.ticket, 1);
		if (t == (int)gridDim.x - 1) {
			*done.ticket = 0;
			__threadfence_system();
			*reinterpret_cast<volatile int *>(done.host_flag) = done.seq;
		}
	}
}

template <int D, int CPW>
__global__ void __launch_bounds__(NB_THREADS) count_within_kernel(const double *__restrict__ xs,
                                                                  int n, int npad,
                                                                  const double *__restrict__ yy,
                                                                  int m, double T, int stop_at,
                                                                  int *__restrict__ counts,
                                                                  const CountDone done)
{
	const int lane = threadIdx.x & 31;
	const long long wid = ((long long)blockIdx.x * NB_THREADS + threadIdx.x) >> 5;
	const long long j0 = wid * CPW;
	if (j0 >= m) {
		count_done(done);
		return;
	}
	double y[CPW][D];
	int cnt[CPW];
#pragma unroll
	for (int c = 0; c < CPW; ++c) {
		const long long j = (j0 + c < m) ? j0 + c : (long long)m - 1;
		cnt[c] = 0;
#pragma unroll
		for (int k = 0; k < D; ++k) y[c][k] = __ldg(yy + j * D + k);
	}
	// two chunks of 32 members per iteration: the loads of the second are in flight while the
	// first is tested (ncu on the one-chunk loop: 25 % long-scoreboard stalls, FP64 pipe 64 %)
	for (int base = 0; base < n; base += 64) {
		const int i0 = base + lane, i1 = base + 32 + lane;
		const bool valid0 = i0 < n, valid1 = i1 < n;
		double x0[D], x1[D];
#pragma unroll
		for (int k = 0; k < D; ++k) {
			x0[k] = valid0 ? xs[(size_t)k * npad + i0] : 0.0;
			x1[k] = valid1 ? xs[(size_t)k * npad + i1] : 0.0;
		}
		bool all_done = stop_at > 0;
#pragma unroll
		for (int c = 0; c < CPW; ++c) {
			const double d = sqdist_reg<D>(x0, y[c]);
			const unsigned hits = __ballot_sync(0xffffffffu, valid0 && d < T);
			cnt[c] += __popc(hits);
			all_done = all_done && cnt[c] >= stop_at;
		}
		if (all_done) break;   // warp-uniform: cnt comes from the ballot
		if (base + 32 >= n) break;
		all_done = stop_at > 0;
#pragma unroll
		for (int c = 0; c < CPW; ++c) {
			const double d = sqdist_reg<D>(x1, y[c]);
			const unsigned hits = __ballot_sync(0xffffffffu, valid1 && d < T);
			cnt[c] += __popc(hits);
			all_done = all_done && cnt[c] >= stop_at;
		}
		if (all_done) break;
	}
	if (lane == 0) {
#pragma unroll
		for (int c = 0; c < CPW; ++c)
			if (j0 + c < m) counts[j0 + c] = cnt[c];
	}
	count_done(done);
}

// any dimensionality: coordinates streamed per channel
__global__ void __launch_bounds__(NB_THREADS) count_within_generic_kernel(
    const double *__restrict__ xs, int n, int npad, int ndim, const double *__restrict__ yy, int m,
    double T, int stop_at, int *__restrict__ counts, const CountDone done)
{
	const int lane = threadIdx.x & 31;
	const long long j = ((long long)blockIdx.x * NB_THREADS + threadIdx.x) >> 5;
	if (j >= m) {
		count_done(done);
		return;
	}
	const double *y = yy + j * ndim;
	int cnt = 0;
	for (int base = 0; base < n; base += 32) {
		const int i = base + lane;
		const bool valid = i < n;
		double d = 0.0;
		if (valid) {
			double t = __dsub_rn(xs[i], __ldg(y));
			d = __dmul_rn(t, t);
			for (int k = 1; k < ndim; ++k) {
				t = __dsub_rn(xs[(size_t)k * npad + i], __ldg(y + k));
				d = __dadd_rn(d, __dmul_rn(t, t));
			}
		}
		cnt += __popc(__ballot_sync(0xffffffffu, valid && d < T));
		if (stop_at > 0 && cnt >= stop_at) break;
	}
	if (lane == 0) counts[j] = cnt;
	count_done(done);
}

// Full counts of many candidates against many members (no early exit): lanes over CANDIDATES,
// C of them per lane in registers, the members read once per warp with uniform (broadcast)
// loads, four in flight.  Per pair test 8 FP64 instructions (the same separately rounded
// sub/mul/add sequence) and nothing else on the FP64 pipe: `d < T` is decided on the bit
// patterns -- d is a sum of squares, so it is +0, positive, +inf or NaN, and for T >= +0 the
// unsigned order of the bits is the order of the values (NaN compares above everything, like
// the floating-point test).  No ballot, no popcount: a predicated integer add per pair.
// The warp-per-candidate kernel above spends a ninth FP64 slot on the compare and waits for
// per-lane member loads (ncu, round 1: FP64 pipe 64 %, 25 % long-scoreboard stalls); it stays
// for `any` / countmax queries, where its per-candidate early exit is the point.
// ncu (profiles/r02_count_tile.md): FP64 pipe 73 %, issue slots 61 % busy, no memory stalls left.
// An FP64 instruction holds the issue port of its sub-partition for two cycles, so the bound is
// 2 x 8 + 3 (two ISETP + the add) + ~1.5 (loads, loop) = 20.5 issue cycles per pair test and
// lane = 2.9 ms for 5e9 pairs; fusing the multiply-adds would break bit-exactness.
// grid = (candidate tiles, member ranges): partial counts are added with integer atomics.
constexpr int CT_THREADS = 128;

template <int D, int C>
__global__ void __launch_bounds__(CT_THREADS) count_tile_kernel(const double *__restrict__ xs, int n, int npad,
                                                                 const double *__restrict__ yy, int m,
                                                                 unsigned long long Tbits, int per_range,
                                                                 int *__restrict__ counts)
{
	const long long j0 = ((long long)blockIdx.x * CT_THREADS + threadIdx.x) * C;
	double y[C][D];
	int cnt[C];
#pragma unroll
	for (int c = 0; c < C; ++c) {
		const long long j = j0 + c < m ? j0 + c : (long long)m - 1;
		cnt[c] = 0;
#pragma unroll
		for (int k = 0; k < D; ++k) y[c][k] = __ldg(yy + j * D + k);
	}
	const int i0 = blockIdx.y * per_range;
	const int i1 = min(n, i0 + per_range);
	int i = i0;
	for (; i + 4 <= i1; i += 4) {
		double x[4][D];
#pragma unroll
		for (int u = 0; u < 4; ++u)
#pragma unroll
			for (int k = 0; k < D; ++k) x[u][k] = __ldg(xs + (size_t)k * npad + i + u);
#pragma unroll
		for (int u = 0; u < 4; ++u)
#pragma unroll
			for (int c = 0; c < C; ++c) {
				const double d = sqdist_reg<D>(x[u], y[c]);
				cnt[c] += (unsigned long long)__double_as_longlong(d) < Tbits ? 1 : 0;
			}
	}
	for (; i < i1; ++i) {
		double x[D];
#pragma unroll
		for (int k = 0; k < D; ++k) x[k] = __ldg(xs + (size_t)k * npad + i);
#pragma unroll
		for (int c = 0; c < C; ++c) {
			const double d = sqdist_reg<D>(x, y[c]);
			cnt[c] += (unsigned long long)__double_as_longlong(d) < Tbits ? 1 : 0;
		}
	}
#pragma unroll
	for (int c = 0; c < C; ++c)
		if (j0 + c < m && cnt[c]) atomicAdd(counts + j0 + c, cnt[c]);
}

// pair tests from which the tiled kernel is taken (below, the launch is latency either way)
constexpr long long CT_MIN_PAIRS = 1LL << 24;

template <int D>
static int launch_count_tile(const double *xs, int n, int npad, const double *yy, int m, double T,
                             int *counts, int sm_count, cudaStream_t st)
{
	constexpr int C = D <= 4 ? 4 : 2;
	const int tiles = ceil_div(m, CT_THREADS * C);
	// member ranges: ~64 CTAs per SM (nine are resident: the last, partial wave is then a few per
	// cent of the run, not half of it -- 1372 CTAs on 1332 slots measured 3.39 ms), at least 256
	// members each
	int ranges = ceil_div((long long)sm_count * 64, tiles);
	if (ranges > n / 256) ranges = n / 256;
	if (ranges < 1) ranges = 1;
	const int per_range = (int)round_up(ceil_div(n, ranges), 4);
	ranges = ceil_div(n, per_range);
	MDNS_CUDA(cudaMemsetAsync(counts, 0, (size_t)m * sizeof(int), st));
	unsigned long long tb;
	memcpy(&tb, &T, sizeof tb);
	count_tile_kernel<D, C><<<dim3(tiles, ranges), CT_THREADS, 0, st>>>(xs, n, npad, yy, m, tb, per_range, counts);
	MDNS_LAUNCHED("count_tile_kernel");
	return MDNS_OK;
}

template <int D>
static int launch_count_d(const double *xs, int n, int npad, const double *yy, int m, double T,
                          int stop_at, int *counts, int sm_count, cudaStream_t st, const CountDone &done)
{
	// full counts of a large problem: lanes over candidates (T >= +0 and not NaN: bit-pattern compare)
	if (!done.ticket && stop_at == 0 && (long long)n * m >= CT_MIN_PAIRS && T >= 0.0)
		return launch_count_tile<D>(xs, n, npad, yy, m, T, counts, sm_count, st);
	const long long warps_wanted = (long long)sm_count * 16;
	if (m >= 4 * warps_wanted) {
		const int warps = ceil_div(m, 4);
		count_within_kernel<D, 4><<<ceil_div(warps, NB_THREADS / 32), NB_THREADS, 0, st>>>(
		    xs, n, npad, yy, m, T, stop_at, counts, done);
	} else {
		count_within_kernel<D, 1><<<ceil_div(m, NB_THREADS / 32), NB_THREADS, 0, st>>>(
		    xs, n, npad, yy, m, T, stop_at, counts, done);
	}
	MDNS_LAUNCHED("count_within_kernel");
	return MDNS_OK;
}

// ticket / host_flag / seq: see CountDone (ticket == nullptr: an ordinary launch)
int launch_count_within(const double *xs, int n, int npad, int ndim, const double *yy, int m,
                        double T, int stop_at, int *counts, int sm_count, cudaStream_t st, int *ticket,
                        int *host_flag, int seq)
{
	if (m <= 0) return MDNS_OK;
	const CountDone done = {ticket, host_flag, seq};
	switch (ndim) {
	case 1: return launch_count_d<1>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	case 2: return launch_count_d<2>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	case 3: return launch_count_d<3>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	case 4: return launch_count_d<4>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	case 5: return launch_count_d<5>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	case 6: return launch_count_d<6>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	case 7: return launch_count_d<7>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	case 8: return launch_count_d<8>(xs, n, npad, yy, m, T, stop_at, counts, sm_count, st, done);
	default:
		count_within_generic_kernel<<<ceil_div(m, NB_THREADS / 32), NB_THREADS, 0, st>>>(
		    xs, n, npad, ndim, yy, m, T, stop_at, counts, done);
		MDNS_LAUNCHED("count_within_generic_kernel");
		return MDNS_OK;
	}
}

// ------------------------------------------- candidate generation (fused) ---
// the ball draws of RadFriendsRegion.generate
// (radfriendsregion.py:156-178) fused with the neighbour count on the device.  Proposal p picks a
// member, a direction (normalised normal vector) and a radius r*u^(1/D), counts the members
// within r of the proposed point and keeps it with probability 1/count -- a uniform draw from
// the union of balls.  The random numbers come from a counter-based generator (Philox4x32-10
// keyed by the seed, counter = proposal index), so the result depends only on (seed, first
// proposal index), not on the launch shape.  This is NOT the numpy stream of the reference:
// parity with radfriendsregion.py is statistical (tests/test_gpu_region.py), the seed-exact
// mirror stays clustering/radfriendsregion.py on the host RNG.
struct Philox {
	uint32_t c[4], k[2];
	__device__ __forceinline__ void round()
	{
		const uint32_t lo0 = 0xD2511F53u * c[0], hi0 = __umulhi(0xD2511F53u, c[0]);
		const uint32_t lo1 = 0xCD9E8D57u * c[2], hi1 = __umulhi(0xCD9E8D57u, c[2]);
		const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
		c[0] = n0;
		c[1] = lo1;
		c[2] = n2;
		c[3] = lo0;
		k[0] += 0x9E3779B9u;
		k[1] += 0xBB67AE85u;
	}
};

// four 32-bit words for (seed, proposal, block)
__device__ __forceinline__ uint4 philox4(unsigned long long seed, unsigned long long idx, uint32_t block)
{
	Philox p;
	p.c[0] = (uint32_t)idx;
	p.c[1] = (uint32_t)(idx >> 32);
	p.c[2] = block;
	p.c[3] = 0x5eed5eedu;
	p.k[0] = (uint32_t)seed;
	p.k[1] = (uint32_t)(seed >> 32);
#pragma unroll
	for (int r = 0; r < 10; ++r) p.round();
	return make_uint4(p.c[0], p.c[1], p.c[2], p.c[3]);
}

// uniform in (0, 1) from 53 random bits
__device__ __forceinline__ double u01(uint32_t a, uint32_t b)
{
	const unsigned long long bits = ((unsigned long long)a << 21) ^ (unsigned long long)(b >> 11);
	return ((double)(bits & ((1ULL << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
}

template <int D>
__global__ void __launch_bounds__(NB_THREADS) region_generate_kernel(
    const double *__restrict__ xs, int n, int npad, double r, double T, unsigned long long seed,
    unsigned long long first, int m, double *__restrict__ points, uint8_t *__restrict__ keep,
    int *__restrict__ nnear_out)
{
	const int lane = threadIdx.x & 31;
	const long long j = ((long long)blockIdx.x * NB_THREADS + threadIdx.x) >> 5;   // proposal
	if (j >= m) return;
	const unsigned long long idx = first + (unsigned long long)j;
	// every lane computes the same proposal (cheap), then the warp shares the member scan
	const uint4 w0 = philox4(seed, idx, 0);
	const int centre = (int)(((unsigned long long)w0.x * (unsigned long long)n) >> 32);
	const double urad = u01(w0.y, w0.z);
	double y[D];
	double norm2 = 0.0;
#pragma unroll
	for (int k = 0; k < D; k += 2) {
		const uint4 w = philox4(seed, idx, 1 + k / 2);
		const double u1 = u01(w.x, w.y), u2 = u01(w.z, w.w);
		const double rad = sqrt(-2.0 * log(u1));
		double s, c;
		sincospi(2.0 * u2, &s, &c);
		y[k] = rad * c;
		norm2 += y[k] * y[k];
		if (k + 1 < D) {
			y[k + 1] = rad * s;
			norm2 += y[k + 1] * y[k + 1];
		}
	}
	const uint4 wc = philox4(seed, idx, 1 + (D + 1) / 2);
	const double coin = u01(wc.x, wc.y);
	const double scale = r * pow(urad, 1.0 / D) / sqrt(norm2);
#pragma unroll
	for (int k = 0; k < D; ++k) y[k] = xs[(size_t)k * npad + centre] + y[k] * scale;
	// The count only thins the proposals (no bit-exact contract with cneighbors.c here), so the
	// squared distance uses fused multiply-adds: 2D instead of 3D-1 FP64 instructions per pair.
	// Two chunks of 32 members are in flight per iteration, as in count_within_kernel.
	int cnt = 0;
	for (int base = 0; base < n; base += 64) {
		const int i0 = base + lane, i1 = base + 32 + lane;
		const bool valid0 = i0 < n, valid1 = i1 < n;
		double x0[D], x1[D];
#pragma unroll
		for (int k = 0; k < D; ++k) {
			x0[k] = valid0 ? xs[(size_t)k * npad + i0] : 0.0;
			x1[k] = valid1 ? xs[(size_t)k * npad + i1] : 0.0;
		}
		double d0 = 0.0, d1 = 0.0;
#pragma unroll
		for (int k = 0; k < D; ++k) {
			const double t0 = x0[k] - y[k], t1 = x1[k] - y[k];
			d0 = __fma_rn(t0, t0, d0);
			d1 = __fma_rn(t1, t1, d1);
		}
		cnt += __popc(__ballot_sync(0xffffffffu, valid0 && d0 < T));
		cnt += __popc(__ballot_sync(0xffffffffu, valid1 && d1 < T));
	}
	if (lane == 0) {
#pragma unroll
		for (int k = 0; k < D; ++k) points[(size_t)j * D + k] = y[k];
		// the centre itself is within r, so cnt >= 1 up to rounding at the surface
		keep[j] = (cnt <= 1 || coin * cnt < 1.0) ? 1 : 0;
		nnear_out[j] = cnt;
	}
}

__global__ void __launch_bounds__(256) gather_points_kernel(const double *__restrict__ points, int D,
                                                            const int *__restrict__ idx, int n,
                                                            double *__restrict__ out)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= n * D) return;
	out[i] = points[(size_t)idx[i / D] * D + i % D];
}

int launch_region_generate(const double *xs, int n, int npad, int ndim, double r, double T,
                           unsigned long long seed, unsigned long long first, int m, double *points,
                           uint8_t *keep, int *nnear, cudaStream_t st)
{
	if (m <= 0) return MDNS_OK;
	const int blocks = ceil_div(m, NB_THREADS / 32);
#define MDNS_GEN_CASE(D)                                                                       \
	case D:                                                                                \
		region_generate_kernel<D><<<blocks, NB_THREADS, 0, st>>>(xs, n, npad, r, T, seed, first, m, \
		                                                        points, keep, nnear);         \
		break
	switch (ndim) {
		MDNS_GEN_CASE(1);
		MDNS_GEN_CASE(2);
		MDNS_GEN_CASE(3);
		MDNS_GEN_CASE(4);
		MDNS_GEN_CASE(5);
		MDNS_GEN_CASE(6);
		MDNS_GEN_CASE(7);
		MDNS_GEN_CASE(8);
	default:
		set_error("device candidate generation supports 1..8 dimensions, not %d", ndim);
		return MDNS_EINVAL;
	}
#undef MDNS_GEN_CASE
	MDNS_LAUNCHED("region_generate_kernel");
	return MDNS_OK;
}

int launch_gather_points(const double *points, int ndim, const int *idx, int n, double *out,
                         cudaStream_t st)
{
	if (n <= 0) return MDNS_OK;
	gather_points_kernel<<<ceil_div((long long)n * ndim, 256), 256, 0, st>>>(points, ndim, idx, n, out);
	MDNS_LAUNCHED_HELPER("gather_points_kernel");
	return MDNS_OK;
}

// ------------------------------------------- per-axis ("SupFriends") distance ---
// clustering/neighbors.py:22-73.  Two data-parallel pieces; the sequential growth of the per-axis
// box (neighbors.py:44-58) stays on the host, which only visits the uncovered points.

// neighbors.py:24-25: index of every member's nearest other member (euclidean; the reference
// sorts the cdist row and takes the second entry).  One warp per member, generic in the
// dimension; ties go to the smaller index.
__global__ void __launch_bounds__(NB_THREADS) nn_index_kernel(const double *__restrict__ xs, int n,
                                                              int npad, int ndim,
                                                              int *__restrict__ nearest)
{
	const int lane = threadIdx.x & 31;
	const long long i = ((long long)blockIdx.x * NB_THREADS + threadIdx.x) >> 5;
	if (i >= n) return;
	double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
	int arg = -1;
	for (int base = 0; base < n; base += 32) {
		const int j = base + lane;
		if (j < n && j != i) {
			double t = __dsub_rn(xs[j], xs[i]);
			double d = __dmul_rn(t, t);
			for (int k = 1; k < ndim; ++k) {
				t = __dsub_rn(xs[(size_t)k * npad + j], xs[(size_t)k * npad + i]);
				d = __dadd_rn(d, __dmul_rn(t, t));
			}
			if (d < best) {       // j increases within a lane: the first minimum is kept
				best = d;
				arg = j;
			}
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const double ob = __shfl_xor_sync(0xffffffffu, best, o);
		const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
		if (oa >= 0 && (arg < 0 || ob < best || (ob == best && oa < arg))) {
			best = ob;
			arg = oa;
		}
	}
	if (lane == 0) nearest[i] = arg;
}

// neighbors.py:40-43: covered[q] = any listed reference member j with |x_qk - x_jk| < md_k on
// every axis k.  One warp per listed query, ballot early exit.
__global__ void __launch_bounds__(NB_THREADS) axis_covered_kernel(
    const double *__restrict__ xs, int npad, int ndim, const double *__restrict__ md,
    const int *__restrict__ query, int nq, const int *__restrict__ ref, int nr,
    uint8_t *__restrict__ covered)
{
	const int lane = threadIdx.x & 31;
	const long long q = ((long long)blockIdx.x * NB_THREADS + threadIdx.x) >> 5;
	if (q >= nq) return;
	const int i = query[q];
	bool found = false;
	for (int base = 0; base < nr; base += 32) {
		const int r = base + lane;
		bool close = r < nr;
		if (close) {
			const int j = ref[r];
			for (int k = 0; k < ndim; ++k) {
				const double dist = fabs(__dsub_rn(xs[(size_t)k * npad + i], xs[(size_t)k * npad + j]));
				close = close && dist < md[k];
			}
		}
		if (__ballot_sync(0xffffffffu, close) != 0u) {
			found = true;
			break;
		}
	}
	if (lane == 0) covered[q] = found ? 1 : 0;
}

int launch_nn_index(const double *xs, int n, int npad, int ndim, int *nearest, cudaStream_t st)
{
	if (n <= 0) return MDNS_OK;
	nn_index_kernel<<<ceil_div(n, NB_THREADS / 32), NB_THREADS, 0, st>>>(xs, n, npad, ndim, nearest);
	MDNS_LAUNCHED("nn_index_kernel");
	return MDNS_OK;
}

int launch_axis_covered(const double *xs, int npad, int ndim, const double *md, const int *query,
                        int nq, const int *ref, int nr, uint8_t *covered, cudaStream_t st)
{
	if (nq <= 0) return MDNS_OK;
	axis_covered_kernel<<<ceil_div(nq, NB_THREADS / 32), NB_THREADS, 0, st>>>(xs, npad, ndim, md, query,
	                                                                        nq, ref, nr, covered);
	MDNS_LAUNCHED("axis_covered_kernel");
	return MDNS_OK;
}

// ------------------------------------------------------- single-point test ---
// cneighbors.c:77-92: members are spread over the whole grid; any hit raises the flag.
__global__ void __launch_bounds__(NB_THREADS) within_single_kernel(const double *__restrict__ xs,
                                                                   int n, int npad, int ndim,
                                                                   const double *__restrict__ y,
                                                                   double T, int *__restrict__ flag)
{
	bool hit = false;
	for (long long i = (long long)blockIdx.x * NB_THREADS + threadIdx.x; i < n;
	     i += (long long)gridDim.x * NB_THREADS) {
		double t = __dsub_rn(xs[i], __ldg(y));
		double d = __dmul_rn(t, t);
		for (int k = 1; k < ndim; ++k) {
			t = __dsub_rn(xs[(size_t)k * npad + i], __ldg(y + k));
			d = __dadd_rn(d, __dmul_rn(t, t));
		}
		hit = hit || d < T;
	}
	if (__ballot_sync(0xffffffffu, hit) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

int launch_within_single(const double *xs, int n, int npad, int ndim, const double *y, double T,
                         int *flag, cudaStream_t st)
{
	if (n <= 0) return MDNS_OK;
	int blocks = ceil_div(n, NB_THREADS);
	if (blocks > 1024) blocks = 1024;
	within_single_kernel<<<blocks, NB_THREADS, 0, st>>>(xs, n, npad, ndim, y, T, flag);
	MDNS_LAUNCHED("within_single_kernel");
	return MDNS_OK;
}

// ------------------------------------------------- bootstrap: round lists ---
// One CTA per round: ordered compaction of the un-chosen ("query") and chosen
// ("reference") sample indices of chosen[i*nboot + b] (cneighbors.c:146,150).
__global__ void __launch_bounds__(1024) bootstrap_lists_kernel(const double *__restrict__ chosen,
                                                               int n, int nboot,
                                                               int *__restrict__ qidx,
                                                               int *__restrict__ ridx,
                                                               int *__restrict__ counts,
                                                               unsigned long long *__restrict__ nearest)
{
	__shared__ int warp_sums[32];
	const int b = blockIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int carry_r = 0;
	for (int base = 0; base < n; base += 1024) {
		const int i = base + threadIdx.x;
		const bool in = i < n;
		const int isref = (in && chosen[(size_t)i * nboot + b] != 0.0) ? 1 : 0;
		int x = isref;
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
		const int rpos = carry_r + before + x - isref;   // references before i
		if (in) {
			if (isref)
				ridx[(size_t)b * n + rpos] = i;
			else
				qidx[(size_t)b * n + (i - rpos)] = i;    // queries before i = i - refs before i
			nearest[(size_t)b * n + i] = BITS_1E300;
		}
		carry_r += all;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		counts[b] = n - carry_r;
		counts[nboot + b] = carry_r;
	}
}

int launch_bootstrap_lists(const double *chosen, int n, int nboot, int *qidx, int *ridx,
                           int *counts, unsigned long long *nearest, cudaStream_t st)
{
	bootstrap_lists_kernel<<<nboot, 1024, 0, st>>>(chosen, n, nboot, qidx, ridx, counts, nearest);
	MDNS_LAUNCHED("bootstrap_lists_kernel");
	return MDNS_OK;
}

__global__ void all_pairs_lists_kernel(int n, int *__restrict__ qidx, int *__restrict__ ridx,
                                       int *__restrict__ counts,
                                       unsigned long long *__restrict__ nearest)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) {
		qidx[i] = i;
		ridx[i] = i;
		nearest[i] = BITS_1E300;
	}
	if (i == 0) {
		counts[0] = n;
		counts[1] = n;
	}
}

int launch_all_pairs_lists(int n, int *qidx, int *ridx, int *counts, unsigned long long *nearest,
                           cudaStream_t st)
{
	all_pairs_lists_kernel<<<ceil_div(n, 256), 256, 0, st>>>(n, qidx, ridx, counts, nearest);
	MDNS_LAUNCHED("all_pairs_lists_kernel");
	return MDNS_OK;
}

// ------------------------------------------- nearest reference per query ---
// grid = (query tiles, rounds, reference splits).  Each thread owns one query sample of its
// round; reference samples are staged through shared memory in tiles and broadcast.  The
// partial minimum of each reference split is merged with atomicMin on the bit pattern
// (non-negative doubles order like unsigned integers; NaN patterns sort above 1e300 and so
// are ignored exactly like the `d < nearest_d` test of cneighbors.c:58,157).
constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 256;

template <int D>
__global__ void __launch_bounds__(NN_THREADS) nn_min_kernel(const double *__restrict__ xs, int n,
                                                            int npad, const int *__restrict__ qidx,
                                                            const int *__restrict__ ridx,
                                                            const int *__restrict__ counts,
                                                            int nrounds, int exclude_self,
                                                            unsigned long long *__restrict__ nearest)
{
	__shared__ double tile[D][NN_TILE];
	__shared__ int tile_idx[NN_TILE];
	const int b = blockIdx.y;
	const int nq = counts[b], nr = counts[nrounds + b];
	const int q = blockIdx.x * NN_THREADS + threadIdx.x;
	if (blockIdx.x * NN_THREADS >= nq) return;   // whole CTA idle for this round
	const bool valid = q < nq;
	const int qi = valid ? qidx[(size_t)b * n + q] : 0;
	double x[D];
#pragma unroll
	for (int k = 0; k < D; ++k) x[k] = valid ? xs[(size_t)k * npad + qi] : 0.0;

	// this CTA's slice of the reference list
	const int per = (nr + gridDim.z - 1) / gridDim.z;
	const int r_begin = blockIdx.z * per;
	const int r_end = min(nr, r_begin + per);
	double best = 1e300;
	for (int t0 = r_begin; t0 < r_end; t0 += NN_TILE) {
		const int tn = min(NN_TILE, r_end - t0);
		__syncthreads();
		for (int s = threadIdx.x; s < tn; s += NN_THREADS) {
			const int ri = ridx[(size_t)b * n + t0 + s];
			tile_idx[s] = ri;
#pragma unroll
			for (int k = 0; k < D; ++k) tile[k][s] = xs[(size_t)k * npad + ri];
		}
		__syncthreads();
		if (valid) {
			for (int s = 0; s < tn; ++s) {
				double y[D];
#pragma unroll
				for (int k = 0; k < D; ++k) y[k] = tile[k][s];
				const double d = sqdist_reg<D>(x, y);
				if (exclude_self && tile_idx[s] == qi) continue;
				if (d < best) best = d;
			}
		}
	}
	if (valid && best < 1e300)
		atomicMin(nearest + (size_t)b * n + qi, (unsigned long long)__double_as_longlong(best));
}

__global__ void __launch_bounds__(NN_THREADS) nn_min_generic_kernel(
    const double *__restrict__ xs, int n, int npad, int ndim, const int *__restrict__ qidx,
    const int *__restrict__ ridx, const int *__restrict__ counts, int nrounds, int exclude_self,
    unsigned long long *__restrict__ nearest)
{
	const int b = blockIdx.y;
	const int nq = counts[b], nr = counts[nrounds + b];
	const int q = blockIdx.x * NN_THREADS + threadIdx.x;
	if (q >= nq) return;
	const int qi = qidx[(size_t)b * n + q];
	const int per = (nr + gridDim.z - 1) / gridDim.z;
	const int r_begin = blockIdx.z * per;
	const int r_end = min(nr, r_begin + per);
	double best = 1e300;
	for (int s = r_begin; s < r_end; ++s) {
		const int ri = ridx[(size_t)b * n + s];
		if (exclude_self && ri == qi) continue;
		double t = __dsub_rn(xs[qi], xs[ri]);
		double d = __dmul_rn(t, t);
		for (int k = 1; k < ndim; ++k) {
			t = __dsub_rn(xs[(size_t)k * npad + qi], xs[(size_t)k * npad + ri]);
			d = __dadd_rn(d, __dmul_rn(t, t));
		}
		if (d < best) best = d;
	}
	if (best < 1e300)
		atomicMin(nearest + (size_t)b * n + qi, (unsigned long long)__double_as_longlong(best));
}

template <int D>
static int launch_nn_d(const double *xs, int n, int npad, const int *qidx, const int *ridx,
                       const int *counts, int nrounds, int exclude_self,
                       unsigned long long *nearest, dim3 grid, cudaStream_t st)
{
	nn_min_kernel<D><<<grid, NN_THREADS, 0, st>>>(xs, n, npad, qidx, ridx, counts, nrounds,
	                                             exclude_self, nearest);
	MDNS_LAUNCHED("nn_min_kernel");
	return MDNS_OK;
}

int launch_nn_min(const double *xs, int n, int npad, int ndim, const int *qidx, const int *ridx,
                  const int *counts, int nrounds, int exclude_self, unsigned long long *nearest,
                  int sm_count, cudaStream_t st)
{
	if (n <= 0 || nrounds <= 0) return MDNS_OK;
	const int qtiles = ceil_div(n, NN_THREADS);   // upper bound; idle CTAs exit at once
	// split the reference list so that the grid covers the chip a few times over
	long long ctas = (long long)qtiles * nrounds;
	int split = 1;
	const int max_split = ceil_div(n, NN_TILE);
	while (ctas * split < (long long)sm_count * 8 && split < max_split) split *= 2;
	if (split > max_split) split = max_split;
	if (split < 1) split = 1;
	dim3 grid(qtiles, nrounds, split);
#define MDNS_NN_CASE(D) \
	case D: return launch_nn_d<D>(xs, n, npad, qidx, ridx, counts, nrounds, exclude_self, nearest, grid, st)
	switch (ndim) {
		MDNS_NN_CASE(1);
		MDNS_NN_CASE(2);
		MDNS_NN_CASE(3);
		MDNS_NN_CASE(4);
		MDNS_NN_CASE(5);
		MDNS_NN_CASE(6);
		MDNS_NN_CASE(7);
		MDNS_NN_CASE(8);
	default:
		nn_min_generic_kernel<<<grid, NN_THREADS, 0, st>>>(xs, n, npad, ndim, qidx, ridx, counts,
		                                                  nrounds, exclude_self, nearest);
		MDNS_LAUNCHED("nn_min_generic_kernel");
		return MDNS_OK;
	}
#undef MDNS_NN_CASE
}

// ---------------------------------------------------------------- finalize ---
// max over rounds b and query samples of nearest (bit patterns order like the doubles).
// skip_first drops original sample 0 (cneighbors.c:162).  A round without queries
// contributes 0 (cneighbors.c:142), so the running maximum starts at +0.0.
__global__ void __launch_bounds__(1024) nn_finalize_kernel(const int *__restrict__ qidx,
                                                           const int *__restrict__ counts, int n,
                                                           int nrounds, int skip_first,
                                                           const unsigned long long *__restrict__ nearest,
                                                           double *__restrict__ result)
{
	__shared__ unsigned long long warp_max[32];
	unsigned long long best = 0ULL;
	for (int b = 0; b < nrounds; ++b) {
		const int nq = counts[b];
		for (int q = threadIdx.x; q < nq; q += 1024) {
			const int qi = qidx[(size_t)b * n + q];
			if (skip_first && qi == 0) continue;
			const unsigned long long v = nearest[(size_t)b * n + qi];
			best = v > best ? v : best;
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const unsigned long long y = __shfl_xor_sync(0xffffffffu, best, o);
		best = y > best ? y : best;
	}
	if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = best;
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 0; w < 32; ++w) best = warp_max[w] > best ? warp_max[w] : best;
		*result = __longlong_as_double((long long)best);
	}
}

int launch_nn_finalize(const int *qidx, const int *counts, int n, int nrounds, int skip_first,
                       const unsigned long long *nearest, double *result, cudaStream_t st)
{
	nn_finalize_kernel<<<1, 1024, 0, st>>>(qidx, counts, n, nrounds, skip_first, nearest, result);
	MDNS_LAUNCHED("nn_finalize_kernel");
	return MDNS_OK;
}

}  // namespace mdns
