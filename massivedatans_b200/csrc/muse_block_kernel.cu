// muse_block_kernel.cu -- MUSE scaled chi-square (cmuselike.c:48-64) for long spectra.
//
// cmuselike needs every data set twice: first s1 = sum y*m/v and s2 = 1e-10 + sum m^2/v
// (cmuselike.c:51-56), then chi = sum (y - s*m)^2/v with s = s1/s2 (:57-61).  A row of the
// MUSE cube (3600 channels, data + inverse variance = 57.6 KB) fits neither registers nor a
// sub-warp, and streaming it twice makes the second pass an L2 re-read (measured: 45 % of
// the HBM roofline).  Here one CTA owns one data set at a time:
//
//   * thread 0 keeps a ring of row slots (y row + w row) filled with two bulk-TMA copies per
//     data set (cp.async.bulk -> UBLKCP, mbarrier transaction counts): each row is fetched
//     from HBM exactly once, three rows are in flight per SM;
//   * the 512 threads copy their channel pairs of the slot into registers (NF fragments per
//     thread, rows up to 8192 channels) and release the slot at once, so the ring refills
//     while both passes run out of registers; KT candidate spectra are register-blocked per
//     pass and come from L1/L2 through the read-only path.  Longer rows (NF = 0) re-read
//     the slot from shared memory in both passes;
//   * block-wide FP64 reductions: warp butterfly + fixed-order sum of the warp partials,
//     so the result does not depend on scheduling;
//   * G = 2: the CTA is split into two groups of 256 threads that work on alternate data sets
//     of the CTA's list with their own named barriers.  With one data set at a time the two
//     passes and their two block-wide reductions form a latency chain of ~1.8 us per row, longer
//     than the 1.3 us the row takes to arrive at the SM's share of HBM bandwidth (measured on the
//     4223 x 3600 cube: 52 us = 0.71 of the roofline); two rows in flight hide it.
//
// The resident W holds 1/v (computed once at upload with a correctly rounded division), so
// the kernel multiplies where the reference divides: y*m/v -> y*(m*w).  Differences to the
// reference are rounding-level (asserted < 1e-12 relative; contract 1e-9).
#include "kernels.cuh"

namespace mdns {

constexpr int MB_THREADS = 512;
constexpr int MB_WARPS = MB_THREADS / 32;
constexpr int MB_STAGES = 3;
constexpr size_t MB_SMEM_LIMIT = 220 * 1024;

// named barrier of one thread group (barrier 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int group, int nthreads)
{
	asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(nthreads) : "memory");
}

// sum over the TG threads of one group; `red` is one of the group's two alternating buffers
template <int N, int NCOL, int TG>
__device__ __forceinline__ void group_sum(double (&v)[N], double (*red)[MB_WARPS][NCOL], int group,
                                          int gtid)
{
	constexpr int GW = TG / 32;
	const int warp = gtid >> 5, lane = gtid & 31;
#pragma unroll
	for (int i = 0; i < N; ++i) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v[i] += shfl_xor_f64(v[i], o);
	}
	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < N; ++i) (*red)[warp][i] = v[i];
	}
	group_sync(group, TG);
#pragma unroll
	for (int i = 0; i < N; ++i) {
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < GW; ++w) s += (*red)[w][i];
		v[i] = s;
	}
}

template <int KT, int NF, int G>
__global__ void __launch_bounds__(MB_THREADS, 1) muse_block_kernel(const LikeArgs a)
{
	constexpr int TG = MB_THREADS / G;            // threads per group
	extern __shared__ __align__(128) unsigned char smem_raw[];
	constexpr int PERIOD = (G == 1 || MB_STAGES % 2 == 0) ? MB_STAGES : 2 * MB_STAGES;
	static_assert(G == 1 || G == 2, "one or two data sets in flight");
	__shared__ uint64_t full_bar[G][MB_STAGES];
	__shared__ double red[G][2][MB_WARPS][2 * KT];
	const int nfrag = (a.nx + 1) >> 1;
	const int mfp = a.mpitch >> 1;
	const uint32_t row_bytes = (uint32_t)a.pitch * 8u;
	const size_t slot_bytes = 2 * (size_t)row_bytes;
	const int group = threadIdx.x / TG, gtid = threadIdx.x % TG;

	if (threadIdx.x == 0) {
#pragma unroll
		for (int s = 0; s < MB_STAGES; ++s)
#pragma unroll
			for (int g = 0; g < G; ++g) mbar_init(&full_bar[g][s], 1);
		mbar_fence_init();
	}
	__syncthreads();

	// rows of this CTA: r = blockIdx.x + it * gridDim.x; group g takes it = g, g + G, ...
	const long long nmine = a.n_rows > (long long)blockIdx.x
	                            ? (a.n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x
	                            : 0;
	auto issue = [&](long long it) {
		const long long r = (long long)blockIdx.x + it * gridDim.x;
		const long long row = a.active ? (long long)a.active[r] : r;
		const int slot = (int)(it % MB_STAGES);
		unsigned char *dst = smem_raw + slot * slot_bytes;
		uint64_t *bar = &full_bar[it % G][slot];      // the barrier of the group that will read it
		mbar_expect_tx(bar, 2 * row_bytes);
		tma_load_1d(dst, a.Y + row * a.pitch, row_bytes, bar);
		tma_load_1d(dst + row_bytes, a.W + row * a.pitch, row_bytes, bar);
	};
	if (threadIdx.x == 0) {
		for (long long it = 0; it < MB_STAGES && it < nmine; ++it) issue(it);
	}

	const double2 *model = reinterpret_cast<const double2 *>(a.model);
	int flip = 0;
	for (long long it = group; it < nmine; it += G) {
		const int slot = (int)(it % MB_STAGES);
		const long long r = (long long)blockIdx.x + it * gridDim.x;
		const long long row = a.active ? (long long)a.active[r] : r;
		// Every (group, slot) pair has its own barrier: with two groups a slot alternates
		// between them, and a parity wait is only sound for a waiter that consumed the
		// immediately preceding phase of the barrier it waits on (a shared per-slot barrier
		// let a group that was two phases behind pass on the phase in between: wrong rows,
		// then a launch failure).  Row `it` returns to the same (group, slot) every PERIOD rows.
		mbar_wait(&full_bar[group][slot], (uint32_t)((it / PERIOD) & 1));
		const double2 *sy = reinterpret_cast<const double2 *>(smem_raw + slot * slot_bytes);
		const double2 *sw = reinterpret_cast<const double2 *>(smem_raw + slot * slot_bytes + row_bytes);

		// NF > 0: the row moves to registers and the slot is recycled immediately
		constexpr int NR = NF > 0 ? NF : 1;
		double2 ry[NR], rw[NR];
		if (NF > 0) {
#pragma unroll
			for (int i = 0; i < NR; ++i) {
				const int f = gtid + i * TG;
				ry[i] = rw[i] = make_double2(0.0, 0.0);
				if (f < nfrag) {
					ry[i] = sy[f];
					rw[i] = sw[f];
				}
			}
			group_sync(group, TG);   // the whole group has its fragments: refill the slot
			if (gtid == 0 && it + MB_STAGES < nmine) issue(it + MB_STAGES);
		}
		const int niter = NF > 0 ? NF : (nfrag + TG - 1) / TG;

		// Batches (K > KT): every KT candidates cost two passes over the row's registers, two
		// block-wide reductions and 2 x KT model spectra streamed from L1/L2.  Tried: all first
		// passes of 16 candidates, one barrier, all second passes, one barrier (same bits) --
		// no faster (K = 4 on the 40 000 x 3600 cube 0.88 vs 0.84 ms): batches are bounded by
		// re-streaming the model spectra from L2 for every row pair (K x 28.8 KB x 2 passes,
		// ~11 TB/s), not by the barriers, and this order lets pass 2 find its spectra in L1.
		for (int k0 = 0; k0 < a.K; k0 += KT) {
			// ---- pass 1: s1 = sum y*m*w, s2 = sum m*m*w (cmuselike.c:51-56)
			double acc[2 * KT];
#pragma unroll
			for (int i = 0; i < 2 * KT; ++i) acc[i] = 0.0;
#pragma unroll
			for (int i = 0; i < niter; ++i) {
				const int f = gtid + i * TG;
				if (f < nfrag) {
					const double2 y = NF > 0 ? ry[NF > 0 ? i : 0] : sy[f];
					const double2 w = NF > 0 ? rw[NF > 0 ? i : 0] : sw[f];
#pragma unroll
					for (int k = 0; k < KT; ++k) {
						const double2 m = __ldg(model + (size_t)(k0 + k) * mfp + f);
						const double t0 = m.x * w.x, t1 = m.y * w.y;
						acc[2 * k] = fma(y.x, t0, acc[2 * k]);
						acc[2 * k] = fma(y.y, t1, acc[2 * k]);
						acc[2 * k + 1] = fma(m.x, t0, acc[2 * k + 1]);
						acc[2 * k + 1] = fma(m.y, t1, acc[2 * k + 1]);
					}
				}
			}
			group_sum<2 * KT, 2 * KT, TG>(acc, &red[group][flip], group, gtid);
			flip ^= 1;
			double s[KT];
#pragma unroll
			for (int k = 0; k < KT; ++k) s[k] = acc[2 * k] / (acc[2 * k + 1] + 1e-10);
			// ---- pass 2: chi = sum (y - s*m)^2 * w (cmuselike.c:58-61)
			double chi[KT];
#pragma unroll
			for (int k = 0; k < KT; ++k) chi[k] = 0.0;
#pragma unroll
			for (int i = 0; i < niter; ++i) {
				const int f = gtid + i * TG;
				if (f < nfrag) {
					const double2 y = NF > 0 ? ry[NF > 0 ? i : 0] : sy[f];
					const double2 w = NF > 0 ? rw[NF > 0 ? i : 0] : sw[f];
#pragma unroll
					for (int k = 0; k < KT; ++k) {
						const double2 m = __ldg(model + (size_t)(k0 + k) * mfp + f);
						const double r0 = fma(-s[k], m.x, y.x);
						const double r1 = fma(-s[k], m.y, y.y);
						chi[k] = fma(r0 * r0, w.x, chi[k]);
						chi[k] = fma(r1 * r1, w.y, chi[k]);
					}
				}
			}
			group_sum<KT, 2 * KT, TG>(chi, &red[group][flip], group, gtid);
			flip ^= 1;
			if (gtid == 0) {
#pragma unroll
				for (int k = 0; k < KT; ++k)
					if (k0 + k < a.K) a.out[(long long)(k0 + k) * a.out_stride + row] = -0.5 * chi[k];
			}
		}
		if (NF == 0) {
			group_sync(group, TG);   // the group is done with the slot: refill it
			if (gtid == 0 && it + MB_STAGES < nmine) issue(it + MB_STAGES);
		}
	}
}

bool muse_block_fits(const LikeArgs &a)
{
	return a.W != nullptr && MB_STAGES * 2 * (size_t)a.pitch * 8 <= MB_SMEM_LIMIT &&
	       (size_t)a.pitch * 8 < (1u << 19);   // mbarrier tx-count range
}

template <int KT, int NF, int G>
static int launch_muse_block_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	const size_t smem = MB_STAGES * 2 * (size_t)a.pitch * 8;
	auto kern = muse_block_kernel<KT, NF, G>;
	MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, MB_THREADS, smem));
	if (occ < 1) {
		set_error("muse block kernel does not fit (%zu bytes of shared memory)", smem);
		return MDNS_EINVAL;
	}
	long long gx = a.n_rows;
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	kern<<<(unsigned)gx, MB_THREADS, smem, st>>>(a);
	MDNS_LAUNCHED("muse_block_kernel");
	return MDNS_OK;
}

template <int KT>
static int launch_muse_block_k(const LikeArgs &a, int groups, int sm_count, cudaStream_t st)
{
	const int nfrag = (a.nx + 1) >> 1;
	if (groups == 2) {
		// two data sets in flight per CTA, 256 threads each
		const int per_thread = ceil_div(nfrag, MB_THREADS / 2);
		if (per_thread <= 2) return launch_muse_block_inst<KT, 2, 2>(a, sm_count, st);
		if (per_thread <= 4) return launch_muse_block_inst<KT, 4, 2>(a, sm_count, st);
		if (per_thread <= 8) return launch_muse_block_inst<KT, 8, 2>(a, sm_count, st);
		return launch_muse_block_inst<KT, 0, 2>(a, sm_count, st);
	}
	const int per_thread = ceil_div(nfrag, MB_THREADS);
	if (per_thread <= 1) return launch_muse_block_inst<KT, 1, 1>(a, sm_count, st);
	if (per_thread <= 2) return launch_muse_block_inst<KT, 2, 1>(a, sm_count, st);
	if (per_thread <= 4) return launch_muse_block_inst<KT, 4, 1>(a, sm_count, st);
	if (per_thread <= 8) return launch_muse_block_inst<KT, 8, 1>(a, sm_count, st);
	return launch_muse_block_inst<KT, 0, 1>(a, sm_count, st);
}

int launch_muse_block(const LikeArgs &a, int ktile, int groups, int sm_count, cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	int kt = ktile;
	// 512 threads cap the kernel at 128 registers: four candidates per pass spill once the
	// row fragments live in registers, so the automatic choice stops at two
	if (kt != 1 && kt != 2 && kt != 4) kt = a.K >= 2 ? 2 : 1;
	if (groups != 1 && groups != 2) groups = 2;
	switch (kt) {
	case 1: return launch_muse_block_k<1>(a, groups, sm_count, st);
	case 2: return launch_muse_block_k<2>(a, groups, sm_count, st);
	default: return launch_muse_block_k<4>(a, groups, sm_count, st);
	}
}

}  // namespace mdns
