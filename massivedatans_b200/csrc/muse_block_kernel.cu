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
//   * block-wide FP64 reductions: warp butterfly + fixed-order sum of the 8 warp partials,
//     so the result does not depend on scheduling.
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

template <int N, int NCOL>
__device__ __forceinline__ void block_sum(double (&v)[N], double (*red)[MB_WARPS][NCOL])
{
	// red points at one of two alternating scratch buffers (see caller)
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
	for (int i = 0; i < N; ++i) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v[i] += shfl_xor_f64(v[i], o);
	}
	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < N; ++i) (*red)[warp][i] = v[i];
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < N; ++i) {
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < MB_WARPS; ++w) s += (*red)[w][i];
		v[i] = s;
	}
}

template <int KT, int NF>
__global__ void __launch_bounds__(MB_THREADS, 1) muse_block_kernel(const LikeArgs a)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ uint64_t full_bar[MB_STAGES];
	__shared__ double red[2][MB_WARPS][2 * KT];
	const int nfrag = (a.nx + 1) >> 1;
	const int mfp = a.mpitch >> 1;
	const uint32_t row_bytes = (uint32_t)a.pitch * 8u;
	const size_t slot_bytes = 2 * (size_t)row_bytes;

	if (threadIdx.x == 0) {
#pragma unroll
		for (int s = 0; s < MB_STAGES; ++s) mbar_init(&full_bar[s], 1);
		mbar_fence_init();
	}
	__syncthreads();

	// rows of this CTA: r = blockIdx.x + it * gridDim.x
	const long long nmine = a.n_rows > (long long)blockIdx.x
	                            ? (a.n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x
	                            : 0;
	auto issue = [&](long long it) {
		const long long r = (long long)blockIdx.x + it * gridDim.x;
		const long long row = a.active ? (long long)a.active[r] : r;
		const int slot = (int)(it % MB_STAGES);
		unsigned char *dst = smem_raw + slot * slot_bytes;
		mbar_expect_tx(&full_bar[slot], 2 * row_bytes);
		tma_load_1d(dst, a.Y + row * a.pitch, row_bytes, &full_bar[slot]);
		tma_load_1d(dst + row_bytes, a.W + row * a.pitch, row_bytes, &full_bar[slot]);
	};
	if (threadIdx.x == 0) {
		for (long long it = 0; it < MB_STAGES && it < nmine; ++it) issue(it);
	}

	const double2 *model = reinterpret_cast<const double2 *>(a.model);
	int flip = 0;
	for (long long it = 0; it < nmine; ++it) {
		const int slot = (int)(it % MB_STAGES);
		const long long r = (long long)blockIdx.x + it * gridDim.x;
		const long long row = a.active ? (long long)a.active[r] : r;
		mbar_wait(&full_bar[slot], (uint32_t)((it / MB_STAGES) & 1));
		const double2 *sy = reinterpret_cast<const double2 *>(smem_raw + slot * slot_bytes);
		const double2 *sw = reinterpret_cast<const double2 *>(smem_raw + slot * slot_bytes + row_bytes);

		// NF > 0: the row moves to registers and the slot is recycled immediately
		constexpr int NR = NF > 0 ? NF : 1;
		double2 ry[NR], rw[NR];
		if (NF > 0) {
#pragma unroll
			for (int i = 0; i < NR; ++i) {
				const int f = threadIdx.x + i * MB_THREADS;
				ry[i] = rw[i] = make_double2(0.0, 0.0);
				if (f < nfrag) {
					ry[i] = sy[f];
					rw[i] = sw[f];
				}
			}
			__syncthreads();   // everybody has its fragments: refill the slot
			if (threadIdx.x == 0 && it + MB_STAGES < nmine) issue(it + MB_STAGES);
		}
		const int niter = NF > 0 ? NF : (nfrag + MB_THREADS - 1) / MB_THREADS;

		for (int k0 = 0; k0 < a.K; k0 += KT) {
			// ---- pass 1: s1 = sum y*m*w, s2 = sum m*m*w (cmuselike.c:51-56)
			double acc[2 * KT];
#pragma unroll
			for (int i = 0; i < 2 * KT; ++i) acc[i] = 0.0;
#pragma unroll
			for (int i = 0; i < niter; ++i) {
				const int f = threadIdx.x + i * MB_THREADS;
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
			block_sum<2 * KT, 2 * KT>(acc, &red[flip]);
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
				const int f = threadIdx.x + i * MB_THREADS;
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
			block_sum<KT, 2 * KT>(chi, &red[flip]);
			flip ^= 1;
			if (threadIdx.x == 0) {
#pragma unroll
				for (int k = 0; k < KT; ++k)
					if (k0 + k < a.K) a.out[(long long)(k0 + k) * a.out_stride + row] = -0.5 * chi[k];
			}
		}
		if (NF == 0) {
			__syncthreads();   // everybody is done with the slot: refill it
			if (threadIdx.x == 0 && it + MB_STAGES < nmine) issue(it + MB_STAGES);
		}
	}
}

bool muse_block_fits(const LikeArgs &a)
{
	return a.W != nullptr && MB_STAGES * 2 * (size_t)a.pitch * 8 <= MB_SMEM_LIMIT &&
	       (size_t)a.pitch * 8 < (1u << 19);   // mbarrier tx-count range
}

template <int KT, int NF>
static int launch_muse_block_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	const size_t smem = MB_STAGES * 2 * (size_t)a.pitch * 8;
	auto kern = muse_block_kernel<KT, NF>;
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
static int launch_muse_block_k(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	const int nfrag = (a.nx + 1) >> 1;
	const int per_thread = ceil_div(nfrag, MB_THREADS);
	if (per_thread <= 1) return launch_muse_block_inst<KT, 1>(a, sm_count, st);
	if (per_thread <= 2) return launch_muse_block_inst<KT, 2>(a, sm_count, st);
	if (per_thread <= 4) return launch_muse_block_inst<KT, 4>(a, sm_count, st);
	if (per_thread <= 8) return launch_muse_block_inst<KT, 8>(a, sm_count, st);
	return launch_muse_block_inst<KT, 0>(a, sm_count, st);
}

int launch_muse_block(const LikeArgs &a, int ktile, int sm_count, cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	int kt = ktile;
	// 512 threads cap the kernel at 128 registers: four candidates per pass spill once the
	// row fragments live in registers, so the automatic choice stops at two
	if (kt != 1 && kt != 2 && kt != 4) kt = a.K >= 2 ? 2 : 1;
	switch (kt) {
	case 1: return launch_muse_block_k<1>(a, sm_count, st);
	case 2: return launch_muse_block_k<2>(a, sm_count, st);
	default: return launch_muse_block_k<4>(a, sm_count, st);
	}
}

}  // namespace mdns
