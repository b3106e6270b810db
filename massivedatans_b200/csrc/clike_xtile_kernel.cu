// clike_xtile_kernel.cu -- candidate-batch chi-square in the expanded form, register-blocked.
//
//     sum_j (m_kj - y_ij)^2 = Syy_i - 2 * Sym_ik + Smm_k
//
// with Syy_i = sum_j y_ij^2 resident per data set (computed once at upload), Smm_k = sum_j m_kj^2
// per candidate (one tiny launch per batch) and the cross term Sym_ik = sum_j m_kj * y_ij computed
// here: ONE FP64 FMA per (element, candidate) where the direct form of clike.c:64-76 needs a
// subtraction and an FMA.  The direct-form tile kernel (clike_tile_kernel.cu) runs into the FP64
// pipe at K ~ 8 candidates per pass of the data (ncu: FP64 pipe 75 % at K = 16, 0.54 ms against
// an HBM floor of 0.26 ms); this kernel halves the FP64 work.
//
// Structure (all-active rows only, like clike_tile_kernel):
//   * one producer thread streams [256 data sets] x [16 channels] boxes of the resident row
//     matrix into a shared-memory ring with tiled tensor-TMA copies (UTMALDG, 128-byte swizzle,
//     mbarrier transaction counts) and, once per CTA, the KT model spectra of this pass with a
//     bulk-TMA copy;
//   * every consumer lane owns R data sets of the tile (rows r, r + 256/R, ...): a model value
//     pair is fetched from shared memory ONCE per warp (all lanes read the same address:
//     broadcast, one wavefront) and feeds 2*R FMAs.  A first version read the model through
//     the constant bank; ncu showed the issue slots split 1:1 between LDCU and DFMA (FP64 pipe
//     40 %), and with R > 1 ptxas fell back to per-lane LDC, which was slower still;
//   * KT*R accumulators per lane, no cross-lane reduction, coalesced logL stores.
//
// Accuracy.  Everything is FP64, but the three sums cancel when a candidate fits high
// signal-to-noise data.  With sequential FP64 summation over C channels the absolute error of
// Syy - 2 Sym + Smm is bounded by about (2C+4) * 2^-53 * (Syy + Smm), so a result is kept only
// when that bound is below the tolerance relative to the result itself (a.xp_guard =
// (2C+4) * 2^-53 / xp_tol; keep iff chi2 >= xp_guard * (Syy + Smm); default xp_tol 1e-10, the
// parity contract is 1e-9).  The data sets of every other pair are appended to a list and
// recomputed in the direct form by a small follow-up launch (xtile_fixup_kernel; rare: counted in
// a.xp_redo[0], and data that needs it for more than 2 % of its rows is sent back to the direct
// kernel by the host).  Keeping the fix-up out of line keeps it out of the main kernel's
// register budget (135 instead of 167 registers at KT = 16).
#include <cuda.h>

#include "kernels.cuh"

namespace mdns {

constexpr int XT_ROWS = 256;                    // data sets per tile = rows of one TMA box
constexpr int XT_BOX_CH = 16;                   // channels per box row = 128 bytes (swizzle span)
constexpr int XT_STAGE_BYTES = XT_ROWS * XT_BOX_CH * 8;
constexpr int XT_MAX_COUNTERS = 1024;           // ints behind a.xp_redo: total + one per pass
constexpr int slab_counter_base_dev = 1024, slab_counter_count_dev = 512;   // = slab_dmma_kernel.cu
#ifndef XT_UNROLL
#define XT_UNROLL 2
#endif
constexpr int XT_UNROLL_PAIRS = XT_UNROLL;   // channel pairs unrolled in the inner loop (registers)

__device__ __forceinline__ void xt_mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void xt_tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0,
                                               int c1, uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	    : "memory");
}

// Direct-form recomputation of the data sets the expanded form could not vouch for in this
// pass: one warp per listed data set, lanes across channels, candidates k0 .. k0+kt_valid-1.
// counter[0] accumulates the number of recomputed rows (host feedback), counter[1 + pass] is
// the length of this pass's list.
__global__ void __launch_bounds__(256) xtile_fixup_kernel(const LikeArgs a, const int k0,
                                                          const int kt_valid, const int pass)
{
	// launched as a programmatic dependent of the likelihood kernel: set up while that one is
	// still running, and nothing is touched before it has completed
	pdl_trigger();
	pdl_wait();
	// (slab_dmma_kernel hands out its slabs through a counter behind the list lengths: back to
	// zero for the next launch of this pass)
	if (blockIdx.x == 0 && threadIdx.x == 0 && pass < slab_counter_count_dev) a.xp_redo[slab_counter_base_dev + pass] = 0;
	const int n = a.xp_redo[1 + pass];
	if (n == 0) return;
	const int lane = threadIdx.x & 31;
	const int warps = gridDim.x * 8;
	const double inv = a.scale / a.noise2;
	for (int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += warps) {
		const int gr = a.xp_list[e];               // (compacted) row within this launch
		const double *yrow = a.Y + (long long)(a.active ? a.active[gr] : gr) * a.pitch;
		for (int k = 0; k < kt_valid; ++k) {
			const double *m = a.model + (size_t)(k0 + k) * a.mpitch;
			double s = 0.0;
			for (int j = lane; j < a.nx; j += 32) {
				const double d = m[j] - yrow[j];
				s = fma(d, d, s);
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
			if (lane == 0) {
				const double val = s * inv;
				if (a.out) a.out[(long long)(k0 + k) * a.out_stride + gr] = val;
				// the fused accept test of rows_dmma_kernel skipped this pair
				if (a.lmins && a.counts && val > a.lmins[gr]) atomicAdd(a.counts + k0 + k, 1);
			}
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.xp_redo, n);
}

template <int KT, int R, int STAGES>
__global__ void __launch_bounds__(XT_ROWS / R + 32) clike_xtile_kernel(
    const __grid_constant__ CUtensorMap tmap, const LikeArgs a, const int k0, const int kt_valid,
    const int pass)
{
	constexpr int LANE_ROWS = XT_ROWS / R;             // distance between the rows of one lane
	constexpr int CONSUMER_WARPS = LANE_ROWS / 32;
	// two partial sums per pair (even / odd channels) while the accumulators fit comfortably
	constexpr int NACC = KT * R <= 16 ? 2 : 1;
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], model_bar;
	// the swizzle pattern is a function of the shared-memory address: align the ring to 1 KB
	unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	const double *sm_model = reinterpret_cast<const double *>(ring + STAGES * XT_STAGE_BYTES);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int pitch_even = (int)a.pitch;       // channels incl. the zero pad of an odd count
	const int nchunks = (pitch_even + XT_BOX_CH - 1) / XT_BOX_CH;
	const int ntiles = (a.n_rows + XT_ROWS - 1) / XT_ROWS;

	if (threadIdx.x == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], CONSUMER_WARPS);
		}
		mbar_init(&model_bar, 1);
		mbar_fence_init();
	}
	__syncthreads();

	if (warp == CONSUMER_WARPS) {
		// ===================== producer (one elected thread) =====================
		if (lane == 0) {
			// the KT model spectra of this pass: rows k0 .. k0+KT-1 of the padded buffer
			const uint32_t mbytes = (uint32_t)KT * (uint32_t)a.mpitch * 8u;
			mbar_expect_tx(&model_bar, mbytes);
			tma_load_1d(const_cast<double *>(sm_model), a.model + (size_t)k0 * a.mpitch, mbytes,
			            &model_bar);
			int it = 0;
			for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
				const int r0 = a.row0 + tile * XT_ROWS;
				for (int c = 0; c < nchunks; ++c, ++it) {
					const int stage = it % STAGES;
					const uint32_t round = (uint32_t)(it / STAGES);
					mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);   // first round passes
					// out-of-bounds parts of a box are zero-filled and still counted
					mbar_expect_tx(&full_bar[stage], XT_STAGE_BYTES);
					xt_tma_load_2d(ring + (size_t)stage * XT_STAGE_BYTES, &tmap, c * XT_BOX_CH, r0,
					               &full_bar[stage]);
				}
			}
		}
	} else {
		// ===================== consumer warps =====================
		const int r_local = warp * 32 + lane;
		const int sw = r_local & 7;                // 128-byte swizzle: chunk ^= row % 8
		const double inv = a.scale / a.noise2;
		mbar_wait(&model_bar, 0);
		int it = 0;
		for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
			double acc[R][NACC][KT];
#pragma unroll
			for (int i = 0; i < R; ++i)
#pragma unroll
				for (int k = 0; k < KT; ++k) acc[i][0][k] = acc[i][NACC - 1][k] = 0.0;
			for (int c = 0; c < nchunks; ++c, ++it) {
				const int stage = it % STAGES;
				const uint32_t round = (uint32_t)(it / STAGES);
				mbar_wait(&full_bar[stage], round & 1u);
				const unsigned char *rowb = ring + (size_t)stage * XT_STAGE_BYTES + r_local * 128;
				const int jbase = c * XT_BOX_CH;
				const int left = pitch_even - jbase;     // valid channels of this box (even)
				const double *mrow = sm_model + jbase;
				if (left >= XT_BOX_CH) {
#pragma unroll XT_UNROLL_PAIRS
					for (int u = 0; u < XT_BOX_CH / 2; ++u) {
						double2 y[R];
#pragma unroll
						for (int i = 0; i < R; ++i)
							y[i] = *reinterpret_cast<const double2 *>(rowb + i * (LANE_ROWS * 128) +
							                                          ((u ^ sw) << 4));
#pragma unroll
						for (int k = 0; k < KT; ++k) {
							const double2 m =
							    *reinterpret_cast<const double2 *>(mrow + k * a.mpitch + 2 * u);
#pragma unroll
							for (int i = 0; i < R; ++i) {
								acc[i][0][k] = fma(m.x, y[i].x, acc[i][0][k]);
								acc[i][NACC - 1][k] = fma(m.y, y[i].y, acc[i][NACC - 1][k]);
							}
						}
					}
				} else {
					for (int u = 0; u < left / 2; ++u) {
						double2 y[R];
#pragma unroll
						for (int i = 0; i < R; ++i)
							y[i] = *reinterpret_cast<const double2 *>(rowb + i * (LANE_ROWS * 128) +
							                                          ((u ^ sw) << 4));
#pragma unroll
						for (int k = 0; k < KT; ++k) {
							const double2 m =
							    *reinterpret_cast<const double2 *>(mrow + k * a.mpitch + 2 * u);
#pragma unroll
							for (int i = 0; i < R; ++i) {
								acc[i][0][k] = fma(m.x, y[i].x, acc[i][0][k]);
								acc[i][NACC - 1][k] = fma(m.y, y[i].y, acc[i][NACC - 1][k]);
							}
						}
					}
				}
				__syncwarp();
				if (lane == 0) xt_mbar_arrive(&empty_bar[stage]);   // this warp is done with the stage
			}
#pragma unroll
			for (int i = 0; i < R; ++i) {
				const long long gr = (long long)tile * XT_ROWS + i * LANE_ROWS + r_local;
				const bool live = gr < a.n_rows;
				const double syy = live ? __ldg(a.syy + a.row0 + gr) : 0.0;
				bool redo = false;
#pragma unroll
				for (int k = 0; k < KT; ++k) {
					const double sym = NACC == 2 ? acc[i][0][k] + acc[i][NACC - 1][k] : acc[i][0][k];
					const double smm = __ldg(a.smm + k0 + k);
					const double chi = syy + fma(-2.0, sym, smm);
					const bool ok = chi >= a.xp_guard * (syy + smm);   // false for NaN too
					if (live && k < kt_valid) {
						if (ok)
							a.out[(long long)(k0 + k) * a.out_stride + gr] = chi * inv;
						else
							redo = true;
					}
				}
				if (redo) a.xp_list[atomicAdd(a.xp_redo + 1 + pass, 1)] = (int)gr;
			}
		}
	}
}

// ---- host side ---------------------------------------------------------------------------
int launch_xtile_fixup(const LikeArgs &a, int k0, int kv, int pass, int sm_count, cudaStream_t st);

static size_t xtile_smem(int kt, int stages, int mpitch)
{
	return (size_t)stages * XT_STAGE_BYTES + 1024 + (size_t)kt * mpitch * 8;
}

template <int KT, int R, int STAGES>
static int launch_xtile_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	constexpr int THREADS = XT_ROWS / R + 32;
	const size_t smem = xtile_smem(KT, STAGES, a.mpitch);
	auto kern = clike_xtile_kernel<KT, R, STAGES>;
	MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
	if (occ < 1) {
		set_error("expanded tile kernel does not fit (%zu bytes of shared memory)", smem);
		return MDNS_EINVAL;
	}
	CUtensorMap tm;
	memcpy(&tm, a.tmap256, sizeof tm);
	const int ntiles = ceil_div(a.n_rows, XT_ROWS);
	long long gx = ntiles;
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	const int npass = ceil_div(a.K, KT);
	if (npass + 1 > XT_MAX_COUNTERS) {
		set_error("expanded tile kernel: %d passes exceed the counter block", npass);
		return MDNS_EINVAL;
	}
	// list lengths of this launch's passes (counter[0], the running total, is left alone)
	if (!a.xp_counters_clear)
		MDNS_CUDA(cudaMemsetAsync(a.xp_redo + 1, 0, (size_t)npass * sizeof(int), st));
	for (int k0 = 0, pass = 0; k0 < a.K; k0 += KT, ++pass) {
		const int kv = a.K - k0 < KT ? a.K - k0 : KT;
		kern<<<(unsigned)gx, THREADS, smem, st>>>(tm, a, k0, kv, pass);
		MDNS_LAUNCHED("clike_xtile_kernel");
		const int rc = launch_xtile_fixup(a, k0, kv, pass, sm_count, st);
		if (rc != MDNS_OK) return rc;
	}
	return MDNS_OK;
}

// the direct-form fix-up of one pass (shared with clike_dmma_kernel.cu)
int launch_xtile_fixup(const LikeArgs &a, int k0, int kv, int pass, int sm_count, cudaStream_t st)
{
	int fix_blocks = ceil_div(a.n_rows, 8);
	if (fix_blocks > 2 * sm_count) fix_blocks = 2 * sm_count;
	launch_pdl(xtile_fixup_kernel, dim3(fix_blocks), dim3(256), 0, st, true, a, k0, kv, pass);
	MDNS_LAUNCHED_HELPER("xtile_fixup_kernel");
	return MDNS_OK;
}

bool xtile_fits(const LikeArgs &a, int kt, int stages)
{
	return a.tmap256 && !a.active && a.syy && a.smm && a.xp_redo && a.xp_list &&
	       xtile_smem(kt, stages, a.mpitch) <= 220 * 1024 &&
	       (size_t)kt * a.mpitch * 8 < (1u << 20);   // mbarrier tx-count range
}

// kt in {8, 16, 32}; lane_rows in {2, 4}; stages in {2, 3}
int xtile_counter_capacity() { return XT_MAX_COUNTERS; }

int launch_clike_xtile(const LikeArgs &a, int kt, int lane_rows, int stages, int sm_count,
                       cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	if (!xtile_fits(a, kt, stages)) {
		set_error("expanded tile kernel: needs all-active rows, the resident row sums and %zu bytes "
		          "of shared memory", xtile_smem(kt, stages, a.mpitch));
		return MDNS_EINVAL;
	}
#define MDNS_XT(KK, RR, SS) \
	if (kt == KK && lane_rows == RR && stages == SS) return launch_xtile_inst<KK, RR, SS>(a, sm_count, st)
	MDNS_XT(8, 2, 2);
	MDNS_XT(8, 2, 3);
	MDNS_XT(8, 4, 2);
	MDNS_XT(8, 4, 3);
	MDNS_XT(16, 2, 2);
	MDNS_XT(16, 2, 3);
	MDNS_XT(16, 4, 2);
	MDNS_XT(16, 4, 3);
	MDNS_XT(32, 2, 2);
	MDNS_XT(32, 2, 3);
#undef MDNS_XT
	set_error("unsupported expanded tile-kernel shape kt=%d lane_rows=%d stages=%d", kt, lane_rows,
	          stages);
	return MDNS_EINVAL;
}

}  // namespace mdns
