// clike_dmma_kernel.cu -- candidate-batch chi-square, expanded form, cross term on the FP64
// tensor path (mma.sync m8n8k4 f64 -> SASS DMMA).
//
//     sum_j (m_kj - y_ij)^2 = Syy_i - 2 * Sym_ik + Smm_k ,   Sym = Y * M^T
//
// Same algebra, guard and fix-up as clike_xtile_kernel.cu (see there for the error bound); what
// changes is how the K x C x N contraction Sym is issued.  tools/dmma_peak.cu measured the DMMA
// path of the B200 at 18.5 T FMA/s -- the same rate as the FP64 FMA pipe, and the two do not add
// up (they share the unit) -- so the tensor path buys no FLOPs.  What it buys is operand
// traffic: one DMMA performs 8 x 8 x 4 FMAs from ONE double of Y and ONE double of M per lane,
// where the FMA form needs a shared-memory model fetch per 4 FMAs.  ncu on the FMA form at
// K = 16 showed the shared-memory pipe busier than the FP64 pipe (55 % vs 47 %, one warp per
// scheduler); here a warp tile of 32 data sets x 8*NC candidates needs 4 + NC 64-bit fragment
// loads per 4 channels for 4*NC DMMAs.
//
// tcgen05 has no FP64 kind, so this is the only tensor path that keeps the 1e-9 contract; the
// TF32/BF16 variants of the north star's split cannot (cancellation).
//
// Layout:
//   * producer thread: tensor-TMA ring; every stage holds a [256 data sets] x [16 channels] box
//     of the resident rows and the matching [KT candidates] x [16 channels] box of the model
//     batch (both 128-byte swizzled, one mbarrier transaction).  The model slice is re-fetched
//     from L2 per tile (KT*128 B beside 32 KB of data), which keeps the shared-memory footprint
//     independent of the channel count: a first version kept the whole model batch in shared
//     memory and fell to one CTA per SM at 1000 channels (0.67 of the HBM roofline at K = 16);
//   * 8 consumer warps, warp w owns data sets [32w, 32w+32) of the tile as 4 row tiles of 8.
//     Fragment (m8n8k4, f64): A[g][t] = Y[row(g)][4*ks + t], B[t][g] = M[8*nc + g][4*ks + t],
//     g = lane / 4, t = lane % 4.  The logical row g is mapped to the physical row
//     perm(g) = 0,2,4,6,1,3,5,7 so that each half-warp touches rows of equal parity: with the
//     128-byte swizzle (16-byte chunk index ^= row % 8) its 16 lanes then hit 8 distinct chunk
//     columns and the 64-bit loads are conflict free; the candidates of a B fragment are
//     permuted the same way;
//   * D[g][2t + {0,1}] accumulates over all channels in registers (2 doubles per tile).
#include <cuda.h>

#include "kernels.cuh"

namespace mdns {

constexpr int DM_ROWS = 256;                    // data sets per tile = rows of one TMA box
constexpr int DM_BOX_CH = 16;                   // channels per box row = 128 bytes (swizzle span)
constexpr int DM_STAGE_BYTES = DM_ROWS * DM_BOX_CH * 8;

__device__ __forceinline__ void dm_mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void dm_tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0,
                                               int c1, uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	    : "memory");
}

// four arbitrary rows of the 2-D tensor (tile::gather4, sm_100): columns [c0, c0+16) of rows
// r0..r3 land as four consecutive 128-byte rows, swizzled like a tiled box
__device__ __forceinline__ void dm_tma_gather4(void *smem_dst, const CUtensorMap *tmap, int c0,
                                               int r0, int r1, int r2, int r3, uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

// NC candidate tiles (of 8) and DM_MR row tiles (of 8 data sets) per consumer warp;
// DM_ROWS / (8 * DM_MR) consumer warps per CTA.
// GATHER: the active data sets are listed in a.active (masked batch); the 64 lanes of TWO
// producer warps fill a stage with 64 gather4 copies of four listed rows each (`tmap` then is
// the one-row-box descriptor of the resident matrix).  The gather is bound by the issue rate of
// these 512-byte copies per producer warp (one warp per CTA: 0.215 ms for 5e5 active rows at
// K = 8, twice that with one CTA per SM), hence the second warp.  Consumers do not change: the
// rows of a tile land in list order, results go to the compacted slots.
template <int NC, int STAGES, int DM_MR, bool GATHER>
__global__ void __launch_bounds__(DM_ROWS / (8 * DM_MR) * 32 + (GATHER ? 64 : 32)) clike_dmma_kernel(
    const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap mmap,
    const LikeArgs a, const int k0, const int kt_valid, const int pass)
{
	constexpr int KT = NC * 8;
	constexpr int DM_WARPS = DM_ROWS / (8 * DM_MR);             // consumer warps
	constexpr int MODEL_BYTES = KT * DM_BOX_CH * 8;             // model slice of one stage
	constexpr int STAGE_BYTES = DM_STAGE_BYTES + MODEL_BYTES;   // multiple of 1 KB
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
	// the swizzle pattern is a function of the shared-memory address: align the ring to 1 KB
	unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int pitch_even = (int)a.pitch;       // channels incl. the zero pad of an odd count
	const int nchunks = (pitch_even + DM_BOX_CH - 1) / DM_BOX_CH;
	const int ntiles = (a.n_rows + DM_ROWS - 1) / DM_ROWS;

	if (threadIdx.x == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], DM_WARPS);
		}
		mbar_fence_init();
	}
	__syncthreads();

	if (warp >= DM_WARPS) {
		// ===================== producer =====================
		if (GATHER) {
			const int pl = (warp - DM_WARPS) * 32 + lane;     // 0..63: one group of four rows each
			int it = 0;
			for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
				// this lane's four listed rows (clamped at the end of the list: the surplus rows
				// of the last tile are computed and dropped)
				int rows[4];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const long long r = (long long)tile * DM_ROWS + pl * 4 + j;
					rows[j] = a.active[r < a.n_rows ? r : a.n_rows - 1];
				}
				for (int c = 0; c < nchunks; ++c, ++it) {
					const int stage = it % STAGES;
					const uint32_t round = (uint32_t)(it / STAGES);
					unsigned char *dst = ring + (size_t)stage * STAGE_BYTES;
					if (pl == 0) {
						mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);   // first round passes
						mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
						dm_tma_load_2d(dst + DM_STAGE_BYTES, &mmap, c * DM_BOX_CH, k0, &full_bar[stage]);
					}
					// the two producer warps meet on named barrier 1 once the stage is free
					__syncwarp();
					asm volatile("bar.sync 1, 64;" ::: "memory");
					dm_tma_gather4(dst + pl * 512, &tmap, c * DM_BOX_CH, rows[0], rows[1], rows[2],
					               rows[3], &full_bar[stage]);
				}
			}
		} else if (lane == 0) {
			int it = 0;
			for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
				const int r0 = a.row0 + tile * DM_ROWS;
				for (int c = 0; c < nchunks; ++c, ++it) {
					const int stage = it % STAGES;
					const uint32_t round = (uint32_t)(it / STAGES);
					mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);   // first round passes
					// out-of-bounds parts of a box are zero-filled and still counted
					mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
					unsigned char *dst = ring + (size_t)stage * STAGE_BYTES;
					dm_tma_load_2d(dst, &tmap, c * DM_BOX_CH, r0, &full_bar[stage]);
					dm_tma_load_2d(dst + DM_STAGE_BYTES, &mmap, c * DM_BOX_CH, k0, &full_bar[stage]);
				}
			}
		}
	} else {
		// ===================== consumer warps =====================
		const int g = lane >> 2, t = lane & 3;
		const int pr = ((g & 3) << 1) | (g >> 2);      // physical row of logical row g
		// byte offset of this lane's A element inside a stage, per channel step ks: row * 128 +
		// ((2*ks + t/2) ^ pr) * 16 + (t & 1) * 8 ; the row tiles of the warp are 1 KB apart
		const int a_row_off = (warp * (8 * DM_MR) + pr) * 128 + (t & 1) * 8;
		const int a_chunk = t >> 1;
		// B element: candidate perm(g) of the tile (same permutation, same swizzle), 128 B per row
		const int b_row_off = DM_STAGE_BYTES + pr * 128 + (t & 1) * 8;
		const double inv = a.scale / a.noise2;
		int it = 0;
		for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
			double acc[DM_MR][NC][2];
#pragma unroll
			for (int mr = 0; mr < DM_MR; ++mr)
#pragma unroll
				for (int nc = 0; nc < NC; ++nc) acc[mr][nc][0] = acc[mr][nc][1] = 0.0;
			for (int c = 0; c < nchunks; ++c, ++it) {
				const int stage = it % STAGES;
				const uint32_t round = (uint32_t)(it / STAGES);
				mbar_wait(&full_bar[stage], round & 1u);
				const unsigned char *sbase = ring + (size_t)stage * STAGE_BYTES;
#pragma unroll
				for (int ks = 0; ks < DM_BOX_CH / 4; ++ks) {
					double fa[DM_MR], fb[NC];
					const int choff = ((2 * ks + a_chunk) ^ pr) << 4;
#pragma unroll
					for (int mr = 0; mr < DM_MR; ++mr)
						fa[mr] = *reinterpret_cast<const double *>(sbase + a_row_off + mr * 1024 + choff);
#pragma unroll
					for (int nc = 0; nc < NC; ++nc)
						fb[nc] = *reinterpret_cast<const double *>(sbase + b_row_off + nc * 1024 + choff);
#pragma unroll
					for (int mr = 0; mr < DM_MR; ++mr)
#pragma unroll
						for (int nc = 0; nc < NC; ++nc)
							dmma_8x8x4(acc[mr][nc][0], acc[mr][nc][1], fa[mr], fb[nc]);
				}
				__syncwarp();
				if (lane == 0) dm_mbar_arrive(&empty_bar[stage]);   // this warp is done with the stage
			}
			// epilogue: lane holds D[row(g)][2t + {0,1}] of every (row tile, candidate tile)
#pragma unroll
			for (int mr = 0; mr < DM_MR; ++mr) {
				const long long gr = (long long)tile * DM_ROWS + warp * (8 * DM_MR) + mr * 8 + pr;
				const bool live = gr < a.n_rows;
				const double syy = !live ? 0.0
				                   : GATHER ? __ldg(a.syy + a.active[gr])
				                            : __ldg(a.syy + a.row0 + gr);
				bool redo = false;
#pragma unroll
				for (int nc = 0; nc < NC; ++nc) {
#pragma unroll
					for (int i = 0; i < 2; ++i) {
						const int col = 2 * t + i;      // logical column -> physical candidate
						const int k = nc * 8 + (((col & 3) << 1) | (col >> 2));
						const double smm = __ldg(a.smm + k0 + k);
						const double chi = syy + fma(-2.0, acc[mr][nc][i], smm);
						const bool ok = chi >= a.xp_guard * (syy + smm);   // false for NaN too
						if (live && k < kt_valid) {
							if (ok)
								a.out[(long long)(k0 + k) * a.out_stride + gr] = chi * inv;
							else
								redo = true;
						}
					}
				}
				// the four lanes of a group share the data set: list it once
				// (no short-circuit: every lane must take part in both shuffles)
				int flag = redo ? 1 : 0;
				flag |= __shfl_xor_sync(0xffffffffu, flag, 1);
				flag |= __shfl_xor_sync(0xffffffffu, flag, 2);
				redo = flag != 0;
				if (redo && t == 0) a.xp_list[atomicAdd(a.xp_redo + 1 + pass, 1)] = (int)gr;
			}
		}
	}
}

// ---- host side ---------------------------------------------------------------------------
static size_t dmma_smem(int kt, int stages)
{
	return (size_t)stages * (DM_STAGE_BYTES + (size_t)kt * DM_BOX_CH * 8) + 1024;
}

int make_row_tensor_map_box(void *out, const double *Y, long long n_rows, long long pitch,
                            int box_rows);

int launch_xtile_fixup(const LikeArgs &a, int k0, int kv, int pass, int sm_count, cudaStream_t st);

template <int NC, int STAGES, int DM_MR, bool GATHER>
static int launch_dmma_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	constexpr int KT = NC * 8;
	constexpr int THREADS = DM_ROWS / (8 * DM_MR) * 32 + (GATHER ? 64 : 32);
	const size_t smem = dmma_smem(KT, STAGES);
	auto kern = clike_dmma_kernel<NC, STAGES, DM_MR, GATHER>;
	MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
	if (occ < 1) {
		set_error("DMMA tile kernel does not fit (%zu bytes of shared memory)", smem);
		return MDNS_EINVAL;
	}
	CUtensorMap tm, mm;
	memcpy(&tm, GATHER ? a.tmap_gather : a.tmap256, sizeof tm);
	// the model batch [Kpad][mpitch] as boxes of KT candidates x 16 channels (Kpad is a multiple
	// of 32, so a box never leaves the buffer)
	const long long kpad = (long long)round_up(a.K, KT_MAX);
	int rcm = make_row_tensor_map_box(&mm, a.model, kpad, a.mpitch, KT);
	if (rcm != MDNS_OK) return rcm;
	const int ntiles = ceil_div(a.n_rows, DM_ROWS);
	long long gx = ntiles;
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	const int npass = ceil_div(a.K, KT);
	if (npass + 1 > xtile_counter_capacity()) {
		set_error("DMMA tile kernel: %d passes exceed the counter block", npass);
		return MDNS_EINVAL;
	}
	// list lengths of this launch's passes (counter[0], the running total, is left alone)
	if (!a.xp_counters_clear)
		MDNS_CUDA(cudaMemsetAsync(a.xp_redo + 1, 0, (size_t)npass * sizeof(int), st));
	for (int k0 = 0, pass = 0; k0 < a.K; k0 += KT, ++pass) {
		const int kv = a.K - k0 < KT ? a.K - k0 : KT;
		kern<<<(unsigned)gx, THREADS, smem, st>>>(tm, mm, a, k0, kv, pass);
		MDNS_LAUNCHED(GATHER ? "clike_dmma_kernel(gather)" : "clike_dmma_kernel");
		const int rc = launch_xtile_fixup(a, k0, kv, pass, sm_count, st);
		if (rc != MDNS_OK) return rc;
	}
	return MDNS_OK;
}

bool dmma_fits(const LikeArgs &a, int kt, int stages)
{
	if (stages > 10) stages -= 10;
	return (a.active ? a.tmap_gather != nullptr : a.tmap256 != nullptr) && a.syy && a.smm &&
	       a.xp_redo && a.xp_list && dmma_smem(kt, stages) <= 220 * 1024;
}

// kt in {8, 16, 32}; stages in {2, 3, 4}, + 10 for 16 consumer warps of 16 data sets each
// instead of 8 warps of 32
int launch_clike_dmma(const LikeArgs &a, int kt, int stages, int sm_count, cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	if (!dmma_fits(a, kt, stages)) {
		set_error("DMMA tile kernel: needs the tensor maps, the resident row sums and %zu bytes of "
		          "shared memory", dmma_smem(kt, stages));
		return MDNS_EINVAL;
	}
#define MDNS_DM(KK, SS)                                                                              \
	if (kt == KK && stages == SS)                                                                \
		return a.active ? launch_dmma_inst<KK / 8, SS, 4, true>(a, sm_count, st)             \
		                : launch_dmma_inst<KK / 8, SS, 4, false>(a, sm_count, st);           \
	if (kt == KK && stages == SS + 10)                                                           \
		return a.active ? launch_dmma_inst<KK / 8, SS, 2, true>(a, sm_count, st)             \
		                : launch_dmma_inst<KK / 8, SS, 2, false>(a, sm_count, st)
	MDNS_DM(8, 2);
	MDNS_DM(8, 3);
	MDNS_DM(8, 4);
	MDNS_DM(16, 2);
	MDNS_DM(16, 3);
	MDNS_DM(16, 4);
	MDNS_DM(32, 2);
	MDNS_DM(32, 3);
	MDNS_DM(32, 4);
#undef MDNS_DM
	set_error("unsupported DMMA tile-kernel shape kt=%d stages=%d", kt, stages);
	return MDNS_EINVAL;
}

}  // namespace mdns
