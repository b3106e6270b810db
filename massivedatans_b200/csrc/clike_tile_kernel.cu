// clike_tile_kernel.cu -- candidate-batch chi-square kernel: tensor-TMA ring + lane-per-data-set.
//
// For K >= ~4 candidates per pass the lanes-across-channels kernels (likelihood_kernels.cu)
// are held back by shared-memory bandwidth (every lane needs its own model value) and by
// exposed global-load latency (ncu: 42 % long-scoreboard stalls, FP64 pipe 53 % busy).  This
// kernel turns the mapping around for the all-active case (contiguous resident rows):
//
//   * one producer thread streams [128 data sets] x [16 channels] boxes of the resident
//     row matrix into a ring of shared-memory stages with tiled tensor-TMA copies
//     (cp.async.bulk.tensor.2d -> SASS UTMALDG), completion tracked by mbarrier
//     transaction counts; the ring keeps tens of KB per SM in flight without registers;
//   * the boxes land with the hardware 128-byte swizzle, so a consumer lane that owns ONE
//     data set (one 128-byte box row) reads its 16-byte chunks conflict-free although all
//     32 lanes of the warp read the same channel pair of 32 different rows;
//   * the channel index is therefore warp-uniform and the KT model spectra are read from the
//     constant bank (uniform loads, LDCU): no shared-memory bandwidth, no per-lane
//     registers; KT accumulators per lane, no cross-lane reduction, coalesced logL stores.
//
// Arithmetic per (element, candidate) is the same DADD + DFMA as in clike_rows_kernel; the
// per-lane summation runs over the channels in order, which is the reference's own order
// (clike.c:64-76).  A first version of this kernel used one 1-D bulk copy per data-set
// segment (works for gathered rows too) but the producer warp could not issue them fast
// enough (0.78 ms per pass at N=1e6 against 0.35 ms of the block kernel); masked batches
// therefore stay on clike_block_kernel.
#include <cuda.h>

#include <cstdlib>

#include "kernels.cuh"

namespace mdns {

constexpr int TILE_BOX_CH = 16;                // channels per box row = 128 bytes (swizzle span)
constexpr int TILE_CMODEL = 6400;              // doubles of model spectra in the constant bank

// model spectra of the current pass: c_model[k*mpitch + j].  Indexed with warp-uniform
// expressions of kernel parameters and loop counters only, so that ptxas keeps the loads on
// the uniform datapath (LDCU); a double2 view made it fall back to per-lane LDC and cost 40 %.
__constant__ double c_model[TILE_CMODEL];

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1,
                                            uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	    : "memory");
}

// TILE_ROWS data sets per tile (one consumer lane each), NBOX boxes (16 channels each) per
// ring stage
template <int KT, int NBOX, int STAGES, int TILE_ROWS>
__global__ void __launch_bounds__(TILE_ROWS + 32) clike_tile_kernel(
    const __grid_constant__ CUtensorMap tmap, const LikeArgs a, const int k0, const int kt_valid)
{
	constexpr int TILE_CONSUMER_WARPS = TILE_ROWS / 32;
	constexpr int TILE_BOX_BYTES = TILE_ROWS * TILE_BOX_CH * 8;
	constexpr int CW = TILE_BOX_CH * NBOX;
	constexpr int STAGE_BYTES = TILE_BOX_BYTES * NBOX;
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
	// the swizzle pattern is a function of the shared-memory address: align the ring to 1 KB
	unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int pitch_even = (int)a.pitch;       // channels incl. the zero pad of an odd count
	const int nchunks = (pitch_even + CW - 1) / CW;
	const int ntiles = (a.n_rows + TILE_ROWS - 1) / TILE_ROWS;

	if (threadIdx.x == 0) {
#pragma unroll
		for (int s = 0; s < STAGES; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], TILE_CONSUMER_WARPS);
		}
		mbar_fence_init();
	}
	__syncthreads();

	if (warp == TILE_CONSUMER_WARPS) {
		// ===================== producer (one elected thread) =====================
		if (lane == 0) {
			int it = 0;
			for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
				const int r0 = a.row0 + tile * TILE_ROWS;
				for (int c = 0; c < nchunks; ++c, ++it) {
					const int stage = it % STAGES;
					const uint32_t round = (uint32_t)(it / STAGES);
					mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);   // first round passes
					// out-of-bounds parts of a box are zero-filled and still counted
					mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
					for (int b = 0; b < NBOX; ++b)
						tma_load_2d(ring + (size_t)stage * STAGE_BYTES + b * TILE_BOX_BYTES, &tmap,
						            c * CW + b * TILE_BOX_CH, r0, &full_bar[stage]);
				}
			}
		}
	} else {
		// ===================== consumer warps =====================
		const int r_local = warp * 32 + lane;
		const int sw = r_local & 7;                // 128-byte swizzle: chunk ^= row % 8
		const double inv = a.scale / a.noise2;
		int it = 0;
		for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
			const long long gr = (long long)tile * TILE_ROWS + r_local;
			// two partial sums per candidate (even / odd channels) while registers allow
			constexpr int NACC = KT <= 8 ? 2 : 1;
			double acc[NACC][KT];
#pragma unroll
			for (int k = 0; k < KT; ++k) acc[0][k] = acc[NACC - 1][k] = 0.0;
			for (int c = 0; c < nchunks; ++c, ++it) {
				const int stage = it % STAGES;
				const uint32_t round = (uint32_t)(it / STAGES);
				mbar_wait(&full_bar[stage], round & 1u);
				const unsigned char *rowb = ring + (size_t)stage * STAGE_BYTES + r_local * 128;
#pragma unroll
				for (int b = 0; b < NBOX; ++b) {
					const int jbase = c * CW + b * TILE_BOX_CH;
					const int left = pitch_even - jbase;     // valid channels of this box (even)
					if (left >= TILE_BOX_CH) {
#pragma unroll
						for (int u = 0; u < TILE_BOX_CH / 2; ++u) {
							const double2 y = *reinterpret_cast<const double2 *>(
							    rowb + b * TILE_BOX_BYTES + ((u ^ sw) << 4));
#pragma unroll
							for (int k = 0; k < KT; ++k) {
								const double d0 = c_model[k * a.mpitch + jbase + 2 * u] - y.x;
								const double d1 = c_model[k * a.mpitch + jbase + 2 * u + 1] - y.y;
								acc[0][k] = fma(d0, d0, acc[0][k]);
								acc[NACC - 1][k] = fma(d1, d1, acc[NACC - 1][k]);
							}
						}
					} else if (left > 0) {
						for (int u = 0; u < left / 2; ++u) {
							const double2 y = *reinterpret_cast<const double2 *>(
							    rowb + b * TILE_BOX_BYTES + ((u ^ sw) << 4));
#pragma unroll
							for (int k = 0; k < KT; ++k) {
								const double d0 = c_model[k * a.mpitch + jbase + 2 * u] - y.x;
								const double d1 = c_model[k * a.mpitch + jbase + 2 * u + 1] - y.y;
								acc[0][k] = fma(d0, d0, acc[0][k]);
								acc[NACC - 1][k] = fma(d1, d1, acc[NACC - 1][k]);
							}
						}
					}
				}
				__syncwarp();
				if (lane == 0) mbar_arrive(&empty_bar[stage]);   // this warp is done with the stage
			}
			if (gr < a.n_rows) {
#pragma unroll
				for (int k = 0; k < KT; ++k)
					if (k < kt_valid)
						a.out[(long long)(k0 + k) * a.out_stride + gr] =
						    (NACC == 2 ? acc[0][k] + acc[NACC - 1][k] : acc[0][k]) * inv;
			}
		}
	}
}

// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Describe the resident row matrix Y[n_rows][pitch] (doubles) to the TMA unit: boxes of
// tile_rows x 16 channels, 128-byte swizzle, zero fill outside.  `out` receives 128 bytes.
int make_row_tensor_map(void *out, const double *Y, long long n_rows, long long pitch, int tile_rows)
{
	static EncodeTiledFn encode = nullptr;
	if (!encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult q;
		MDNS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
		if (!fn || q != cudaDriverEntryPointSuccess) {
			set_error("cuTensorMapEncodeTiled is not available from this driver");
			return MDNS_ECUDA;
		}
		encode = (EncodeTiledFn)fn;
	}
	static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
	const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)n_rows};
	const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(double)};
	const cuuint32_t box[2] = {TILE_BOX_CH, (cuuint32_t)tile_rows};
	const cuuint32_t estr[2] = {1, 1};
	// L2 promotion of the box rows (experiment knob MDNS_TMAP_L2 = 0 / 64 / 128 / 256)
	CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;   // measured best
	if (const char *e = getenv("MDNS_TMAP_L2")) {
		const int v = atoi(e);
		promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
		      : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
		      : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
		                 : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
	}
	CUtensorMap tm;
	const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)Y, dims, strides, box,
	                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
	                          promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("cuTensorMapEncodeTiled failed with code %d (rows %lld, pitch %lld)", (int)r,
		          n_rows, pitch);
		return MDNS_ECUDA;
	}
	memcpy(out, &tm, sizeof tm);
	return MDNS_OK;
}

// the same descriptor for any box height (model batches: KT candidates x 16 channels)
int make_row_tensor_map_box(void *out, const double *Y, long long n_rows, long long pitch,
                            int box_rows)
{
	return make_row_tensor_map(out, Y, n_rows, pitch, box_rows);
}

template <int KT, int NBOX, int STAGES, int TILE_ROWS>
static int launch_tile_inst(const LikeArgs &a, const void *tmap, int sm_count, cudaStream_t st)
{
	constexpr int TILE_THREADS = TILE_ROWS + 32;
	constexpr int TILE_BOX_BYTES = TILE_ROWS * TILE_BOX_CH * 8;
	const size_t smem = (size_t)STAGES * NBOX * TILE_BOX_BYTES + 1024;   // + alignment slack
	auto kern = clike_tile_kernel<KT, NBOX, STAGES, TILE_ROWS>;
	MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TILE_THREADS, smem));
	if (occ < 1) {
		set_error("clike tile kernel does not fit (%zu bytes of shared memory)", smem);
		return MDNS_EINVAL;
	}
	double *cm = nullptr;
	MDNS_CUDA(cudaGetSymbolAddress((void **)&cm, c_model));
	CUtensorMap tm;
	memcpy(&tm, tmap, sizeof tm);
	const int ntiles = ceil_div(a.n_rows, TILE_ROWS);
	long long gx = ntiles;
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	for (int k0 = 0; k0 < a.K; k0 += KT) {
		const int kv = a.K - k0 < KT ? a.K - k0 : KT;
		// candidates [k0, k0+KT) of the padded model buffer -> constant bank (stream ordered)
		MDNS_CUDA(cudaMemcpyAsync(cm, a.model + (size_t)k0 * a.mpitch,
		                          (size_t)KT * a.mpitch * sizeof(double), cudaMemcpyDeviceToDevice,
		                          st));
		kern<<<(unsigned)gx, TILE_THREADS, smem, st>>>(tm, a, k0, kv);
		MDNS_LAUNCHED("clike_tile_kernel");
	}
	return MDNS_OK;
}

// kt in {4, 8, 16, 32}; nbox in {1, 2}; stages in {3, 4, 6}; tile_rows in {128, 256}.
// Requires a.active == nullptr and the tensor map built for the same tile_rows.
int launch_clike_tile(const LikeArgs &a, const void *tmap, int kt, int nbox, int stages,
                      int tile_rows, int sm_count, cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	if (a.active || !tmap) {
		set_error("tile kernel needs the contiguous all-active layout");
		return MDNS_EINVAL;
	}
	if ((long long)kt * a.mpitch > TILE_CMODEL) {
		set_error("tile kernel: %d candidates x %d channels exceed the constant bank", kt, a.mpitch);
		return MDNS_EINVAL;
	}
#define MDNS_TILE(KK, BB, SS, RR)                                         \
	if (kt == KK && nbox == BB && stages == SS && tile_rows == RR) \
		return launch_tile_inst<KK, BB, SS, RR>(a, tmap, sm_count, st)
#define MDNS_TILE_K(BB, SS, RR) \
	MDNS_TILE(4, BB, SS, RR);   \
	MDNS_TILE(8, BB, SS, RR);   \
	MDNS_TILE(16, BB, SS, RR);  \
	MDNS_TILE(32, BB, SS, RR)
	MDNS_TILE_K(1, 4, 128);
	MDNS_TILE_K(1, 6, 128);
	MDNS_TILE_K(2, 3, 128);
	MDNS_TILE_K(1, 3, 256);
	MDNS_TILE_K(1, 4, 256);
	MDNS_TILE_K(2, 3, 256);
#undef MDNS_TILE_K
#undef MDNS_TILE
	set_error("unsupported tile-kernel shape kt=%d nbox=%d stages=%d rows=%d", kt, nbox, stages,
	          tile_rows);
	return MDNS_EINVAL;
}

int tile_constant_capacity() { return TILE_CMODEL; }

}  // namespace mdns
