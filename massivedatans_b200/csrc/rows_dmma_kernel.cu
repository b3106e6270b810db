// rows_dmma_kernel.cu -- the K x C x N contraction  S[k][i] = sum_j A[i][j] * B[k][j]  of resident
// rows A (data sets) against a staged batch B (candidate spectra) on the FP64 tensor path
// (mma.sync m8n8k4 f64 -> SASS DMMA), work split evenly over the CTAs in units of
// (256-row tile, 16-channel chunk) -- "stream-K" -- and two epilogues:
//
//   EPI_CLIKE  expanded chi-square of clike.c:64-76,  sum_j (m_kj - y_ij)^2 = Syy_i - 2 S_ik + Smm_k,
//              with the cancellation guard + direct-form fix-up list of clike_xtile_kernel.cu and the
//              accept test of hiermetriclearn.py:193 (`L > Lmins`) fused in: per-candidate counts of
//              accepting data sets leave the kernel next to (or instead of) the logL matrix;
//   EPI_RAW    S itself, for one or two row matrices in one launch, the second one contracted with
//              the SQUARED batch: cmuselike.c:48-64 in expanded form needs S1 = (y/v) . m and
//              S2 = (1/v) . m^2 (muse_xp_finalize_kernel turns them into chi2).
//
// Why stream-K.  The first tensor-path kernel (clike_dmma_kernel.cu, round 1) gave whole tiles to
// CTAs round robin.  489 tiles on 296 resident CTAs (125 000 data sets x 1000 channels, one GPU's
// share of BASELINE configs[3]) left the second wave 65 % full: 0.80 of the HBM roofline at K = 16
// against 0.94 at K = 8; and a MUSE cube (4223 x 3600) has 17 tiles for 148 SMs.  Here CTA b
// owns the contiguous unit range [b*T/G, (b+1)*T/G) of the T = tiles * chunks units, so every CTA
// streams the same number of bytes whatever the shape.  A tile cut by a CTA boundary is summed from
// per-CTA partial accumulators parked in a workspace (at most two per CTA: the head of its first
// tile and the tail of its last); the CTA whose partial arrives last (one ticket per tile) adds them
// in CTA order -- a fixed order, so results do not depend on timing -- and runs the epilogue.
// Nobody spins on anybody: no co-residency assumption, no deadlock.
//
// Data path as before: a producer thread feeds a ring of tensor-TMA boxes ([256 rows] x [16
// channels] of A plus the matching [KT candidates] x [16 channels] box of B per stage, 128-byte
// swizzle, one mbarrier transaction); masked batches are fed by 64 tile::gather4 copies per stage
// from two producer warps.  Consumer warp w owns rows [8*MR*w, 8*MR*(w+1)) of the tile; fragment
// loads are conflict free thanks to the row permutation 0,2,4,6,1,3,5,7 (see clike_dmma_kernel.cu).
#include <cuda.h>

#include "kernels.cuh"

namespace mdns {

constexpr int RD_ROWS = 256;                    // data sets per tile = rows of one TMA box
constexpr int RD_BOX_CH = 16;                   // channels per box row = 128 bytes (swizzle span)
constexpr int RD_STAGE_BYTES = RD_ROWS * RD_BOX_CH * 8;
constexpr int EPI_CLIKE = 0, EPI_RAW = 1;
constexpr int RD_BLOCK_CH = 64;                 // summation block of the raw contraction (channels)

__device__ __forceinline__ void rd_mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void rd_tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0,
                                               int c1, uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ void rd_tma_gather4(void *smem_dst, const CUtensorMap *tmap, int c0,
                                               int r0, int r1, int r2, int r3, uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ void rd_dmma(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

// The schedule.  VT = nmat * ntiles virtual tiles of nch chunks each, G CTAs.
//   bulk       CTA b takes the whole tiles b, b + G, b + 2G, ... of the first `waves` full waves
//              (neighbouring CTAs stream neighbouring rows, as the round-1 kernel did);
//   remainder  the last VT - waves*G tiles (between G and 2G-1 of them when VT >= G) are cut
//              into G equal unit ranges: CTA b owns units [b*T/G, (b+1)*T/G) of the T = rem*nch.
struct RdSched {
	long long T;       // units of the remainder
	int G;             // CTAs
	int nch;           // chunks per tile
	int waves;         // whole-tile rounds before the remainder
	__device__ __forceinline__ long long begin(int b) const { return (long long)b * T / G; }
	// the CTA whose range holds unit u
	__device__ __forceinline__ int owner(long long u) const { return (int)(((u + 1) * G - 1) / T); }
};

__device__ __forceinline__ RdSched rd_schedule(int nmat, int ntiles, int nch, int G)
{
	RdSched sc;
	const int VT = nmat * ntiles;
	sc.G = G;
	sc.nch = nch;
	sc.waves = VT / G > 1 ? VT / G - 1 : 0;
	sc.T = (long long)(VT - sc.waves * G) * nch;
	return sc;
}

// what a CTA keeps of its schedule (shared memory: the hot loop has no registers to spare)
struct RdMine {
	long long u0, u1;      // unit range within the remainder
	int waves;             // whole tiles before it
	int lv_first, nseg;    // first remainder tile touched (remainder-local) and how many
};

// segment `seg` of this CTA: virtual tile, chunk range, and whether it is a cut tile
__device__ __forceinline__ void rd_segment(const RdMine &m, int seg, int G, int nch, int &vt, int &cs,
                                           int &ce)
{
	if (seg < m.waves) {
		vt = blockIdx.x + seg * G;
		cs = 0;
		ce = nch;
		return;
	}
	const int lv = m.lv_first + (seg - m.waves);
	const long long ub = (long long)lv * nch;
	cs = (int)((m.u0 > ub ? m.u0 : ub) - ub);
	ce = (int)((m.u1 < ub + nch ? m.u1 : ub + nch) - ub);
	vt = m.waves * G + lv;
}

// ---- epilogue of the expanded chi-square, specialised on what the tile needs ------------------
// FULL: every row of the tile and every candidate of the pass is valid (no predicates);
// COUNT: fused accept test into the packed per-lane counters cntp[nc] (two 16-bit counts);
// STORE: the logL matrix is wanted.
template <int NC, int MR, bool GATHER, bool FULL, bool COUNT, bool STORE>
__device__ __forceinline__ void rd_epilogue_clike(const double (&acc)[MR][NC][2], const LikeArgs &a,
                                                  const double *s_smm, const double *s_syy,
                                                  const double *s_lm, int tile, int warp, int pr, int t,
                                                  int kp0, int kp1, int k0, int kt_valid, int pass,
                                                  double inv, unsigned (&cntp)[NC])
{
	unsigned redo_mask = 0;
#pragma unroll
	for (int mr = 0; mr < MR; ++mr) {
		const int lr = warp * (8 * MR) + mr * 8 + pr;            // row within the tile
		const long long gr = (long long)tile * RD_ROWS + lr;
		const bool live = FULL || gr < a.n_rows;
		// Syy (and the threshold) of the tile's rows were prefetched into shared memory while the
		// tile was being contracted (rd_prefetch_rows): no exposed load latency here
		const double syy = live ? s_syy[lr] : 0.0;
		double lm = 0.0;
		if (COUNT) lm = live ? s_lm[lr] : __longlong_as_double(0x7ff0000000000000LL);
		double *o0 = nullptr, *o1 = nullptr;
		if (STORE) {
			o0 = a.out + (long long)(k0 + kp0) * a.out_stride + gr;
			o1 = a.out + (long long)(k0 + kp1) * a.out_stride + gr;
		}
		bool redo = false;
#pragma unroll
		for (int nc = 0; nc < NC; ++nc) {
#pragma unroll
			for (int i = 0; i < 2; ++i) {
				const int kp = i ? kp1 : kp0;
				const double smm = s_smm[nc * 8 + kp];
				const double chi = syy + fma(-2.0, acc[mr][nc][i], smm);
				const bool ok = chi >= a.xp_guard * (syy + smm);   // false for NaN too
				const bool valid = FULL || (live && nc * 8 + kp < kt_valid);
				const double val = chi * inv;
				if (STORE && valid && ok) (i ? o1 : o0)[(long long)nc * 8 * a.out_stride] = val;
				if (COUNT && valid && ok && val > lm) cntp[nc] += 1u << (16 * i);
				redo = redo || (valid && !ok);
			}
		}
		if (redo) redo_mask |= 1u << mr;
	}
	// the four lanes of a group share the data set: list it once
	redo_mask |= __shfl_xor_sync(0xffffffffu, redo_mask, 1);
	redo_mask |= __shfl_xor_sync(0xffffffffu, redo_mask, 2);
	if (redo_mask && t == 0) {
#pragma unroll
		for (int mr = 0; mr < MR; ++mr)
			if (redo_mask & (1u << mr))
				a.xp_list[atomicAdd(a.xp_redo + 1 + pass, 1)] =
				    (int)((long long)tile * RD_ROWS + warp * (8 * MR) + mr * 8 + pr);
	}
}

// Per-row epilogue inputs of a tile (Syy, accept threshold) fetched asynchronously into shared
// memory (cp.async, no registers held across the contraction); every warp fetches and later
// reads only its own rows, so a warp-level wait is all the synchronisation needed.
__device__ __forceinline__ void rd_cp_async8(double *smem_dst, const double *gmem_src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
	             : "memory");
}

template <int MR, bool GATHER>
__device__ __forceinline__ void rd_prefetch_rows(const LikeArgs &a, double *s_syy, double *s_lm, int tile,
                                                 int warp, int pr, int t)
{
	if (t == 0) {
#pragma unroll
		for (int mr = 0; mr < MR; ++mr) {
			const int lr = warp * (8 * MR) + mr * 8 + pr;
			const long long gr = (long long)tile * RD_ROWS + lr;
			if (gr < a.n_rows) {
				rd_cp_async8(s_syy + lr, GATHER ? a.syy + a.active[gr] : a.syy + a.row0 + gr);
				if (a.lmins) rd_cp_async8(s_lm + lr, a.lmins + gr);
			}
		}
	}
	asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NC, int STAGES, int MR, bool GATHER, int EPI>
__global__ void __launch_bounds__(RD_ROWS / (8 * MR) * 32 + (GATHER ? 64 : 32), (NC == 2 && MR == 4) ? 2 : 1) rows_dmma_kernel(
    const __grid_constant__ CUtensorMap tmapA0, const __grid_constant__ CUtensorMap tmapA1,
    const __grid_constant__ CUtensorMap tmapB0, const __grid_constant__ CUtensorMap tmapB1,
    const LikeArgs a, const int k0, const int kt_valid, const int pass, const int nmat)
{
	constexpr int KT = NC * 8;
	constexpr int WARPS = RD_ROWS / (8 * MR);                   // consumer warps
	constexpr int CTHREADS = WARPS * 32;
	constexpr int MODEL_BYTES = KT * RD_BOX_CH * 8;
	constexpr int STAGE_BYTES = RD_STAGE_BYTES + MODEL_BYTES;   // multiple of 1 KB
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint64_t full_bar[STAGES], empty_bar[STAGES];
	__shared__ double s_smm[KT];
	__shared__ double s_syy[RD_ROWS], s_lm[RD_ROWS];
	__shared__ int s_counts[KT];
	__shared__ int s_last;
	__shared__ RdMine s_mine;
	unsigned char *ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int ntiles = (a.n_rows + RD_ROWS - 1) / RD_ROWS;
	const int nch = ((int)a.pitch + RD_BOX_CH - 1) / RD_BOX_CH;
	const int G = gridDim.x;

	pdl_trigger();      // the fix-up / finalize launch behind this one waits for the whole grid itself
	if (threadIdx.x == 0) {
		const RdSched sc = rd_schedule(nmat, ntiles, nch, G);
		RdMine m;
		m.waves = sc.waves;
		m.u0 = sc.begin(blockIdx.x);
		m.u1 = sc.begin(blockIdx.x + 1);
		m.lv_first = (int)(m.u0 / nch);
		m.nseg = m.u1 > m.u0 ? (int)((m.u1 - 1) / nch) - m.lv_first + 1 : 0;
		s_mine = m;
#pragma unroll
		for (int s = 0; s < STAGES; ++s) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], WARPS);
		}
		mbar_fence_init();
	}
	// (programmatic dependent launch: everything below reads the model kernel's output)
	pdl_wait();
	if (threadIdx.x < KT) {
		s_counts[threadIdx.x] = 0;
		s_smm[threadIdx.x] = EPI == EPI_CLIKE ? a.smm[k0 + threadIdx.x] : 0.0;
	}
	__syncthreads();
	const int nseg = s_mine.waves + s_mine.nseg;

	if (warp >= WARPS) {
		// ===================== producer =====================
		int it = 0;
		if (GATHER) {
			const int pl = (warp - WARPS) * 32 + lane;     // 0..63: one group of four rows each
			for (int seg = 0; seg < nseg; ++seg) {
				int vt, cs, ce;
				rd_segment(s_mine, seg, G, nch, vt, cs, ce);
				const int mat = vt >= ntiles ? 1 : 0;
				const int tile = vt - mat * ntiles;
				const CUtensorMap *ta = mat ? &tmapA1 : &tmapA0;
				const CUtensorMap *tb = mat ? &tmapB1 : &tmapB0;
				int rows[4];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const long long r = (long long)tile * RD_ROWS + pl * 4 + j;
					rows[j] = a.active[r < a.n_rows ? r : a.n_rows - 1];
				}
				for (int c = cs; c < ce; ++c, ++it) {
					const int stage = it % STAGES;
					const uint32_t round = (uint32_t)(it / STAGES);
					unsigned char *dst = ring + (size_t)stage * STAGE_BYTES;
					if (pl == 0) {
						mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);
						mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
						rd_tma_load_2d(dst + RD_STAGE_BYTES, tb, c * RD_BOX_CH, k0, &full_bar[stage]);
					}
					__syncwarp();
					asm volatile("bar.sync 1, 64;" ::: "memory");
					rd_tma_gather4(dst + pl * 512, ta, c * RD_BOX_CH, rows[0], rows[1], rows[2],
					               rows[3], &full_bar[stage]);
				}
			}
		} else if (lane == 0) {
			for (int seg = 0; seg < nseg; ++seg) {
				int vt, cs, ce;
				rd_segment(s_mine, seg, G, nch, vt, cs, ce);
				const int mat = vt >= ntiles ? 1 : 0;
				const int tile = vt - mat * ntiles;
				const CUtensorMap *ta = mat ? &tmapA1 : &tmapA0;
				const CUtensorMap *tb = mat ? &tmapB1 : &tmapB0;
				const int r0 = a.row0 + tile * RD_ROWS;
				for (int c = cs; c < ce; ++c, ++it) {
					const int stage = it % STAGES;
					const uint32_t round = (uint32_t)(it / STAGES);
					mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);
					// out-of-bounds parts of a box are zero-filled and still counted
					mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
					unsigned char *dst = ring + (size_t)stage * STAGE_BYTES;
					rd_tma_load_2d(dst, ta, c * RD_BOX_CH, r0, &full_bar[stage]);
					rd_tma_load_2d(dst + RD_STAGE_BYTES, tb, c * RD_BOX_CH, k0, &full_bar[stage]);
				}
			}
		}
	} else {
		// ===================== consumer warps =====================
		const int g = lane >> 2, t = lane & 3;
		const int pr = ((g & 3) << 1) | (g >> 2);      // physical row of logical row g
		const int tid = warp * 32 + lane;
		const int a_row_off = (warp * (8 * MR) + pr) * 128 + (t & 1) * 8;
		const int a_chunk = t >> 1;
		const int b_row_off = RD_STAGE_BYTES + pr * 128 + (t & 1) * 8;
		// logical columns 2t, 2t+1 of a candidate tile -> physical candidates (same permutation)
		const int kp0 = (((2 * t) & 3) << 1) | ((2 * t) >> 2);
		const int kp1 = (((2 * t + 1) & 3) << 1) | ((2 * t + 1) >> 2);
		const double inv = a.scale / a.noise2;
		unsigned cntp[NC];
#pragma unroll
		for (int nc = 0; nc < NC; ++nc) cntp[nc] = 0;
		int it = 0;
		for (int seg = 0; seg < nseg; ++seg) {
			int vt, cs, ce;
			rd_segment(s_mine, seg, G, nch, vt, cs, ce);
			const int mat = vt >= ntiles ? 1 : 0;
			const int tile = vt - mat * ntiles;
			double acc[MR][NC][2];
#pragma unroll
			for (int mr = 0; mr < MR; ++mr)
#pragma unroll
				for (int nc = 0; nc < NC; ++nc) acc[mr][nc][0] = acc[mr][nc][1] = 0.0;
			if (EPI == EPI_CLIKE) rd_prefetch_rows<MR, GATHER>(a, s_syy, s_lm, tile, warp, pr, t);
			// raw contractions (long MUSE spectra) are summed in blocks of RD_BLOCK_CH channels: the
			// rounding-error bound of a sum of C terms drops from ~C*u to ~(B + C/B)*u, which is
			// what lets the expanded cmuselike form vouch for well-fitting candidates (muse_xp.cu)
			double acc2[EPI == EPI_RAW ? MR : 1][EPI == EPI_RAW ? NC : 1][2];
			if (EPI == EPI_RAW) {
#pragma unroll
				for (int mr = 0; mr < MR; ++mr)
#pragma unroll
					for (int nc = 0; nc < NC; ++nc) acc2[mr][nc][0] = acc2[mr][nc][1] = 0.0;
			}
			for (int c = cs; c < ce; ++c, ++it) {
				const int stage = it % STAGES;
				const uint32_t round = (uint32_t)(it / STAGES);
				mbar_wait(&full_bar[stage], round & 1u);
				const unsigned char *sbase = ring + (size_t)stage * STAGE_BYTES;
#pragma unroll
				for (int ks = 0; ks < RD_BOX_CH / 4; ++ks) {
					double fa[MR], fb[NC];
					const int choff = ((2 * ks + a_chunk) ^ pr) << 4;
#pragma unroll
					for (int mr = 0; mr < MR; ++mr)
						fa[mr] = *reinterpret_cast<const double *>(sbase + a_row_off + mr * 1024 + choff);
#pragma unroll
					for (int nc = 0; nc < NC; ++nc) {
						fb[nc] = *reinterpret_cast<const double *>(sbase + b_row_off + nc * 1024 + choff);
						// second pair of the raw contraction: the SQUARED batch (cmuselike's sum m^2/v)
						if (EPI == EPI_RAW && mat) fb[nc] *= fb[nc];
					}
#pragma unroll
					for (int mr = 0; mr < MR; ++mr)
#pragma unroll
						for (int nc = 0; nc < NC; ++nc)
							rd_dmma(acc[mr][nc][0], acc[mr][nc][1], fa[mr], fb[nc]);
				}
				__syncwarp();
				if (lane == 0) rd_mbar_arrive(&empty_bar[stage]);
				if (EPI == EPI_RAW && ((c - cs) % (RD_BLOCK_CH / RD_BOX_CH)) == RD_BLOCK_CH / RD_BOX_CH - 1) {
#pragma unroll
					for (int mr = 0; mr < MR; ++mr)
#pragma unroll
						for (int nc = 0; nc < NC; ++nc)
#pragma unroll
							for (int i = 0; i < 2; ++i) {
								acc2[mr][nc][i] += acc[mr][nc][i];
								acc[mr][nc][i] = 0.0;
							}
				}
			}
			if (EPI == EPI_RAW) {
#pragma unroll
				for (int mr = 0; mr < MR; ++mr)
#pragma unroll
					for (int nc = 0; nc < NC; ++nc)
#pragma unroll
						for (int i = 0; i < 2; ++i) acc[mr][nc][i] += acc2[mr][nc][i];
			}
			if (cs != 0 || ce != nch) {
				// a cut tile: park this CTA's partial sums; the last of the tile's CTAs to get here
				// adds all of them up in CTA order and carries on to the epilogue
				const int lv = vt - s_mine.waves * G;        // remainder-local tile
				double *slot = a.ws + (size_t)(blockIdx.x * 2 + (seg == s_mine.waves ? 0 : 1)) * (RD_ROWS * KT);
#pragma unroll
				for (int mr = 0; mr < MR; ++mr)
#pragma unroll
					for (int nc = 0; nc < NC; ++nc)
#pragma unroll
						for (int i = 0; i < 2; ++i)
							__stcg(slot + ((mr * NC + nc) * 2 + i) * CTHREADS + tid, acc[mr][nc][i]);
				__threadfence();
				asm volatile("bar.sync 2, %0;" ::"n"(CTHREADS) : "memory");
				const RdSched sc = rd_schedule(nmat, ntiles, nch, G);
				const long long ub = (long long)lv * nch;
				const int b_first = sc.owner(ub), b_last = sc.owner(ub + nch - 1);
				if (tid == 0) {
					// release / acquire through ONE thread, as a grid barrier does it: the fence after
					// the CTA barrier is cumulative over the other threads' stores, the one after the
					// atomic orders the reads of the other CTAs' slots behind it.  (Fences by the
					// storing threads alone, before the barrier, let a few rows of a slot arrive late
					// on the B200: the finalising CTA then saw the previous launch's values.)
					__threadfence();
					const int old = atomicAdd(a.tickets + vt, 1);
					__threadfence();
					const int last = old == b_last - b_first ? 1 : 0;
					if (last) a.tickets[vt] = 0;      // ready for the next launch
					s_last = last;
				}
				asm volatile("bar.sync 2, %0;" ::"n"(CTHREADS) : "memory");
				if (!s_last) continue;
				__threadfence();
#pragma unroll
				for (int mr = 0; mr < MR; ++mr)
#pragma unroll
					for (int nc = 0; nc < NC; ++nc) acc[mr][nc][0] = acc[mr][nc][1] = 0.0;
				for (int bb = b_first; bb <= b_last; ++bb) {
					const int first_of_bb = (int)(sc.begin(bb) / nch) == lv ? 0 : 1;
					const double *src = a.ws + (size_t)(bb * 2 + first_of_bb) * (RD_ROWS * KT);
#pragma unroll
					for (int mr = 0; mr < MR; ++mr)
#pragma unroll
						for (int nc = 0; nc < NC; ++nc)
#pragma unroll
							for (int i = 0; i < 2; ++i)
								acc[mr][nc][i] += __ldcg(src + ((mr * NC + nc) * 2 + i) * CTHREADS + tid);
				}
			}
			// ---- epilogue: lane holds S[row(g)][2t + {0,1}] of every (row tile, candidate tile)
			if (EPI == EPI_RAW) {
				double *outp = mat ? a.out_b : a.out;
#pragma unroll
				for (int mr = 0; mr < MR; ++mr) {
					const long long gr = (long long)tile * RD_ROWS + warp * (8 * MR) + mr * 8 + pr;
					const bool live = gr < a.n_rows;
#pragma unroll
					for (int nc = 0; nc < NC; ++nc)
#pragma unroll
						for (int i = 0; i < 2; ++i) {
							const int k = nc * 8 + (i ? kp1 : kp0);
							if (live && k < kt_valid)
								outp[(long long)(k0 + k) * a.out_stride + gr] = acc[mr][nc][i];
						}
				}
				continue;
			}
			const bool full = kt_valid == KT && (long long)(tile + 1) * RD_ROWS <= a.n_rows;
			asm volatile("cp.async.wait_all;" ::: "memory");
			__syncwarp();
#define RD_EPI(FULL, COUNT, STORE)                                                                       \
	rd_epilogue_clike<NC, MR, GATHER, FULL, COUNT, STORE>(acc, a, s_smm, s_syy, s_lm, tile, warp, pr, t, \
	                                                      kp0, kp1, k0, kt_valid, pass, inv, cntp)
			if (a.counts) {
				if (a.out) {
					if (full) RD_EPI(true, true, true); else RD_EPI(false, true, true);
				} else {
					if (full) RD_EPI(true, true, false); else RD_EPI(false, true, false);
				}
				// the packed counters hold 16 bits: flush long before they can overflow
				if ((seg & 2047) == 2047) {
#pragma unroll
					for (int nc = 0; nc < NC; ++nc) {
						if (cntp[nc] & 0xffffu) atomicAdd(&s_counts[nc * 8 + kp0], (int)(cntp[nc] & 0xffffu));
						if (cntp[nc] >> 16) atomicAdd(&s_counts[nc * 8 + kp1], (int)(cntp[nc] >> 16));
						cntp[nc] = 0;
					}
				}
			} else {
				if (full) RD_EPI(true, false, true); else RD_EPI(false, false, true);
			}
#undef RD_EPI
		}
		if (EPI == EPI_CLIKE && a.counts) {
#pragma unroll
			for (int nc = 0; nc < NC; ++nc) {
				if (cntp[nc] & 0xffffu) atomicAdd(&s_counts[nc * 8 + kp0], (int)(cntp[nc] & 0xffffu));
				if (cntp[nc] >> 16) atomicAdd(&s_counts[nc * 8 + kp1], (int)(cntp[nc] >> 16));
			}
		}
	}
	if (EPI == EPI_CLIKE && a.counts) {
		__syncthreads();
		if (threadIdx.x < KT) {
			const int c = s_counts[threadIdx.x];
			if (c) atomicAdd(a.counts + k0 + threadIdx.x, c);
		}
	}
}

// ---- host side ---------------------------------------------------------------------------
static size_t rd_smem(int kt, int stages)
{
	return (size_t)stages * (RD_STAGE_BYTES + (size_t)kt * RD_BOX_CH * 8) + 1024;
}

int make_row_tensor_map_box(void *out, const double *Y, long long n_rows, long long pitch,
                            int box_rows);
int launch_xtile_fixup(const LikeArgs &a, int k0, int kv, int pass, int sm_count, cudaStream_t st);

size_t rows_dmma_workspace_doubles(int sm_count)
{
	// two partial tiles per resident CTA (at most 2 CTAs per SM... the occupancy query may say 3
	// for the smallest shapes), KT <= 32
	return (size_t)2 * sm_count * 3 * RD_ROWS * 32;
}

template <int NC, int STAGES, int MR, bool GATHER, int EPI>
static int launch_rd_inst(const LikeArgs &a, int nmat, int sm_count, cudaStream_t st)
{
	constexpr int KT = NC * 8;
	constexpr int THREADS = RD_ROWS / (8 * MR) * 32 + (GATHER ? 64 : 32);
	const size_t smem = rd_smem(KT, STAGES);
	auto kern = rows_dmma_kernel<NC, STAGES, MR, GATHER, EPI>;
	MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int occ = 0;
	MDNS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
	if (occ < 1) {
		set_error("rows_dmma_kernel does not fit (%zu bytes of shared memory)", smem);
		return MDNS_EINVAL;
	}
	if (occ > 3) occ = 3;      // workspace slots
	CUtensorMap ta0, ta1, tb0, tb1;
	memcpy(&ta0, GATHER ? a.tmap_gather : a.tmap256, sizeof ta0);
	memcpy(&ta1, nmat > 1 ? (GATHER ? a.tmap_gather_b : a.tmap256_b) : (GATHER ? a.tmap_gather : a.tmap256),
	       sizeof ta1);
	// the staged batch [Kpad][mpitch] as boxes of KT candidates x 16 channels (Kpad is a multiple
	// of 32, so a box never leaves the buffer)
	const long long kpad = (long long)round_up(a.K, KT_MAX);
	int rc = make_row_tensor_map_box(&tb0, a.model, kpad, a.mpitch, KT);
	if (rc != MDNS_OK) return rc;
	tb1 = tb0;      // the second pair contracts with the same batch, squared on the fly
	const int ntiles = ceil_div(a.n_rows, RD_ROWS);
	const int nch = ceil_div(a.pitch, RD_BOX_CH);
	const long long T = (long long)nmat * ntiles * nch;
	// at least 8 chunks (256 KB of rows) per CTA on average; problems too small to give every
	// resident CTA 64 chunks run one CTA per SM (half as many partial tiles to exchange)
	long long gx = (T + 7) / 8;
	const long long resident = (long long)sm_count * occ;
	if (gx > resident) gx = resident;
	if (gx > sm_count && T / gx < 64) gx = T / 64 > sm_count ? T / 64 : sm_count;
	if (gx < 1) gx = 1;
	if (const char *e = getenv("MDNS_RD_GX")) {      // experiment knob: CTAs of the launch
		const long long v = atoll(e);
		if (v >= 1 && v <= resident) gx = v;
	}
	const int npass = ceil_div(a.K, KT);
	if (EPI == EPI_CLIKE) {
		if (npass + 1 > xtile_counter_capacity()) {
			set_error("rows_dmma_kernel: %d passes exceed the counter block", npass);
			return MDNS_EINVAL;
		}
		// list lengths of this launch's passes (counter[0], the running total, is left alone)
		if (!a.xp_counters_clear)
			MDNS_CUDA(cudaMemsetAsync(a.xp_redo + 1, 0, (size_t)npass * sizeof(int), st));
	}
	for (int k0 = 0, pass = 0; k0 < a.K; k0 += KT, ++pass) {
		const int kv = a.K - k0 < KT ? a.K - k0 : KT;
		launch_pdl(kern, dim3((unsigned)gx), dim3(THREADS), smem, st,
		           EPI == EPI_CLIKE && (pass > 0 || a.xp_counters_clear), ta0, ta1, tb0, tb1, a, k0, kv, pass, nmat);
		MDNS_LAUNCHED(EPI == EPI_RAW ? (GATHER ? "rows_dmma_kernel(raw,gather)" : "rows_dmma_kernel(raw)")
		                             : (GATHER ? "rows_dmma_kernel(gather)" : "rows_dmma_kernel"));
		if (EPI == EPI_CLIKE) {
			rc = launch_xtile_fixup(a, k0, kv, pass, sm_count, st);
			if (rc != MDNS_OK) return rc;
		}
	}
	return MDNS_OK;
}

bool rows_dmma_fits(const LikeArgs &a, int kt, int stages)
{
	// (16 candidates with 16 consumer warps is not instantiated: it needs a 56-register cap to keep
	// two CTAs per SM, spills under it, was no faster than 8 warps x 32 data sets -- and gave wrong
	// sums in a few rows of a cut tile about once in four launches on the B200, the only shape to)
	static const int shapes[][2] = {{8, 3}, {8, 13}, {8, 14}, {16, 3}, {32, 3}, {32, 2}};
	bool known = false;
	for (auto &sh : shapes) known = known || (sh[0] == kt && sh[1] == stages);
	if (!known) return false;
	if (stages > 10) stages -= 10;
	return (a.active ? a.tmap_gather != nullptr : a.tmap256 != nullptr) && a.ws && a.tickets &&
	       rd_smem(kt, stages) <= 220 * 1024;
}

// kt in {8, 16, 32}; stages in {2, 3, 4}, + 10 for 16 consumer warps of 16 data sets each
// instead of 8 warps of 32.  raw: S itself into a.out (and a.out_b for the second pair, nmat = 2).
int launch_rows_dmma(const LikeArgs &a, int kt, int stages, bool raw, int nmat, int sm_count,
                     cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	if (!rows_dmma_fits(a, kt, stages) || (!raw && !(a.syy && a.smm && a.xp_redo && a.xp_list))) {
		set_error("rows_dmma_kernel: needs the tensor maps, the workspace, the resident row sums and "
		          "%zu bytes of shared memory", rd_smem(kt, stages));
		return MDNS_EINVAL;
	}
#define MDNS_RD(KK, SS, MM)                                                                          \
	if (kt == KK && stages == SS + (MM == 2 ? 10 : 0)) {                                         \
		if (raw)                                                                             \
			return a.active ? launch_rd_inst<KK / 8, SS, MM, true, EPI_RAW>(a, nmat, sm_count, st)   \
			                : launch_rd_inst<KK / 8, SS, MM, false, EPI_RAW>(a, nmat, sm_count, st); \
		return a.active ? launch_rd_inst<KK / 8, SS, MM, true, EPI_CLIKE>(a, 1, sm_count, st)   \
		                : launch_rd_inst<KK / 8, SS, MM, false, EPI_CLIKE>(a, 1, sm_count, st); \
	}
	MDNS_RD(8, 3, 4)
	MDNS_RD(8, 3, 2)
	MDNS_RD(8, 4, 2)
	MDNS_RD(16, 3, 4)
	MDNS_RD(32, 3, 4)
	MDNS_RD(32, 2, 4)
#undef MDNS_RD
	set_error("unsupported rows_dmma_kernel shape kt=%d stages=%d", kt, stages);
	return MDNS_EINVAL;
}

}  // namespace mdns
