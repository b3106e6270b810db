// slab_dmma_kernel.cu -- expanded chi-square of clike.c:64-76 on the FP64 tensor path for SHORT
// spectra (the candidate batch fits in shared memory beside the row rings): every warp streams
// its own slabs of rows through its own TMA ring.
//
// Why (round 2, profiles/r02_rows_dmma_k16.md + r02_nx_sweep.json).  rows_dmma_kernel gives a
// CTA tiles of 256 rows: all 8 consumer warps wait for the same 32 KB stage, the producer refills
// it when the slowest warp has let go, and every 13 chunks (200 channels) the whole CTA stops for
// the epilogue.  At 16 candidates the FP64 tensor work is 65-70 % of the memory time, and those
// couplings leave the consumers waiting for data 25 % of the time while DRAM runs at 75 %:
// 0.82-0.86 of the roofline at 1e6 x 200 against 0.98 at 8 candidates.  Two more losses are
// specific to a pitch of 200 doubles: rows start on odd multiples of 64 bytes, so every
// 128-byte box row straddles two lines (nx = 192: 0.915, 208: 0.88, 200: 0.82 at equal bytes),
// and the 13th chunk is half padding but costs a full chunk of DMMAs.
//
// Here nothing is shared between warps but the candidate batch:
//   * the batch (KT candidates x all channels, 128-byte-swizzled boxes of 16 channels) is loaded
//     ONCE per CTA and stays;
//   * warp w owns NSLOT slots of 4 KB (32 rows x 16 channels).  Lane 0 issues the TMA copy of the
//     box NSLOT units ahead the moment the warp has finished a slot: no producer warp, no empty
//     barriers, a slot is out of flight only for the 32 DMMAs that read it;
//   * every warp starts with a fixed slab (interleaved over the CTAs), further slabs are handed
//     out by an atomic counter (grabbed one slab ahead, so the round trip is never waited for);
//     a warp's epilogue overlaps the other 15 warps' contraction;
//   * P = 2 ("row pairs") when pitch = 8 mod 16 doubles: the matrix is addressed as [N/2] rows
//     of 2*pitch doubles, whose boxes ARE line-aligned.  Boxes 0..S-1 belong to the even row,
//     box S holds the even row's last 8 channels and the odd row's first 8, boxes S+1..2S the
//     rest of the odd row, 8 channels out of step with the batch boxes -- the B fragments of
//     the odd row come from the matching half boxes.  No padding is read or multiplied: 12.5
//     chunks of DMMAs per row instead of 13.
// Channel order and the grouping into k = 4 steps are those of rows_dmma_kernel's uncut tiles,
// so the two kernels agree bit for bit there.
#include <cuda.h>

#include "kernels.cuh"

namespace mdns {

constexpr int SL_WARPS = 16;
constexpr int SL_THREADS = SL_WARPS * 32;
constexpr int SL_ROWS = 32;                     // (super-)rows of a slab = rows of one TMA box
constexpr int SL_SLOT_BYTES = SL_ROWS * 128;    // 32 rows x 16 channels
constexpr int SL_COUNTER_BASE = 1024;           // slab counters live at xp_redo[1024 + pass],
constexpr int SL_MAX_PASSES = 512;              // behind the fix-up counters (xtile_counter_capacity)

__device__ __forceinline__ void sl_tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1,
                                               uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ void sl_tma_gather4(void *smem_dst, const CUtensorMap *tmap, int c0, int r0, int r1,
                                               int r2, int r3, uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ void sl_dmma(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

// k-steps KS0..KS1-1 of one 16-channel box of rows against the batch.  SHIFT = 0: batch box
// `bbox` covers the same channels.  SHIFT = 1 (odd row of a pair): the box's channels are
// 16 j - 8 .. 16 j + 7, i.e. the second half of batch box j - 1 and the first half of box j
// (`bbox` points at box j).
template <int NC, int KS0, int KS1, int SHIFT>
__device__ __forceinline__ void sl_box(double (&acc)[4][NC][2], const unsigned char *aslot,
                                       const unsigned char *bbox, int bbox_bytes, int a_off, int b_off,
                                       int a_chunk, int pr)
{
#pragma unroll
	for (int ks = KS0; ks < KS1; ++ks) {
		double fa[4], fb[NC];
		const int choff = ((2 * ks + a_chunk) ^ pr) << 4;
#pragma unroll
		for (int mr = 0; mr < 4; ++mr)
			fa[mr] = *reinterpret_cast<const double *>(aslot + a_off + mr * 1024 + choff);
		const int ksb = SHIFT ? ((ks + 2) & 3) : ks;
		const unsigned char *bp = (SHIFT && ks < 2) ? bbox - bbox_bytes : bbox;
		const int choffb = ((2 * ksb + a_chunk) ^ pr) << 4;
#pragma unroll
		for (int nc = 0; nc < NC; ++nc)
			fb[nc] = *reinterpret_cast<const double *>(bp + b_off + nc * 1024 + choffb);
#pragma unroll
		for (int mr = 0; mr < 4; ++mr)
#pragma unroll
			for (int nc = 0; nc < NC; ++nc) sl_dmma(acc[mr][nc][0], acc[mr][nc][1], fa[mr], fb[nc]);
	}
}

// Epilogue of one accumulator set: rows gr0 + mr * rstep.  Same arithmetic as
// rd_epilogue_clike (rows_dmma_kernel.cu): chi = Syy - 2 S + Smm, guard, optional store, fused
// accept test into packed 16-bit counters, fix-up list.
template <int NC, bool COUNT, bool STORE, bool GATHER>
__device__ __forceinline__ void sl_epilogue(const double (&acc)[4][NC][2], const LikeArgs &a,
                                            const double *s_smm, long long gr0, int rstep, int t,
                                            int kp0, int kp1, int k0, int kt_valid, int pass, double inv,
                                            unsigned (&cntp)[NC])
{
	double syy[4], lm[4];
#pragma unroll
	for (int mr = 0; mr < 4; ++mr) {
		const long long gr = gr0 + mr * rstep;
		const bool live = gr < a.n_rows;
		// (masked batches: gr is the slot in the compacted order, Syy lives with the shard's rows)
		syy[mr] = live ? __ldg(a.syy + (GATHER ? (long long)a.active[gr] : a.row0 + gr)) : 0.0;
		lm[mr] = 0.0;
		if (COUNT) lm[mr] = live ? __ldg(a.lmins + gr) : __longlong_as_double(0x7ff0000000000000LL);
	}
	unsigned redo_mask = 0;
#pragma unroll
	for (int mr = 0; mr < 4; ++mr) {
		const long long gr = gr0 + mr * rstep;
		const bool live = gr < a.n_rows;
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
				const double chi = syy[mr] + fma(-2.0, acc[mr][nc][i], smm);
				const bool ok = chi >= a.xp_guard * (syy[mr] + smm);   // false for NaN too
				const bool valid = live && nc * 8 + kp < kt_valid;
				const double val = chi * inv;
				if (STORE && valid && ok) (i ? o1 : o0)[(long long)nc * 8 * a.out_stride] = val;
				if (COUNT && valid && ok && val > lm[mr]) cntp[nc] += 1u << (16 * i);
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
		for (int mr = 0; mr < 4; ++mr)
			if (redo_mask & (1u << mr))
				a.xp_list[atomicAdd(a.xp_redo + 1 + pass, 1)] = (int)(gr0 + mr * rstep);
	}
}

// one box of a slab on its way: plain rows (one tiled copy) or listed rows (masked batches: eight
// gather4 copies of four listed rows each, issued by lanes 0..7 with the indices they hold)
template <bool GATHER>
__device__ __forceinline__ void sl_issue(unsigned char *slot, uint64_t *bar, const CUtensorMap *tmap, int box,
                                         int slab, const int (&rows)[4], int lane)
{
	if (lane == 0) mbar_expect_tx(bar, SL_SLOT_BYTES);
	if (GATHER) {
		__syncwarp();
		if (lane < 8) sl_tma_gather4(slot + lane * 512, tmap, box * 16, rows[0], rows[1], rows[2], rows[3], bar);
	} else if (lane == 0) {
		sl_tma_load_2d(slot, tmap, box * 16, slab * SL_ROWS, bar);
	}
}

// the four listed rows lane l < 8 fetches for slab `slab` (clamped behind the end of the list:
// the duplicates are never looked at)
__device__ __forceinline__ void sl_rows_of(const LikeArgs &a, int slab, int nslabs, int lane, int (&rows)[4])
{
	if (lane < 8 && slab < nslabs) {
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const long long r = (long long)slab * SL_ROWS + lane * 4 + j;
			rows[j] = a.active[r < a.n_rows ? r : a.n_rows - 1];
		}
	}
}

template <int NC, int P, int NSLOT, bool GATHER>
__global__ void __launch_bounds__(SL_THREADS, 1) slab_dmma_kernel(
    const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const LikeArgs a,
    const int k0, const int kt_valid, const int pass)
{
	constexpr int KT = NC * 8;
	constexpr int BBOX_BYTES = KT * 128;                      // batch box: KT candidates x 16 channels
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint64_t batch_bar, full_bar[SL_WARPS * NSLOT];
	__shared__ double s_smm[KT];
	__shared__ int s_counts[KT];
	unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int pitch = (int)a.pitch;
	const int nchb = (pitch + 15) >> 4;                       // batch boxes
	const int S = P == 2 ? (pitch - 8) >> 4 : 0;              // P = 2: the straddling box
	const int nbox = P == 2 ? 2 * S + 1 : nchb;               // row boxes per (super-)row
	const int nsup = (a.n_rows + P - 1) / P;
	const int nslabs = (nsup + SL_ROWS - 1) / SL_ROWS;
	unsigned char *batch = base;
	unsigned char *ring = base + (size_t)nchb * BBOX_BYTES + (size_t)warp * NSLOT * SL_SLOT_BYTES;
	uint64_t *bars = full_bar + warp * NSLOT;
	int *ctr = a.xp_redo + SL_COUNTER_BASE + pass;

	// (programmatic dependent launch: the fix-up launch behind this one may be set up now; it
	// does nothing before its own pdl_wait())
	pdl_trigger();
	if (threadIdx.x == 0) {
		mbar_init(&batch_bar, 1);
		for (int s = 0; s < SL_WARPS * NSLOT; ++s) mbar_init(&full_bar[s], 1);
		mbar_fence_init();
	}
	__syncthreads();

	// ---- this warp's slabs: `cur` is being contracted, `nxt` comes after it, `grab` is the one
	// after that, asked for at the start of `cur` and looked at when `cur` is done
	// The first slab of every warp is fixed -- warp w of CTA b starts with slab b + G w, so a
	// launch with fewer slabs than warps still spreads over all the SMs -- the rest come from the
	// counter (which therefore counts from G * 16).
	// Nothing up to pdl_wait() depends on the model kernel this launch may be overlapping with:
	// the rows are resident data, the slab counter was handed back by the previous fix-up launch.
	const int first_dynamic = (int)gridDim.x * SL_WARPS;
	int cur = (int)blockIdx.x + (int)gridDim.x * warp, nxt = 0, grab = 0;
	if (lane == 0) nxt = first_dynamic + atomicAdd(ctr, 1);
	nxt = __shfl_sync(0xffffffffu, nxt, 0);
	// the copy cursor runs NSLOT boxes ahead of the contraction: box `pb` of slab (ahead ? nxt : cur)
	int pb = 0;
	bool ahead = false;
	int rows_cur[4] = {0, 0, 0, 0}, rows_nxt[4] = {0, 0, 0, 0};
	if (GATHER) {
		sl_rows_of(a, cur, nslabs, lane, rows_cur);
		sl_rows_of(a, nxt, nslabs, lane, rows_nxt);
	}
	if (cur < nslabs) {
		for (int s = 0; s < NSLOT; ++s) {
			sl_issue<GATHER>(ring + s * SL_SLOT_BYTES, &bars[s], &tmapA, pb, cur, rows_cur, lane);
			if (++pb == nbox) {
				// (NSLOT <= nbox: only ever after the last slot)
				pb = 0;
				ahead = true;
			}
		}
	}

	// ---- the candidate batch, its Smm and the list counters are the model kernel's output
	pdl_wait();
	if (threadIdx.x == 0) {
		mbar_expect_tx(&batch_bar, (uint32_t)(nchb * BBOX_BYTES));
		for (int c = 0; c < nchb; ++c)
			sl_tma_load_2d(batch + (size_t)c * BBOX_BYTES, &tmapB, c * 16, k0, &batch_bar);
	}
	if (threadIdx.x < KT) {
		s_counts[threadIdx.x] = 0;
		s_smm[threadIdx.x] = a.smm[k0 + threadIdx.x];
	}
	__syncthreads();

	const int g = lane >> 2, t = lane & 3;
	const int pr = ((g & 3) << 1) | (g >> 2);      // physical row of logical row g (conflict-free loads)
	const int a_off = pr * 128 + (t & 1) * 8;
	const int a_chunk = t >> 1;
	const int b_off = a_off;
	const int kp0 = (((2 * t) & 3) << 1) | ((2 * t) >> 2);
	const int kp1 = (((2 * t + 1) & 3) << 1) | ((2 * t + 1) >> 2);
	const double inv = a.scale / a.noise2;
	unsigned cntp[NC];
#pragma unroll
	for (int nc = 0; nc < NC; ++nc) cntp[nc] = 0;
	int slot = 0;
	uint32_t phase = 0;
	int done = 0;
	mbar_wait(&batch_bar, 0);

	// one box: wait for it, `body`, hand the slot back to the copy cursor
#define SL_STEP(body)                                                                               \
	do {                                                                                        \
		mbar_wait(&bars[slot], phase);                                                      \
		const unsigned char *aslot = ring + slot * SL_SLOT_BYTES;                           \
		body;                                                                               \
		__syncwarp();                                                                       \
		const int ps = ahead ? nxt : cur;                                                   \
		if (ps < nslabs) {                                                                  \
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                \
			if (ahead)                                                                  \
				sl_issue<GATHER>(ring + slot * SL_SLOT_BYTES, &bars[slot], &tmapA, pb, ps, rows_nxt, lane); \
			else                                                                        \
				sl_issue<GATHER>(ring + slot * SL_SLOT_BYTES, &bars[slot], &tmapA, pb, ps, rows_cur, lane); \
		}                                                                                   \
		if (++pb == nbox) {                                                                 \
			pb = 0;                                                                     \
			ahead = true;                                                               \
		}                                                                                   \
		if (++slot == NSLOT) {                                                              \
			slot = 0;                                                                   \
			phase ^= 1u;                                                                \
		}                                                                                   \
	} while (0)

	while (cur < nslabs) {
		if (lane == 0) grab = first_dynamic + atomicAdd(ctr, 1);
		double acc[P][4][NC][2];
#pragma unroll
		for (int p = 0; p < P; ++p)
#pragma unroll
			for (int mr = 0; mr < 4; ++mr)
#pragma unroll
				for (int nc = 0; nc < NC; ++nc) acc[p][mr][nc][0] = acc[p][mr][nc][1] = 0.0;
		if (P == 2) {
			for (int j = 0; j < S; ++j)
				SL_STEP((sl_box<NC, 0, 4, 0>(acc[0], aslot, batch + (size_t)j * BBOX_BYTES, BBOX_BYTES, a_off,
				                              b_off, a_chunk, pr)));
			SL_STEP((sl_box<NC, 0, 2, 0>(acc[0], aslot, batch + (size_t)S * BBOX_BYTES, BBOX_BYTES, a_off, b_off,
			                              a_chunk, pr),
			         sl_box<NC, 2, 4, 1>(acc[P - 1], aslot, batch, BBOX_BYTES, a_off, b_off, a_chunk, pr)));
			for (int j = 1; j <= S; ++j)
				SL_STEP((sl_box<NC, 0, 4, 1>(acc[P - 1], aslot, batch + (size_t)j * BBOX_BYTES, BBOX_BYTES, a_off,
				                              b_off, a_chunk, pr)));
		} else {
			for (int j = 0; j < nbox - 1; ++j)
				SL_STEP((sl_box<NC, 0, 4, 0>(acc[0], aslot, batch + (size_t)j * BBOX_BYTES, BBOX_BYTES, a_off,
				                              b_off, a_chunk, pr)));
			// the last box: skip the k-steps that are all padding
			if (pitch - 16 * (nbox - 1) <= 8)
				SL_STEP((sl_box<NC, 0, 2, 0>(acc[0], aslot, batch + (size_t)(nbox - 1) * BBOX_BYTES, BBOX_BYTES,
				                              a_off, b_off, a_chunk, pr)));
			else
				SL_STEP((sl_box<NC, 0, 4, 0>(acc[0], aslot, batch + (size_t)(nbox - 1) * BBOX_BYTES, BBOX_BYTES,
				                              a_off, b_off, a_chunk, pr)));
		}
		// ---- epilogue: lane holds S[row(g)][2t + {0,1}] of every (row group, candidate tile)
#pragma unroll
		for (int p = 0; p < P; ++p) {
			const long long gr0 = ((long long)cur * SL_ROWS + pr) * P + p;
			if (a.counts) {
				if (a.out)
					sl_epilogue<NC, true, true, GATHER>(acc[p], a, s_smm, gr0, 8 * P, t, kp0, kp1, k0, kt_valid, pass,
					                            inv, cntp);
				else
					sl_epilogue<NC, true, false, GATHER>(acc[p], a, s_smm, gr0, 8 * P, t, kp0, kp1, k0, kt_valid, pass,
					                             inv, cntp);
			} else {
				sl_epilogue<NC, false, true, GATHER>(acc[p], a, s_smm, gr0, 8 * P, t, kp0, kp1, k0, kt_valid, pass, inv,
				                             cntp);
			}
		}
		// the packed counters hold 16 bits (a slab adds at most 4 P per half): flush in time
		if (a.counts && (++done & 2047) == 0) {
#pragma unroll
			for (int nc = 0; nc < NC; ++nc) {
				if (cntp[nc] & 0xffffu) atomicAdd(&s_counts[nc * 8 + kp0], (int)(cntp[nc] & 0xffffu));
				if (cntp[nc] >> 16) atomicAdd(&s_counts[nc * 8 + kp1], (int)(cntp[nc] >> 16));
				cntp[nc] = 0;
			}
		}
		cur = nxt;
		nxt = __shfl_sync(0xffffffffu, grab, 0);
		ahead = false;
		if (GATHER) {
#pragma unroll
			for (int j = 0; j < 4; ++j) rows_cur[j] = rows_nxt[j];
			sl_rows_of(a, nxt, nslabs, lane, rows_nxt);
		}
	}
#undef SL_STEP
	if (a.counts) {
#pragma unroll
		for (int nc = 0; nc < NC; ++nc) {
			if (cntp[nc] & 0xffffu) atomicAdd(&s_counts[nc * 8 + kp0], (int)(cntp[nc] & 0xffffu));
			if (cntp[nc] >> 16) atomicAdd(&s_counts[nc * 8 + kp1], (int)(cntp[nc] >> 16));
		}
		__syncthreads();
		if (threadIdx.x < KT) {
			const int c = s_counts[threadIdx.x];
			if (c) atomicAdd(a.counts + k0 + threadIdx.x, c);
		}
	}
}

// ---- host side ---------------------------------------------------------------------------
int make_row_tensor_map_box(void *out, const double *Y, long long n_rows, long long pitch, int box_rows);
int launch_xtile_fixup(const LikeArgs &a, int k0, int kv, int pass, int sm_count, cudaStream_t st);

static int slab_pair_mode(const LikeArgs &a)
{
	if (a.active) return 1;      // listed rows: no pairs
	// row pairs: pitch = 8 mod 16 doubles, and the launch starts on a line boundary
	return (a.pitch % 16 == 8 && ((uintptr_t)a.Y & 127) == 0 && (a.row0 & 1) == 0) ? 2 : 1;
}

static size_t slab_smem(const LikeArgs &a, int kt, int nslot)
{
	const size_t nchb = (size_t)(a.pitch + 15) / 16;
	return 1024 + nchb * kt * 128 + (size_t)SL_WARPS * nslot * SL_SLOT_BYTES;
}

long long slab_dmma_slabs(const LikeArgs &a)
{
	const int P = slab_pair_mode(a);
	return (((long long)a.n_rows + P - 1) / P + SL_ROWS - 1) / SL_ROWS;
}

int slab_counter_base() { return SL_COUNTER_BASE; }
int slab_counter_count() { return SL_MAX_PASSES; }

// kt in {8, 16}; nslot in {2, 3}
bool slab_dmma_fits(const LikeArgs &a, int kt, int nslot)
{
	if ((kt != 8 && kt != 16) || (nslot != 2 && nslot != 3)) return false;
	if (!a.Y || !a.syy || !a.smm || !a.xp_redo || !a.xp_list) return false;
	if (a.active && !a.tmap_gather) return false;
	if (a.pitch % 2 || a.pitch < 16 * nslot || a.mpitch != a.pitch) return false;
	const int npass = ceil_div(a.K, kt);
	if (npass > SL_MAX_PASSES || npass + 1 > xtile_counter_capacity()) return false;
	// 227 KB per SM, 1 KB of it reserved, a little static shared memory
	return slab_smem(a, kt, nslot) <= 225 * 1024;
}

template <int NC, int P, int NSLOT, bool GATHER>
static int launch_slab_inst(const LikeArgs &a, int sm_count, cudaStream_t st)
{
	constexpr int KT = NC * 8;
	const size_t smem = slab_smem(a, KT, NSLOT);
	auto kern = slab_dmma_kernel<NC, P, NSLOT, GATHER>;
	MDNS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CUtensorMap ta, tb;
	// the rows of THIS launch as (super-)rows of P * pitch doubles, boxes of 32 x 16 channels
	const long long nsup = ((long long)a.n_rows + P - 1) / P;
	int rc = MDNS_OK;
	if (GATHER)
		memcpy(&ta, a.tmap_gather, sizeof ta);      // one-row boxes of the whole shard
	else
		rc = make_row_tensor_map_box(&ta, a.Y, nsup, (long long)P * a.pitch, SL_ROWS);
	if (rc != MDNS_OK) return rc;
	const long long kpad = (long long)round_up(a.K, KT_MAX);
	rc = make_row_tensor_map_box(&tb, a.model, kpad, a.mpitch, KT);
	if (rc != MDNS_OK) return rc;
	const long long nslabs = (nsup + SL_ROWS - 1) / SL_ROWS;
	// one CTA per SM even when the slabs would fit in fewer: 16 warps on one SM share its FP64
	// tensor pipe, 16 warps on 16 SMs do not (1e5 x 200 on 98 CTAs: 0.075 ms)
	long long gx = nslabs < sm_count ? nslabs : sm_count;
	if (gx < 1) gx = 1;
	const int npass = ceil_div(a.K, KT);
	if (!a.xp_counters_clear)
		MDNS_CUDA(cudaMemsetAsync(a.xp_redo + 1, 0, (size_t)npass * sizeof(int), st));
	for (int k0 = 0, pass = 0; k0 < a.K; k0 += KT, ++pass) {
		const int kv = a.K - k0 < KT ? a.K - k0 : KT;
		// behind the model kernel that also reset the counters (first pass) or behind the
		// previous pass's fix-up launch: a programmatic dependent launch
		launch_pdl(kern, dim3((unsigned)gx), dim3(SL_THREADS), smem, st, pass > 0 || a.xp_counters_clear, ta, tb, a,
		           k0, kv, pass);
		MDNS_LAUNCHED(GATHER ? "slab_dmma_kernel(gather)" : "slab_dmma_kernel");
		// (the fix-up launch also hands the slab counter of this pass back at zero)
		rc = launch_xtile_fixup(a, k0, kv, pass, sm_count, st);
		if (rc != MDNS_OK) return rc;
	}
	return MDNS_OK;
}

int launch_slab_dmma(const LikeArgs &a, int kt, int nslot, int sm_count, cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	if (!slab_dmma_fits(a, kt, nslot)) {
		set_error("slab_dmma_kernel: needs the resident row sums (and the gather map for listed rows) and %zu bytes of shared memory",
		          slab_smem(a, kt, nslot));
		return MDNS_EINVAL;
	}
	const int P = slab_pair_mode(a);
#define MDNS_SL(KK, PP, SS)                                                                      \
	if (kt == KK && P == PP && nslot == SS) {                                                    \
		if (PP == 1 && a.active) return launch_slab_inst<KK / 8, 1, SS, true>(a, sm_count, st);  \
		return launch_slab_inst<KK / 8, PP, SS, false>(a, sm_count, st);                         \
	}
	MDNS_SL(16, 2, 3)
	MDNS_SL(16, 1, 3)
	MDNS_SL(16, 2, 2)
	MDNS_SL(16, 1, 2)
	MDNS_SL(8, 2, 3)
	MDNS_SL(8, 1, 3)
	MDNS_SL(8, 2, 2)
	MDNS_SL(8, 1, 2)
#undef MDNS_SL
	set_error("unsupported slab_dmma_kernel shape kt=%d slots=%d", kt, nslot);
	return MDNS_EINVAL;
}

}  // namespace mdns
