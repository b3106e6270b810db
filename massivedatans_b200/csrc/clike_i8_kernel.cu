// clike_i8_kernel.cu -- the north star's tcgen05 experiment: the cross term of the expanded
// chi-square,  Sym = Y * M^T  (N x C x K),  on the 5th-generation tensor cores.
//
// tcgen05 has no FP64 kind, and one TF32/BF16 product cannot hold the 1e-9 contract (the
// cancellation in Syy - 2 Sym + Smm amplifies the operand rounding).  What the tensor cores CAN do
// exactly is integer work: kind::i8 multiplies signed 8-bit operands into 32-bit integer
// accumulators in TMEM without any rounding.  So the FP64 operands are cut into digits
// (the "Ozaki scheme"):
//
//   y_ij = 2^ey_i * sum_{s=1..S} q_s(i,j) 2^-(7s-1) + tail ,  q_s in [-64, 64]   (rows scaled by a
//   m_kj = 2^em_k * sum_{t=1..S} q_t(k,j) 2^-(7t-1) + tail                         power of two)
//
//   Sym_ik = 2^(ey_i + em_k) * sum_{p=2..P} 2^-(7p-2) * G_p(i,k) ,
//   G_p(i,k) = sum_{s+t=p} sum_j q_s(i,j) q_t(k,j)         -- exact in int32
//
// With S = 7 digits and the pairs s + t <= P = 8 that is 28 integer MMAs per (128 data sets x 64
// candidates x 32 channels); the pairs of equal weight share one TMEM accumulator, 7 accumulators
// of 64 columns = 448 of the 512 TMEM columns.  What is dropped (pairs beyond P, the tails) is
// bounded by 4 C (S 2^-7(P-1) + 2^(1-7S)) (Syy + Smm) = 1.3e-11 (Syy + Smm) at C = 200, measured
// 6.5e-14 (tools/ozaki_emulate.py, the CPU emulation of exactly this arithmetic); the epilogue's
// guard keeps a result only if that bound is below the tolerance relative to chi2 and flags every
// other data set for the direct-form fix-up, like the FP64 tensor path.
//
// Warp roles (one CTA per SM, 320 threads): warps 0-7 epilogue (two per TMEM lane quarter:
// tcgen05.ld 32x32b, exact 64-bit recombination, guard, logL stores), warp 8 TMA producer,
// warp 9 TMEM allocation + MMA issue (one elected lane: tcgen05.mma.cta_group::1.kind::i8, operands
// through 128-byte-swizzled K-major shared-memory descriptors, completion signalled with
// tcgen05.commit on mbarriers).  Per 128-channel block the 7 candidate digit tiles (8 KB each) sit
// in one of two B slots, the 7 data digit tiles (16 KB each) stream through a ring of 6 A slots.
//
// The digit planes of the resident rows are built once (1 byte per digit and channel: 7/8 of the
// FP64 matrix, rows padded to 128 channels) and only when this path is asked for (tuning lanes = 5).
#include <cuda.h>

#include "kernels.cuh"

namespace mdns {

constexpr int I8_S = 7;                 // digits per operand
constexpr int I8_P = 8;                 // pairs s + t <= P
constexpr int I8_NACC = I8_P - 1;       // accumulators (weights p = 2..P)
constexpr int I8_M = 128;               // data sets per tile (UMMA M)
constexpr int I8_N = 64;                // candidates per tile (UMMA N)
constexpr int I8_KB = 128;              // channels (bytes) per block = one 128-byte swizzle row
constexpr int I8_A_BYTES = I8_M * I8_KB;    // 16 KB
constexpr int I8_B_BYTES = I8_N * I8_KB;    // 8 KB
constexpr int I8_A_SLOTS = 6;
constexpr int I8_B_SLOTS = 2;
constexpr int I8_EPI_WARPS = 8;             // two per TMEM lane quarter: 32 of the 64 columns each
constexpr int I8_THREADS = (I8_EPI_WARPS + 2) * 32;
constexpr size_t I8_SMEM = (size_t)I8_A_SLOTS * I8_A_BYTES + (size_t)I8_B_SLOTS * I8_S * I8_B_BYTES + 1024;

// ---- digit planes -----------------------------------------------------------------------------
// One warp per row: exponent from the largest magnitude, then S signed 7-bit digits per element
// (all operations exact: scalings by powers of two, round to nearest, differences of neighbours).
// planes[s][row][j] int8, row-major with `cp` bytes per row (zero beyond nx); scale[row] = 2^e.
__global__ void __launch_bounds__(256) i8_split_kernel(const double *__restrict__ rows, long long n_rows,
                                                       long long pitch, int nx, int8_t *__restrict__ planes,
                                                       long long plane_rows, int cp,
                                                       double *__restrict__ scale,
                                                       uint8_t *__restrict__ clear_flags, long long nflags,
                                                       int *__restrict__ zero_word)
{
	if (zero_word && blockIdx.x == 0 && threadIdx.x == 0) *zero_word = 0;
	const int lane = threadIdx.x & 31;
	const long long warps = (long long)gridDim.x * 8;
	for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += warps) {
		const double *p = rows + r * pitch;
		double amax = 0.0;
		for (int j = lane; j < nx; j += 32) amax = fmax(amax, fabs(p[j]));
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, shfl_xor_f64(amax, o));
		int e = 0;
		if (amax > 0.0 && amax < 1e300) {
			frexp(amax, &e);          // amax = f 2^e, f in [0.5, 1)
			e += 1;                   // |y| 2^-e < 0.5
		}
		const double inv = ldexp(1.0, -e);
		if (lane == 0) scale[r] = ldexp(1.0, e);
		for (int j = lane; j < cp; j += 32) {
			double v = j < nx ? p[j] * inv * 64.0 : 0.0;
			if (!(fabs(v) <= 64.0)) v = 0.0;       // NaN / inf rows: the guard sends them to the fix-up
#pragma unroll
			for (int s = 0; s < I8_S; ++s) {
				const double q = rint(v);
				planes[((long long)s * plane_rows + r) * cp + j] = (int8_t)(int)q;
				v = (v - q) * 128.0;
			}
		}
	}
	if (clear_flags)
		for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nflags; i += (long long)gridDim.x * 256)
			clear_flags[i] = 0;
}

int launch_i8_split(const double *rows, long long n_rows, long long pitch, int nx, int8_t *planes,
                    long long plane_rows, int cp, double *scale, uint8_t *clear_flags, long long nflags,
                    int *zero_word, cudaStream_t st)
{
	long long blocks = (n_rows + 7) / 8;
	if (clear_flags && blocks < 148 * 4) blocks = 148 * 4;      // enough threads to clear the flags quickly
	if (blocks > 148 * 16) blocks = 148 * 16;
	if (blocks < 1) blocks = 1;
	i8_split_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, n_rows, pitch, nx, planes, plane_rows, cp, scale,
	                                                   clear_flags, nflags, zero_word);
	MDNS_LAUNCHED_HELPER("i8_split_kernel");
	return MDNS_OK;
}

// ---- tcgen05 helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t i8_smem_desc(uint32_t smem_addr)
{
	// K-major operand tile in the canonical 128-byte-swizzle layout (rows of 128 bytes, 8-row
	// groups 1024 bytes apart): start address >> 4, LBO (unused when swizzled) = 1, SBO = 1024 >> 4,
	// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
	return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
	       ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// D[tmem] (+)= A[smem] * B[smem]^T, 128 x 64 x 32, signed 8-bit operands, 32-bit integer accumulation
__device__ __forceinline__ void i8_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                       uint32_t accumulate)
{
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
	    "}\n" ::"r"(tmem_d),
	    "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
	    : "memory");
}

// the same with the accumulate flag known at compile time
__device__ __forceinline__ void i8_mma_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc)
{
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "setp.eq.b32 p, 0, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%4, %4, %4, %4}, p;\n\t"
	    "}\n" ::"r"(tmem_d),
	    "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(0u)
	    : "memory");
}

__device__ __forceinline__ void i8_commit(uint64_t *bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
	                 smem_u32(bar))
	             : "memory");
}

__device__ __forceinline__ void i8_tma_load(void *smem_dst, const CUtensorMap *tmap, int c0, int c1,
                                            uint64_t *bar)
{
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
	    "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
	    "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
	    : "memory");
}

__device__ __forceinline__ void i8_tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
	      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
	      "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
	      "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
	      "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr)
	    : "memory");
}

__device__ __forceinline__ void i8_tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
	      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
	      "=r"(r[15])
	    : "r"(taddr)
	    : "memory");
}

struct I8Args {
	int n_rows, K, nct, nrt, nkb;
	long long plane_rows_y;       // rows per digit plane of the data (multiple of 128)
	int plane_rows_m;             // rows per digit plane of the batch (multiple of 64)
	const double *scale_y, *scale_m, *syy, *smm;
	double *out;
	long long out_stride;
	uint8_t *flags;               // rows the guard could not vouch for
	double guard, inv;
	int dbg;                      // measurement knob (MDNS_I8_DBG): 1 = no recombination/stores, 2 = no TMEM reads either
};

__global__ void __launch_bounds__(I8_THREADS, 1) clike_i8_kernel(const __grid_constant__ CUtensorMap tmapY,
                                                                 const __grid_constant__ CUtensorMap tmapM,
                                                                 const I8Args a)
{
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	__shared__ uint64_t full_a[I8_A_SLOTS], empty_a[I8_A_SLOTS], full_b[I8_B_SLOTS], empty_b[I8_B_SLOTS];
	__shared__ uint64_t tmem_full, tmem_empty;
	__shared__ uint32_t s_tmem;
	unsigned char *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	unsigned char *smem_a = base;
	unsigned char *smem_b = base + (size_t)I8_A_SLOTS * I8_A_BYTES;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int ntiles = a.nrt * a.nct;

	if (threadIdx.x == 0) {
		for (int i = 0; i < I8_A_SLOTS; ++i) {
			mbar_init(&full_a[i], 1);
			mbar_init(&empty_a[i], 1);
		}
		for (int i = 0; i < I8_B_SLOTS; ++i) {
			mbar_init(&full_b[i], 1);
			mbar_init(&empty_b[i], 1);
		}
		mbar_init(&tmem_full, 1);
		mbar_init(&tmem_empty, I8_EPI_WARPS);
		mbar_fence_init();
	}
	if (warp == I8_EPI_WARPS + 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
		             "r"(512u)
		             : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = s_tmem;

	if (warp == I8_EPI_WARPS) {
		// ===================== TMA producer =====================
		if (lane == 0) {
			int ia = 0, ib = 0;
			for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
				const int rt = tile / a.nct, ct = tile - rt * a.nct;
				for (int kb = 0; kb < a.nkb; ++kb, ++ib) {
					const int sb = ib % I8_B_SLOTS;
					mbar_wait(&empty_b[sb], ((ib / I8_B_SLOTS) & 1) ^ 1);
					mbar_expect_tx(&full_b[sb], I8_S * I8_B_BYTES);
					for (int t = 0; t < I8_S; ++t)
						i8_tma_load(smem_b + ((size_t)sb * I8_S + t) * I8_B_BYTES, &tmapM, kb * I8_KB,
						            t * a.plane_rows_m + ct * I8_N, &full_b[sb]);
					for (int s = 0; s < I8_S; ++s, ++ia) {
						const int sa = ia % I8_A_SLOTS;
						mbar_wait(&empty_a[sa], ((ia / I8_A_SLOTS) & 1) ^ 1);
						mbar_expect_tx(&full_a[sa], I8_A_BYTES);
						i8_tma_load(smem_a + (size_t)sa * I8_A_BYTES, &tmapY, kb * I8_KB,
						            (int)(s * a.plane_rows_y + (long long)rt * I8_M), &full_a[sa]);
					}
				}
			}
		}
	} else if (warp == I8_EPI_WARPS + 1) {
		// ===================== MMA issuer =====================
		if (lane == 0) {
			// instruction descriptor: D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), both
			// K-major, N = 64 (>> 3 at bit 17), M = 128 (>> 4 at bit 24)
			const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_N >> 3) << 17) |
			                       ((uint32_t)(I8_M >> 4) << 24);
			int ia = 0, ib = 0, it = 0;
			for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
				mbar_wait(&tmem_empty, (it & 1) ^ 1);      // the epilogue has drained the accumulators
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				for (int kb = 0; kb < a.nkb; ++kb, ++ib) {
					const int sb = ib % I8_B_SLOTS;
					mbar_wait(&full_b[sb], (ib / I8_B_SLOTS) & 1);
					// descriptors differ only in the start-address field (bytes >> 4): everything below
					// is additions of compile-time constants, the single issuing thread must not be
					// the bottleneck (a first version built every descriptor from scratch: 125 cycles
					// per MMA against a 32-cycle floor)
					const uint64_t db0 = i8_smem_desc(smem_u32(smem_b + (size_t)sb * I8_S * I8_B_BYTES));
#pragma unroll
					for (int s = 0; s < I8_S; ++s, ++ia) {
						const int sa = ia % I8_A_SLOTS;
						mbar_wait(&full_a[sa], (ia / I8_A_SLOTS) & 1);
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						const uint64_t da0 = i8_smem_desc(smem_u32(smem_a + (size_t)sa * I8_A_BYTES));
						// digit s+1 of the data pairs with the digits t+1 <= P - (s+1) of the batch
#pragma unroll
						for (int t = 0; t < I8_S; ++t) {
							if (t + s + 2 > I8_P) continue;
							const uint32_t d = tmem + (uint32_t)((s + t) * I8_N);     // accumulator p - 2
#pragma unroll
							for (int ks = 0; ks < I8_KB / 32; ++ks) {
								const uint64_t da = da0 + (uint64_t)(ks * 2);
								const uint64_t db = db0 + (uint64_t)(t * (I8_B_BYTES >> 4) + ks * 2);
								if (s == 0 && ks == 0)
									i8_mma(d, da, db, idesc, kb != 0 ? 1u : 0u);
								else
									i8_mma_acc(d, da, db, idesc);
							}
						}
						i8_commit(&empty_a[sa]);       // arrives when the MMAs above have read the slot
					}
					i8_commit(&empty_b[sb]);
				}
				i8_commit(&tmem_full);
			}
		}
	} else {
		// ===================== epilogue warps 0..7 =====================
		// warp w reads TMEM lanes 32 (w % 4) .. +31 (its rows) and the columns 32 (w / 4) .. +31 of
		// every accumulator (its candidates).  The seven integer sums of a (row, candidate) pair are
		// recombined exactly in two 64-bit integers (weights 2^-26 and 2^-54 relative to the digit
		// scale) -- two int->double conversions per result instead of seven.
		const int q = warp & 3, half = warp >> 2;
		int it = 0;
		for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
			const int rt = tile / a.nct, ct = tile - rt * a.nct;
			const long long row = (long long)rt * I8_M + q * 32 + lane;
			const bool live = row < a.n_rows;
			const double sy = live ? a.scale_y[row] : 0.0;
			const double syy = live ? a.syy[row] : 0.0;
			// this lane's candidate of the warp's 32: broadcast by shuffle when its column comes up
			const int kmine = ct * I8_N + half * 32 + lane;
			const double smm_l = kmine < a.K ? __ldg(a.smm + kmine) : 0.0;
			const double sm_l = kmine < a.K ? __ldg(a.scale_m + kmine) : 0.0;
			mbar_wait(&tmem_full, it & 1);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			bool redo = false;
			if (a.dbg == 2) {
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty)) : "memory");
				continue;
			}
#pragma unroll 1
			for (int c16 = 0; c16 < 2; ++c16) {
				uint32_t g[I8_NACC][16];
				const uint32_t col0 = (uint32_t)(half * 32 + c16 * 16);
#pragma unroll
				for (int p = 0; p < I8_NACC; ++p)
					i8_tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(p * I8_N) + col0, g[p]);
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				if (c16 == 1) {
					// everything this warp needs is in registers: hand TMEM back to the MMA warp
					asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
					__syncwarp();
					if (lane == 0)
						asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty)) : "memory");
				}
				if (a.dbg == 1) {
					if ((g[0][0] ^ g[3][7] ^ g[6][15]) == 0x7fffffffu) redo = true;      // keep the loads alive
					continue;
				}
#pragma unroll
				for (int j = 0; j < 16; ++j) {
					const int kl = half * 32 + c16 * 16 + j, k = ct * I8_N + kl;
					// accumulators p = 0..6 carry the weights 2^-12, 2^-19, ..., 2^-54
					const long long ihi = ((long long)(int)g[0][j] << 14) + ((long long)(int)g[1][j] << 7) + (int)g[2][j];
					const long long ilo = ((long long)(int)g[3][j] << 21) + ((long long)(int)g[4][j] << 14) +
					                      ((long long)(int)g[5][j] << 7) + (int)g[6][j];
					const double s26 = __longlong_as_double((long long)(1023 - 26) << 52);
					const double s54 = __longlong_as_double((long long)(1023 - 54) << 52);
					const double dot = fma((double)ilo, s54, (double)ihi * s26);
					const double smm = __shfl_sync(0xffffffffu, smm_l, c16 * 16 + j);
					const double sm = __shfl_sync(0xffffffffu, sm_l, c16 * 16 + j);
					if (k < a.K) {                              // uniform over the warp
						const double sym = dot * (sy * sm);
						const double chi = syy + fma(-2.0, sym, smm);
						const bool ok = chi >= a.guard * (syy + smm);   // false for NaN too
						if (live) {
							if (ok)
								a.out[(long long)k * a.out_stride + row] = chi * a.inv;
							else
								redo = true;
						}
					}
				}
			}
			if (redo) a.flags[row] = 1;
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == I8_EPI_WARPS + 1) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
	}
}

// rows the guard flagged -> list (16 flags per thread, one 128-bit load); the direct-form
// recomputation of the listed rows is xtile_fixup_kernel's, shared with the FP64 tensor path
__global__ void __launch_bounds__(256) i8_collect_kernel(const uint8_t *__restrict__ flags, long long n_rows,
                                                         int *__restrict__ list, int *__restrict__ list_len)
{
	const long long base = ((long long)blockIdx.x * 256 + threadIdx.x) * 16;
	if (base >= n_rows) return;
	const uint4 w = *reinterpret_cast<const uint4 *>(flags + base);      // the buffer is padded to 128 rows
	if ((w.x | w.y | w.z | w.w) == 0) return;
	const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
	for (int q = 0; q < 4; ++q)
#pragma unroll
		for (int b = 0; b < 4; ++b)
			if (((words[q] >> (8 * b)) & 0xffu) && base + 4 * q + b < n_rows)
				list[atomicAdd(list_len, 1)] = (int)(base + 4 * q + b);
}

int launch_xtile_fixup(const LikeArgs &a, int k0, int kv, int pass, int sm_count, cudaStream_t st);

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnI8)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static int i8_tensor_map(CUtensorMap *tm, const int8_t *planes, long long rows, int cp, int box_rows)
{
	static EncodeTiledFnI8 encode = nullptr;
	if (!encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult q;
		MDNS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
		if (!fn || q != cudaDriverEntryPointSuccess) {
			set_error("cuTensorMapEncodeTiled is not available from this driver");
			return MDNS_ECUDA;
		}
		encode = (EncodeTiledFnI8)fn;
	}
	const cuuint64_t dims[2] = {(cuuint64_t)cp, (cuuint64_t)rows};
	const cuuint64_t strides[1] = {(cuuint64_t)cp};
	const cuuint32_t box[2] = {(cuuint32_t)I8_KB, (cuuint32_t)box_rows};
	const cuuint32_t estr[2] = {1, 1};
	const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)planes, dims, strides, box, estr,
	                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
	                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("cuTensorMapEncodeTiled (digit planes) failed with code %d", (int)r);
		return MDNS_ECUDA;
	}
	return MDNS_OK;
}

int i8_plane_pitch(int nx) { return (int)round_up(nx, I8_KB); }
long long i8_plane_rows(long long n) { return (long long)round_up((size_t)n, I8_M); }
int i8_batch_rows(int K) { return (int)round_up(K, I8_N); }
int i8_digits() { return I8_S; }
double i8_guard(int nx, double tol)
{
	const double bound = 4.0 * nx * (I8_S * ldexp(1.0, -7 * (I8_P - 1)) + ldexp(1.0, 1 - 7 * I8_S));
	return bound / tol + (2.0 * nx + 4.0) * 1.1102230246251565e-16 / tol;
}

// a: the usual arguments (all rows active); planes_y / scale_y: resident digit planes of the rows;
// planes_m / scale_m: scratch for the digit planes of the staged batch (built here)
int launch_clike_i8(const LikeArgs &a, const int8_t *planes_y, const double *scale_y, long long plane_rows_y,
                    int8_t *planes_m, double *scale_m, uint8_t *flags, double tol, int sm_count,
                    cudaStream_t st)
{
	if (a.n_rows <= 0 || a.K <= 0) return MDNS_OK;
	if (a.active) {
		set_error("the tcgen05 experiment runs on all-active rows only");
		return MDNS_EINVAL;
	}
	const int cp = i8_plane_pitch(a.nx);
	const int mrows = i8_batch_rows(a.K);
	// digit planes of the staged batch (+ reset of the row flags)
	int rc = launch_i8_split(a.model, mrows < (int)round_up(a.K, KT_MAX) ? mrows : (long long)round_up(a.K, KT_MAX),
	                         a.mpitch, a.nx, planes_m, mrows, cp, scale_m, flags, a.n_rows, a.xp_redo + 1, st);
	if (rc != MDNS_OK) return rc;
	CUtensorMap ty, tm;
	if ((rc = i8_tensor_map(&ty, planes_y, (long long)I8_S * plane_rows_y, cp, I8_M)) != MDNS_OK) return rc;
	if ((rc = i8_tensor_map(&tm, planes_m, (long long)I8_S * mrows, cp, I8_N)) != MDNS_OK) return rc;
	I8Args g;
	g.n_rows = a.n_rows;
	g.K = a.K;
	g.nct = mrows / I8_N;
	g.nrt = ceil_div(a.n_rows, I8_M);
	g.nkb = cp / I8_KB;
	g.plane_rows_y = plane_rows_y;
	g.plane_rows_m = mrows;
	g.scale_y = scale_y + a.row0;
	g.scale_m = scale_m;
	g.syy = a.syy + a.row0;
	g.smm = a.smm;
	g.out = a.out;
	g.out_stride = a.out_stride;
	g.flags = flags;
	g.guard = i8_guard(a.nx, tol);
	g.inv = a.scale / a.noise2;
	{
		const char *e = getenv("MDNS_I8_DBG");
		g.dbg = e ? atoi(e) : 0;
	}
	MDNS_CUDA(cudaFuncSetAttribute(clike_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8_SMEM));
	long long gx = (long long)g.nrt * g.nct;
	if (gx > sm_count) gx = sm_count;
	clike_i8_kernel<<<(unsigned)gx, I8_THREADS, I8_SMEM, st>>>(ty, tm, g);
	MDNS_LAUNCHED("clike_i8_kernel");
	// flagged rows -> list -> direct-form recomputation of all K candidates of those rows
	i8_collect_kernel<<<ceil_div(ceil_div(a.n_rows, 16), 256), 256, 0, st>>>(flags, a.n_rows, a.xp_list,
	                                                                         a.xp_redo + 1);
	MDNS_LAUNCHED_HELPER("i8_collect_kernel");
	LikeArgs f = a;
	f.lmins = nullptr;
	f.counts = nullptr;
	if ((rc = launch_xtile_fixup(f, 0, a.K, 0, sm_count, st)) != MDNS_OK) return rc;
	return MDNS_OK;
}

}  // namespace mdns
