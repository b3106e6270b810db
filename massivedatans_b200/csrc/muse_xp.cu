// muse_xp.cu -- cmuselike.c:48-64 in expanded form, for batches of K model spectra.
//
//   chi_ik = sum_j (y_ij - s m_kj)^2 / v_ij ,  s = S1 / (1e-10 + S2)
//          = Swyy_i - 2 s S1_ik + s^2 S2_ik ,
//   S1_ik = sum_j (y_ij / v_ij) m_kj ,  S2_ik = sum_j (1 / v_ij) m_kj^2 ,  Swyy_i = sum_j y_ij^2 / v_ij
//
// S1 and S2 are two K x C x N contractions of RESIDENT matrices (y/v and 1/v, built once at
// upload) with the staged spectra and their squares: one launch of rows_dmma_kernel in raw mode
// streams both matrices once -- 16 bytes per element, the algorithmic minimum of cmuselike -- on
// the FP64 tensor path, whatever K is.  The direct form (muse_block_kernel) has to walk every row
// twice per candidate because s must be known before the residuals can be squared; for K
// candidates it re-streams the K spectra from L2 for every row pair in both passes (round 1:
// 0.36 of the HBM roofline at K = 4, 0.11 at K = 16 on the 4223 x 3600 cube).
//
// Accuracy.  The three terms cancel when a candidate fits high signal-to-noise data.  The raw
// contraction is summed in blocks of B = 64 channels (rows_dmma_kernel.cu), so each of the three
// sums carries at most about (B + C/B + 32) * 2^-53 relative to the sum of the magnitudes of its
// terms, and |2 s S1| <= Swyy + s^2 S2 (Cauchy-Schwarz); a result is kept only if
//     chi >= guard * (Swyy + s^2 S2) ,  guard = 2 (B + C/B + 32) 2^-53 / tol ,
// everything else (and every NaN) is recomputed on the spot in the direct two-pass form by the warp
// that owns the row (muse_xp_finalize_kernel).  Data whose candidates need that for more than 2 %
// of the rows switch the path off (host feedback), exactly like the clike expanded form.
#include "kernels.cuh"

namespace mdns {

// YW = Y o W and Swyy = sum W Y^2 per row, once at upload (one warp per row)
__global__ void __launch_bounds__(256) muse_prepare_kernel(const double *__restrict__ Y,
                                                           const double *__restrict__ W, long long n_rows,
                                                           long long pitch, int nfrag,
                                                           double *__restrict__ YW,
                                                           double *__restrict__ swyy)
{
	const int lane = threadIdx.x & 31;
	const long long warps = (long long)gridDim.x * 8;
	for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += warps) {
		const double2 *py = reinterpret_cast<const double2 *>(Y + r * pitch);
		const double2 *pw = reinterpret_cast<const double2 *>(W + r * pitch);
		double2 *po = reinterpret_cast<double2 *>(YW + r * pitch);
		double s0 = 0.0, s1 = 0.0;
		for (int f = lane; f < nfrag; f += 32) {
			const double2 y = __ldg(py + f), w = __ldg(pw + f);
			double2 o;
			o.x = y.x * w.x;
			o.y = y.y * w.y;
			po[f] = o;
			s0 = fma(o.x, y.x, s0);
			s1 = fma(o.y, y.y, s1);
		}
		double s = s0 + s1;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += shfl_xor_f64(s, o);
		if (lane == 0) swyy[r] = s;
	}
}

int launch_muse_prepare(const double *Y, const double *W, long long n_rows, long long pitch, int nx,
                        double *YW, double *swyy, cudaStream_t st)
{
	if (n_rows <= 0) return MDNS_OK;
	long long blocks = (n_rows + 7) / 8;
	if (blocks > 148 * 16) blocks = 148 * 16;
	muse_prepare_kernel<<<(unsigned)blocks, 256, 0, st>>>(Y, W, n_rows, pitch, (nx + 1) >> 1, YW, swyy);
	MDNS_LAUNCHED_HELPER("muse_prepare_kernel");
	return MDNS_OK;
}

// chi from the raw sums, one warp per row (lanes across the candidates); what the guard cannot
// vouch for is recomputed on the spot in the direct two-pass form (cmuselike.c:48-64 as written)
// by the same warp, lanes across the channels.  counter[0] accumulates the recomputed rows.
__global__ void __launch_bounds__(256) muse_xp_finalize_kernel(const LikeArgs a,
                                                               const double *__restrict__ S1,
                                                               const double *__restrict__ S2,
                                                               double guard, int *__restrict__ redo_total)
{
	// a programmatic dependent of the raw contraction: set up while it runs, S1 / S2 read after it
	pdl_wait();
	const int lane = threadIdx.x & 31;
	const long long warps = (long long)gridDim.x * 8;
	for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < a.n_rows; r += warps) {
		const int row = a.active ? a.active[r] : (int)r;
		const double sw = a.swyy[row];
		bool any_redo = false;
		for (int kb = 0; kb < a.K; kb += 32) {
			const int k = kb + lane;
			bool redo = false;
			if (k < a.K) {
				const double s1 = S1[(long long)k * a.n_rows + r], s2 = S2[(long long)k * a.n_rows + r];
				const double s = s1 / (s2 + 1e-10);
				const double t = s * s * s2;
				const double chi = sw + fma(-2.0 * s, s1, t);
				if (chi >= guard * (sw + t))          // false for NaN too
					a.out[(long long)k * a.out_stride + row] = -0.5 * chi;
				else
					redo = true;
			}
			unsigned todo = __ballot_sync(0xffffffffu, redo);
			any_redo = any_redo || todo != 0;
			const double *y = a.Y + (long long)row * a.pitch, *w = a.W + (long long)row * a.pitch;
			while (todo) {
				const int kk = kb + __ffs(todo) - 1;
				todo &= todo - 1;
				const double *m = a.model + (size_t)kk * a.mpitch;
				double s1 = 0.0, s2 = 0.0;
				for (int j = lane; j < a.nx; j += 32) {
					const double t = m[j] * w[j];
					s1 = fma(y[j], t, s1);
					s2 = fma(m[j], t, s2);
				}
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) {
					s1 += shfl_xor_f64(s1, o);
					s2 += shfl_xor_f64(s2, o);
				}
				const double s = s1 / (s2 + 1e-10);
				double chi = 0.0;
				for (int j = lane; j < a.nx; j += 32) {
					const double d = fma(-s, m[j], y[j]);
					chi = fma(d * d, w[j], chi);
				}
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) chi += shfl_xor_f64(chi, o);
				if (lane == 0) a.out[(long long)kk * a.out_stride + row] = -0.5 * chi;
			}
		}
		if (any_redo && lane == 0) atomicAdd(redo_total, 1);
	}
}

int launch_muse_xp_finalize(const LikeArgs &a, const double *S1, const double *S2, double guard,
                            int *redo_total, int sm_count, cudaStream_t st)
{
	if (a.n_rows <= 0) return MDNS_OK;
	int blocks = ceil_div(a.n_rows, 8);
	if (blocks > 8 * sm_count) blocks = 8 * sm_count;
	launch_pdl(muse_xp_finalize_kernel, dim3(blocks), dim3(256), 0, st, true, a, S1, S2, guard, redo_total);
	MDNS_LAUNCHED_HELPER("muse_xp_finalize_kernel");
	return MDNS_OK;
}

double muse_xp_guard(int nx, double tol)
{
	const double B = 64.0;     // RD_BLOCK_CH
	return 2.0 * (B + nx / B + 32.0) * 1.1102230246251565e-16 / tol;
}

}  // namespace mdns
