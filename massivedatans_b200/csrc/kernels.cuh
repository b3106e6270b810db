// kernels.cuh -- host-callable launchers of the sm_100a kernels (internal).
#pragma once
#include "common.cuh"

namespace mdns {

// ---------------------------------------------------------------- layout ---
// Resident layout ("data-set-major rows"): data set i of a shard owns the row
//   Y[i*pitch .. i*pitch + nx)   (doubles), pitch = round_up(nx, 2)
// i.e. every row starts on a 16-byte boundary (128-bit loads, bulk-TMA) and an odd
// channel count is padded with one zero.  No wider padding: ncu showed that padding
// rows to 128 bytes is fetched from DRAM anyway (1664 instead of 1600 bytes per row).
// The host matrix is channel-major (clike.c:72: yy[i + j*ndata]); it is
// transposed once at upload.  Model spectra use the same pitch and padding.
constexpr int ROW_ALIGN = 2;    // doubles (16 bytes)
constexpr int XP_MIN_K = 3;     // smallest all-active batch that takes the expanded form by itself
constexpr int XP_MIN_K_MASKED = 5;   // and the smallest masked one (gather4-fed tensor path)
constexpr int MUSE_XP_MIN_K = 3; // smallest batch of spectra that takes the expanded cmuselike form
constexpr int KT_MAX = 32;      // model buffers are padded to a multiple of this many candidates

// in: channel-major chunk in[j*ld_in + c], j < nx, c < nb  (device staging)
// out: rows out[c*pitch + j]; recip != 0 stores 1/v (inverse variance)
int launch_transpose_rows(const double *in, size_t ld_in, int nx, int nb, double *out,
                          size_t pitch, int recip, cudaStream_t st);

// clike.c:65 -- model[k*mpitch + j] = A_k*exp(-0.5*((mu_k - x_j)/sig_k)^2); zero padding
// for j >= nx and for k in [K, Kpad).
// smm (may be null): sum of squares of every model row; counters (may be null): the expanded
// kernels' counter block, whose per-pass entries [1, 1+ncounters) are reset.
int launch_line_model(const double *x, int nx, const double *params, int K, int Kpad,
                      double *model, int mpitch, double *smm, int *counters, int ncounters,
                      cudaStream_t st);
// zero-pad host-provided spectra: src[K][nx] -> model[Kpad][mpitch]
int launch_pad_spectra(const double *src, int nx, int K, int Kpad, double *model, int mpitch,
                       cudaStream_t st);

// mask bytes -> ordered list of active row indices (stable compaction; the rank
// of i in the mask is the output slot k of clike.c:67-74).
// scratch must hold ceil(n/4096)+1 ints.
int launch_compact_mask(const uint8_t *mask, int n, int *scratch, int *active, int *n_active_dev,
                        cudaStream_t st);

struct LikeArgs {
	const double *Y = nullptr;      // rows
	const double *W;      // inverse-variance rows (muse) or nullptr
	long long pitch;      // doubles
	int nx;
	const int *active;    // nullptr = identity
	int n_rows;           // number of active rows
	const double *model;  // [Kpad][mpitch]
	int mpitch;
	int K;
	double noise2;        // noise^2 (clike)
	double scale;         // clike: multiplies the chi-square sum
	double *out;          // clike: [K][out_stride] compacted; muse: [K][out_stride] by row index
	long long out_stride;
	const void *tmap;     // host copies of the rows' CUtensorMaps for 128- and 256-row tiles
	const void *tmap256;  // (tile kernel) or nullptr
	const void *tmap_gather;  // one-row boxes of the whole shard for tile::gather4, or nullptr
	int row0;             // first row of this launch within the shard (tile kernel coordinates)
	// one candidate given by value (K = 1 fast path of clike_rows_kernel): the kernel builds the
	// model spectrum itself, no parameter upload and no model kernel
	int inline_model;
	const double *x;      // wavelength grid
	double line_A, line_mu, line_sig;
	// small batches (clike_small_kernel): the staged parameter points (A, mu, sig) on the device;
	// every CTA builds the spectra itself, no model kernel ran.  nullptr = spectra in `model`
	const double *params = nullptr;
	// expanded form (clike_xtile_kernel): Syy - 2 Sym + Smm
	const double *syy;    // resident sum of squares of every row of the shard, or nullptr
	const double *smm;    // [Kpad] sum of squares of every model spectrum
	double xp_guard;      // keep a result iff chi2 >= xp_guard*(Syy+Smm), else direct form
	int *xp_redo;         // device counters: [0] rows recomputed in the direct form so far,
	                      // [1 + pass] list length of each pass of the current launch
	int *xp_list;         // rows (within the launch) to recompute, capacity = rows of the shard
	bool xp_counters_clear;  // the per-pass counters were already reset by the model kernel
	// accept test fused into the likelihood epilogues (hiermetriclearn.py:193 `L > Lmins`):
	// counts[k] += number of rows of this launch whose value exceeds lmins[row]
	const double *lmins = nullptr;   // [n_rows], aligned with the launch's (compacted) rows
	int *counts = nullptr;           // [Kpad] device counters (zeroed by the caller)
	// stream-K partial sums of rows_dmma_kernel: workspace slots and one ticket per tile
	double *ws = nullptr;
	int *tickets = nullptr;
	// second (rows, batch) pair of the raw contraction (MUSE: 1/v rows against squared spectra)
	const void *tmap256_b = nullptr;
	const void *tmap_gather_b = nullptr;
	const double *model_b = nullptr;
	double *out_b = nullptr;
	const double *swyy = nullptr;    // MUSE expanded form: resident sum of y^2/v per row
};

struct Tuning {
	bool allow_expanded = true;   // false: automatic choice never takes the expanded form
	int lanes = 0;   // lanes per data set: 8 or 32; 1 = tile kernel (unroll = channels per
	                 // stage 16/32, rows = ring stages 3/4/6, all-active only), 2 = expanded-form
	                 // tile kernel (unroll = data sets per lane 2/4, rows = ring stages 2/3,
	                 // all-active only), 3 = expanded form on the FP64 tensor path (DMMA;
	                 // ktile 8/16/32, rows = ring stages 2/3/4); 6 = per-warp slabs on the tensor
	                 // path; 7 = small batches in one launch (clike_small_kernel); 0 = auto
	int unroll = 0;  // 128-bit fragments in flight per lane (0 = auto)
	int ktile = 0;   // candidates per pass (0 = auto)
	int rows = 0;    // data sets per lane group in the block kernel: 1, 2, 4 (0 = auto)
};

// *accept_fused (may be null): set to 1 when the chosen kernel applied a.lmins / a.counts itself
int launch_clike(const LikeArgs &a, const Tuning &t, int sm_count, cudaStream_t st,
                 int *accept_fused = nullptr);
int launch_muse(const LikeArgs &a, const Tuning &t, int sm_count, cudaStream_t st);
// small batches in one launch (clike_small_kernel): candidates per slice, shared-memory fit
int clike_small_ktile(int K);
bool clike_small_fits(int K, int mpitch);
// one speculative pass over a handful of active data sets in ONE launch: candidates by value,
// decision and the accepted vector written to pinned host memory (likelihood_kernels.cu)
bool draw_small_fits(const LikeArgs &a);
size_t draw_small_host_bytes();
int launch_draw_small(const LikeArgs &a, const double *params, int seq, int *host_block, cudaStream_t st);
// counts[k] = #{r : L[k*stride + r] > lmins[r]}  (hiermetriclearn.py:193 on the device)
int launch_accept_count(const double *L, long long stride, int n, int K, const double *lmins,
                        int *counts, cudaStream_t st, bool zero = true);
// the accept decision on the device (see likelihood_kernels.cu): layout of the `sel` block
constexpr int SEL_FIRST = 0, SEL_COUNT = 1, SEL_REDO = 2, SEL_COUNTS = 4;
int launch_select_first(const int *counts, int K, const int *redo, int *sel, cudaStream_t st);
int launch_gather_selected(const double *L, long long stride, int r0, int n, const int *sel,
                           double *out, cudaStream_t st);
int launch_selected_flags(const double *L, long long stride, int n, const int *sel,
                          const double *lmins, uint8_t *flags, cudaStream_t st);
int launch_gather_selected_values(const double *L, long long stride, const int *sel, const int *idx,
                                  const int *n_dev, int n_max, double *out, cudaStream_t st);
// sparse accept: flags[r] = L[r] > lmins[r]; out[i] = L[idx[i]]
int launch_accept_flags(const double *L, int n, const double *lmins, uint8_t *flags, cudaStream_t st);
int launch_gather_values(const double *L, const int *idx, int n, double *out, cudaStream_t st);
// MUSE-type, one CTA per data set, rows staged once through a bulk-TMA ring (muse_block_kernel.cu)
bool muse_block_fits(const LikeArgs &a);
// groups: 1 or 2 data sets in flight per CTA (0 = automatic)
int launch_muse_block(const LikeArgs &a, int ktile, int groups, int sm_count, cudaStream_t st);
// lane-per-data-set kernel fed by a bulk-TMA ring (clike_tile_kernel.cu)
int launch_clike_tile(const LikeArgs &a, const void *tmap, int kt, int nbox, int stages,
                      int tile_rows, int sm_count, cudaStream_t st);
// expanded form Syy - 2 Sym + Smm, register-blocked over data sets (clike_xtile_kernel.cu);
// kt in {8, 16, 32}, lane_rows in {2, 4}, stages in {2, 3}
bool xtile_fits(const LikeArgs &a, int kt, int stages);
int xtile_counter_capacity();
// per-warp slabs with the candidate batch resident in shared memory (slab_dmma_kernel.cu): short
// spectra, all rows active; kt in {8, 16}, nslot in {2, 3}.  Its slab counters live behind the
// fix-up counters: [slab_counter_base(), + slab_counter_count())
bool slab_dmma_fits(const LikeArgs &a, int kt, int nslot);
int launch_slab_dmma(const LikeArgs &a, int kt, int nslot, int sm_count, cudaStream_t st);
long long slab_dmma_slabs(const LikeArgs &a);   // slabs this launch would be cut into
int slab_counter_base();
int slab_counter_count();
// the same expanded form with the cross term on the FP64 tensor path (clike_dmma_kernel.cu);
// kt in {8, 16, 32}, stages in {2, 3, 4}
bool dmma_fits(const LikeArgs &a, int kt, int stages);
int launch_clike_dmma(const LikeArgs &a, int kt, int stages, int sm_count, cudaStream_t st);
int launch_clike_xtile(const LikeArgs &a, int kt, int lane_rows, int stages, int sm_count,
                       cudaStream_t st);
// stream-K contraction on the FP64 tensor path (rows_dmma_kernel.cu): the clike epilogue with the
// fused accept test, or (raw) the contraction(s) themselves for the MUSE expanded form
bool rows_dmma_fits(const LikeArgs &a, int kt, int stages);
size_t rows_dmma_workspace_doubles(int sm_count);
int launch_rows_dmma(const LikeArgs &a, int kt, int stages, bool raw, int nmat, int sm_count,
                     cudaStream_t st);
// cmuselike in expanded form (muse_xp.cu): resident y/v rows + sum y^2/v per row at upload; per
// batch rows_dmma_kernel in raw mode (S1, S2; the second contraction squares the spectra on the
// fly) and the finalize kernel (chi, guard, direct-form recomputation in place)
int launch_muse_prepare(const double *Y, const double *W, long long n_rows, long long pitch, int nx,
                        double *YW, double *swyy, cudaStream_t st);
int launch_muse_xp_finalize(const LikeArgs &a, const double *S1, const double *S2, double guard,
                            int *redo_total, int sm_count, cudaStream_t st);
double muse_xp_guard(int nx, double tol);
// the tcgen05 experiment (clike_i8_kernel.cu): cross term on the INT8 tensor path from 7-bit digit
// planes of the FP64 operands (exact integer products, FP64 recombination)
int i8_plane_pitch(int nx);
long long i8_plane_rows(long long n);
int i8_batch_rows(int K);
int i8_digits();
double i8_guard(int nx, double tol);
int launch_i8_split(const double *rows, long long n_rows, long long pitch, int nx, int8_t *planes,
                    long long plane_rows, int cp, double *scale, uint8_t *clear_flags, long long nflags,
                    int *zero_word, cudaStream_t st);
int launch_clike_i8(const LikeArgs &a, const int8_t *planes_y, const double *scale_y, long long plane_rows_y,
                    int8_t *planes_m, double *scale_m, uint8_t *flags, double tol, int sm_count,
                    cudaStream_t st);
// out[r] = sum_j rows[r*pitch + j]^2 (rows: resident data sets or padded model spectra)
int launch_row_sumsq(const double *rows, long long n_rows, long long pitch, int nx, double *out,
                     cudaStream_t st);
// 128-byte CUtensorMap describing the resident rows Y[n_rows][pitch] (tile kernel)
int make_row_tensor_map(void *out, const double *Y, long long n_rows, long long pitch,
                        int tile_rows);
int tile_constant_capacity();

// ------------------------------------------------------------ neighbours ---
// members in SoA layout xs[k*npad + i]
// ticket (device int, zero) / host_flag (pinned) / seq: the last CTA to finish writes seq to
// *host_flag after every count is visible to the host (small calls polled by the host); nullptr:
// an ordinary launch
int launch_count_within(const double *xs, int n, int npad, int ndim, const double *yy, int m,
                        double T, int stop_at, int *counts, int sm_count, cudaStream_t st,
                        int *ticket = nullptr, int *host_flag = nullptr, int seq = 0);
// m ball draws of RadFriendsRegion.generate fused with the neighbour count: points[m][ndim],
// keep[m] (accepted with probability 1/nnear), nnear[m]; Philox keyed by (seed, first + j)
int launch_region_generate(const double *xs, int n, int npad, int ndim, double r, double T,
                           unsigned long long seed, unsigned long long first, int m, double *points,
                           uint8_t *keep, int *nnear, cudaStream_t st);
int launch_gather_points(const double *points, int ndim, const int *idx, int n, double *out,
                         cudaStream_t st);
// per-axis SupFriends distance (neighbors.py:22-73): nearest other member of every member; and
// "is some listed reference member inside the per-axis box md around the listed query member"
int launch_nn_index(const double *xs, int n, int npad, int ndim, int *nearest, cudaStream_t st);
int launch_axis_covered(const double *xs, int npad, int ndim, const double *md, const int *query,
                        int nq, const int *ref, int nr, uint8_t *covered, cudaStream_t st);
int launch_within_single(const double *xs, int n, int npad, int ndim, const double *y, double T,
                         int *flag, cudaStream_t st);
// chosen[n][nboot] doubles -> per round: query list (un-chosen i) and reference list (chosen j).
// lists: qidx[b*n + p], ridx[b*n + p]; counts[b] (queries), counts[nboot + b] (references);
// nearest[b*n + p] is initialised to the bit pattern of 1e300 (cneighbors.c:148).
int launch_bootstrap_lists(const double *chosen, int n, int nboot, int *qidx, int *ridx,
                           int *counts, unsigned long long *nearest, cudaStream_t st);
// mode 1: one round, every sample queries every other sample (cneighbors.c:50-64).
int launch_all_pairs_lists(int n, int *qidx, int *ridx, int *counts, unsigned long long *nearest,
                           cudaStream_t st);
int launch_nn_min(const double *xs, int n, int npad, int ndim, const int *qidx, const int *ridx,
                  const int *counts, int nrounds, int exclude_self, unsigned long long *nearest,
                  int sm_count, cudaStream_t st);
// max over rounds and queries of nearest (skip_first: ignore original sample 0, cneighbors.c:162)
int launch_nn_finalize(const int *qidx, const int *counts, int n, int nrounds, int skip_first,
                       const unsigned long long *nearest, double *result, cudaStream_t st);

}  // namespace mdns
