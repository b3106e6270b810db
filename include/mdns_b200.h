/*
 * mdns_b200.h -- C ABI of libmdns_b200.so, the B200-native (sm_100a) hot path of
 * massivedatans.  Plain C types only (pointers, sizes, doubles); loaded with
 * ctypes exactly like the reference's clike.so / cmuselike.so / cneighbors.so.
 *
 * Each entry point names the reference interface it replaces; paths are
 * relative to the upstream JohannesBuchner/massivedatans tree.
 *
 * Conventions
 *   - every function returning int returns 0 (MDNS_OK) on success and a
 *     negative MDNS_E* code on failure; mdns_last_error() then holds a message
 *     (thread-local).  The reference C always returns 0 and its callers ignore
 *     the value (sample.py:106, neighbors.py:141); here failures are loud.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with MDNS_ECUDA.
 *   - host matrices use the reference layout: channel-major, data-set index
 *     fastest, element (channel j, data set i) at m[i + j*ndata]
 *     (clike.c:72, cmuselike.c:53).  Masks are 1-byte C bool / numpy.bool_
 *     arrays (sample.py:94); any non-zero byte means "active".
 *   - all arithmetic is IEEE FP64.
 */
#ifndef MDNS_B200_H
#define MDNS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDNS_OK        0
#define MDNS_EINVAL   -1   /* bad argument                                  */
#define MDNS_ECUDA    -2   /* CUDA runtime / launch failure, or no device   */
#define MDNS_ENOMEM   -3   /* host or device allocation failed              */
#define MDNS_ESTATE   -4   /* call sequence error (e.g. launch before stage)*/

/* ---- process-wide ---------------------------------------------------- */
const char *mdns_last_error(void);
int         mdns_version(void);          /* 100*major + minor */
int         mdns_device_count(void);     /* CUDA devices visible; 0 if none */
/* Kernels launched by this library since load (bench.py's "gpu_launches"). */
int64_t     mdns_launch_count(void);
/* Name of the kernel launched most recently by this library (measurement aid). */
const char *mdns_last_kernel(void);
/* Host helper of the neighbour path: the smallest double T such that
 * sqrt(T) >= r in IEEE arithmetic, so that  sqrt(d) < r  <=>  d < T  exactly
 * (the compare at cneighbors.c:88,109 without a sqrt in the device loop). */
double      mdns_sqrt_threshold(double r);
/* Pinned host memory for result vectors (D2H lands without a bounce copy). */
void       *mdns_host_alloc(int64_t bytes);
int         mdns_host_free(void *p);

/* ---- resident data sets ---------------------------------------------- */
typedef struct mdns_dataset mdns_dataset;

/*
 * Upload the data once and keep it resident in HBM (replaces passing x, yy
 * [and vv] host pointers on every call: sample.py:106, musefuse.py:534).
 *   x   [nx]          wavelength grid, may be NULL when only *_spectra calls are used
 *   yy  [nx*ndata]    data, reference layout
 *   vv  [nx*ndata]    per-element variance (cmuselike.c:36) or NULL for the
 *                     scalar-noise likelihood of clike.c
 *   devices/ndevices  CUDA ordinals to shard the data sets over, contiguous
 *                     ranges of the data-set index; NULL/0 = current device 0.
 * The host arrays are copied; they may be freed after the call returns.
 */
int mdns_dataset_create(const double *x, const double *yy, const double *vv,
                        int ndata, int nx, const int *devices, int ndevices,
                        mdns_dataset **out);
/*
 * The same from files, without a host copy of the matrix (the loader half of sample.py:27-31,
 * which reads the whole `y` into RAM: 8 GB at BASELINE configs[3]).  y_path / v_path: .npy files
 * (numpy.save) holding the reference-layout matrix [nx][ndata], float64, C order; v_path may be
 * NULL.  Column blocks are read with pread into a 64 MB pinned buffer, copied to the device and
 * transposed there into the resident rows; every shard reads only its own columns.
 * (The reference stores HDF5; h5py's `numpy.save(path, f['y'][()])` converts once.)
 */
int mdns_dataset_create_from_npy(const double *x, const char *y_path, const char *v_path,
                                 const int *devices, int ndevices, mdns_dataset **out);
int mdns_dataset_destroy(mdns_dataset *ds);
int mdns_dataset_info(const mdns_dataset *ds, int *ndata, int *nx, int *nshards,
                      int64_t *resident_bytes);

/*
 * One-call evaluation with HOST inputs and outputs (the drop-in path).
 *
 * mdns_clike_eval_params: K parameter points params[k] = (A, mu, sig); the
 * line model A*exp(-0.5*((mu-x_j)/sig)^2) (clike.c:65) is generated on the
 * device, then for every active data set i (rank r among the active ones,
 * clike.c:67-74):
 *     Lout[k*n_act + r] = scale * sum_j ((ypred_kj - y_ij)/noise)^2
 * scale = 1 gives what clike.c leaves in Lout, scale = -0.5 folds in
 * sample.py:108.  mask == NULL means all data sets.  Lout must hold
 * lout_capacity >= K*n_act doubles; *n_act_out (may be NULL) receives n_act.
 * Replaces `like` of clike.c:34-40 called K times.
 */
int mdns_clike_eval_params(mdns_dataset *ds, const double *params, int K,
                           double noise, double scale, const uint8_t *mask,
                           double *Lout, int64_t lout_capacity, int *n_act_out);
/* Same with K caller-provided model spectra ypred[K][nx] (the model() of
 * musefuse.py:222-284 stays on the host). */
int mdns_clike_eval_spectra(mdns_dataset *ds, const double *ypred, int K,
                            double noise, double scale, const uint8_t *mask,
                            double *Lout, int64_t lout_capacity, int *n_act_out);
/*
 * MUSE scaled chi-square (cmuselike.c:48-64): for every active data set i
 *     Lout[k*ndata + i] = -0.5 * sum_j (y_ij - s*m_kj)^2 / v_ij ,
 *     s = (sum_j y m / v) / (1e-10 + sum_j m^2 / v)
 * un-compacted; entries of inactive data sets are left untouched
 * (cmuselike.c:49,62).  Needs a data set created with vv.
 * Replaces `like` of cmuselike.c:34-38 called K times.
 */
int mdns_muse_eval_spectra(mdns_dataset *ds, const double *ypred, int K,
                           const uint8_t *mask, double *Lout);

/*
 * MUSE stellar-population model on the device (musefuse.py:222-284 `model`), so that a batch of
 * K parameter points becomes K staged model spectra without a spectrum crossing PCIe.
 *   grids[nZ][nages][nwave]  template grids, one [nages][nwave] array per metallicity
 *                            (musefuse.py:176-185 `grid`), non-negative
 *   Zs[nZ]                   metallicity nodes, increasing (musefuse.py:188)
 *   ages[nages]              template ages in yr, increasing (musefuse.py:191)
 *   model_wavelength[nwave]  template wavelengths, strictly increasing, same unit as
 *                            `wavelength` (musefuse.py:206-207)
 *   calzetti[nwave]          attenuation curve on the template grid (musefuse.py:208-217)
 *   wavelength[nx]           data wavelength grid; nx = channels of the data set
 *   norm_index               template channel the spectrum is normalised at (2050, :255)
 * Everything is copied to every device of `ds`.
 */
typedef struct mdns_muse_model mdns_muse_model;
int mdns_muse_model_create(mdns_dataset *ds, const double *grids, int nZ, const double *Zs,
                           int nages, const double *ages, int nwave,
                           const double *model_wavelength, const double *calzetti,
                           const double *wavelength, int nx, int norm_index,
                           mdns_muse_model **out);
int mdns_muse_model_destroy(mdns_muse_model *m);
/* params[K][5] = (Z, SFtau, sfage, z, EBV), the arguments of model() (SFtau in yr, i.e.
 * 10**logSFtau, musefuse.py:524).  Builds the K spectra and stages them like
 * mdns_stage_spectra; follow with mdns_set_mask + mdns_muse_launch + mdns_fetch.
 * nonzero[K] (may be NULL): numpy.any(ypred) per spectrum, the guard of musefuse.py:528-530.
 * MDNS_EINVAL if a metallicity lies below Zs[0] (the reference raises IndexError there). */
int mdns_muse_model_stage(mdns_muse_model *m, const double *params, int K, int *nonzero);
/* The K spectra of the last mdns_muse_model_stage, ypred_out[K][nx] (parity checks). */
int mdns_muse_model_spectra(mdns_muse_model *m, double *ypred_out);

/*
 * Staged interface (inputs resident in HBM; used for device-side timing and
 * by callers that keep the mask across many candidates, e.g. one
 * draw_constrained call, hiermetriclearn.py:173-211).
 *   set_mask     upload + compact the mask (NULL = all); returns n_act
 *   stage_params upload K (A,mu,sig) triples          (clike-type)
 *   stage_spectra upload K model spectra [K][nx]      (either type)
 *   launch       enqueue the kernels only (asynchronous, no copies)
 *   fetch        D2H of the last launch's result + synchronise; layout as in
 *                the one-call functions (compacted for clike, full for muse)
 *   sync         wait for all enqueued work
 */
int mdns_set_mask(mdns_dataset *ds, const uint8_t *mask, int *n_act_out);
int mdns_stage_params(mdns_dataset *ds, const double *params, int K);
int mdns_stage_spectra(mdns_dataset *ds, const double *ypred, int K);
int mdns_clike_launch(mdns_dataset *ds, double noise, double scale);
int mdns_muse_launch(mdns_dataset *ds);
/* launch + fetch of the clike-type result in one call: the active rows are processed in
 * chunks and the D2H copy of each finished chunk overlaps the next chunk's kernel. */
int mdns_clike_launch_fetch(mdns_dataset *ds, double noise, double scale, double *Lout,
                            int64_t lout_capacity);
/* Speculative batch of the constrained draw (hiermetriclearn.py:181-196: one candidate at a
 * time until `numpy.any(L > Lmins)`): scores the K staged candidates in one pass and applies
 * the accept test on the device.  Lmins[n_act] is aligned with the compacted active order.
 * accept_counts[K] (may be NULL) = data sets with scale*chi2 > Lmins per candidate;
 * *first_k = first candidate with a non-zero count, or -1; its logL vector is copied to Lout
 * (contents unspecified when none accepts).  Only K ints and one vector cross PCIe -- and no
 * vector at all when no candidate is accepted.  The accept test runs inside the likelihood
 * kernels and the choice of first_k on the device; the accepted candidate's vector is then
 * fetched from its row of the device-side logL matrix. */
/* Thresholds may also be staged once per constrained draw (they are constant while candidates
 * are tried, hiermetriclearn.py:173-211): mdns_set_thresholds after mdns_set_mask, then
 * mdns_clike_first_accept with Lmins = NULL for every batch of candidates. */
int mdns_set_thresholds(mdns_dataset *ds, const double *Lmins);
int mdns_clike_first_accept(mdns_dataset *ds, double noise, double scale, const double *Lmins,
                            int *accept_counts, int *first_k, double *Lout,
                            int64_t lout_capacity);
/* One call per speculative pass of a constrained draw: mdns_set_mask (a repeated mask is
 * recognised), mdns_set_thresholds (skipped when a short Lmins equals the staged one),
 * mdns_stage_params and mdns_clike_first_accept.  *n_act_out = active data sets; Lmins and Lout
 * must both hold lout_capacity >= n_act doubles (checked before either is touched).  With at most
 * 64 active data sets and at most 32 candidates the pass is ONE kernel launch (candidates by
 * value, decision and vector written to pinned memory; the logL are bit-identical to the general
 * path's); after such a pass the candidates are not staged on the device: stage again before
 * mdns_clike_launch / mdns_clike_first_accept. */
int mdns_clike_draw_pass(mdns_dataset *ds, const uint8_t *mask, const double *Lmins,
                         const double *params, int K, double noise, double scale,
                         int *accept_counts, int *first_k, double *Lout, int64_t lout_capacity,
                         int *n_act_out);
/* The same decision, returning only what the sampler consumes (multi_nested_sampler.py:482-485
 * pushes the new point onto the shelves of the data sets with Lj[j] > Lmins[j]): idx_out[0..n)
 * = positions j (in the compacted active order, increasing) of the data sets the first accepted
 * candidate is accepted for, val_out = their logL.  12 bytes per accepting data set cross PCIe
 * instead of the whole vector.  capacity = entries idx_out / val_out can hold (n_act is always
 * enough). */
int mdns_clike_first_accept_sparse(mdns_dataset *ds, double noise, double scale, const double *Lmins,
                                   int *accept_counts, int *first_k, int32_t *idx_out,
                                   double *val_out, int64_t capacity, int *n_out);
/* Two-step form: accept_counts[K] alone (of this process's data sets; summed over the ranks when
 * a communicator is attached, see mdns_comm_init), then -- if wanted -- the logL vector of any
 * candidate k of the same launch.  For callers whose consumer lives on the device (the live-point
 * table) or that run the exchange between processes themselves. */
int mdns_clike_accept_counts(mdns_dataset *ds, double noise, double scale, const double *Lmins,
                             int *accept_counts);
int mdns_fetch_candidate(mdns_dataset *ds, int k, double *Lout, int64_t lout_capacity);
int mdns_fetch(mdns_dataset *ds, double *Lout, int64_t lout_capacity);
int mdns_sync(mdns_dataset *ds);
/* Experiment knob: the dense first-accept pass can score the active data sets in row chunks; the
 * accept counts so far go down after every chunk and the rows of a finished chunk start their
 * way to the host as soon as the counts name a candidate.  0 = automatic (one launch: measured
 * faster), 1 = one chunk, up to 7. */
int mdns_set_draw_chunks(mdns_dataset *ds, int nchunks);

/* ---- one process per GPU: the exchange step ---------- */
/*
 * Data sets are independent (clike.c:68-74), so ranks never exchange data.  The one global
 * decision of the path is `numpy.any(L > Lmins)` (hiermetriclearn.py:193): with a communicator
 * attached, mdns_clike_first_accept[_sparse] / mdns_clike_draw_pass add the K per-candidate accept
 * counts up over the ranks with ncclAllReduce on the data set's stream (NVLink / NVSwitch) before
 * the first accepted candidate is chosen on the device -- every rank returns the same first_k and
 * the GLOBAL counts, and its own slice of the logL vector.  Every rank must make the same calls
 * (also ranks whose mask leaves them no active data set).  NCCL is bound with dlopen when the
 * first of these functions is called; nothing else in the library needs it.
 *   unique_id  MDNS_UNIQUE_ID_BYTES bytes from ncclGetUniqueId, created by one rank and handed to
 *              the others by the caller (massivedatans_b200.sharding does it over TCP)
 *   init       collective (ncclCommInitRank); the data set must live on ONE device
 */
#define MDNS_UNIQUE_ID_BYTES 128
int mdns_comm_unique_id(void *id_out);
int mdns_comm_init(mdns_dataset *ds, const void *id, int nranks, int rank);
int mdns_comm_destroy(mdns_dataset *ds);
int mdns_comm_info(const mdns_dataset *ds, int *nranks, int *rank);
/* values[n] (host, n <= 1024) := sum (op 0) or max (op 1) over the ranks; doubles as a barrier.
 * No communicator: values unchanged. */
int mdns_comm_allreduce(mdns_dataset *ds, double *values, int n, int op);
/* Device-consumer exchange: every rank contributes the logL vector of candidate k of its last
 * clike launch (n_act entries); afterwards every rank holds all of them in rank order on its
 * device and, if Lall != NULL, compacted in Lall (n_per_rank[r] entries of rank r; may be NULL).
 * ncclAllGather on the data set's stream. */
int mdns_comm_allgather_candidate(mdns_dataset *ds, int k, double *Lall, int64_t capacity,
                                  int *n_per_rank);
/* CUDA-event stopwatch on the data set's own streams (max over shards). */
int mdns_timer_start(mdns_dataset *ds);
int mdns_timer_stop(mdns_dataset *ds, float *elapsed_ms);
/* Measurement aid: evict the L2 by overwriting a 256 MB scratch buffer on every shard's stream
 * (between timed iterations of problems that fit in the 126 MB L2). */
int mdns_flush_l2(mdns_dataset *ds);
/* Kernel-variant override for experiments: lanes per data set (0 = auto),
 * fragments in flight per lane (0 = auto), candidates per pass (0 = auto),
 * data sets per lane group (0 = auto; > 1 selects the register-blocked kernel).
 * lanes = 7: parameter-point batches of any size in ONE launch (every CTA builds the spectra,
 * direct form) -- the automatic choice only from 5 candidates and up to 1e5 model x data-set
 * evaluations. */
int mdns_set_tuning(mdns_dataset *ds, int lanes, int unroll, int ktile, int rows);
/* Expanded form of the candidate-batch kernels (automatic for K >= 3 over >= 8192 data sets when
 * all are active, for K >= 5 over >= 4096 active data sets of a masked batch; batches of up to 1e5
 * evaluations take the one-launch direct-form kernel instead):
 *     sum_j (m_j - y_j)^2 = Syy - 2*Sym + Smm ,  Syy resident per data set,
 * one FP64 FMA per (element, candidate) instead of two operations.  FP64 throughout; a
 * result is kept only when the rounding-error bound of the three sums is below rel_tol
 * (default 1e-10; the parity contract of clike.c:64-76 is 1e-9), anything else is recomputed
 * in the direct form inside the same launch.  enable = 0 keeps every launch on the direct
 * kernels; rel_tol <= 0 leaves the tolerance unchanged.  Data that needs more than 2 % of its
 * rows recomputed switches itself back to the direct form. */
int mdns_set_expanded(mdns_dataset *ds, int enable, double rel_tol);
int mdns_expanded_stats(const mdns_dataset *ds, int *enabled, int64_t *redo_rows);

/* ---- live-point likelihood table (next row of the hot path) ------------ */
/*
 * `live_pointsL[nlive_points][ndata]` of multi_nested_sampler.py:111 resident
 * next to the data it was scored on (same devices, same data-set ranges), so
 * that the per-iteration reductions of the sampler run where the numbers are.
 * Row-major like the numpy array (row = live point).  All results are
 * selections and therefore bit-identical to numpy's.
 */
typedef struct mdns_livetable mdns_livetable;
int mdns_livetable_create(mdns_dataset *ds, int nlive, mdns_livetable **out);
int mdns_livetable_destroy(mdns_livetable *t);
int mdns_livetable_upload(mdns_livetable *t, const double *L);      /* [nlive][ndata] */
int mdns_livetable_download(mdns_livetable *t, double *L);
/* rows [row0, row0+K) := the K logL vectors of the data set's last clike launch
 * (every data set active): the initial population of
 * multi_nested_sampler.py:91-111 without leaving the device. */
int mdns_livetable_fill_from_launch(mdns_livetable *t, mdns_dataset *ds, int row0);
/* prepare(), multi_nested_sampler.py:134-137 and :531: per data set the minimum,
 * the row of its first occurrence (numpy.argmin) and the maximum over the live
 * points.  Any output may be NULL. */
int mdns_livetable_colstats(mdns_livetable *t, double *Lmins, int64_t *Lmini, double *Lmax);
/* The column minima become the accept thresholds of `ds` (every data set active) on the
 * devices, without a host round trip: table -> thresholds -> mdns_clike_first_accept with
 * Lmins = NULL. */
int mdns_livetable_stage_thresholds(mdns_livetable *t, mdns_dataset *ds);
/* advance, multi_nested_sampler.py:520-524: table[rows[d]][d] = values[d] for
 * every data set d with rows[d] >= 0. */
int mdns_livetable_replace(mdns_livetable *t, const int64_t *rows, const double *values);
/* Lmins_higher, multi_nested_sampler.py:438-447 (find_nsmallest :44-47): for the
 * j-th listed data set d = indices[j] (increasing) with n = shelf_offsets[j+1] -
 * shelf_offsets[j] queued likelihoods shelf_values[shelf_offsets[j] ...],
 * out[j] = numpy.partition(concatenate(table[:, d], shelf), n)[n]. */
int mdns_livetable_lmins_higher(mdns_livetable *t, const int *indices, int nidx,
                                const int64_t *shelf_offsets, const double *shelf_values,
                                double *out);

/*
 * Subset partition, generate_subsets_graph / generate_subsets_nograph
 * (multi_nested_sampler.py:204-355): groups of data sets that share live points,
 * i.e. connected components of the data set <-> live point graph.
 *   upload_points   live_pointsp[nlive][ndata] (int64 indices into the point pile,
 *                   multi_nested_sampler.py:111); kept as int32 on the first device
 *   replace_points  live_pointsp[rows[d]][d] = ids[d] for rows[d] >= 0 (:523)
 *   subsets         labels[d] = smallest data-set index of d's group (-1 outside the
 *                   mask; NULL mask = all); groups in increasing label order are the
 *                   order the reference yields them in (:239).  npoints = size of the
 *                   point pile (every id < npoints).
 */
int mdns_livetable_upload_points(mdns_livetable *t, const int64_t *live_pointsp);
int mdns_livetable_replace_points(mdns_livetable *t, const int64_t *rows, const int64_t *ids);
int mdns_livetable_subsets(mdns_livetable *t, const uint8_t *data_mask, int64_t npoints,
                           int32_t *labels, int *ncomponents, int *nrounds);

/* ---- RadFriends neighbour tests --------------------------------------- */
typedef struct mdns_region mdns_region;

/* A region owns the resident member set (live-point union, metric space). */
int mdns_region_create(int device, mdns_region **out);
int mdns_region_destroy(mdns_region *rg);
/* CUDA-event stopwatch on the region's stream (measurement aid). */
int mdns_region_timer_start(mdns_region *rg);
int mdns_region_timer_stop(mdns_region *rg, float *elapsed_ms);
/* xx[n][ndim] row-major (neighbors.py:100 argtype); upload once per region
 * (radfriendsregion.py:59-70 builds one region from `members`). */
int mdns_region_set_members(mdns_region *rg, const double *xx, int n, int ndim);
/* cneighbors.c:95-119: out[j] (caller-initialised, normally zeros) gains one
 * per member within maxdistance of candidate yy[j]; with countmax > 0 the
 * scan of candidate j stops once out[j] >= countmax.  Bit-exact. */
int mdns_region_count_within(mdns_region *rg, double maxdistance,
                             const double *yy, int m, double *out, int countmax);
/* cneighbors.c:77-92: *result = 1 if any member is within maxdistance of y. */
/* Candidate generation on the device (the ball draws of RadFriendsRegion.generate,
 * radfriendsregion.py:156-178, fused with the neighbour count): proposals
 * first_proposal .. first_proposal+nproposals-1 of the stream keyed by `seed` (counter-based
 * Philox, not the numpy stream: statistical parity only); the accepted points (uniform in the
 * union of balls of radius maxdistance around the members) land in points_out[n_out][ndim], in
 * proposal order.  capacity = points points_out can hold (nproposals is always enough). */
int mdns_region_generate(mdns_region *rg, double maxdistance, uint64_t seed, uint64_t first_proposal,
                         int nproposals, double *points_out, int64_t capacity, int *n_out);
int mdns_region_is_within(mdns_region *rg, double maxdistance, const double *y,
                          int *result);
/* Per-axis "SupFriends" distance, clustering/neighbors.py:22-73 (find_maxdistance).
 * neighbors.py:24-25: nearest_out[i] = index of the member nearest to member i (euclidean,
 * i itself excluded).  Needs >= 2 members. */
int mdns_region_nearest_index(mdns_region *rg, int *nearest_out);
/* neighbors.py:40-43: covered_out[q] = 1 if some member ref[r] lies inside the per-axis box
 * maxdistance[ndim] around member query[q] (|x_qk - x_rk| < maxdistance[k] for every axis k).
 * The sequential box growth of neighbors.py:44-58 stays with the caller, which only has to
 * visit the members reported uncovered. */
int mdns_region_axis_covered(mdns_region *rg, const double *maxdistance, const int *query, int nq,
                             const int *ref, int nr, uint8_t *covered_out);
/* cneighbors.c:125-179: chosen[n][nboot] float64 0/1 (round index fastest). */
int mdns_region_bootstrapped_maxdistance(mdns_region *rg, const double *chosen,
                                         int nboot, double *result);
/* cneighbors.c:32-75. */
int mdns_region_most_distant_nearest_neighbor(mdns_region *rg, double *result);

/*
 * One-shot forms with the reference's exact C signatures (cneighbors.c:32-34,
 * 77-79, 95-98, 125-130).  They run on a process-wide region on device 0 and
 * skip the member upload when xx is byte-identical to the previous call's.
 * On failure they print mdns_last_error() to stderr and return NaN / -1.
 */
double mdns_most_distant_nearest_neighbor(const void *xx, int nsamples, int ndim);
int    mdns_is_within_distance_of(const void *xx, int nsamples, int ndim,
                                  double maxdistance, const void *y);
int    mdns_count_within_distance_of(const void *xx, int nsamples, int ndim,
                                     double maxdistance, const void *yy,
                                     int nothers, void *out, int countmax);
double mdns_bootstrapped_maxdistance(const void *xx, int nsamples, int ndim,
                                     const void *chosen, int nbootstraps);

/*
 * One-shot likelihoods with the reference's exact C signatures (clike.c:34-40,
 * cmuselike.c:34-38).  The data matrix is made resident on first sight of
 * (pointer, shape) and re-used afterwards.  The reference re-reads its arguments
 * on every call (clike.c:72), so a cached matrix is re-validated on every call by
 * hashing ALL of it (several host threads; ~25 ms per GB): any in-place edit is
 * seen and the matrix is uploaded again.  At most two matrices stay resident
 * (least recently used is dropped).  Callers that never modify a matrix in place
 * can opt out of the full hash -- mdns_legacy_trust(1) or MDNS_LEGACY_TRUST=1:
 * 256 probes only -- or, better, use the resident interface above.
 * mdns_clike_like accumulates into Lout as clike.c:72 does.
 */
int mdns_clike_like(const void *x, const void *yy, int ndata, int nx, double A,
                    double mu, double sig, double noise_level,
                    const void *data_mask, void *Lout);
int mdns_cmuselike_like(const void *yy, const void *vv, const void *ypred,
                        const void *data_mask, int ndata, int nx, void *Lout);
/* Drop every data set cached by the two functions above. */
int mdns_legacy_reset(void);
/* trust = 1: cached matrices are re-validated with 256 probes instead of a full hash (the caller
 * vouches that they are not edited in place); 0: full hash (default); anything else: query.
 * Returns the previous setting. */
int mdns_legacy_trust(int trust);

#ifdef __cplusplus
}
#endif
#endif /* MDNS_B200_H */
