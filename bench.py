#!/usr/bin/env python
"""bench.py -- headline measurement of the massivedatans hot path on B200.

Metric (BASELINE.json): model x data-set logL evaluations per second at N = 1e4..1e6 on
1/2/4/8 GPUs, and % of the HBM roofline of the batched likelihood kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU)

One "step" = one pass of the batched likelihood over one batch: `--candidates` parameter
points scored against every data set of the GPU's resident shard (gensimple_horns-style
synthetic spectra, C=200 channels, N=1e6 data sets per GPU -> 1.6 GB, far beyond the 126 MB
L2, so every step streams from HBM).  Data sets are sharded contiguously, one shard per
rank, no data-path collective: weak scaling.  One process per GPU; the ranks meet through the
shim's own NCCL communicator (mdns_comm_*, no framework in the loop): barrier + max over ranks
for the timings, and the all-reduce of the K accept counts inside the first-accept passes.

Keys of the JSON line (rank 0 prints exactly one):
  value      evals/s with inputs (data, mask, parameter points) resident in HBM, device-timed
             with CUDA events on the shim's stream, max over ranks
  roofline   algorithmic bytes per step / device time per step vs MEASURED_PEAKS.json hbm_gbs;
             `sustained_frac`: the same over a >= 2 s back-to-back loop (power-capped clocks);
             traffic = dram bytes of the dominant kernel from the committed ncu capture
  e2e        the same metric through the public Python callable with HOST buffers: per step the
             mask and parameter points go host->device and the K x N logL matrix comes back
             (PCIe-bound).  e2e.first_accept: the same batch through the device-side form of the
             constrained draw's loop (hiermetriclearn.py:181-196): K parameter points in, the
             GLOBAL accept counts (ncclAllReduce over the ranks inside the shim) + this rank's
             slice of the accepted candidate's logL vector out; .sparse: only the accepting data
             sets' indices and logL; .device_consumer: counts only, the vector stays in HBM for
             the resident live-point table (mdns_livetable_*)
  configs3   BASELINE configs[3] as written: gen_realistic-style spectra, 1e6 data sets x 1000
             channels IN TOTAL, strong-scaled (1e6 / N per GPU), device-timed + roofline
  sweep_n    (N = 1) the metric at N = 1e4, 1e5, 1e6 data sets for K = 1 and 16, L2 flushed
             between timed launches where the data fits the L2
  muse       (N = 1) cmuselike path on a cube of the reference shape, roofline + reference C
  neighbors  (N = 1) RadFriends pair tests/s at the shapes the sampler issues + reference C
  cpu_baseline  the reference's own unmodified clike.so (oracle/_ref; the serial build
             sample.py:81-84 loads, its OpenMP variant is racy, clike.c:32), one contiguous
             data-set shard per host thread, all host cores, bounded sample
The reference arm (--impl reference) times that same CPU implementation per step at the stated N.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'model x data-set logL evaluations per second'
UNIT = 'evals/s'
HBM_FALLBACK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--ndata', type=int, default=1000000, help='data sets per GPU')
    ap.add_argument('--nx', type=int, default=200, help='channels')
    ap.add_argument('--candidates', type=int, default=16, help='parameter points per step')
    ap.add_argument('--mask', default='all', choices=['all', 'half', 'sparse', 'prefix'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--tuning', default='', help='lanes,unroll,ktile override (experiments)')
    ap.add_argument('--sweep', action='store_true',
                    help='also time K = 1 ... 400 device-side and add them under "sweep"')
    ap.add_argument('--no-expanded', action='store_true',
                    help='keep candidate batches on the direct-form kernels')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip configs3 / sweep_n / muse / neighbors (kernel experiments)')
    ap.add_argument('--configs3-ndata', type=int, default=1000000,
                    help='TOTAL data sets of the configs[3] strong-scaling line')
    ap.add_argument('--sustained-s', type=float, default=2.0)
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return rank, world, local


class ClockSampler(object):
    """Samples SM clock and throttle reasons of one GPU through NVML while work runs."""

    def __init__(self, index, period=0.01):
        self.samples = []
        self.reasons = set()
        self.period = period
        self.stop_flag = threading.Event()
        self.thread = None
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:               # NVML unavailable: report nulls
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = (('hw_slowdown', 0x8), ('sw_power_cap', 0x4), ('sw_thermal_slowdown', 0x20),
                     ('hw_thermal_slowdown', 0x40), ('hw_power_brake_slowdown', 0x80),
                     ('applications_clocks_setting', 0x2), ('sync_boost', 0x10))
            for name, bit in names:
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self.stop_flag.is_set():
            self._once()
            time.sleep(self.period)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nv is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        self._once()
        self.stop_flag.set()
        self.thread.join()
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.sm_max,
                'reasons': sorted(self.reasons), 'samples': len(s)}


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return HBM_FALLBACK_GBS, 'fallback (B200_PROFILING.md)'


def ncu_traffic(nx, ndata, candidates):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if the
    capture was taken on this very configuration (profiles/clike_traffic.json)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'clike_traffic.json')) as f:
            t = json.load(f)
        if (t.get('nx'), t.get('ndata'), t.get('candidates')) == (nx, ndata, candidates):
            return float(t['dram_bytes_per_launch'])
    except Exception:
        pass
    return None


def algorithmic_bytes(n_act, ndata, nx, K):
    # data rows once + K model spectra + K logL vectors + mask
    return n_act * nx * 8 + K * nx * 8 + K * n_act * 8 + ndata


def muse_bytes(n_act, ndata, nx, K):
    # y and 1/v rows once each + K spectra + K logL vectors + mask
    return n_act * nx * 16 + K * nx * 8 + K * n_act * 8 + ndata


def make_inputs(args, rank):
    from massivedatans_b200 import synth
    x, y, _ = synth.horns(args.ndata, nx=args.nx, legacy=False, seed=1000 + rank)
    pts = synth.parameter_points(args.candidates, seed=7)
    log_pts = pts.copy()
    log_pts[:, 2] = numpy.log10(pts[:, 2])       # what the reference callable receives
    mask = synth.masks(args.ndata, seed=11)[args.mask]
    return x, y, pts, log_pts, mask


def host_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


class ReferencePool(object):
    """The reference's own clike.so (oracle/_ref, the serial build sample.py:81-84 loads --
    its OpenMP variant is racy, clike.c:32) driven with every host thread it can use: the
    sample is cut into contiguous data-set shards (copied once, outside the timed region) and
    each host thread scores all candidates against its own shard.  ctypes releases the GIL
    for the duration of each call, so the shards run concurrently; the C code is unmodified."""

    def __init__(self, x, y, mask, threads):
        from concurrent.futures import ThreadPoolExecutor
        nx, ndata = y.shape
        self.threads = max(1, min(threads, ndata))
        bounds = numpy.linspace(0, ndata, self.threads + 1).astype(int)
        self.x = x
        self.shards = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            m = numpy.ascontiguousarray(mask[a:b])
            self.shards.append((numpy.ascontiguousarray(y[:, a:b]), m, numpy.zeros(int(m.sum()))))
        self.n_act = int(mask.sum())
        self.pool = ThreadPoolExecutor(self.threads)

    def _work(self, shard, pts):
        from oracle import ref
        y, m, out = shard
        for p in pts:
            out[:] = 0
            ref.clike(self.x, y, p[0], p[1], p[2], 0.01, m, Lout=out)
        return out[0] if out.size else 0.0

    def step(self, pts):
        list(self.pool.map(lambda sh: self._work(sh, pts), self.shards))
        return len(pts) * self.n_act

    def close(self):
        self.pool.shutdown()


def cpu_reference_rate(x, y, pts, mask, budget_s, threads, min_reps=1):
    """Time the reference's clike.so on candidates x data sets; returns evals/s."""
    pool = ReferencePool(x, y, mask, threads)
    pool.step(pts[:1])                      # touch the pages, load the library
    done = 0
    reps = 0
    t0 = time.perf_counter()
    while True:
        done += pool.step(pts)
        reps += 1
        if reps >= min_reps and time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    pool.close()
    return done / dt, dt, reps


REF_NOTE = ('reference clike.so (gcc -O3; the serial build sample.py:81-84 loads, its OpenMP '
            'variant is racy, clike.c:32), unmodified, one contiguous data-set shard per host '
            'thread')


def run_reference(args):
    """--impl reference: the reference's own CPU implementation at the stated N, rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from massivedatans_b200 import synth
    n = args.ndata
    x, y, _ = synth.horns(n, nx=args.nx, legacy=False, seed=1000)
    pts = synth.parameter_points(args.candidates, seed=7)
    mask = synth.masks(n, seed=11)[args.mask]
    n_act = int(mask.sum())
    threads = host_threads()
    pool = ReferencePool(x, y, mask, threads)
    # bounded: every step is the full workload (K candidates x N data sets, 0.2-0.5 s on a
    # 16-64 core host); cap the number of steps so that the arm ends within a few minutes
    t0 = time.perf_counter()
    pool.step(pts)
    one = time.perf_counter() - t0
    steps = max(1, min(args.steps, int(120.0 / max(one, 1e-3))))
    for _ in range(max(0, min(args.warmup, 3) - 1)):
        pool.step(pts)
    t0 = time.perf_counter()
    for _ in range(steps):
        pool.step(pts)
    dt = time.perf_counter() - t0
    pool.close()
    value = steps * args.candidates * n_act / dt
    sample = ('%d data sets x %d channels x %d candidates per step (the stated workload of ONE GPU), '
              '%d timed steps on %d host threads, %s'
              % (n, args.nx, args.candidates, steps, pool.threads, REF_NOTE))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': pool.threads, 'kind': 'reference',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {
        'workload': ('gensimple_horns-style spectra: %d data sets per GPU x %d channels, mask=%s, '
                     '%d candidate parameter points per step (BASELINE configs[2] shape at the '
                     "metric's N = 1e6; configs[3] as written is the `configs3` sub-object)"
                     % (args.ndata, args.nx, args.mask, args.candidates)),
        'ndata_per_gpu': args.ndata, 'nx': args.nx, 'candidates_per_step': args.candidates,
        'mask': args.mask,
        'l2': 'inputs larger than L2 (%.2f GB resident per GPU vs 126 MB L2), no flush needed'
              % (args.ndata * args.nx * 8 / 1e9),
        'sharding': 'contiguous data-set ranges, one shard per GPU, no data-path collective',
    }


def device_time(ds, steps, warm=3, flush=False):
    """ms per launch of the staged batch, CUDA events on the shim's stream.  flush: evict the L2
    before every timed launch and time each launch on its own (small problems)."""
    for _ in range(warm):
        ds.launch_clike(0.01, -0.5)
    ds.sync()
    if not flush:
        ds.timer_start()
        for _ in range(steps):
            ds.launch_clike(0.01, -0.5)
        return ds.timer_stop() / steps
    total = 0.0
    for _ in range(steps):
        ds.flush_l2()
        ds.timer_start()
        ds.launch_clike(0.01, -0.5)
        total += ds.timer_stop()
    return total / steps


def sustained_time(ds, seconds, first_guess_ms):
    """ms per launch over a back-to-back loop of at least `seconds` (the board settles at its
    power cap), CUDA events around the whole loop."""
    n = max(50, int(seconds * 1e3 / max(first_guess_ms, 1e-3)))
    ds.timer_start()
    for _ in range(n):
        ds.launch_clike(0.01, -0.5)
    return ds.timer_stop() / n, n


def run_ours(args):
    rank, world, local = dist_env()
    distributed = world > 1
    inproc_devices = None
    if distributed:
        n_gpus = world
        device = local
    else:
        n_gpus = args.gpus
        device = 0
        if n_gpus > 1:        # not under torchrun: one process drives N shards
            inproc_devices = list(range(n_gpus))

    from massivedatans_b200 import _lib, sharding, synth
    from massivedatans_b200.likelihood import ResidentDataset, make_multi_loglikelihood
    lib = _lib.load()
    _lib.require_device()
    # pinned result buffers next to this rank's GPU (NUMA): D2H of all ranks in parallel
    bound = sharding.bind_near_device(device) if distributed else False

    # ---- inputs ------------------------------------------------------------
    if inproc_devices:
        import copy
        big = copy.copy(args)
        big.ndata = args.ndata * n_gpus
        x, y, pts, log_pts, mask = make_inputs(big, 0)
        f = make_multi_loglikelihood(x, y, 0.01, devices=inproc_devices)
    else:
        x, y, pts, log_pts, mask = make_inputs(args, rank)
        f = make_multi_loglikelihood(x, y, 0.01, devices=[device])
    ds = f.dataset
    if distributed:
        # (NCCL prints its version banner to stdout at the first init when NCCL_DEBUG asks for it:
        # rank 0's stdout carries the one JSON line and nothing else)
        os.environ.pop('NCCL_DEBUG', None)
        if os.environ.get('MDNS_NCCL_DEBUG'):
            os.environ['NCCL_DEBUG'] = os.environ['MDNS_NCCL_DEBUG']
        sharding.init_comm_from_env(ds)

    def barrier():
        ds.sync()
        if distributed:
            ds.comm_allreduce([0.0])

    def max_over_ranks(v):
        return float(ds.comm_allreduce([v], op='max')[0]) if distributed else v

    def sum_over_ranks(v):
        return numpy.asarray(ds.comm_allreduce(numpy.asarray(v, dtype=float))) if distributed \
            else numpy.asarray(v, dtype=float)

    if args.tuning:
        ds.set_tuning(*[int(v) for v in args.tuning.split(',')])
    if args.no_expanded:
        ds.set_expanded(False)
    ndata_local = y.shape[1]
    n_act = int(mask.sum())
    K = args.candidates
    sampler = ClockSampler(device)
    sampler.start()
    peak, peak_src = hbm_peak()

    # ---- device-resident timing (value) -------------------------------------
    ds.set_mask(mask)
    ds.stage_params(pts)
    for _ in range(max(args.warmup, 3)):
        ds.launch_clike(0.01, -0.5)
    ds.sync()
    # the GPU has been idle during the upload: let the clocks come up (~0.1 s of launches) before
    # the timed steps, so that `value` is not a measurement of the ramp
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.1:
        for _ in range(20):
            ds.launch_clike(0.01, -0.5)
        ds.sync()
    barrier()
    launches0 = lib.mdns_launch_count()
    ds.timer_start()
    for _ in range(args.steps):
        ds.launch_clike(0.01, -0.5)
    ms = ds.timer_stop()
    launches = lib.mdns_launch_count() - launches0
    kernel_name = lib.mdns_last_kernel().decode()
    barrier()
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    evals_per_step_all = K * n_act * (n_gpus if distributed else 1)
    value = evals_per_step_all / (ms_per_step * 1e-3)
    # the same over >= 2 s back to back: the board sits at its power cap
    sustained = None
    if args.sustained_s > 0:
        s_ms, s_n = sustained_time(ds, args.sustained_s, ms_per_step)
        s_ms = max_over_ranks(s_ms)
        sustained = {'ms_per_step': s_ms, 'steps': s_n, 'seconds': s_ms * s_n * 1e-3}
        if K == 16 and not distributed and not inproc_devices:
            # the alternative operating point: two passes of 8 candidates
            ds.stage_params(pts[:8])
            t8 = device_time(ds, 20)
            s8_ms, s8_n = sustained_time(ds, args.sustained_s / 2, t8)
            sustained['two_passes_of_8'] = {'burst_ms_per_16': 2 * t8, 'sustained_ms_per_16': 2 * s8_ms}
            ds.stage_params(pts)

    # ---- optional device-side sweep over the candidate count ----------------
    sweep = None
    if args.sweep and not distributed:
        sweep = []
        for Ks in (1, 2, 4, 8, 16, 32, 64, 400):
            ds.stage_params(synth.parameter_points(Ks, seed=7))
            t = device_time(ds, max(3, min(args.steps, 4000 // Ks)))
            gbs = algorithmic_bytes(n_act, ndata_local, args.nx, Ks) / (t * 1e-3) / 1e9
            sweep.append({'candidates': Ks, 'ms_per_step': t, 'evals_per_s': Ks * n_act / (t * 1e-3),
                          'hbm_gbs': gbs, 'hbm_frac': gbs / peak,
                          'kernel': lib.mdns_last_kernel().decode()})
        ds.stage_params(pts)

    # ---- end to end through the public callable (host buffers) -------------
    log_list = [tuple(p) for p in log_pts]
    for _ in range(3):
        L = f.batch(log_list, mask) if K > 1 else f(log_list[0], mask)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        L = f.batch(log_list, mask) if K > 1 else f(log_list[0], mask)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    L = numpy.array(L, copy=True).reshape((K, n_act))     # off the recycled pinned block
    assert L.size == K * n_act and numpy.isfinite(L).all()
    e2e_value = evals_per_step_all * args.steps / e2e_s

    # the speculative batch of the constrained draw (hiermetriclearn.py:181-196) through
    # ResidentDataset.begin_draw / draw_batch: mask and thresholds are constant during one
    # draw_constrained call (hiermetriclearn.py:173-211) and are staged once, outside the timed
    # loop; per step the K parameter points go in, the accept counts (summed over the ranks inside
    # the shim) and the logL vector of the first accepted candidate come out.  Thresholds are set
    # half way between the best and the second-best candidate wherever the LAST candidate wins
    # (and out of reach elsewhere): all K are consumed, as in the reference's one-at-a-time loop,
    # and no decision hinges on the last bits of a logL.
    fa = None
    if K > 1:
        wins = sum_over_ranks(numpy.bincount(numpy.argmax(L, axis=0), minlength=K))
        order = numpy.argsort(wins, kind='stable')          # most frequent winner last, all ranks
        pts_fa = numpy.ascontiguousarray(pts[order])
        L_fa = L[order]
        srt = numpy.sort(L_fa, axis=0)
        top, second = srt[-1], srt[-2]
        out_of_reach = top + 1e-6 * numpy.abs(top) + 1e-6
        sure = (numpy.argmax(L_fa, axis=0) == K - 1) & (top - second > 1e-6 * numpy.abs(top))
        Lmins = numpy.where(sure, 0.5 * (top + second), out_of_reach)
        want_counts = sum_over_ranks((L_fa > Lmins).sum(axis=1)).astype(int)
        want_k = int(numpy.nonzero(want_counts)[0][0]) if want_counts.any() else -1

        def timed(fn):
            for _ in range(3):
                r = fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                r = fn()
            dt = time.perf_counter() - t0
            barrier()
            return max_over_ranks(dt), r

        ds.begin_draw(mask, Lmins)
        fa_s, (k_acc, L_acc, counts) = timed(lambda: ds.draw_batch(pts_fa, 0.01))
        ok = bool(k_acc == want_k and numpy.array_equal(counts, want_counts) and
                  (k_acc < 0 or numpy.allclose(L_acc, L_fa[k_acc], rtol=1e-12, atol=0)))
        # the common outcome of a speculative batch in a long rejection chain: nothing accepted,
        # nothing but the counts comes back
        ds.begin_draw(mask, out_of_reach)
        rej_s, (k_rej, _, c_rej) = timed(lambda: ds.draw_batch(pts_fa, 0.01))
        ok = ok and bool(k_rej == -1 and not numpy.any(c_rej))
        ds.begin_draw(mask, Lmins)
        # counts only: the consumer is on the device (live-point table)
        dc_s, dcounts = timed(lambda: ds.draw_counts(pts_fa, 0.01))
        ok = ok and bool(numpy.array_equal(dcounts, want_counts))
        # the sparse form: only the accepting data sets of the accepted candidate come back.
        # Thresholds raised so that the accepted candidate wins for ~1 % of the data sets (late in
        # a run a new point is accepted for few data sets; early for most of them -- the dense
        # form above is that case).
        margin = numpy.where(sure, L_fa[K - 1] - Lmins, -1.0)
        cut = float(numpy.quantile(margin[sure], 0.98)) if sure.any() else 0.0
        Lmins_s = numpy.where(margin > cut * 1.01, Lmins, out_of_reach)
        ds.begin_draw(mask, Lmins_s)
        fs_s, (ks, js, Ljs, cs) = timed(lambda: ds.draw_batch_sparse(pts_fa, 0.01))
        want_j = numpy.nonzero(L_fa[K - 1] > Lmins_s)[0]
        any_global = sum_over_ranks([len(want_j)])[0] > 0
        sparse_ok = bool(ks == (K - 1 if any_global else -1) and
                         (ks < 0 or (numpy.array_equal(js, want_j) and
                                     numpy.allclose(Ljs, L_fa[K - 1][want_j], rtol=1e-12, atol=0))))
        nsh = n_gpus if distributed else 1
        fa = {'value': evals_per_step_all * args.steps / fa_s, 'unit': UNIT,
              'ms_per_step': 1e3 * fa_s / args.steps, 'accepted_candidate': int(k_acc),
              'matches_full_matrix': ok,
              'h2d_bytes_per_step': K * 24 * nsh,
              'd2h_bytes_per_step': (n_act * 8 + K * 4) * nsh,
              'staged_once_per_draw_bytes': (ndata_local + n_act * 8) * nsh,
              'exchange': ('ncclAllReduce of the K accept counts on the shard stream inside '
                           'mdns_clike_first_accept (communicator: mdns_comm_init)'
                           if distributed else 'single process: none'),
              'api': 'ResidentDataset.begin_draw(data_mask, Lmins) once, then '
                     'draw_batch(params, noise) per step (accept test fused into the likelihood '
                     'kernel, decision on the device; the rows of the accepted candidate are '
                     'fetched only once a chunk decision names one, overlapped with the next chunk)',
              'none_accepted': {'value': evals_per_step_all * args.steps / rej_s, 'unit': UNIT,
                                'ms_per_step': 1e3 * rej_s / args.steps,
                                'd2h_bytes_per_step': K * 4 * nsh,
                                'note': 'same call, thresholds out of reach: every candidate rejected'},
              'sparse': {'value': evals_per_step_all * args.steps / fs_s, 'unit': UNIT,
                         'ms_per_step': 1e3 * fs_s / args.steps,
                         'accepting_data_sets': int(len(want_j)), 'matches_full_matrix': sparse_ok,
                         'd2h_bytes_per_step': (int(len(want_j)) * 12 + K * 4) * nsh,
                         'api': 'draw_batch_sparse(params, noise): indices and logL of the data '
                                'sets the accepted candidate is accepted for '
                                '(multi_nested_sampler.py:482-485); global decision'},
              'device_consumer': {'value': evals_per_step_all * args.steps / dc_s, 'unit': UNIT,
                                  'ms_per_step': 1e3 * dc_s / args.steps,
                                  'd2h_bytes_per_step': K * 4 * nsh,
                                  'api': 'accept counts only (no candidate accepted / the vector is '
                                         'consumed on the device by mdns_livetable_*)'}}
    shards = (n_gpus if distributed else 1)
    # bytes that really cross PCIe per step: the K parameter points; an all-true mask is
    # recognised by a host scan and needs no device list, a partial mask is uploaded once and
    # recognised (memcmp) when it comes again -- both checks run inside the timed region
    h2d = K * 24 * shards
    d2h = K * n_act * 8 * shards

    # ---- roofline of the dominant kernel (per GPU) ---------------------------
    per_gpu_n = ndata_local // (len(inproc_devices) if inproc_devices else 1)
    per_gpu_act = n_act // (len(inproc_devices) if inproc_devices else 1)
    bytes_per_launch = algorithmic_bytes(per_gpu_act, per_gpu_n, args.nx, K)
    achieved = bytes_per_launch / (ms_per_step * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': ncu_traffic(args.nx, per_gpu_n, K),
                'kernel': kernel_name, 'algorithmic_bytes_per_launch': bytes_per_launch,
                'peak_source': peak_src,
                'note': 'duration = whole step (line model + dominant kernel + fix-up launch), '
                        'CUDA events on the shim stream'}
    if sustained:
        roofline['sustained_frac'] = bytes_per_launch / (sustained['ms_per_step'] * 1e-3) / 1e9 / peak
        roofline['sustained'] = sustained
        if 'two_passes_of_8' in sustained:
            b8 = 2 * algorithmic_bytes(per_gpu_act, per_gpu_n, args.nx, 8)
            sustained['two_passes_of_8']['note'] = (
                'evals/s is what counts: 16 candidates in one pass take %.3f ms sustained, in two '
                'passes of 8 %.3f ms (the data is streamed twice, %.2f GB instead of %.2f GB)'
                % (sustained['ms_per_step'], sustained['two_passes_of_8']['sustained_ms_per_16'],
                   b8 / 1e9, bytes_per_launch / 1e9))

    extras = {}
    if not args.no_extras:
        # free the headline data set before the 8 GB of configs[3]
        fa_keep = fa
        f.dataset.close()
        del f, ds, y, L
        extras['configs3'] = bench_configs3(args, rank, world, local, peak)
        if rank == 0 and n_gpus == 1:
            extras['sweep_n'] = bench_sweep_n(args, peak)
            extras['muse'] = bench_muse(peak)
            extras['neighbors'] = bench_neighbors()
        fa = fa_keep
    clocks = sampler.stop()

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            n_s = min(args.ndata, 500000)
            xs, ys, _ = synth.horns(n_s, nx=args.nx, legacy=False, seed=1000)
            ms_ = synth.masks(n_s, seed=11)[args.mask]
            threads = host_threads()
            rate, dt, reps = cpu_reference_rate(xs, ys, pts, ms_, budget_s=10.0, threads=threads)
            rate1, dt1, reps1 = cpu_reference_rate(xs, ys[:, :100000].copy(), pts,
                                                   numpy.ascontiguousarray(ms_[:100000]),
                                                   budget_s=3.0, threads=1)
            cpu = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'reference',
                   'sample': '%d data sets x %d channels x %d candidates x %d repetitions '
                             '(%.1f s), %s' % (n_s, args.nx, K, reps, dt, REF_NOTE),
                   'single_thread_value': rate1}
        except Exception as e:      # the oracle must exist; say why if it does not
            cpu = {'value': None, 'unit': UNIT, 'cores': 1, 'kind': 'reference',
                   'sample': 'failed: %s' % e}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args), 'roofline': roofline,
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': 1e3 * e2e_s / args.steps,
                    'mask_bytes_scanned_on_host_per_step': ndata_local * shards,
                    'api': 'massivedatans_b200.likelihood.make_multi_loglikelihood(...)'
                           + ('.batch' if K > 1 else '') + '(params, data_mask), host numpy in/out',
                    'first_accept': fa},
            'gpu_launches': int(launches), 'clocks': clocks,
            'ranks': {'processes': world, 'numa_bound': bool(bound),
                      'plumbing': 'NCCL communicator inside libmdns_b200.so (dlopen), id handed over '
                                  'by TCP on MASTER_ADDR; no torch in the process'},
        }
        line.update(extras)
        if sweep is not None:
            line['sweep'] = sweep
        print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def bench_configs3(args, rank, world, local, peak):
    """BASELINE configs[3] as written (gen_realistic.py:16-50 with the 10 000 cap lifted): 1e6 data
    sets x 1000 channels in total, strong-scaled: every rank holds and scores 1e6 / world of them."""
    from massivedatans_b200 import sharding, synth
    from massivedatans_b200.likelihood import ResidentDataset
    total = args.configs3_ndata
    i0, n = sharding.shard_ranges(total, world)[rank]
    nx = 1000
    t0 = time.perf_counter()
    x, y = synth.realistic_fast(n, nx=nx, seed=1 + rank, threads=min(host_threads(), 32))
    gen_s = time.perf_counter() - t0
    ds = ResidentDataset(x, y, devices=[local])
    del y
    if world > 1:
        sharding.init_comm_from_env(ds, port=int(os.environ.get('MASTER_PORT', '29500')) + 41)
    out = {'workload': 'gen_realistic-style spectra: %d data sets x %d channels in TOTAL, '
                       '%d per GPU (strong scaling), all active' % (total, nx, n),
           'scaling': 'strong', 'ndata_total': total, 'ndata_per_gpu': n, 'nx': nx,
           'host_generation_s': gen_s, 'by_candidates': []}
    ds.set_mask(None)
    for K in (1, 8, 16):
        ds.stage_params(synth.parameter_points(K, seed=7))
        t = device_time(ds, 20 if n >= 250000 else 50)
        ds.sync()
        if world > 1:
            t = float(ds.comm_allreduce([t], op='max')[0])
        b = algorithmic_bytes(n, n, nx, K)
        out['by_candidates'].append({
            'candidates': K, 'ms_per_step': t, 'value': K * total / (t * 1e-3), 'unit': UNIT,
            'hbm_gbs_per_gpu': b / (t * 1e-3) / 1e9, 'roofline_frac': b / (t * 1e-3) / 1e9 / peak,
            'kernel': ds._lib.mdns_last_kernel().decode()})
    head = out['by_candidates'][-1]
    out.update({'value': head['value'], 'unit': UNIT, 'ms_per_step': head['ms_per_step'],
                'roofline_frac': head['roofline_frac'], 'candidates_per_step': 16})
    # the accept pass with the exchange, sparse result (what the sampler consumes)
    K = 16
    pts = synth.parameter_points(K, seed=7)
    L = ds.loglike_batch(pts, None, 0.01)
    best = numpy.array(L.max(axis=0))
    ds.begin_draw(None, best - 1e-6 * numpy.abs(best))       # the best candidate of every data set accepts
    for _ in range(3):
        ds.draw_batch_sparse(pts, 0.01)
    ds.sync()
    if world > 1:
        ds.comm_allreduce([0.0])
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        ks, js, Ljs, cs = ds.draw_batch_sparse(pts, 0.01)
    dt = time.perf_counter() - t0
    if world > 1:
        dt = float(ds.comm_allreduce([dt], op='max')[0])
    out['first_accept_sparse'] = {'ms_per_step': 1e3 * dt / reps, 'value': K * total * reps / dt,
                                  'unit': UNIT, 'accepted_candidate': int(ks),
                                  'accepting_data_sets_this_rank': int(len(js))}
    ds.close()
    return out


def bench_sweep_n(args, peak):
    """The metric at N = 1e4, 1e5, 1e6 (K = 1 and 16), one GPU.  1e4 x 200 (16 MB) and 1e5 x 200
    (160 MB) fit (mostly) in the 126 MB L2: timed launch by launch with the L2 flushed in between,
    and once more back to back (L2-hot) -- the latency-bound regime of the small configs."""
    from massivedatans_b200 import synth
    from massivedatans_b200.likelihood import ResidentDataset
    rows = []
    for n in (10000, 100000, 1000000):
        x, y, _ = (synth.horns(n, nx=args.nx, legacy=False, seed=n) if n != 10000 else
                   synth.nothing(n, nx=args.nx, legacy=False) + (None,))
        ds = ResidentDataset(x, y)
        ds.set_mask(None)
        for K in (1, 16):
            ds.stage_params(synth.parameter_points(K, seed=7))
            steps = 40 if n < 1000000 else 20
            cold = device_time(ds, steps, flush=True)
            hot = device_time(ds, steps)
            b = algorithmic_bytes(n, n, args.nx, K)
            rows.append({'ndata': n, 'candidates': K, 'ms_l2_flushed': cold, 'ms_back_to_back': hot,
                         'value': K * n / (cold * 1e-3), 'value_back_to_back': K * n / (hot * 1e-3),
                         'unit': UNIT, 'roofline_frac_flushed': b / (cold * 1e-3) / 1e9 / peak,
                         'resident_mb': n * args.nx * 8 / 1e6,
                         'kernel': ds._lib.mdns_last_kernel().decode()})
        ds.close()
        del y
    return rows


def bench_muse(peak):
    """cmuselike path on a cube of the reference shape (4223 spectra x 3600 channels, y and
    variance: 243 MB, beyond the L2), K = 1, 4, 16 model spectra, all spectra active."""
    from massivedatans_b200 import synth
    from massivedatans_b200.likelihood import ResidentDataset
    from oracle import ref
    ndata, nspec = synth.MUSE_NDATA, synth.MUSE_NSPEC
    y, v, t = synth.muse(ndata=ndata, nspec=nspec)
    ds = ResidentDataset(None, y, variance=v)
    mask = numpy.ones(ndata, dtype=bool)
    ds.set_mask(mask)
    out = {'workload': 'MUSE-shaped cube %d x %d (synth.muse)' % (ndata, nspec), 'by_candidates': []}
    for K in (1, 4, 16):
        ypreds = numpy.array([synth.muse_template(nspec, phase=0.1 * k) for k in range(K)])
        ds.stage_spectra(ypreds)
        for _ in range(3):
            ds.launch_muse()
        ds.sync()
        ds.timer_start()
        reps = 50
        for _ in range(reps):
            ds.launch_muse()
        tms = ds.timer_stop() / reps
        b = muse_bytes(ndata, ndata, nspec, K)
        out['by_candidates'].append({'candidates': K, 'ms_per_step': tms,
                                     'value': K * ndata / (tms * 1e-3), 'unit': UNIT,
                                     'hbm_gbs': b / (tms * 1e-3) / 1e9,
                                     'roofline_frac': b / (tms * 1e-3) / 1e9 / peak,
                                     'kernel': ds._lib.mdns_last_kernel().decode()})
    # parity on the spot + the reference C beside it
    Lout = numpy.zeros((1, ndata))
    ds.muse_loglike(t, mask, Lout)
    want = ref.cmuselike(y, v, t, mask)
    out['max_rel_err_vs_reference'] = float(numpy.max(numpy.abs(Lout[0] - want) / numpy.abs(want)))
    t0 = time.perf_counter()
    ref.cmuselike(y, v, t, mask)
    serial = time.perf_counter() - t0
    os.environ.setdefault('OMP_NUM_THREADS', str(host_threads()))
    ref.cmuselike(y, v, t, mask, parallel=True)
    t0 = time.perf_counter()
    for _ in range(3):
        ref.cmuselike(y, v, t, mask, parallel=True)
    par = (time.perf_counter() - t0) / 3
    out['cpu_baseline'] = {'kind': 'reference', 'unit': UNIT,
                           'serial_value': ndata / serial, 'serial_ms': 1e3 * serial,
                           'openmp_value': ndata / par, 'openmp_ms': 1e3 * par,
                           'cores': int(os.environ['OMP_NUM_THREADS']),
                           'sample': 'cmuselike.so / cmuselike-parallel.so (oracle/_ref), one call each'}
    ds.close()
    return out


def bench_neighbors():
    """RadFriends neighbour tests at the shapes the sampler issues (radfriendsregion.py:124: 1000
    candidates per round; hiermetriclearn.py:106,130: 10 000), members resident on the device:
    device time of one count / any call (CUDA events on the region's stream, includes the H2D of
    the candidates and the D2H of the counts) and the reference C on the host."""
    import ctypes
    from massivedatans_b200 import _lib, synth
    from oracle import ref
    lib = _lib.load()
    rows = []
    for n, m in ((400, 1000), (5000, 10000), (50000, 100000)):
        xx, yy = synth.members_and_candidates(n, m, 3)
        chosen = synth.bootstrap_chosen(n, 10, numpy.random.RandomState(1))
        rg = ctypes.c_void_p()
        _lib.check(lib.mdns_region_create(0, ctypes.byref(rg)), 'mdns_region_create')
        _lib.check(lib.mdns_region_set_members(rg, xx.ctypes.data, n, 3), 'mdns_region_set_members')
        r = ctypes.c_double()
        _lib.check(lib.mdns_region_bootstrapped_maxdistance(rg, chosen.ctypes.data, 10, ctypes.byref(r)),
                   'mdns_region_bootstrapped_maxdistance')
        radius = r.value
        row = {'members': n, 'candidates': m, 'ndim': 3, 'radius': radius}
        for name, countmax in (('count', 0), ('any', 1)):
            out = numpy.zeros(m)
            ms = ctypes.c_float()
            reps = 20 if n <= 5000 else 5
            for _ in range(2):
                out[:] = 0
                lib.mdns_region_count_within(rg, radius, yy.ctypes.data, m, out.ctypes.data, countmax)
            total = 0.0
            for _ in range(reps):
                out[:] = 0
                lib.mdns_region_timer_start(rg)
                _lib.check(lib.mdns_region_count_within(rg, radius, yy.ctypes.data, m, out.ctypes.data,
                                                        countmax), 'mdns_region_count_within')
                lib.mdns_region_timer_stop(rg, ctypes.byref(ms))
                total += ms.value
            t = total / reps
            want = ref.count_within_distance_of_raw(xx, radius, yy, numpy.zeros(m), countmax)
            t0 = time.perf_counter()
            ref.count_within_distance_of_raw(xx, radius, yy, numpy.zeros(m), countmax)
            cpu = time.perf_counter() - t0
            row[name] = {'ms': t, 'pair_tests_per_s': n * m / (t * 1e-3) if countmax == 0 else None,
                         'bit_exact': bool(numpy.array_equal(out, want)),
                         'reference_ms': 1e3 * cpu, 'speedup': cpu / (t * 1e-3),
                         'kernel': lib.mdns_last_kernel().decode()}
        lib.mdns_region_timer_start(rg)
        _lib.check(lib.mdns_region_bootstrapped_maxdistance(rg, chosen.ctypes.data, 10, ctypes.byref(r)),
                   'mdns_region_bootstrapped_maxdistance')
        ms = ctypes.c_float()
        lib.mdns_region_timer_stop(rg, ctypes.byref(ms))
        row['bootstrapped_maxdistance'] = {'ms': ms.value, 'rounds': 10}
        if n <= 5000:
            t0 = time.perf_counter()
            want_r = ref.bootstrapped_maxdistance_chosen(xx, chosen)
            row['bootstrapped_maxdistance'].update({'reference_ms': 1e3 * (time.perf_counter() - t0),
                                                    'bit_exact': bool(want_r == radius)})
        rows.append(row)
        lib.mdns_region_destroy(rg)
    return rows


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
