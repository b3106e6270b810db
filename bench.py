#!/usr/bin/env python
"""bench.py -- headline measurement of the massivedatans hot path on B200.

Metric (BASELINE.json): model x data-set logL evaluations per second, and % of the HBM
roofline of the batched likelihood kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU)

One "step" = one pass of the batched likelihood over one batch: `--candidates` parameter
points scored against every data set of the GPU's resident shard (gensimple_horns-style
synthetic spectra, C=200 channels, N=1e6 data sets per GPU -> 1.6 GB, far beyond the 126 MB
L2, so every step streams from HBM).  Data sets are sharded contiguously, one shard per
rank, no data-path collective: weak scaling.

Keys of the JSON line (rank 0 prints exactly one):
  value      evals/s with inputs (data, mask, parameter points) resident in HBM, device-timed
             with CUDA events on the shim's stream, max over ranks
  e2e        the same metric through the public Python callable with HOST buffers: per step
             the mask and parameter points go host->device and the K x N logL matrix comes back
             (PCIe-bound).  e2e.first_accept: the same batch through the device-side form of
             the constrained draw's loop (hiermetriclearn.py:181-196): K parameter points in,
             K accept counts + the accepted candidate's logL vector out (for N > 1 with the
             all-reduce of the K counts between the ranks); .sparse: only the accepting data
             sets' indices and logL come back
  roofline   algorithmic bytes per step / device time per step vs MEASURED_PEAKS.json hbm_gbs;
             traffic = dram bytes of the dominant kernel from the committed ncu capture
  cpu_baseline  the reference's own unmodified clike.so (oracle/_ref; the serial build
             sample.py:81-84 loads, its OpenMP variant is racy, clike.c:32), one contiguous
             data-set shard per host thread, all host cores, bounded sample
The reference arm (--impl reference) times that same CPU implementation per step.
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
    ap.add_argument('--ref-ndata', type=int, default=100000,
                    help='data sets per step of the CPU reference arm (bounded sample)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--tuning', default='', help='lanes,unroll,ktile override (experiments)')
    ap.add_argument('--sweep', action='store_true',
                    help='also time K = 1 ... 400 device-side and add them under "sweep"')
    ap.add_argument('--no-expanded', action='store_true',
                    help='keep candidate batches on the direct-form kernels')
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
    """--impl reference: the reference's own CPU implementation, rank 0 only."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from massivedatans_b200 import synth
    n = args.ref_ndata
    x, y, _ = synth.horns(n, nx=args.nx, legacy=False, seed=1000)
    pts = synth.parameter_points(args.candidates, seed=7)
    mask = synth.masks(n, seed=11)[args.mask]
    n_act = int(mask.sum())
    threads = host_threads()
    pool = ReferencePool(x, y, mask, threads)
    for _ in range(max(args.warmup, 1)):
        pool.step(pts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.step(pts)
    dt = time.perf_counter() - t0
    pool.close()
    value = args.steps * args.candidates * n_act / dt
    sample = ('%d data sets x %d channels x %d candidates per step on %d host threads, %s'
              % (n, args.nx, args.candidates, pool.threads, REF_NOTE))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
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
                     '%d candidate parameter points per step (BASELINE configs[2]/[3] shape)'
                     % (args.ndata, args.nx, args.mask, args.candidates)),
        'ndata_per_gpu': args.ndata, 'nx': args.nx, 'candidates_per_step': args.candidates,
        'mask': args.mask,
        'l2': 'inputs larger than L2 (%.2f GB resident per GPU vs 126 MB L2), no flush needed'
              % (args.ndata * args.nx * 8 / 1e9),
        'sharding': 'contiguous data-set ranges, one shard per GPU, no data-path collective',
    }


def run_ours(args):
    rank, world, local = dist_env()
    import torch
    import torch.distributed as dist
    distributed = world > 1
    inproc_devices = None
    if distributed:
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        n_gpus = world
        device = local
    else:
        n_gpus = args.gpus
        device = 0
        if n_gpus > 1:        # not under torchrun: one process drives N shards
            inproc_devices = list(range(n_gpus))

    from massivedatans_b200 import _lib
    from massivedatans_b200.likelihood import ResidentDataset, make_multi_loglikelihood
    lib = _lib.load()
    _lib.require_device()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not distributed:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    if args.tuning:
        ds.set_tuning(*[int(v) for v in args.tuning.split(',')])
    if args.no_expanded:
        ds.set_expanded(False)
    ndata_local = y.shape[1]
    n_act = int(mask.sum())
    K = args.candidates
    sampler = ClockSampler(device)
    sampler.start()

    # ---- device-resident timing (value) -------------------------------------
    ds.set_mask(mask)
    ds.stage_params(pts)
    for _ in range(max(args.warmup, 3)):
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

    # ---- optional device-side sweep over the candidate count ----------------
    sweep = None
    if args.sweep and not distributed:
        from massivedatans_b200 import synth
        sweep = []
        peak_gbs = hbm_peak()[0]
        for Ks in (1, 2, 4, 8, 16, 32, 64, 400):
            ds.stage_params(synth.parameter_points(Ks, seed=7))
            for _ in range(3):
                ds.launch_clike(0.01, -0.5)
            ds.sync()
            reps = max(3, min(args.steps, 4000 // Ks))
            ds.timer_start()
            for _ in range(reps):
                ds.launch_clike(0.01, -0.5)
            t = ds.timer_stop() / reps
            gbs = algorithmic_bytes(n_act, ndata_local, args.nx, Ks) / (t * 1e-3) / 1e9
            sweep.append({'candidates': Ks, 'ms_per_step': t, 'evals_per_s': Ks * n_act / (t * 1e-3),
                          'hbm_gbs': gbs, 'hbm_frac': gbs / peak_gbs,
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
    # loop; per step the K parameter points go in, the accept counts and the logL vector of the
    # first accepted candidate come out.  Thresholds are set so that only the LAST candidate is
    # accepted: all K are consumed, as in the reference's one-at-a-time loop.
    fa = None
    if K > 1:
        wins = numpy.bincount(numpy.argmax(L, axis=0), minlength=K)
        if distributed:       # the same candidate order on every rank
            tw = torch.as_tensor(wins, dtype=torch.int64).cuda()
            dist.all_reduce(tw)
            wins = tw.cpu().numpy()
        order = numpy.argsort(wins, kind='stable')          # most frequent winner last
        pts_fa = numpy.ascontiguousarray(pts[order])
        L_fa = L[order]
        Lmins = numpy.max(L_fa[:K - 1], axis=0)
        ds.begin_draw(mask, Lmins)
        from massivedatans_b200 import sharding

        def fa_step():
            if not distributed:
                return ds.draw_batch(pts_fa, 0.01)
            # one process per GPU: local counts -> all-reduce of K integers over NCCL (the one
            # exchange step of the sharded path) -> every rank fetches the globally first
            # accepted candidate from its own shard
            c = ds.draw_counts(pts_fa, 0.01)
            k, tot = sharding.global_first_accepted(c, device=torch.device('cuda', local))
            return k, (ds.fetch_candidate(k) if k >= 0 else None), tot

        for _ in range(3):
            k_acc, L_acc, counts = fa_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            k_acc, L_acc, counts = fa_step()
        fa_s = time.perf_counter() - t0
        barrier()
        fa_s = max_over_ranks(fa_s)
        # the sparse form: only the accepting data sets of the accepted candidate come back.
        # Thresholds raised so that the accepted candidate wins for ~1 % of the data sets (late in
        # a run a new point is accepted for few data sets; early for most of them -- the dense
        # form above is that case).
        margin = L_fa[K - 1] - Lmins
        Lmins_s = Lmins + max(float(numpy.quantile(margin, 0.99)), 0.0)
        ds.begin_draw(mask, Lmins_s)
        for _ in range(3):
            ks, js, Ljs, cs = ds.draw_batch_sparse(pts_fa, 0.01)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ks, js, Ljs, cs = ds.draw_batch_sparse(pts_fa, 0.01)
        fs_s = time.perf_counter() - t0
        barrier()
        fs_s = max_over_ranks(fs_s)
        want_j = numpy.nonzero(L_fa[K - 1] > Lmins_s)[0]
        sparse_ok = bool(ks == (K - 1 if len(want_j) else -1) and
                         (ks < 0 or (numpy.array_equal(js, want_j) and
                                     numpy.array_equal(Ljs, L_fa[K - 1][want_j]))))
        want_k = K - 1 if (distributed or (L_fa[K - 1] > Lmins).any()) else -1
        ok = bool(k_acc == want_k and (k_acc < 0 or numpy.array_equal(L_acc, L_fa[K - 1])))
        nsh = n_gpus if distributed else 1
        fa = {'value': evals_per_step_all * args.steps / fa_s, 'unit': UNIT,
              'ms_per_step': 1e3 * fa_s / args.steps, 'accepted_candidate': int(k_acc),
              'matches_full_matrix': ok,
              'h2d_bytes_per_step': K * 24 * nsh,
              'd2h_bytes_per_step': (n_act * 8 + K * 4) * nsh,
              'staged_once_per_draw_bytes': (ndata_local + n_act * 8) * nsh,
              'api': ('ResidentDataset.begin_draw(data_mask, Lmins) once, then per step draw_counts + '
                      'NCCL all-reduce of the K accept counts + fetch_candidate'
                      if distributed else
                      'ResidentDataset.begin_draw(data_mask, Lmins) once, then '
                      'draw_batch(params, noise) per step'),
              'sparse': {'value': evals_per_step_all * args.steps / fs_s, 'unit': UNIT,
                         'ms_per_step': 1e3 * fs_s / args.steps,
                         'accepting_data_sets': int(len(want_j)), 'matches_full_matrix': sparse_ok,
                         'd2h_bytes_per_step': (int(len(want_j)) * 12 + K * 4) * nsh,
                         'api': 'draw_batch_sparse(params, noise): indices and logL of the data '
                                'sets the accepted candidate is accepted for '
                                '(multi_nested_sampler.py:482-485)'}}
    shards = (n_gpus if distributed else 1)
    # bytes that really cross PCIe per step: the K parameter points; an all-true mask is
    # recognised by a host scan and needs no device list, a partial mask is uploaded once and
    # recognised (memcmp) when it comes again -- both checks run inside the timed region
    mask_bytes_uploaded_each_step = 0
    h2d = (K * 24 + mask_bytes_uploaded_each_step) * shards
    d2h = K * n_act * 8 * shards
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (per GPU) ---------------------------
    peak, peak_src = hbm_peak()
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

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            n_s = min(ndata_local, 500000)
            ys = numpy.ascontiguousarray(y[:, :n_s])
            threads = host_threads()
            rate, dt, reps = cpu_reference_rate(x, ys, pts, numpy.ascontiguousarray(mask[:n_s]),
                                                budget_s=10.0, threads=threads)
            rate1, dt1, reps1 = cpu_reference_rate(x, ys[:, :100000].copy(), pts,
                                                   numpy.ascontiguousarray(mask[:100000]),
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
        }
        if sweep is not None:
            line['sweep'] = sweep
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
