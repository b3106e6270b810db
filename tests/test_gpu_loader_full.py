"""The loader half of SURVEY §8 f4 at the size it exists for: BASELINE configs[3] (gen_realistic,
1e6 data sets x 1000 channels = 8 GB) made resident from a .npy file without the matrix ever
being in host memory (sample.py:27-31 reads all of `y` into RAM).

Default size is one GPU's share of configs[3] on eight GPUs (125 000 x 1000, 1 GB: seconds);
MDNS_LOADER_FULL=1 runs the whole 8 GB matrix and, run alone in a fresh process, records the
peak host RSS in gpurun_out/r02_loader.json:

    MDNS_LOADER_FULL=1 python -m pytest tests/test_gpu_loader_full.py -q -m gpu
"""
import json
import os
import time

import numpy
import pytest

from conftest import rel_err
from massivedatans_b200 import synth
from massivedatans_b200.likelihood import ResidentDataset

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _vm(key):
    for line in open('/proc/self/status'):
        if line.startswith(key + ':'):
            return int(line.split()[1]) / 1024.0       # MB
    return float('nan')


def _write_realistic_npy(path, N, nx, sample, block=8):
    """gen_realistic.py:16-50-style spectra written channel block by channel block ([nx, N] float64,
    numpy.save layout); only `block` channels x N are in memory at a time.  Returns x and the
    columns listed in `sample` ([nx, len(sample)])."""
    x = numpy.linspace(400, 800, nx)
    rs = numpy.random.RandomState(7)
    z = rs.beta(2, 30, size=N) * 2
    rest_wave = 440
    width_broad = 10 ** rs.normal(3, 0.2, size=N) * rest_wave / 300000
    width_narrow = 10 ** rs.normal(1, 0.2, size=N) * rest_wave / 300000
    signal_level = 1. / (rs.power(1, size=N) * 100 + 2)
    height_broad = numpy.where(rs.uniform(size=N) < 0.5, 10 ** rs.normal(0, 0.2, size=N),
                               10 ** rs.normal(-2, 0.2, size=N)) * signal_level
    kept = numpy.empty((nx, len(sample)))
    with open(path, 'wb') as f:
        numpy.lib.format.write_array_header_1_0(
            f, {'descr': '<f8', 'fortran_order': False, 'shape': (nx, N)})
        for c0 in range(0, nx, block):
            c1 = min(nx, c0 + block)
            d = rest_wave - x[c0:c1, None] / (1. + z[None, :])
            ym = height_broad * numpy.exp(-0.5 * (d / width_broad) ** 2)
            ym += signal_level * numpy.exp(-0.5 * (d / width_narrow) ** 2)
            rg = numpy.random.Generator(numpy.random.Philox(key=11, counter=[0, 0, 0, c0]))
            ym += rg.standard_normal(size=ym.shape) * synth.NOISE_LEVEL
            kept[c0:c1] = ym[:, sample]
            f.write(numpy.ascontiguousarray(ym).tobytes())
    return x, kept


def test_configs3_matrix_goes_file_to_hbm_without_a_host_copy(tmp_path, oracle_port):
    full = os.environ.get('MDNS_LOADER_FULL', '') not in ('', '0')
    N, nx = (1000000, 1000) if full else (125000, 1000)
    rs = numpy.random.RandomState(3)
    # sampled data sets: both ends, tile and column-block boundaries of the loader, random ones
    sample = numpy.unique(numpy.concatenate([
        [0, 1, 255, 256, 257, N - 257, N - 256, N - 2, N - 1, 8191, 8192, 8193],
        rs.randint(0, N, size=1500)]))
    path = os.path.join(os.environ.get('MDNS_LOADER_DIR', str(tmp_path)), 'configs3_y.npy')
    t0 = time.perf_counter()
    x, kept = _write_realistic_npy(path, N, nx, sample)
    t_write = time.perf_counter() - t0
    try:
        assert os.path.getsize(path) >= N * nx * 8
        try:                                   # peak-RSS counter back to the current RSS
            with open('/proc/self/clear_refs', 'w') as f:
                f.write('5')
        except OSError:
            pass
        rss_before = _vm('VmRSS')
        t0 = time.perf_counter()
        ds = ResidentDataset.from_npy(x, path)
        t_load = time.perf_counter() - t0
        hwm_after = _vm('VmHWM')
        assert (ds.nx, ds.ndata) == (nx, N)
        # the contract: the matrix never sits in host memory (64 MB pinned block + the CUDA
        # context, whatever the matrix size)
        grown = hwm_after - rss_before
        assert grown < 1024.0, 'host memory grew by %.0f MB while loading' % grown
        # parity on the sampled data sets, K = 16 (the tensor path) and K = 1 (the row kernel)
        mask = numpy.zeros(N, dtype=bool)
        mask[sample] = True
        worst = 0.0
        for K in (16, 1):
            pts = synth.parameter_points(K, seed=5)
            got_all = numpy.array(ds.loglike_batch(pts, None, synth.NOISE_LEVEL, scale=1.0)[:, sample])
            got_masked = numpy.array(ds.loglike_batch(pts, mask, synth.NOISE_LEVEL, scale=1.0))
            want = numpy.array([oracle_port.clike(x, kept, p[0], p[1], p[2], synth.NOISE_LEVEL,
                                                  numpy.ones(len(sample), dtype=bool)) for p in pts])
            worst = max(worst, rel_err(got_all, want), rel_err(got_masked, want))
            assert rel_err(got_all, want) < 1e-10 and rel_err(got_masked, want) < 1e-10
        hwm_end = _vm('VmHWM')
        ds.close()
        if full:
            out = {'ndata': N, 'nx': nx, 'file_gb': os.path.getsize(path) / 1e9,
                   'write_file_s': t_write, 'load_s': t_load,
                   'load_gb_per_s': os.path.getsize(path) / 1e9 / t_load,
                   'host_rss_before_load_mb': rss_before, 'host_peak_rss_after_load_mb': hwm_after,
                   'host_peak_rss_end_mb': hwm_end, 'sampled_data_sets': int(len(sample)),
                   'max_rel_err_vs_oracle': worst,
                   'note': 'VmHWM of the pytest process (python + numpy + CUDA context + the '
                           'generator\'s channel blocks); the 8 GB matrix itself only exists in the '
                           'file and in HBM'}
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(ROOT, 'gpurun_out', 'r02_loader.json'), 'w') as f:
                json.dump(out, f, indent=1)
    finally:
        if os.path.exists(path):
            os.remove(path)
