"""bench.py contract checks that need no GPU: the reference arm (the reference's own clike.so on
the host cores) prints one JSON line with the keys the driver reads, and the roofline helper
uses the algorithmic bytes of SURVEY.md section 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                                   '--steps', '1', '--warmup', '1', '--ndata', '3000'],
                                  cwd=ROOT, timeout=300).decode()
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['e2e']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] >= 1
    assert d['config']['candidates_per_step'] == 16 and 'workload' in d['config']
    # the arm runs the N its config line states (round 1 timed a 10x smaller sample under that label)
    assert d['config']['ndata_per_gpu'] == 3000 and '3000 data sets' in d['cpu_baseline']['sample']
    assert d['gpu_launches'] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                                   '--gpus', '2', '--steps', '1', '--warmup', '1'],
                                  cwd=ROOT, env=env, timeout=120).decode()
    assert out.strip() == ''


def test_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    # n_act*C*8 + K*C*8 + K*n_act*8 + N ; 1609 B per evaluation at K=1, C=200
    assert bench.algorithmic_bytes(1, 1, 200, 1) == 200 * 8 + 200 * 8 + 8 + 1
    assert bench.algorithmic_bytes(10 ** 6, 10 ** 6, 200, 1) / 10 ** 6 == 1609.0016
    assert bench.algorithmic_bytes(10 ** 6, 10 ** 6, 200, 16) == 1729025600


def test_muse_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    # cmuselike-type: y and 1/v each read once -> 57 609 B per evaluation at K=1, C=3600
    assert bench.muse_bytes(1, 1, 3600, 1) == 3600 * 16 + 3600 * 8 + 8 + 1
    assert abs(bench.muse_bytes(4223, 4223, 3600, 1) / 4223.0 - 57609.0) < 8.0


def test_realistic_fast_matches_the_generator_distributions():
    sys.path.insert(0, ROOT)
    import numpy
    from massivedatans_b200 import synth
    x, y = synth.realistic_fast(3000, nx=1000, seed=1, threads=3)
    x2, y2, _ = synth.realistic(3000, nx=1000, seed=1)
    assert y.shape == y2.shape == (1000, 3000) and y.flags['C_CONTIGUOUS'] and numpy.array_equal(x, x2)
    # same line parameters (drawn up front from the same RandomState), independent noise streams:
    # the spectra agree to the noise level
    assert numpy.abs(y - y2).max() < 12 * synth.NOISE_LEVEL
    assert abs(numpy.std(y - y2) / (numpy.sqrt(2) * synth.NOISE_LEVEL) - 1) < 0.05
    # chunk boundaries and thread count do not change the result
    x3, y3 = synth.realistic_fast(3000, nx=1000, seed=1, threads=1)
    assert numpy.array_equal(y, y3)
