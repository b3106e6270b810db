#!/usr/bin/env python
"""Programmatic dependent launch of the kernels of a step (model kernel -> likelihood kernel ->
fix-up / finalize) against fully serialised launches (MDNS_NO_PDL=1), same box, alternating
processes.   python tools/r2_pdl_ab.py [rounds]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker():
    import numpy
    import bench
    from massivedatans_b200 import _lib, synth
    from massivedatans_b200.likelihood import ResidentDataset
    lib = _lib.load()
    peak = bench.hbm_peak()[0]
    rows = []
    for n in (10000, 100000, 300000, 1000000):
        x, y, _ = synth.horns(n, nx=200, legacy=False, seed=1000)
        ds = ResidentDataset(x, y)
        ds.set_mask(None)
        for K in (8, 16, 32):
            ds.stage_params(synth.parameter_points(K, seed=7))
            steps = 60 if n < 1000000 else 100
            t = bench.device_time(ds, steps, flush=True)
            tb = bench.device_time(ds, steps, flush=False)
            b = bench.algorithmic_bytes(n, n, 200, K)
            rows.append({'what': 'clike', 'n': n, 'K': K, 'ms_flushed': t, 'ms_back_to_back': tb,
                         'frac_flushed': b / (t * 1e-3) / 1e9 / peak, 'frac_b2b': b / (tb * 1e-3) / 1e9 / peak,
                         'kernel': lib.mdns_last_kernel().decode()})
        ds.close()
    ndata, nspec = synth.MUSE_NDATA, synth.MUSE_NSPEC
    y, v, t = synth.muse(ndata=ndata, nspec=nspec)
    ds = ResidentDataset(None, y, variance=v)
    ds.set_mask(numpy.ones(ndata, dtype=bool))
    for K in (4, 16):
        ypreds = numpy.array([synth.muse_template(nspec, phase=0.1 * k) for k in range(K)])
        ds.stage_spectra(ypreds)
        for _ in range(3):
            ds.launch_muse()
        ds.sync()
        ds.timer_start()
        for _ in range(50):
            ds.launch_muse()
        ms = ds.timer_stop() / 50
        b = bench.muse_bytes(ndata, ndata, nspec, K)
        rows.append({'what': 'muse', 'n': ndata, 'K': K, 'ms_back_to_back': ms,
                     'frac_b2b': b / (ms * 1e-3) / 1e9 / peak, 'kernel': lib.mdns_last_kernel().decode()})
    ds.close()
    print('ROWS ' + json.dumps(rows), flush=True)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    # the switch of the 'serial' arm (default: MDNS_NO_PDL=1); e.g. MDNS_ONE_LAUNCH=0 compares the
    # step with a model kernel against the one whose likelihood kernel builds the spectra itself
    var, val = (sys.argv[2].split('=') + ['1'])[:2] if len(sys.argv) > 2 else ('MDNS_NO_PDL', '1')
    res = {'pdl': [], 'serial': [], 'serial_arm': '%s=%s' % (var, val)}
    for r in range(rounds):
        for arm in ('serial', 'pdl'):
            env = dict(os.environ)
            env.pop(var, None)
            if arm == 'serial':
                env[var] = val
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker'], env=env,
                               capture_output=True, text=True, timeout=600)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith('ROWS ')]
            if p.returncode != 0 or not line:
                print(arm, 'FAILED', p.returncode, p.stdout[-2000:], p.stderr[-2000:], flush=True)
                continue
            res[arm].append(json.loads(line[0][5:]))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'r2_pdl_ab.json' if var == 'MDNS_NO_PDL' else 'r2_ab_%s.json' % var.lower()), 'w'), indent=1)
    if res['pdl'] and res['serial']:
        for i, row in enumerate(res['pdl'][0]):
            def best(arm, key):
                vals = [run[i].get(key) for run in res[arm] if run[i].get(key) is not None]
                return min(vals) if vals else float('nan')
            print('%-5s n=%-8d K=%-3d %-24s flushed %.4f -> %.4f ms   back to back %.4f -> %.4f ms' % (
                row['what'], row['n'], row['K'], row['kernel'], best('serial', 'ms_flushed'), best('pdl', 'ms_flushed'),
                best('serial', 'ms_back_to_back'), best('pdl', 'ms_back_to_back')), flush=True)


if __name__ == '__main__':
    if '--worker' in sys.argv:
        worker()
    else:
        main()
