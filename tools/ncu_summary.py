#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into small committed summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_k8.md "<cmd>"
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_clike_tile_k8.md "<cmd>"
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
    'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed',
    'sm__ops_path_tensor_src_int8.avg.pct_of_peak_sustained_elapsed',
    'sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith('=='))]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    d = OrderedDict()
    for r in rows[1:]:
        if len(r) > vi:
            d.setdefault(r[ki], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in d.values())
    with open(dst, 'w') as f:
        f.write('# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n')
        f.write('command: `%s`\n\nPer-launch times are cold-cache and serialised: compare SHARES.\n\n' % cmd)
        f.write('| kernel | launches | mean ns | share of GPU time |\n|---|---|---|---|\n')
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write('| `%s` | %d | %.0f | %.3f |\n' % (k.split('(')[0][:90], len(v), sum(v) / len(v), sum(v) / tot))


def full(src, dst, cmd):
    out = subprocess.check_output(['ncu', '-i', src, '--page', 'raw', '--csv']).decode()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write('# ncu --set full (clock-control none) summary\n\ncommand: `%s`\n\n' % cmd)
        name_i = hdr.index('Kernel Name')
        for r in rows[2:]:
            f.write('## %s\n\n| metric | value | unit |\n|---|---|---|\n' % r[name_i].split('(')[0])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write('| %s | %s | %s |\n' % (k, r[i], units[i]))
            f.write('\nissue-stall samples (smsp__pcsamp_warps_issue_stalled_*):\n\n')
            st = []
            for i, h in enumerate(hdr):
                if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
                    try:
                        st.append((float(r[i].replace(',', '')), h.split('issue_stalled_')[1]))
                    except ValueError:
                        pass
            tot = sum(v for v, _ in st) or 1
            for v, h in sorted(st, reverse=True)[:8]:
                f.write('* %s: %.0f (%.1f %%)\n' % (h, v, 100 * v / tot))
            f.write('\n')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
