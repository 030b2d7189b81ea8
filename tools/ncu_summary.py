#!/usr/bin/env python
"""Summaries of ncu captures for profiles/: launch-list shares from a --metrics gpu__time_duration.sum CSV, and a metric table
from a --set full report (ncu -i REP --page raw --csv).
    python tools/ncu_summary.py launches gpurun_out/x_launches.csv
    python tools/ncu_summary.py full gpurun_out/x_full.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or not r[iv].replace('.', '').replace(',', '').isdigit():
            continue
        n = r[ik].split('(')[0]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(',', ''))
    tot = sum(a[1] for a in agg.values())
    print('| kernel | launches | avg us | share |\n|---|---|---|---|')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| {n[:90]} | {c} | {t / c / 1e3:.1f} | {100 * t / tot:.1f}% |')


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index('Kernel Name')
    kernels = rows[2:]
    print('| metric | unit | ' + ' | '.join(r[ik].split('(')[0][:40] for r in kernels) + ' |')
    print('|---|---|' + '---|' * len(kernels))
    for m in WANT:
        if m in hdr:
            i = hdr.index(m)
            print(f'| {m} | {units[i]} | ' + ' | '.join(r[i] for r in kernels) + ' |')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
