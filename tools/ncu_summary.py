#!/usr/bin/env python
"""Summarise .ncu-rep captures into a small CSV under profiles/ (run here, no GPU needed).
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/r1_x.csv
"""
import csv
import io
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__cluster_size',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct']

w = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--print-units', 'base'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    idx = [h.index(k) for k in WANT if k in h]
    if first:
        w.writerow(['report'] + ['%s [%s]' % (h[i], units[i]) if units[i] else h[i] for i in idx])
        first = False
    for r in rows[2:]:
        w.writerow([rep.split('/')[-1]] + [r[i] for i in idx])
