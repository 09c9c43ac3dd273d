#!/usr/bin/env python
"""Per-kernel one-line summary of an .ncu-rep (reads `ncu --page raw --csv`)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [('gpu__time_duration.sum', 'us'), ('smsp__inst_executed.sum', 'Minst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('smsp__thread_inst_executed_per_inst_executed.ratio', 'thr/inst'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'),
        ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'noinst'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long_sb'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'short_sb'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'wait'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'barrier'),
        ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'branch'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'math_thr'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'mio_thr'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'lg_thr'),
        ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'dispatch'),
        ('dram__bytes_read.sum', 'dramR'), ('dram__bytes_write.sum', 'dramW')]
print('%-34s' % 'kernel' + ''.join('%9s' % n for _, n in want))
tot = 0.0
for r in rows[2:]:
    name = r[col['Kernel Name']]
    name = name.replace('rmd_eval_kernel', 'eval').replace('(anonymous namespace)::', '')[:33]
    vals = []
    for k, n in want:
        v = r[col[k]] if k in col else ''
        try:
            f = float(v.replace(',', ''))
            if n == 'us':
                unit = rows[1][col[k]]
                f = f * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit, 1)
                tot += f
            if n == 'Minst':
                f /= 1e6
            if n in ('dramR', 'dramW'):
                unit = rows[1][col[k]]
                f = f * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(unit, 1)
            vals.append('%9.2f' % f)
        except ValueError:
            vals.append('%9s' % v[:8])
    print('%-34s' % name + ''.join(vals))
print('total us', tot)
