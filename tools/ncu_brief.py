#!/usr/bin/env python
"""Prints the handful of ncu metrics that decide the next step for a latency/issue-bound kernel,
and (with --src) the hottest SASS lines with their average active threads.
Usage: tools/ncu_brief.py REP [--src N] [--json OUT]"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']
out = {}
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h):
        v = data[0][i]
        try:
            if 'issue_stalled' in h and float(v) < 0.05: continue
        except ValueError: pass
        out[h] = f"{v} {units[i]}"
        print(f"{h:84s} {v} {units[i]}")
n = 0
if '--src' in sys.argv:
    n = int(sys.argv[sys.argv.index('--src') + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[hi[0]]
    ia, ie, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    ist = hdr.index("Warp Stall Sampling (All Samples)")
    d = [r for r in rows[hi[0] + 1:] if len(r) == len(hdr) and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in d); tots = sum(int(r[ist] or 0) for r in d)
    print("total warp inst", tot, "static", len(d), "samples", tots)
    for k, r in enumerate(d):
        e, t, s = int(r[ie]), int(r[it]), int(r[ist] or 0)
        if e > tot / n or s > tots / n:
            print(f"{k:5d} {e/1e6:8.1f}M thr {t/max(e,1):5.1f} stall {100*s/max(tots,1):5.2f}%  {r[ia][:100]}")
if '--json' in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index('--json') + 1], 'w'), indent=1)
