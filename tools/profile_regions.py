#!/usr/bin/env python
"""Reads an .ncu-rep (raw + source pages) and prints key metrics and a per-region
instruction accounting of pathtrace_accumulate_kernel. Usage: tools_profile_regions.py REP"""
import collections, csv, io, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
summary = {}
for i, h in enumerate(hdr):
    if h in want:
        summary[h] = {"unit": units[i], "per_launch": [r[i] for r in data]}
        print(f"{h:82s} {units[i]:10s} {[r[i] for r in data]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[hi[0]]
ia, ie, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
end = hi[1] - 1 if len(hi) > 1 else len(rows)
d = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in d)
votes = [k for k, r in enumerate(d) if 'VOTE' in r[ia]]
iters = int(d[votes[0]][ie])
lds = [k for k, r in enumerate(d) if 'LDS' in r[ia]]
print("total warp inst", tot, "loop iterations", iters, "per iteration", round(tot / iters, 1))
regions = {}
def region(a, b, name):
    e = sum(int(r[ie]) for r in d[a:b]); t = sum(int(r[it]) for r in d[a:b])
    regions[name] = {"static": b - a, "warp_inst_per_iter": round(e / iters, 1), "share_pct": round(e / tot * 100, 1), "avg_threads": round(t / max(e, 1), 1)}
    print(f"{name:16s} static {b-a:5d} warp-inst/iter {e/iters:7.1f} share {e/tot*100:5.1f}% avg thr {t/max(e,1):5.1f}")
# the scan ends where the first shading LDS begins (minus the finalize prologue)
first_scan = votes[1]
region(votes[0], votes[1], "regeneration")
region(first_scan, lds[0] - 8, "scan")
region(lds[0] - 8, len(d), "finalize+shade")
op = collections.Counter()
for r in d[first_scan:lds[0] - 8]:
    s = r[ia].strip().split(); m = s[0] if not s[0].startswith('@') else s[1]; op[m.split('.')[0]] += int(r[ie])
print("scan mix per iter:", {k: round(v / iters, 1) for k, v in op.most_common(16)})
if len(sys.argv) > 2:
    json.dump({"metrics": summary, "loop_iterations": iters, "warp_inst_per_iteration": round(tot / iters, 1), "regions": regions,
               "scan_mix_per_iteration": {k: round(v / iters, 1) for k, v in op.most_common(16)}}, open(sys.argv[2], "w"), indent=1)
