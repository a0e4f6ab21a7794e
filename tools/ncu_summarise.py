#!/usr/bin/env python
"""tools/ncu_summarise.py REP [REP ...]: one line of JSON per captured launch with the metrics bench.py and DESIGN.md
quote (duration, warp instructions, issue-slot utilisation, threads per instruction, cache hit rates, DRAM bytes, the
L1 data pipe, registers), straight from `ncu -i REP --page raw --csv`."""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "time",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_utilization_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "sm__cycles_active.avg": "sm_cycles_active",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
         "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}

for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        out = {"report": rep.split("/")[-1], "kernel": r[hdr.index("Kernel Name")].replace("void ", "").split("(")[0].replace("unnamed>::", "")}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] != "":
                v = float(r[i].replace(",", ""))
                u = units[i]
                if WANT[h].startswith("dram_bytes") or WANT[h] == "time":
                    v *= SCALE.get(u, 1.0)
                out[WANT[h] + ("_s" if WANT[h] == "time" else "")] = round(v, 9) if WANT[h] == "time" else round(v, 3)
        print(json.dumps(out))
