#!/usr/bin/env python3
"""Prints the handful of ncu metrics we track from a .ncu-rep (raw page)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-85s %-12s %s" % (k, units[i], " | ".join(r[i] for r in data)))

# --json OUT KERNEL_SUBSTR BLOCKS K : the per-launch means bench.py's roofline object reads (profiles/ncu_kmap16.json)
if "--json" in sys.argv:
    import json
    a = sys.argv.index("--json")
    out, sub, blocks, K = sys.argv[a + 1], sys.argv[a + 2], int(sys.argv[a + 3]), int(sys.argv[a + 4])
    col = {k: hdr.index(k) for k in hdr}
    sel = [r for r in data if sub in r[col["Kernel Name"]]]
    # the retry launches that follow every k_map16 launch (a few microseconds, nothing to do unless a tracked pass failed) are
    # the same kernel: keep the launches that did a pass's worth of work
    def _inst(r):
        return float(r[col["smsp__inst_executed.sum"]].replace(",", ""))
    top = max(_inst(r) for r in sel)
    sel = [r for r in sel if _inst(r) >= 0.5 * top]

    def val(r, k):                                     # value in base units (ncu prints Gbyte / Mbyte / Kbyte per column)
        v = float(r[col[k]].replace(",", ""))
        u = units[col[k]].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    def mean(k):
        return sum(val(r, k) for r in sel) / len(sel)
    js = {"kernel": sub, "launches_captured": len(sel), "blocks_per_launch": blocks, "K": K,
          "dram_bytes_per_launch": mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum"),
          "dram_bytes_read_per_launch": mean("dram__bytes_read.sum"), "dram_bytes_write_per_launch": mean("dram__bytes_write.sum"),
          "warp_instructions_per_launch": mean("smsp__inst_executed.sum"),
          "duration_us_under_ncu": mean("gpu__time_duration.sum"),
          "utilisation_pct": {"issue_slots": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                              "alu_pipe": mean("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                              "fma_pipe": mean("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                              "dram_of_ncu_peak": mean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                              "warps_active": mean("sm__warps_active.avg.pct_of_peak_sustained_active")},
          "source": "ncu --set full --clock-control none capture %s (tools/ncu_summary.py --json)" % rep}
    json.dump(js, open(out, "w"), indent=1)
    print("wrote", out)
