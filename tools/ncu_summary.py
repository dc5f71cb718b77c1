"""Summarise an ncu --page raw --csv export: time, DRAM bytes, issue/occupancy, FP64 pipe, top stall reasons per kernel.
Usage: ncu -i X.ncu-rep --page raw --csv > X.csv; python tools/ncu_summary.py X.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__average_warp_latency_per_inst_issued.ratio"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:90], "grid", r[idx.get("Grid Size", 0)], "block", r[idx.get("Block Size", 0)])
    for w in want:
        if w in idx and r[idx[w]] != "":
            print("   %-70s %s %s" % (w, r[idx[w]], units[idx[w]]))
    top = sorted(((float(r[idx[h]].replace(",", "") or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                  for h in stall), reverse=True)[:6]
    print("   stalls (warps per issue-active cycle):", ", ".join("%s %.2f" % (n, v) for v, n in top))
