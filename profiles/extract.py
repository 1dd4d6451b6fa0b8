"""Summarise `ncu --set full` reports into one CSV row per kernel (run here, no GPU needed):
    python profiles/extract.py gpurun_out/r1_ncu_*.ncu-rep > profiles/r1_ncu_full_summary.csv"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct"]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "barrier", "math_pipe_throttle", "wait", "not_selected",
               "dispatch_stall", "lg_throttle", "mio_throttle", "membar", "branch_resolving"]

out = csv.writer(sys.stdout)
hdr = None
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        cols = ["report", "kernel"] + KEYS + ["stall_" + s for s in STALL_NAMES]
        if hdr is None:
            hdr = cols
            out.writerow(cols)
        vals = [rep.split("/")[-1], d.get("Kernel Name", "")]
        for k in KEYS:
            v = d.get(k, "")
            vals.append((v + " " + u.get(k, "")).strip() if v else "")
        for s in STALL_NAMES:
            vals.append(d.get(STALLS % s, ""))
        out.writerow(vals)
