#!/usr/bin/env python3
"""Per-launch summary of `ncu --set full` captures of the sequence front-end kernels -> profiles/r02_seq_kernels.csv.

    python tools/ncu_seq_summary.py profiles/r02_seq_kernels.csv gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]

Columns: source report, kernel, grid, duration (us), tensor-pipe % (sm__pipe_tensor_cycles_active, of peak sustained active),
XU-pipe % (MUFU: sm__inst_executed_pipe_xu), DRAM read / write bytes, DRAM % of peak, SM clock (GHz), registers per thread."""
import csv
import io
import os
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "duration_us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("gpc__cycles_elapsed.avg.per_second", "sm_clock"), ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid")]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    rows_out = []
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        for r in data:
            rec = {"report": os.path.basename(rep), "kernel": r[col["Kernel Name"]].split("(")[0]}
            for m, name in COLS:
                if m in col:
                    rec[name] = r[col[m]] + (" " + units[col[m]] if units[col[m]] and name in ("dram_read", "dram_write", "sm_clock", "duration_us") else "")
            rows_out.append(rec)
    keys = ["report", "kernel"] + [n for _, n in COLS]
    with open(out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        for r in rows_out:
            w.writerow(r)
    print(f"{len(rows_out)} launches -> {out}")


if __name__ == "__main__":
    main()
