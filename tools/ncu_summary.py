#!/usr/bin/env python3
"""Summarise an Nsight Compute report of one training step into profiles/ artefacts.

    python tools/ncu_summary.py gpurun_out/prof_r01_step.ncu-rep profiles/r01_step

writes  <out>_kernels.csv   one row per launch: step-kernel name, grid, duration, DRAM bytes, DRAM %, tensor-pipe %
        <out>_traffic.json  {kernel name: dram_read+dram_write bytes per launch}  (bench.py reads profiles/traffic.json)

The launches of one training step always come in the order of fnd_train_step (csrc/fnd_api.cu); the n-th
fnd_gemm_kernel launch after prep_kernel is named accordingly.
"""
import csv
import io
import json
import subprocess
import sys

STEP_ORDER = ["prep", "gemm_proj", "gemm_qkv", "assemble_fwd", "gemm_fuse0", "gemm_fuse1", "gemm_pre0", "gemm_pre1", "head",
              "dgrad_pre1", "dgrad_pre0", "dgrad_fuse1", "dgrad_fuse0", "assemble_bwd", "dgrad_qkv", "wgrad_all", "adamw"]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for h, i in list(col.items()):            # some metrics carry a section prefix ("TPC.TriageCompute.<metric>")
        col.setdefault(h.split(".TriageCompute.")[-1], i)

    def val(r, name):
        if name not in col or r[col[name]] in ("", "n/a", "no data"):
            return None
        return float(r[col[name]].replace(",", "")) * UNIT_SCALE.get(units[col[name]], 1.0)

    # name the launches: find the first prep_kernel, then follow the step order
    names = [None] * len(data)
    kn = [r[col["Kernel Name"]] for r in data]
    pos = None
    for i, k in enumerate(kn):
        if k.startswith("prep_kernel"):
            pos = 0
        if pos is not None:
            names[i] = STEP_ORDER[pos % len(STEP_ORDER)]
            pos += 1
    # launches before the first prep: the tail of the previous step
    first = next((i for i, n in enumerate(names) if n), len(data))
    for i in range(first):
        names[i] = STEP_ORDER[len(STEP_ORDER) - first + i] if first <= len(STEP_ORDER) else "?"

    traffic, out_rows = {}, []
    for r, n in zip(data, names):
        rd, wr = val(r, "dram__bytes_read.sum") or 0.0, val(r, "dram__bytes_write.sum") or 0.0
        traffic.setdefault(n, []).append(rd + wr)
        out_rows.append([n, r[col["Kernel Name"]].split("(")[0], r[col["Grid Size"]], r[col["Block Size"]],
                         val(r, "gpu__time_duration.sum"), rd, wr,
                         val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         val(r, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
                         val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                         val(r, "launch__registers_per_thread"), val(r, "lts__t_bytes.sum")])
    with open(out + "_kernels.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["step_kernel", "cuda_kernel", "grid", "block", "duration_us", "dram_read_B", "dram_write_B", "dram_pct_peak",
                    "tensor_pipe_pct", "warps_active_pct", "regs", "l2_bytes"])
        w.writerows(out_rows)
    with open(out + "_traffic.json", "w") as f:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, f, indent=1, sort_keys=True)
    tot = sum(r[4] for r in out_rows if r[4])
    print(f"{len(out_rows)} launches, {tot:.1f} us under ncu (cold cache, serialised)")
    for r in out_rows:
        print(f"  {r[0]:14s} {r[2]:>12s} {r[4]:8.2f} us  share {100 * r[4] / tot:5.1f}%  dram {(r[5] + r[6]) / 1e6:8.2f} MB")


if __name__ == "__main__":
    main()
