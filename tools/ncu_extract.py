#!/usr/bin/env python
"""Summarise an Nsight Compute report for profiles/: per kernel (first profiled launch of each name)
duration, DRAM bytes, executed fp64 instruction counts (-> executed FLOP/s), pipe / issue utilisation,
occupancy and the top warp-stall reasons.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_extract.py /tmp/raw.csv "command that was profiled" > profiles/r02_ncu_summary.json
"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma",
    "sm__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul",
    "sm__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_fp64.sum": "warp_instructions_fp64_pipe",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__pcsamp_sample_buffer_full": None,
}
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|"
                   r"smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"command": sys.argv[2] if len(sys.argv) > 2 else "", "kernels": {}}
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        short = re.sub(r"\(.*$", "", name).strip()
        short = re.sub(r"^void\s+", "", short)
        short = short.replace("c8::", "")
        if short in out["kernels"]:
            continue
        k = {"full_name": name[:200]}
        stalls = {}
        for h, i in col.items():
            v = num(r[i]) if i < len(r) else None
            if v is None:
                continue
            if h in WANT and WANT[h]:
                u = units[i]
                key = WANT[h]
                if key == "duration_ns":
                    v *= {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}.get(u, 1)
                if key.endswith("_bytes"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                k[key] = v
            m = STALL.match(h)
            if m:
                stalls[m.group(1) or m.group(2)] = v
        if "dram_read_bytes" in k:
            k["dram_bytes"] = k["dram_read_bytes"] + k.get("dram_write_bytes", 0.0)
        if "dfma" in k and "duration_ns" in k:
            fl = 2 * k["dfma"] + k.get("dmul", 0) + k.get("dadd", 0)
            k["executed_fp64_flops"] = fl
            k["executed_fp64_tflops"] = fl / k["duration_ns"] * 1e-3
        k["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out["kernels"][short] = k
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
