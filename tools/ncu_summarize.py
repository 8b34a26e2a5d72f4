#!/usr/bin/env python
"""Condense `ncu --set full` reports (gpurun_out/*.ncu-rep, scratch) into the tracked summary profiles/r02_ncu_summary.json:
per kernel the duration, DRAM bytes, pipe utilisation, L2 hit rate, registers, grid -- the numbers DESIGN.md quotes.

    python tools/ncu_summarize.py gpurun_out/r02_ncu_*.ncu-rep > profiles/r02_ncu_summary.json
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "dmma_pipe_pct_of_active_cycles",
    "sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed": "fp64_tensor_ops_pct_of_peak_over_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_of_active_cycles",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__occupancy_limit_registers": "occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def main():
    out = {}
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            rec = {}
            for h, u, v in zip(hdr, units, vals):
                if h == "Kernel Name":
                    rec["kernel"] = v
                elif h in KEEP:
                    try:
                        x = float(v.replace(",", ""))
                    except ValueError:
                        continue
                    if u in SCALE and KEEP[h] in ("duration", "dram_read", "dram_write"):
                        x *= SCALE[u]
                    rec[KEEP[h]] = x
            if "duration" in rec and rec.get("dram_read") is not None:
                rec["dram_bytes"] = rec["dram_read"] + rec.get("dram_write", 0.0)
                rec["dram_GBps"] = rec["dram_bytes"] / rec["duration"] / 1e9
            out[os.path.basename(path).replace(".ncu-rep", "")] = rec
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
