#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
ONE bench.py step: launches, serialised time, share of the step, DRAM bytes; and the DRAM traffic of the Cholesky launch sequence,
which bench.py reports as `roofline.traffic`.

    python tools/launch_summary.py gpurun_out/r02_launches_c4.csv "<the ncu command>" [k] > profiles/r02_launches_c4_summary.json
(k: summarise the k-th eval of the list only -- bench.py --steps 1 --warmup 1 runs warm-up, timed, e2e and check evals; k = 2 is the timed one)
"""
import collections
import csv
import json
import re
import sys

CHOL = ("potrf_tile_kernel2", "gemm_tile_kernel_v2", "gemm_direct2_kernel", "chain_column_kernel", "ozaki_update_kernel", "ozaki_slice_kernel", "ozaki_scale_kernel")


def main():
    path, command = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per_id = collections.OrderedDict()
    for row in rd:
        if len(row) != len(hdr):
            continue
        rec = per_id.setdefault(row[ix["ID"]], {"name": row[ix["Kernel Name"]]})
        val = float(row[ix["Metric Value"]].replace(",", ""))
        unit = row[ix["Metric Unit"]]
        name = row[ix["Metric Name"]]
        if name == "gpu__time_duration.sum":
            rec["ms"] = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        elif name.startswith("dram__bytes"):
            rec["dram"] = rec.get("dram", 0.0) + val * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
    # one step = from the `occurrence`-th launch of the projection kernel (the first kernel of an eval) up to the next one
    recs = list(per_id.values())
    occurrence = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    if occurrence > 0:
        starts = [i for i, r in enumerate(recs) if "project_dmma_kernel" in r["name"]]
        lo = starts[occurrence - 1]
        hi = starts[occurrence] if occurrence < len(starts) else len(recs)
        recs = recs[lo:hi]
    per_kernel = collections.OrderedDict()
    for rec in recs:
        key = re.sub(r"\(.*$", "", rec["name"]).strip()
        k = per_kernel.setdefault(key, {"launches": 0, "total_ms": 0.0, "dram_bytes": 0.0})
        k["launches"] += 1
        k["total_ms"] += rec.get("ms", 0.0)
        k["dram_bytes"] += rec.get("dram", 0.0)
    total = sum(k["total_ms"] for k in per_kernel.values())
    for k in per_kernel.values():
        k["share_pct"] = round(100.0 * k["total_ms"] / total, 4) if total else None
        k["total_ms"] = round(k["total_ms"], 5)
    chol = [n for n in per_kernel if any(c in n for c in CHOL)]
    N, m = 16384, 64
    out = {
        "command": command, "round": 2,
        "config": "BASELINE config 4: OILMM p=64 m=64 N=16384, one full eval (the timed step), streams=1; final round-2 code",
        "per_kernel": per_kernel, "total_ms": round(total, 3),
        "cholesky_sequence": {
            "launches": sum(per_kernel[n]["launches"] for n in chol),
            "dram_bytes_per_step": sum(per_kernel[n]["dram_bytes"] for n in chol),
            "serialized_ms": round(sum(per_kernel[n]["total_ms"] for n in chol), 5),
            "share_pct_of_step": round(100.0 * sum(per_kernel[n]["total_ms"] for n in chol) / total, 3) if total else None,
            "algorithmic_flops": m * N ** 3 / 3.0, "kernels": chol},
    }
    out["dram_bytes_per_launch"] = out["cholesky_sequence"]["dram_bytes_per_step"]
    for key, alg, label in (("kmat_sym_kernel", m * 8.0 * N * (N + 1) / 2, "kmat"), ("fwd_sweep_kernel", m * 8.0 * N * (N + 1) / 2, "fwd_sweep"),
                            ("bwd_sweep_kernel", m * 8.0 * N * (N + 1) / 2, "bwd_sweep"), ("project_dmma_kernel", 8.0 * (64 + 64) * N, "project")):
        for n, k in per_kernel.items():
            if key in n and k["total_ms"] > 0:
                out[label] = {"ms": k["total_ms"], "dram_bytes": k["dram_bytes"], "algorithmic_bytes": alg,
                              "achieved_GBs_algorithmic": alg / (k["total_ms"] * 1e-3) / 1e9}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
