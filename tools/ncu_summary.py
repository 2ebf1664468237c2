#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few
per-launch numbers bench.py / DESIGN.md quote.  Usage:
   python tools/ncu_summary.py <rep> <key> [<rep> <key> ...]  -> prints JSON
"""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:80]}
        for k, name in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                v = float(r[i].replace(",", ""))
                d[name] = v * UNIT.get(units[i], 1)
        if "dram_read" in d:
            d["dram_bytes_per_launch"] = d["dram_read"] + d["dram_write"]
        launches.append(d)
    return launches


if __name__ == "__main__":
    res = {}
    args = sys.argv[1:]
    for rep, key in zip(args[0::2], args[1::2]):
        ls = summarize(rep)
        res[key] = dict(ls[-1], launches_captured=len(ls))
    print(json.dumps(res, indent=1))
