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


def by_kernel(rep, min_ms=0.5):
    """Per kernel name: launches captured and the AVERAGE of every metric over them (only
    launches longer than min_ms: the analysis / skipped-fallback launches are not the point)."""
    import re

    agg = {}
    for d in summarize(rep):
        if d.get("duration", 0) * 1e3 < min_ms:
            continue
        name = re.sub(r"\(.*", "", d["kernel"]).replace("void ", "").strip()
        agg.setdefault(name, []).append(d)
    out = {}
    for name, ls in agg.items():
        keys = [k for k in ls[0] if k != "kernel"]
        out[name] = {k: sum(x.get(k, 0.0) for x in ls) / len(ls) for k in keys}
        out[name]["launches_captured"] = len(ls)
        out[name]["duration_ms_each"] = [round(x["duration"] * 1e3, 3) for x in ls]
    return out


if __name__ == "__main__":
    if sys.argv[1] == "--by-kernel":
        print(json.dumps(by_kernel(sys.argv[2]), indent=1))
        sys.exit(0)
    res = {}
    args = sys.argv[1:]
    for rep, key in zip(args[0::2], args[1::2]):
        ls = summarize(rep)
        res[key] = dict(ls[-1], launches_captured=len(ls))
    print(json.dumps(res, indent=1))
