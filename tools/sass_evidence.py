#!/usr/bin/env python
"""Instruction evidence per kernel from the built library (no GPU needed):
   python tools/sass_evidence.py > profiles/r2_sass_evidence.txt
Counts the SASS mnemonics that prove which hardware path a kernel uses: DMMA (FP64 tensor cores,
mma.sync.m8n8k4.f64), HMMA...TF32 (mma.sync.m16n8k8 tf32), UTMALDG / UTMASTG (TMA), LDGSTS
(cp.async), 256-bit global stores, SYNCS (mbarrier), cluster barriers."""
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
lib = os.path.join(ROOT, "photon_weave_b200", "lib", "libpwb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = {"DMMA": r"\bDMMA", "TF32_MMA": r"HMMA\.[0-9]+\.F32\.TF32", "TMA": r"\bUTMA(LDG|STG)", "cp.async": r"\bLDGSTS",
        "STG.256": r"\bSTG\.E\.(ENL2\.)?256|\bSTG\.E\.STRONG\.\w+\.256|STG\S*\.256", "LDG.128": r"\bLDG\S*\.128",
        "mbarrier": r"\bSYNCS", "cluster": r"\bUCGABAR|\bBAR\.ARV\.\w*CLUSTER|CGABAR"}
print("# SASS evidence (cuobjdump -sass photon_weave_b200/lib/libpwb200.so, sm_100a): instruction counts per kernel")
print("# DMMA = mma.sync.m8n8k4.f64 (FP64 tensor cores); TF32_MMA = mma.sync.m16n8k8 tf32 (complex64 3xTF32 kernels);")
print("# TMA = cp.async.bulk.tensor; cp.async = LDGSTS; STG.256 = 256-bit global stores; mbarrier = SYNCS")
name, counts, rows = None, {}, []
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, counts))
        name, counts = m.group(1), {k: 0 for k in pats}
        continue
    if name:
        for k, p in pats.items():
            if re.search(p, line):
                counts[k] += 1
if name:
    rows.append((name, counts))
demangled = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
for (raw, c), dn in zip(rows, demangled):
    if not any(c[k] for k in ("DMMA", "TF32_MMA", "TMA", "cp.async", "STG.256", "mbarrier", "cluster")):
        continue
    short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("pw::", "").replace("(anonymous namespace)::", "")
    print(f"{short[:78]:78s} " + "  ".join(f"{k}={v}" for k, v in c.items() if v))
