#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel in an .ncu-rep
(captured with --import-source on; the library built with -lineinfo).
   python tools/ncu_lines.py <rep> <cubin> <mangled-name-substring> <source.cu> [top]
Aligns `ncu --page source --csv` (SASS order) with `nvdisasm -g -c` line markers."""
import csv, io, re, subprocess, sys

rep, cubin, name, src = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and name in l)
end = next((i for i in range(start + 1, len(sass)) if sass[i].startswith(".text.")), len(sass))
cur, seq = None, []
base = src.split("/")[-1]
for l in sass[start:end]:
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(base) else -1
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
# <rep> may also be the CSV already exported on the GPU box (`ncu -i x.ncu-rep --page source --csv`)
out = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h0 = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr, data = rows[h0], rows[h0 + 1:]
reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h]
ia, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
assert len(seq) == len(data), (len(seq), len(data))
agg, st, why = {}, {}, {}
for (ln, _), r in zip(seq, data):
    agg[ln] = agg.get(ln, 0) + int(r[ia])
    st[ln] = st.get(ln, 0) + int(r[ist])
    w = why.setdefault(ln, {})
    for i in reasons:
        if r[i] not in ("", "0"):
            w[hdr[i]] = w.get(hdr[i], 0) + int(float(r[i]))
tot, tst = sum(agg.values()), sum(st.values())
lines = open(src).read().split("\n")
print(f"instructions {tot}, stall samples {tst}")
for ln, c in sorted(st.items(), key=lambda x: -x[1])[:top]:
    text = lines[ln - 1].strip()[:100] if ln and ln > 0 else "(other file)"
    top3 = " ".join(f"{k[6:]}={v}" for k, v in sorted(why.get(ln, {}).items(), key=lambda x: -x[1])[:3])
    print(f"stall {c / tst:6.1%}  inst {agg[ln] / tot:6.1%}  L{ln}: {text}  [{top3}]")
