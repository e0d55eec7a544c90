#!/usr/bin/env python
"""Top stall-sample instructions of one kernel of an .ncu-rep (needs -lineinfo / --import-source on).
usage: ncu_src.py report.ncu-rep <kernel regex> [launch-skip] [topN]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:2])
hdr = rows[1]
src, si, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and not c.endswith("_not_issued")]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) > si and r[0].startswith("0x"):
        data.append(r)
tot = sum(float(r[si] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
for k, r in sorted(enumerate(data), key=lambda kr: -float(kr[1][si] or 0))[:topn]:
    st = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {float(r[si]):7.0f} {100*float(r[si])/tot:5.1f}% exec={r[ie]:>8}  {r[src].strip()[:70]:70s} {' '.join(f'{n[6:]}={v:.0f}' for v,n in st if v>0)}")
