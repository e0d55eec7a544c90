#!/usr/bin/env python
"""gpurun_out/field_ncu.csv (the ncu pass of tools/gpu_r02_run6.sh over tools/field_ncu_target.py) -> profiles/field_traffic.json:
DRAM bytes per launch of the field kernels, read by bench.py into the `field` entries of its JSON line.

    python tools/make_field_traffic.py [csv] [tag]"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "field_ncu.csv")
tag = sys.argv[2] if len(sys.argv) > 2 else "r02_final"
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
d = collections.OrderedDict()
for r in rows:
    d.setdefault((r[0], r[4]), {})[r[12]] = float(r[14].replace(",", ""))
out = {"source": f"profiles/{tag}_field_ncu.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none; "
                 "tools/field_ncu_target.py: 4096^2 x 6 f64 field, 16384 / 2000 stamps)"}
for (_, k), v in d.items():
    n = None
    if "extract_bulk_kernel<double, double>" in k:
        n = "extract_f64"
    elif "extract_bulk_kernel<double, float>" in k:
        n = "extract_f64_to_f32"
    elif "sqdiff_partial_kernel<double>" in k:
        n = "field_mse"
    elif "window_axpy_rows_kernel<double, float" in k:
        n = "window_axpy_f64_inplace" if ("(bool)1" in k or ", 1>" in k) else "window_axpy_f64"
    if n and n not in out and "dram__bytes_read.sum" in v:
        out[n] = {"kernel": k.split("(const")[0].replace("void ", "").strip(), "dram_bytes_per_launch": int(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]),
                  "dram_read": int(v["dram__bytes_read.sum"]), "dram_write": int(v["dram__bytes_write.sum"]), "ncu_us": v["gpu__time_duration.sum"] / 1e3}
json.dump(out, open(os.path.join(ROOT, "profiles", "field_traffic.json"), "w"), indent=1)
import shutil

shutil.copy(src, os.path.join(ROOT, "profiles", f"{tag}_field_ncu.csv"))
print({k: (v["dram_bytes_per_launch"], round(v["ncu_us"], 1)) for k, v in out.items() if k != "source"})
