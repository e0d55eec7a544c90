#!/usr/bin/env python
"""Turn gpurun_out/{bench.json, launches.csv, prof_full_raw.csv, bench_ref.json} into the tracked evidence under profiles/:
   <tag>_bench.json, <tag>_bench_ref.json, <tag>_launches.csv, <tag>_full_raw_summary.csv, traffic.json, <tag>_summary.md
usage: tools/make_profile_summary.py r01_final"""
import csv, json, os, sys, shutil, collections

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
bench = json.loads(open(os.path.join(G, "bench.json")).read().strip().splitlines()[-1])
json.dump(bench, open(os.path.join(P, f"{tag}_bench.json"), "w"), indent=1)
try:
    ref = json.loads(open(os.path.join(G, "bench_ref.json")).read().strip().splitlines()[-1])
    json.dump(ref, open(os.path.join(P, f"{tag}_bench_ref.json"), "w"), indent=1)
except Exception as e:
    ref = {"error": repr(e)}
# ---- launch list (timed region of `bench.py --steps 2`) ------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
tot = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("void ", "")
    t, n = tot.get(k, (0.0, 0))
    tot[k] = (t + float(r[vi].replace(",", "")) / 1e6, n + 1)  # ns -> ms
all_ms = sum(t for t, _ in tot.values())
# ---- full capture: per-kernel metrics of one chunk --------------------------------------------------------
raw = list(csv.reader(open(os.path.join(G, "prof_full_raw.csv"))))
h = raw[0]
def col(r, n):
    return r[h.index(n)] if n in h else ""
names = ["enc_bn_pack", "enc_conv1", "enc_conv2", "enc_conv3", "enc_conv4", "enc_conv5", "enc_conv6", "enc_conv7", "enc_conv8", "enc_dense", "latent",
         "dec_dense1", "dec_dense2", "dec_convT1", "dec_convT2", "dec_convT3", "dec_convT4", "dec_convT5", "dec_convT6", "dec_convT7", "dec_convT8", "dec_head"]
keep = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic"]
body = raw[2:2 + len(names)]
prec = bench["config"]["precision"]
traffic = {"source": f"profiles/{tag}_full_raw_summary.csv (ncu --set full --clock-control none, one launch of 1024 stamps, precision {prec})", prec: {}}
with open(os.path.join(P, f"{tag}_full_raw_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["layer", "kernel"] + keep)
    for nm, r in zip(names, body):
        kn = col(r, "Kernel Name").split("(")[0].replace("void ", "")
        w.writerow([nm, kn] + [col(r, k) for k in keep])
        try:
            def tobytes(v, unit):
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            u = raw[1]
            db = tobytes(col(r, "dram__bytes_read.sum"), u[h.index("dram__bytes_read.sum")]) + tobytes(col(r, "dram__bytes_write.sum"), u[h.index("dram__bytes_write.sum")])
            traffic[prec][nm] = {"dram_bytes": int(db), "stamps": 1024, "kernel": kn}
        except Exception:
            pass
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
# ---- summary ----------------------------------------------------------------------------------------------
L = bench["layers"]
ev_ms = sum(l["ms"] for l in L)
with open(os.path.join(P, f"{tag}_summary.md"), "w") as f:
    f.write(f"# {tag} — B200, precision {bench['config']['precision']}, {bench['config']['stamps_per_gpu_per_step']} stamps/step\n\n")
    f.write(f"bench.py (no profiler): value = {bench['value']:.0f} stamps/s, e2e = {bench['e2e']['value']:.0f} stamps/s, {bench['ms_per_step']:.3f} ms/step, "
            f"launches/step = {bench['gpu_launches'] // bench['steps']}, clocks {bench['clocks']}\n\n")
    if "value" in ref:
        f.write(f"reference arm (`bench.py --impl reference`, oracle port on {ref['cpu_baseline']['cores']} host threads): {ref['value']:.0f} stamps/s\n\n")
    if "cpu_baseline" in bench:
        f.write(f"cpu_baseline in the same run: {bench['cpu_baseline']}\n\n")
    f.write("## ncu launch list of the timed region (`--metrics gpu__time_duration.sum --clock-control none`, 2 steps)\n\nPer-launch times are cold-cache / serialised: compare SHARES.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        f.write(f"| `{k}` | {n} | {t:.3f} | {100 * t / all_ms:.1f}% |\n")
    f.write("\n## CUDA-event per-layer times averaged over the timed steps (ms per step) and algorithmic TFLOP/s\n\n| layer | ms | share | TFLOP/s | frac of sustained bf16 peak |\n|---|---|---|---|---|\n")
    for l in L:
        f.write(f"| {l['layer']} | {l['ms']:.3f} | {100 * l['ms'] / ev_ms:.1f}% | {l['tflops']} | {l['frac']} |\n")
    f.write(f"\nroofline object: {json.dumps(bench['roofline'])}\n\n")
    if "field" in bench:
        f.write(f"field kernels: {json.dumps(bench['field'])}\n\n")
    if "alt_precision" in bench:
        f.write(f"other precisions: {json.dumps(bench['alt_precision'])}\n\n")
    f.write("## one chunk under `ncu --set full` (see the _full_raw_summary.csv next to this file)\n\n| layer | kernel | us | tensor pipe active % | TC smem wavefronts % | LSU wavefronts % | L2 % | DRAM % | DRAM MB |\n|---|---|---|---|---|---|---|---|---|\n")
    for nm, r in zip(names, body):
        kn = col(r, "Kernel Name").split("(")[0].replace("void ", "")
        db = traffic[prec].get(nm, {}).get("dram_bytes", 0) / 1e6
        f.write(f"| {nm} | `{kn}` | {col(r, keep[0])} | {col(r, keep[2])[:5]} | {col(r, keep[3])[:5]} | {col(r, keep[4])[:5]} | {col(r, keep[5])[:5]} | {col(r, keep[6])[:5]} | {db:.0f} |\n")
    # ---- the device detector: per-kernel times / DRAM bytes of one 4096^2 detection (tools/detect_ncu_target.py under ncu)
    dn = os.path.join(G, "detect_ncu.csv")
    if os.path.exists(dn):
        shutil.copy(dn, os.path.join(P, f"{tag}_detect_ncu.csv"))
        rows = [r for r in csv.reader(open(dn)) if len(r) > 10 and r[0].isdigit()]
        d = collections.OrderedDict()
        for r in rows:
            d.setdefault((int(r[0]), r[4].split("(")[0].replace("void ", "")), {})[r[12]] = float(r[14].replace(",", ""))
        f.write("\n## device detector: one detection of a 4096^2 x 6 f64 field with 2000 sources (`ncu --metrics gpu__time_duration.sum,dram__bytes_*`; "
                "16 launches = one call; CUDA-event time of the call without a profiler: see `field.detect.ms` above)\n\n"
                "| kernel | us | DRAM read MB | DRAM write MB |\n|---|---|---|---|\n")
        tot_us = 0.0
        # the capture window need not start at a call boundary: one row per kernel (first occurrence), in pipeline order
        order = ["det_band_kernel", "det_mesh_kernel", "det_mesh_fill_kernel", "det_mesh_median_kernel", "det_mesh_rank_kernel", "det_mesh_final_kernel",
                 "det_nodes_kernel", "det_foreground_kernel", "det_filter", "det_ccl_merge_kernel", "det_ccl_flatten_stats_kernel", "det_mark_kernel",
                 "det_count_kernel", "det_scan_kernel", "det_scatter_kernel", "det_moments_kernel"]
        first = collections.OrderedDict()
        for (i, k), v in d.items():
            first.setdefault(k, v)
        d = collections.OrderedDict(((j, k), v) for j, (k, v) in enumerate(sorted(first.items(), key=lambda kv: next((n for n, o in enumerate(order) if kv[0].startswith(o)), 99))))
        for (i, k), v in d.items():
            us = v.get("gpu__time_duration.sum", 0.0) / 1e3
            tot_us += us
            f.write(f"| `{k}` | {us:.1f} | {v.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {v.get('dram__bytes_write.sum', 0) / 1e6:.1f} |\n")
        f.write(f"| sum | {tot_us:.1f} | | |\n")
print("wrote profiles/%s_*" % tag)
