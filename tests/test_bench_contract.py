"""The reference arm of bench.py (`--impl reference`, the oracle port on the host cores) runs without a GPU: check the JSON
line the driver parses — one line, the contract keys, the arm's own cpu_baseline / e2e objects."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, DBV_REF_SAMPLE="32", OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "stamps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", DBV_REF_SAMPLE="32")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_recorded_gpu_line_and_ncu_launch_list_agree():
    """The committed evidence of the GPU arm (profiles/r02_final_bench.json, written by bench.py on the B200) carries the contract keys,
    and the dominant kernel's share of the step agrees with the ncu launch list of the same command (profiles/r02_final_launches.csv:
    per-launch times are cold-cache and serialised, so shares are compared, not times)."""
    import collections
    import csv

    line = json.load(open(os.path.join(ROOT, "profiles", "r02_final_bench.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    r = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] < line["value"]
    assert line["gpu_launches"] > 0 and line["cpu_baseline"]["kind"] == "port" and "workload" in line["config"]
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    rows = [x for x in csv.reader(open(os.path.join(ROOT, "profiles", "r02_final_launches.csv"))) if x and not x[0].startswith("==")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = collections.Counter()
    for x in rows[1:]:
        tot[x[ki].split("(")[0].replace("void ", "")] += float(x[vi].replace(",", ""))
    name = r["kernel"].split(" ")[0]  # e.g. tc_pairh_kernel<128>
    share = tot[name] / sum(tot.values())
    assert share == max(tot.values()) / sum(tot.values()), "the roofline's kernel is the one with the largest share in the ncu list"
    assert abs(share - r["share_of_step"]) < 0.02, (share, r["share_of_step"])
