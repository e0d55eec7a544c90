#!/bin/bash
# round 2, GPU call 10 (8 GPUs): aggregate host<->device copy ceiling of the box (8 ranks copying at once), tiled field pass
# with the prepared exchange plan + its phase breakdown, 8-GPU bench line
O=gpurun_out/r02n; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/pcie_probe.py > $O/pcie_probe_8gpu.log 2> $O/pcie_probe_8gpu.err; echo "pcie rc=$?"; tail -n 1 $O/pcie_probe_8gpu.log
timeout 300 python tools/pcie_probe.py > $O/pcie_probe_1gpu.log 2>/dev/null; tail -n 1 $O/pcie_probe_1gpu.log
timeout 600 $TR --master-port 29542 tools/field_tiled_nccl.py 4096 2000 > $O/field_tiled_8gpu.json 2> $O/field_tiled_8gpu.err; echo "tiled rc=$?"; tail -n 1 $O/field_tiled_8gpu.json | cut -c1-700
timeout 600 $TR --master-port 29543 tools/field_tiled_breakdown.py 4096 2000 > $O/field_tiled_breakdown_8gpu.json 2> /dev/null; echo "breakdown rc=$?"; tail -n 1 $O/field_tiled_breakdown_8gpu.json | cut -c1-900
timeout 900 $TR --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02n/bench_8gpu.json').read().strip().splitlines()[-1])
print("8gpu value",round(b['value']),"e2e",round(b['e2e']['value']), "f64", b['e2e'].get('pageable_f64_input',{}).get('value'))
print('field_tiled', {kk:vv for kk,vv in (b.get('field_tiled') or {}).items() if kk not in ('api','collectives','timing')})
PY
