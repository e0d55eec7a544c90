#!/bin/bash
# round 2, GPU call 5 (2 GPUs): tests, tiled field over NCCL, 2-GPU bench line
O=gpurun_out/r02e; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 6 $O/tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/field_tiled_nccl.py 4096 2000 > $O/field_tiled_2gpu.json 2> $O/field_tiled_2gpu.err; echo "tiled rc=$?"; tail -n 2 $O/field_tiled_2gpu.json | cut -c1-900
timeout 600 python tools/field_tiled_nccl.py 4096 2000 > $O/field_tiled_1gpu.json 2> $O/field_tiled_1gpu.err; echo "tiled 1gpu rc=$?"; tail -n 1 $O/field_tiled_1gpu.json | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02e/bench_2gpu.json').read().strip().splitlines()[-1])
print("2gpu value",round(b['value']),"e2e",round(b['e2e']['value']), "f64", b['e2e'].get('pageable_f64_input',{}).get('value'))
print('field_tiled', {kk:vv for kk,vv in (b.get('field_tiled') or {}).items() if kk not in ('api','collectives','timing')})
PY
