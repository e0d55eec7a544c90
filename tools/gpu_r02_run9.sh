#!/bin/bash
# round 2, GPU call 9 (8 GPUs): tiled field over NCCL at 8 ranks, 8-GPU bench line (stamps/s + the all-ranks field_tiled extra)
O=gpurun_out/r02m; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/field_tiled_nccl.py 4096 2000 > $O/field_tiled_8gpu.json 2> $O/field_tiled_8gpu.err; echo "tiled rc=$?"; tail -n 1 $O/field_tiled_8gpu.json | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02m/bench_8gpu.json').read().strip().splitlines()[-1])
print("8gpu value",round(b['value']),"e2e",round(b['e2e']['value']), "f64", b['e2e'].get('pageable_f64_input',{}).get('value'))
print('field_tiled', {kk:vv for kk,vv in (b.get('field_tiled') or {}).items() if kk not in ('api','collectives','timing')})
PY
