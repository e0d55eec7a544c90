#!/bin/bash
# round 2, GPU call 13 (2 GPUs): the tiled detector again after its exact path moved onto the device (all-gather of the detection band),
# and the 2-GPU bench line with the detect_tiled extra (all ranks taking part)
O=gpurun_out/r02v; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 tools/detect_tiled_nccl.py 4096 2000 > $O/detect_tiled_2gpu.json 2> $O/detect_tiled_2gpu.err; echo "tiled detect rc=$?"; tail -n 1 $O/detect_tiled_2gpu.json | cut -c1-1500; grep -v "^\s*$" $O/detect_tiled_2gpu.err | grep -v "\*\*\*\|OMP_NUM" | tail -n 8 | cut -c1-300
timeout 600 $TR --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench2 rc=$?"; grep -v "\*\*\*\|OMP_NUM\|^\s*$" $O/bench_2gpu.err | tail -n 5 | cut -c1-300
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02v/bench_2gpu.json').read().strip().splitlines()[-1])
print("2gpu value",round(b['value']),"e2e",round(b['e2e']['value']))
ft=b.get('field_tiled') or {}
print('field_tiled', {kk:vv for kk,vv in ft.items() if kk not in ('api','collectives','timing','detect_tiled')})
print('detect_tiled', {kk:vv for kk,vv in (ft.get('detect_tiled') or {}).items() if kk not in ('collectives','timing')})
PY
