#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -x -k "fp32 or sampling or encoder_decoder" > gpurun_out/net_fp32.log 2>&1; echo "fp32 tests rc=$?"; tail -n 6 gpurun_out/net_fp32.log
for v in 1 0; do
DBV_SIMT_TILED=$v timeout 600 python bench.py --precision fp32 --steps 2 --warmup 3 --no-extras --batch 1024 > gpurun_out/fp32_$v.json 2> gpurun_out/fp32_$v.err; echo "fp32 bench (tiled=$v) rc=$?"
python - <<PY
import json
b=json.loads(open('gpurun_out/fp32_$v.json').read().strip().splitlines()[-1])
print("fp32 value",round(b['value']),"ms/step",round(b['ms_per_step'],3))
print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.2f}" for l in b['layers']))
PY
done
