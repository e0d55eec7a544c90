#!/bin/bash
# whole GPU test suite (as the driver runs it) + the fp32-tier throughput
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/all_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -n 6 gpurun_out/all_tests.log
timeout 600 python bench.py --precision fp32 --steps 2 --warmup 3 --no-extras --batch 1024 > gpurun_out/fp32.json 2> gpurun_out/fp32.err; echo "fp32 bench rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/fp32.json').read().strip().splitlines()[-1])
print("fp32 value",round(b['value']),"ms/step",round(b['ms_per_step'],3))
print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.2f}" for l in b['layers']))
PY
