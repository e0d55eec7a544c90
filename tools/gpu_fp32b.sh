#!/bin/bash
mkdir -p gpurun_out
for v in "DBV_SIMT_KC=16 DBV_SIMT_KC_SMALL=16" "DBV_SIMT_KC=32 DBV_SIMT_KC_SMALL=8"; do
env $v timeout 600 python bench.py --precision fp32 --steps 2 --warmup 3 --no-extras --batch 1024 > gpurun_out/fp32_v.json 2> gpurun_out/fp32_v.err; echo "fp32 bench ($v) rc=$?"
python - <<PY
import json
b=json.loads(open('gpurun_out/fp32_v.json').read().strip().splitlines()[-1])
print("fp32 value",round(b['value']),"ms/step",round(b['ms_per_step'],3))
print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.2f}" for l in b['layers']))
PY
done
