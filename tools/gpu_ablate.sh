#!/bin/bash
mkdir -p gpurun_out
for skip in ${SKIPS:-0 1 2 3}; do
  DBV_HALO_SKIP=$skip timeout 300 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/abl_$skip.json 2>gpurun_out/abl_$skip.err
  python - $skip <<'PY'
import json,sys
b=json.loads(open(f'gpurun_out/abl_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print("skip",sys.argv[1],"ms/step",round(b['ms_per_step'],2)," ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.2f}" for l in b['layers'] if l['layer'] in ('enc_conv1','enc_conv2','enc_conv3','dec_convT6','dec_convT7','dec_convT8','dec_head')))
PY
done
