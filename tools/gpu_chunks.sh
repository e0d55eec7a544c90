#!/bin/bash
mkdir -p gpurun_out
for ch in 64 128 256 512 1024 2048; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --chunk $ch > gpurun_out/chunk_$ch.json 2>gpurun_out/chunk_$ch.err
  python - $ch <<'PY'
import json,sys
b=json.loads(open(f'gpurun_out/chunk_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print("chunk",sys.argv[1],"ms/step",round(b['ms_per_step'],2),"value",round(b['value']),"e2e",round(b['e2e']['value'])," ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.2f}" for l in b['layers']))
PY
done
