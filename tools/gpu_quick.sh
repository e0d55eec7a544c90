#!/bin/bash
# quick correctness + speed check of the tensor-core network path: parity tests, then a short bench with per-layer times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu -x -k "${TESTK:-tensor_core or emulating or bf16x3 or x3}" > gpurun_out/quick_tests.log 2>&1; echo "tests rc=$?"; tail -n ${TAILN:-6} gpurun_out/quick_tests.log
DBV_VERBOSE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/quick.json 2> gpurun_out/quick.err; echo "bench rc=$?"
grep "halo plan" gpurun_out/quick.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/quick.json').read().strip().splitlines()[-1])
print("value",round(b['value']),"ms/step",round(b['ms_per_step'],3))
print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.2f}" for l in b['layers']))
PY
