#!/bin/bash
# quick iteration: tensor-core parity tests + short bench (no extras)
mkdir -p gpurun_out
DBV_VERBOSE=1 timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu -k "tensor_core or emulating or chunking or cfg2" -s -x > gpurun_out/iter_tests.log 2>&1; echo "tests rc=$?"
grep -E "halo plan|per-layer|err/peak|passed|failed|Error|error|assert" gpurun_out/iter_tests.log | cut -c1-1500 | head -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/iter_bench.err
python - <<'PY'
import json
try:
    b=json.loads(open('gpurun_out/iter_bench.json').read().strip().splitlines()[-1])
    print("value",round(b['value']),"e2e",round(b['e2e']['value']),"ms/step",round(b['ms_per_step'],2),"net frac",b['network']['frac_of_bf16_sustained'])
    print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.2f}" for l in b['layers']))
except Exception as e: print("no bench json", e)
PY
