#!/bin/bash
mkdir -p gpurun_out
DBV_PDL=1 timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -x -k "tensor_core or chunking or ragged or cfg2 or real_dc2" > gpurun_out/net_pdl.log 2>&1; echo "PDL tests rc=$?"; tail -n 3 gpurun_out/net_pdl.log
for v in 0 1 0 1; do
DBV_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/pdl_$v.json 2> gpurun_out/pdl_$v.err; echo "bench PDL=$v rc=$?"
python - <<PY
import json
b=json.loads(open('gpurun_out/pdl_$v.json').read().strip().splitlines()[-1])
print("PDL=$v value",round(b['value']),"ms/step",round(b['ms_per_step'],3),"e2e",round(b['e2e']['value']))
PY
done
