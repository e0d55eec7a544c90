#!/bin/bash
# round 2, GPU call 2: flat MMA issue tables + parity-class N-concatenation in the halo kernel
O=gpurun_out/r02b; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 5 $O/tests.log
DBV_VERBOSE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-extras > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python tools/halo_clocks.py mixed 4096 > $O/halo_clocks_mixed.json 2> $O/halo_clocks_mixed.err; echo "clocks rc=$?"
DEBVADER_B200_LIB=$PWD/debvader_b200/libdebvader_b200_ablate.so DBV_NO_CONCAT=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras > $O/noconcat.json 2> $O/noconcat.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --precision bf16x3 > $O/bf16x3.json 2> $O/bf16x3.err
timeout 300 python tools/profile_field_api.py > $O/profile_field_api.txt 2>&1
python - <<'PY'
import json
for f in ("bench","noconcat","bf16x3"):
    try:
        b=json.loads(open(f'gpurun_out/r02b/{f}.json').read().strip().splitlines()[-1])
        print(f,"value",round(b['value']),"e2e",round(b['e2e']['value']))
        print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.3f}" for l in b['layers']))
    except Exception as e: print(f,"ERR",e)
PY
head -45 gpurun_out/r02b/profile_field_api.txt
