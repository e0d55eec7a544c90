#!/bin/bash
# round 2, GPU call 3: alpha prefetch, per-launch carve-out, 3-deep halo ring, window_axpy load fix, fp16 overflow watch
O=gpurun_out/r02c; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 5 $O/tests.log
A=$PWD/debvader_b200/libdebvader_b200_ablate.so
for rep in 1 2; do
  DBV_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > $O/bench_$rep.json 2> $O/bench_$rep.err; echo "bench rc=$?"
  DEBVADER_B200_LIB=$A DBV_HALO_CARVEOUT=100 timeout 300 python bench.py --steps 10 --warmup 3 --no-extras > $O/carve100_$rep.json 2> $O/carve100_$rep.err
done
timeout 300 python tools/halo_clocks.py mixed 4096 > $O/halo_clocks_mixed.json 2> $O/halo_clocks_mixed.err; echo "clocks rc=$?"
timeout 300 python tools/bench_field.py > $O/bench_field.log 2>&1
python - <<'PY'
import json
for f in ("bench_1","carve100_1","bench_2","carve100_2"):
    try:
        b=json.loads(open(f'gpurun_out/r02c/{f}.json').read().strip().splitlines()[-1])
        print(f,"value",round(b['value']),"e2e",round(b['e2e']['value']))
        print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.3f}" for l in b['layers']))
    except Exception as e: print(f,"ERR",e)
PY
grep "halo plan" gpurun_out/r02c/bench_1.err
grep -i "window_axpy\|extract N=16384" gpurun_out/r02c/bench_field.log
