#!/bin/bash
# round 2, GPU call 4: new parity tests, tuner on 2368 stamps, ncu --set full with source of the halo kernels
O=gpurun_out/r02d; mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 5 $O/tests.log
for rep in 1 2; do
  DBV_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > $O/bench_$rep.json 2> $O/bench_$rep.err; echo "bench rc=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "bench full rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-extras --batch 2048"
$CMD > $O/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"tc_halo_kernel" -c 7 -o $O/prof_halo -f $CMD > $O/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 $O/ncu_full.log | cut -c1-200
python - <<'PY'
import json
for f in ("bench_1","bench_2","bench_full"):
    try:
        b=json.loads(open(f'gpurun_out/r02d/{f}.json').read().strip().splitlines()[-1])
        print(f,"value",round(b['value']),"e2e",round(b['e2e']['value']))
        print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.3f}" for l in b['layers']))
        if f=="bench_full":
            fl=b.get('field',{})
            for k in ('window_axpy_f64','window_axpy_f64_inplace','ms_per_field_kernels','ms_per_field','cfg1_dc2_field'):
                print(k, {kk:vv for kk,vv in fl.get(k,{}).items() if kk not in ('note','includes','api','field')})
            print('field_tiled', {kk:vv for kk,vv in (b.get('field_tiled') or {}).items() if kk not in ('api','collectives','timing')})
            print('e2e', b['e2e']['value'], b['e2e'].get('pageable_f64_input',{}).get('value'))
    except Exception as e: print(f,"ERR",e)
PY
grep "halo plan" gpurun_out/r02d/bench_1.err
