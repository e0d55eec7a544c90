#!/bin/bash
# round 2, GPU call 16: smoke + the bench line (with extras) of the final code (the detector changed after call 14; the network kernels did not:
# the ncu launch list / full capture of call 14 stay valid)
O=gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/smoke.log
python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -2 $O/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("value",round(b['value']),"e2e",round(b['e2e']['value']),"f64 e2e",b['e2e'].get('pageable_f64_input',{}).get('value'))
r=b['roofline']; print("roofline", {k:v for k,v in r.items() if k in ('kernel','achieved','frac','share_of_step')})
f=b.get('field',{})
for k in ('ms_per_field','cfg1_dc2_field','detect','iterative_device_detector'):
    print(k, {kk:vv for kk,vv in (f.get(k) or {}).items() if kk not in ('note','includes','api','field','traffic')})
print('detect_tiled', {kk:vv for kk,vv in ((b.get('field_tiled') or {}).get('detect_tiled') or {}).items() if kk not in ('collectives','timing')})
PY
