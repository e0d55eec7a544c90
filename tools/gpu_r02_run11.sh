#!/bin/bash
# round 2, GPU call 11: the device detector — GPU parity tests against the oracle, per-kernel times / DRAM bytes,
# and a bench line with the new roofline grouping (dominant __global__ function) and the `detect` extras
O=gpurun_out/r02t; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_detect.py -x -q -m gpu > $O/detect_tests.log 2>&1; echo "detect tests rc=$?"; tail -n 25 $O/detect_tests.log
timeout 300 python tools/detect_ncu_target.py > $O/detect_plain.log 2>&1; echo "detect plain rc=$?"; tail -n 2 $O/detect_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"det_" --launch-skip 19 -c 19 --csv --log-file $O/detect_ncu.csv python tools/detect_ncu_target.py > $O/detect_ncu.log 2>&1; echo "detect ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02t/detect_ncu.csv')) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((int(r[0]),r[4].split('(')[0]),{})[r[12]]=float(r[14].replace(',',''))
for (i,k),v in sorted(d.items()): print(i,k,{a:round(b,1) for a,b in v.items()})
PY
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -n 3 $O/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02t/bench.json').read().strip().splitlines()[-1])
print("value",round(b['value']),"e2e",round(b['e2e']['value']))
r=b['roofline']; print("roofline", {k:v for k,v in r.items() if k not in ('kernels','traffic_detail')})
for k in r['kernels']: print("  ", k)
f=b.get('field',{})
for k in ('detect','iterative_device_detector','ms_per_field','cfg1_dc2_field'):
    print(k, {kk:vv for kk,vv in (f.get(k) or {}).items() if kk not in ('note','includes','api','field')})
PY
