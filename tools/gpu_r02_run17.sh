#!/bin/bash
# round 2, GPU call 17: smoke (now with the detector check), detector GPU tests, detector timing + per-kernel ncu list after the batched tile loads
O=gpurun_out/r02x; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python -m pytest tests/test_gpu_detect.py -x -q -m gpu > $O/detect_tests.log 2>&1; echo "detect tests rc=$?"; tail -n 3 $O/detect_tests.log
timeout 300 python tools/detect_ncu_target.py > $O/detect_plain.log 2>&1; echo "detect plain rc=$?"; tail -n 1 $O/detect_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"det_" --launch-skip 16 -c 16 --csv --log-file $O/detect_ncu.csv python tools/detect_ncu_target.py > $O/detect_ncu.log 2>&1; echo "detect ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02x/detect_ncu.csv')) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((int(r[0]),r[4].split('(')[0]),{})[r[12]]=float(r[14].replace(',',''))
tot=0
for (i,k),v in sorted(d.items()):
    tot+=v.get('gpu__time_duration.sum',0); print(i,k,round(v.get('gpu__time_duration.sum',0)/1e3,1))
print("sum us", tot/1e3)
PY
