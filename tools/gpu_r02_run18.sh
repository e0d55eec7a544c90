#!/bin/bash
# round 2, GPU call 18: detector timing + per-kernel ncu list with the largest shared-memory carve-out for det_mesh_kernel; detector tests
O=gpurun_out/r02y; mkdir -p $O
timeout 300 python tools/detect_ncu_target.py > $O/detect_plain.log 2>&1; echo "detect plain rc=$?"; tail -n 1 $O/detect_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"det_" --launch-skip 16 -c 16 --csv --log-file $O/detect_ncu.csv python tools/detect_ncu_target.py > $O/detect_ncu.log 2>&1; echo "detect ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_detect.py -x -q -m gpu > $O/detect_tests.log 2>&1; echo "detect tests rc=$?"; tail -n 3 $O/detect_tests.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02y/detect_ncu.csv')) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((int(r[0]),r[4].split('(')[0]),{})[r[12]]=float(r[14].replace(',',''))
tot=0
for (i,k),v in sorted(d.items()):
    tot+=v.get('gpu__time_duration.sum',0); print(i,k,round(v.get('gpu__time_duration.sum',0)/1e3,1), round(v.get('sm__warps_active.avg.pct_of_peak_sustained_active',0),1))
print("sum us", tot/1e3)
PY
