#!/bin/bash
# round 2, GPU call 14: the whole evidence record of the final code — GPU suite (incl. the device detector), smoke, bench (+reference
# arm), ncu launch list, ncu --set full of one chunk, field-kernel and detector-kernel DRAM bytes / times, PCIe probe
bash tools/gpu_final.sh
O=gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"extract_bulk_kernel|window_axpy|sqdiff_partial|axpy_bin" --launch-skip 10 -c 12 --csv --log-file $O/field_ncu.csv python tools/field_ncu_target.py > $O/field_ncu.log 2>&1; echo "field ncu rc=$?"
timeout 300 python tools/detect_ncu_target.py > $O/detect_plain.log 2>&1; echo "detect plain rc=$?"; tail -n 1 $O/detect_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"det_" --launch-skip 20 -c 20 --csv --log-file $O/detect_ncu.csv python tools/detect_ncu_target.py > $O/detect_ncu.log 2>&1; echo "detect ncu rc=$?"
timeout 300 python tools/pcie_probe.py > $O/pcie_probe_1gpu.log 2>&1; echo "pcie rc=$?"
python - <<'PY'
import json, csv
b=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("value",round(b['value']),"e2e",round(b['e2e']['value']),"f64 e2e",b['e2e'].get('pageable_f64_input',{}).get('value'))
r=b['roofline']; print("roofline", {k:v for k,v in r.items() if k not in ('kernels','traffic_detail','slowest_layer')})
print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.3f}" for l in b['layers']))
f=b.get('field',{})
for k in ('extract_f64','extract_f64_to_f32','window_axpy_f64','window_axpy_f64_inplace','ms_per_field_kernels','ms_per_field','cfg1_dc2_field','detect','iterative_device_detector'):
    print(k, {kk:vv for kk,vv in (f.get(k) or {}).items() if kk not in ('note','includes','api','field','traffic')})
ft=b.get('field_tiled') or {}
print('field_tiled', {kk:vv for kk,vv in ft.items() if kk not in ('api','collectives','timing','detect_tiled')})
print('detect_tiled', {kk:vv for kk,vv in (ft.get('detect_tiled') or {}).items() if kk not in ('collectives','timing')})
rows=[r for r in csv.reader(open('gpurun_out/detect_ncu.csv')) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((int(r[0]),r[4].split('(')[0]),{})[r[12]]=float(r[14].replace(',',''))
tot=0
for (i,k),v in sorted(d.items()):
    tot+=v.get('gpu__time_duration.sum',0); print(i,k,{a:round(b,1) for a,b in v.items()})
print("detector kernels, sum us", tot/1e3)
PY
