#!/bin/bash
# round 2, GPU call 6 (after the container was re-created): the whole evidence record of the current code —
# GPU suite, smoke, bench (+reference arm), ncu launch list, ncu --set full of one chunk, field-kernel DRAM bytes, PCIe probe.
bash tools/gpu_final.sh
O=gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"extract_bulk_kernel|window_axpy|sqdiff_partial|axpy_bin" --launch-skip 10 -c 12 --csv --log-file $O/field_ncu.csv python tools/field_ncu_target.py > $O/field_ncu.log 2>&1; echo "field ncu rc=$?"
timeout 300 python tools/pcie_probe.py > $O/pcie_probe_1gpu.log 2>&1; echo "pcie rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("value",round(b['value']),"e2e",round(b['e2e']['value']),"f64 e2e",b['e2e'].get('pageable_f64_input',{}).get('value'))
print("roofline", b['roofline'])
print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.3f}" for l in b['layers']))
f=b.get('field',{})
for k in ('extract_f64','extract_f64_to_f32','window_axpy_f64','window_axpy_f64_inplace','ms_per_field_kernels','ms_per_field','cfg1_dc2_field'):
    print(k, {kk:vv for kk,vv in (f.get(k) or {}).items() if kk not in ('note','includes','api','field')})
print('field_tiled', {kk:vv for kk,vv in (b.get('field_tiled') or {}).items() if kk not in ('api','collectives','timing')})
PY
