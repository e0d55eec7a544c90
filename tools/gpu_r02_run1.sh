#!/bin/bash
# round 2, GPU call 1: whole GPU suite, bench line, clock64 breakdown of the halo kernels, MMA/epilogue ablations, field-kernel DRAM bytes
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi -L > $O/gpu.txt
timeout 1200 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 5 $O/tests.log
DBV_VERBOSE=1 timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python tools/halo_clocks.py mixed 4096 > $O/halo_clocks_mixed.json 2> $O/halo_clocks_mixed.err; echo "clocks rc=$?"
timeout 300 python tools/halo_clocks.py bf16x3 4096 > $O/halo_clocks_bf16x3.json 2> $O/halo_clocks_bf16x3.err
for skip in 1 2; do
  DEBVADER_B200_LIB=$PWD/debvader_b200/libdebvader_b200_ablate.so DBV_HALO_SKIP=$skip timeout 300 python bench.py --steps 5 --warmup 3 --no-extras > $O/abl_$skip.json 2> $O/abl_$skip.err
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"extract_bulk_kernel|window_axpy_kernel|sqdiff_partial|axpy_bin" --launch-skip 10 -c 12 --csv --log-file $O/field_ncu.csv python tools/field_ncu_target.py > $O/field_ncu.log 2>&1; echo "field ncu rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r02a/bench.json').read().strip().splitlines()[-1])
print("value",round(b['value']),"e2e",round(b['e2e']['value']),"f64 e2e",b['e2e'].get('pageable_f64_input',{}).get('value'))
print(" ".join(f"{l['layer'].replace('enc_','e').replace('dec_','d')}={l['ms']:.3f}" for l in b['layers']))
f=b.get('field',{})
for k in ('extract_f64','extract_f64_to_f32','window_axpy_f64','window_axpy_f64_inplace','ms_per_field_kernels','ms_per_field','cfg1_dc2_field'):
    print(k, f.get(k))
print('field_tiled', b.get('field_tiled'))
PY
