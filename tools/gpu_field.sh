#!/bin/bash
# field-operator check on the GPU box: parity tests of the field kernels, then the micro-benchmark with the tuning knobs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_field_ops.py -q -m gpu -x > gpurun_out/field_ops.log 2>&1; echo "field ops tests rc=$?"; tail -n 8 gpurun_out/field_ops.log
timeout 600 python tools/bench_field.py > gpurun_out/bench_field.log 2>&1; echo "bench_field rc=$?"; cat gpurun_out/bench_field.log
for v in "DBV_EXTRACT_BULK=0" "DBV_EXTRACT_ROWS=8" "DBV_EXTRACT_ROWS=16" "DBV_EXTRACT_ROWS=5" "DBV_AXPY_BINS=0"; do
  echo "--- $v"; env $v timeout 300 python tools/bench_field.py 2>&1 | grep -E "^extract|^window_axpy" 
done
timeout 900 python -m pytest tests/test_gpu_field_deblend.py -q -m gpu -x > gpurun_out/field_deblend.log 2>&1; echo "field deblend tests rc=$?"; tail -n 5 gpurun_out/field_deblend.log
