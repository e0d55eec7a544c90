#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_field_ops.py -q -m gpu -x > gpurun_out/field_ops.log 2>&1; echo "field ops tests rc=$?"; tail -n 4 gpurun_out/field_ops.log
timeout 600 python tools/bench_field.py > gpurun_out/bench_field.log 2>&1; echo "bench_field rc=$?"; grep -E "window_axpy|spline|sub-pixel|position" gpurun_out/bench_field.log
for v in "DBV_AXPY_BINS=0"; do
  echo "--- $v"; env $v timeout 300 python tools/bench_field.py 2>&1 | grep -E "^window_axpy|sub-pixel" 
done
