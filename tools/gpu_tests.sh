#!/bin/bash
# Run the GPU test files one by one, least risky first, each under its own timeout, logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
run() { # name, timeout, cmd...
  local name=$1; local t=$2; shift 2
  echo "=== $name ==="
  timeout $t "$@" > gpurun_out/$name.log 2>&1; local rc=$?
  echo "$name rc=$rc"; tail -n ${TAILN:-15} gpurun_out/$name.log
}
run field_ops 600 python -m pytest tests/test_gpu_field_ops.py -q -m gpu -x
run probe 180 python -m pytest tests/test_gpu_network.py -q -m gpu -k "probe" -s
run net_fp32 600 python -m pytest tests/test_gpu_network.py -q -m gpu -k "fp32 or sampling" -s
run net_tc 600 python -m pytest tests/test_gpu_network.py -q -m gpu -k "tensor_core or emulating" -s
run net_rest 900 python -m pytest tests/test_gpu_network.py -q -m gpu -k "not probe and not fp32 and not sampling and not tensor_core and not emulating" -s
run field_deblend 600 python -m pytest tests/test_gpu_field_deblend.py -q -m gpu -s
