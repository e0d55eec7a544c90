#!/bin/bash
# round-end check the way the driver does it: whole GPU suite, smoke(), then the evidence record
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/all_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -n 3 gpurun_out/all_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
bash tools/gpu_record.sh
timeout 300 python tools/bench_field.py > gpurun_out/bench_field.log 2>&1; echo "bench_field rc=$?"
