#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_network.py -q -m gpu -x -k "chunking or ragged or cfg2" > gpurun_out/net_host.log 2>&1; echo "host-path tests rc=$?"; tail -n 3 gpurun_out/net_host.log
python tools/e2e_breakdown.py 2>&1 | grep -E "deblend\(|growing|HOST_PIECE"
