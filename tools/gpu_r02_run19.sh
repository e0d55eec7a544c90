#!/bin/bash
# round 2, GPU call 19: ncu --set full (source-level) capture of det_mesh_kernel
O=gpurun_out/r02z; mkdir -p $O
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"det_mesh_kernel" --launch-skip 1 -c 1 -o $O/det_mesh -f python tools/detect_ncu_target.py > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -n 2 $O/ncu.log | cut -c1-200
ls -la $O
