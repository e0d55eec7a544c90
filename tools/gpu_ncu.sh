#!/bin/bash
# one `ncu --set full` capture of the halo / simt kernels of one chunk (plain run first, as the recipe requires)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-extras --batch 512"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-tc_halo_kernel|simt_conv}" -s ${SKIP:-0} -c ${COUNT:-8} -o gpurun_out/prof_${TAG:-halo} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
