#!/bin/bash
# one `ncu --set full` capture of ONE timed step (the profiler API window in bench.py excludes the autotuner)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-extras --batch ${BATCH:-1024}"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"${KREGEX:-tc_halo_kernel|tc_conv_kernel|bn_pack8|simt_conv|latent}" -c ${COUNT:-23} -o gpurun_out/prof_${TAG:-full} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
