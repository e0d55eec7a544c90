#!/bin/bash
# Record the round's evidence: bench (with extras), ncu launch list, one ncu --set full capture of the top kernels.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 93 -c 186 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-extras --batch 1024"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_halo_kernel|tc_conv_kernel" -s 69 -c 20 -o gpurun_out/prof_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
