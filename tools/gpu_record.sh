#!/bin/bash
# Record the round's evidence on the GPU box: bench (with extras), the reference arm, the ncu launch list of the timed
# region and one `ncu --set full` capture of one chunk of every kernel.  Outputs in gpurun_out/; tools/make_profile_summary.py
# turns them into the tracked files under profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$?"
# launch list: per-launch durations of the two timed steps only (bench.py brackets them with cudaProfilerStart/Stop)
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"
# full capture: one step of ONE chunk (batch 1024 = chunk), every kernel once
CMD="python bench.py --steps 1 --warmup 3 --no-extras --batch 1024"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -c 30 -o gpurun_out/prof_full -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2>/dev/null; wc -l gpurun_out/prof_full_raw.csv
