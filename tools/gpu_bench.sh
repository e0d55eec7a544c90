#!/bin/bash
# bench + ncu launch list (run plain first, as the profiling recipe requires); outputs in gpurun_out/
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 7000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 168 -c 336 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
