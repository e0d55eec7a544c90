#!/bin/bash
# round 2, GPU call 7: whole GPU suite (no -x: every failure listed), field micro-benchmark with the warp-per-row window kernel
# (A/B against the generic kernel through the ablation build), field-kernel DRAM bytes, clock64 breakdown of the halo kernels
O=gpurun_out/r02g; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 8 $O/tests.log
timeout 600 python tools/bench_field.py > $O/bench_field.log 2>&1; echo "bench_field rc=$?"; cat $O/bench_field.log
echo "--- generic window kernel (ablation build, DBV_AXPY_GENERIC=1)"
DEBVADER_B200_LIB=$PWD/debvader_b200/libdebvader_b200_ablate.so DBV_AXPY_GENERIC=1 timeout 300 python tools/bench_field.py 2>&1 | grep -E "^window_axpy"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"extract_bulk_kernel|window_axpy|sqdiff_partial|axpy_bin" --launch-skip 10 -c 12 --csv --log-file $O/field_ncu.csv python tools/field_ncu_target.py > $O/field_ncu.log 2>&1; echo "field ncu rc=$?"
timeout 300 python tools/halo_clocks.py mixed 4096 > $O/halo_clocks_mixed.json 2> $O/halo_clocks_mixed.err; echo "clocks rc=$?"; tail -c 3000 $O/halo_clocks_mixed.json
