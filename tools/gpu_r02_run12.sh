#!/bin/bash
# round 2, GPU call 12 (2 GPUs): the tiled device detector over NCCL — bit-identity with the single-GPU list on every rank, the
# assembled-field path on a footprint wider than the halo, the tiled iterative loop with it (4096^2 field, 2000 sources)
O=gpurun_out/r02u; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29551 tools/detect_tiled_nccl.py 4096 2000 > $O/detect_tiled_2gpu.json 2> $O/detect_tiled_2gpu.err; echo "tiled detect rc=$?"; tail -n 1 $O/detect_tiled_2gpu.json | cut -c1-1500; grep -v "^\s*$" $O/detect_tiled_2gpu.err | tail -n 12 | cut -c1-300
