#!/bin/bash
# round 2, GPU call 8: the fp32 tier on tensor cores (precision="fp32tc": segmented tcgen05 accumulation promoted in fp32 registers)
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -k "tensor_core_path or cfg2 or fp32_path or tcgen05_gemm_probe" > $O/tests.log 2>&1; echo "tests rc=$?"; grep -E "err/peak|per-layer|passed|failed|Error" $O/tests.log | cut -c1-1500 | tail -20
timeout 600 python - > $O/fp32tc_rate.txt 2>&1 <<'PY'
import torch, time
from debvader_b200.model.model import load_deblender
x = torch.randn((4096, 59, 59, 6), device="cuda") * 0.3
mean = torch.empty_like(x); std = torch.empty_like(x)
for prec in ("fp32tc", "fp16x3", "fp32"):
    net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234", precision=prec)
    for _ in range(2): net.deblend_into(x, mean, std)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): net.deblend_into(x, mean, std)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(prec, f"{ms:.3f} ms per 4096 stamps = {4096 / ms * 1e3:.0f} stamps/s")
    net.set_profiling(True)
    net.deblend_into(x, mean, std); torch.cuda.synchronize()
    print("   ", " ".join(f"{n.replace('enc_','e').replace('dec_','d')}={t:.3f}" for n, t in net.layer_times()))
    net.close()
PY
echo "rate rc=$?"; cat $O/fp32tc_rate.txt | cut -c1-900
