#!/bin/bash
# Pass r02d (ONE GPU): multi-device context with eager module loading, at-scale parity against the reference, new bench line.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_group.py -q -x > gpurun_out/r02d_pytest_group.log 2>&1
echo "== group: $(tail -1 gpurun_out/r02d_pytest_group.log)"
grep -E "FAILED|Error|assert" gpurun_out/r02d_pytest_group.log | head -20
AMIE_B200_DEVICES=0,0 timeout 300 python -m pytest tests/test_gpu_e2e.py -q > gpurun_out/r02d_pytest_e2e_group.log 2>&1
echo "== e2e with AMIE_B200_DEVICES=0,0: $(tail -1 gpurun_out/r02d_pytest_e2e_group.log)"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k at_scale > gpurun_out/r02d_pytest_at_scale.log 2>&1
echo "== at scale: $(tail -1 gpurun_out/r02d_pytest_at_scale.log)"
grep -E "DOF, reference" gpurun_out/r02d_pytest_at_scale.log
free -g | head -2
timeout 1000 python bench.py > gpurun_out/r02d_bench_1gpu.json 2> gpurun_out/r02d_bench_1gpu.err
tail -c 3000 gpurun_out/r02d_bench_1gpu.json
tail -5 gpurun_out/r02d_bench_1gpu.err
