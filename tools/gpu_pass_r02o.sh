#!/bin/bash
# Pass r02o (TWO GPUs, short): the shipped kernels on the torchrun path, and the multi-device assembly / elimination /
# field rows across two DISTINCT devices (peer copies over NVLink).
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_group.py -q --timeout 60 -k "dev01 and (assembly or boundary or field_recovery or errors)" > gpurun_out/r02o_pytest_group_2gpu.log 2>&1
echo "== group rows on devices 0,1: $(tail -1 gpurun_out/r02o_pytest_group_2gpu.log)"
grep -E "FAILED|^E  " gpurun_out/r02o_pytest_group_2gpu.log | cut -c1-300 | head -8
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02o_bench_2gpu_torchrun.json 2> gpurun_out/r02o_bench_2gpu_torchrun.err
python -c "
import json; d=json.loads(open('gpurun_out/r02o_bench_2gpu_torchrun.json').read().strip().splitlines()[-1]); print('torchrun x2', d['value'], d['e2e'], d['roofline']['frac'], d['x_checksum'], d.get('x_checksum_rel_to_1gpu'), d.get('nit'))"
tail -2 gpurun_out/r02o_bench_2gpu_torchrun.err | cut -c1-300
