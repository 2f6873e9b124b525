#!/bin/bash
# Pass r02g (TWO GPUs): the multi-device context across distinct devices, the torchrun path beside it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 400 python -m pytest tests/test_gpu_group.py tests/test_gpu_dist.py -q --timeout 150 > gpurun_out/r02g_pytest_group_dist.log 2>&1
echo "== group + dist on 2 GPUs: $(tail -1 gpurun_out/r02g_pytest_group_dist.log)"
grep -E "FAILED|^E  " gpurun_out/r02g_pytest_group_dist.log | head -10
AMIE_B200_DEVICES=0,1 timeout 400 python -m pytest tests/test_gpu_e2e.py -q -s --timeout 350 > gpurun_out/r02g_pytest_e2e_2gpu.log 2>&1
echo "== e2e (FeatureTree drop-in) with AMIE_B200_DEVICES=0,1: $(tail -1 gpurun_out/r02g_pytest_e2e_2gpu.log)"
grep -E "tripoint:|e2e 2d|e2e 3d|FAILED|^E  " gpurun_out/r02g_pytest_e2e_2gpu.log | cut -c1-300 | head
timeout 200 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "enrichment" > gpurun_out/r02g_pytest_enrich.log 2>&1
echo "== enrichment-like rows: $(tail -1 gpurun_out/r02g_pytest_enrich.log)"
timeout 400 python bench.py --gpus 2 --single-process --no-cpu --no-upload > gpurun_out/r02g_bench_2gpu_single.json 2> gpurun_out/r02g_bench_2gpu_single.err
python -c "
import json; d=json.loads(open('gpurun_out/r02g_bench_2gpu_single.json').read().strip().splitlines()[-1]); print('single-process x2', d['value'], d['e2e'], d['roofline']['frac'], d['x_checksum'], d['config'].get('parallelism'))"
tail -2 gpurun_out/r02g_bench_2gpu_single.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02g_bench_2gpu_torchrun.json 2> gpurun_out/r02g_bench_2gpu_torchrun.err
python -c "
import json; d=json.loads(open('gpurun_out/r02g_bench_2gpu_torchrun.json').read().strip().splitlines()[-1]); print('torchrun x2', d['value'], d['e2e'], d['roofline']['frac'], d['x_checksum'], d['x_checksum_rel_to_1gpu'], d['nit'])"
tail -2 gpurun_out/r02g_bench_2gpu_torchrun.err
