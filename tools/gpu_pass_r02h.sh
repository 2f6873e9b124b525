#!/bin/bash
# Pass r02h (EIGHT GPUs, short): ONE context over 8 devices -- the group tests across distinct GPUs and the bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index --format=csv,noheader | wc -l
timeout 240 python -m pytest tests/test_gpu_group.py -q --timeout 100 -k "dev01234567 or dev0123" > gpurun_out/r02h_pytest_group_8gpu.log 2>&1
echo "== group on 4 and 8 GPUs: $(tail -1 gpurun_out/r02h_pytest_group_8gpu.log)"
grep -E "FAILED|^E  " gpurun_out/r02h_pytest_group_8gpu.log | head -6
for n in 8 4; do
timeout 240 python bench.py --gpus $n --single-process --no-cpu --no-upload > gpurun_out/r02h_bench_${n}gpu_single.json 2> gpurun_out/r02h_bench_${n}gpu_single.err
python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_${n}gpu_single.json').read().strip().splitlines()[-1]); print('single-process x$n', d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['launch_ms'], d['x_checksum'], d['clocks'])"
tail -2 gpurun_out/r02h_bench_${n}gpu_single.err
done
