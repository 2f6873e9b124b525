#!/bin/bash
# Pass r02n (ONE GPU, short): the bench line with the reworked same-size sample (restart from the GPU's own truncated solution
# when nit(eps) jumps over the window).
mkdir -p gpurun_out
timeout 420 python bench.py --steps 1 --warmup 0 --no-e2e --no-upload > gpurun_out/r02n_bench_1gpu.json 2> gpurun_out/r02n_bench_1gpu.err
tail -3 gpurun_out/r02n_bench_1gpu.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r02n_bench_1gpu.json').read().strip().splitlines()[-1]); c=d['cpu_baseline']; print('bench', d['value'], d['roofline']['frac'], 'cpu', c['value'], c['same_config'], json.dumps(c['same_size_pair'])[:900])"
