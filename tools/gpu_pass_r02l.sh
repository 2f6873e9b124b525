#!/bin/bash
# Pass r02l (ONE GPU): producer with unconditional row-pointer loads, one more stage per pipeline (variant 6), the 2x2 column
# mapping as default (variant 4 = row mapping); assembly / elimination / field recovery on multi-device contexts (parts on
# device 0); 3d-1000 end to end.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_group.py -q --timeout 120 > gpurun_out/r02l_pytest_group.log 2>&1
echo "== group: $(tail -1 gpurun_out/r02l_pytest_group.log)"
grep -E "^E  |FAILED|Error" gpurun_out/r02l_pytest_group.log | head -12
timeout 300 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_recovery.py tests/test_gpu_variants.py tests/test_gpu_parity.py -q --timeout 120 \
    -k "not at_scale and not full_size and not throughput" > gpurun_out/r02l_pytest_rows.log 2>&1
echo "== assembly / recovery / variants / parity: $(tail -1 gpurun_out/r02l_pytest_rows.log)"
grep -E "^E  |FAILED" gpurun_out/r02l_pytest_rows.log | head -10
timeout 200 python -m pytest tests/test_gpu_e2e.py -q -s --timeout 150 -k "1000" > gpurun_out/r02l_pytest_e2e.log 2>&1
echo "== e2e 3d-1000: $(tail -1 gpurun_out/r02l_pytest_e2e.log)"
grep -E "e2e 3d|^E  " gpurun_out/r02l_pytest_e2e.log | cut -c1-400
PROBE_VARIANTS=0,6,100,106 timeout 200 python tools/gpu_probe.py S3-hex:256 S3-tet:256 > gpurun_out/r02l_probe_s3.json 2> gpurun_out/r02l_probe_s3.err
cat gpurun_out/r02l_probe_s3.json
PROBE_VARIANTS=0,4,6,100,104,106 timeout 200 python tools/gpu_probe.py S2-tri:4096 > gpurun_out/r02l_probe_tri.json 2> gpurun_out/r02l_probe_tri.err
cat gpurun_out/r02l_probe_tri.json
for v in 0 6; do
    timeout 300 python bench.py --spmv-variant $v --steps 1 --warmup 1 --no-cpu --no-e2e --no-upload > gpurun_out/r02l_bench_hex_v$v.json 2> gpurun_out/r02l_bench_hex_v$v.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02l_bench_hex_v$v.json').read().strip().splitlines()[-1]); print('hex v$v', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'], d['x_checksum'], d['config']['iterations_per_step'])"
done
