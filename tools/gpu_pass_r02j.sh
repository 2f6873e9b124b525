#!/bin/bash
# Pass r02j (ONE GPU): x gather of the 3x3 pipeline with lane <-> element (default) against lane <-> block (variant 4):
# parity, isolated launches, inside the solve, ncu of the new launch inside the solve; assembly gather with prefetch.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py -q --timeout 120 -k "not at_scale and not full_size" > gpurun_out/r02j_pytest_parity.log 2>&1
echo "== parity + variants: $(tail -1 gpurun_out/r02j_pytest_parity.log)"
grep -E "^E  |FAILED" gpurun_out/r02j_pytest_parity.log | head -10
PROBE_VARIANTS=0,4,100,104,0,4,100,104 timeout 200 python tools/gpu_probe.py S3-hex:256 > gpurun_out/r02j_probe_hex.json 2> gpurun_out/r02j_probe_hex.err
cat gpurun_out/r02j_probe_hex.json
PROBE_VARIANTS=0,4,100,104 timeout 200 python tools/gpu_probe.py S3-tet:256 > gpurun_out/r02j_probe_tet.json 2> gpurun_out/r02j_probe_tet.err
cat gpurun_out/r02j_probe_tet.json
for v in 0 4; do
    timeout 300 python bench.py --spmv-variant $v --steps 1 --warmup 1 --no-cpu --no-e2e --no-upload > gpurun_out/r02j_bench_hex_v$v.json 2> gpurun_out/r02j_bench_hex_v$v.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02j_bench_hex_v$v.json').read().strip().splitlines()[-1]); print('hex v$v', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'], d['x_checksum'], d['config']['iterations_per_step'])"
done
for v in 0 4; do
    timeout 300 python bench.py --preset S3-tet --spmv-variant $v --steps 1 --warmup 1 --no-cpu --no-e2e --no-upload > gpurun_out/r02j_bench_tet_v$v.json 2> gpurun_out/r02j_bench_tet_v$v.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02j_bench_tet_v$v.json').read().strip().splitlines()[-1]); print('tet v$v', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'], d['x_checksum'], d['config']['iterations_per_step'])"
done
timeout 200 python tools/probe_next_rows.py assembly > gpurun_out/r02j_probe_assembly.jsonl 2> gpurun_out/r02j_probe_assembly.err
cut -c1-420 gpurun_out/r02j_probe_assembly.jsonl
timeout 120 python -m pytest tests/test_gpu_assembly.py -q --timeout 100 > gpurun_out/r02j_pytest_assembly.log 2>&1
echo "== assembly: $(tail -1 gpurun_out/r02j_pytest_assembly.log)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_s3_rt" -s 40 -c 2 -o gpurun_out/r02j_prof_spmv_insolve \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-upload > gpurun_out/r02j_prof_spmv.log 2>&1
ls -la gpurun_out/r02j_prof_spmv_insolve.ncu-rep
