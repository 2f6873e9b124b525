#!/bin/bash
# Pass r02f (ONE GPU): re-run of what failed for test-side reasons, SpMV A/B (elected barrier polls), bench lines for
# tetrahedra and 2D, ncu captures (SpMV in the solve, assembly).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_group.py -q --timeout 120 > gpurun_out/r02f_pytest_group.log 2>&1
echo "== group: $(tail -1 gpurun_out/r02f_pytest_group.log)"
timeout 400 python -m pytest tests/test_gpu_e2e.py -q -s --timeout 300 -k tripoint > gpurun_out/r02f_pytest_tripoint.log 2>&1
echo "== tripoint: $(tail -1 gpurun_out/r02f_pytest_tripoint.log)"
grep -E "tripoint:|^E  " gpurun_out/r02f_pytest_tripoint.log | cut -c1-400 | head -8
timeout 200 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "enrichment" > gpurun_out/r02f_pytest_enrich.log 2>&1
echo "== enrichment-like rows: $(tail -1 gpurun_out/r02f_pytest_enrich.log)"
grep -E "^E  " gpurun_out/r02f_pytest_enrich.log | head -5
# SpMV A/B on one box: isolated launches, then inside the solve
PROBE_VARIANTS=0,4,100,104 timeout 200 python tools/gpu_probe.py S3-hex:256 > gpurun_out/r02f_probe_hex.json 2> gpurun_out/r02f_probe_hex.err
cat gpurun_out/r02f_probe_hex.json
PROBE_VARIANTS=0,5,100,105 timeout 200 python tools/gpu_probe.py S3-tet:256 > gpurun_out/r02f_probe_tet.json 2> gpurun_out/r02f_probe_tet.err
cat gpurun_out/r02f_probe_tet.json
for v in 0 4; do
    timeout 300 python bench.py --spmv-variant $v --steps 1 --warmup 1 --no-cpu --no-e2e --no-upload > gpurun_out/r02f_bench_hex_v$v.json 2> gpurun_out/r02f_bench_hex_v$v.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02f_bench_hex_v$v.json').read().strip().splitlines()[-1]); print('hex v$v', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'])"
done
# the other shapes of BASELINE.json's configs as bench lines
timeout 600 python bench.py --preset S3-tet --steps 1 --warmup 3 --no-same-size > gpurun_out/r02f_bench_S3tet256.json 2> gpurun_out/r02f_bench_S3tet256.err
python -c "
import json; d=json.loads(open('gpurun_out/r02f_bench_S3tet256.json').read().strip().splitlines()[-1]); print('tet', d['value'], d['e2e'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None)"
timeout 900 python bench.py --preset S2-tri --steps 1 --warmup 3 --no-same-size > gpurun_out/r02f_bench_S2tri4096.json 2> gpurun_out/r02f_bench_S2tri4096.err
python -c "
import json; d=json.loads(open('gpurun_out/r02f_bench_S2tri4096.json').read().strip().splitlines()[-1]); print('tri', d['value'], d['e2e'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'] if d['cpu_baseline'] else None)"
# ncu: the SpMV as it runs inside the solve (traffic for bench.py's roofline.traffic), the assembly gather
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_s3_rt" -s 40 -c 2 -o gpurun_out/r02f_prof_spmv_insolve \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-upload > gpurun_out/r02f_prof_spmv.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_gather|k_place_elements|k_dirichlet" -c 12 \
    -o gpurun_out/r02f_prof_assembly python tools/probe_next_rows.py assembly > gpurun_out/r02f_prof_assembly.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
