#!/bin/bash
# Second GPU pass of round 2 (ONE GPU): the multi-device context on one box (parts side by side on device 0), the
# team-of-warps SpMV variants against the shipped kernel, compute-sanitizer logs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py -q -x > gpurun_out/r02b_pytest_group.log 2>&1
echo "== group: $(tail -1 gpurun_out/r02b_pytest_group.log)"
tail -30 gpurun_out/r02b_pytest_group.log | head -28
# the drop-in binary with the devices taken from the environment (two parts on device 0)
AMIE_B200_DEVICES=0,0 timeout 600 python -m pytest tests/test_gpu_e2e.py -q > gpurun_out/r02b_pytest_e2e_group.log 2>&1
echo "== e2e with AMIE_B200_DEVICES=0,0: $(tail -1 gpurun_out/r02b_pytest_e2e_group.log)"
# SpMV A/B, isolated launches (plain and in-solve form)
PROBE_VARIANTS=0,4,5,100,104,105 timeout 300 python tools/gpu_probe.py S3-hex:256 > gpurun_out/r02b_probe_hex.json 2> gpurun_out/r02b_probe_hex.err
cat gpurun_out/r02b_probe_hex.json
PROBE_VARIANTS=0,6,7,4,100,106,107,104 timeout 300 python tools/gpu_probe.py S3-tet:256 > gpurun_out/r02b_probe_tet.json 2> gpurun_out/r02b_probe_tet.err
cat gpurun_out/r02b_probe_tet.json
# inside the solve, under the sustained-load power cap
for v in 4 5; do
    timeout 300 python bench.py --spmv-variant $v --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_bench_hex_v$v.json 2> gpurun_out/r02b_bench_hex_v$v.err
    tail -c 700 gpurun_out/r02b_bench_hex_v$v.json
done
for v in 6 7; do
    timeout 300 python bench.py --preset S3-tet --spmv-variant $v --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_bench_tet_v$v.json 2> gpurun_out/r02b_bench_tet_v$v.err
    tail -c 700 gpurun_out/r02b_bench_tet_v$v.json
done
# sanitizer: shared-memory hazards of the pipelines, out-of-bounds accesses, one device and two parts
for tool in racecheck memcheck; do
    SANITIZE_MAXIT=12 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py S3-hex 20 > gpurun_out/r02b_sanitizer_${tool}_S3hex20.log 2>&1
    echo "== $tool S3-hex-20: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02b_sanitizer_${tool}_S3hex20.log | tail -1)"
    SANITIZE_MAXIT=12 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py S2-tri 40 > gpurun_out/r02b_sanitizer_${tool}_S2tri40.log 2>&1
    echo "== $tool S2-tri-40: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02b_sanitizer_${tool}_S2tri40.log | tail -1)"
done
SANITIZE_MAXIT=12 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_case.py S3-hex 20 0,0 > gpurun_out/r02b_sanitizer_memcheck_group.log 2>&1
echo "== memcheck group 0,0: $(grep -E 'ERROR SUMMARY' gpurun_out/r02b_sanitizer_memcheck_group.log | tail -1)"
