#!/bin/bash
# First GPU pass of round 2 (ONE GPU, ~20 min): everything written after round 1's GPU budget was spent gets its first
# run on hardware, then the probes and captures that decide which prepared variants become defaults.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_pass_r02.sh'      (about 20 min of box time)
mkdir -p gpurun_out
# 1. the GPU tests that have not run on a B200 yet, one file at a time so that one failure does not hide the others
for f in precond sequence variants recovery; do
    timeout 400 python -m pytest tests/test_gpu_$f.py -q > gpurun_out/r02_pytest_$f.log 2>&1
    echo "== $f: $(tail -1 gpurun_out/r02_pytest_$f.log)"
done
# 2. the rows next to the solve: defaults vs prepared variants (same bits, times)
timeout 200 python tools/probe_next_rows.py > gpurun_out/r02_probe_next_rows.jsonl 2> gpurun_out/r02_probe_next_rows.err
cat gpurun_out/r02_probe_next_rows.jsonl
timeout 300 python tools/probe_renumber.py 96 > gpurun_out/r02_probe_renumber.json 2> gpurun_out/r02_probe_renumber.err
cat gpurun_out/r02_probe_renumber.json
# 3. one full capture per kernel of those rows (the assembly part of the probe only: -k filters, -c counts MATCHING launches)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_gather|k_dirichlet" -c 20 \
    -o gpurun_out/r02_prof_assembly python tools/probe_next_rows.py assembly > gpurun_out/r02_prof_assembly.log 2>&1
# 4. the whole suite and the default bench line
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
tail -c 600 gpurun_out/r02_bench_1gpu.json
# 5. A/B: p.q as its own pass after a plain SpMV (the fused form costs ~8 % of the SpMV, a pass over p and q ~2 %)
AMIE_B200_SPLIT_DOT=1 timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/r02_bench_1gpu_split_dot.json 2> gpurun_out/r02_bench_1gpu_split_dot.err
tail -c 600 gpurun_out/r02_bench_1gpu_split_dot.json
# 6. the other shapes of BASELINE.json's configs as bench lines (tetrahedra: 15 blocks per row; 2D: 2x2 blocks)
timeout 400 python bench.py --preset S3-tet --mesh-n 256 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_bench_S3tet256.json 2> gpurun_out/r02_bench_S3tet256.err
tail -c 400 gpurun_out/r02_bench_S3tet256.json
timeout 600 python bench.py --preset S2-tri --mesh-n 4096 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_bench_S2tri4096.json 2> gpurun_out/r02_bench_S2tri4096.err
tail -c 400 gpurun_out/r02_bench_S2tri4096.json
