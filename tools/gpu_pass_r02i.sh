#!/bin/bash
# Pass r02i (ONE GPU, short): pinned DOF set of the 3D end-to-end case, assembly gather with cached loads + its ncu capture.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_e2e.py -q -s --timeout 150 -k featuretree > gpurun_out/r02i_pytest_e2e.log 2>&1
echo "== e2e: $(tail -1 gpurun_out/r02i_pytest_e2e.log)"
grep -E "e2e 2d|e2e 3d|^E  " gpurun_out/r02i_pytest_e2e.log | cut -c1-400
timeout 200 python tools/probe_next_rows.py assembly > gpurun_out/r02i_probe_assembly.jsonl 2> gpurun_out/r02i_probe_assembly.err
cat gpurun_out/r02i_probe_assembly.jsonl | cut -c1-900
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_gather" -c 3 \
    -o gpurun_out/r02i_prof_gather python tools/probe_next_rows.py assembly > gpurun_out/r02i_prof_gather.log 2>&1
ls -la gpurun_out/r02i_prof_gather.ncu-rep
timeout 300 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_precond.py -q --timeout 120 > gpurun_out/r02i_pytest_rows.log 2>&1
echo "== assembly + precond: $(tail -1 gpurun_out/r02i_pytest_rows.log)"
