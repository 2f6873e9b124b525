#!/bin/bash
# Second GPU pass of round 1 (ONE GPU, a few minutes): validates the field-recovery row, takes first timings and an
# ncu capture of the assembly / elimination / field kernels, then re-runs the whole GPU suite.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_recovery.py -q > gpurun_out/r01b_pytest_recovery.log 2>&1
tail -3 gpurun_out/r01b_pytest_recovery.log
timeout 200 python tools/probe_next_rows.py > gpurun_out/r01b_probe_next_rows.jsonl 2> gpurun_out/r01b_probe_next_rows.err
cat gpurun_out/r01b_probe_next_rows.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_assemble_gather|k_dirichlet|k_element_fields" -c 8 \
    -o gpurun_out/r01b_prof_next_rows python tools/probe_next_rows.py > gpurun_out/r01b_prof_next_rows.log 2>&1
ls -la gpurun_out | tail -5
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r01b_pytest_gpu.log 2>&1
tail -3 gpurun_out/r01b_pytest_gpu.log
