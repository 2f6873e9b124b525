#!/bin/bash
# Pass r02e (ONE GPU): group tests with per-test timeouts, block preconditioners, the reworked assembly, tripoint, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py -v --durations=0 --timeout 120 > gpurun_out/r02e_pytest_group.log 2>&1
echo "== group: $(tail -1 gpurun_out/r02e_pytest_group.log)"
grep -E "PASSED|FAILED|Timeout|ERROR" gpurun_out/r02e_pytest_group.log | cut -c1-150 | head -40
grep -E "^[0-9.]+s (call|setup)" gpurun_out/r02e_pytest_group.log | head -8
timeout 300 python -m pytest tests/test_gpu_precond.py -q --timeout 120 -s > gpurun_out/r02e_pytest_precond.log 2>&1
echo "== precond: $(tail -1 gpurun_out/r02e_pytest_precond.log)"
grep -E "block-Jacobi|FAILED|Error" gpurun_out/r02e_pytest_precond.log | head
timeout 400 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_variants.py tests/test_gpu_recovery.py tests/test_gpu_sequence.py -q --timeout 120 > gpurun_out/r02e_pytest_rows.log 2>&1
echo "== assembly/variants/recovery/sequence: $(tail -1 gpurun_out/r02e_pytest_rows.log)"
grep -E "FAILED|Error" gpurun_out/r02e_pytest_rows.log | head
timeout 500 python -m pytest tests/test_gpu_e2e.py -q -s --timeout 400 > gpurun_out/r02e_pytest_e2e.log 2>&1
echo "== e2e: $(tail -1 gpurun_out/r02e_pytest_e2e.log)"
grep -E "tripoint:|e2e 2d|e2e 3d|FAILED|Error" gpurun_out/r02e_pytest_e2e.log | cut -c1-400 | head
timeout 200 python tools/probe_next_rows.py > gpurun_out/r02e_probe_next_rows.jsonl 2> gpurun_out/r02e_probe_next_rows.err
cat gpurun_out/r02e_probe_next_rows.jsonl
tail -3 gpurun_out/r02e_probe_next_rows.err
timeout 900 python bench.py > gpurun_out/r02e_bench_1gpu.json 2> gpurun_out/r02e_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02e_bench_1gpu.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["roofline"]["frac"], d["e2e"]["value"])
print(json.dumps(d["cpu_baseline"])[:1500])
print(d["upload"])
PY
tail -3 gpurun_out/r02e_bench_1gpu.err
