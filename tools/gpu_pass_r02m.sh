#!/bin/bash
# Pass r02m (ONE GPU), the closing pass of round 2 with the shipped kernels: ncu evidence (launch list of a bench step and
# the full capture of the in-solve SpMV), one bench line, then the whole GPU suite.
mkdir -p gpurun_out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02m_launches_S3hex256.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-upload > gpurun_out/r02m_launches_bench.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_s3_rt" -s 40 -c 2 -o gpurun_out/r02m_prof_spmv_insolve \
    python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-upload > gpurun_out/r02m_prof_spmv.log 2>&1
ls -la gpurun_out/r02m_prof_spmv_insolve.ncu-rep gpurun_out/r02m_launches_S3hex256.csv
timeout 600 python bench.py --steps 1 --warmup 1 > gpurun_out/r02m_bench_1gpu.json 2> gpurun_out/r02m_bench_1gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r02m_bench_1gpu.json').read().strip().splitlines()[-1]); print('bench', d['value'], 'e2e', d['e2e'], 'roofline', d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'], 'cpu', (d['cpu_baseline'] or {}).get('value'))"
for p in S3-tet S2-tri; do W=1; [ $p = S2-tri ] && W=0
    timeout 300 python bench.py --preset $p --steps 1 --warmup $W --no-cpu --no-e2e --no-upload > gpurun_out/r02m_bench_$p.json 2> gpurun_out/r02m_bench_$p.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02m_bench_$p.json').read().strip().splitlines()[-1]); print('$p', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['clocks'], d['config']['iterations_per_step'])"
done
timeout 1100 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/r02m_pytest_gpu.log 2>&1
echo "== pytest -m gpu: $(tail -1 gpurun_out/r02m_pytest_gpu.log)"
grep -E "^E  |FAILED|Timeout" gpurun_out/r02m_pytest_gpu.log | cut -c1-300 | head -20
