#!/bin/bash
# Pass r02k (ONE GPU): the tuned x gather (parity + isolated launches), the 2x2 pipeline with the column mapping (variant 5)
# against the row mapping, the new end-to-end cases (3d-1000, XFEM-enriched asr2d), ncu of the 2x2 kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -q --timeout 120 \
    -k "column_mapping or spmv_matches or ragged or enrichment or test_pcg_parity or other_strides" > gpurun_out/r02k_pytest_parity.log 2>&1
echo "== parity subset: $(tail -1 gpurun_out/r02k_pytest_parity.log)"
grep -E "^E  |FAILED" gpurun_out/r02k_pytest_parity.log | head -10
timeout 400 python -m pytest tests/test_gpu_e2e.py -q -s --timeout 300 -k featuretree > gpurun_out/r02k_pytest_e2e.log 2>&1
echo "== e2e: $(tail -1 gpurun_out/r02k_pytest_e2e.log)"
grep -E "e2e 2d|e2e 3d|e2e asr|^E  " gpurun_out/r02k_pytest_e2e.log | cut -c1-400
PROBE_VARIANTS=0,100 timeout 200 python tools/gpu_probe.py S3-hex:256 S3-tet:256 > gpurun_out/r02k_probe_s3.json 2> gpurun_out/r02k_probe_s3.err
cat gpurun_out/r02k_probe_s3.json
PROBE_VARIANTS=0,5,100,105,0,5,100,105 timeout 200 python tools/gpu_probe.py S2-tri:4096 > gpurun_out/r02k_probe_tri.json 2> gpurun_out/r02k_probe_tri.err
cat gpurun_out/r02k_probe_tri.json
for v in 100 105; do
    PROBE_VARIANTS=$v timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_s2_rt" -s 3 -c 2 -o gpurun_out/r02k_prof_spmv_s2_v$v \
        python tools/gpu_probe.py S2-tri:4096 > gpurun_out/r02k_prof_spmv_s2_v$v.log 2>&1
done
ls -la gpurun_out/r02k_prof_spmv_s2_v*.ncu-rep
