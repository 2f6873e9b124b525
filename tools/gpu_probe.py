#!/usr/bin/env python3
"""Quick device probe: SpMV GB/s (algorithmic bytes / event time) for a few sizes and presets."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
cases = [("S3-hex", 128), ("S3-hex", 200), ("S3-tet", 200), ("S2-tri", 2048), ("S2-tri", 4096)]
if len(sys.argv) > 1:
    cases = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
for preset, n in cases:
    syn = pkg.Synth(preset, n)
    asm = pkg.Assembly(device=0)
    t0 = time.time()
    syn.to_device(asm)
    gen = time.time() - t0
    st = asm.stats()
    out = {"case": f"{preset}-{n}", "ndof": st.ndof, "nnzb": st.nnzb, "gen_s": round(gen, 2), "algo_MB": st.spmv_algorithmic_bytes / 1e6}
    variants = [0, 100] if st.stride == 3 else [8, 0, 100]
    if os.environ.get('PROBE_VARIANTS'):
        variants = [int(v) for v in os.environ['PROBE_VARIANTS'].split(',')]
    import numpy as np
    asm.upload_x0(np.random.default_rng(0).standard_normal(st.ndof))
    asm.spmv_resident(reps=1, variant=1 if st.stride == 3 else 8)
    yref = asm.download_vector(1)
    for v in variants:
        ms = asm.spmv_resident(reps=20, variant=v)
        err = float(np.abs(asm.download_vector(1) - yref).max() / np.abs(yref).max())
        if err > 1e-13:
            out[f"ERR_v{v}"] = err
        out[f"spmv_ms_v{v}"] = round(ms, 4)
        out[f"spmv_GBs_v{v}"] = round(st.spmv_algorithmic_bytes / ms / 1e6, 1)
    print(json.dumps(out), flush=True)
    asm.close()
