#!/usr/bin/env python3
"""Latency of small solves (the sizes of the reference's own drivers: main_tripoint 9 k DOF,
main_3d_benchmark 26 k DOF): microseconds per PCG iteration with and without CUDA-graph batches."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
for preset, n in [("S2-tri", 67), ("S3-tet", 21), ("S3-hex", 32), ("S3-hex", 64)]:
    syn = pkg.Synth(preset, n)
    asm = syn.assembly(device=0)
    asm.sync_matrix()
    asm.upload_rhs(asm.getForces())
    out = {"case": f"{preset}-{n}", "ndof": syn.nb * syn.stride}
    for graph in (0, 1):
        asm.set_option("graph", graph)
        for rep in range(3):
            asm.upload_x0(None)
            t0 = time.time()
            ok, nit, err, rho = asm.pcg_resident(nssor=32)
            wall = time.time() - t0
        st = asm.stats()
        out[f"graph{graph}"] = {"nit": nit, "solve_ms": round(st.solve_ms, 3), "us_per_it": round(1e3 * st.solve_ms / max(1, nit), 2), "wall_ms": round(1e3 * wall, 3)}
    print(json.dumps(out), flush=True)
    asm.close()
