#!/usr/bin/env python3
"""One small solve per call, for compute-sanitizer (racecheck / memcheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py S3-hex 20 [devices, e.g. 0,0]
PCG + BiCGStab through the C-ABI with host buffers; prints one JSON line."""
import json
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

pkg = g.load_package()
preset, n = sys.argv[1], int(sys.argv[2])
devices = [int(d) for d in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
syn = pkg.Synth(preset, n)
rs, ci, arr, b = syn.rows()
A = pkg.CoordinateIndexedSparseMatrix(rs, ci, syn.stride, arr)
asm = pkg.Assembly(A, b, devices=devices) if devices else pkg.Assembly(A, b, device=0)
maxit = int(os.environ.get("SANITIZE_MAXIT", "-1"))
cg = pkg.ConjugateGradient(asm)
cg.nssor = 32
ok = cg.solve(None, None, 1e-10, maxit)
bi = pkg.BiConjugateGradientStabilized(asm)
okb = bi.solve(None, None, 1e-10, maxit if maxit > 0 else -1)
print(json.dumps({"case": f"{preset}-{n}", "devices": devices, "pcg": [bool(ok), int(cg.nit)], "bicgstab": [bool(okb), int(bi.nit)]}))
asm.close()
