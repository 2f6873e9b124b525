#!/usr/bin/env python3
"""Gather-kernel variants of amie_b200_assemble on a 64^3-node hexahedral grid (250 k elements, 1.74 GB algorithmic):
time, and bit-equality of the assembled values between variants (variant 0 is the one the parity tests pin).
PROBE_NCU=1: one full launch per variant + the elimination, for an ncu capture."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as g  # noqa: E402
from probe_next_rows import hex_grid, pattern, PEAK  # noqa: E402

pkg = g.load_package()
ncu = bool(os.environ.get("PROBE_NCU"))
n, s = 64, 3
ids = hex_grid(n)
nb = n ** 3
rs, ci = pattern(ids, nb)
rng = np.random.default_rng(0)
ke = rng.standard_normal((ids.shape[0], 8, 8, 9)) * 10.0 ** rng.integers(-3, 4, (ids.shape[0], 1, 1, 1))
scales = rng.uniform(0.5, 2.0, ids.shape[0])
nnzb = int(ci.size)
alg = ke.size * 8 + ke.size // 9 * 4 + nnzb * (8 * s * s + 4)
asm = pkg.Assembly(device=0)
asm.set_structure_only(s, rs, ci)
asm.set_elements(ids)
out = dict(elements=int(ids.shape[0]), nnzb=nnzb, algorithmic_bytes=alg)
arrays = {}
cnt = ids.shape[0] // 100
for variant in (0, 1):
    asm.set_option("assemble_variant", variant)
    full, part = [], []
    for _ in range(1 if ncu else 4):
        asm.update_elements(0, ke, scales)
        asm.assemble()
        full.append(asm.stats().assemble_ms)
    if not ncu:
        arrays[variant] = asm.download_matrix()[2]
        for r in range(3):                      # a damage step touching 1 % of the elements
            asm.update_elements(r * cnt, ke[r * cnt:(r + 1) * cnt] * 0.5, scales[r * cnt:(r + 1) * cnt])
            asm.assemble()
            part.append(asm.stats().assemble_ms)
        arrays[(variant, "part")] = asm.download_matrix()[2]
    out[f"variant{variant}"] = dict(full_ms=min(full), gbs=alg / (min(full) * 1e-3) / 1e9,
                                    frac_of_peak=alg / (min(full) * 1e-3) / 1e9 / PEAK,
                                    one_percent_ms=min(part) if part else None)
if ncu:
    asm.upload_rhs(np.zeros(nb * s))
    fix = np.arange(0, nb * s, 97, dtype=np.uint32)
    asm.set_boundary_conditions(fix, np.ones(fix.size))
else:
    out["same_bits_full"] = bool(np.array_equal(arrays[0].view(np.uint64), arrays[1].view(np.uint64)))
    out["same_bits_after_partial_updates"] = bool(np.array_equal(arrays[(0, "part")].view(np.uint64), arrays[(1, "part")].view(np.uint64)))
    out["nonzero_values"] = int(np.count_nonzero(arrays[0]))
print(json.dumps(out), flush=True)
asm.close()
