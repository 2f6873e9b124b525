import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
n = 64
syn = pkg.Synth("S3-hex", n)
for mode in ("plain-ctx", "dist-ctx"):
    asm = pkg.Assembly(device=0)
    if mode == "dist-ctx":
        asm.dist_init(0, 1, pkg.nccl_unique_id(), np.array([0, syn.nb], np.uint64))
        asm.dist_synth_to_device(syn)
    else:
        syn.to_device(asm)
    for variant in (1, 0, 1, 0, 2):
        asm.set_option("spmv_variant", variant)
        asm.upload_x0(None)
        r = asm.pcg_resident(nssor=32)
        x = asm.download_x()
        print(mode, "variant", variant, r, "x checksum", float(np.abs(x).sum()), flush=True)
    asm.close()
